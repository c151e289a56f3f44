#!/usr/bin/env python
"""Per-kernel table of the LAST step of an ncu launch list (`--metrics gpu__time_duration.sum --csv`).
usage: launch_table.py launches.csv <launches per step>      -> markdown table on stdout
       launch_table.py launches.csv <launches per step> skip -> prints how many launches precede the last step"""
import collections, csv, sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
for r in rd:
    if len(r) > vi and r[mi] == "gpu__time_duration.sum":
        rows.append((r[ki], float(r[vi].replace(",", ""))))
nl = int(sys.argv[2])
if len(sys.argv) > 3 and sys.argv[3] == "skip":
    print(max(0, len(rows) - nl))
    sys.exit(0)
step = rows[-nl:]
tot = sum(t for _, t in step)
agg = collections.OrderedDict()
for k, t in step:
    k = k.split("(")[0]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += t
print(f"{len(rows)} launches captured; last step = {nl} launches, {tot / 1e3:.1f} us under ncu (cold-cache, serialised: compare shares)\n")
print("| kernel | launches | us | share |\n|---|---:|---:|---:|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k[:70]}` | {n} | {t / 1e3:.1f} | {t / tot:.3f} |")
