#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (markdown)."""
import collections
import csv
import sys


def main(path, title=""):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        n += 1
        a = agg.setdefault(row["Kernel Name"].split("(")[0][:70], [0, 0.0])
        a[0] += 1
        a[1] += float(row["Metric Value"].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print(f"# {title or path}\n\n{n} launches captured; per-launch times are cold-cache and serialised under ncu — compare SHARES.\n")
    print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {a[0]} | {a[1] / 1e3:.1f} | {a[1] / a[0] / 1e3:.1f} | {a[1] / tot:.3f} |")


if __name__ == "__main__":
    main(*sys.argv[1:])
