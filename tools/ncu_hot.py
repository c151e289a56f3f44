#!/usr/bin/env python
"""Hot SASS lines + instruction mix of one kernel from `ncu -i X.ncu-rep --page source --csv` (stdin or file).
usage: ncu_hot.py src.csv <kernel substring> [top]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
want, top_n = sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name, hdr = rows[i][1], rows[i + 1]
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            body.append(rows[j]); j += 1
        i = j
        if want not in name or "Source" not in hdr:
            continue
        si, src, ex = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source"), hdr.index("Instructions Executed")
        tot, totex = sum(int(r[si] or 0) for r in body), sum(int(r[ex] or 0) for r in body)
        print(name[:80], "| samples", tot, "| SASS lines", len(body), "| warp-instr", totex)
        c = collections.Counter()
        for r in body:
            t = r[src].strip().split()
            if t:
                c[(t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]] += int(r[ex] or 0)
        print("mix:", [(k, round(v / max(totex, 1), 3)) for k, v in c.most_common(14)])
        for k, r in sorted(sorted(enumerate(body), key=lambda x: -int(x[1][si] or 0))[:top_n], key=lambda x: x[0]):
            print(f"{k:5d} {r[si]:>6s} {r[ex]:>9s}  {r[src].strip()[:100]}")
        break
    else:
        i += 1
