#!/usr/bin/env python
"""Markdown summary of an `ncu --set full` report: ncu -i X.ncu-rep --page raw --csv | python tools/ncu_summary.py"""
import csv
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("lts__t_bytes.sum", "L2 bytes"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"), ("launch__registers_per_thread", "regs"),
        ("sm__inst_executed.avg.per_cycle_elapsed", "IPC")]
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
print("| kernel | " + " | ".join(k for _, k in KEYS) + " |\n|---|" + "---:|" * len(KEYS))
for r in rows[2:]:
    cells = []
    for name, _ in KEYS:
        try:
            i = hdr.index(name)
            cells.append(f"{float(r[i]):.1f} {units[i]}".strip())
        except (ValueError, IndexError):
            cells.append("?")
    print(f"| `{r[hdr.index('Kernel Name')][:40]}` | " + " | ".join(cells) + " |")
