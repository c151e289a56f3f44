#!/bin/bash
# One-GPU profiling pass for profiles/: bench JSON (not under a profiler), ncu launch list, ncu --set full of
# one step's kernels.  Usage (on the GPU box): bash tools/profile_round.sh <tag>
tag=${1:-rXX}
out=gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err
tail -2 $out/${tag}_bench_1gpu.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 90 -c 120 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-render > $out/${tag}_ncu_list.log 2>&1
# one full step: skip warm-up launches (3 steps x ~33 launches incl. memsets), capture ~36 kernels
timeout 900 ncu --set full --clock-control none --import-source on -s 70 -c 30 -o $out/${tag}_full -f \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-render --no-l2-flush > $out/${tag}_ncu_full.log 2>&1
ncu -i $out/${tag}_full.ncu-rep --page raw --csv > $out/${tag}_full_raw.csv 2>/dev/null
ls -la $out/${tag}_full.ncu-rep | awk '{print $5}'
bash tools/other_cfgs.sh > $out/${tag}_other_cfgs.log 2>&1
tail -12 $out/${tag}_other_cfgs.log
