#!/bin/bash
# usage (on the GPU box, from the repo root): bash tools/profile_round.sh <tag> [workload]
# Writes under gpurun_out/<tag>_*: the bench JSON line (never under a profiler), the ncu launch list of the
# same command (repo kernels only, `-k regex:^k_`), and one `--set full` capture of every kernel of the last step.
set -u
TAG=${1:-r02}
WL=${2:-lego_256}
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 20 --warmup 5 --workload $WL > $OUT/${TAG}_bench_1gpu.json 2> $OUT/${TAG}_bench.err
tail -c 400 $OUT/${TAG}_bench_1gpu.json
NL=$(python -c "import json,sys; d=json.load(open('$OUT/${TAG}_bench_1gpu.json')); print(d['gpu_launches']//d['steps'])")
echo "launches per step: $NL"
CMD="python bench.py --steps 1 --warmup 3 --no-render --no-cpu-baseline --no-config3 --no-packed --workload $WL"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
    $CMD > $OUT/${TAG}_ncu_bench.log 2>&1
SKIP=$(python tools/launch_table.py $OUT/${TAG}_launches.csv $NL skip)
python tools/launch_table.py $OUT/${TAG}_launches.csv $NL > $OUT/${TAG}_launches.md
ncu --set full --clock-control none --import-source on -k regex:^k_ -s $SKIP -c $NL -o $OUT/${TAG}_full -f \
    $CMD > $OUT/${TAG}_ncu_full.log 2>&1
ncu -i $OUT/${TAG}_full.ncu-rep --page raw --csv > $OUT/${TAG}_full_raw.csv 2>/dev/null
python tools/ncu_summary.py < $OUT/${TAG}_full_raw.csv > $OUT/${TAG}_ncu_full.md
cat $OUT/${TAG}_launches.md | head -30
ls -la $OUT | tail -12
