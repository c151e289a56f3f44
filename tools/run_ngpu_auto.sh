#!/bin/bash
# usage: bash tools/run_ngpu_auto.sh <tag> <N>   -> the default (--exchange auto) bench line at N GPUs
TAG=$1; N=$2
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 20 --warmup 5 > $OUT/${TAG}_bench_${N}gpu_auto.json 2> $OUT/${TAG}_bench_${N}gpu_auto.err
echo "rc=$?"
python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench_${N}gpu_auto.json"))
    print("value %.3f M rays/s  ms %.4f  e2e %.3f M" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6))
    print("parallelism:", d["config"]["parallelism"]); print("tuning:", d["config"]["exchange_tuning"])
    print("config3:", d["config3"]["value"], d["config3"]["ms_per_step"], d["config3"]["exchange_tuning"])
    print("packed:", d["packed_layout"]["value"]); print("parity:", {k: v for k, v in d["parity_check"].items() if k.startswith(("max", "adam"))})
    print("render:", {k: (v["mpix_per_s"], v["ms_per_frame"]) for k, v in d["render"].items() if isinstance(v, dict)})
except Exception as e:
    print("no json:", e); print(open("$OUT/${TAG}_bench_${N}gpu_auto.err").read()[-3000:])
PY
