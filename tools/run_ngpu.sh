#!/bin/bash
# usage (GPU box, repo root): bash tools/run_ngpu.sh <tag> <N> [extra bench args]   -> gpurun_out/<tag>_bench_<N>gpu*.json
set -u
TAG=$1; N=$2; shift 2
OUT=gpurun_out; mkdir -p $OUT
run() {  # name, extra args...
  local name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline "$@" > $OUT/${TAG}_bench_${N}gpu_${name}.json 2> $OUT/${TAG}_bench_${N}gpu_${name}.err
  echo "== $name rc=$?"; python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench_${N}gpu_${name}.json"))
    print("value %.3f M rays/s  ms %.4f  e2e %.3f M" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6))
    print("parallelism:", d["config"]["parallelism"])
    print("stages:", d["stages_ms"])
    if "config3" in d: print("config3:", {k: d["config3"][k] for k in ("value", "ms_per_step", "R_per_gpu", "exchange")}, d["config3"]["stages_ms"])
    if "packed_layout" in d: print("packed:", d["packed_layout"]["value"], d["packed_layout"]["ms_per_step"])
    if "parity_check" in d: print("parity:", d["parity_check"])
    if "render" in d: print("render:", {k: (v["mpix_per_s"], v["ms_per_frame"]) for k, v in d["render"].items() if isinstance(v, dict)})
except Exception as e:
    print("no json:", e)
    print(open("$OUT/${TAG}_bench_${N}gpu_${name}.err").read()[-3000:])
PY
}
run auto "$@"
run p2p --exchange peer-p2p --no-render --no-config3 --no-packed "$@"
run nccl --exchange nccl --no-render --no-config3 --no-packed "$@"
