#!/bin/bash
# 2-GPU check of the in-kernel-sync exchange: worker (parity + timings), bench with kernel / barrier ordering.
tag=${1:-rXX}; n=${2:-2}
out=gpurun_out
tr() { timeout 70 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
PEER_WORKER_OUT=$out/${tag}_peer_${n}gpu.json tr 29750 tests/peer_worker.py > $out/${tag}_peer_${n}gpu.log 2>&1; tail -2 $out/${tag}_peer_${n}gpu.log | cut -c1-1500
tr 29751 bench.py --gpus $n --steps 20 --warmup 5 --no-render > $out/${tag}_bench_${n}gpu_kernelsync.json 2> $out/${tag}_bench_${n}gpu_kernelsync.err
TENSORF_PEER_SYNC=barrier tr 29752 bench.py --gpus $n --steps 20 --warmup 5 --no-render > $out/${tag}_bench_${n}gpu_barriersync.json 2> $out/${tag}_bench_${n}gpu_barriersync.err
python - <<PY
import json
for f in ("kernelsync", "barriersync"):
    f = "$out/${tag}_bench_${n}gpu_" + f + ".json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["config"]["parallelism"])
    except Exception as e:
        print(f, "FAILED", e)
PY
