#!/bin/bash
# Multi-GPU pass: peer tests + worker timings, bench with each exchange.  Usage: bash tools/run_2gpu.sh <tag> [ngpus]
tag=${1:-rXX}; n=${2:-2}
out=gpurun_out
timeout 300 python -m pytest tests/test_gpu_peer.py -x -q -m gpu 2>&1 | tail -5
tr() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
PEER_WORKER_OUT=$out/${tag}_peer_${n}gpu.json tr 29750 tests/peer_worker.py > $out/${tag}_peer_${n}gpu.log 2>&1; tail -2 $out/${tag}_peer_${n}gpu.log
tr 29751 bench.py --gpus $n --steps 20 --warmup 5 > $out/${tag}_bench_${n}gpu.json 2> $out/${tag}_bench_${n}gpu.err
tail -3 $out/${tag}_bench_${n}gpu.err
for ex in peer-p2p peer-overlap peer-multicast nccl; do
  tr 29752 bench.py --gpus $n --steps 20 --warmup 5 --exchange $ex --no-render > $out/${tag}_bench_${n}gpu_$ex.json 2> $out/${tag}_bench_${n}gpu_$ex.err
done
python - <<PY
import json
for f in ("", "_peer-p2p", "_peer-overlap", "_peer-multicast", "_nccl"):
    f = "$out/${tag}_bench_${n}gpu" + f + ".json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["config"]["parallelism"])
    except Exception as e:
        print(f, "FAILED", e)
PY
