#!/bin/bash
# Short N-GPU validation (charged N x): peer worker (parity + exchange timings) and one bench line.  Usage: <tag> <ngpus>
tag=${1:-rXX}; n=${2:-4}
out=gpurun_out
tr() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
PEER_WORKER_OUT=$out/${tag}_peer_${n}gpu.json tr 29750 tests/peer_worker.py > $out/${tag}_peer_${n}gpu.log 2>&1; tail -2 $out/${tag}_peer_${n}gpu.log
tr 29751 bench.py --gpus $n --steps 20 --warmup 5 --no-render > $out/${tag}_bench_${n}gpu.json 2> $out/${tag}_bench_${n}gpu.err
tail -2 $out/${tag}_bench_${n}gpu.err; cat $out/${tag}_bench_${n}gpu.json | cut -c1-600
