#!/bin/bash
run() { echo "== $1"; env $2 timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-render --workload dozer_128 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stages_ms']; print(round(d['ms_per_step'],4), {k:s[k] for k in ('density_select','density_scatter','appearance_scatter','appearance_gather')})"; }
run base "X=1"
run seg8 "TENSORF_SEG_LEN=8"
run seg32 "TENSORF_SEG_LEN=32"
run seg64 "TENSORF_SEG_LEN=64"
run app32 "TENSORF_SEG_LEN_APP=32"
run app128 "TENSORF_SEG_LEN_APP=128"
