#!/bin/bash
run() { echo "== $1"; env $2 TENSORF_B200_LIB=$3 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-render 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stages_ms']; print(round(d['ms_per_step'],4), {k:s[k] for k in ('density_select','density_scatter','appearance_scatter','appearance_gather')})"; }
for m in 2 3 4; do run sw$m "X=1" tools/lib_sw$m.so; done
run sw3seg8 "TENSORF_SEG_LEN=8" tools/lib_sw3.so
run sw3seg32 "TENSORF_SEG_LEN=32" tools/lib_sw3.so
run sw3app32 "TENSORF_SEG_LEN_APP=32" tools/lib_sw3.so
run sw2seg8 "TENSORF_SEG_LEN=8" tools/lib_sw2.so
