#!/bin/bash
run() { echo "== $1"; env $2 TENSORF_B200_LIB=$3 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-render 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stages_ms']; print(round(d['ms_per_step'],4), {k:s[k] for k in ('density_select','density_scatter','appearance_scatter','appearance_gather')})"; }
for m in 4 6 12; do run ds_minb$m "X=1" tools/lib_ds$m.so; done
run seg8 "TENSORF_SEG_LEN=8" tensorf-jax_b200/tensorf_b200/libtensorf_b200.so
run seg32 "TENSORF_SEG_LEN=32" tensorf-jax_b200/tensorf_b200/libtensorf_b200.so
run app32 "TENSORF_SEG_LEN_APP=32" tensorf-jax_b200/tensorf_b200/libtensorf_b200.so
run app128 "TENSORF_SEG_LEN_APP=128" tensorf-jax_b200/tensorf_b200/libtensorf_b200.so
