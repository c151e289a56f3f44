#!/bin/bash
for w in lego_221 lego_300 lego_300_4096 dozer_128 dozer_300; do
  echo "== $w"; timeout 280 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-render --workload $w 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), 'rays/s', round(d['ms_per_step'],3),'ms', 'step-frac', round(d['roofline_step']['frac'],3), d['stages_ms'])"
done
