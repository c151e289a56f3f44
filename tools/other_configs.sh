#!/bin/bash
# usage (GPU box, repo root): bash tools/other_configs.sh <tag>  -> gpurun_out/<tag>_other_configs.txt (one line per workload)
TAG=${1:-r02}
OUT=gpurun_out/${TAG}_other_configs.txt
: > $OUT
for wl in lego_221 lego_300 lego_300_4096 dozer_128 dozer_300; do
  python bench.py --workload $wl --steps 10 --warmup 3 --no-render --no-cpu-baseline --no-config3 --no-packed > /tmp/oc.json 2> /tmp/oc.err || { echo "$wl FAILED" >> $OUT; tail -3 /tmp/oc.err >> $OUT; continue; }
  python - "$wl" >> $OUT <<'PY'
import json, sys
d = json.load(open("/tmp/oc.json"))
print(sys.argv[1], d["config"]["workload"], "rays/s %.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"],
      "roofline_step %.3f" % d["roofline_step"]["frac"], "stages", d["stages_ms"])
PY
done
cat $OUT
