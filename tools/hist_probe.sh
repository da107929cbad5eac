#!/bin/bash
# histogram kernel knobs: run length per lane and CTAs per level
for cfg in "64 12" "64 16" "48 12" "80 12" "64 20"; do
  set -- $cfg
  FALWA_HIST_RUN=$1 FALWA_HIST_CTAS=$2 python bench.py --no-cpu-baseline --steps 3 --batch 8 > gpurun_out/h_$1_$2.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/h_$1_$2.json'))
print("run=$1 ctas=$2 value=%.1f area_hist3 ms/snap=%.3f" % (d['value'], d['kernels']['area_hist3_kernel']['ms_per_step']/8))
PY
done
