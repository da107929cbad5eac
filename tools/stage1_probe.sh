#!/bin/bash
# stage-1 tile shape sweep: time of interp_qgpv_kernel per snapshot for different CTA shapes (bx x 256/bx)
for bx in 32 64 128 256; do
  FALWA_STAGE1_BX=$bx python bench.py --no-cpu-baseline --steps 3 --batch 4 > gpurun_out/s1_$bx.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/s1_$bx.json'))
print("bx=$bx value=%.1f interp_qgpv ms/snap=%.3f" % (d['value'], d['kernels']['interp_qgpv_kernel']['ms_per_step']/4))
PY
done
