#!/bin/bash
# window kernel: warps-per-CTA sweep (bench lines only)
for nw in "$@"; do
  FALWA_WIN_WARPS=$nw python bench.py --no-cpu-baseline --steps 5 > gpurun_out/nw${nw}.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/nw${nw}.json").read().strip().splitlines()[-1])
print("NW=${nw} value %.1f  window %.3f ms/step"%(d["value"], d["kernels"]["lwa_window_kernel"]["ms_per_step"]))
PY
done
