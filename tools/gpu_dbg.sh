#!/bin/bash
for d in 0 1 2; do
  FALWA_WIN_DEBUG=$d FALWA_WIN_WARPS=24 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/dbg$d.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/dbg$d.json").read().strip().splitlines()[-1])
print("DEBUG=$d window %.3f ms/step"%(d["kernels"]["lwa_window_kernel"]["ms_per_step"]))
PY
done
