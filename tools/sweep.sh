#!/bin/bash
# usage: tools/sweep.sh KERNEL_SUBSTR "ENV1=a ENV2=b" "ENV1=c" ...   -> ms/snapshot of that kernel per setting
K=$1; shift
for cfg in "$@"; do
  env $cfg python bench.py --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/sweep.json 2>gpurun_out/sweep.err || tail -3 gpurun_out/sweep.err
  python - "$K" "$cfg" <<'PY'
import json,sys
d=json.load(open('gpurun_out/sweep.json')); B=d['config']['snapshots_per_step_per_gpu']
for k,v in d['kernels'].items():
    if sys.argv[1] in k: print(f"{sys.argv[2]:40s} {k:28s} {v['ms_per_step']/B:.3f} ms/snap   total {d['value']:.1f} snap/s")
PY
done
