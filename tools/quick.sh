#!/bin/bash
# quick GPU iteration: flux/pipeline parity tests + bench kernel table.  usage: tools/quick.sh [pytest -k expr] [bench args...]
K=${1:-"flux or lwa or pipeline"}; shift
python -m pytest tests -m gpu -x -q -k "$K" 2>&1 | tail -4
python bench.py --no-cpu-baseline --steps 3 "$@" > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err || tail -20 gpurun_out/quick_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/quick_bench.json'))
print(f"value={d['value']:.1f} snap/s  ms/step={d['ms_per_step']:.2f}  e2e={d['e2e']['value']:.1f}  B={d['config']['snapshots_per_step_per_gpu']} launches={d['gpu_launches']}")
B=d['config']['snapshots_per_step_per_gpu']
for k,v in d['kernels'].items():
    print(f"  {k:34s} n={v['launches_per_step']:5.1f} ms/snap={v['ms_per_step']/B:7.3f} share={v['share']*100:5.1f}%")
print('roofline', {k:(round(v,4) if isinstance(v,float) else v) for k,v in d['roofline'].items()})
PY
