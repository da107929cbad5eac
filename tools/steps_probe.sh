for st in 3 5 8; do python bench.py --no-cpu-baseline --steps $st > gpurun_out/sp.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/sp.json')); print('steps', $st, 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'clocks', d['clocks'])"; done
