for b in 2 8 12; do python bench.py --no-cpu-baseline --steps 4 --batch $b > gpurun_out/bp.json 2>gpurun_out/bp.err || tail -2 gpurun_out/bp.err; python -c "
import json; d=json.load(open('gpurun_out/bp.json')); print('batch', $b, 'value', round(d['value'],1), 'ms/snapshot', round(d['ms_per_step']/$b,3), 'e2e', round(d['e2e']['value'],1))"; done
