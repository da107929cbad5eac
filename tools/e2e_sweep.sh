#!/bin/bash
for c in 1 2 4; do
  python bench.py --no-cpu-baseline --steps 3 --chunk $c > gpurun_out/e2e_$c.json 2>gpurun_out/e2e_$c.err || tail -3 gpurun_out/e2e_$c.err
  python -c "
import json; d=json.load(open('gpurun_out/e2e_$c.json')); print('chunk $c', 'value', round(d['value'],1), 'e2e', d['e2e'])"
done
