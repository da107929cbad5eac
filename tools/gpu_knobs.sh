#!/bin/bash
# usage: tools/gpu_knobs.sh "ENV1=a ENV2=b" "ENV1=c" ...   one short bench run per environment setting
for e in "$@"; do
  echo "== $e"
  env $e python bench.py --kernels-only --steps 3 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('   %.1f snapshots/s  '%d['value'] + '  '.join('%s %.3f'%(k.replace('_kernel',''),v) for k,v in d['kernels'].items()))"
done
