#!/bin/bash
# per-kernel times of the pipeline on the analytic waves and on the upsampled real snapshot, windowed sums vs interval scan
for a in "--data analytic --method 0" "--data upsampled --method 0" "--data upsampled --method 2" "--data analytic --method 2"; do
  echo "== $a"
  python bench.py --kernels-only --steps 3 --warmup 2 $a 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('   %.1f snapshots/s  '%d['value'] + '  '.join('%s %.3f'%(k.replace('_kernel',''),v) for k,v in d['kernels'].items()))"
done
free -g | head -2
