#!/bin/bash
# first GPU contact of lwa_prefix.cu: Section A parity (1.5 degree, real + analytic, NaN), then full size, then timings
timeout 300 python -m pytest tests/test_gpu_f90_modules.py -q -x -k "prefix" 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_fullsize_oracle.py -q -x -k "prefix or upsampled" 2>&1 | tail -15
for a in "--data analytic" "--data upsampled" "FORCE1 --data analytic" "FORCE0 --data upsampled"; do
  echo "== $a"
  e=""; case "$a" in FORCE1*) e="FALWA_LWA_FORCE=1"; a="${a#FORCE1 }";; FORCE0*) e="FALWA_LWA_FORCE=0"; a="${a#FORCE0 }";; esac
  env $e timeout 200 python bench.py --kernels-only --steps 3 --warmup 2 $a 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('   %.1f snapshots/s  '%d['value'] + '  '.join('%s %.3f'%(k.replace('_kernel',''),v) for k,v in d['kernels'].items()))"
done
