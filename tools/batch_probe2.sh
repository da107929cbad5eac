for b in 8 9 16 18; do
  echo "== batch $b"
  python bench.py --kernels-only --steps 3 --warmup 2 --batch $b 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
b=$b
print('   %.1f snapshots/s  '%d['value'] + '  '.join('%s %.4f'%(k.replace('_kernel',''),v/b) for k,v in list(d['kernels'].items())[:6]))"
done
