B="python bench.py --steps 40 --warmup 3 --no-e2e --no-cpu"
for extra in "--regconst 1 --maxreg 102" "--regconst 1 --maxreg 85" "--regconst 0 --maxreg 85" "--regconst 0 --maxreg 72"; do
  echo "== $extra"
  $B $extra 2>&1 | tail -1 | python -c "
import sys,json,numpy as np
d=json.loads(sys.stdin.read())
r=d['roofline']; k=d['kernel_ms_each']; print('value %.3e frac %.3f regs %d blocks/SM %d spill %d first5 %.3f last10 %.3f min %.3f' % (d['value'], r['frac'], r['registers'], r['blocks_per_sm'], r['spill_bytes'], np.mean(k[1:6]), np.mean(k[-10:]), min(k)))"
done
