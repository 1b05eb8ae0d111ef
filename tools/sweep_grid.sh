B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --workload exphydro_grid_2048x2048_x_60d"
for extra in "--generic-route" "" "--grid-tc 1" "--grid-tc 3" "--grid-tc 2 --grid-bh 32" "--grid-tc 3 --grid-bh 32" "--grid-tc 4 --grid-bh 32" "--grid-tc 2 --grid-bw 64 --grid-bh 8" "--grid-tc 2 --grid-bh 8"; do
  echo "== $extra"
  $B $extra 2>&1 | tail -1 | python -c "
import sys,json,numpy as np
try:
    d=json.loads(sys.stdin.read())
    r=d['roofline']; print('value %.3e frac %.3f kernel_ms %.3f regs %d blocks/SM %d spill %d' % (d['value'], r['frac'], r['kernel_ms'], r['registers'], r['blocks_per_sm'], r['spill_bytes']))
except Exception as e: print('FAILED', e)"
done
