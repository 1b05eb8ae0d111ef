#!/bin/bash
B="timeout 200 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --workload m50_50k_basins_x_3650d"
for extra in "" "--block 32" "--block 64" "--block 96" "--block 32 --stages 2" "--block 64 --maxreg 168"; do
  echo "== $extra"
  $B $extra 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('value %.3e kernel_ms %.3f regs %d blocks/SM %d spill %d block %d' % (d['value'], r['kernel_ms'], r['registers'], r['blocks_per_sm'], r['spill_bytes'], r['block']))
except Exception as e: print('FAILED', e)"
done
