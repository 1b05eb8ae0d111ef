#!/bin/bash
# wave grid kernel: parity tests, then 2048^2 x 60 d bench for the grid paths
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "wave or fused_grid or grid_model" 2>&1 | tail -5
B="timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --workload exphydro_grid_2048x2048_x_60d"
show() { tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    r=d['roofline']; print('value %.3e frac %.3f kernel_ms %.3f regs %d blocks/SM %d spill %d tc %s' % (d['value'], r['frac'], r['kernel_ms'], r['registers'], r['blocks_per_sm'], r['spill_bytes'], d['config'].get('steps_per_launch')))
except Exception as e: print('FAILED', e)"; }
for extra in "${@}"; do
  echo "== $extra"
  $B $extra 2>&1 | show
done
