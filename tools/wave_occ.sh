#!/bin/bash
# committed static wave kernel: occupancy experiments (record tiles, register budget, forcing stages)
B="timeout 120 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --workload exphydro_grid_2048x2048_x_60d --grid-kernel wave"
show() { tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    r=d['roofline']; print('value %.3e frac %.3f kernel_ms %.3f regs %d blocks/SM %d spill %d smem %d tc %s' % (d['value'], r['frac'], r['kernel_ms'], r['registers'], r['blocks_per_sm'], r['spill_bytes'], r['dyn_smem'], d['config'].get('steps_per_launch')))
except Exception as e: print('FAILED', e)"; }
echo "== base";  $B 2>&1 | show
echo "== obuf2"; HB_WAVE_OBUF=2 $B 2>&1 | show
echo "== obuf2 maxreg 100"; HB_WAVE_OBUF=2 $B --maxreg 100 2>&1 | show
echo "== obuf2 maxreg 100 stages 2"; HB_WAVE_OBUF=2 $B --maxreg 100 --stages 2 2>&1 | show
echo "== obuf2 maxreg 84 stages 2"; HB_WAVE_OBUF=2 $B --maxreg 84 --stages 2 2>&1 | show
echo "== obuf2 maxreg 100 tc 8"; HB_WAVE_OBUF=2 $B --maxreg 100 --grid-tc 8 2>&1 | show
echo "== obuf2 maxreg 84 stages 2 tc 8"; HB_WAVE_OBUF=2 $B --maxreg 84 --stages 2 --grid-tc 8 2>&1 | show
