#!/bin/bash
# wave grid kernel: where does the time go?  (probes with HB_EXTRA_DEFINES produce invalid results by design)
B="timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --workload exphydro_grid_2048x2048_x_60d --grid-kernel wave"
show() { tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    r=d['roofline']; print('value %.3e frac %.3f kernel_ms %.3f regs %d blocks/SM %d spill %d tc %s' % (d['value'], r['frac'], r['kernel_ms'], r['registers'], r['blocks_per_sm'], r['spill_bytes'], d['config'].get('steps_per_launch')))
except Exception as e: print('FAILED', e)"; }
for tc in 1 2 4; do echo "== tc $tc"; $B --grid-tc $tc 2>&1 | show; done
echo "== no waits (SPIN_LIMIT 0)"; HB_EXTRA_DEFINES="#define HB_SPIN_LIMIT 0" $B 2>&1 | show
echo "== no fence"; HB_EXTRA_DEFINES="#define HB_PROBE_NO_FENCE 1" $B 2>&1 | show
echo "== no fence tc 4"; HB_EXTRA_DEFINES="#define HB_PROBE_NO_FENCE 1" $B --grid-tc 4 2>&1 | show
echo "== block 64"; $B --block 64 2>&1 | show
echo "== block 256"; $B --block 256 2>&1 | show
timeout 900 ncu --set full --import-source on --clock-control none -k hb_grid_wave -s 5 -c 1 -o gpurun_out/wave_tc16 -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --workload exphydro_grid_2048x2048_x_60d --grid-kernel wave > gpurun_out/wave_ncu.log 2>&1
tail -3 gpurun_out/wave_ncu.log
