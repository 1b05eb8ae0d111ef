#!/bin/bash
mkdir -p gpurun_out
run() { timeout 300 python bench.py --workload exphydro_grid_2048x2048_x_60d "$@" --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys,os
d=json.loads(sys.stdin.readline()); r=d['roofline']; print('$*', 'ms', round(d['ms_per_step'],3), 'frac', round(r['frac'],3), 'regs', r.get('registers'), 'blocks/SM', r.get('blocks_per_sm'), 'spill', r.get('spill_bytes'), 'tc', d['config'].get('grid_steps_per_launch'))" | tee -a gpurun_out/r2_wave_occ2.txt; }
run --maxreg 255
run --maxreg 255 --grid-tc 4
run --maxreg 170
run --maxreg 170 --grid-tc 6
run --maxreg 128
run --block 64
run --block 64 --maxreg 255
run --block 256
