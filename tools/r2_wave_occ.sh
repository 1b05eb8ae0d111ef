#!/bin/bash
mkdir -p gpurun_out
run() { timeout 300 python bench.py --workload exphydro_grid_2048x2048_x_60d "$@" --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys,os
d=json.loads(sys.stdin.readline()); r=d['roofline']; print(os.environ.get('HB_WAVE_OBUF','-'), os.environ.get('HB_EXTRA_DEFINES','-'), '$*', 'ms', round(d['ms_per_step'],3), 'frac', round(r['frac'],3), 'regs', r.get('registers'), 'blocks/SM', r.get('blocks_per_sm'), 'spill', r.get('spill_bytes'))" | tee -a gpurun_out/r2_wave_occ.txt; }
export HB_EXTRA_DEFINES="#define HB_USLOTS 3"
run
HB_WAVE_OBUF=2 run
HB_WAVE_OBUF=2 run --maxreg 100
HB_WAVE_OBUF=2 run --maxreg 100 --grid-tc 12
HB_WAVE_OBUF=2 run --maxreg 84
HB_WAVE_OBUF=2 run --maxreg 84 --grid-tc 12
export HB_EXTRA_DEFINES="#define HB_USLOTS 2"
HB_WAVE_OBUF=2 run --maxreg 100
