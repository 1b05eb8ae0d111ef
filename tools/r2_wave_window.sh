#!/bin/bash
mkdir -p gpurun_out
run() { timeout 300 python bench.py "$@" --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']; print('$*', 'ms', round(d['ms_per_step'],3), 'frac', round(r['frac'],3), d['config'].get('grid_split',''), d['config'].get('grid_steps_per_launch',''))" | tee -a gpurun_out/r2_wave_window.txt; }
for tc in 4 8 12 16 32; do run --workload exphydro_grid_512x512_x_60d --grid-tc $tc; done
for tc in 4 6 8 12; do run --workload exphydro_grid_2048x2048_x_60d --grid-tc $tc; done
for pn in 2 4 8; do for tc in 6 10 15 20; do run --workload exphydro_grid_2048x2048_x_60d --grid-panels $pn --grid-tc $tc; done; done
