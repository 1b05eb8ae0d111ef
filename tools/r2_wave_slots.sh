#!/bin/bash
mkdir -p gpurun_out
w=exphydro_grid_2048x2048_x_60d
for d in 8 6 5 3 2; do
  HB_EXTRA_DEFINES="#define HB_USLOTS $d" timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']; print('$w uslots $d', 'ms', round(d['ms_per_step'],3), 'frac', round(r['frac'],3), 'regs', r.get('registers'), 'spill', r.get('spill_bytes'), d['kernel_ms_each'])" | tee -a gpurun_out/r2_wave_slots.txt
done
