#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_guards.py -m gpu -x -q -k "grid or wave or panel or strip" 2>&1 | tail -5 | tee gpurun_out/r2_wave_repoll_tests.log
for w in exphydro_grid_2048x2048_x_60d exphydro_grid_512x512_x_60d; do
for d in 1 0; do
  HB_EXTRA_DEFINES="#define HB_WATCH_REPOLL $d" timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']; print('$w repoll $d', 'ms', round(d['ms_per_step'],3), 'frac', round(r['frac'],3), 'regs', r.get('registers'), 'spill', r.get('spill_bytes'), d['kernel_ms_each'])" | tee -a gpurun_out/r2_wave_repoll.txt
done; done
