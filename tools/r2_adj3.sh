#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_adjoint.py tests/test_gpu_calibration.py -m gpu -x -q -k "adjoint or hybrid or reference" 2>&1 | tail -3 | tee gpurun_out/r2_adj3_tests.log
for mb in 3 2; do
  HB_ADJ_MINBLOCKS=$mb timeout 300 python bench.py --workload gradient_m50_adjoint --steps 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('adjoint minblocks=$mb ms', round(d['ms_per_step'],2), 'regs', d.get('registers'), 'spill', d.get('spill_bytes'), 'blocks', d.get('blocks_per_sm'), 'fwd', round(d.get('forward_trajectory_ms',0),2))" | tee -a gpurun_out/r2_adj3.txt
done
