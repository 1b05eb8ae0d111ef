#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gradient.py tests/test_gpu_adjoint.py tests/test_gpu_calibration.py -m gpu -x -q -k "not fuzz" 2>&1 | tail -3 | tee gpurun_out/r2_logt_tests.log
for w in hbv_ensemble_1M_members_x_7305d hbv_100k_basins_x_730d; do
  timeout 300 python bench.py --workload $w --steps 8 --warmup 3 --no-cpu --no-e2e --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']; print('$w', 'ms', round(d['ms_per_step'],3), 'value', '%.4g' % d['value'], 'frac', round(r['frac'],3), 'regs', r.get('registers'), 'blocks', r.get('blocks_per_sm'))" | tee -a gpurun_out/r2_logt.txt
done
