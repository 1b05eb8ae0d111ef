#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gradient.py tests/test_gpu_adjoint.py tests/test_gpu_tc.py tests/test_gpu_calibration.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2_kbank_tests.log
for w in hbv_ensemble_1M_members_x_7305d gr4j_uh_100k_basins_x_3650d exphydro_1M_basins_x_365d hbv_100k_basins_x_730d m50_50k_basins_x_3650d exphydro_grid_2048x2048_x_60d; do
  timeout 300 python bench.py --workload $w --steps 8 --warmup 3 --no-cpu --no-e2e --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']; print('$w', 'ms', round(d['ms_per_step'],3), 'value', '%.4g' % d['value'], 'frac', round(r['frac'],3), 'regs', r.get('registers'), 'clocks', d['clocks'].get('reasons'))" | tee -a gpurun_out/r2_kbank.txt
done
timeout 300 python bench.py --workload gradient_m50_adjoint --steps 3 --no-cpu 2>/dev/null | tail -c 400 | tee -a gpurun_out/r2_kbank.txt
