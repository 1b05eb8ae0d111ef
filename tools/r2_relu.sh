#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gradient.py tests/test_gpu_adjoint.py -m gpu -x -q -k "not fuzz" 2>&1 | tail -4 | tee gpurun_out/r2_relu_tests.log
for w in hbv_ensemble_1M_members_x_7305d gr4j_uh_100k_basins_x_3650d exphydro_1M_basins_x_365d hbv_100k_basins_x_730d; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-e2e --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']; print('$w', 'ms', round(d['ms_per_step'],3), 'value', '%.4g' % d['value'], 'frac', round(r['frac'],3), 'regs', r.get('registers'))" | tee -a gpurun_out/r2_relu.txt
done
