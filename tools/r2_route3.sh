#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "route or grid or generic or routed" 2>&1 | tail -12 | tee gpurun_out/r2_route3_tests.log
for rp in 3 2; do
  timeout 300 python bench.py --workload exphydro_grid_2048x2048_x_60d --grid-kernel generic --route-passes $rp --steps 5 --no-cpu > gpurun_out/r2_bench_grid_generic_p$rp.json 2> gpurun_out/r2_bench_grid_generic_p$rp.err
  tail -c 1100 gpurun_out/r2_bench_grid_generic_p$rp.json; tail -3 gpurun_out/r2_bench_grid_generic_p$rp.err
done
