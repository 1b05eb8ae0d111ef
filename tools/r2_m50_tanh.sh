#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/r2_fp64_pipe 2>&1 | tee gpurun_out/round2_fp64_pipe_microbench.txt
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_adjoint.py tests/test_gpu_gradient.py -m gpu -x -q -k "tc or m50 or neural or adjoint or gradient" 2>&1 | tail -8 | tee gpurun_out/r2_m50_tanh_tests.log
timeout 300 python bench.py --workload m50_50k_basins_x_3650d --steps 5 --no-cpu --no-e2e > gpurun_out/r2_bench_m50_f64_tanh10.json 2>&1; tail -c 700 gpurun_out/r2_bench_m50_f64_tanh10.json
HB_EXTRA_DEFINES="#define HB_TANH_EXACT 1" timeout 300 python bench.py --workload m50_50k_basins_x_3650d --steps 5 --no-cpu --no-e2e 2>&1 | tail -c 300
