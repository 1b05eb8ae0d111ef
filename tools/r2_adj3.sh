#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_adjoint.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2_adjoint_tests.log
timeout 900 python bench.py --workload gradient_m50_adjoint --steps 3 --warmup 2 > gpurun_out/r2_bench_adjoint_m50.json 2> gpurun_out/r2_bench_adjoint_m50.err
tail -c 1200 gpurun_out/r2_bench_adjoint_m50.json; tail -3 gpurun_out/r2_bench_adjoint_m50.err
