#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_calibration.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2_gpu_tests3.log
python bench.py --workload irf_route --steps 5 > gpurun_out/r2_bench_irf_tree.json 2> gpurun_out/r2_bench_irf_tree.err
tail -c 1800 gpurun_out/r2_bench_irf_tree.json; tail -3 gpurun_out/r2_bench_irf_tree.err
