#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_guards.py -m gpu -x -q -k "irf or conv or composition or segment or tree or raster" 2>&1 | tail -3 | tee gpurun_out/r2_irf_final_tests.log
timeout 300 python bench.py --workload irf_route --steps 5 --no-cpu > gpurun_out/round2_bench_irf_river_tree_tt32.json 2>/dev/null; tail -c 600 gpurun_out/round2_bench_irf_river_tree_tt32.json
timeout 300 python bench.py --workload irf_route_ragged --steps 5 --no-cpu > gpurun_out/round2_bench_irf_ragged_tt32.json 2>/dev/null; tail -c 300 gpurun_out/round2_bench_irf_ragged_tt32.json
