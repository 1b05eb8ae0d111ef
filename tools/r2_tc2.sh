#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -x -q -k "tc or float32 or m50 or neural" 2>&1 | tail -25 | tee gpurun_out/r2_tc_m50.log
for nt in 0 1; do
  timeout 300 python bench.py --workload m50_50k_basins_x_3650d --precision f32 --nn-tensor $nt --steps 5 --no-cpu > gpurun_out/r2_bench_m50_f32_nt$nt.json 2> gpurun_out/r2_bench_m50_f32_nt$nt.err
  tail -c 900 gpurun_out/r2_bench_m50_f32_nt$nt.json; tail -3 gpurun_out/r2_bench_m50_f32_nt$nt.err
done
timeout 300 python bench.py --workload m50_50k_basins_x_3650d --steps 5 --no-cpu --no-e2e > gpurun_out/r2_bench_m50_f64.json 2>&1; tail -c 700 gpurun_out/r2_bench_m50_f64.json
