#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_adjoint.py -m gpu -x -q -k "m50 or neural or adjoint or float32" 2>&1 | tail -4 | tee gpurun_out/r2_m50_recheck.log
timeout 300 python bench.py --workload m50_50k_basins_x_3650d --steps 5 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('m50 f64 kernel_ms', d['roofline']['kernel_ms'], d['kernel_ms_each'])" | tee gpurun_out/r2_m50_after.txt
s=$(date +%s); timeout 900 python bench.py --impl reference > gpurun_out/round2_bench_reference_default.json 2> gpurun_out/round2_bench_reference_default.err; echo "reference arm rc=$? wall=$(( $(date +%s) - s ))s" | tee gpurun_out/round2_default_bench_wall.txt
s=$(date +%s); timeout 900 python bench.py > gpurun_out/round2_bench_default_final.json 2> gpurun_out/round2_bench_default_final.err; echo "default arm rc=$? wall=$(( $(date +%s) - s ))s" | tee -a gpurun_out/round2_default_bench_wall.txt
tail -c 1500 gpurun_out/round2_bench_default_final.json
