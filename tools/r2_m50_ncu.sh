#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --workload m50_50k_basins_x_3650d --steps 2 --warmup 1 --no-cpu --no-e2e"
# how long does a step take with 1, 2, 3 resident warps per sub-partition (148 SMs x 4 x 32 = 18 944 basins per layer of warps)?
for n in 18944 37888 50000 56832; do
  timeout 200 python bench.py --workload m50_50k_basins_x_3650d --basins $n --steps 3 --warmup 1 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('basins', d['config']['basins_per_gpu'], 'kernel_ms', d['roofline']['kernel_ms'])" | tee -a gpurun_out/round2_m50_warps_per_subpartition.txt
done
timeout 300 $CMD > gpurun_out/r2_m50_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hb_stepper -s 1 -c 1 -o gpurun_out/round2_m50_f64 -f $CMD > gpurun_out/ncu_m50.log 2>&1
tail -3 gpurun_out/ncu_m50.log
