#!/bin/bash
# round 2 evidence: one bench JSON line per BASELINE config (1 GPU), launch list of the default bench, DRAM traffic of the
# generic routed path's three passes, ncu --set full of the M50 Float32 tcgen05 kernel and of the adjoint kernel.
mkdir -p gpurun_out
B="python bench.py --no-cpu --no-secondary"
$B --workload gr4j_uh_100k_basins_x_3650d --steps 5 --no-e2e > gpurun_out/round2_bench_config2_gr4j.json 2>/dev/null
$B --workload hbv_ensemble_1M_members_x_7305d --steps 3 > gpurun_out/round2_bench_config3_hbv_ensemble.json 2>/dev/null
$B --workload m50_50k_basins_x_3650d --steps 5 --no-e2e > gpurun_out/round2_bench_config4_m50_f64.json 2>/dev/null
$B --workload exphydro_grid_2048x2048_x_60d --steps 5 > gpurun_out/round2_bench_config5_grid2048_wave.json 2>/dev/null
$B --workload exphydro_grid_4096x4096_total_x_60d --steps 3 > gpurun_out/round2_bench_config5_grid4096_1gpu.json 2>/dev/null
python bench.py --workload exphydro_single_basin_365d --steps 5 > gpurun_out/round2_bench_config1_single_basin.json 2>/dev/null
for f in gpurun_out/round2_bench_config*.json; do echo $f; cut -c1-400 $f; done
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-secondary"
$CMD > gpurun_out/plain_default.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/round2_launches_exphydro_1M.csv $CMD > gpurun_out/ncu_launches.log 2>&1
CMDG="python bench.py --workload exphydro_grid_2048x2048_x_60d --grid-kernel generic --steps 1 --warmup 3 --no-cpu --no-secondary"
$CMDG > gpurun_out/plain_generic.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 189 -c 63 --csv --log-file gpurun_out/round2_generic_route_3pass_traffic.csv $CMDG > gpurun_out/ncu_generic.log 2>&1
CMDM="python bench.py --workload m50_50k_basins_x_3650d --precision f32 --nn-tensor 1 --steps 1 --warmup 3 --no-cpu --no-secondary"
$CMDM > gpurun_out/plain_m50f32.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hb_stepper -s 3 -c 1 -o gpurun_out/round2_m50_f32_tcgen05 -f $CMDM > gpurun_out/ncu_m50f32.log 2>&1
tail -2 gpurun_out/ncu_m50f32.log
CMDA="python tools/r2_adj_small.py"
$CMDA > gpurun_out/plain_adj.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hb_adjoint -c 1 -o gpurun_out/round2_m50_adjoint -f $CMDA > gpurun_out/ncu_adj.log 2>&1
tail -2 gpurun_out/ncu_adj.log; cat gpurun_out/plain_adj.log | tail -1
