# Round-1 profile capture (run under gpurun): every ncu run is preceded by a plain run of the same command.
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_1M.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r01_exphydro_1M.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain_1M_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hb_stepper -s 4 -c 1 -o gpurun_out/prof_r01_exphydro_1M $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
CMD2="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --workload m50_50k_basins_x_3650d"
$CMD2 > gpurun_out/plain_m50.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hb_stepper -s 4 -c 1 -o gpurun_out/prof_r01_m50 $CMD2 > gpurun_out/ncu_m50.log 2>&1
tail -2 gpurun_out/ncu_m50.log
CMD3="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --workload exphydro_grid_2048x2048_x_60d"
$CMD3 > gpurun_out/plain_grid.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hb_route_step -s 200 -c 2 -o gpurun_out/prof_r01_route $CMD3 > gpurun_out/ncu_route.log 2>&1
tail -2 gpurun_out/ncu_route.log
CMD4="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --workload gr4j_uh_100k_basins_x_3650d"
$CMD4 > gpurun_out/plain_gr4j.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hb_stepper -s 4 -c 1 -o gpurun_out/prof_r01_gr4j $CMD4 > gpurun_out/ncu_gr4j.log 2>&1
tail -2 gpurun_out/ncu_gr4j.log
