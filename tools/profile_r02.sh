#!/bin/bash
# Round-2 profile capture: every ncu run follows a plain run of the same command that exited 0.
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_1M.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_exphydro_1M.csv $CMD > gpurun_out/ncu_launches.log 2>&1
CMDG="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --workload exphydro_grid_2048x2048_x_60d"
$CMDG > gpurun_out/plain_grid.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_grid_wave.csv $CMDG > gpurun_out/ncu_launches_grid.log 2>&1
CMDR="python bench.py --workload raster_ingest_4096 --steps 2"
$CMDR > gpurun_out/plain_raster.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hb_pack_kernel -s 3 -c 1 -o gpurun_out/r02_pack $CMDR > gpurun_out/ncu_pack.log 2>&1
tail -2 gpurun_out/ncu_pack.log
