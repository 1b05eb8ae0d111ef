#!/bin/bash
# final captures after the forcing-ring fence: launch list + ncu --set full of the headline kernel and of the wave kernel
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_1M.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_exphydro_1M.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain_1M_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hb_stepper -s 4 -c 1 -o gpurun_out/r02_exphydro_1M -f $CMD > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
CMDG="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --workload exphydro_grid_2048x2048_x_60d"
$CMDG > gpurun_out/plain_grid.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_grid_wave.csv $CMDG > gpurun_out/ncu_launches_grid.log 2>&1
$CMDG > gpurun_out/plain_grid_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hb_grid_wave -s 5 -c 1 -o gpurun_out/r02_wave_final -f $CMDG > gpurun_out/ncu_wave.log 2>&1
tail -1 gpurun_out/ncu_wave.log
