#!/bin/bash
# ncu --set full of hb_stepper for the HBV ensemble (objective mode), M50 and GR4J workloads (each after a plain run)
for w in hbv_ensemble_1M_members_x_7305d m50_50k_basins_x_3650d gr4j_uh_100k_basins_x_3650d; do
  CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --workload $w"
  $CMD > gpurun_out/plain_$w.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:hb_stepper -s 3 -c 1 -o gpurun_out/r02_$w -f $CMD > gpurun_out/ncu_$w.log 2>&1
  tail -1 gpurun_out/ncu_$w.log
done
