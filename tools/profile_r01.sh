# Round-1 profile capture (run under gpurun): plain run first, then ncu on the same command.
set -x
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_1M.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r01_exphydro_1M.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain_1M_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hb_stepper -s 4 -c 1 -o gpurun_out/prof_r01_exphydro_1M $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
for w in gr4j_uh_100k_basins_x_3650d hbv_100k_basins_x_730d m50_50k_basins_x_3650d; do
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --workload $w 2>&1 | tail -1 > gpurun_out/bench_$w.json
  python -c "
import json,sys
d=json.load(open('gpurun_out/bench_$w.json')); r=d['roofline']
print('$w', 'value %.3e frac %.3f kernel_ms %.3f regs %d blocks/SM %d spill %d' % (d['value'], r['frac'], r['kernel_ms'], r['registers'], r['blocks_per_sm'], r['spill_bytes']))"
done
python -c "
import hydromodels_jl_b200 as hm, ctypes
hm.engine.init(0)
v=ctypes.c_double(); print('fp64 probe rc', hm.engine.lib().hb_probe_fp64_tflops(ctypes.byref(v)), 'TFLOP/s', v.value)"
