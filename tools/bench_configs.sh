for w in exphydro_grid_512x512_x_60d exphydro_grid_2048x2048_x_60d hbv_ensemble_1M_members_x_7305d gr4j_uh_100k_basins_x_3650d; do
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --workload $w > gpurun_out/bench_$w.log 2>&1
  tail -1 gpurun_out/bench_$w.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('$w', 'value %.3e frac %.3f kernel_ms %.3f regs %d blocks/SM %d spill %d' % (d['value'], r['frac'], r['kernel_ms'], r['registers'], r['blocks_per_sm'], r['spill_bytes']))
except Exception as e:
    print('$w FAILED', e)
"
done
tail -5 gpurun_out/bench_exphydro_grid_512x512_x_60d.log | cut -c1-400
