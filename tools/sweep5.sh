python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -2
for w in exphydro_1M_basins_x_365d gr4j_uh_100k_basins_x_3650d hbv_100k_basins_x_730d m50_50k_basins_x_3650d hbv_ensemble_1M_members_x_7305d exphydro_grid_2048x2048_x_60d; do
  python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --workload $w 2>&1 | tail -1 | python -c "
import sys,json,numpy as np
d=json.loads(sys.stdin.read()); r=d['roofline']; k=d['kernel_ms_each']; print('$w value %.3e frac %.3f min %.3f mean %.3f regs %d blocks %d' % (d['value'], r['frac'], min(k), np.mean(k), r['registers'], r['blocks_per_sm']))"
done
