python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for w in exphydro_1M_basins_x_365d m50_50k_basins_x_3650d gr4j_uh_100k_basins_x_3650d hbv_ensemble_1M_members_x_7305d; do
  echo "== $w"
  python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu --workload $w 2>&1 | tail -1 | python -c "
import sys,json,numpy as np
d=json.loads(sys.stdin.read())
r=d['roofline']; k=d['kernel_ms_each']; print('value %.3e frac %.3f regs %d blocks/SM %d spill %d first5 %.3f last10 %.3f min %.3f' % (d['value'], r['frac'], r['registers'], r['blocks_per_sm'], r['spill_bytes'], np.mean(k[1:6]), np.mean(k[-10:]), min(k)))"
done
