python __graft_entry__.py --smoke 2>&1 | tail -1
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_final.log 2>&1; tail -1 gpurun_out/bench_final.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.4e frac %.3f traffic %s kernel_ms %.3f e2e %.3e cpu %.3e (%d cores) launches %d clocks %s' % (d['value'], r['frac'], r['traffic'], r['kernel_ms'], d['e2e']['value'], d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['gpu_launches'], d['clocks']))"
python bench.py --impl reference 2>&1 | tail -1 | cut -c1-160
for w in gr4j_uh_100k_basins_x_3650d gr4j_uh_100k_basins_x_3650d; do
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --workload $w 2>&1 | tail -1 | python -c "
import sys,json,numpy as np
d=json.loads(sys.stdin.read()); r=d['roofline']; k=d['kernel_ms_each']; print('$w value %.3e frac %.3f min %.3f mean %.3f' % (d['value'], r['frac'], min(k), np.mean(k)))"
done
