#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/round2_gpu_tests_full.log
for w in exphydro_1M_basins_x_365d gr4j_uh_100k_basins_x_3650d hbv_100k_basins_x_730d exphydro_64k_basins_x_365d; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-e2e --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']; print('$w', 'ms', round(d['ms_per_step'],3), 'frac', round(r['frac'],3), 'regs', r.get('registers'), 'block', r.get('block'), 'blocks', r.get('blocks_per_sm'), 'smem', r.get('dyn_smem'))" | tee -a gpurun_out/r2_newdefaults.txt
done
timeout 300 python bench.py --workload exphydro_grid_2048x2048_x_60d --grid-kernel generic --steps 5 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('generic 3-pass ms', round(d['ms_per_step'],3))" | tee -a gpurun_out/r2_newdefaults.txt
s=$(date +%s); timeout 900 python bench.py > gpurun_out/round2_bench_default_final.json 2> gpurun_out/round2_bench_default_final.err; echo "default arm rc=$? wall=$(( $(date +%s) - s ))s" | tee -a gpurun_out/r2_newdefaults.txt
python -c "
import json
d=json.loads(open('gpurun_out/round2_bench_default_final.json').read().strip().splitlines()[-1])
print('default: value %.4g frac %.3f e2e %.4g pcie_frac %.3f pageable %.4g secondary %.2f ms' % (d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['pcie_frac'], d['e2e']['pageable']['value'], d['secondary']['ms_per_step']))" | tee -a gpurun_out/r2_newdefaults.txt
