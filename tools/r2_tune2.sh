#!/bin/bash
mkdir -p gpurun_out
run() { timeout 300 python bench.py --workload $W "$@" --steps 8 --warmup 3 --no-cpu --no-e2e --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']; print('$W $*', 'ms', round(d['ms_per_step'],3), 'frac', round(r['frac'],3), 'regs', r.get('registers'), 'blocks', r.get('blocks_per_sm'), 'smem', r.get('dyn_smem'))" | tee -a gpurun_out/r2_tune2.txt; }
for W in exphydro_1M_basins_x_365d gr4j_uh_100k_basins_x_3650d hbv_100k_basins_x_730d; do
 for b in 128 64; do for s in 4 6 8; do run --block $b --stages $s; done; done
done
