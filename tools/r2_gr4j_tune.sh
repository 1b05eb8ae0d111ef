#!/bin/bash
mkdir -p gpurun_out
run() { timeout 300 python bench.py --workload $W "$@" --steps 8 --warmup 3 --no-cpu --no-e2e --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']; print('$W $*', 'ms', round(d['ms_per_step'],3), 'frac', round(r['frac'],3), 'regs', r.get('registers'), 'blocks', r.get('blocks_per_sm'), 'smem', r.get('dyn_smem'), 'spill', r.get('spill_bytes'))" | tee -a gpurun_out/r2_gr4j_tune.txt; }
W=gr4j_uh_100k_basins_x_3650d
run
run --stages 8
run --stages 2
run --maxreg 100
run --maxreg 100 --stages 8
run --block 64
run --block 256
W=hbv_100k_basins_x_730d
run
run --stages 8
run --maxreg 100
run --block 64
