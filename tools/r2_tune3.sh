#!/bin/bash
mkdir -p gpurun_out
run() { timeout 300 python bench.py --workload $W "$@" --steps 5 --warmup 3 --no-cpu --no-e2e --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']; print('$W $*', 'ms', round(d['ms_per_step'],3), 'regs', r.get('registers'), 'block', r.get('block'), 'blocks', r.get('blocks_per_sm'))" | tee -a gpurun_out/r2_tune3.txt; }
W=hbv_ensemble_1M_members_x_7305d
run; run --block 64; run --block 256
W=m50_50k_basins_x_3650d
run; run --block 64; run --block 96
