#!/bin/bash
# BASELINE config 5: one 4096 x 4096 grid on NG GPUs, row strips vs column panels
NG=${1:-8}
TR="timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $NG --steps 5 --warmup 3 --no-cpu --workload exphydro_grid_4096x4096_total_x_60d --grid-kernel wave"
show() { tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    r=d['roofline']; print('n_gpus %d value %.3e frac/gpu %.3f ms/step %.3f %s' % (d['n_gpus'], d['value'], r['frac'], d['ms_per_step'], d['config'].get('scaling_note', '')))
except Exception as e: print('FAILED', e)"; }
echo "== strips"; $TR --grid-split rows --grid-tc 8 2>&1 | show
echo "== panels"; $TR --grid-split cols 2>&1 | show
