#!/bin/bash
# N GPUs (default 2): sharded tests, then the grid workload as row strips and as column panels, then the default workload
NG=${1:-2}
timeout 900 python -m pytest tests/test_gpu_multi.py -q -x -m gpu 2>&1 | tail -4
TR="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $NG --steps 5 --warmup 3 --no-cpu"
show() { tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    r=d['roofline']; print('n_gpus %d value %.3e frac %.3f kernel_ms %.3f ms/step %.3f %s' % (d['n_gpus'], d['value'], r['frac'], r['kernel_ms'], d['ms_per_step'], d['config'].get('grid_split', '')))
except Exception as e: print('FAILED', e)"; }
echo "== grid wave, row strips"; $TR --workload exphydro_grid_2048x2048_x_60d --grid-kernel wave --grid-split rows 2>&1 | show
echo "== grid wave, column panels"; $TR --workload exphydro_grid_2048x2048_x_60d --grid-kernel wave --grid-split cols 2>&1 | show
echo "== exphydro 1M"; $TR --no-e2e 2>&1 | show
