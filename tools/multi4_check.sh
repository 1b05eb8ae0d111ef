#!/bin/bash
# 4 GPUs: sharded tests (interior ranks), grid wave and default workload under torchrun
timeout 900 python -m pytest tests/test_gpu_multi.py -q -x -m gpu 2>&1 | tail -4
TR="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu"
show() { tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    r=d['roofline']; print('n_gpus %d value %.3e frac %.3f kernel_ms %.3f ms/step %.3f e2e %s' % (d['n_gpus'], d['value'], r['frac'], r['kernel_ms'], d['ms_per_step'], (d.get('e2e') or {}).get('value')))
except Exception as e: print('FAILED', e)"; }
echo "== grid wave x4"; $TR --workload exphydro_grid_2048x2048_x_60d --grid-kernel wave 2>&1 | show
echo "== exphydro 1M x4"; $TR 2>&1 | show
