#!/bin/bash
# A/B of the opt-in early poll of the wave grid kernel (HB_PREFETCH_POLL): parity first, then the 2048^2 x 60 d bench line
export HB_EXTRA_DEFINES="#define HB_PREFETCH_POLL 1"
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x -k "wave or wide_grid or fullsize or grid" 2>&1 | tail -3
show() { tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d.get('roofline') or {}
print('value %.3e  frac %s kernel_ms %s' % (d['value'], r.get('frac'), r.get('kernel_ms')))"; }
for rep in 1 2; do
echo "== prefetch"; timeout 200 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --workload exphydro_grid_2048x2048_x_60d 2>&1 | show
echo "== baseline"; HB_EXTRA_DEFINES="" timeout 200 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --workload exphydro_grid_2048x2048_x_60d 2>&1 | show
done
