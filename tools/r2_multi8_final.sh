#!/bin/bash
# end of round 2: the driver's scaling command at N = 8 on the final tree
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/round2_bench_default_8gpu_final.json 2> gpurun_out/round2_bench_default_8gpu_final.err
echo "rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/round2_bench_default_8gpu_final.json').read().strip().splitlines()[-1])
e=d['e2e']; s=d['secondary']
print('value %.4g frac %.3f e2e %.4g pcie_frac %.3f secondary ms %.3f exch %.4f clocks %s' % (d['value'], d['roofline']['frac'], e['value'], e.get('pcie_frac',0), s['ms_per_step'], s['exchange_share'], d['clocks']))
PY
tail -3 gpurun_out/round2_bench_default_8gpu_final.err
