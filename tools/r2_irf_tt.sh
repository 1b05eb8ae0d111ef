#!/bin/bash
mkdir -p gpurun_out
for tt in 32; do
 for w in irf_route irf_route_ragged; do
  HB_IRF_TT=$tt timeout 300 python bench.py --workload $w --steps 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$w TT=$tt ms', round(d['ms_per_step'],3), {k:(round(v,3) if isinstance(v,float) else v) for k,v in d.get('roofline',{}).items() if k in ('frac','achieved','kernel_ms')})" | tee -a gpurun_out/r2_irf_tt.txt
 done
 HB_IRF_TT=$tt timeout 600 python -m pytest -q tests/test_gpu_parity.py -m gpu -k "irf or conv or segment or tree" 2>&1 | tail -1
done
