#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rasters.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2_ingest_tests.log
timeout 600 python bench.py --workload raster_ingest_4096 --steps 5 > gpurun_out/round2_bench_raster_ingest.json 2> gpurun_out/round2_bench_raster_ingest.err; tail -3 gpurun_out/round2_bench_raster_ingest.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/round2_bench_raster_ingest.json').read().strip().splitlines()[-1])
k=d['roofline']['kernels']
for n,v in k.items():
    if n!='ingest_from_host': print(n, round(v['ms'],3), round(v['frac_of_hbm_peak'],3))
print(json.dumps(k['ingest_from_host'], indent=None)[:900])
PY
