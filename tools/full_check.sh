#!/bin/bash
# full GPU test suite, every bench workload (one line each), default bench line with e2e + cpu baseline
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
show() { tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d.get('roofline') or {}
    print('value %.3e %s  frac %s kernel_ms %s  e2e %s' % (d['value'], d['unit'], r.get('frac'), r.get('kernel_ms'), (d.get('e2e') or {}).get('value')))
except Exception as e: print('FAILED', e)"; }
for w in gr4j_uh_100k_basins_x_3650d hbv_ensemble_1M_members_x_7305d m50_50k_basins_x_3650d exphydro_grid_2048x2048_x_60d; do
  echo "== $w"; timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --workload $w 2>&1 | show
done
echo "== raster ingest"; timeout 200 python bench.py --workload raster_ingest_4096 --steps 5 > gpurun_out/bench_raster.json 2>gpurun_out/bench_raster.err; python -c "
import json; d=json.load(open('gpurun_out/bench_raster.json')); print({k:(round(v['ms'],3), round(v['frac_of_hbm_peak'],3)) for k,v in d['roofline']['kernels'].items()})"
echo "== calibration"; timeout 300 python bench.py --workload calibrate_exphydro --steps 20 > gpurun_out/bench_calib.json 2>gpurun_out/bench_calib.err; tail -c 600 gpurun_out/bench_calib.json
echo "== default bench"; timeout 600 python bench.py > gpurun_out/bench_default.json 2>gpurun_out/bench_default.err; cat gpurun_out/bench_default.json | show
