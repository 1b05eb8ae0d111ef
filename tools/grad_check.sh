#!/bin/bash
# gradient tests first (new code), then the whole GPU suite (the stepper template changed), then two short bench lines
timeout 600 python -m pytest tests/test_gpu_gradient.py -q -x 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_gradient.py 2>&1 | tail -5
timeout 300 python bench.py --workload gradient_exphydro --steps 3 --warmup 1 > gpurun_out/bench_gradient.json 2> gpurun_out/bench_gradient.err; tail -c 1200 gpurun_out/bench_gradient.json; tail -3 gpurun_out/bench_gradient.err
timeout 200 python bench.py --workload irf_route --steps 3 > gpurun_out/bench_irf.json 2> gpurun_out/bench_irf.err; tail -c 700 gpurun_out/bench_irf.json
