#!/bin/bash
# end-of-session validation: smoke(), whole GPU suite, default bench line (with e2e + cpu baseline), reference arm
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 2500 gpurun_out/bench_default.json; tail -2 gpurun_out/bench_default.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 900 gpurun_out/bench_reference.json
