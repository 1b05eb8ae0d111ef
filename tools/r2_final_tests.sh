#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s -k "parity_without" 2>&1 | grep "\[parity\]" | tee gpurun_out/round2_parity_without_floor.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/round2_gpu_tests_full.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/round2_smoke.log
