#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/r2_tc_probe.log
