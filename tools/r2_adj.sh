#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_adjoint.py -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/r2_adjoint_tests.log
