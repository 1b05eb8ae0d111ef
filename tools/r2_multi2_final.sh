#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/round2_multi_gpu_tests_2gpu_final.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/round2_bench_default_2gpu_final.json 2> gpurun_out/round2_bench_default_2gpu_final.err
echo "bench rc=$?"; tail -c 1200 gpurun_out/round2_bench_default_2gpu_final.json
