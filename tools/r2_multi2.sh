#!/bin/bash
# round 2: 2-GPU check of the library communicator (NCCL + peer-store halo exchange) and the bench with its config-5 secondary line
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_2gpu.txt 2>&1; free -g >> gpurun_out/r2_topo_2gpu.txt; nproc >> gpurun_out/r2_topo_2gpu.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2_multi_tests_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
tail -c 2500 gpurun_out/r2_bench_2gpu.json; tail -5 gpurun_out/r2_bench_2gpu.err
