#!/bin/bash
# round 2: 4 GPUs — multi-GPU tests with interior ranks (two neighbours: strips / panels, torch / NCCL / peer-store exchange),
# the windowed end-to-end path on a small workload, and the default bench line at N = 4.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/round2_box_4gpu.txt 2>&1; free -g >> gpurun_out/round2_box_4gpu.txt; nproc >> gpurun_out/round2_box_4gpu.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/round2_multi_gpu_tests_4gpu.log
HB_E2E_WINDOWS=3 timeout 300 python bench.py --workload exphydro_64k_basins_x_365d --steps 3 --no-cpu --no-secondary > gpurun_out/round2_bench_64k_windowed_e2e.json 2> gpurun_out/round2_bench_64k_windowed_e2e.err
tail -c 1500 gpurun_out/round2_bench_64k_windowed_e2e.json | cut -c1-1500; tail -3 gpurun_out/round2_bench_64k_windowed_e2e.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/round2_bench_default_4gpu.json 2> gpurun_out/round2_bench_default_4gpu.err
tail -c 3000 gpurun_out/round2_bench_default_4gpu.json; tail -5 gpurun_out/round2_bench_default_4gpu.err
