#!/bin/bash
# round 2: the driver's scaling command at N = 8 (headline + e2e with the concurrent PCIe probe + config-5 secondary)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/round2_box_8gpu.txt 2>&1; free -g >> gpurun_out/round2_box_8gpu.txt; nproc >> gpurun_out/round2_box_8gpu.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/round2_bench_default_8gpu.json 2> gpurun_out/round2_bench_default_8gpu.err
tail -c 3500 gpurun_out/round2_bench_default_8gpu.json; tail -5 gpurun_out/round2_bench_default_8gpu.err
