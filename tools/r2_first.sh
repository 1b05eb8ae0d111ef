#!/bin/bash
# round 2, first GPU call: ring / staging tests, box inventory, default bench with the new end-to-end leg
mkdir -p gpurun_out
{ nproc; free -g; nvidia-smi topo -m; lscpu | head -25; nvidia-smi --query-gpu=name,pcie.link.gen.max,pcie.link.width.max --format=csv; } > gpurun_out/r2_box.txt 2>&1
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
tail -c 3000 gpurun_out/r2_bench_default.json; tail -5 gpurun_out/r2_bench_default.err
