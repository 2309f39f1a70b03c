#!/bin/bash
# round 2, GPU trip: the whole -m gpu suite on the restored tree, then BASELINE configs[4] -- the rmsc03 parameter grid,
# 16 points x 1024 seeds = 16 384 instances in one batch on one GPU
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
( time timeout 840 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2l_gpu_tests.log 2>&1; tail -4 gpurun_out/r2l_gpu_tests.log
timeout 540 python bench.py --workload rmsc03_grid --instances 16384 --steps 2 --warmup 1 --e2e-steps 1 --also '' \
    > gpurun_out/r2l_bench_grid_x16384.json 2> gpurun_out/r2l_bench_grid_x16384.err; tail -c 1500 gpurun_out/r2l_bench_grid_x16384.json; tail -3 gpurun_out/r2l_bench_grid_x16384.err
