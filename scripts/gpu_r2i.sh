#!/bin/bash
# round 2, GPU trip: validation -- micro-optimised lane handlers, warp engine at bench sizes, bench.py, the whole GPU suite
mkdir -p gpurun_out
timeout 900 python scripts/gpu_lane_perf.py sparse_zi_1000 40960 lane 2>&1 | grep -v Warning | tail -1 | cut -c1-330
timeout 600 python scripts/gpu_lane_prof.py sparse_zi_1000 8192 2>&1 | grep -v Warning | tail -1
timeout 900 python scripts/gpu_lane_perf.py rmsc03_2h 2048 warp 2>&1 | grep -v Warning | tail -1 | cut -c1-330
timeout 900 python scripts/gpu_lane_perf.py rmsc03_2h 3072 warp 2>&1 | grep -v Warning | tail -1 | cut -c1-330
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; tail -c 3000 gpurun_out/r2i_bench.json; tail -3 gpurun_out/r2i_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2i_bench_ref.json 2> gpurun_out/r2i_bench_ref.err; tail -c 600 gpurun_out/r2i_bench_ref.json
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
