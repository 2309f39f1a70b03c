#!/bin/bash
# round 2, GPU trip: lane engine v2 (CTA regrouping by handler) -- parity, throughput, one ncu capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_lane.py -x -q 2>&1 | tail -4
timeout 600 python scripts/gpu_lane_perf.py sparse_zi_100 16384 lane 2>&1 | grep -v Warning | tail -1
timeout 900 python scripts/gpu_lane_perf.py sparse_zi_100 65536 lane 2>&1 | grep -v Warning | tail -1
timeout 900 python scripts/gpu_lane_perf.py sparse_zi_1000 4096 lane warp 2>&1 | grep -v Warning | tail -3
timeout 900 python scripts/gpu_lane_perf.py sparse_zi_1000 16384 lane 2>&1 | grep -v Warning | tail -1
timeout 900 python scripts/gpu_lane_perf.py sparse_zi_1000 32768 lane 2>&1 | grep -v Warning | tail -1
timeout 900 python scripts/gpu_lane_perf.py rmsc03_2h 4096 lane 2>&1 | grep -v Warning | tail -1
timeout 900 python scripts/gpu_lane_perf.py rmsc03_2h 16384 lane 2>&1 | grep -v Warning | tail -1
timeout 600 python scripts/gpu_lane_perf.py sparse_zi_100 8192 lane > gpurun_out/r2c_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:lane_sim_kernel -c 1 -o gpurun_out/r2c_lane_zi100x8192 \
    python scripts/gpu_lane_perf.py sparse_zi_100 8192 lane > gpurun_out/r2c_ncu.log 2>&1
tail -1 gpurun_out/r2c_plain.log; tail -2 gpurun_out/r2c_ncu.log
