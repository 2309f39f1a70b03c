#!/bin/bash
# round 2, GPU trip: seeded sweeps on the device, occupancy scaling of the lane engine, one ncu capture of its kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lane.py -x -q -k "seeded" 2>&1 | tail -5
timeout 600 python scripts/gpu_lane_perf.py sparse_zi_100 16384 lane 2>&1 | grep -v Warning | tail -1
timeout 900 python scripts/gpu_lane_perf.py sparse_zi_100 65536 lane 2>&1 | grep -v Warning | tail -1
timeout 900 python scripts/gpu_lane_perf.py sparse_zi_1000 16384 lane warp 2>&1 | grep -v Warning | tail -3
timeout 900 python scripts/gpu_lane_perf.py sparse_zi_1000 32768 lane 2>&1 | grep -v Warning | tail -1
timeout 900 python scripts/gpu_lane_perf.py rmsc03_2h 4096 lane warp 2>&1 | grep -v Warning | tail -3
timeout 900 python scripts/gpu_lane_perf.py rmsc03_2h 16384 lane 2>&1 | grep -v Warning | tail -1
timeout 600 python scripts/gpu_lane_perf.py sparse_zi_100 8192 lane > gpurun_out/r2b_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:lane_sim_kernel -c 1 -o gpurun_out/r2b_lane_zi100x8192 \
    python scripts/gpu_lane_perf.py sparse_zi_100 8192 lane > gpurun_out/r2b_ncu.log 2>&1
tail -2 gpurun_out/r2b_plain.log; tail -3 gpurun_out/r2b_ncu.log
