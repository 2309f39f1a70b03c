#!/bin/bash
# round 2, GPU trip: near buffer in front of the event heap -- parity, phase breakdown, throughput
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_lane.py tests/test_dense_oracle.py tests/test_book_replay.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python scripts/gpu_lane_prof.py sparse_zi_1000 8192 2>&1 | grep -v Warning | tail -1
timeout 600 python scripts/gpu_lane_prof.py sparse_zi_1000 32768 2>&1 | grep -v Warning | tail -1
timeout 600 python scripts/gpu_lane_prof.py rmsc03_2h 4096 2>&1 | grep -v Warning | tail -1
timeout 900 python scripts/gpu_lane_perf.py sparse_zi_1000 45056 lane 2>&1 | grep -v Warning | tail -1 | cut -c1-330
timeout 900 python scripts/gpu_lane_perf.py rmsc03_2h 16384 lane 2>&1 | grep -v Warning | tail -1 | cut -c1-330
timeout 600 python scripts/gpu_lane_perf.py sparse_zi_100 65536 lane 2>&1 | grep -v Warning | tail -1 | cut -c1-330
