#!/bin/bash
# round 2, first GPU trip: lane-engine parity, then lane vs warp throughput
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; free -g | head -2; nproc
timeout 1500 python -m pytest tests/test_gpu_lane.py -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "untraced" 2>&1 | tail -8
for n in 1024 4096; do timeout 900 python scripts/gpu_lane_perf.py sparse_zi_1000 $n lane warp 2>&1 | grep -v Warning | tail -4; done
timeout 900 python scripts/gpu_lane_perf.py sparse_zi_1000 8192 lane 2>&1 | grep -v Warning | tail -3
timeout 900 python scripts/gpu_lane_perf.py sparse_zi_100 16384 lane warp 2>&1 | grep -v Warning | tail -4
timeout 900 python scripts/gpu_lane_perf.py rmsc03_2h 512 lane warp 2>&1 | grep -v Warning | tail -4
