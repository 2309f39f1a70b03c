#!/bin/bash
# round 2, GPU trip: a warp of the lane CTA dedicated to the heavy handler classes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lane.py -x -q -k "golden or sweep or philox or pause" 2>&1 | tail -2
timeout 600 python scripts/gpu_lane_prof.py sparse_zi_1000 8192 2>&1 | grep -v Warning | tail -1
timeout 900 python scripts/gpu_lane_perf.py sparse_zi_1000 45056 lane 2>&1 | grep -v Warning | tail -1 | cut -c1-330
timeout 600 python scripts/gpu_lane_perf.py sparse_zi_100 65536 lane 2>&1 | grep -v Warning | tail -1 | cut -c1-330
timeout 900 python scripts/gpu_lane_perf.py rmsc03_2h 16384 lane 2>&1 | grep -v Warning | tail -1 | cut -c1-330
