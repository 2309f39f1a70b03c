#!/bin/bash
# round 2, GPU trip: lane engine v2.2 (run-ahead of light events) -- parity, run-ahead sweep, phase breakdown
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_lane.py -x -q 2>&1 | tail -3
for ra in 0 3 6 12 24; do
  echo "run_ahead=$ra"
  ABIDES_LANE_RUN_AHEAD=$ra timeout 900 python scripts/gpu_lane_perf.py sparse_zi_1000 16384 lane 2>&1 | grep -v Warning | tail -1 | cut -c1-330
done
for ra in 0 6 24; do
  echo "run_ahead=$ra"
  ABIDES_LANE_RUN_AHEAD=$ra timeout 900 python scripts/gpu_lane_perf.py rmsc03_2h 8192 lane 2>&1 | grep -v Warning | tail -1 | cut -c1-330
done
timeout 900 python scripts/gpu_lane_perf.py sparse_zi_1000 32768 lane 2>&1 | grep -v Warning | tail -1 | cut -c1-330
timeout 900 python scripts/gpu_lane_perf.py sparse_zi_100 65536 lane 2>&1 | grep -v Warning | tail -1 | cut -c1-330
timeout 600 python scripts/gpu_lane_prof.py sparse_zi_1000 8192 2>&1 | grep -v Warning | tail -1
timeout 600 python scripts/gpu_lane_prof.py rmsc03_2h 4096 2>&1 | grep -v Warning | tail -1
