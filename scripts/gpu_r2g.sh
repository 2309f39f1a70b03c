#!/bin/bash
# round 2, GPU trip: inlined hot helpers in the lane kernels (A/B against out-of-line), memory diet -> more instances
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_lane.py -x -q 2>&1 | tail -3
for lib in libabides_b200_prof.so libabides_b200_prof_noinl.so; do
  ABIDES_B200_LIB=abides_b200/csrc/$lib timeout 600 python scripts/gpu_lane_prof.py sparse_zi_1000 8192 2>&1 | grep -v Warning | tail -1
  ABIDES_B200_LIB=abides_b200/csrc/$lib timeout 600 python scripts/gpu_lane_prof.py sparse_zi_1000 32768 2>&1 | grep -v Warning | tail -1
  ABIDES_B200_LIB=abides_b200/csrc/$lib timeout 600 python scripts/gpu_lane_prof.py rmsc03_2h 4096 2>&1 | grep -v Warning | tail -1
done
ABIDES_LANE_CTA=64 ABIDES_B200_LIB=abides_b200/csrc/libabides_b200_prof_l64.so timeout 600 python scripts/gpu_lane_prof.py sparse_zi_1000 32768 2>&1 | grep -v Warning | tail -1
timeout 900 python scripts/gpu_lane_perf.py sparse_zi_1000 32768 lane 2>&1 | grep -v Warning | tail -1 | cut -c1-330
timeout 900 python scripts/gpu_lane_perf.py sparse_zi_1000 45056 lane 2>&1 | grep -v Warning | tail -1 | cut -c1-330
timeout 900 python scripts/gpu_lane_perf.py rmsc03_2h 32768 lane 2>&1 | grep -v Warning | tail -1 | cut -c1-330
timeout 600 python scripts/gpu_lane_perf.py sparse_zi_100 131072 lane 2>&1 | grep -v Warning | tail -1 | cut -c1-330
