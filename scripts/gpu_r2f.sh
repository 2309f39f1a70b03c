#!/bin/bash
# round 2, GPU trip: CTA size of the lane engine (phase breakdown per size), Philox throughput mode
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lane.py -x -q -k "philox or sweep or seeded" 2>&1 | tail -3
for c in 32 64 128 256; do
  lib=abides_b200/csrc/libabides_b200_prof_l$c.so; [ $c = 128 ] && lib=abides_b200/csrc/libabides_b200_prof.so
  ABIDES_LANE_CTA=$c ABIDES_B200_LIB=$lib timeout 600 python scripts/gpu_lane_prof.py sparse_zi_1000 8192 2>&1 | grep -v Warning | tail -1
  ABIDES_LANE_CTA=$c ABIDES_B200_LIB=$lib timeout 600 python scripts/gpu_lane_prof.py sparse_zi_1000 32768 2>&1 | grep -v Warning | tail -1
done
for c in 32 64; do
  ABIDES_LANE_CTA=$c ABIDES_B200_LIB=abides_b200/csrc/libabides_b200_prof_l$c.so timeout 600 python scripts/gpu_lane_prof.py rmsc03_2h 4096 2>&1 | grep -v Warning | tail -1
done
ABIDES_PERF_RNG=philox timeout 900 python scripts/gpu_lane_perf.py sparse_zi_1000 32768 lane 2>&1 | grep -v Warning | tail -1 | cut -c1-330
ABIDES_PERF_RNG=philox timeout 900 python scripts/gpu_lane_perf.py sparse_zi_1000 98304 lane 2>&1 | grep -v Warning | tail -1 | cut -c1-330
