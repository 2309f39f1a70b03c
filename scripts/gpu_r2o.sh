#!/bin/bash
# round 2, the last two GPU-minutes: one run each of the lane engine on sparse_zi_1000 x 45 056 with as much L1 as the SM
# can give (ABIDES_LANE_SMEM_CARVEOUT=0), with 256-instance CTAs, and as committed
mkdir -p gpurun_out
L=$PWD/abides_b200/csrc
export ABIDES_PERF_RUNS=1
ABIDES_LANE_SMEM_CARVEOUT=0 timeout 33 python scripts/gpu_lane_perf.py sparse_zi_1000 45056 lane 2>&1 | grep -v Warning | tail -1 | cut -c1-300 | tee gpurun_out/r2o_carveout0.json
ABIDES_B200_LIB=$L/libabides_b200_l256.so timeout 33 python scripts/gpu_lane_perf.py sparse_zi_1000 45056 lane 2>&1 | grep -v Warning | tail -1 | cut -c1-300 | tee gpurun_out/r2o_cta256.json
timeout 33 python scripts/gpu_lane_perf.py sparse_zi_1000 45056 lane 2>&1 | grep -v Warning | tail -1 | cut -c1-300 | tee gpurun_out/r2o_default.json
echo "elapsed $SECONDS s"
