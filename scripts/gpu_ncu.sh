#!/bin/bash
# ncu capture of sim_kernel on a 1024-instance batch: $1 = tag, $2 = config (default sparse_zi_100)
TAG=${1:-rX}; CFG=${2:-sparse_zi_100}
mkdir -p gpurun_out
timeout 300 python scripts/gpu_zi1000.py 1024 $CFG > gpurun_out/${TAG}_plain.log 2>&1 || exit 1
tail -8 gpurun_out/${TAG}_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sim_kernel -c 1 -o gpurun_out/${TAG}_${CFG}_x1024 \
  python scripts/gpu_zi1000.py 1024 $CFG > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
