#!/bin/bash
# Round-1 profiling recipe (run under gpurun).  Each ncu run is preceded by the same command without ncu.
set -x
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --e2e-steps 1"
$BENCH > gpurun_out/r1_bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r1_launches.csv $BENCH > gpurun_out/r1_ncu_launches.log 2>&1
SMALL="python scripts/gpu_zi1000.py 1024 sparse_zi_100"
$SMALL > gpurun_out/r1_small_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sim_kernel -c 1 -o gpurun_out/r1_zi100x1024 $SMALL > gpurun_out/r1_ncu_full.log 2>&1
BIG="python scripts/gpu_zi1000.py 1024 sparse_zi_1000"
$BIG > gpurun_out/r1_big_plain.log 2>&1 && \
ncu --section WarpStateStats --section SchedulerStats --section SourceCounters --section InstructionStats --section MemoryWorkloadAnalysis --section Occupancy --clock-control none --import-source on -k regex:sim_kernel -c 1 -o gpurun_out/r1_zi1000x1024 $BIG > gpurun_out/r1_ncu_big.log 2>&1
ls -la gpurun_out
