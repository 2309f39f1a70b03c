#!/bin/bash
# Round profiling recipe (run under gpurun): plain bench first, then the ncu launch list of the same command.
TAG=${1:-r1}
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --e2e-steps 1"
timeout 900 $BENCH > gpurun_out/${TAG}_bench_plain.json 2> gpurun_out/${TAG}_bench_plain.err || exit 1
cat gpurun_out/${TAG}_bench_plain.json
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH \
  > gpurun_out/${TAG}_ncu_launches.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_launches.log
