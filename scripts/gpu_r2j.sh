#!/bin/bash
# round 2, GPU trip: confirm the restored lane code, pick the warp-engine bench size, ncu evidence for the bench kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lane.py -x -q -k "golden or sweep" 2>&1 | tail -2
timeout 900 python scripts/gpu_lane_perf.py sparse_zi_1000 45056 lane 2>&1 | grep -v Warning | tail -1 | cut -c1-330
timeout 900 python scripts/gpu_lane_perf.py rmsc03_2h 2368 warp 2>&1 | grep -v Warning | tail -1 | cut -c1-330
# launch list of the bench command (every kernel with its device time; cold-cache, serialised: compare shares)
timeout 900 python bench.py --steps 2 --warmup 1 --also '' > gpurun_out/r2j_bench_plain.json 2> gpurun_out/r2j_bench_plain.err && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2j_launches.csv \
    python bench.py --steps 2 --warmup 1 --also '' > gpurun_out/r2j_ncu_launches.log 2>&1
tail -c 400 gpurun_out/r2j_bench_plain.json; tail -2 gpurun_out/r2j_ncu_launches.log
# the dominant kernel of the headline workload at a size ncu can replay: rmsc03 (2 h) x 296 on the warp engine
timeout 600 python scripts/gpu_lane_perf.py rmsc03_2h 296 warp > gpurun_out/r2j_plain_warp.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:sim_kernel -c 1 -o gpurun_out/r2j_warp_rmsc03x296 \
    python scripts/gpu_lane_perf.py rmsc03_2h 296 warp > gpurun_out/r2j_ncu_warp.log 2>&1
tail -1 gpurun_out/r2j_plain_warp.log | cut -c1-300; tail -2 gpurun_out/r2j_ncu_warp.log
# the lane kernel on the `also` workload: DRAM traffic and lane occupancy (a metric subset: the launch takes 10 s)
timeout 600 python scripts/gpu_lane_perf.py sparse_zi_1000 8192 lane > gpurun_out/r2j_plain_lane.log 2>&1 && \
timeout 1500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum,sm__inst_executed.avg.per_cycle_elapsed,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,launch__registers_per_thread \
    --clock-control none -k regex:lane_sim_kernel -c 1 --csv --log-file gpurun_out/r2j_lane_zi1000x8192_metrics.csv \
    python scripts/gpu_lane_perf.py sparse_zi_1000 8192 lane > gpurun_out/r2j_ncu_lane.log 2>&1
tail -1 gpurun_out/r2j_plain_lane.log | cut -c1-300; tail -12 gpurun_out/r2j_lane_zi1000x8192_metrics.csv | cut -c1-200
