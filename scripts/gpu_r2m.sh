#!/bin/bash
# round 2, GPU trip: the -m gpu suite with the log-archive tests, the bench line of this build, BASELINE configs[4]
# (16 384-instance rmsc03 grid in one batch: AUTO takes the lane engine, the warp engine's tables would not fit), and
# the DRAM traffic of the two bench kernels at their bench sizes (a metric subset: one or two replays of a 7 s / 15 s launch)
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2m_gpu_tests.log 2>&1; tail -4 gpurun_out/r2m_gpu_tests.log
timeout 400 python bench.py --steps 3 --warmup 3 --also '' > gpurun_out/r2m_bench_default.json 2> gpurun_out/r2m_bench_default.err; tail -c 700 gpurun_out/r2m_bench_default.json; tail -2 gpurun_out/r2m_bench_default.err
timeout 500 python bench.py --workload rmsc03_grid --instances 16384 --steps 1 --warmup 1 --e2e-steps 1 --also '' \
    > gpurun_out/r2m_bench_grid_x16384.json 2> gpurun_out/r2m_bench_grid_x16384.err; tail -c 1200 gpurun_out/r2m_bench_grid_x16384.json; tail -3 gpurun_out/r2m_bench_grid_x16384.err
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed.avg.per_cycle_elapsed,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,lts__t_bytes.sum,l1tex__t_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active,gcc__cache_requests_type_instruction.sum,launch__registers_per_thread
timeout 240 ncu --metrics $M --clock-control none -k regex:sim_kernel -c 1 --csv --log-file gpurun_out/r2m_warp_rmsc03x2368_metrics.csv \
    python scripts/gpu_lane_perf.py rmsc03_2h 2368 warp > gpurun_out/r2m_ncu_warp.log 2>&1; tail -14 gpurun_out/r2m_warp_rmsc03x2368_metrics.csv | cut -d, -f13- | cut -c1-120
timeout 300 ncu --metrics $M --clock-control none -k regex:lane_sim_kernel -c 1 --csv --log-file gpurun_out/r2m_lane_zi1000x45056_metrics.csv \
    python scripts/gpu_lane_perf.py sparse_zi_1000 45056 lane > gpurun_out/r2m_ncu_lane.log 2>&1; tail -14 gpurun_out/r2m_lane_zi1000x45056_metrics.csv | cut -d, -f13- | cut -c1-120
