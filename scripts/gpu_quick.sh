#!/bin/bash
# quick perf check: throughput of both ZI workloads + i-cache counters of sim_kernel
mkdir -p gpurun_out
timeout 300 python scripts/gpu_zi1000.py 1024 sparse_zi_1000 2>&1 | grep -E "kernel ms|msg/s|mismatch" | tr '\n' ' '; echo
timeout 120 python scripts/gpu_zi1000.py 1024 sparse_zi_100 2>&1 | grep -E "kernel ms|msg/s|mismatch" | tr '\n' ' '; echo
timeout 600 ncu --metrics sm__icc_request_hit_rate.pct,sm__icc_requests.sum,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio \
  --clock-control none -k regex:sim_kernel -c 1 python scripts/gpu_zi1000.py 1024 sparse_zi_100 2>&1 | grep -E "icc|gcc|inst_executed|time_duration|no_instruction"
