#!/bin/bash
# round 2, last GPU trip: the lane engine with and without thread-local staging of the hot state (A/B on the `also`
# workload, sparse_zi_1000 x 45 056), then the whole -m gpu suite on the fastest build, then -- if time is left -- the
# DRAM traffic of that build's kernel at the same size (the "after" of profiles/r2_lane_zi1000x45056_metrics.csv)
mkdir -p gpurun_out
L=$PWD/abides_b200/csrc
for v in "" _nostage _nostage1; do
  ABIDES_B200_LIB=$L/libabides_b200$v.so timeout 75 python scripts/gpu_lane_perf.py sparse_zi_1000 45056 lane 2>&1 | grep -v Warning | tail -1 > gpurun_out/r2n_ab_zi1000$v.json
  cut -c1-260 gpurun_out/r2n_ab_zi1000$v.json
done
BEST=$(python - <<'PY'
import json
best, rate = "", 0.0
for v in ("", "_nostage", "_nostage1"):
    try:
        d = json.loads(open("gpurun_out/r2n_ab_zi1000%s.json" % v).read())
        if d["nonzero_status"] == 0 and d["msg_per_s"] > rate * (1.0 if v == "" else 1.03):     # a variant must win by 3 %
            best, rate = v, d["msg_per_s"]
    except Exception:
        pass
print(best)
PY
)
echo "fastest build: libabides_b200$BEST.so" | tee gpurun_out/r2n_best.txt
( time ABIDES_B200_LIB=$L/libabides_b200$BEST.so timeout 345 python -m pytest tests -m gpu -q ) > gpurun_out/r2n_gpu_tests.log 2>&1; tail -5 gpurun_out/r2n_gpu_tests.log
if [ $SECONDS -lt 520 ]; then
  M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum,smsp__inst_executed.sum
  ABIDES_B200_LIB=$L/libabides_b200$BEST.so timeout $((625 - SECONDS)) ncu --metrics $M --clock-control none -k regex:lane_sim_kernel -c 1 --csv \
      --log-file gpurun_out/r2n_lane_zi1000x45056_metrics.csv python scripts/gpu_lane_perf.py sparse_zi_1000 45056 lane > gpurun_out/r2n_ncu_lane.log 2>&1
  tail -5 gpurun_out/r2n_lane_zi1000x45056_metrics.csv | cut -d, -f13- | cut -c1-100
fi
echo "elapsed $SECONDS s"
