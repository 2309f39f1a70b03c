#!/bin/bash
# round 2, two GPUs: the bench under torchrun (weak and strong scaling) and the box-wide sweep driver
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2k_bench_n2.json 2> gpurun_out/r2k_bench_n2.err; tail -c 900 gpurun_out/r2k_bench_n2.json; tail -3 gpurun_out/r2k_bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --gpus 2 --steps 2 --warmup 1 --scaling strong --instances 4096 > gpurun_out/r2k_bench_n2_strong.json 2> gpurun_out/r2k_bench_n2_strong.err; tail -c 600 gpurun_out/r2k_bench_n2_strong.json
timeout 600 python -m pytest tests/test_gpu_lane.py -x -q -k "chunked" 2>&1 | tail -2
timeout 600 python - <<'PY' 2>&1 | grep -v Warning | tail -3
import sys, time; sys.path.insert(0, ".")
from abides_b200 import runtime
seeds = runtime.parallel_seeds(5, 4736)
t = time.time(); out = runtime.run_sweep("rmsc03", seeds, ["-t", "ABM", "-d", "20200603"], devices=(0, 1), chunk_instances=1184)
print("run_sweep rmsc03 x 4736 on 2 GPUs: %.1f s wall, %d messages, placement %s" % (time.time() - t, int(out.res["ttl_messages"].sum()), [(p[0], p[2], round(p[3])) for p in out.placement]))
PY
