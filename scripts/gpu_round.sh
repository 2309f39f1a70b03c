#!/bin/bash
# one GPU round-trip, as the driver does it at round end: smoke, the whole GPU suite, the bench with the driver's flags
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warning | tail -3
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/round_bench.json 2> gpurun_out/round_bench.err ) 2>&1 | grep real
tail -c 1500 gpurun_out/round_bench.json; tail -2 gpurun_out/round_bench.err
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/round_bench_ref.json 2> gpurun_out/round_bench_ref.err ) 2>&1 | grep real
tail -c 500 gpurun_out/round_bench_ref.json
