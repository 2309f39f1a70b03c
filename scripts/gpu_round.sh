#!/bin/bash
# one GPU round-trip: parity tests, then throughput checks and the phase profile (each under its own timeout)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python scripts/gpu_zi1000.py 1024 sparse_zi_1000 2>&1 | tail -12
timeout 120 python scripts/gpu_zi1000.py 1024 sparse_zi_100 2>&1 | grep -E "run s|msg/s|mismatch"
timeout 300 python scripts/gpu_prof.py 1024 sparse_zi_1000 2>&1 | tail -22
