#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 | tee gpurun_out/r2y_parity.log
VARIANTS=6:1,4:1 SIZES=1,30,74,148 timeout 200 python scripts/gpu_latency_table.py 2>&1 | tail -3 | tee gpurun_out/r2y_latency.log
