#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python scripts/gpu_br7_groups.py 2>&1 | tail -12 | tee gpurun_out/r2b_groups.log
VARIANTS=7:8,3:6,3:4,4:1,6:1 SIZES=30,74,148,592,888,1184,2368 timeout 300 python scripts/gpu_latency_table.py 2>&1 | tail -8 | tee gpurun_out/r2b_latency.log
