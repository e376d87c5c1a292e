#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "variant7 or gate" 2>&1 | tail -3 | tee gpurun_out/r2l_parity.log
VARIANTS=7:8 SIZES=1184,2368 timeout 200 python scripts/gpu_latency_table.py 2>&1 | tail -2 | tee gpurun_out/r2l_latency.log
( PT_CASES=7:1184 timeout 300 python scripts/gpu_phase_timing.py
  PT_TAG=_J4 PT_CASES=7:1184 B200FHE_BR7_GROUP=4 timeout 300 python scripts/gpu_phase_timing.py
  PT_TAG=_J2 PT_CASES=7:1184 B200FHE_BR7_GROUP=2 timeout 300 python scripts/gpu_phase_timing.py ) 2>&1 | grep -v "^+" | tee gpurun_out/r2l_phase_timing.log
