#!/bin/bash
set -x
mkdir -p gpurun_out
( timeout 300 python scripts/gpu_phase_timing.py
  PT_TAG=_J4 PT_CASES=7:1184 B200FHE_BR7_GROUP=4 timeout 300 python scripts/gpu_phase_timing.py
  PT_TAG=_J2 PT_CASES=7:1184 B200FHE_BR7_GROUP=2 timeout 300 python scripts/gpu_phase_timing.py ) 2>&1 | grep -v "^+" | tee gpurun_out/r02_phase_timing.log
