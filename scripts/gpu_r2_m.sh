#!/bin/bash
set -x
mkdir -p gpurun_out
( PT_CASES=7:1184 timeout 300 python scripts/gpu_phase_timing.py
  PT_TAG=_J4 PT_CASES=7:1184 B200FHE_BR7_GROUP=4 timeout 300 python scripts/gpu_phase_timing.py ) 2>&1 | grep -v "^+" | grep "CTA\|loop of\|ms per launch" | tee gpurun_out/r2m_phase_timing.log
