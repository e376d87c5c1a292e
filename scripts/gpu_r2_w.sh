#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ref_link.py -x -q 2>&1 | tail -15 | tee gpurun_out/r2w_reflink.log
