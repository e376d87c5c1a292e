#!/bin/bash
# round-2 GPU call F: 80-bit flavour parity on the GPU + parameter sweep at N = 1
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_flavour80.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2f_pytest80.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 | tee gpurun_out/r2f_pytest128.log
NGPU=1 bash scripts/sweep_params.sh --steps 3 2>&1 | tail -6 | cut -c1-900
