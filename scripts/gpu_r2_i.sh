#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 180 python scripts/gpu_br8_probe.py 2>&1 | tail -6 | tee gpurun_out/r2i_br8.log
timeout 120 oracle/_ref/cufhe_test_gate_gpu 2>&1 | tail -40 | tee gpurun_out/r2i_cufhe.log
