#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4 | tee gpurun_out/r2n_parity.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-netlist --no-cpu-baseline 2> gpurun_out/r2n_bench.err | tail -1 | tee gpurun_out/r2n_bench.log | cut -c1-700
