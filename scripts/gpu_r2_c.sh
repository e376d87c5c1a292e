#!/bin/bash
# round-2 GPU call C: compiled schedule (CUDA graph replay) correctness + bench with the netlist legs
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_netlist.py -x -q 2>&1 | tail -8 | tee gpurun_out/r2c_pytest.log
timeout 1200 python bench.py --steps 5 --warmup 3 2> gpurun_out/r2c_bench.err | tail -1 | tee gpurun_out/r2c_bench.log
tail -5 gpurun_out/r2c_bench.err
