#!/bin/bash
# phase timing of the two rotation kernels + ncu --set full of the bench's own rotation launch (traffic per launch)
set -x
mkdir -p gpurun_out
timeout 300 python scripts/gpu_phase_timing.py 2>&1 | tail -45 | tee gpurun_out/r02_phase_timing.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:br7_kernel -c 1 -o gpurun_out/prof_br7_bench -f \
    python bench.py --steps 1 --warmup 0 --no-netlist --no-cpu-baseline > gpurun_out/ncu_br7_bench.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_br7_bench.ncu-rep > gpurun_out/r02_br7_bench_launch.txt 2>&1 || true
tail -5 gpurun_out/ncu_br7_bench.log
