#!/bin/bash
# One GPU call: parity tests, bench, ncu launch list, ncu full capture of the hot kernel.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -2 | tee gpurun_out/bench_ref.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:br_kernel -s 1 -c 1 -o gpurun_out/prof_br \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch 2048 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
