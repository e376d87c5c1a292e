#!/bin/bash
# round-2 GPU call A: parity of the 16-warp shape, its wave time vs the 12-warp shape, bench, ncu full capture
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5 | tee gpurun_out/r2a_pytest.log
VARIANTS=7:8,3:6,3:4,4:1,6:1 SIZES=30,74,148,592,888,1184,2368 timeout 300 python scripts/gpu_latency_table.py 2>&1 | tail -8 | tee gpurun_out/r2a_latency.log
cp gpurun_out/latency_table.json gpurun_out/r2a_latency_table.json
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2a_bench_auto.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --variant 3 --jobs-per-cta 6 2>&1 | tail -1 | tee gpurun_out/r2a_bench_v3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:br7_kernel -s 1 -c 1 -f -o gpurun_out/prof_br7 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch 1184 --variant 7 --jobs-per-cta 8 > gpurun_out/ncu_br7.log 2>&1
ls -la gpurun_out | tail -8
