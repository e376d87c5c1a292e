#!/bin/bash
# One GPU call: parity tests, bench (both arms), netlist runs, ncu launch list, ncu full capture of the hot kernel.
set -x
mkdir -p gpurun_out
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.log
for c in mux-ram-8-16-16 cahp-pearl-mux cahp-ruby-mux; do
  timeout 300 python scripts/multi_gpu_netlist.py --case $c --cycles ${CYCLES:-10} 2>&1 | tail -1 | tee gpurun_out/net_${c}_n1.log
done
timeout 300 python scripts/gpu_latency_table.py 2>&1 | tail -6
./scripts/microbench/pipes > gpurun_out/microbench_pipes.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
# dominant launch of the default step: br3_kernel<6>, 7992 jobs
timeout 900 ncu --set full --clock-control none --import-source on -k regex:br3_kernel -s 1 -c 1 -f -o gpurun_out/prof_br_auto \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
# latency shapes at the sizes they serve
VARIANT=4 NB=148 GLIST=1 bash scripts/gpu_ncu.sh > /dev/null 2>&1
VARIANT=6 NB=74 GLIST=1 bash scripts/gpu_ncu.sh > /dev/null 2>&1
ls -la gpurun_out
