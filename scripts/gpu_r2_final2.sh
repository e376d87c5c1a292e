#!/bin/bash
# round-2 final single-GPU call: smoke, whole GPU suite, both bench arms, ncu launch list + DRAM traffic of the bench's launch
set -x
mkdir -p gpurun_out
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2 | tee gpurun_out/r02_smoke.log
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 2>/dev/null | tail -1 | tee gpurun_out/r02_bench_ref.log | cut -c1-400
B200FHE_NO_CALIBRATE=1 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,launch__grid_size --clock-control none \
    -k regex:br7_kernel --csv --log-file gpurun_out/r02_br7_traffic.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-netlist > gpurun_out/r02_ncu_traffic.log 2>&1
python scripts/traffic_from_ncu.py gpurun_out/r02_br7_traffic.csv 8192 gpurun_out/br_kernel_traffic.json
cp gpurun_out/br_kernel_traffic.json profiles/br_kernel_traffic.json
timeout 1500 python bench.py --steps 20 --warmup 5 2> gpurun_out/r02_bench.err | tail -1 | tee gpurun_out/r02_bench.log | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-netlist > gpurun_out/r02_ncu_bench.log 2>&1
ls -la gpurun_out | tail -12
NGPU=1 bash scripts/sweep_params.sh --steps 3 2>/dev/null | cut -c1-200
