#!/bin/bash
set -x
mkdir -p gpurun_out
B200FHE_L2_PERSIST=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-netlist --no-cpu-baseline 2> gpurun_out/r2v_bench.err | tail -1 > gpurun_out/r2v_bench.log
head -c 300 gpurun_out/r2v_bench.log; grep "L2 persisting" gpurun_out/r2v_bench.err
B200FHE_L2_PERSIST=1 B200FHE_NO_CALIBRATE=1 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,launch__grid_size --clock-control none \
    -k regex:br7_kernel --csv --log-file gpurun_out/r2v_traffic.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-netlist > gpurun_out/r2v_ncu.log 2>&1
python scripts/traffic_from_ncu.py gpurun_out/r2v_traffic.csv 8192 gpurun_out/r2v_traffic.json
