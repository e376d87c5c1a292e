#!/bin/bash
set -x
mkdir -p gpurun_out
B200FHE_NO_CALIBRATE=1 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,launch__grid_size --clock-control none \
    -k regex:br7_kernel --csv --log-file gpurun_out/r02_br7_traffic.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-netlist > gpurun_out/r02_ncu_traffic.log 2>&1
python scripts/traffic_from_ncu.py gpurun_out/r02_br7_traffic.csv 8192 gpurun_out/br_kernel_traffic.json
