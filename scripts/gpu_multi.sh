#!/bin/bash
# multi-GPU checks (run under gpurun --gpus N)
N=${1:-2}
CYC=${CYCLES:-3}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_n$N.log
for c in ${CASES:-mux-ram-8-16-16 cahp-pearl-mux cahp-ruby-mux}; do
  timeout 600 $TR scripts/multi_gpu_netlist.py --case $c --cycles $CYC 2>&1 | tail -1 | tee gpurun_out/net_${c}_n$N.log
done
