#!/bin/bash
# round-2 GPU call G (8 GPUs): bench with the netlist leg on 8 ranks (iyokan-b200, one process per GPU)
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --netlist-cases cahp-pearl-mux,cahp-ruby-mux,mux-ram-8-16-16 \
    2> gpurun_out/r2g_bench8.err | tail -1 | tee gpurun_out/r2g_bench8.log | cut -c1-300
tail -5 gpurun_out/r2g_bench8.err
