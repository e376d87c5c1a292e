#!/bin/bash
# round-2 GPU call D (2 GPUs): NCCL all-gathers inside the per-clock CUDA graph, iyokan-b200 one process per GPU
set -x
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2d_bench2.err | tail -1 | tee gpurun_out/r2d_bench2.log
tail -15 gpurun_out/r2d_bench2.err
B200FHE_NETLIST_HOST=python timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2d_bench2py.err | tail -1 | tee gpurun_out/r2d_bench2py.log
tail -5 gpurun_out/r2d_bench2py.err
timeout 600 python -m pytest tests/test_gpu_netlist.py -x -q -k two_gpus 2>&1 | tail -5 | tee gpurun_out/r2d_pytest2.log
