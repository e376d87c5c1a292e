#!/bin/bash
# debug build of the CUDA library with per-phase cycle counters (scripts/gpu_phase_timing.py)
set -e
cd "$(dirname "$0")/.."
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr -shared -Xcompiler -fPIC \
     -DB200FHE_PHASE_TIMING ${EXTRA_DEFS} -o iyokan_b200/csrc/libb200fhe_timing.so iyokan_b200/csrc/b200fhe.cu -ldl
