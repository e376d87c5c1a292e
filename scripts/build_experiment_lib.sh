#!/bin/bash
# Builds iyokan_b200/csrc/libb200fhe_exp.so with the experimental rotation shapes compiled in
# (br8_kernel: 4-CTA clusters, br9_kernel: 4-point threads); load it with B200FHE_LIB=... (scripts/gpu_br8_probe.py,
# scripts/gpu_br9_probe.py).  The shipped library leaves them out: both measured slower than br6_kernel.
set -e
cd "$(dirname "$0")/.."
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr -shared -Xcompiler -fPIC \
     -DB200FHE_WITH_BR8 -DB200FHE_WITH_BR9 ${EXTRA_DEFS} -o iyokan_b200/csrc/libb200fhe_exp.so iyokan_b200/csrc/b200fhe.cu -ldl
