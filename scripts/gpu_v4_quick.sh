#!/bin/bash
# quick loop for the latency kernels: parity tests + timings at a few batch sizes
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "${TESTS:-variant4 or variant5}" 2>&1 | tail -3
VARIANTS=${VARIANTS:-4:1,5:1} SIZES=${SIZES:-1,37,74,148,296} timeout 300 python scripts/gpu_latency_table.py 2>&1 | tail -4
