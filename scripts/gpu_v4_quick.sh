#!/bin/bash
# quick loop for the latency kernel: parity test + timings at a few batch sizes
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "variant4" 2>&1 | tail -3
VARIANTS=${VARIANTS:-4:1} SIZES=${SIZES:-1,148,296,2368} timeout 300 python scripts/gpu_latency_table.py 2>&1 | tail -4
