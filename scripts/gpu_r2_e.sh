#!/bin/bash
# round-2 GPU call E: whole GPU suite (incl. the reference's test0 suite on the plugin) + default bench
set -x
mkdir -p gpurun_out
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2 | tee gpurun_out/r2e_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2e_pytest.log
