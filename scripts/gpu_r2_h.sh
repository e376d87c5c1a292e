#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 180 python scripts/gpu_br8_probe.py 2>&1 | tail -6 | tee gpurun_out/r2h_br8.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --netlist-cases cahp-pearl-mux,mux-ram-8-16-16 2> gpurun_out/r2h_bench2.err | tail -1 > gpurun_out/r2h_bench2.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2h_bench2.log").read().strip().splitlines()[-1])
for x in [d["netlist"]] + d.get("netlist_more", []):
    print({k: x[k] for k in ("case", "n_gpus", "s_per_cycle", "collectives_per_cycle", "model_s_per_cycle", "outputs_ok")})
PY
