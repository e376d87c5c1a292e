#!/bin/bash
# bench (HomNAND shards + netlist leg) and the config-5 sweep on $NGPU ranks
N=${NGPU:-2}
set -x
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + N)) \
    bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --netlist-cases cahp-pearl-mux,cahp-ruby-mux,mux-ram-8-16-16 \
    2> gpurun_out/r02_bench_n$N.err | tail -1 > gpurun_out/r02_bench_n$N.log
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_bench_n$N.log").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["e2e"]["value"], d["outputs_ok"])
for x in [d["netlist"]] + d.get("netlist_more", []):
    print({k: x[k] for k in ("case", "n_gpus", "s_per_cycle", "bootstraps_per_s", "collectives_per_cycle", "exchanged_bytes_per_cycle", "model_s_per_cycle", "outputs_ok")})
PY
NGPU=$N bash scripts/sweep_params.sh --steps 3 --no-cpu 2>&1 | cut -c1-330
