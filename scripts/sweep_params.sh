#!/bin/bash
# config 5: both parameter flavours, Hom{NAND,MUX}, on ${NGPU:-1} GPU(s); JSON lines -> gpurun_out/sweep_n${NGPU:-1}.jsonl
N=${NGPU:-1}
mkdir -p gpurun_out
: > gpurun_out/sweep_n$N.jsonl
for fl in "" 80; do
  if [ "$N" = 1 ]; then
    B200FHE_FLAVOUR=$fl timeout 900 python scripts/sweep_params.py "$@" | grep '^{' >> gpurun_out/sweep_n$N.jsonl
  else
    B200FHE_FLAVOUR=$fl timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
        --master-port $((29600 + N)) scripts/sweep_params.py "$@" 2>/dev/null | grep '^{' >> gpurun_out/sweep_n$N.jsonl
  fi
done
cat gpurun_out/sweep_n$N.jsonl
