#!/bin/bash
# ncu full captures of the blind-rotation kernel for a few jobs-per-CTA settings
mkdir -p gpurun_out
for g in ${GLIST:-4 2}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:br[23456]?_kernel -s 1 -c 1 -f -o gpurun_out/prof_br_v${VARIANT:-1}_G$g \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch ${NB:-1184} --jobs-per-cta $g --variant ${VARIANT:-1} > gpurun_out/ncu_full_v${VARIANT:-1}_G$g.log 2>&1
done
ls -la gpurun_out
