#!/bin/bash
# ncu full captures of the blind-rotation kernel for a few jobs-per-CTA settings
mkdir -p gpurun_out
for g in ${GLIST:-4 2}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:br_kernel -s 1 -c 1 -f -o gpurun_out/prof_br_G$g \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch ${NB:-1184} --jobs-per-cta $g > gpurun_out/ncu_full_G$g.log 2>&1
done
ls -la gpurun_out
