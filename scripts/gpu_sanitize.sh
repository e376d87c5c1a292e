#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 500 compute-sanitizer --tool $tool --print-limit 20 python scripts/gpu_sanitize.py > gpurun_out/r02_sanitize_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|exact|MISMATCH|gates|agree|Error|hazard" gpurun_out/r02_sanitize_$tool.log | sort | uniq -c | head -20
done
