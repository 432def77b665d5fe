#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  q=""; [ $tool != memcheck ] && q="--quick"
  ( time timeout 900 compute-sanitizer --tool $tool --print-limit 20 python probes/sanitize_small.py $q ) > gpurun_out/r2_sanitizer_$tool.txt 2>&1
  tail -5 gpurun_out/r2_sanitizer_$tool.txt
done
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2_pytest_gpu_b.txt
