#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_complex.py -x -q -m gpu -k "lauum or complex" 2>&1 | tail -12 | tee gpurun_out/r2_pytest_lauum2.txt
for tool in memcheck racecheck synccheck; do
  q=""; [ $tool != memcheck ] && q="--quick"
  ( time timeout 900 compute-sanitizer --tool $tool --print-limit 20 python probes/sanitize_small.py $q ) > gpurun_out/r2_sanitizer_$tool.txt 2>&1
  tail -6 gpurun_out/r2_sanitizer_$tool.txt
done
