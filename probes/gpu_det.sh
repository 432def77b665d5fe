#!/bin/bash
mkdir -p gpurun_out
for args in "float16 LLNS 8192 8192" "float16 LLNS 1024 512" "float16 LLNS 256 256" "float16 LLNS 128 256" "float16 RLNS 8192 8192" "float32 LLNS 8192 8192" "float16 LLNS 8192 8192 tc_bn=256" "float16 LLNM 8192 8192"; do
  timeout 300 python probes/tc_determinism.py $args 2>&1 | tail -1 | cut -c1-700
done | tee gpurun_out/tc_det.txt
