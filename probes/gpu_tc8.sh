#!/bin/bash
mkdir -p gpurun_out
timeout 900 python probes/tc_probe.py gemm:f32:NN gemm:f32:TN gemm:f32:NT gemm:f32:NT@tc_bn=128 trx:f32 2>&1 | tee gpurun_out/tc_probe8.txt | cut -c1-200 | awk 'NR%2==0'
timeout 600 python probes/tc_probe.py time:f32:16384:16384:LLNS time:f32:16384:16384:LUTM time:f32:8192:16384:RLNS 2>&1 | tee gpurun_out/tc_time8.txt | cut -c1-1300
timeout 900 python -m pytest tests -m gpu -q -k "float32 or gemm or low_precision or tensor_core" 2>&1 | tail -6 | tee gpurun_out/pytest_gpu8.txt
