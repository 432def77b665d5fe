#!/bin/bash
mkdir -p gpurun_out
timeout 900 python probes/tc_probe.py gemm:f16:NN@tc_bn=128 gemm:f16:TN@tc_bn=128 gemm:f16:NT@tc_bn=128 gemm:f32:NN@tc_bn=128 gemm:f32:TN@tc_bn=128 gemm:f32:NT@tc_bn=128 trx:f16 trx:f32 trx:f16@tc_bn=128 trx:f32@tc_bn=128 trx:f32@tf32_raw_hi=1 gemm:f32:NT@tf32_raw_hi=1 2>&1 | tee gpurun_out/tc_probe3.txt | cut -c1-400
timeout 600 python probes/tc_probe.py time:f16:16384:16384:LLNS time:f32:16384:16384:LUTM time:f32:16384:16384:LUTM@tf32_raw_hi=1 time:f16:8192:16384:RLNS 2>&1 | tee gpurun_out/tc_time3.txt | cut -c1-1500
