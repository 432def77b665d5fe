#!/bin/bash
mkdir -p gpurun_out
timeout 900 python probes/tc_probe.py gemm:f16:NN@tc_cg=2 gemm:f16:TN@tc_cg=2 gemm:f16:NT@tc_cg=2 gemm:f32:NN@tc_cg=2 gemm:f32:NT@tc_cg=2 trx:f16@tc_cg=2 trx:f32@tc_cg=2 2>&1 | tee gpurun_out/tc_probe10.txt | cut -c1-200 | awk 'NR%3==0'
timeout 600 python probes/tc_probe.py time:f16:16384:16384:LLNS time:f32:16384:16384:LLNS time:f16:32768:16384:RLNS 2>&1 | tee gpurun_out/tc_time10.txt | cut -c1-700
