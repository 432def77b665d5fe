#!/bin/bash
mkdir -p gpurun_out
timeout 900 python probes/tc_probe.py gemm:f16:NN@tc_cg=2 gemm:f16:TN@tc_cg=2 gemm:f16:NT@tc_cg=2 gemm:f32:NN@tc_cg=2 gemm:f32:TN@tc_cg=2 gemm:f32:NT@tc_cg=2 2>&1 | tee gpurun_out/tc_probe9.txt | cut -c1-330
timeout 600 python probes/tc_probe.py trx:f16@tc_cg=2 trx:f32@tc_cg=2 2>&1 | tee -a gpurun_out/tc_probe9.txt | cut -c1-300
timeout 600 python probes/tc_probe.py time:f16:16384:16384:LLNS time:f16:16384:16384:LLNS@tc_cg=1 time:f32:16384:16384:LLNS time:f16:32768:16384:RLNS 2>&1 | tee gpurun_out/tc_time9.txt | cut -c1-1200
