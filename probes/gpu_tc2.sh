#!/bin/bash
mkdir -p gpurun_out
timeout 600 python probes/tc_probe.py gemm:f32:NN gemm:f32:NT trx:f32 2>&1 | tee gpurun_out/tc_probe2.txt | cut -c1-600
timeout 600 python probes/tc_probe.py time:f16:16384:16384:LLNS time:f32:16384:16384:LUTM time:f32:16384:16384:LLNS 2>&1 | tee gpurun_out/tc_time2.txt | cut -c1-1500
