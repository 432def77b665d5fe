#!/bin/bash
mkdir -p gpurun_out
timeout 600 python probes/tc_probe.py time:f16:16384:16384:LLNS time:f16:16384:16384:LLNS@pdl=0 time:f32:16384:16384:LLNS time:f16:8192:16384:RLNS time:f16:16384:16384:LLNM 2>&1 | grep -v breakdown | tee gpurun_out/tc_time6.txt | cut -c1-400
timeout 900 python probes/tc_probe.py trx:f16 trx:f32 gemm:f16:NN gemm:f32:NT 2>&1 | tee gpurun_out/tc_probe6.txt | cut -c1-200 | tail -12
timeout 600 python probes/tc_determinism.py float16 LLNS 8192 8192 | cut -c1-300
timeout 600 python probes/tc_determinism.py float32 RLNS 4096 8192 | cut -c1-300
