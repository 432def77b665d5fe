#!/bin/bash
mkdir -p gpurun_out
timeout 900 python probes/tc_probe.py gemm:f32:NN gemm:f32:TN gemm:f32:NT@tc_bn=128 gemm:f16:NN trx:f16 trx:f32 trx:f32@tc_chunk_k=128 2>&1 | tee gpurun_out/tc_probe4.txt | cut -c1-300
timeout 600 python probes/tc_probe.py time:f32:16384:16384:LUTM time:f32:16384:16384:LUTM@tc_chunk_k=256 time:f32:16384:16384:LUTM@tc_chunk_k=0 time:f16:16384:16384:LLNS 2>&1 | tee gpurun_out/tc_time4.txt | cut -c1-1500
timeout 900 python -m pytest tests -m gpu -q -k "gemm_add_sub or low_precision or tensor_core" 2>&1 | tail -8 | tee gpurun_out/pytest_gpu4.txt
