#!/bin/bash
mkdir -p gpurun_out
python probes/tc_phases.py f16 2>&1 | tail -3 | cut -c1-100,330-560
python probes/tc_phases.py f32 2>&1 | tail -3 | cut -c1-100,330-560
timeout 900 python probes/tc_probe.py gemm:f16:NN gemm:f16:NT@tc_bn=128 gemm:f32:TN gemm:f32:NT trx:f16 trx:f32 2>&1 | tee gpurun_out/tc_probe7.txt | cut -c1-200 | awk 'NR%3==0'
timeout 600 python probes/tc_probe.py time:f16:16384:16384:LLNS time:f32:16384:16384:LLNS time:f16:8192:16384:RLNS time:f16:16384:16384:LLNM time:f32:16384:16384:LUTM 2>&1 | tee gpurun_out/tc_time7.txt | cut -c1-1300
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu7.txt
