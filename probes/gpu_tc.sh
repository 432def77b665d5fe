#!/bin/bash
# GPU visit for the tcgen05 path: isolated variant probe, timings, then the parity suite.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | head -1
timeout 900 python probes/tc_probe.py 2>&1 | tee gpurun_out/tc_probe.txt | cut -c1-400
timeout 600 python probes/tc_probe.py time:f16:16384:16384:LLNS time:f32:16384:16384:LUTM time:f16:8192:32768:RLNS 2>&1 | tee gpurun_out/tc_time.txt | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
