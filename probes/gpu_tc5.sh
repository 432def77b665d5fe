#!/bin/bash
mkdir -p gpurun_out
timeout 900 python probes/tc_probe.py trx:f16 trx:f32 2>&1 | tee gpurun_out/tc_probe5.txt | cut -c1-300
timeout 600 python probes/tc_probe.py time:f32:16384:16384:LUTM time:f16:16384:16384:LLNM time:f16:8192:16384:RUTM time:f32:8192:16384:RLNM 2>&1 | tee gpurun_out/tc_time5.txt | cut -c1-900
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu5.txt
