#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu20.txt
timeout 600 python probes/sweep_variants.py --n 16384 --m 16384 --dtypes float64 --cases LLNS,RLNS,RUTS,RLNM,RUNM 2>&1 | tee gpurun_out/sweep_r20.txt
timeout 300 python probes/tc_probe.py --child "time:f32:16384:16384:LUTM" 2>&1 | head -2 | cut -c1-300 | tee -a gpurun_out/sweep_r20.txt
