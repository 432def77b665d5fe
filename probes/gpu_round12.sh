#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "persistent or cta_pair or block_inverse or low_precision or tensor_core" 2>&1 | tail -12 | tee gpurun_out/pytest_gpu12.txt
for opt in tc_persist=1 tc_persist=0; do
  timeout 300 python probes/tc_probe.py --child "time:f16:16384:16384:LLNS@$opt" 2>&1 | tee -a gpurun_out/tc_time_r12.txt | cut -c1-900
done
timeout 300 python probes/sweep_variants.py --n 32768 --m 16384 --dtypes float16 --cases RLNS 2>&1 | tee -a gpurun_out/sweep_r12.txt
timeout 300 python probes/sweep_variants.py --n 16384 --m 16384 --dtypes float16 --cases LLNM,RUNS,LUTS,RLTM 2>&1 | tee -a gpurun_out/sweep_r12.txt
