#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "block_inverse or low_precision or tensor_core or gated" 2>&1 | tail -5 | tee gpurun_out/pytest_gpu8.txt
for opt in inv_overlap=1 inv_overlap=0; do
  timeout 300 python probes/sweep_variants.py --n 16384 --m 16384 --dtypes float16,float32 --cases LLNS,RUNS --opt $opt 2>&1 | sed "s/^/$opt /" | tee -a gpurun_out/sweep_r8.txt
done
timeout 300 python probes/sweep_variants.py --n 32768 --m 16384 --dtypes float16 --cases RLNS 2>&1 | tee -a gpurun_out/sweep_r8.txt
