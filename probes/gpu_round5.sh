#!/bin/bash
# GPU visit 5: block-inverse leaves -- parity of the low-precision tests, then timing sweep over inv_block
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "block_inverse or low_precision or tensor_core" 2>&1 | tail -15 | tee gpurun_out/pytest_gpu5.txt
for ib in 128 512 1024 2048; do
  timeout 300 python probes/sweep_variants.py --n 16384 --m 16384 --dtypes float16,float32 --cases LLNS,RLNS,LUTS --opt inv_block=$ib 2>&1 | sed "s/^/ib=$ib /" | tee -a gpurun_out/sweep_r5.txt
done
timeout 300 python probes/sweep_variants.py --n 32768 --m 16384 --dtypes float16 --cases RLNS --opt inv_block=1024 2>&1 | tee -a gpurun_out/sweep_r5.txt
timeout 300 python probes/sweep_variants.py --n 32768 --m 16384 --dtypes float16 --cases RLNS --opt inv_block=2048 2>&1 | tee -a gpurun_out/sweep_r5.txt
