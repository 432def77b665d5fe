#!/bin/bash
mkdir -p gpurun_out
for ck in 512 1024 2048 4096 0; do
  timeout 200 python probes/sweep_variants.py --n 16384 --m 16384 --dtypes float32 --cases LUTM,LLNS --opt tc_chunk_k=$ck 2>&1 | sed "s/^/chunk_k=$ck /"
done | tee gpurun_out/sweep_chunk.txt
