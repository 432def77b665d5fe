#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "options_round_trip" 2>&1 | tail -3
for ib in 1024 2048; do timeout 200 python probes/sweep_variants.py --n 16384 --m 16384 --dtypes float16 --cases LLNS --opt inv_block=$ib 2>&1 | sed "s/^/ib=$ib /"; done | tee gpurun_out/sweep_ib_b.txt
for ib in 128 256 512 1024; do timeout 200 python probes/sweep_variants.py --n 16384 --m 16384 --dtypes float32 --cases LLNS --opt inv_block=$ib 2>&1 | sed "s/^/ib=$ib /"; done | tee -a gpurun_out/sweep_ib_b.txt
timeout 200 python probes/sweep_variants.py --n 32768 --m 16384 --dtypes float16 --cases RLNS --opt inv_block=2048 2>&1 | sed "s/^/ib=2048 /" | tee -a gpurun_out/sweep_ib_b.txt
