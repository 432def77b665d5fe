#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "wide_pair or options_round" 2>&1 | tail -6 | tee gpurun_out/pytest_wide.txt
for opt in tc_wide_k=4096 tc_wide_k=0 tc_wide_k=2048 tc_wide_k=8192; do
timeout 200 python probes/sweep_variants.py --n 32768 --m 16384 --dtypes float16 --cases RLNS --opt $opt 2>&1 | sed "s/^/$opt /"
timeout 200 python probes/sweep_variants.py --n 16384 --m 16384 --dtypes float16 --cases LLNS --opt $opt 2>&1 | sed "s/^/$opt /"
done | tee gpurun_out/sweep_wide.txt
