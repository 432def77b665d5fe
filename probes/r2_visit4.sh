#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conditioning.py tests/test_gpu_boundary.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2_pytest_cond2.txt
timeout 300 python probes/slab_phases.py 2048 16384 S 2>&1 | tail -22 | tee gpurun_out/r2_slab_phases2.txt
timeout 300 python probes/sweep_variants.py --n 16384 --m 16384 --dtypes float16 --cases LLNS,RLNS 2>&1 | tail -8 | tee gpurun_out/r2_lowprec_guard2.txt
timeout 300 python probes/sweep_variants.py --n 16384 --m 16384 --dtypes float16 --cases LLNS,RLNS --opt inv_guard=0 2>&1 | tail -8 | tee -a gpurun_out/r2_lowprec_guard2.txt
