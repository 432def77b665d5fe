#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conditioning.py tests/test_gpu_boundary.py -x -q -m gpu -s 2>&1 | tail -70 | tee gpurun_out/r2_pytest_cond.txt
timeout 300 python probes/slab_phases.py 2048 16384 S 2>&1 | tee gpurun_out/r2_slab_phases.txt
timeout 300 python probes/sweep_variants.py --n 16384 --m 16384 --dtypes float16,float32 --cases LLNS,RLNS 2>&1 | tail -8 | tee gpurun_out/r2_lowprec_guard.txt
timeout 300 python probes/sweep_variants.py --n 16384 --m 16384 --dtypes float16,float32 --cases LLNS,RLNS --opt inv_guard=0 2>&1 | tail -8 | tee -a gpurun_out/r2_lowprec_guard.txt
