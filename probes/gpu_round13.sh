#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "persistent or block_inverse or tensor_core_path" 2>&1 | tail -6 | tee gpurun_out/pytest_gpu13.txt
timeout 300 python probes/tc_probe.py --child "time:f16:16384:16384:LLNS" 2>&1 | tee -a gpurun_out/tc_time_r13.txt | cut -c1-900
timeout 300 python probes/tc3_phases.py 2>&1 | head -1 | cut -c1-1200
timeout 300 python probes/gemm_sweep.py 16384 | tee gpurun_out/gemm_sweep_f16_b.txt
timeout 300 python probes/sweep_variants.py --n 32768 --m 16384 --dtypes float16 --cases RLNS 2>&1 | tee -a gpurun_out/sweep_r13.txt
