#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fp64 or slab or unit_diagonal or opposite or right_side or gated or lu or host_buffer or concurrent" 2>&1 | tail -15 | tee gpurun_out/r2_pytest_slab.txt
timeout 300 python probes/slab_phases.py 2048 16384 S 2>&1 | head -20 | tee gpurun_out/r2_slab_phases3.txt
timeout 300 python probes/time_rectrxm.py --n 16384 --m 16384 --macro 2048 --streams 1,4,8 --reps 3 2>&1 | tee gpurun_out/r2_macro_sweep2.txt
