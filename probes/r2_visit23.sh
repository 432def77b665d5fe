#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_boundary.py -x -q -m gpu -k "fp64 or slab or unit_diagonal or opposite or right_side or gated or lu or host_buffer or concurrent or streaming" 2>&1 | tail -8 | tee gpurun_out/r2_pytest_slab2b.txt
timeout 300 python probes/slab_phases.py 2048 16384 S 112 2>&1 | tail -6
timeout 300 python probes/time_rectrxm.py --n 16384 --m 16384 --macro -1,4096 --streams 0,1 --reps 3 2>&1 | tee gpurun_out/r2_slab2_pipe.txt
timeout 300 python probes/time_rectrxm.py --n 16384 --m 16384 --uplo U --macro -1 --streams 0 --reps 2 2>&1 | tee -a gpurun_out/r2_slab2_pipe.txt
timeout 300 python probes/time_rectrxm.py --n 32768 --m 8192 --macro -1 --streams 0 --reps 2 2>&1 | tee -a gpurun_out/r2_slab2_pipe.txt
