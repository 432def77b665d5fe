#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fp64 or slab or unit_diagonal or opposite or right_side or gated or lu or host_buffer or concurrent" 2>&1 | tail -12 | tee gpurun_out/r2_pytest_slab2.txt
for o in slab_kind=0 slab_kind=1; do
timeout 300 python probes/time_rectrxm.py --n 2048 --m 16384 --macro 2048 --streams 1 --reps 3 --opt $o 2>&1 | tee -a gpurun_out/r2_slab2_sweep.txt
timeout 300 python probes/time_rectrxm.py --n 16384 --m 16384 --macro 2048,4096 --streams 1,4 --reps 3 --opt $o 2>&1 | tee -a gpurun_out/r2_slab2_sweep.txt
done
timeout 300 python probes/time_rectrxm.py --n 32768 --m 8192 --func M --macro 2048,4096 --streams 0 --reps 2 2>&1 | tee -a gpurun_out/r2_slab2_sweep.txt
