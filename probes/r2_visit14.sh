#!/bin/bash
mkdir -p gpurun_out
for w in 128 64 0; do
timeout 300 python probes/time_rectrxm.py --n 32768 --m 8192 --func M --macro 2048 --streams 0 --reps 2 --opt slab_w=$w 2>&1 | tee -a gpurun_out/r2_c5_slabw.txt
timeout 300 python probes/time_rectrxm.py --n 32768 --m 8192 --func S --macro 2048 --streams 0 --reps 2 --opt slab_w=$w 2>&1 | tee -a gpurun_out/r2_c5_slabw.txt
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -x -q -m gpu -k "fp64 or slab or C5 or C2" 2>&1 | tail -4
