#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fp64 or slab or unit_diagonal or opposite or right_side or gated or lu or host_buffer or concurrent" 2>&1 | tail -6 | tee gpurun_out/r2_pytest_slab64.txt
for w in 64 128; do
timeout 300 python probes/time_rectrxm.py --n 16384 --m 16384 --macro 2048 --streams 1,4 --reps 3 --opt slab_w=$w 2>&1 | tee -a gpurun_out/r2_slabw_sweep.txt
done
timeout 300 python probes/time_rectrxm.py --n 16384 --m 16384 --macro 1024,4096 --streams 4 --reps 3 --opt slab_w=64 2>&1 | tee -a gpurun_out/r2_slabw_sweep.txt
timeout 300 python probes/time_rectrxm.py --n 2048 --m 16384 --macro 2048 --streams 1 --reps 3 --opt slab_w=64 2>&1 | tee -a gpurun_out/r2_slabw_sweep.txt
timeout 300 python probes/time_rectrxm.py --n 2048 --m 16384 --macro 2048 --streams 1 --reps 3 --opt slab_w=128 2>&1 | tee -a gpurun_out/r2_slabw_sweep.txt
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python probes/sanitize_small.py --quick 2>&1 | grep -v "Host Frame\|libcuda\|python3\|libc.so\|libffi\|_ctypes" | tail -8 | tee gpurun_out/r2_sanitizer_racecheck.txt
