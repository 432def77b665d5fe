#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:slab_f64 -s 1 -c 1 -o gpurun_out/prof_slab -f \
    python probes/time_rectrxm.py --n 4096 --m 16384 --macro 4096 --reps 1 > gpurun_out/ncu_slab.log 2>&1
tail -3 gpurun_out/ncu_slab.log
