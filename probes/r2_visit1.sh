#!/bin/bash
# round 2, visit 1: baseline data for the fused slab kernel (ncu full capture, macro / streams sweep)
mkdir -p gpurun_out
python -c "import torch; print(torch.cuda.get_device_name(0))"
timeout 600 python probes/time_rectrxm.py --n 16384 --m 16384 --macro 512,1024,2048,4096 --streams 1,4 --reps 2 2>&1 | tee gpurun_out/r2_macro_sweep.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:slab_f64 -s 1 -c 1 -o gpurun_out/r2_prof_slab -f \
    python probes/time_rectrxm.py --n 2048 --m 16384 --macro 2048 --reps 1 > gpurun_out/r2_ncu_slab.log 2>&1
tail -3 gpurun_out/r2_ncu_slab.log
ls -la gpurun_out/r2_prof_slab.ncu-rep
python probes/ncu_summarise.py rep gpurun_out/r2_prof_slab.ncu-rep gpurun_out/r2_ncu_slab_summary.csv
cat gpurun_out/r2_ncu_slab_summary.csv
