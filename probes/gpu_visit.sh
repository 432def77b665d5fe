#!/bin/bash
# final single-GPU visit of the round: full parity suite, headline bench, Float16 breakdown, ncu captures of the wide pair kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_final.txt
timeout 900 python bench.py 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-300
timeout 300 python probes/tc_probe.py --child "time:f16:16384:16384:LLNS" 2>&1 | tee gpurun_out/tc_time_final.txt | cut -c1-900
timeout 300 python probes/tc_probe.py --child "time:f16:32768:16384:RLNS" 2>&1 | tee -a gpurun_out/tc_time_final.txt | cut -c1-900
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc4 -s 0 -c 1 -o gpurun_out/prof_tc4 -f \
    python probes/tc_probe.py --child "time:f16:16384:16384:LLNS@inv_overlap=0" > gpurun_out/ncu_tc4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_f16_final.csv \
    python probes/tc_probe.py --child "time:f16:16384:16384:LLNS@inv_overlap=0" > gpurun_out/ncu_l_final.log 2>&1
ls -la gpurun_out | tail -3
