#!/bin/bash
mkdir -p gpurun_out
C="time:f16:16384:16384:LLNS@inv_overlap=0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_f16_ib.csv \
    python probes/tc_probe.py --child "$C" > gpurun_out/ncu_l9.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 0 -c 2 -o gpurun_out/prof_f16_mid -f \
    python probes/tc_probe.py --child "$C" > gpurun_out/ncu_f9a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 1 -c 1 -o gpurun_out/prof_f16_leaf -f \
    python probes/tc_probe.py --child "$C" > gpurun_out/ncu_f9b.log 2>&1
ls -la gpurun_out | tail -5
