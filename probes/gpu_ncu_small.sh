#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 0 -c 4 -o gpurun_out/prof_tc_small -f \
    python probes/tc_probe.py --child time:f16:16384:16384:LLNS > gpurun_out/ncu_small.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 0 -c 2 -o gpurun_out/prof_tc_trmm_f16 -f \
    python probes/tc_probe.py --child time:f16:16384:16384:LLNM > gpurun_out/ncu_trmm.log 2>&1
ls -la gpurun_out/*.ncu-rep
