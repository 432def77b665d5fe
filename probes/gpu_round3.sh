#!/bin/bash
# GPU visit: parity suite, headline bench, timings of the other BASELINE configs, ncu launch lists + full captures of the tcgen05 kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.txt
timeout 600 python probes/tc_probe.py time:f16:16384:16384:LLNS time:f32:16384:16384:LUTM time:f16:32768:16384:RLNS time:f32:16384:16384:LLNS 2>&1 | tee gpurun_out/tc_time_r3.txt | cut -c1-1200
timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_tc|diag_prep" -c 300 --csv --log-file gpurun_out/launches_tc_f16.csv \
    python probes/tc_probe.py --child time:f16:16384:16384:LLNS > gpurun_out/ncu_launches_tc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 127 -c 1 -o gpurun_out/prof_tc_f16_top -f \
    python probes/tc_probe.py --child time:f16:16384:16384:LLNS > gpurun_out/ncu_full_tc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc|diag_prep" -s 0 -c 3 -o gpurun_out/prof_tc_f32_trmm -f \
    python probes/tc_probe.py --child time:f32:16384:16384:LUTM > gpurun_out/ncu_full_tc32.log 2>&1
ls -la gpurun_out | tail -8
