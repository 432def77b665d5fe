#!/bin/bash
mkdir -p gpurun_out
for ib in 128 512 1024; do
  timeout 300 python probes/tc_probe.py --child "time:f16:16384:16384:LLNS@inv_block=$ib" 2>&1 | sed "s/^/ib=$ib /" | tee -a gpurun_out/tc_time_r6.txt
  timeout 300 python probes/tc_probe.py --child "time:f16:16384:256:LLNS@inv_block=$ib" 2>&1 | sed "s/^/ib=$ib /" | tee -a gpurun_out/tc_time_r6.txt
done
timeout 300 python probes/tc_probe.py --child "time:f32:16384:16384:LLNS@inv_block=1024" 2>&1 | sed "s/^/ib=1024 /" | tee -a gpurun_out/tc_time_r6.txt
timeout 300 python probes/tc_probe.py --child "time:f32:16384:256:LLNS@inv_block=1024" 2>&1 | sed "s/^/ib=1024 /" | tee -a gpurun_out/tc_time_r6.txt
timeout 300 python probes/tc_probe.py --child "time:f32:16384:256:LLNS@inv_block=128" 2>&1 | sed "s/^/ib=128 /" | tee -a gpurun_out/tc_time_r6.txt
