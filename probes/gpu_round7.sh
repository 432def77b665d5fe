#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "block_inverse or low_precision or tensor_core" 2>&1 | tail -5 | tee gpurun_out/pytest_gpu7.txt
for opt in inv_block=1024,inv_dup=1 inv_block=1024,inv_dup=0 inv_block=1024,streams=2 inv_block=512,streams=2; do
  timeout 300 python probes/tc_probe.py --child "time:f16:16384:16384:LLNS@$opt" 2>&1 | tee -a gpurun_out/tc_time_r7.txt | cut -c1-900
done
timeout 300 python probes/tc_probe.py --child "time:f16:32768:16384:RLNS@inv_block=1024" 2>&1 | tee -a gpurun_out/tc_time_r7.txt | cut -c1-900
timeout 300 python probes/tc_probe.py --child "time:f32:16384:16384:LLNS@inv_block=512" 2>&1 | tee -a gpurun_out/tc_time_r7.txt | cut -c1-900
