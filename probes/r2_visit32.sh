#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_getrf.py tests/test_gpu_parity.py -x -q -k "getrf or laswp or lu" > gpurun_out/pytest_getrf.txt 2>&1
echo "getrf tests exit $?"
tail -3 gpurun_out/pytest_getrf.txt
timeout 240 python probes/time_getrf.py > gpurun_out/time_getrf.txt 2>&1
echo "time exit $?"
cut -c1-175 gpurun_out/time_getrf.txt | tail -8
for c in 8 0; do echo "NLA_GETRF_CLUSTER=$c"; NLA_GETRF_CLUSTER=$c timeout 120 python probes/time_getrf.py 4096 16384 2>&1 | tail -4 | cut -c1-150; done
