#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_getrf.py tests/test_gpu_parity.py -x -q -k "getrf or laswp or lu" > gpurun_out/pytest_getrf.txt 2>&1
echo "getrf tests exit $?"
tail -5 gpurun_out/pytest_getrf.txt
timeout 240 python probes/time_getrf.py > gpurun_out/time_getrf.txt 2>&1
echo "time exit $?"
tail -8 gpurun_out/time_getrf.txt
for r in 128 256 512; do echo "NLA_GETRF_ROWS=$r"; NLA_GETRF_ROWS=$r timeout 120 python probes/time_getrf.py 16384 2>&1 | tail -2 | cut -c1-200; done
