#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_getrf.py -x -q > gpurun_out/pytest_getrf.txt 2>&1
echo "getrf tests exit $?"
tail -15 gpurun_out/pytest_getrf.txt
timeout 240 python probes/time_getrf.py > gpurun_out/time_getrf.txt 2>&1
echo "time exit $?"
tail -8 gpurun_out/time_getrf.txt
