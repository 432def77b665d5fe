#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_getrf.py tests/test_gpu_parity.py -x -q -k "getrf or leaf_entry" > gpurun_out/pytest_getrf.txt 2>&1
echo "tests exit $?"; tail -3 gpurun_out/pytest_getrf.txt
timeout 200 python probes/small_n_latency.py > gpurun_out/small_n_latency.txt 2>&1
echo "latency exit $?"; tail -42 gpurun_out/small_n_latency.txt | cut -c1-170
