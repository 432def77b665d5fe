#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_getrf.py tests/test_gpu_parity.py -x -q -k "getrf or laswp or lu" > gpurun_out/pytest_getrf.txt 2>&1
echo "getrf tests exit $?"
tail -5 gpurun_out/pytest_getrf.txt
timeout 240 python probes/time_getrf.py > gpurun_out/time_getrf.txt 2>&1
echo "time exit $?"
cut -c1-175 gpurun_out/time_getrf.txt | tail -8
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/getrf_launches4.csv python probes/getrf_one.py 16384 f64 > gpurun_out/getrf_ncu.log 2>&1; python probes/agg_launches.py gpurun_out/getrf_launches4.csv
