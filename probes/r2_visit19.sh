#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_boundary.py -x -q -m gpu -k "streaming or host" 2>&1 | tail -15 | tee gpurun_out/r2_pytest_stream.txt
timeout 300 python probes/time_host_opts2.py 2>&1 | head -3 | tee gpurun_out/r2_host_stream.txt
