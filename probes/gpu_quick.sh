#!/bin/bash
# quick GPU visit: parity tests then option sweeps
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.txt
python probes/time_rectrxm.py "$@" 2>&1 | tail -12 | tee gpurun_out/sweep.txt
