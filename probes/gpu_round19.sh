#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "recursive_lu or unit_diagonal" 2>&1 | tail -15 | tee gpurun_out/pytest_gpu19.txt
