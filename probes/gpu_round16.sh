#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "unit_diagonal or reference_grid or leaf_entry or fused_slab or low_precision" 2>&1 | tail -8 | tee gpurun_out/pytest_gpu16.txt
