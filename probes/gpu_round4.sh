#!/bin/bash
# GPU visit 4: parity suite, headline bench, all-variant sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.txt
timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-900
timeout 900 python probes/sweep_variants.py 2>&1 | tee gpurun_out/sweep_r4.txt
