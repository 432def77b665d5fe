#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_full.txt 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu_full.txt
timeout 400 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
echo "bench exit $?"; cut -c1-1500 gpurun_out/bench_1gpu.json; tail -3 gpurun_out/bench_1gpu.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
