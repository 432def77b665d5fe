#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list + full capture of the dominant kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json
tail -3 gpurun_out/bench.err
if [ "$1" != "nosmoke" ]; then python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2; fi
if [ "$2" != "noncu" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_f64|leaf_kernel|gemm_simt" -c 520 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_f64_tma -s 63 -c 1 -o gpurun_out/prof_gemm_top -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:leaf_kernel -s 5 -c 1 -o gpurun_out/prof_leaf -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_leaf.log 2>&1
ls -la gpurun_out
fi
