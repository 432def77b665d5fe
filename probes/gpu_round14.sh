#!/bin/bash
# GPU visit 14: full parity suite, headline bench, variant sweep, ncu launch list of the bench step, ncu full capture of the persistent kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu14.txt
timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-300
timeout 900 python probes/sweep_variants.py --n 16384 --m 16384 --dtypes float32,float16 2>&1 | tee gpurun_out/sweep_r14.txt | cut -c1-200
timeout 300 python probes/sweep_variants.py --n 32768 --m 16384 --dtypes float16 --cases RLNS 2>&1 | tee -a gpurun_out/sweep_r14.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_step.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc3 -s 1 -c 2 -o gpurun_out/prof_tc3 -f \
    python probes/tc_probe.py --child "time:f16:16384:16384:LLNS@inv_overlap=0" > gpurun_out/ncu_tc3.log 2>&1
ls -la gpurun_out | tail -4
