#!/bin/bash
# round 2, visit 2: new boundary / scale tests, whole GPU suite, bench with the extra legs, reference arm timing, ncu capture of the top update
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_scale.py -x -q -m gpu -s 2>&1 | tail -25 | tee gpurun_out/r2_pytest_new.txt
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r2_pytest_gpu.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -3 gpurun_out/r2_bench.err; cut -c1-600 gpurun_out/r2_bench.json
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; tail -4 gpurun_out/r2_bench_ref.err; cut -c1-400 gpurun_out/r2_bench_ref.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f64_tma -s 3 -c 1 -o gpurun_out/r2_prof_gemm_top -f \
    python probes/time_rectrxm.py --n 16384 --m 16384 --macro 2048 --streams 1 --reps 0 > gpurun_out/r2_ncu_gemm.log 2>&1
tail -2 gpurun_out/r2_ncu_gemm.log
python probes/ncu_summarise.py rep gpurun_out/r2_prof_gemm_top.ncu-rep gpurun_out/r2_ncu_gemm_top_summary.csv; head -12 gpurun_out/r2_ncu_gemm_top_summary.csv
