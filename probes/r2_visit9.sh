#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "lauum or gemm or tensor_core or low_precision or block_inverse" 2>&1 | tail -12 | tee gpurun_out/r2_pytest_lauum.txt
timeout 300 python probes/sweep_variants.py --n 32768 --m 16384 --dtypes float16 --cases RLNS --reps 3 2>&1 | tail -3 | tee gpurun_out/r2_c4slice.txt
timeout 300 python probes/sweep_variants.py --n 32768 --m 16384 --dtypes float16 --cases RLNS --reps 3 --opt inv_guard=0 2>&1 | tail -3 | tee -a gpurun_out/r2_c4slice.txt
timeout 300 python probes/sweep_variants.py --n 16384 --m 16384 --dtypes float16,float32 --cases LLNS,LUTM --reps 3 2>&1 | tail -5 | tee -a gpurun_out/r2_c4slice.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; tail -3 gpurun_out/r2_bench_b.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_b.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['roofline']['leaf_tflops'], d['roofline']['whole_step_frac_of_peak'], {k:v for k,v in d['e2e'].items() if k in ('value','ms_per_step')})
for x in d['extra'] or []: print({k:v for k,v in x.items() if k in('config','value','ms_per_step','backward_error','error')})
PY
