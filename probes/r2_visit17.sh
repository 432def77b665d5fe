#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2_pytest_gpu_c.txt
timeout 300 python probes/time_rectrxm.py --n 16384 --m 16384 --macro 4096,8192 --streams 4 --reps 3 2>&1 | tee gpurun_out/r2_macro8192.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; tail -3 gpurun_out/r2_bench_c.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_c.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['leaf_tflops'], d['roofline']['whole_step_frac_of_peak'], {k:v for k,v in d['e2e'].items() if k in ('value','ms_per_step','host_link_ceiling','floor_ms_per_step','limiter')})
for x in d['extra'] or []: print({k:v for k,v in x.items() if k in('config','value','ms_per_step','backward_error','error')}, x.get('sustained',{}).get('value'))
PY
