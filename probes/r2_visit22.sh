#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2_pytest_gpu_d.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; tail -3 gpurun_out/r2_bench_d.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_d.json') if l.startswith('{')][-1])
r=d['roofline']
print(d['value'], d['ms_per_step'], d['gpu_launches'], r['kernel'][:30], r['achieved'], r['frac'], r['whole_step_frac_of_peak'], {k:v for k,v in d['e2e'].items() if k in ('value','ms_per_step','floor_ms_per_step','limiter')})
for x in d['extra'] or []: print({k:v for k,v in x.items() if k in('config','value','ms_per_step','backward_error','error')}, x.get('sustained',{}).get('value'))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:slab2_f64 -s 1 -c 1 -o gpurun_out/r2_prof_slab2 -f \
    python probes/time_rectrxm.py --n 16384 --m 16384 --macro -1 --streams 1 --reps 1 > gpurun_out/r2_ncu_slab2.log 2>&1
tail -2 gpurun_out/r2_ncu_slab2.log
python probes/ncu_summarise.py rep gpurun_out/r2_prof_slab2.ncu-rep gpurun_out/r2_ncu_slab2_summary.csv; cat gpurun_out/r2_ncu_slab2_summary.csv
