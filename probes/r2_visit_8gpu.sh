#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -14 > gpurun_out/r2_topo_8gpu.txt; nproc >> gpurun_out/r2_topo_8gpu.txt; free -g | head -2 >> gpurun_out/r2_topo_8gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551"
timeout 900 $TR bench.py --gpus 8 --steps 5 --warmup 3 2> gpurun_out/r2_bench8.err > gpurun_out/r2_bench_8gpu.json; tail -5 gpurun_out/r2_bench8.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_8gpu.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], {k:v for k,v in d['e2e'].items() if k!='api'})
for x in d['extra'] or []: print({k:v for k,v in x.items() if k in('config','value','ms_per_step','backward_error','error','sustained')})
PY
