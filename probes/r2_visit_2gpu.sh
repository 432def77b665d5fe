#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -12 | tee gpurun_out/r2_topo_2gpu.txt
timeout 600 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_conditioning.py -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r2_pytest_2gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551"
timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 2> gpurun_out/r2_bench2.err > gpurun_out/r2_bench_2gpu.json; tail -5 gpurun_out/r2_bench2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_2gpu.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], {k:v for k,v in d['e2e'].items() if k!='api'})
for x in d['extra'] or []: print({k:v for k,v in x.items() if k in('config','value','ms_per_step','backward_error','error')})
PY
