#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551"
for gm in 2048 4096 2048; do
NLA_GATED_MACRO=$gm timeout 120 $TR bench.py --gpus 8 --steps 8 --warmup 3 --no-extra --no-e2e 2> gpurun_out/r2_bench8.err > gpurun_out/r2_bench_8gpu_gm$gm.json
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_8gpu_gm$gm.json') if l.startswith('{')][-1])
print($gm, d['value'], d['ms_per_step'], d['gpu_launches'])
PY
done
