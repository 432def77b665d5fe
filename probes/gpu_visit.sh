#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601"
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 2> gpurun_out/bench2.err | tee gpurun_out/bench_2gpu_final.json | cut -c1-200
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_2gpu_final.json') if l.startswith('{')][0]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step'])"
timeout 120 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | cut -c1-200
