#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "pipelined_host or gated or host_buffer" 2>&1 | tail -5 | tee gpurun_out/pytest_gpu11.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 2> gpurun_out/bench2.err | tee gpurun_out/bench_2gpu_pipelined.json | cut -c1-200
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_2gpu_pipelined.json') if l.startswith('{')][0]); print(d['value'], d['ms_per_step'], d['e2e'])"
tail -3 gpurun_out/bench2.err
