#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "pipelined_host" 2>&1 | tail -3
python probes/time_host_pipe.py 2>&1 | tail -5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531"
timeout 600 $TR bench.py --gpus 4 --steps 5 --warmup 3 2> gpurun_out/bench4.err | tee gpurun_out/bench_4gpu.json | cut -c1-200
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_4gpu.json') if l.startswith('{')][0]); print(d['value'], d['ms_per_step'], d['e2e'])"
