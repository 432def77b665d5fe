#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551"
timeout 600 $TR bench.py --gpus 8 --steps 5 --warmup 3 2> gpurun_out/bench8.err | tee gpurun_out/bench_8gpu.json | cut -c1-200
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_8gpu.json') if l.startswith('{')][0]); print(d['value'], d['ms_per_step'], d['e2e'], d['clocks'])"
tail -2 gpurun_out/bench8.err
timeout 300 $TR probes/bench_configs.py --config C4 2>/dev/null | tee gpurun_out/configs_8gpu.txt
timeout 300 $TR probes/bench_configs.py --config C5 2>/dev/null | tee -a gpurun_out/configs_8gpu.txt
