#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29591"
timeout 200 $TR probes/bench_configs.py --config C4 --triangle 1 2>/dev/null | tee gpurun_out/configs_8gpu_c.txt
timeout 200 $TR probes/bench_configs.py --config C4 --triangle 0 2>/dev/null | tee -a gpurun_out/configs_8gpu_c.txt
