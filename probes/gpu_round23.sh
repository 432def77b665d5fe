#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571"
timeout 300 $TR probes/bench_configs.py --config C4 2>/dev/null | tee gpurun_out/configs_8gpu_b.txt
timeout 300 $TR probes/bench_configs.py --config C4 --no-pipeline 2>/dev/null | tee -a gpurun_out/configs_8gpu_b.txt
