#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521"
timeout 600 $TR probes/bench_configs.py --config C4 2>/dev/null | tee -a gpurun_out/configs_2gpu.txt
timeout 600 $TR probes/bench_configs.py --config C4 --no-pipeline 2>/dev/null | tee -a gpurun_out/configs_2gpu.txt
timeout 600 $TR probes/bench_configs.py --config C5 2>/dev/null | tee -a gpurun_out/configs_2gpu.txt
timeout 600 $TR probes/bench_configs.py --config C5 --no-pipeline 2>/dev/null | tee -a gpurun_out/configs_2gpu.txt
