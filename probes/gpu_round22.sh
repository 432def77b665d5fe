#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "gated or pipelined" 2>&1 | tail -6 | tee gpurun_out/pytest_gpu22.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561"
timeout 300 $TR probes/bench_configs.py --config C4 2>/dev/null | tee gpurun_out/configs_2gpu_b.txt
