#!/bin/bash
# 2-GPU visit: weak-scaling bench with the panel-pipelined broadcast of A vs one blocking broadcast; C5 shape on one GPU
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 2> gpurun_out/bench2.err | tee gpurun_out/bench_2gpu_pipelined.json | cut -c1-700
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-pipeline --no-e2e 2> gpurun_out/bench2b.err | tee gpurun_out/bench_2gpu_blocking.json | cut -c1-400
timeout 600 python probes/sweep_variants.py --n 32768 --m 8192 --dtypes float64 --cases LLNM,LLNS --reps 2 2>&1 | tee gpurun_out/sweep_c5.txt
tail -3 gpurun_out/bench2.err
