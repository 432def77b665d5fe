#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_parity.py -x -q -m gpu -k "gated or streaming or multi_gpu" 2>&1 | tail -4 | tee gpurun_out/r2_pytest_gated.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551"
for gm in 2048 4096; do
NLA_GATED_MACRO=$gm timeout 150 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-extra --no-e2e 2> gpurun_out/r2_bench2.err > gpurun_out/r2_bench_2gpu_gm$gm.json; tail -2 gpurun_out/r2_bench2.err | cut -c1-200
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_2gpu_gm$gm.json') if l.startswith('{')][-1])
print($gm, d['value'], d['ms_per_step'], d['gpu_launches'])
PY
done
