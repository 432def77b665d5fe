"""Time nla_getrf2 (recursive LU on the device) against torch.linalg.lu_factor (cuSOLVER) -- context only, not a bench line."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import __graft_entry__ as ge

nla = ge.load_package()
h = nla.default_handle(0)
out = []
sizes = [int(a) for a in sys.argv[1:]] or [4096, 8192, 16384]
for dtype in (torch.float64, torch.float32):
    for n in sizes:
        g = torch.Generator(device="cuda").manual_seed(n)
        A0 = (torch.rand(n, n, device="cuda", dtype=dtype, generator=g) - 0.5).t()      # column-major view
        flops = 2.0 / 3.0 * n ** 3
        times = []
        for it in range(4):
            A = A0.clone(memory_format=torch.preserve_format)
            assert A.stride() == (1, n)
            torch.cuda.synchronize()
            h.launch_count(reset=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _, ipiv, info = nla.getrf2(A)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        launches = h.launch_count()
        ref = []
        for it in range(3):
            A = A0.clone(memory_format=torch.preserve_format)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            LU, piv = torch.linalg.lu_factor(A)
            e1.record()
            torch.cuda.synchronize()
            ref.append(e0.elapsed_time(e1))
        # residual of ours on the last run
        rec = {"dtype": str(dtype), "n": n, "ms": min(times), "tflops": flops / min(times) / 1e9, "launches": launches,
               "cusolver_ms": min(ref), "cusolver_tflops": flops / min(ref) / 1e9, "info": int(info.item()),
               "pivots_equal_cusolver": bool(torch.equal(ipiv.to(torch.int32), piv.to(torch.int32)))}
        print(json.dumps(rec), flush=True)
        out.append(rec)
if len(sys.argv) == 1:
    json.dump(out, open("gpurun_out/time_getrf.json", "w"), indent=1)
