"""Per-call cost at the reference's own test sizes (n <= 256) and a little above: host time to enqueue one call (the call is asynchronous)
and device time per call, warm handle.  VERDICT r01 weak #10: 'per-call host work is unmeasured for small n'."""
import json, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import __graft_entry__ as ge
nla = ge.load_package()
h = nla.default_handle(0)
out = []
for dtype in (torch.float64, torch.float32, torch.float16):
    for n, m in ((16, 8), (64, 64), (128, 128), (256, 256), (512, 256), (1024, 1024), (2048, 2048)):
        A = (torch.rand(n, n, device="cuda", dtype=torch.float32) - 0.5).to(dtype)
        A = (torch.tril(A) + torch.eye(n, device="cuda", dtype=dtype) * n ** 0.5).t().contiguous().t()
        B = (torch.rand(m, n, device="cuda", dtype=torch.float32) - 0.5).to(dtype).t()       # n x m column-major
        for func in "SM":
            for _ in range(5):
                nla.unified_rectrxm("L", "L", "N", 1.0, func, A, B)
            torch.cuda.synchronize()
            reps = 200
            h.launch_count(reset=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter(); e0.record()
            for _ in range(reps):
                nla.unified_rectrxm("L", "L", "N", 1.0, func, A, B)
            e1.record(); t1 = time.perf_counter()
            torch.cuda.synchronize()
            rec = {"dtype": str(dtype).split(".")[1], "n": n, "m": m, "func": func, "host_us_per_call": (t1 - t0) / reps * 1e6,
                   "device_us_per_call": e0.elapsed_time(e1) / reps * 1e3, "launches_per_call": h.launch_count() / reps}
            out.append(rec); print(json.dumps(rec), flush=True)
json.dump(out, open("gpurun_out/small_n_latency.json", "w"), indent=1)
