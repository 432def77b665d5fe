# Probe: times unified_rectrxm on device-resident synthetic inputs (scaled recipe) for option sweeps.
import argparse, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=16384); ap.add_argument("--m", type=int, default=16384)
ap.add_argument("--side", default="L"); ap.add_argument("--uplo", default="L"); ap.add_argument("--trans", default="N"); ap.add_argument("--func", default="S")
ap.add_argument("--dtype", default="float64"); ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--streams", default="1"); ap.add_argument("--leaf", default="128"); ap.add_argument("--macro", default="1024")
ap.add_argument("--opt", default="", help="k=v,k=v handle options")
a = ap.parse_args()
nla = ge.load_package(); h = nla.default_handle(0)
for kv in filter(None, a.opt.split(",")):
    k, v = kv.split("="); h.set_option(k, int(v))
dt = getattr(torch, a.dtype); n, m = a.n, a.m
g = torch.Generator(device="cuda").manual_seed(1)
A = (2 * torch.rand(n, n, dtype=torch.float32, device="cuda", generator=g) - 1).to(dt) / n ** 0.5
A = (torch.tril(A, -1) if a.uplo == "L" else torch.triu(A, 1)) + torch.diag(1 + torch.rand(n, dtype=torch.float32, device="cuda", generator=g).to(dt))
A = A.t().contiguous().t()
shape = (n, m) if a.side == "L" else (m, n)
B0 = (torch.rand(shape, dtype=torch.float32, device="cuda", generator=g) + 1).to(dt).t().contiguous().t()
X = B0.clone(memory_format=torch.preserve_format)
for st in [int(s) for s in a.streams.split(",")]:
  for macro in [int(s) for s in a.macro.split(",")]:
    for leaf in [int(s) for s in a.leaf.split(",")]:
        h.set_option("streams", st); h.set_option("leaf", leaf); h.set_option("macro", macro)
        ts = []
        for r in range(a.reps + 1):
            X.copy_(B0); torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            h.launch_count(reset=True)
            e0.record(); nla.unified_rectrxm(a.side, a.uplo, a.trans, 1.0, a.func, A, X); e1.record(); torch.cuda.synchronize()
            if r > 0: ts.append(e0.elapsed_time(e1))
        ms = min(ts); fl = n * n * m
        print(json.dumps({"n": n, "m": m, "case": a.side + a.uplo + a.trans + a.func, "dtype": a.dtype, "streams": st, "leaf": leaf, "macro": macro, "ms_min": round(ms, 3),
                          "ms_all": [round(t, 2) for t in ts], "tflops": round(fl / ms * 1e-9, 2), "launches": h.launch_count(), "opt": a.opt}), flush=True)
