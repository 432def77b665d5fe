# Probe: times every side/uplo/trans/func variant of unified_rectrxm for the three element types on device-resident
# synthetic inputs (scaled recipe) and checks the backward error on a slab of 256 vectors.  One JSON line per case.
import argparse, itertools, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=8192); ap.add_argument("--m", type=int, default=8192)
ap.add_argument("--dtypes", default="float64,float32,float16"); ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--cases", default="")
ap.add_argument("--opt", default="", help="k=v,k=v handle options")
a = ap.parse_args()
nla = ge.load_package(); h = nla.default_handle(0)
for kv in filter(None, a.opt.split(",")):
    k, v = kv.split("="); h.set_option(k, int(v))
n, m = a.n, a.m
cases = a.cases.split(",") if a.cases else ["".join(c) for c in itertools.product("LR", "LU", "NT", "SM")]
for dname in a.dtypes.split(","):
    dt = getattr(torch, dname)
    for case in cases:
        side, uplo, trans, func = case
        g = torch.Generator(device="cuda").manual_seed(1)
        A = (2 * torch.rand(n, n, dtype=torch.float32, device="cuda", generator=g) - 1) / n ** 0.5
        A = (torch.tril(A, -1) if uplo == "L" else torch.triu(A, 1)) + torch.diag(1 + torch.rand(n, dtype=torch.float32, device="cuda", generator=g))
        A = A.to(dt).t().contiguous().t()
        shape = (n, m) if side == "L" else (m, n)
        B0 = (torch.rand(shape, dtype=torch.float32, device="cuda", generator=g) + 1).to(dt).t().contiguous().t()
        X = B0.clone(memory_format=torch.preserve_format)
        ts = []
        for r in range(a.reps + 1):
            X.copy_(B0); torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            h.launch_count(reset=True)
            e0.record(); nla.unified_rectrxm(side, uplo, trans, 1.0, func, A, X); e1.record(); torch.cuda.synchronize()
            if r > 0: ts.append(e0.elapsed_time(e1))
        ms = min(ts)
        Ad = (torch.tril(A) if uplo == "L" else torch.triu(A)).double()
        opA = Ad.t() if trans != "N" else Ad
        if side == "L":
            Xs, Bs = X[:, :256].double(), B0[:, :256].double()
            R = (opA @ Xs - Bs) if func == "S" else (Xs - opA @ Bs)
        else:
            Xs, Bs = X[:256, :].double(), B0[:256, :].double()
            R = (Xs @ opA - Bs) if func == "S" else (Xs - Bs @ opA)
        den = (torch.linalg.norm(opA) * torch.linalg.norm(Xs) + torch.linalg.norm(Bs)) if func == "S" else torch.linalg.norm(opA) * torch.linalg.norm(Bs)
        print(json.dumps({"dtype": dname, "case": case, "n": n, "m": m, "ms": round(ms, 3), "tflops": round(n * n * m / ms * 1e-9, 1),
                          "launches": h.launch_count(), "err": float(torch.linalg.norm(R) / den)}), flush=True)
        del Ad, opA, R, A, B0, X
