# Probe for the tcgen05 path: every (dtype, majorness) GEMM variant through nla_gemm_update, then unified_rectrxm in
# Float16 / Float32 for all side/uplo/trans/func.  Each group runs in its own subprocess with a timeout so that a trap in
# one variant does not hide the others.  Prints one JSON line per case.
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(group):
    import numpy as np, torch
    import __graft_entry__ as ge
    nla = ge.load_package(); h = nla.default_handle(0)
    dts = {"f16": (np.float16, torch.float16), "f32": (np.float32, torch.float32)}
    group, _, opts = group.partition("@")   # e.g. trx:f32@tc_bn=128,tf32_raw_hi=1
    for kv in filter(None, opts.split(",")):
        k, v = kv.split("="); h.set_option(k, int(v))
    kind, dname = group.split(":")[:2]
    group = group + ("@" + opts if opts else "")
    npdt, tdt = dts[dname]
    rng = np.random.RandomState(0)

    def rel(a, b):
        a = a.astype(np.float64); b = b.astype(np.float64)
        return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))

    if kind == "gemm":
        ta, tb = group.split(":")[2][:2]
        for (M, N, K) in [(128, 256, 64), (128, 256, 128), (256, 512, 256), (200, 300, 96), (1024, 640, 512), (384, 1000, 1024)]:
            A = rng.rand(M, K).astype(npdt) - 0.5; B = rng.rand(K, N).astype(npdt) - 0.5; C = rng.rand(M, N).astype(npdt)
            Ain = np.asfortranarray(A.T.copy() if ta == "T" else A); Bin = np.asfortranarray(B.T.copy() if tb == "T" else B)
            dA, dB, dC = nla.colmajor(Ain), nla.colmajor(Bin), nla.colmajor(np.asfortranarray(C))
            h.launch_count(reset=True)
            nla.GEMM_ADD(dA, dB, dC, transa=ta, transb=tb); torch.cuda.synchronize()
            want = C.astype(np.float64) + A.astype(np.float64) @ B.astype(np.float64)
            print(json.dumps({"case": group, "MNK": [M, N, K], "rel": rel(nla.to_numpy(dC), want), "launches": h.launch_count()}), flush=True)
    elif kind == "trx":
        from oracle import reference_port as rp
        import itertools
        sizes = [(128, 256), (256, 64), (700, 96), (1024, 512)]
        for (n, m) in sizes:
            worst = {}
            for side, uplo, trans, func in itertools.product("LR", "LU", "NT", "SM"):
                A, B0 = rp.make_inputs(n, m, side, uplo, npdt, seed=n + m, recipe="scaled")
                dA, dB = nla.colmajor(A), nla.colmajor(B0)
                nla.unified_rectrxm(side, uplo, trans, 1.0, func, dA, dB); torch.cuda.synchronize()
                err = rp.error_metric(side, uplo, trans, 1.0, func, A, B0, nla.to_numpy(dB))
                worst[side + uplo + trans + func] = float(err)
            k = max(worst, key=worst.get)
            print(json.dumps({"case": group, "n": n, "m": m, "worst": k, "err": worst[k],
                              "bad": {c: e for c, e in worst.items() if not (e < (1e-5 if dname == "f32" else 1e-2))}}), flush=True)
    elif kind == "time":
        n = int(group.split(":")[2]); m = int(group.split(":")[3]); case = group.split(":")[4][:4]
        side, uplo, trans, func = case
        g = torch.Generator(device="cuda").manual_seed(1)
        A = (2 * torch.rand(n, n, dtype=torch.float32, device="cuda", generator=g) - 1) / n ** 0.5
        A = (torch.tril(A, -1) if uplo == "L" else torch.triu(A, 1)) + torch.diag(1 + torch.rand(n, dtype=torch.float32, device="cuda", generator=g))
        A = A.to(tdt).t().contiguous().t()
        shape = (n, m) if side == "L" else (m, n)
        B0 = (torch.rand(shape, dtype=torch.float32, device="cuda", generator=g) + 1).to(tdt).t().contiguous().t()
        X = B0.clone(memory_format=torch.preserve_format)
        for streams in (1, 0):
            h.set_option("streams", streams)
            ts = []
            for r in range(4):
                X.copy_(B0); torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                h.launch_count(reset=True)
                e0.record(); nla.unified_rectrxm(side, uplo, trans, 1.0, func, A, X); e1.record(); torch.cuda.synchronize()
                if r > 0: ts.append(e0.elapsed_time(e1))
            ms = min(ts)
            # backward error on a slab of 256 vectors, FP64
            Ad = (torch.tril(A) if uplo == "L" else torch.triu(A)).double()
            opA = Ad.t() if trans != "N" else Ad
            if side == "L":
                Xs, Bs = X[:, :256].double(), B0[:, :256].double()
                R = (opA @ Xs - Bs) if func == "S" else (Xs - opA @ Bs)
                den = (torch.linalg.norm(opA) * torch.linalg.norm(Xs) + torch.linalg.norm(Bs)) if func == "S" else torch.linalg.norm(opA) * torch.linalg.norm(Bs)
            else:
                Xs, Bs = X[:256, :].double(), B0[:256, :].double()
                R = (Xs @ opA - Bs) if func == "S" else (Xs - Bs @ opA)
                den = (torch.linalg.norm(opA) * torch.linalg.norm(Xs) + torch.linalg.norm(Bs)) if func == "S" else torch.linalg.norm(opA) * torch.linalg.norm(Bs)
            print(json.dumps({"case": group, "streams": streams, "ms_min": round(ms, 3), "tflops": round(n * n * m / ms * 1e-9, 1), "launches": h.launch_count(),
                              "err": float(torch.linalg.norm(R) / den)}), flush=True)
            del Ad, opA, R
        # per-launch breakdown (one stream): kind, algorithmic flops -> total ms, TFLOP/s
        h.set_option("streams", 1); h.set_option("profile", 1)
        X.copy_(B0); nla.unified_rectrxm(side, uplo, trans, 1.0, func, A, X); torch.cuda.synchronize()
        prof = h.profile_read(); h.set_option("profile", 0)
        agg = {}
        for k, f, ms in prof:
            if k == 2: continue   # span record
            key = ("gemm" if k == 1 else "leaf", f)
            c = agg.setdefault(key, [0, 0.0]); c[0] += 1; c[1] += ms
        rows = [{"kind": k, "gflop_each": round(f * 1e-9, 1), "count": c, "ms_total": round(ms, 3), "tflops": round(f * c / ms * 1e-9, 1)} for (k, f), (c, ms) in sorted(agg.items(), key=lambda kv: -kv[0][1])]
        print(json.dumps({"case": group, "breakdown": rows}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2]); sys.exit(0)
    groups = sys.argv[1:] or (["gemm:%s:%s" % (d, t) for d in ("f16", "f32") for t in ("NN", "TN", "NT")] + ["trx:f16", "trx:f32"])
    for g in groups:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", g], capture_output=True, text=True, timeout=300)
            out = r.stdout.strip()
            print(out if out else json.dumps({"case": g, "no_output": True}), flush=True)
            if r.returncode != 0:
                print(json.dumps({"case": g, "rc": r.returncode, "stderr": r.stderr[-600:]}), flush=True)
        except subprocess.TimeoutExpired:
            print(json.dumps({"case": g, "timeout": True}), flush=True)
