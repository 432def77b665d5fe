# Probe: Float16 update GEMM C -= A*B through nla_gemm_update for the shapes of a solve's levels, persistent pair kernel
# (gemm_tc3) vs one-tile kernels (gemm_tc2 / gemm_tc).  One JSON line per shape.
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
nla = ge.load_package(); h = nla.default_handle(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
for MK in (512, 1024, 2048, 4096, 8192, 16384):
    M = K = MK
    A = (torch.rand(K, M, dtype=torch.float32, device="cuda") - 0.5).half().t()      # column-major M x K
    B = (torch.rand(N, K, dtype=torch.float32, device="cuda") - 0.5).half().t()      # column-major K x N
    C = torch.rand(N, M, dtype=torch.float32, device="cuda").half().t()              # column-major M x N
    out = {"M": M, "N": N, "K": K}
    for persist in (1, 0):
        h.set_option("tc_persist", persist)
        ts = []
        for r in range(6):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); nla._gemm(C, A, B, -1); e1.record(); torch.cuda.synchronize()
            if r > 1: ts.append(e0.elapsed_time(e1))
        ms = min(ts)
        out["persist%d_us" % persist] = round(ms * 1e3, 1); out["persist%d_tflops" % persist] = round(2.0 * M * N * K / ms * 1e-9, 1)
    print(json.dumps(out), flush=True)
    del A, B, C
h.set_option("tc_persist", 1)
