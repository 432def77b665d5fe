# Probe: per-CTA phase timestamps of ONE persistent pair-kernel launch (gemm_tc3): where a mid-size update spends its time.
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as ge
nla = ge.load_package(); h = nla.default_handle(0)
N = 16384
dbg = torch.zeros(64 * 148, dtype=torch.int64, device="cuda")
for MK in (1024, 2048):
    M = K = MK
    A = (torch.rand(K, M, device="cuda") - 0.5).half().t(); B = (torch.rand(N, K, device="cuda") - 0.5).half().t(); C = torch.rand(N, M, device="cuda").half().t()
    for r in range(3): nla._gemm(C, A, B, -1)
    torch.cuda.synchronize(); dbg.zero_(); h.set_option("tc_dbg", dbg.data_ptr())
    nla._gemm(C, A, B, -1); torch.cuda.synchronize(); h.set_option("tc_dbg", 0)
    d = dbg.view(148, 64).cpu().numpy().astype(np.int64)
    t0 = d[:, 0][d[:, 0] > 0].min()
    def stat(col):
        v = d[:, col][d[:, col] > 0]
        return None if len(v) == 0 else [round(float(x - t0) / 1e3, 2) for x in (v.min(), np.median(v), v.max())]
    out = {"M": M, "K": K, "N": N, "entry": stat(0), "prologue": stat(1)}
    for j in range(6):
        out["tile%d" % j] = {"first_stage": stat(4 + 4 * j), "mma_issued": stat(5 + 4 * j), "acc_done": stat(6 + 4 * j), "drained": stat(7 + 4 * j)}
    print(json.dumps(out), flush=True)
    # one cluster in detail (CTA 0 = leader of cluster 0, CTA 1 = its peer)
    for cta in (0, 1, 146):
        print("cta", cta, [round(float(x - t0) / 1e3, 2) if x > 0 else None for x in d[cta, :28]], flush=True)
