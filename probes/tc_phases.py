# Probe: per-CTA phase timestamps (globaltimer) of single tcgen05 kernel launches: leaf and small-K update.
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as ge
nla = ge.load_package(); h = nla.default_handle(0)
dt = torch.float16 if (len(sys.argv) < 2 or sys.argv[1] == "f16") else torch.float32
n, m = 4096, 16384
g = torch.Generator(device="cuda").manual_seed(1)
A = ((2 * torch.rand(n, n, device="cuda", generator=g) - 1) / n ** 0.5).tril(-1) + torch.diag(1 + torch.rand(n, device="cuda", generator=g))
PAD = int(sys.argv[2]) if len(sys.argv) > 2 else 0
A = A.to(dt).t().contiguous().t()
B0 = nla.empty_colmajor(n, m, dt, ld=n + PAD)
B0.copy_((torch.rand(n, m, device="cuda", generator=g) + 1).to(dt))
print("ld", B0.stride(1), "pad", PAD)
dbg = torch.zeros(8 * 8192, dtype=torch.int64, device="cuda")
def phases(label, fn, ncta):
    fn(); torch.cuda.synchronize()          # warm
    dbg.zero_(); h.set_option("tc_dbg", dbg.data_ptr())
    fn(); torch.cuda.synchronize(); h.set_option("tc_dbg", 0)
    d = dbg[:8 * ncta].view(ncta, 8).cpu().numpy().astype(np.int64)
    t0 = d[:, 0].min()
    names = ["entry", "prologue", "first_stage", "mma_issued", "acc_done", "drain_done", "tmem_free"]
    out = {"kernel": label, "ctas": ncta}
    for i, nm in enumerate(names):
        col = d[:, i][d[:, i] > 0]
        out[nm + "_us(min/med/max)"] = [round(float(x - t0) / 1e3, 2) for x in (col.min(), np.median(col), col.max())] if len(col) else None
    print(json.dumps(out), flush=True)
X = nla.empty_colmajor(n, m, dt, ld=n + PAD); X.copy_(B0)
h.set_option("streams", 1)
# leaf: one 128-block solve against all RHS (nla_trsm-like through unified_rectrxm with n=128)
A128 = A[:128, :128]; X128 = X[:128, :]
phases("leaf n=128 (prep + leaf GEMM K=128, BN auto)", lambda: nla.unified_rectrxm("L", "L", "N", 1.0, "S", A128, X128), 128)
# small-K update: C(128 x m) -= A(128x128) * B(128 x m)
C = X[128:256, :]; Bop = X[:128, :]; Aop = A[128:256, :128]
phases("update M=128 K=128", lambda: nla.GEMM_SUB(C, Aop, Bop), 128)
C = X[1024:2048, :]; Bop = X[:1024, :]; Aop = A[1024:2048, :1024]
phases("update M=1024 K=1024", lambda: nla.GEMM_SUB(C, Aop, Bop), 512)
