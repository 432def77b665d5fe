"""Where does the Float32 LU residual come from?  n = 3000 under different option sets (LAPACK sgetrf on the same matrix: 7.8e-6)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import __graft_entry__ as ge
nla = ge.load_package()
h = nla.default_handle(0)
m = n = 3000
rng = np.random.RandomState(7 * m + n)
A0 = (rng.rand(m, n) - 0.5).astype(np.float32)
def resid():
    dA = nla.colmajor(A0)
    _, ipiv, info = nla.getrf2(dA)
    torch.cuda.synchronize()
    LU = nla.to_numpy(dA).astype(np.float64); piv = ipiv.cpu().numpy() - 1
    L = np.tril(LU, -1) + np.eye(n); U = np.triu(LU)
    PA = A0.astype(np.float64).copy()
    for i, p in enumerate(piv):
        if p != i: PA[[i, p]] = PA[[p, i]]
    return np.linalg.norm(PA - L @ U) / np.linalg.norm(A0)
import time
def solve_err_time(mode):
    # C3-like Float32 solve: n = 8192, m = 8192, backward error and time
    n2 = 8192
    g = torch.Generator(device="cuda").manual_seed(3)
    T = (torch.rand(n2, n2, device="cuda", dtype=torch.float32, generator=g) - 0.5)
    T = torch.tril(T) + torch.eye(n2, device="cuda") * n2 ** 0.5
    T = T.t().contiguous().t()
    B = (torch.rand(n2, n2, device="cuda", dtype=torch.float32, generator=g) - 0.5).t().contiguous().t()
    ts = []
    for it in range(4):
        X = B.clone(memory_format=torch.preserve_format)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        nla.unified_rectrxm("L", "L", "N", 1.0, "S", T, X); torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    R = T.double() @ X.double() - B.double()
    be = float(torch.linalg.norm(R) / (torch.linalg.norm(T.double()) * torch.linalg.norm(X.double())))
    return min(ts), be
for mode in (2, 1, 0):
    h.set_option("tf32_raw_hi", mode)
    ms, be = solve_err_time(mode)
    print("tf32_raw_hi", mode, "LU n=3000 resid", resid(), "| solve 8192x8192 ms", ms, "backward err", be, flush=True)
h.set_option("tf32_raw_hi", 1)
