"""Float32: error of the tensor-core update (3xTF32) and of the inverse-based solve with an LU factor's unit-lower L, each against
the FMA kernels (force_simt) and an FP64 reference."""
import sys
import numpy as np
import torch
from scipy.linalg import lu_factor
sys.path.insert(0, ".")
import __graft_entry__ as ge
nla = ge.load_package()
h = nla.default_handle(0)
rng = np.random.RandomState(1)
n = 3000; n1 = 1504
A0 = (rng.rand(n, n) - 0.5).astype(np.float32)
lu, piv = lu_factor(A0.astype(np.float64))
L11 = (np.tril(lu[:n1, :n1], -1) + np.eye(n1)).astype(np.float32)
print("cond(L11) =", np.linalg.cond(L11.astype(np.float64)), "max|U| =", np.abs(np.triu(lu)).max())
B = (rng.rand(n1, n - n1) - 0.5).astype(np.float32)
Xref = np.linalg.solve(L11.astype(np.float64), B.astype(np.float64))
for simt in (0, 1):
    h.set_option("force_simt", simt)
    for ib in ((0, 128) if not simt else (0,)):
        h.set_option("inv_block", ib)
        dL = nla.colmajor(L11); dB = nla.colmajor(B)
        nla.trsm("L", "L", "N", "U", dL, dB, 1.0)
        torch.cuda.synchronize()
        X = nla.to_numpy(dB).astype(np.float64)
        res = np.linalg.norm(L11.astype(np.float64) @ X - B) / (np.linalg.norm(L11) * np.linalg.norm(X))
        fwd = np.linalg.norm(X - Xref) / np.linalg.norm(Xref)
        print("trsm unit-lower n1=1504: simt", simt, "inv_block", ib, "residual", res, "forward", fwd, flush=True)
    h.set_option("inv_block", 0)
    # update C -= A21 * X with the magnitudes of an LU step
    A21 = np.tril(lu[n1:, :n1]).astype(np.float32) if False else lu[n1:, :n1].astype(np.float32)
    U12 = Xref.astype(np.float32)
    C = (rng.rand(n - n1, n - n1) - 0.5).astype(np.float32)
    want = C.astype(np.float64) - A21.astype(np.float64) @ U12.astype(np.float64)
    dC = nla.colmajor(C)
    nla.GEMM_SUB(dC, nla.colmajor(A21), nla.colmajor(U12))
    torch.cuda.synchronize()
    got = nla.to_numpy(dC).astype(np.float64)
    print("update K=1504: simt", simt, "error / ||want||", np.linalg.norm(got - want) / np.linalg.norm(want),
          " error / (|A||B|)", np.linalg.norm(got - want) / np.linalg.norm(np.abs(A21.astype(np.float64)) @ np.abs(U12.astype(np.float64))), flush=True)
h.set_option("force_simt", 0)
# one panel level deeper: small K updates
for K in (32, 64, 128, 512):
    A = (rng.rand(2048, K) - 0.5).astype(np.float32); Bm = (rng.rand(K, 2048) - 0.5).astype(np.float32); C = (rng.rand(2048, 2048) - 0.5).astype(np.float32)
    want = C.astype(np.float64) - A.astype(np.float64) @ Bm.astype(np.float64)
    for simt in (0, 1):
        h.set_option("force_simt", simt)
        dC = nla.colmajor(C); nla.GEMM_SUB(dC, nla.colmajor(A), nla.colmajor(Bm)); torch.cuda.synchronize()
        print("update K", K, "simt", simt, np.linalg.norm(nla.to_numpy(dC) - want) / np.linalg.norm(want), flush=True)
h.set_option("force_simt", 0)
