# Small end-to-end pass over the newer kernels (block inverses, persistent pair kernel incl. edge tiles, dup epilogue, transposed
# right side, laswp, unit diagonal) meant to be run under `compute-sanitizer --tool memcheck`.
import itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as ge
from oracle import reference_port as rp
nla = ge.load_package(); h = nla.default_handle(0)
QUICK = "--quick" in sys.argv   # racecheck / synccheck are ~100x slower than memcheck: one size, the variants that differ in kernels
worst = 0.0
for dtype, tol in ((np.float16, 1e-2), (np.float32, 1e-5), (np.float64, 1e-13)):
    for (n, m) in (((1304, 200),) if QUICK else ((1300, 200), (2304, 136))):
        for side, uplo, trans, func in ([("L", "L", "N", "S"), ("L", "U", "T", "S"), ("R", "L", "N", "S"), ("L", "L", "N", "M"), ("R", "U", "T", "M")] if QUICK
                                        else itertools.product("LR", "LU", "NT", "SM")):
            A, B0 = rp.make_inputs(n, m, side, uplo, dtype, seed=3, recipe="scaled")
            dA, dB = nla.colmajor(A), nla.colmajor(B0)
            nla.unified_trxm(side, uplo, trans, "N", 1.5, func, dA, dB); torch.cuda.synchronize()
            err = rp.error_metric(side, uplo, trans, 1.5, func, A, B0, nla.to_numpy(dB))
            assert err < tol, (dtype, n, m, side, uplo, trans, func, err)
            worst = max(worst, err / tol)
# GEMM shapes with ragged edges through the persistent kernel
rng = np.random.RandomState(0)
for (M, N, K) in ((384, 520, 320), (1000, 1544, 200)):
    A = (rng.rand(M, K) - 0.5).astype(np.float16); B = (rng.rand(K, N) - 0.5).astype(np.float16); C = rng.rand(M, N).astype(np.float16)
    dC = nla.colmajor(np.asfortranarray(C)); nla.GEMM_SUB(dC, nla.colmajor(np.asfortranarray(A)), nla.colmajor(np.asfortranarray(B))); torch.cuda.synchronize()
    want = C.astype(np.float64) - A.astype(np.float64) @ B.astype(np.float64)
    assert np.linalg.norm(nla.to_numpy(dC) - want) / np.linalg.norm(want) < 1e-3
# laswp
A = rng.rand(300, 70); dA = nla.colmajor(A); piv = torch.from_numpy(rng.randint(1, 301, size=300).astype(np.int64)).cuda()
nla.laswp(dA, 1, 300, piv, 1); torch.cuda.synchronize()
ref = A.copy()
for i, p in enumerate(piv.cpu().numpy()):
    ref[[i, p - 1]] = ref[[p - 1, i]]
assert np.array_equal(nla.to_numpy(dA), ref)
# lauum (triangle-masked GEMM epilogues) and the forced substitution fallback of the conditioning guard
for dtype, tol in ((np.float64, 1e-13), (np.float32, 3e-5), (np.float16, 1e-2)):
    n = 520
    F = ((rng.rand(n, n) - 0.5) / np.sqrt(n) + np.eye(n)).astype(dtype)
    for uplo in "LU":
        T = np.tril(F) if uplo == "L" else np.triu(F)
        dA = nla.colmajor(T); nla.lauum(uplo, dA, 256); torch.cuda.synchronize()
        T64 = T.astype(np.float64); want = T64.T @ T64 if uplo == "L" else T64 @ T64.T
        mask = np.tril(np.ones((n, n), bool)) if uplo == "L" else np.triu(np.ones((n, n), bool))
        assert np.linalg.norm(nla.to_numpy(dA)[mask] - want[mask]) / np.linalg.norm(want[mask]) < tol
h.set_option("inv_guard_kappa", 1)
for dtype, tol in ((np.float16, 1e-2), (np.float32, 1e-5)):
    for side, uplo in (("L", "L"), ("R", "U")):
        A, B0 = rp.make_inputs(1304, 200, side, uplo, dtype, seed=4, recipe="scaled")   # 1304: a 16-byte pitch, i.e. the tensor-core path
        dA, dB = nla.colmajor(A), nla.colmajor(B0)
        nla.unified_rectrxm(side, uplo, "N", 1.0, "S", dA, dB); torch.cuda.synchronize()
        fb = h.get_option("inv_fallbacks")
        assert fb == 2, (dtype, side, uplo, fb)
        assert rp.error_metric(side, uplo, "N", 1.0, "S", A, B0, nla.to_numpy(dB)) < tol
h.set_option("inv_guard_kappa", 0)
# recursive LU: cooperative panel kernel (cross-CTA polling), planned laswp, unit-lower solves, updates
from scipy.linalg import lu_factor
for dtype, m, n in ((np.float64, 700, 520), (np.float32, 300, 420)) if not QUICK else ((np.float64, 330, 200),):
    A0 = (rng.rand(m, n) - 0.5).astype(dtype)
    dA = nla.colmajor(A0)
    _, ipiv, info = nla.getrf2(dA); torch.cuda.synchronize()
    lu_ref, piv_ref = lu_factor(A0.astype(np.float64), check_finite=False)
    assert int(info.item()) == 0
    if dtype == np.float64:
        assert np.array_equal(ipiv.cpu().numpy() - 1, piv_ref[:min(m, n)])
        assert np.linalg.norm(nla.to_numpy(dA) - lu_ref) / np.linalg.norm(lu_ref) < 1e-11
print("sanitize_small ok, worst err/tol", worst)
