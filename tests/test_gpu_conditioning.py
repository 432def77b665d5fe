"""Numerics of the Float32 / Float16 solve leaves under controlled conditioning (ADVICE medium).

The default low-precision solve multiplies by explicitly inverted diagonal blocks of order 1024 (csrc/tri_inv.cuh), which is only
conditionally backward stable (error ~ eps * cond(block)); the reference's leaf is a substitution (src/trsm.jl:15-27), which is backward
stable for any triangular matrix.  The library therefore guards every inverted block on the device (csrc/tri_guard.cuh) and solves the
rejected ones by substitution.  These tests
  * run matrix families whose diagonal blocks have cond ~ 1e1 .. 1e6 with right-hand sides that make the inverse-based leaf cancel
    (B = op(A) * X_true with X_true = O(1): then |inv T| |b| >> |x|), compare default / guard off / 128-blocks / true substitution
    (force_simt) and require: wherever substitution meets the tolerance, the default path meets it too;
  * force the fallback on benign matrices (threshold 1) for every solve variant and check it against the inverse-based result.
Tolerances (north_star): Float32 1e-5, Float16 1e-2, normwise backward error in FP64."""
import itertools

import numpy as np
import pytest

from oracle import reference_port as rp

pytestmark = pytest.mark.gpu

TOL = {np.float32: 1e-5, np.float16: 1e-2}


def same_sign(n, c, seed=0):
    """Lower triangular, O(1) diagonal, all off-diagonal entries negative: inv(T) is entrywise positive and grows like exp(c sqrt(n));
    cond_2 of a 1024-block: c = 0.5 -> 6e1, 1.0 -> 6e3, 1.5 -> 5e5."""
    r = np.random.RandomState(seed)
    return np.tril(-r.rand(n, n) * c / np.sqrt(n), -1) + np.diag(1 + r.rand(n))


def qr_tri(n, kappa, seed=0):
    """Triangular factor of a matrix with geometrically spaced singular values: cond_2(T) = kappa exactly."""
    r = np.random.RandomState(seed)
    U, _ = np.linalg.qr(r.randn(n, n)); V, _ = np.linalg.qr(r.randn(n, n))
    s = kappa ** (-np.arange(n) / (n - 1.0))
    R = np.linalg.qr((U * s) @ V.T)[1]
    R = R * np.sign(np.diag(R))[:, None]
    return R.T.copy()


def berr(side, A64, X, B064):
    X = X.astype(np.float64)
    R = (A64 @ X if side == "L" else X @ A64) - B064
    return np.linalg.norm(R) / (np.linalg.norm(A64) * np.linalg.norm(X) + np.linalg.norm(B064))


def run(nla, side, uplo, trans, alpha, A, B0, diag="N"):
    import torch

    dA, dB = nla.colmajor(A), nla.colmajor(B0)
    nla.trsm(side, uplo, trans, diag, dA, dB, alpha)
    torch.cuda.synchronize()
    return nla.to_numpy(dB)


FAMILIES = [("same_sign", 0.5), ("same_sign", 1.0), ("same_sign", 1.5), ("qr", 1e2), ("qr", 1e4), ("qr", 1e6)]


@pytest.mark.parametrize("dtype", [np.float32, np.float16])
@pytest.mark.parametrize("fam,par", FAMILIES)
def test_ill_conditioned_diagonal_blocks(nla, gpu, dtype, fam, par):
    n, m = 2048, 96
    L = same_sign(n, par) if fam == "same_sign" else qr_tri(n, par)
    rng = np.random.RandomState(3)
    rows = []
    try:
        for side, uplo, trans in [("L", "L", "N"), ("L", "U", "T"), ("R", "L", "N"), ("R", "U", "N")]:
            Ast = np.asfortranarray((L if uplo == "L" else L.T).astype(dtype))          # stored matrix
            op = Ast.astype(np.float64) if trans == "N" else Ast.astype(np.float64).T     # op(A) in FP64 from the ROUNDED entries
            Xt = 2 * rng.rand(n, m) - 1 if side == "L" else 2 * rng.rand(m, n) - 1
            with np.errstate(over="ignore"):
                B0 = np.asfortranarray((op @ Xt if side == "L" else Xt @ op).astype(dtype))
            if not np.isfinite(B0).all():
                continue
            B064 = B0.astype(np.float64)
            cond_blk = max(np.linalg.cond(op[o:o + 1024, o:o + 1024]) for o in range(0, n, 1024))
            res = {}
            for label, opts in (("default", {}), ("guard_off", {"inv_guard": 0}), ("ib128", {"inv_block": 128}), ("subst", {"force_simt": 1})):
                for k, v in opts.items():
                    gpu.set_option(k, v)
                try:
                    X = run(nla, side, uplo, trans, 1.0, Ast, B0)
                    res[label] = berr(side, op, X, B064) if np.isfinite(X).all() else float("inf")
                    if label == "default":
                        res["fallbacks"] = gpu.get_option("inv_fallbacks")
                finally:
                    for k in opts:
                        gpu.set_option(k, {"inv_guard": 1, "inv_block": 0, "force_simt": 0}[k])
            rows.append((side + uplo + trans, cond_blk, res))
            tol = TOL[dtype]
            # the contract: wherever the reference's arithmetic (substitution) meets the tolerance, so does the default path
            if res["subst"] < tol:
                assert res["default"] < tol, (fam, par, side, uplo, trans, cond_blk, res)
            # and the guard never makes things worse than the unguarded inverse path by more than rounding noise
            if np.isfinite(res["guard_off"]) and res["guard_off"] < tol:
                assert res["default"] < tol, (fam, par, side, uplo, trans, res)
    finally:
        for k, v in (("inv_guard", 1), ("inv_block", 0), ("force_simt", 0)):
            gpu.set_option(k, v)
    for r in rows:
        print(f"{np.dtype(dtype).name} {fam}({par:g}) {r[0]} cond(1024-block) {r[1]:.1e}: " + ", ".join(f"{k} {v:.2e}" if isinstance(v, float) else f"{k} {v}" for k, v in r[2].items()))


@pytest.mark.parametrize("dtype", [np.float32, np.float16])
def test_forced_substitution_fallback_all_variants(nla, gpu, dtype):
    """inv_guard_kappa = 1 rejects every inverted block: each leaf runs tri_subst_kernel instead of the inverse GEMM.  Every solve
    variant, ragged order (last block 1024 < n mod 1024 != 0, rows not a multiple of 16), alpha != 1, unit diagonal; the result must
    meet the tolerance, agree with the inverse-based result, and the counter must report one fallback per block."""
    n, m = 2600, 200
    gpu.set_option("inv_guard_kappa", 1)
    try:
        for side, uplo, trans, diag in itertools.product("LR", "LU", "NT", "NU"):
            A, B0 = rp.make_inputs(n, m, side, uplo, dtype, seed=n + 5, recipe="scaled")
            A1 = A.copy()
            if diag == "U":
                np.fill_diagonal(A1, 1)
                An = A.copy(); np.fill_diagonal(An, np.nan)     # the stored diagonal must not be read
            else:
                An = A
            got = run(nla, side, uplo, trans, -1.5, An, B0, diag)
            assert gpu.get_option("inv_fallbacks") == -(-n // 1024), (side, uplo, trans, diag)
            assert np.isfinite(got).all(), (side, uplo, trans, diag)
            err = rp.error_metric(side, uplo, trans, -1.5, "S", A1, B0, got)
            assert err < TOL[dtype], (side, uplo, trans, diag, err)
            gpu.set_option("inv_guard_kappa", 0)
            try:
                base = run(nla, side, uplo, trans, -1.5, An, B0, diag)
                assert gpu.get_option("inv_fallbacks") == 0
            finally:
                gpu.set_option("inv_guard_kappa", 1)
            d = np.linalg.norm(got.astype(np.float64) - base.astype(np.float64)) / np.linalg.norm(base.astype(np.float64))
            assert d < (2e-5 if dtype == np.float32 else 1e-2), (side, uplo, trans, diag, d)
    finally:
        gpu.set_option("inv_guard_kappa", 0)


def test_nonfinite_inverse_is_caught_fp16(nla, gpu):
    """A block whose inverse overflows Float16 (entries > 65504) although the solution itself is representable: badly scaled rows.
    T = D * L0 with D = diag(2^-14 ... ): inv(T) = inv(L0) inv(D) has entries ~ 2^14..2^17.  The guard must route it to substitution,
    which solves it to tolerance; with the guard off the inverse path returns non-finite values or misses the tolerance."""
    n, m = 1024, 64
    rng = np.random.RandomState(9)
    L0 = np.tril((2 * rng.rand(n, n) - 1) / np.sqrt(n), -1) + np.diag(1 + rng.rand(n))
    d = np.where(np.arange(n) % 2 == 0, 2.0 ** -17, 1.0)
    T = np.asfortranarray((d[:, None] * L0).astype(np.float16))
    op = T.astype(np.float64)
    Xt = 2 * rng.rand(n, m) - 1
    B0 = np.asfortranarray((op @ Xt).astype(np.float16))
    B064 = B0.astype(np.float64)
    X = run(nla, "L", "L", "N", 1.0, T, B0)
    assert gpu.get_option("inv_fallbacks") == 1
    assert np.isfinite(X).all()
    e_def = berr("L", op, X, B064)
    gpu.set_option("force_simt", 1)
    try:
        e_sub = berr("L", op, run(nla, "L", "L", "N", 1.0, T, B0), B064)
    finally:
        gpu.set_option("force_simt", 0)
    assert e_sub < 1e-2 and e_def < 1e-2, (e_def, e_sub)
    gpu.set_option("inv_guard", 0)
    try:
        Xo = run(nla, "L", "L", "N", 1.0, T, B0)
        e_off = berr("L", op, Xo, B064) if np.isfinite(Xo).all() else float("inf")
    finally:
        gpu.set_option("inv_guard", 1)
    print(f"fp16 badly scaled block: default {e_def:.2e}, substitution {e_sub:.2e}, guard off {e_off:.2e}")
