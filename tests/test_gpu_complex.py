"""Complex element types on the recursive TRSM / TRMM path: ComplexF64 / ComplexF32 with trans = 'C' distinct from
'T', complex alpha, unit diagonal -- through the C ABI (nla_rectrxm_complex), against OpenBLAS ztrsm / ztrmm / ctrsm / ctrmm (the routine
family the reference's tests use as their oracle, test/unified_rectrxm.jl:36-40).  The reference advertises complex support
(README.md:20) and builds Adjoint(A) for 'C' (src/rectrxm.jl:57), but its recursion only accepts real element types (:101), so there is
no reference output to compare with: parity is BLAS + normwise backward error (tolerances of north_star: 1e-13 double, 1e-5 single)."""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def make(n, m, side, uplo, dtype, seed):
    rng = np.random.RandomState(seed)
    A = (rng.rand(n, n) - 0.5 + 1j * (rng.rand(n, n) - 0.5)) / np.sqrt(n)
    A = np.tril(A, -1) if uplo == "L" else np.triu(A, 1)
    A = A + np.diag(1 + rng.rand(n) + 1j * (rng.rand(n) - 0.5))
    shape = (n, m) if side == "L" else (m, n)
    B = rng.rand(*shape) + 1 + 1j * (rng.rand(*shape) - 0.5)
    return np.asfortranarray(A.astype(dtype)), np.asfortranarray(B.astype(dtype))


def blas_ref(side, uplo, trans, diag, alpha, func, A, B0):
    from scipy.linalg import blas

    pre = "z" if A.dtype == np.complex128 else "c"
    f = getattr(blas, pre + ("trsm" if func == "S" else "trmm"))
    return f(A.dtype.type(alpha), A, B0, side=0 if side == "L" else 1, lower=1 if uplo == "L" else 0, trans_a={"N": 0, "T": 1, "C": 2}[trans],
             diag=1 if diag == "U" else 0)


def op_of(A, uplo, trans, diag):
    T = (np.tril(A) if uplo == "L" else np.triu(A)).astype(np.complex128)
    if diag == "U":
        np.fill_diagonal(T, 1)
    return T if trans == "N" else (T.T if trans == "T" else T.conj().T)


def error_metric(side, uplo, trans, diag, alpha, func, A, B0, X):
    T = op_of(A, uplo, trans, diag)
    X, B = X.astype(np.complex128), B0.astype(np.complex128)
    if func == "S":
        R = (T @ X if side == "L" else X @ T) - alpha * B
        return np.linalg.norm(R) / (np.linalg.norm(T) * np.linalg.norm(X) + abs(alpha) * np.linalg.norm(B))
    P = alpha * (T @ B if side == "L" else B @ T)
    return np.linalg.norm(X - P) / (abs(alpha) * np.linalg.norm(T) * np.linalg.norm(B))


def run(nla, side, uplo, trans, diag, alpha, func, A, B0):
    import torch

    dA, dB = nla.colmajor(A), nla.colmajor(B0)
    assert dA.dtype in (torch.complex64, torch.complex128)
    nla.unified_trxm(side, uplo, trans, diag, alpha, func, dA, dB)
    torch.cuda.synchronize()
    return nla.to_numpy(dB)


@pytest.mark.parametrize("dtype,tol", [(np.complex128, 1e-13), (np.complex64, 1e-5)])
@pytest.mark.parametrize("n,m", [(16, 3), (96, 40), (300, 33), (1000, 72), (2048, 264)])
def test_complex_all_variants(nla, gpu, dtype, tol, n, m):
    """Leaf sizes (n <= 128), the recursion with the generic GEMM (odd sizes) and with the tensor-core GEMMs (n = 2048), every
    side / uplo / trans in {N, T, C} / func, complex alpha; 'C' must differ from 'T' and both must match BLAS."""
    alpha = 0.75 - 0.5j
    for side, uplo, trans, func in itertools.product("LR", "LU", "NTC", "SM"):
        A, B0 = make(n, m, side, uplo, dtype, seed=n + m)
        An = A.copy()
        An[np.triu_indices(n, 1) if uplo == "L" else np.tril_indices(n, -1)] = np.nan      # the opposite triangle must never matter
        got = run(nla, side, uplo, trans, "N", alpha, func, An, B0)
        assert np.isfinite(got).all(), (side, uplo, trans, func)
        ref = blas_ref(side, uplo, trans, "N", alpha, func, A, B0)
        rel = np.linalg.norm(got.astype(np.complex128) - ref.astype(np.complex128)) / np.linalg.norm(ref.astype(np.complex128))
        err = error_metric(side, uplo, trans, "N", alpha, func, A, B0, got)
        assert err < tol, (side, uplo, trans, func, err)
        assert rel < (1e-12 if dtype == np.complex128 else 2e-5), (side, uplo, trans, func, rel)
    # 'C' is not 'T' for a matrix with a non-zero imaginary part
    A, B0 = make(n, m, "L", "L", dtype, seed=1)
    xt = run(nla, "L", "L", "T", "N", 1.0, "M", A, B0)
    xc = run(nla, "L", "L", "C", "N", 1.0, "M", A, B0)
    assert np.linalg.norm(xt - xc) / np.linalg.norm(xt) > 1e-2


@pytest.mark.parametrize("dtype,tol", [(np.complex128, 1e-13), (np.complex64, 1e-5)])
def test_complex_unit_diagonal_and_errors(nla, gpu, dtype, tol):
    import torch

    n, m = 640, 136
    for side, uplo, trans, func in itertools.product("LR", "LU", "NC", "SM"):
        A, B0 = make(n, m, side, uplo, dtype, seed=9)
        An = A.copy(); np.fill_diagonal(An, np.nan)          # diag = 'U': the stored diagonal is not read
        got = run(nla, side, uplo, trans, "U", -1.5 + 0.25j, func, An, B0)
        assert np.isfinite(got).all()
        assert error_metric(side, uplo, trans, "U", -1.5 + 0.25j, func, A, B0, got) < tol, (side, uplo, trans, func)
    A, B0 = make(64, 8, "L", "L", dtype, seed=2)
    dA, dB = nla.colmajor(A), nla.colmajor(B0)
    with pytest.raises(nla.NextLAError):
        nla.unified_rectrxm("L", "L", "X", 1.0, "S", dA, dB)
    with pytest.raises(nla.NextLAError):     # complex matrices through a real-only entry point
        nla.LeftLowerTRSM(dA, dB)
    # real alpha, empty problems
    nla.unified_rectrxm("L", "L", "N", 2.0, "S", dA, dB); torch.cuda.synchronize()
    assert error_metric("L", "L", "N", "N", 2.0, "S", A, B0, nla.to_numpy(dB)) < tol
    nla.unified_rectrxm("L", "L", "N", 1.0, "S", nla.colmajor(np.zeros((0, 0), dtype=dtype)), nla.colmajor(np.zeros((0, 4), dtype=dtype)))


def test_complex_large_backward_error(nla, gpu):
    """ComplexF64 at n = m = 4096 (tensor-core GEMM updates, FP64 DMMA): backward error on the GPU with an independent cuBLAS product."""
    import torch

    n = m = 4096
    g = torch.Generator(device="cuda").manual_seed(5)
    re = (2 * torch.rand(n, n, dtype=torch.float64, device="cuda", generator=g) - 1) / n ** 0.5
    im = (2 * torch.rand(n, n, dtype=torch.float64, device="cuda", generator=g) - 1) / n ** 0.5
    A = torch.tril(torch.complex(re, im), -1) + torch.diag(torch.complex(1 + torch.rand(n, dtype=torch.float64, device="cuda", generator=g),
                                                                         torch.rand(n, dtype=torch.float64, device="cuda", generator=g) - 0.5))
    dA = A.t().contiguous().t()
    B0 = torch.complex(torch.rand(n, m, dtype=torch.float64, device="cuda", generator=g) + 1, torch.rand(n, m, dtype=torch.float64, device="cuda", generator=g))
    B0 = B0.t().contiguous().t()
    X = B0.clone(memory_format=torch.preserve_format)
    gpu.launch_count(reset=True)
    nla.unified_rectrxm("L", "L", "C", 1.0, "S", dA, X)
    torch.cuda.synchronize()
    assert gpu.launch_count() > 100
    T = torch.tril(dA).conj().t()
    err = (torch.linalg.norm(T @ X - B0) / (torch.linalg.norm(T) * torch.linalg.norm(X) + torch.linalg.norm(B0))).item()
    assert err < 1e-13, err
