"""The oracle's restatement of the reference's recursive LU (oracle/reference_port.getrf2 <- src/lu.jl:185-299) against LAPACK: the
criterion of the reference's own tests (test/lu.jl:68-91: L*U ~ A[p, :]) on its grid, and pivots / factors / info equal to getrf's."""
import numpy as np
import pytest
from scipy.linalg import lapack

from oracle import reference_port as rp


def run(A0):
    A = np.array(A0, order="F", copy=True)
    ipiv = np.zeros(min(A.shape), dtype=np.int64)
    info = rp.getrf2(A, ipiv)
    return A, ipiv, info


@pytest.mark.parametrize("m,n", [(1, 1), (1, 5), (5, 1), (2, 2), (10, 10), (10, 9), (10, 11), (100, 100), (100, 90), (100, 110), (257, 200), (200, 257)])
def test_oracle_getrf2_matches_lapack_fp64(m, n):
    rng = np.random.RandomState(m * 131 + n)
    A0 = rng.rand(m, n) - 0.5
    LU, ipiv, info = run(A0)
    lu_ref, piv_ref, info_ref = lapack.dgetrf(A0)
    assert info == info_ref == 0
    assert np.array_equal(ipiv - 1, piv_ref)
    assert np.linalg.norm(LU - lu_ref) <= 1e-12 * np.linalg.norm(lu_ref)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_oracle_getrf2_reference_grid(dtype):
    """test/lu.jl:68-91 with the reference's default pivoting: m in 10/100/1000 (1000 only in Float64 here: the oracle recurses to single
    columns in Python), n in m, 0.9 m, 1.1 m; L*U ~ A[p, :] with isapprox's rtol sqrt(eps)."""
    for m in (10, 100) + ((1000,) if dtype == np.float64 else ()):
        for n in (m, m // 10 * 9, m // 10 * 11):
            rng = np.random.RandomState(m + 7 * n)
            A0 = rng.rand(m, n).astype(dtype)
            LU, ipiv, info = run(A0)
            assert info == 0 and LU.dtype == dtype
            k = min(m, n)
            L = np.tril(LU[:, :k], -1).astype(np.float64) + np.eye(m, k)
            U = np.triu(LU[:k, :]).astype(np.float64)
            PA = A0.astype(np.float64).copy()
            for i, p in enumerate(ipiv - 1):
                if p != i:
                    PA[[i, p]] = PA[[p, i]]
            assert np.linalg.norm(L @ U - PA) <= np.sqrt(np.finfo(dtype).eps) * max(np.linalg.norm(L @ U), np.linalg.norm(PA))


@pytest.mark.parametrize("n,zero_col", [(40, 0), (40, 17), (64, 63)])
def test_oracle_getrf2_info_is_the_first_zero_pivot(n, zero_col):
    rng = np.random.RandomState(n + zero_col)
    A0 = rng.rand(n, n) - 0.5
    A0[:, zero_col] = 0.0
    if zero_col + 5 < n:
        A0[:, zero_col + 5] = 0.0
    LU, ipiv, info = run(A0)
    lu_ref, piv_ref, info_ref = lapack.dgetrf(A0)
    assert info == info_ref == zero_col + 1
    assert np.array_equal(ipiv - 1, piv_ref)


def test_oracle_laswp_forward_and_backward():
    rng = np.random.RandomState(3)
    A0 = rng.rand(12, 37)
    ipiv = np.array([3, 2, 9, 12, 5, 11, 7, 8, 10, 10, 12, 12], dtype=np.int64)
    A = A0.copy()
    rp.laswp(A, 2, 9, ipiv, 1)
    want = A0.copy()
    for i in range(2, 10):
        p = ipiv[i - 1]
        want[[i - 1, p - 1]] = want[[p - 1, i - 1]]
    assert np.array_equal(A, want)
    rp.laswp(A, 2, 9, ipiv, -1)                # the reverse walk undoes the forward one
    assert np.array_equal(A, A0)


def test_oracle_getrf2_matches_the_frozen_vectors():
    """tests/golden/getrf2_golden.npz (made by tests/golden/make_getrf2_golden.py): the oracle reproduces its frozen factors, pivots and
    info bit for bit, and they agree with the LAPACK columns of the same file."""
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "getrf2_golden.npz"))
    keys = sorted(k[:-5] for k in g.files if k.endswith("_meta"))
    assert len(keys) >= 13
    for key in keys:
        A0 = g[key + "_A"]
        LU, ipiv, info = run(A0)
        assert np.array_equal(LU, g[key + "_oracle_lu"]) and np.array_equal(ipiv, g[key + "_oracle_ipiv"]) and info == int(g[key + "_info"][0])
        tol = 1e-12 if A0.dtype == np.float64 else 2e-5
        ref = g[key + "_lapack_lu"].astype(np.float64)
        assert np.linalg.norm(LU - ref) <= tol * max(1.0, np.linalg.norm(ref))
