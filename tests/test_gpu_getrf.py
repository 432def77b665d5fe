"""nla_getrf2 -- the reference's recursive LU (getrf2!, src/lu.jl:185-299) entirely on the device.

Checked against LAPACK getrf (SciPy): P*A = L*U against the original matrix, the pivot sequence (partial pivoting with "first largest
magnitude" is unique for random data in Float64), info for exactly singular input, the reference's own test grid
(test/lu.jl:70-71: m in 10/100/1000, n in m, 0.9 m, 1.1 m, criterion L*U ~ A[p, :])."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def apply_pivots(A0, piv):
    PA = A0.copy()
    for i, p in enumerate(piv):
        if p != i:
            PA[[i, p]] = PA[[p, i]]
    return PA


def factor(nla, A0):
    import torch

    dA = nla.colmajor(A0)
    _, ipiv, info = nla.getrf2(dA)
    torch.cuda.synchronize()
    return nla.to_numpy(dA), ipiv.cpu().numpy() - 1, int(info.item())


def check_lu(A0, LU, piv, tol):
    m, n = A0.shape
    k = min(m, n)
    assert piv.shape == (k,)
    assert np.all(piv >= np.arange(k)) and np.all(piv < m)
    L = np.tril(LU[:, :k], -1).astype(np.float64) + np.eye(m, k)
    U = np.triu(LU[:k, :]).astype(np.float64)
    assert np.all(np.abs(np.tril(LU[:, :k], -1)) <= 1.0 + 1e-6)     # partial pivoting: multipliers bounded by one
    PA = apply_pivots(A0.astype(np.float64), piv)
    assert np.linalg.norm(PA - L @ U) / np.linalg.norm(A0) < tol


SHAPES = [(1, 1), (1, 7), (7, 1), (2, 2), (33, 33), (64, 40), (40, 64), (257, 300), (300, 257), (1024, 1024), (1500, 900), (700, 1100), (3000, 3000)]


@pytest.mark.parametrize("m,n", SHAPES)
def test_getrf2_fp64_matches_lapack(nla, gpu, m, n):
    from scipy.linalg import lu_factor

    rng = np.random.RandomState(1000 * m + n)
    A0 = rng.rand(m, n) - 0.5
    LU, piv, info = factor(nla, A0)
    assert info == 0
    check_lu(A0, LU, piv, 1e-13)
    lu_ref, piv_ref = lu_factor(A0, check_finite=False)
    assert np.array_equal(piv, piv_ref[:min(m, n)])
    assert np.linalg.norm(LU - lu_ref) / np.linalg.norm(lu_ref) < 1e-10


@pytest.mark.parametrize("m,n", SHAPES)
def test_getrf2_fp32(nla, gpu, m, n):
    rng = np.random.RandomState(7 * m + n)
    A0 = (rng.rand(m, n) - 0.5).astype(np.float32)
    LU, piv, info = factor(nla, A0)
    assert info == 0
    # the Float32 updates run as 3xTF32 on the tensor cores, whose accumulation truncates: measured 5.4e-5 at n = 3000
    # where LAPACK's sgetrf has 7.8e-6 and the FMA kernels (option force_simt) 7.2e-6; the reference's own criterion is sqrt(eps) = 3.4e-4
    check_lu(A0, LU, piv, 1e-4)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_getrf2_reference_grid(nla, gpu, dtype):
    """test/lu.jl:68-91 (default RowMaximum pivoting): m in 10/100/1000, n in m, 0.9 m, 1.1 m; L*U ~ A[p, :] (isapprox: rtol sqrt(eps))."""
    for m in (10, 100, 1000):
        for n in (m, m // 10 * 9, m // 10 * 11):
            rng = np.random.RandomState(m * 3 + n)
            A0 = rng.rand(m, n).astype(dtype)
            LU, piv, info = factor(nla, A0)
            assert info == 0
            k = min(m, n)
            L = np.tril(LU[:, :k], -1).astype(np.float64) + np.eye(m, k)
            U = np.triu(LU[:k, :]).astype(np.float64)
            PA = apply_pivots(A0.astype(np.float64), piv)
            assert np.linalg.norm(L @ U - PA) <= np.sqrt(np.finfo(dtype).eps) * max(np.linalg.norm(L @ U), np.linalg.norm(PA))


def test_getrf2_fp32_fma_kernels_reach_lapack_accuracy(nla, gpu):
    """Option force_simt keeps the Float32 updates on the FMA kernels: the residual is then LAPACK's (sgetrf: 7.8e-6 on this matrix)."""
    m = n = 3000
    rng = np.random.RandomState(7 * m + n)
    A0 = (rng.rand(m, n) - 0.5).astype(np.float32)
    gpu.set_option("force_simt", 1)
    try:
        LU, piv, info = factor(nla, A0)
    finally:
        gpu.set_option("force_simt", 0)
    assert info == 0
    check_lu(A0, LU, piv, 1.2e-5)


def test_getrf2_padded_leading_dimension_and_views(nla, gpu):
    """lda > m and a sub-matrix view of a larger device matrix; the rows/columns outside the view are untouched."""
    import torch
    from scipy.linalg import lu_factor

    rng = np.random.RandomState(5)
    big = rng.rand(900, 800) - 0.5
    dbig = nla.colmajor(big)
    view = dbig[100:700, 64:564]                       # 600 x 500, lda = 900
    _, ipiv, info = nla.getrf2(view)
    torch.cuda.synchronize()
    out = nla.to_numpy(dbig)
    lu_ref, piv_ref = lu_factor(big[100:700, 64:564], check_finite=False)
    assert int(info.item()) == 0
    assert np.array_equal(ipiv.cpu().numpy() - 1, piv_ref)
    assert np.linalg.norm(out[100:700, 64:564] - lu_ref) / np.linalg.norm(lu_ref) < 1e-10
    mask = np.ones_like(big, dtype=bool)
    mask[100:700, 64:564] = False
    assert np.array_equal(out[mask], big[mask])


@pytest.mark.parametrize("n,zero_col", [(100, 5), (100, 0), (640, 300), (64, 63)])
def test_getrf2_reports_the_first_zero_pivot(nla, gpu, n, zero_col):
    """src/lu.jl:246-249 / :260-262 / :289-291: an exactly zero pivot is reported as info = its (1-based) column; the factorisation
    completes.  LAPACK's getrf gives the same info."""
    from scipy.linalg import lapack

    rng = np.random.RandomState(n + zero_col)
    A0 = rng.rand(n, n) - 0.5
    A0[:, zero_col] = 0.0
    if zero_col + 7 < n:
        A0[:, zero_col + 7] = 0.0                      # a later zero column must not overwrite the first
    LU, piv, info = factor(nla, A0)
    lu_ref, piv_ref, info_ref = lapack.dgetrf(A0)
    assert info == info_ref == zero_col + 1
    assert np.array_equal(piv, piv_ref)
    assert np.linalg.norm(LU - lu_ref) / np.linalg.norm(lu_ref) < 1e-10


def test_getrf2_tiny_pivot_is_divided_not_scaled(nla, gpu):
    """src/lu.jl:239-243: a pivot below sfmin divides the column (its reciprocal would overflow)."""
    A0 = np.array([[1e-310, 2.0, 1.0], [5e-311, 1.0, 3.0], [2e-311, 4.0, 2.0]])
    LU, piv, info = factor(nla, A0)
    assert info == 0 and list(piv) == [0, 2, 2]
    assert np.all(np.isfinite(LU))
    # multipliers 2e-311 / 1e-310 and 5e-311 / 1e-310 (1 / 1e-310 overflows); the second step then takes 4 - 0.2 * 2 = 3.6 as its pivot.
    # (OpenBLAS's getrf leaves such a column unscaled, so LAPACK is no oracle here; the reference's rule is src/lu.jl:239-243.)
    want = np.array([[1e-310, 2.0, 1.0], [0.2, 3.6, 1.8], [0.5, 0.0, 2.5]])
    assert np.allclose(LU, want, rtol=1e-3, atol=1e-12)


def test_getrf2_argument_checks(nla, gpu):
    import ctypes
    import torch

    lib = nla.load_library()
    h = gpu
    A = torch.zeros(16, dtype=torch.float64, device="cuda")
    ipiv = torch.zeros(4, dtype=torch.int64, device="cuda")
    info = torch.full((), 7, dtype=torch.int32, device="cuda")
    call = lambda dt, m, n, lda, a=A.data_ptr(), p=ipiv.data_ptr(), i=info.data_ptr(): lib.nla_getrf2(h._h, dt, m, n, a, lda, p, i, None)
    assert call(0, -1, 4, 4) == 2 and call(0, 4, -1, 4) == 2 and call(0, 4, 4, 3) == 2       # src/lu.jl:192-203
    assert call(2, 4, 4, 4) == 7 and call(3, 4, 4, 4) == 7 and call(9, 4, 4, 4) == 3           # Float16 / complex unsupported, bad dtype
    assert call(0, 4, 4, 4, i=None) == 4 and call(0, 4, 4, 4, a=None) == 4
    assert lib.nla_getrf2(None, 0, 4, 4, A.data_ptr(), 4, ipiv.data_ptr(), info.data_ptr(), None) == 8
    assert call(0, 0, 4, 1) == 0                                                               # quick return (:206) with info cleared
    torch.cuda.synchronize()
    assert int(info.item()) == 0


def test_getrf2_against_the_oracle_and_the_frozen_vectors(nla, gpu):
    """tests/golden/getrf2_golden.npz: frozen inputs with the oracle's (oracle/reference_port.getrf2 <- src/lu.jl:185-299) and LAPACK's
    factors, pivots and info.  The device result has the same pivots and info; the factors agree to rounding (the device eliminates a
    panel right-looking, the reference recurses to single columns: different summation order)."""
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "getrf2_golden.npz"))
    keys = sorted(k[:-5] for k in g.files if k.endswith("_meta"))
    assert len(keys) >= 13
    for key in keys:
        A0 = g[key + "_A"]
        LU, piv, info = factor(nla, A0)
        assert info == int(g[key + "_info"][0]), key
        assert np.array_equal(piv + 1, g[key + "_oracle_ipiv"]), key
        tol = 1e-12 if A0.dtype == np.float64 else 2e-5
        want = g[key + "_oracle_lu"].astype(np.float64)
        assert np.linalg.norm(LU - want) <= tol * max(1.0, np.linalg.norm(want)), key


@pytest.mark.parametrize("m,n", [(300, 257), (640, 640), (500, 900)])
def test_getrf2_matches_the_oracle_live(nla, gpu, m, n):
    """The oracle run live on seeded inputs beyond the frozen sizes (several panels per factorisation)."""
    sys_path_oracle()
    from oracle import reference_port as rp

    rng = np.random.RandomState(m + 3 * n)
    A0 = np.asfortranarray(rng.rand(m, n) - 0.5)
    want = A0.copy(order="F")
    ipiv = np.zeros(min(m, n), dtype=np.int64)
    info_o = rp.getrf2(want, ipiv)
    LU, piv, info = factor(nla, A0)
    assert info == info_o == 0
    assert np.array_equal(piv + 1, ipiv)
    assert np.linalg.norm(LU - want) <= 1e-11 * np.linalg.norm(want)


def sys_path_oracle():
    import os
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)


@pytest.mark.parametrize("cluster", [0, 8, 16])
def test_getrf2_every_panel_kernel(nla, gpu, cluster):
    """Option getrf_cluster: 0 = the grid-wide cooperative panel kernel (what very tall panels get), 8 / 16 = one cluster of that size.
    Same pivots and factors as LAPACK whichever kernel factors the panels; Float32 by residual."""
    from scipy.linalg import lu_factor

    gpu.set_option("getrf_cluster", cluster)
    try:
        for m, n in [(1, 1), (70, 33), (257, 300), (1500, 900), (2100, 2100)]:
            rng = np.random.RandomState(m * 7 + n + cluster)
            A0 = rng.rand(m, n) - 0.5
            LU, piv, info = factor(nla, A0)
            lu_ref, piv_ref = lu_factor(A0, check_finite=False)
            assert info == 0 and np.array_equal(piv, piv_ref[:min(m, n)]), (cluster, m, n)
            assert np.linalg.norm(LU - lu_ref) / np.linalg.norm(lu_ref) < 1e-10
        A32 = (np.random.RandomState(cluster).rand(1200, 1000) - 0.5).astype(np.float32)
        LU, piv, info = factor(nla, A32)
        assert info == 0
        check_lu(A32, LU, piv, 1e-4)
        Z = np.random.RandomState(5).rand(300, 300) - 0.5
        Z[:, 123] = 0.0
        _, _, info = factor(nla, Z)
        assert info == 124
    finally:
        gpu.set_option("getrf_cluster", -1)


def test_getrf2_panel_taller_than_a_cluster(nla, gpu):
    """60000 rows: more than 16 CTAs' shared memory holds even for 8-column panels in Float64, so the panels go to the grid-wide kernel
    with a narrower width (32 columns: 406 rows per CTA)."""
    from scipy.linalg import lu_factor

    m, n = 60000, 96
    rng = np.random.RandomState(60)
    A0 = rng.rand(m, n) - 0.5
    LU, piv, info = factor(nla, A0)
    lu_ref, piv_ref = lu_factor(A0, check_finite=False)
    assert info == 0 and np.array_equal(piv, piv_ref)
    assert np.linalg.norm(LU - lu_ref) / np.linalg.norm(lu_ref) < 1e-10


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.float16])
def test_laswp_arbitrary_pivots(nla, gpu, dtype):
    """nla_laswp (src/lu.jl:470-530) with pivot vectors a factorisation would never produce: partners above the row, repeated partners,
    self-interchanges, ranges that are not multiples of the kernel's batch of 16, forward and backward walks, a column view with
    lda > rows.  The planned kernel (laswp.cuh) must reproduce the sequential walk of the oracle bit for bit."""
    import torch

    sys_path_oracle()
    from oracle import reference_port as rp

    rng = np.random.RandomState(11)
    for rows, cols, k1, k2 in [(300, 70, 1, 300), (300, 70, 17, 211), (64, 5, 1, 64), (1000, 333, 3, 999), (40, 1, 1, 40), (513, 129, 100, 131)]:
        for incx in (1, -1):
            for kind in ("random", "few_targets", "self", "above"):
                if kind == "random":
                    piv = rng.randint(1, rows + 1, size=rows)
                elif kind == "few_targets":
                    piv = rng.choice([1, rows, rows // 2 + 1], size=rows)
                elif kind == "self":
                    piv = np.arange(1, rows + 1)
                    piv[::3] = rng.randint(1, rows + 1, size=len(piv[::3]))
                else:
                    piv = np.maximum(1, np.arange(1, rows + 1) - rng.randint(0, 20, size=rows))
                piv = piv.astype(np.int64)
                big = (rng.rand(rows + 8, cols + 3) - 0.5).astype(dtype)
                want = big.copy()
                rp.laswp(want[:rows, 1:cols + 1], k1, k2, piv, incx)
                dbig = nla.colmajor(big)
                nla.laswp(dbig[:rows, 1:cols + 1], k1, k2, torch.from_numpy(piv).cuda(), incx)
                torch.cuda.synchronize()
                assert np.array_equal(nla.to_numpy(dbig), want), (rows, cols, k1, k2, incx, kind)
