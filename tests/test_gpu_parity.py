"""Parity tests proper: the CUDA library (through its C ABI) against the oracle (CPU restatement of the reference,
oracle/) and against OpenBLAS trsm/trmm -- the reference's own test oracle (test/unified_rectrxm.jl:36-40).

Tolerances (stated by BASELINE.json north_star, written here):
  FP64: relative difference to BLAS < 1e-14 on the reference's grid (its own bar, test/unified_rectrxm.jl:8);
        normwise backward error < 1e-13 elsewhere
  FP32: < 1e-5 (test/trsm.jl:8)        FP16: backward error < 1e-2
"""
import itertools

import numpy as np
import pytest

from oracle import c_port
from oracle import reference_port as rp

pytestmark = pytest.mark.gpu

SIDES, UPLOS, TRANS, FUNCS = "LR", "LU", "NTC", "SM"


def run_gpu(nla, side, uplo, trans, alpha, func, A, B0, ld_pad=0):
    import torch

    dA = nla.colmajor(A)
    if ld_pad:
        rows, cols = B0.shape
        dB = nla.empty_colmajor(rows, cols, getattr(torch, str(B0.dtype)), ld=rows + ld_pad)
        dB.copy_(torch.from_numpy(np.ascontiguousarray(B0)).cuda())
        rowsA = A.shape[0]
        dA2 = nla.empty_colmajor(rowsA, rowsA, getattr(torch, str(A.dtype)), ld=rowsA + ld_pad)
        dA2.copy_(torch.from_numpy(np.ascontiguousarray(A)).cuda())
        dA = dA2
    else:
        dB = nla.colmajor(B0)
    nla.unified_rectrxm(side, uplo, trans, alpha, func, dA, dB)
    torch.cuda.synchronize()
    return nla.to_numpy(dB)


def rel(a, b):
    a = a.astype(np.float64); b = b.astype(np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("n", [16, 32, 128, 256])
@pytest.mark.parametrize("m", [1, 8, 64])
def test_reference_grid_fp64(nla, gpu, n, m):
    """The reference's own test grid and criterion (test/unified_rectrxm.jl:10-44): FP64, alpha = 1, all
    side/uplo/trans/func, relative error vs BLAS < 1e-14; additionally vs the oracle."""
    for side, uplo, trans, func in itertools.product(SIDES, UPLOS, TRANS, FUNCS):
        A, B0 = rp.make_inputs(n, m, side, uplo, np.float64, seed=1000 + n + m)
        got = run_gpu(nla, side, uplo, trans, 1.0, func, A, B0)
        blas = rp.blas_reference(side, uplo, trans, 1.0, func, A, B0)
        want = c_port.unified_rectrxm(side, uplo, trans, 1.0, func, A, B0.copy(order="F"))
        assert rel(got, blas) < 1e-14, (side, uplo, trans, func)
        assert rel(got, want) < 1e-14, (side, uplo, trans, func)


@pytest.mark.parametrize("n,m", [(300, 40), (512, 256), (1000, 72), (1024, 1024), (1536, 130), (2048, 512)])
@pytest.mark.parametrize("alpha", [1.0, -0.5])
def test_recursive_fp64_all_variants(nla, gpu, n, m, alpha):
    """n > leaf: recursion + GEMM updates (tensor-core path when TMA-eligible, generic path otherwise),
    non power-of-two n, alpha != 1 -- everything the reference's tests leave unpinned."""
    for side, uplo, trans, func in itertools.product(SIDES, UPLOS, "NT", FUNCS):
        A, B0 = rp.make_inputs(n, m, side, uplo, np.float64, seed=n + m)
        got = run_gpu(nla, side, uplo, trans, alpha, func, A, B0)
        want = c_port.unified_rectrxm(side, uplo, trans, alpha, func, A, B0.copy(order="F"))
        blas = rp.blas_reference(side, uplo, trans, alpha, func, A, B0)
        err = rp.error_metric(side, uplo, trans, alpha, func, A, B0, got)
        assert err < 1e-13, (side, uplo, trans, func, err)
        assert rel(got, blas) < 1e-13, (side, uplo, trans, func)
        assert rel(got, want) < 1e-13, (side, uplo, trans, func)


@pytest.mark.parametrize("n,m,pad", [(257, 33, 0), (640, 200, 3), (1024, 512, 1), (1024, 512, 2)])
def test_unaligned_and_padded_ld_fp64(nla, gpu, n, m, pad):
    """Odd sizes / odd leading dimensions force the generic GEMM path; an even pad keeps the TMA path with ld > rows."""
    for side, uplo, trans, func in itertools.product(SIDES, UPLOS, "NT", FUNCS):
        A, B0 = rp.make_inputs(n, m, side, uplo, np.float64, seed=5 * n + m)
        got = run_gpu(nla, side, uplo, trans, 2.0, func, A, B0, ld_pad=pad)
        blas = rp.blas_reference(side, uplo, trans, 2.0, func, A, B0)
        assert rel(got, blas) < 1e-13, (side, uplo, trans, func)


def test_force_simt_matches_tensor_path(nla, gpu):
    n, m = 1024, 384
    A, B0 = rp.make_inputs(n, m, "L", "L", np.float64, seed=3, recipe="scaled")
    a = run_gpu(nla, "L", "L", "N", 1.0, "S", A, B0)
    gpu.set_option("force_simt", 1)
    try:
        b = run_gpu(nla, "L", "L", "N", 1.0, "S", A, B0)
    finally:
        gpu.set_option("force_simt", 0)
    assert rel(a, b) < 1e-13


@pytest.mark.parametrize("streams", [2, 4])
def test_concurrent_rhs_slabs(nla, gpu, streams):
    n, m = 1024, 1100
    gpu.set_option("streams", streams)
    try:
        for side, func in itertools.product(SIDES, FUNCS):
            A, B0 = rp.make_inputs(n, m, side, "L", np.float64, seed=11)
            got = run_gpu(nla, side, "L", "N", 1.5, func, A, B0)
            blas = rp.blas_reference(side, "L", "N", 1.5, func, A, B0)
            assert rel(got, blas) < 1e-13
    finally:
        gpu.set_option("streams", 0)


@pytest.mark.parametrize("macro", [0, 128, 256, 512, 4096])
def test_fused_slab_cutoff_option(nla, gpu, macro):
    """`macro` = order of the diagonal blocks handled by the fused slab kernel (0 = recursion down to the 128 leaf;
    >= n = the whole problem in one launch).  Ragged n (multiple of 8 but not of 128) and ragged m included."""
    gpu.set_option("macro", macro)
    try:
        for n, m in [(1024, 256), (1160, 200), (2048 + 72, 136)]:
            for uplo, trans, func in itertools.product(UPLOS, "NT", FUNCS):
                A, B0 = rp.make_inputs(n, m, "L", uplo, np.float64, seed=n + macro)
                got = run_gpu(nla, "L", uplo, trans, -1.5, func, A, B0)
                blas = rp.blas_reference("L", uplo, trans, -1.5, func, A, B0)
                err = rp.error_metric("L", uplo, trans, -1.5, func, A, B0, got)
                assert rel(got, blas) < 1e-13 and err < 1e-14, (n, m, uplo, trans, func, rel(got, blas), err)
    finally:
        gpu.set_option("macro", -1)


@pytest.mark.parametrize("leaf", [16, 32, 64, 128])
def test_leaf_cutoff_option(nla, gpu, leaf):
    gpu.set_option("leaf", leaf)
    try:
        for side, uplo, func in itertools.product(SIDES, UPLOS, FUNCS):
            A, B0 = rp.make_inputs(520, 96, side, uplo, np.float64, seed=leaf)
            got = run_gpu(nla, side, uplo, "N", 1.0, func, A, B0)
            blas = rp.blas_reference(side, uplo, "N", 1.0, func, A, B0)
            assert rel(got, blas) < 1e-13
    finally:
        gpu.set_option("leaf", 0)


def test_opposite_triangle_never_read(nla, gpu):
    """Fill the unreferenced triangle of A with NaN: the result must be unchanged (only `uplo` is read)."""
    n, m = 640, 136
    for side, uplo, trans, func in itertools.product(SIDES, UPLOS, "NT", FUNCS):
        A, B0 = rp.make_inputs(n, m, side, uplo, np.float64, seed=17)
        An = A.copy(order="F")
        idx = np.triu_indices(n, 1) if uplo == "L" else np.tril_indices(n, -1)
        An[idx] = np.nan
        got = run_gpu(nla, side, uplo, trans, 1.0, func, An, B0)
        blas = rp.blas_reference(side, uplo, trans, 1.0, func, A, B0)
        assert np.isfinite(got).all() and rel(got, blas) < 1e-13, (side, uplo, trans, func)


@pytest.mark.parametrize("dtype,tol", [(np.float32, 1e-5), (np.float16, 1e-2)])
@pytest.mark.parametrize("n,m", [(16, 8), (128, 64), (256, 64), (700, 96), (1024, 256)])
def test_low_precision(nla, gpu, dtype, tol, n, m):
    """Float32 (reference tolerance 1e-5, test/trsm.jl:8) and Float16 (1e-2 backward error, FP64 truth on the rounded inputs)."""
    recipe = "scaled" if dtype == np.float16 else "reference"
    for side, uplo, trans, func in itertools.product(SIDES, UPLOS, "NT", FUNCS):
        A, B0 = rp.make_inputs(n, m, side, uplo, dtype, seed=n + 3 * m, recipe=recipe)
        got = run_gpu(nla, side, uplo, trans, 1.0, func, A, B0)
        err = rp.error_metric(side, uplo, trans, 1.0, func, A, B0, got)
        truth = rp.blas_reference(side, uplo, trans, 1.0, func, A, B0)
        assert err < tol, (side, uplo, trans, func, err)
        assert rel(got, truth) < (1e-5 if dtype == np.float32 else 1e-2), (side, uplo, trans, func)


def test_leaf_entry_points_fp32(nla, gpu):
    """test/trsm.jl:10-64: the four TRSM leaves in Float32, n in {16,32,128}, m in {1,8,64}, tolerance 1e-5 vs BLAS;
    plus the four TRMM leaves (untested upstream)."""
    import torch

    fns = {("L", "L"): (nla.LeftLowerTRSM, nla.LeftLowerTRMM), ("L", "U"): (nla.LeftUpperTRSM, nla.LeftUpperTRMM),
           ("R", "L"): (nla.RightLowerTRSM, nla.RightLowerTRMM), ("R", "U"): (nla.RightUpperTRSM, nla.RightUpperTRMM)}
    for n, m in itertools.product([16, 32, 128], [1, 8, 64]):
        for (side, uplo), (fs, fm) in fns.items():
            A, B0 = rp.make_inputs(n, m, side, uplo, np.float32, seed=n * m)
            for f, func in ((fs, "S"), (fm, "M")):
                dA, dB = nla.colmajor(A), nla.colmajor(B0)
                f(dA, dB)
                torch.cuda.synchronize()
                blas = rp.blas_reference(side, uplo, "N", 1.0, func, A, B0)
                assert rel(nla.to_numpy(dB), blas) < 1e-5, (n, m, side, uplo, func)


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 1e-5), (np.float16, 1e-2)])
def test_leaf_entry_points_every_dtype_up_to_the_reference_cap(nla, gpu, dtype, tol):
    """The reference's leaves accept diagonal blocks up to 1024 (src/trsm.jl:9-11); so do the entry points here, in all three element types:
    n <= 128 is one launch of the leaf kernel, larger blocks go through the blocked path.  Against the oracle's leaf and OpenBLAS;
    n = 1025 is rejected."""
    import torch

    fns = {("L", "L"): (nla.LeftLowerTRSM, nla.LeftLowerTRMM), ("L", "U"): (nla.LeftUpperTRSM, nla.LeftUpperTRMM),
           ("R", "L"): (nla.RightLowerTRSM, nla.RightLowerTRMM), ("R", "U"): (nla.RightUpperTRSM, nla.RightUpperTRMM)}
    assert nla.load_library().nla_leaf_max(0) == 1024
    for n, m in [(16, 5), (128, 64), (200, 48), (512, 130), (1024, 96)]:
        for (side, uplo), (fs, fm) in fns.items():
            A, B0 = rp.make_inputs(n, m, side, uplo, dtype, seed=n + m)
            for f, func in ((fs, "S"), (fm, "M")):
                dA, dB = nla.colmajor(A), nla.colmajor(B0)
                gpu.launch_count(reset=True)
                f(dA, dB)
                torch.cuda.synchronize()
                if n <= 128:
                    assert gpu.launch_count() == 1
                got = nla.to_numpy(dB)
                if dtype == np.float64:
                    assert rel(got, rp.blas_reference(side, uplo, "N", 1.0, func, A, B0)) < tol, (n, m, side, uplo, func)
                else:
                    assert rp.error_metric(side, uplo, "N", 1.0, func, A, B0, got) < tol, (n, m, side, uplo, func)
    A, B0 = rp.make_inputs(1025, 4, "L", "L", dtype, seed=1)
    with pytest.raises(nla.NextLAError):
        nla.LeftLowerTRSM(nla.colmajor(A), nla.colmajor(B0))


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-14), (np.float32, 2e-5)])
def test_gemm_add_sub(nla, gpu, dtype, tol):
    """GEMM_ADD!(A,B,C): C += A*B and GEMM_SUB!(A,B,C): A -= B*C (src/matmul.jl:69-81), incl. transposed operands.
    All-positive inputs are the worst case for the truncating accumulation of the tensor cores (one direction): the Float32
    bound here is K * 2^-24-ish per 512-long chunk; the path's own criterion is the 1e-5 backward error of test_low_precision."""
    import torch

    rng = np.random.RandomState(0)
    for (M, N, K) in [(128, 128, 128), (256, 384, 512), (100, 36, 77), (1024, 640, 256)]:
        A = np.asfortranarray(rng.rand(M, K).astype(dtype)); B = np.asfortranarray(rng.rand(K, N).astype(dtype))
        C = np.asfortranarray(rng.rand(M, N).astype(dtype))
        dA, dB, dC = nla.colmajor(A), nla.colmajor(B), nla.colmajor(C)
        nla.GEMM_ADD(dA, dB, dC); torch.cuda.synchronize()
        want = C.astype(np.float64) + A.astype(np.float64) @ B.astype(np.float64)
        assert rel(nla.to_numpy(dC), want) < tol
        dC = nla.colmajor(C)
        nla.GEMM_SUB(dC, dA, dB); torch.cuda.synchronize()
        want = C.astype(np.float64) - A.astype(np.float64) @ B.astype(np.float64)
        assert rel(nla.to_numpy(dC), want) < tol
        At = np.asfortranarray(A.T.copy()); Bt = np.asfortranarray(B.T.copy())
        dC = nla.colmajor(C)
        nla.GEMM_ADD(nla.colmajor(At), dB, dC, transa="T"); torch.cuda.synchronize()
        want = C.astype(np.float64) + A.astype(np.float64) @ B.astype(np.float64)
        assert rel(nla.to_numpy(dC), want) < tol
        dC = nla.colmajor(C)
        nla.GEMM_ADD(dA, nla.colmajor(Bt), dC, transb="T"); torch.cuda.synchronize()
        assert rel(nla.to_numpy(dC), want) < tol


def test_edge_cases_and_errors(nla, gpu):
    import torch

    # empty problems are quick returns
    A = nla.colmajor(np.zeros((0, 0))); B = nla.colmajor(np.zeros((0, 5)))
    nla.unified_rectrxm("L", "L", "N", 1.0, "S", A, B)
    A1, B1 = rp.make_inputs(1, 3, "L", "L", np.float64, seed=1)
    got = run_gpu(nla, "L", "L", "N", 3.0, "S", A1, B1)
    assert np.allclose(got, 3.0 * B1 / A1[0, 0])
    A, B0 = rp.make_inputs(64, 8, "L", "L", np.float64, seed=2)
    dA, dB = nla.colmajor(A), nla.colmajor(B0)
    for bad in [("X", "L", "N", "S"), ("L", "X", "N", "S"), ("L", "L", "X", "S"), ("L", "L", "N", "X")]:
        with pytest.raises(nla.NextLAError):
            nla.unified_rectrxm(bad[0], bad[1], bad[2], 1.0, bad[3], dA, dB)
    with pytest.raises(nla.NextLAError):
        nla.unified_rectrxm("L", "L", "N", 1.0, "S", dA, nla.colmajor(np.zeros((32, 8))))
    with pytest.raises(nla.NextLAError):
        nla.unified_rectrxm("L", "L", "N", 1.0, "S", torch.zeros(4, 4, dtype=torch.float64), torch.zeros(4, 4, dtype=torch.float64))


def test_host_buffer_entry_point(nla, gpu):
    n, m = 768, 200
    for side, func in itertools.product(SIDES, FUNCS):
        A, B0 = rp.make_inputs(n, m, side, "U", np.float64, seed=23)
        B = B0.copy(order="F")
        nla.unified_rectrxm_host(side, "U", "T", 0.75, func, A, B)
        blas = rp.blas_reference(side, "U", "T", 0.75, func, A, B0)
        assert rel(B, blas) < 1e-13


@pytest.mark.parametrize("side,uplo,trans,func", [("L", "L", "N", "S"), ("L", "L", "T", "S"), ("L", "U", "N", "M"), ("R", "L", "N", "S"), ("R", "U", "T", "M")])
def test_host_buffer_pipeline_multi_slab(nla, gpu, side, uplo, trans, func):
    """Large enough for several RHS slabs and several column chunks of A (copy-in / solve / copy-out overlapped)."""
    n, m = 4096, 6000
    A, B0 = rp.make_inputs(n, m, side, uplo, np.float64, seed=31, recipe="scaled")
    blas = rp.blas_reference(side, uplo, trans, 1.25, func, A, B0)
    for slabs in (0, 3):   # automatic (one slab at this m) and three RHS slabs on concurrent compute streams, ragged last slab
        gpu.set_option("host_slabs", slabs)
        try:
            B = B0.copy(order="F")
            nla.unified_rectrxm_host(side, uplo, trans, 1.25, func, A, B, handle=gpu)
        finally:
            gpu.set_option("host_slabs", 0)
        assert rel(B, blas) < 1e-13, slabs
        assert rp.error_metric(side, uplo, trans, 1.25, func, A, B0, B) < 1e-14, slabs


def _gpu_backward_error(torch, side, uplo, trans, alpha, func, dA, dB0, dX):
    T = torch.tril(dA) if uplo == "L" else torch.triu(dA)
    if trans != "N":
        T = T.t()
    nA = torch.linalg.norm(T)
    if func == "S":
        R = (T @ dX if side == "L" else dX @ T) - alpha * dB0
        return (torch.linalg.norm(R) / (nA * torch.linalg.norm(dX) + abs(alpha) * torch.linalg.norm(dB0))).item()
    Pm = alpha * (T @ dB0 if side == "L" else dB0 @ T)
    return (torch.linalg.norm(dX - Pm) / (abs(alpha) * nA * torch.linalg.norm(dB0))).item()


@pytest.mark.parametrize("side,uplo,trans,func", [("L", "L", "N", "S"), ("L", "U", "T", "M"), ("R", "L", "N", "S"), ("R", "U", "T", "M"), ("L", "L", "N", "M")])
def test_large_size_properties_fp64(nla, gpu, side, uplo, trans, func):
    """At a BASELINE-scale size (n = m = 8192; the oracle would take minutes) use size-independent properties:
    backward error (computed on the GPU in FP64 by an independent cuBLAS product), linearity in the right-hand
    sides, and the solve/multiply round trip  M(S(B)) == B."""
    import torch

    n = m = 8192
    g = torch.Generator(device="cuda").manual_seed(1234)
    A = (2 * torch.rand(n, n, dtype=torch.float64, device="cuda", generator=g) - 1) / n ** 0.5
    A = (torch.tril(A, -1) if uplo == "L" else torch.triu(A, 1)) + torch.diag(1 + torch.rand(n, dtype=torch.float64, device="cuda", generator=g))
    dA = A.t().contiguous().t()
    shape = (n, m)
    B0 = (torch.rand(shape, dtype=torch.float64, device="cuda", generator=g) + 1).t().contiguous().t()
    X = B0.clone(memory_format=torch.preserve_format)
    assert X.stride(0) == 1
    nla.unified_rectrxm(side, uplo, trans, 1.25, func, dA, X)
    err = _gpu_backward_error(torch, side, uplo, trans, 1.25, func, dA, B0, X)
    assert err < 1e-13, err
    # linearity: op(2*B) == 2*op(B) exactly (power-of-two scaling commutes with rounding)
    X2 = (2 * B0).t().contiguous().t()
    nla.unified_rectrxm(side, uplo, trans, 1.25, func, dA, X2)
    assert torch.equal(X2, 2 * X)
    # round trip with the inverse operation
    inv = "M" if func == "S" else "S"
    nla.unified_rectrxm(side, uplo, trans, 1 / 1.25, inv, dA, X)
    assert (torch.linalg.norm(X - B0) / torch.linalg.norm(B0)).item() < 1e-11


# ---- Float32 / Float16 tensor-core path (tcgen05 GEMM + prepared diagonal blocks, csrc/gemm_tc.cuh, diag_prep.cuh) ----
TOL = {np.float32: 1e-5, np.float16: 1e-2}   # north_star tolerances (test/trsm.jl:8 for Float32)


@pytest.mark.parametrize("dtype", [np.float32, np.float16])
@pytest.mark.parametrize("tc_bn", [0, 128, 256])
def test_tensor_core_gemm_variants(nla, gpu, dtype, tc_bn):
    """GEMM_ADD!/GEMM_SUB! (src/matmul.jl:69-81) on the tcgen05 kernels: every operand-majorness instantiation (NN, TN, NT),
    both N tiles, ragged M/N/K, against an FP64 product of the same (rounded) inputs."""
    import torch

    rng = np.random.RandomState(5)
    gpu.set_option("tc_bn", tc_bn)
    try:
        for (M, N, K) in [(128, 256, 64), (256, 512, 256), (200, 300, 96), (1024, 640, 512), (136, 72, 1000)]:
            A = (rng.rand(M, K) - 0.5).astype(dtype); B = (rng.rand(K, N) - 0.5).astype(dtype); C = rng.rand(M, N).astype(dtype)
            want = {+1: C.astype(np.float64) + A.astype(np.float64) @ B.astype(np.float64),
                    -1: C.astype(np.float64) - A.astype(np.float64) @ B.astype(np.float64)}
            tol = 2e-5 if dtype == np.float32 else 1e-3
            for ta, tb in (("N", "N"), ("T", "N"), ("N", "T"), ("T", "T")):
                Ain = np.asfortranarray(A.T.copy() if ta == "T" else A); Bin = np.asfortranarray(B.T.copy() if tb == "T" else B)
                dA, dB = nla.colmajor(Ain), nla.colmajor(Bin)
                dC = nla.colmajor(np.asfortranarray(C))
                gpu.launch_count(reset=True)
                nla.GEMM_ADD(dA, dB, dC, transa=ta, transb=tb); torch.cuda.synchronize()
                assert gpu.launch_count() == 1
                assert rel(nla.to_numpy(dC), want[+1]) < tol, (M, N, K, ta, tb)
            dC = nla.colmajor(np.asfortranarray(C))
            nla.GEMM_SUB(dC, nla.colmajor(np.asfortranarray(A)), nla.colmajor(np.asfortranarray(B))); torch.cuda.synchronize()
            assert rel(nla.to_numpy(dC), want[-1]) < tol, (M, N, K)
    finally:
        gpu.set_option("tc_bn", 0)


@pytest.mark.parametrize("dtype", [np.float32, np.float16])
@pytest.mark.parametrize("cg", [1, 2])
def test_cta_pair_kernel(nla, gpu, dtype, cg):
    """The cta_group::2 kernel (csrc/gemm_tc2.cuh: a pair of CTAs shares one M = 256 tcgen05.mma) against the single-CTA kernel
    and FP64 truth: updates with odd tile counts (a pair whose second CTA is past the last M tile), ragged N and K, every
    majorness instantiation, and a full solve / multiply with the pairs forced on."""
    import torch

    rng = np.random.RandomState(11)
    gpu.set_option("tc_cg", cg)
    gpu.set_option("tc_persist", 0)   # Float16 would otherwise take the persistent pair kernel (test_persistent_pair_kernel)
    try:
        for (M, N, K) in [(256, 256, 128), (384, 520, 320), (1000, 1544, 1096), (2048, 4096, 2048)]:
            A = (rng.rand(M, K) - 0.5).astype(dtype); B = (rng.rand(K, N) - 0.5).astype(dtype); C = rng.rand(M, N).astype(dtype)
            want = C.astype(np.float64) - A.astype(np.float64) @ B.astype(np.float64)
            for ta, tb in (("N", "N"), ("T", "N"), ("N", "T")):
                Ain = np.asfortranarray(A.T.copy() if ta == "T" else A); Bin = np.asfortranarray(B.T.copy() if tb == "T" else B)
                dC = nla.colmajor(np.asfortranarray(C))
                nla._gemm(dC, nla.colmajor(Ain), nla.colmajor(Bin), -1, transa=ta, transb=tb); torch.cuda.synchronize()
                assert rel(nla.to_numpy(dC), want) < (2e-5 if dtype == np.float32 else 1e-3), (M, N, K, ta, tb)
        n, m = 1160, 520
        for side, uplo, trans, func in itertools.product(SIDES, UPLOS, "NT", FUNCS):
            A, B0 = rp.make_inputs(n, m, side, uplo, dtype, seed=5, recipe="scaled")
            got = run_gpu(nla, side, uplo, trans, 1.5, func, A, B0)
            assert rp.error_metric(side, uplo, trans, 1.5, func, A, B0, got) < TOL[dtype], (side, uplo, trans, func)
    finally:
        gpu.set_option("tc_cg", 0)
        gpu.set_option("tc_persist", 1)


@pytest.mark.parametrize("dtype", [np.float32, np.float16])
def test_tensor_core_path_matches_simt_path(nla, gpu, dtype):
    """The tcgen05 path and the generic strided kernels (option force_simt) are two independent implementations of the same
    schedule: they must agree to rounding, for every side/uplo/trans/func, with alpha != 1 and a ragged order."""
    n, m = 904, 328
    for side, uplo, trans, func in itertools.product(SIDES, UPLOS, "NT", FUNCS):
        A, B0 = rp.make_inputs(n, m, side, uplo, dtype, seed=77, recipe="scaled")
        got = run_gpu(nla, side, uplo, trans, -0.5, func, A, B0)
        gpu.set_option("force_simt", 1)
        try:
            ref = run_gpu(nla, side, uplo, trans, -0.5, func, A, B0)
        finally:
            gpu.set_option("force_simt", 0)
        assert rel(got, ref) < (5e-6 if dtype == np.float32 else 5e-3), (side, uplo, trans, func)
        assert rp.error_metric(side, uplo, trans, -0.5, func, A, B0, got) < TOL[dtype]
        if func == "M":   # the in-place recursion (reference order) against the default batched out-of-place schedule
            gpu.set_option("trmm_batched", 0)
            try:
                rec = run_gpu(nla, side, uplo, trans, -0.5, func, A, B0)
            finally:
                gpu.set_option("trmm_batched", 1)
            assert rel(got, rec) < (5e-6 if dtype == np.float32 else 5e-3), (side, uplo, trans, func)


@pytest.mark.parametrize("dtype", [np.float32, np.float16])
def test_tensor_core_opposite_triangle_and_fallback(nla, gpu, dtype):
    """NaN in the unreferenced triangle must never reach the result (TMA boxes of neighbouring tiles do touch it); a leading
    dimension that breaks the 16-byte TMA pitch rule silently takes the generic kernels."""
    n, m = 640, 264
    for side, uplo, func in itertools.product(SIDES, UPLOS, FUNCS):
        A, B0 = rp.make_inputs(n, m, side, uplo, dtype, seed=9, recipe="scaled")
        An = A.copy()
        An[np.triu_indices(n, 1) if uplo == "L" else np.tril_indices(n, -1)] = np.nan
        got = run_gpu(nla, side, uplo, "N", 1.0, func, An, B0)
        assert np.isfinite(got).all()
        assert rp.error_metric(side, uplo, "N", 1.0, func, A, B0, got) < TOL[dtype], (side, uplo, func)
        got2 = run_gpu(nla, side, uplo, "N", 1.0, func, A, B0, ld_pad=1)   # odd pitch -> generic path
        assert rp.error_metric(side, uplo, "N", 1.0, func, A, B0, got2) < TOL[dtype], (side, uplo, func)


@pytest.mark.parametrize("dtype", [np.float32, np.float16])
def test_tensor_core_host_buffer_entry_point(nla, gpu, dtype):
    n, m = 1408, 520
    for side, uplo, trans, func in [("L", "L", "N", "S"), ("L", "U", "T", "M"), ("R", "L", "N", "S"), ("R", "U", "N", "M")]:
        A, B0 = rp.make_inputs(n, m, side, uplo, dtype, seed=41, recipe="scaled")
        B = B0.copy(order="F")
        nla.unified_rectrxm_host(side, uplo, trans, 1.0, func, A, B)
        assert rp.error_metric(side, uplo, trans, 1.0, func, A, B0, B) < TOL[dtype], (side, uplo, trans, func)


@pytest.mark.parametrize("dtype_name,side,uplo,trans,func", [("float32", "L", "U", "T", "M"), ("float32", "L", "L", "N", "S"), ("float16", "R", "L", "N", "S"),
                                                             ("float16", "L", "L", "N", "S"), ("float16", "R", "U", "T", "M")])
def test_large_size_properties_low_precision(nla, gpu, dtype_name, side, uplo, trans, func):
    """BASELINE configs 3 (Float32 left/upper/transposed TRMM) and 4 (Float16 right/lower TRSM) at n = 8192 with 8192 right-hand
    sides: backward error in FP64 on the GPU (independent cuBLAS product), exact linearity under power-of-two scaling of the
    right-hand sides, concurrent RHS slabs == single stream bit for bit."""
    import torch

    dt = getattr(torch, dtype_name)
    tol = 1e-5 if dt == torch.float32 else 1e-2
    n = m = 8192
    g = torch.Generator(device="cuda").manual_seed(4321)
    A = (2 * torch.rand(n, n, dtype=torch.float32, device="cuda", generator=g) - 1) / n ** 0.5
    A = (torch.tril(A, -1) if uplo == "L" else torch.triu(A, 1)) + torch.diag(1 + torch.rand(n, dtype=torch.float32, device="cuda", generator=g))
    dA = A.to(dt).t().contiguous().t()
    B0 = (torch.rand((n, m), dtype=torch.float32, device="cuda", generator=g) + 1).to(dt).t().contiguous().t()
    X = B0.clone(memory_format=torch.preserve_format)
    gpu.set_option("streams", 1)
    nla.unified_rectrxm(side, uplo, trans, 1.0, func, dA, X)
    err = _gpu_backward_error(torch, side, uplo, trans, 1.0, func, dA.double(), B0.double(), X.double())
    assert err < tol, err
    X2 = (2 * B0).t().contiguous().t()
    nla.unified_rectrxm(side, uplo, trans, 1.0, func, dA, X2)
    if dt == torch.float32:
        assert torch.equal(X2, 2 * X)
    else:
        # Float16 intermediates may be subnormal (|x| < 6.1e-5), where doubling does not commute with rounding; such a
        # one-subnormal-ulp difference can flip the rounding of a later entry by one ulp.  Measured on B200: <= 138 of 6.7e7
        # entries differ (probes/tc_determinism.py); reruns of the same input are bit-identical.
        assert (X2 != 2 * X).float().mean().item() < 1e-4
        assert (torch.linalg.norm(X2.float() - 2 * X.float()) / torch.linalg.norm(X2.float())).item() < 1e-6
    gpu.set_option("streams", 4)
    try:
        X3 = B0.clone(memory_format=torch.preserve_format)
        nla.unified_rectrxm(side, uplo, trans, 1.0, func, dA, X3)
        torch.cuda.synchronize()
        assert torch.equal(X3, X)
    finally:
        gpu.set_option("streams", 0)


@pytest.mark.parametrize("dtype", [np.float32, np.float16])
@pytest.mark.parametrize("ib", [128, 256, 512, 1024, 2048])
def test_block_inverse_leaves(nla, gpu, dtype, ib):
    """Float32 / Float16 solves with diagonal blocks inverted up to order `inv_block` (tri_inv.cuh: FP64 128-blocks, doubling in
    the next wider type, one triangular GEMM per block): every solve variant, ragged order, alpha != 1, both input recipes,
    against the tolerance and against the 128-wide leaves."""
    n, m = 1500, 200
    try:
        for recipe in (("scaled", "reference") if dtype == np.float32 else ("scaled",)):   # the reference recipe underflows Float16
            for side, uplo, trans in itertools.product(SIDES, UPLOS, "NT"):
                A, B0 = rp.make_inputs(n, m, side, uplo, dtype, seed=ib + 5, recipe=recipe)
                gpu.set_option("inv_block", ib)
                got = run_gpu(nla, side, uplo, trans, 1.5, "S", A, B0)
                err = rp.error_metric(side, uplo, trans, 1.5, "S", A, B0, got)
                assert err < TOL[dtype], (recipe, side, uplo, trans, err)
                gpu.set_option("inv_block", 128)
                base = run_gpu(nla, side, uplo, trans, 1.5, "S", A, B0)
                assert rel(got, base) < (2e-5 if dtype == np.float32 else 1e-2), (recipe, side, uplo, trans)
    finally:
        gpu.set_option("inv_block", 0)


@pytest.mark.parametrize("dtype,pc", [(np.float64, 512), (np.float16, 512), (np.float16, 1024), (np.float32, 1024)])
@pytest.mark.parametrize("side,uplo,trans,func", [("L", "L", "N", "S"), ("L", "U", "N", "S"), ("R", "L", "T", "M"), ("L", "L", "T", "S"), ("R", "U", "N", "S")])
def test_gated_arrival_of_A(nla, gpu, dtype, pc, side, uplo, trans, func):
    """nla_rectrxm_gated: A becomes valid panel by panel (here: copied on a side stream from a pristine matrix, each panel behind a
    spin kernel so that the solve really has to wait), result identical to the plain call."""
    import torch

    # (panels that are a multiple of the block-inverse order let the Float32/Float16 solves prepare each panel's diagonal blocks as it
    #  arrives; otherwise, and for multiplies, they wait for all of A first)
    n, m = 3072 if pc == 1024 else 2048, 384
    A, B0 = rp.make_inputs(n, m, side, uplo, dtype, seed=31, recipe="scaled")
    dA_full = nla.colmajor(A)
    want = run_gpu(nla, side, uplo, trans, 1.25, func, A, B0)
    dA = torch.full_like(dA_full, float("nan"))   # nothing valid until its panel has arrived
    dB = nla.colmajor(B0)
    order = nla.panel_order(side, uplo, trans, func, n, pc)
    side_stream = torch.cuda.Stream()
    events = [torch.cuda.Event() for _ in range(n // pc)]
    torch.cuda.synchronize()
    with torch.cuda.stream(side_stream):
        for p in order:
            torch.cuda._sleep(20_000_000)   # ~10 ms per panel
            dA[:, p * pc:(p + 1) * pc].copy_(dA_full[:, p * pc:(p + 1) * pc])
            events[p].record(side_stream)
    nla.unified_rectrxm_gated(side, uplo, trans, 1.25, func, dA, dB, pc, events)
    torch.cuda.synchronize()
    got = nla.to_numpy(dB)
    assert np.isfinite(got).all()
    assert np.array_equal(got, want), (side, uplo, trans, func)


@pytest.mark.parametrize("side,uplo,trans,func", [("L", "L", "N", "S"), ("R", "U", "N", "M")])
def test_pipelined_host_path_single_rank(nla, gpu, side, uplo, trans, func):
    """sharded.unified_rectrxm_pipelined_host with one rank (no process group): pinned host A and B in, B out, A uploaded panel by
    panel, B streamed by the library's host pipeline gated on the panels of A."""
    import torch
    from importlib import import_module

    sh = import_module(nla.__name__ + ".sharded")
    n, m = 2048, 1100
    A, B0 = rp.make_inputs(n, m, side, uplo, np.float64, seed=12, recipe="scaled")
    want = run_gpu(nla, side, uplo, trans, 0.75, func, A, B0)
    hA = torch.from_numpy(np.ascontiguousarray(A.T)).pin_memory().t()
    hB = torch.from_numpy(np.ascontiguousarray(B0.T)).pin_memory().t()
    dA = torch.full((n, n), float("nan"), dtype=torch.float64, device="cuda").t()
    sh.unified_rectrxm_pipelined_host(side, uplo, trans, 0.75, func, dA, hA, hB, panels=4)
    torch.cuda.synchronize()
    got = np.asfortranarray(hB.numpy())
    assert rel(got, want) < 1e-13   # same schedule; the host pipeline cuts the large updates into 1024-wide pieces


def test_persistent_pair_kernel(nla, gpu):
    """The persistent CTA-pair kernel (csrc/gemm_tc3.cuh, Float16: static tile list per cluster, two TMEM accumulators, 8 drain
    warps) against FP64 truth and against the one-tile kernels: updates with a single tile, odd tile counts, more tiles than
    clusters (several tiles per cluster: ring and accumulator hand-over across tiles), ragged M/N/K, every majorness
    instantiation; then every solve / multiply variant with block-inverse leaves (windows 4 / 5, `dup` epilogue), ragged order."""
    import torch

    dtype = np.float16
    rng = np.random.RandomState(21)
    assert gpu.get_option("tc_persist") == 1
    for (M, N, K) in [(256, 256, 64), (384, 520, 320), (1000, 1544, 1096), (2304, 5000, 200), (4096, 8192, 1024)]:
        A = (rng.rand(M, K) - 0.5).astype(dtype); B = (rng.rand(K, N) - 0.5).astype(dtype); C = rng.rand(M, N).astype(dtype)
        want = C.astype(np.float64) - A.astype(np.float64) @ B.astype(np.float64)
        for ta, tb in (("N", "N"), ("T", "N"), ("N", "T")):
            Ain = np.asfortranarray(A.T.copy() if ta == "T" else A); Bin = np.asfortranarray(B.T.copy() if tb == "T" else B)
            dA, dB = nla.colmajor(Ain), nla.colmajor(Bin)
            dC = nla.colmajor(np.asfortranarray(C))
            nla._gemm(dC, dA, dB, -1, transa=ta, transb=tb); torch.cuda.synchronize()
            got = nla.to_numpy(dC)
            assert rel(got, want) < 1e-3, (M, N, K, ta, tb)
            gpu.set_option("tc_persist", 0)
            try:
                dC2 = nla.colmajor(np.asfortranarray(C))
                nla._gemm(dC2, dA, dB, -1, transa=ta, transb=tb); torch.cuda.synchronize()
            finally:
                gpu.set_option("tc_persist", 1)
            assert rel(got, nla.to_numpy(dC2)) < 1e-3, (M, N, K, ta, tb)
    for n, m in ((1500, 520), (2048, 1024), (3000, 300)):
        for side, uplo, trans, func in itertools.product(SIDES, UPLOS, "NT", FUNCS):
            A, B0 = rp.make_inputs(n, m, side, uplo, dtype, seed=n + 1, recipe="scaled")
            got = run_gpu(nla, side, uplo, trans, 1.5, func, A, B0)
            assert np.isfinite(got).all()
            assert rp.error_metric(side, uplo, trans, 1.5, func, A, B0, got) < TOL[dtype], (n, m, side, uplo, trans, func)


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.float16])
def test_unit_diagonal_trsm_trmm(nla, gpu, dtype):
    """trsm / trmm with diag = 'U' (the reference's wrappers accept the flag and ignore it, src/trsm.jl:186): the
    stored diagonal must not be read (it holds NaN here), and the result must equal the non-unit routine applied to the same
    matrix with an explicit unit diagonal, and OpenBLAS with diag = 'U' -- every side/uplo/trans, leaf, fused-slab / block-inverse
    and recursive sizes, alpha != 1."""
    from scipy.linalg import blas

    tol = {np.float64: 1e-13, np.float32: 1e-5, np.float16: 1e-2}[dtype]
    for n, m in ((96, 40), (640, 136), (2304, 264)):
        for side, uplo, trans, func in itertools.product(SIDES, UPLOS, "NT", FUNCS):
            A, B0 = rp.make_inputs(n, m, side, uplo, dtype, seed=n + 7, recipe="scaled")
            A1 = A.copy(); np.fill_diagonal(A1, 1)
            An = A.copy(); np.fill_diagonal(An, np.nan)
            dA, dB = nla.colmajor(An), nla.colmajor(B0)
            f = nla.trsm if func == "S" else nla.trmm
            f(side, uplo, trans, "U", dA, dB, -1.5)
            import torch
            torch.cuda.synchronize()
            got = nla.to_numpy(dB)
            assert np.isfinite(got).all(), (n, side, uplo, trans, func)
            assert rp.error_metric(side, uplo, trans, -1.5, func, A1, B0, got) < tol, (n, side, uplo, trans, func)
            want = run_gpu(nla, side, uplo, trans, -1.5, func, A1, B0)   # non-unit path on the explicit unit diagonal
            assert rel(got, want) < (1e-13 if dtype == np.float64 else tol), (n, side, uplo, trans, func)
            if dtype == np.float64:
                bf = blas.dtrsm if func == "S" else blas.dtrmm
                ref = bf(-1.5, A, B0, side=0 if side == "L" else 1, lower=1 if uplo == "L" else 0, trans_a=0 if trans == "N" else 1, diag=1)
                assert rel(got, ref) < 1e-13, (n, side, uplo, trans, func)
    with pytest.raises(nla.NextLAError):
        nla.trsm("L", "L", "N", "X", nla.colmajor(np.eye(4)), nla.colmajor(np.ones((4, 2))))


@pytest.mark.parametrize("m,n", [(1024, 1024), (1500, 900), (700, 1100)])
def test_recursive_lu_device_steps(nla, gpu, m, n):
    """the reference's recursive LU (getrf2!, src/lu.jl:185-299) with its laswp + TRSM('L','L','N','U') + GEMM steps
    (:274-280, :297) on the device (nla_laswp, nla_trxm diag = 'U', nla_gemm_update); only the <= 64-column panels are factored on the
    host (LAPACK getrf through SciPy).  Checks P*A = L*U against the original matrix and the pivots against LAPACK's."""
    import torch
    from scipy.linalg import lu_factor

    rng = np.random.RandomState(m + n)
    A0 = rng.rand(m, n) - 0.5
    dA = nla.colmajor(A0)
    k = min(m, n)
    ipiv = torch.zeros(k, dtype=torch.int64, device="cuda")

    def getrf2(Av, pv):
        mm, nn = Av.shape
        if nn <= 64 or mm <= 64:
            lu, piv = lu_factor(nla.to_numpy(Av), check_finite=False)
            Av.copy_(torch.from_numpy(np.ascontiguousarray(lu)).cuda())
            pv.copy_(torch.from_numpy(piv[:min(mm, nn)].astype(np.int64) + 1).cuda())   # 1-based like the reference
            return
        n1 = min(mm, nn) // 2                                    # src/lu.jl:255
        getrf2(Av[:, :n1], pv[:n1])                              # :264
        nla.getrf2_update(Av, n1, pv)                            # :274-280 on the device
        getrf2(Av[n1:, n1:], pv[n1:min(mm, nn)])                 # :284
        pv[n1:min(mm, nn)] += n1                                 # :291-293
        nla.laswp(Av[:, :n1], n1 + 1, min(mm, nn), pv, 1)        # :297

    getrf2(dA, ipiv)
    torch.cuda.synchronize()
    LU = nla.to_numpy(dA)
    piv = ipiv.cpu().numpy() - 1
    L = np.tril(LU[:, :k], -1) + np.eye(m, k)
    U = np.triu(LU[:k, :])
    PA = A0.copy()
    for i, p in enumerate(piv):
        if p != i:
            PA[[i, p]] = PA[[p, i]]
    assert np.linalg.norm(PA - L @ U) / np.linalg.norm(A0) < 1e-13
    lu_ref, piv_ref = lu_factor(A0, check_finite=False)
    assert np.array_equal(piv, piv_ref[:k])                      # same pivot sequence as LAPACK's (partial pivoting is unique here)
    assert rel(LU, lu_ref) < 1e-11


def test_right_side_via_left_fp64(nla, gpu):
    """Float64, side 'R': the default runs the equivalent left-side problem on a transposed copy of B (fused slab kernel); option
    right_via_left = 0 keeps the native right-side schedule.  Both against the oracle, every right-side variant, alpha != 1,
    ragged m, a padded leading dimension, unit diagonal."""
    n, m = 1536, 520
    for uplo, trans, func in itertools.product(UPLOS, "NT", FUNCS):
        A, B0 = rp.make_inputs(n, m, "R", uplo, np.float64, seed=91)
        want = c_port.unified_rectrxm("R", uplo, trans, -0.75, func, A, B0.copy(order="F"))
        assert gpu.get_option("right_via_left") == 1
        gpu.launch_count(reset=True)
        got = run_gpu(nla, "R", uplo, trans, -0.75, func, A, B0, ld_pad=2)
        launches_via_left = gpu.launch_count()
        gpu.set_option("right_via_left", 0)
        try:
            gpu.launch_count(reset=True)
            native = run_gpu(nla, "R", uplo, trans, -0.75, func, A, B0, ld_pad=2)
            launches_native = gpu.launch_count()
        finally:
            gpu.set_option("right_via_left", 1)
        assert launches_via_left < launches_native          # 2 transposes + slab schedule instead of 128-wide leaves
        assert rel(got, want) < 1e-13 and rel(native, want) < 1e-13, (uplo, trans, func)
        assert rp.error_metric("R", uplo, trans, -0.75, func, A, B0, got) < 1e-13


def test_options_round_trip(nla, gpu):
    """Every tunable documented in include/nextla_b200.h can be set and read back; unknown keys and out-of-range values are status
    codes, not crashes; defaults are restored."""
    import re

    header = open(nla.HEADER_PATH).read()
    keys = re.findall(r'^ \*   "([a-z0-9_]+)"', header, flags=re.M)
    assert {"leaf", "macro", "streams", "tc_bn", "tc_cg", "inv_block", "tc_persist", "right_via_left", "trmm_batched", "pdl", "profile"} <= set(keys)
    for k in keys:
        old = gpu.get_option(k)
        assert old >= 0 or (k in ("macro", "getrf_cluster") and old == -1), k   # -1 = automatic
        gpu.set_option(k, old)
        assert gpu.get_option(k) == old, k
    with pytest.raises(nla.NextLAError):
        gpu.set_option("no_such_option", 1)
    for k, bad in (("inv_block", 300), ("tc_bn", 64), ("tc_cg", 7), ("streams", 99), ("tc_persist", 5)):
        old = gpu.get_option(k)
        with pytest.raises(nla.NextLAError):
            gpu.set_option(k, bad)
        assert gpu.get_option(k) == old


def test_wide_pair_kernel(nla, gpu):
    """The 256 x 512 persistent pair kernel (csrc/gemm_tc4.cuh: two accumulators share the A tile, Float16, long updates) forced on
    from K = 64 (option tc_wide_k): updates with one tile, more tiles than clusters, ragged M/N/K and N that is not a multiple of 512,
    every majorness instantiation, against FP64 truth; then every solve / multiply variant so that the `dup` epilogue (the update that
    precedes a block-inverse leaf) and the edge path run through it too."""
    import torch

    dtype = np.float16
    rng = np.random.RandomState(33)
    assert gpu.get_option("tc_wide_k") == 4096
    gpu.set_option("tc_wide_k", 64)
    try:
        for (M, N, K) in [(256, 512, 64), (384, 1100, 320), (1000, 1544, 1096), (4096, 9000, 512), (512, 1024, 4160)]:
            A = (rng.rand(M, K) - 0.5).astype(dtype); B = (rng.rand(K, N) - 0.5).astype(dtype); C = rng.rand(M, N).astype(dtype)
            want = C.astype(np.float64) - A.astype(np.float64) @ B.astype(np.float64)
            for ta, tb in (("N", "N"), ("T", "N"), ("N", "T")):
                Ain = np.asfortranarray(A.T.copy() if ta == "T" else A); Bin = np.asfortranarray(B.T.copy() if tb == "T" else B)
                dC = nla.colmajor(np.asfortranarray(C))
                nla._gemm(dC, nla.colmajor(Ain), nla.colmajor(Bin), -1, transa=ta, transb=tb); torch.cuda.synchronize()
                assert rel(nla.to_numpy(dC), want) < 1e-3, (M, N, K, ta, tb)
        for n, m in ((1500, 520), (3000, 1030)):
            for side, uplo, trans, func in itertools.product(SIDES, UPLOS, "NT", FUNCS):
                A, B0 = rp.make_inputs(n, m, side, uplo, dtype, seed=n + 2, recipe="scaled")
                got = run_gpu(nla, side, uplo, trans, 1.5, func, A, B0)
                assert np.isfinite(got).all()
                assert rp.error_metric(side, uplo, trans, 1.5, func, A, B0, got) < TOL[dtype], (n, m, side, uplo, trans, func)
    finally:
        gpu.set_option("tc_wide_k", 4096)


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-13), (np.float32, 1e-5), (np.float16, 1e-2)])
@pytest.mark.parametrize("n,ib", [(96, 32), (1000, 256), (2048, 1024), (1544, 0), (3072, 1024), (200, 4096)])
def test_lauum_block_loop(nla, gpu, dtype, tol, n, ib):
    """lauum! (src/lauum.jl:52-186) through ONE C entry (nla_lauum): the reference's block loop with the off-diagonal
    steps on this library's trmm / GEMM kernels and the diagonal-block products through triangle-masked GEMM epilogues.  L^H L and
    U U^H against NumPy; the opposite triangle (NaN here) is neither used nor written; ib <= 0 (default 1024), ib > n (one block),
    ragged last block, a padded leading dimension; bad arguments are status codes (the reference throws ArgumentError, :54-60)."""
    import torch

    rng = np.random.RandomState(n + ib)
    F = (rng.rand(n, n) - 0.5).astype(dtype) / np.sqrt(n) + np.eye(n, dtype=dtype)
    for uplo in "LU":
        T = np.tril(F) if uplo == "L" else np.triu(F)
        want = (T.astype(np.float64).T @ T.astype(np.float64)) if uplo == "L" else (T.astype(np.float64) @ T.astype(np.float64).T)
        Ain = T.copy()
        Ain[np.triu_indices(n, 1) if uplo == "L" else np.tril_indices(n, -1)] = np.nan
        if n == 1544:   # padded leading dimension
            dA = nla.empty_colmajor(n, n, getattr(torch, np.dtype(dtype).name), ld=n + 8)
            dA.copy_(torch.from_numpy(np.ascontiguousarray(Ain)).cuda())
        else:
            dA = nla.colmajor(Ain)
        gpu.launch_count(reset=True)
        nla.lauum(uplo, dA, ib)
        torch.cuda.synchronize()
        assert gpu.launch_count() > 0
        got = nla.to_numpy(dA)
        mask = np.tril(np.ones((n, n), bool)) if uplo == "L" else np.triu(np.ones((n, n), bool))
        assert np.isnan(got[~mask]).all()                       # untouched
        assert np.isfinite(got[mask]).all()
        # Float32: the diagonal of L^H L is a large term plus many small same-sign ones -- the worst case for the tensor core's
        # truncating accumulation (~2^-24 per MMA, one direction), hence 3e-5 on the relative error here; the reference's
        # own criterion for lauum (test/lauum.jl:26: ||A - expected||_F / n < 1e-5 single, 1e-12 double) is checked as well
        rel_tol = 3e-5 if dtype == np.float32 else tol
        assert np.linalg.norm(got[mask] - want[mask]) / np.linalg.norm(want[mask]) < rel_tol, uplo
        assert np.linalg.norm(got[mask] - want[mask]) / n < {np.float64: 1e-12, np.float32: 1e-5, np.float16: 1e-2}[dtype], uplo
    with pytest.raises(nla.NextLAError):
        nla.lauum("X", nla.colmajor(np.eye(4, dtype=dtype)), 2)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_lauum_reference_grid(nla, gpu, dtype):
    """The reference's own lauum test (test/lauum.jl:1-29): n in {16, 32, 64, 128}, ib in {2, 4, 8}, both triangles, inputs
    0.5 + rand (upper) / -0.5 + rand (lower) with a ZERO opposite triangle, criterion ||A - expected||_F / n < rtol
    (1e-5 single, 1e-12 double, test/lapack_helpers.jl:14); and its error handling (:31-34)."""
    import torch

    rtol = 1e-12 if dtype == np.float64 else 1e-5
    rng = np.random.RandomState(7)
    for uplo, n, ib in itertools.product("UL", [16, 32, 64, 128], [2, 4, 8]):
        A0 = (np.triu(0.5 + rng.rand(n, n)) if uplo == "U" else np.tril(-0.5 + rng.rand(n, n))).astype(dtype)
        dA = nla.colmajor(A0)
        nla.lauum(uplo, dA, ib)
        torch.cuda.synchronize()
        got = nla.to_numpy(dA).astype(np.float64)
        A64 = A0.astype(np.float64)
        expected = np.triu(A64 @ A64.T) if uplo == "U" else np.tril(A64.T @ A64)
        tri = np.triu(got) if uplo == "U" else np.tril(got)
        assert np.linalg.norm(tri - expected) / n < rtol, (uplo, n, ib)
        other = np.tril(got, -1) if uplo == "U" else np.triu(got, 1)
        assert not other.any()                                   # the zero triangle stays zero
    with pytest.raises(nla.NextLAError):
        nla.lauum("X", nla.colmajor(np.zeros((4, 4), dtype=dtype)), 2)
    lib = nla.load_library()
    assert lib.nla_lauum(gpu._h, b"U", 0, -1, None, 1, 2, None) == 2   # negative n (the reference: ArgumentError)
