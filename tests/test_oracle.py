"""CPU tests of the oracle (test infrastructure): it must reproduce the reference's own acceptance criterion --
agreement with OpenBLAS trsm!/trmm! on the reference's grid within the reference's tolerance -- and the frozen golden
fixtures; the NumPy and the C/OpenMP restatements must agree bit for bit."""
import itertools

import numpy as np
import pytest

from oracle import c_port
from oracle import reference_port as rp
from tests.golden_util import load_cases


def rel(a, b):
    a = a.astype(np.float64); b = b.astype(np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("n", [16, 32, 128, 256])
def test_oracle_meets_reference_criterion_fp64(n):
    """test/unified_rectrxm.jl:10-44: n x m grid, all side/uplo/trans/func, alpha = 1, rel. error vs BLAS < 1e-14."""
    for m in (1, 8, 64):
        for side, uplo, trans, func in itertools.product("LR", "LU", "NTC", "SM"):
            A, B0 = rp.make_inputs(n, m, side, uplo, np.float64, seed=n + m)
            got = c_port.unified_rectrxm(side, uplo, trans, 1.0, func, A, B0.copy(order="F"))
            blas = rp.blas_reference(side, uplo, trans, 1.0, func, A, B0)
            assert rel(got, blas) < 1e-14, (n, m, side, uplo, trans, func)


def test_oracle_trsm_leaves_fp32():
    """test/trsm.jl:10-64: Float32 leaves (n <= 256 never recurses for 'S'), tolerance 1e-5 vs BLAS."""
    for n, m in itertools.product([16, 32, 128, 256], [1, 8, 64]):
        for side, uplo in itertools.product("LR", "LU"):
            A, B0 = rp.make_inputs(n, m, side, uplo, np.float32, seed=n * m)
            got = c_port.unified_rectrxm(side, uplo, "N", 1.0, "S", A, B0.copy(order="F"))
            blas = rp.blas_reference(side, uplo, "N", 1.0, "S", A, B0)
            assert rel(got, blas) < 1e-5


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.float16])
def test_numpy_and_c_restatements_bit_identical(dtype):
    for n, m in [(16, 1), (33, 8), (100, 5), (300, 12)]:
        for side, uplo, trans, func in itertools.product("LR", "LU", "NT", "SM"):
            A, B0 = rp.make_inputs(n, m, side, uplo, dtype, seed=n + m, recipe="scaled" if dtype == np.float16 else "reference")
            a = rp.unified_rectrxm(side, uplo, trans, 0.75, func, A, B0.copy(order="F"))
            b = c_port.unified_rectrxm(side, uplo, trans, 0.75, func, A, B0.copy(order="F"))
            assert np.array_equal(a, b), (n, m, side, uplo, trans, func)


def test_oracle_recursion_above_threshold():
    """n > 256 exercises the TRSM recursion + GEMM_SUB! (src/rectrxm.jl:164,170,184,190) that no reference test reaches;
    pinned by BLAS equivalence and backward error, incl. alpha != 1 and a non power-of-two n."""
    for n, m in [(300, 7), (700, 9), (1024, 16)]:
        for side, uplo, trans, func in itertools.product("LR", "LU", "NT", "SM"):
            A, B0 = rp.make_inputs(n, m, side, uplo, np.float64, seed=n)
            got = c_port.unified_rectrxm(side, uplo, trans, -0.5, func, A, B0.copy(order="F"))
            blas = rp.blas_reference(side, uplo, trans, -0.5, func, A, B0)
            assert rel(got, blas) < 1e-13
            assert rp.error_metric(side, uplo, trans, -0.5, func, A, B0, got) < 1e-14


def test_oracle_against_golden_fixtures():
    cnt = 0
    for c in load_cases():
        got = c_port.unified_rectrxm(c["side"], c["uplo"], c["trans"], c["alpha"], c["func"], c["A"], c["B0"].copy(order="F"))
        assert np.array_equal(got, c["oracle"]), c["key"]          # frozen oracle output, bit exact
        tol = {8: 1e-14, 4: 1e-5, 2: 2e-2}[c["dtype"].itemsize]
        assert rel(got, c["blas"]) < tol, c["key"]                 # and the BLAS answer the reference's tests use
        cnt += 1
    assert cnt >= 140


def test_oracle_never_reads_opposite_triangle():
    A, B0 = rp.make_inputs(64, 4, "L", "L", np.float64, seed=1)
    An = A.copy(order="F"); An[np.triu_indices(64, 1)] = np.nan
    a = c_port.unified_rectrxm("L", "L", "N", 1.0, "S", An, B0.copy(order="F"))
    b = c_port.unified_rectrxm("L", "L", "N", 1.0, "S", A, B0.copy(order="F"))
    assert np.array_equal(a, b)
