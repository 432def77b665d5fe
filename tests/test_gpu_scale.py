"""GPU parity at the sizes BASELINE.json quotes, and against the frozen golden fixtures -- through the C ABI.

(1) `tests/golden/rectrxm_golden.npz` (146 seeded cases; columns: oracle output = literal restatement of the reference, and the
    OpenBLAS trsm!/trmm! output the reference's own tests compare with, test/unified_rectrxm.jl:36-40) checked against the CUDA library,
    not only against the oracle (tests/test_oracle.py does that on CPU).
(2) BASELINE configs C2..C5 at their full per-GPU sizes (n = 16384 / 32768).  The oracle would take hours there, so the gate is the
    size-independent one SURVEY.md 8(d) names: normwise backward error in FP64, computed on the GPU by an INDEPENDENT product
    (cuBLAS DGEMM through torch), with the tolerances of north_star written here:
        Float64 1e-13      Float32 1e-5 (test/trsm.jl:8)      Float16 1e-2
    plus exact linearity in alpha (power-of-two scaling commutes with every rounding) and a NaN-filled opposite triangle.
"""
import numpy as np
import pytest

from tests.golden_util import load_cases

pytestmark = pytest.mark.gpu

TOL = {"float64": 1e-13, "float32": 1e-5, "float16": 1e-2}


def rel(a, b):
    a = a.astype(np.float64); b = b.astype(np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def test_golden_fixtures_on_gpu(nla, gpu):
    """Every frozen case through nla_rectrxm: against the BLAS column within the reference's tolerance per element type, and against the
    oracle column (Float64: the two are the same arithmetic up to summation order, < 1e-13; Float32: < 1e-5; Float16: the oracle
    accumulates in Float16 like the reference, src/matmul.jl:18-19, so only the FP64 truth is a meaningful yardstick there)."""
    import torch

    cnt = 0
    worst = {}
    for c in load_cases():
        dA, dB = nla.colmajor(c["A"]), nla.colmajor(c["B0"])
        nla.unified_rectrxm(c["side"], c["uplo"], c["trans"], c["alpha"], c["func"], dA, dB)
        torch.cuda.synchronize()
        got = nla.to_numpy(dB)
        name = c["dtype"].name
        tol_blas = {"float64": 1e-14, "float32": 1e-5, "float16": 1e-2}[name]
        e_blas = rel(got, c["blas"])
        assert e_blas < tol_blas, (c["key"], name, e_blas)
        if name != "float16":
            e_or = rel(got, c["oracle"])
            assert e_or < (1e-13 if name == "float64" else 1e-5), (c["key"], name, e_or)
        worst[name] = max(worst.get(name, 0.0), e_blas)
        cnt += 1
    assert cnt >= 140
    print("golden on GPU: worst relative difference to BLAS per dtype:", worst)


def _make(torch, n, m, side, uplo, dt, seed, nan_opposite=True):
    """Scaled recipe of SURVEY.md 8(d) generated on the device: strict triangle U(-1,1)/sqrt(n), diagonal U(1,2), B = U(0,1)+1.
    The unreferenced triangle of A holds NaN, so a single stray read poisons the result."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.empty((n, n), dtype=dt, device="cuda").t()              # column-major
    blk = 4096
    for c0 in range(0, n, blk):                                      # generated in column panels: no n x n FP32/FP64 temporaries at n = 32768
        c1 = min(n, c0 + blk)
        P = ((2 * torch.rand(n, c1 - c0, dtype=torch.float32, device="cuda", generator=g) - 1) / n ** 0.5).to(dt)
        A[:, c0:c1].copy_(P)
        del P
    d = (1 + torch.rand(n, dtype=torch.float32, device="cuda", generator=g)).to(dt)
    A.diagonal().copy_(d)
    shape = (n, m) if side == "L" else (m, n)
    B0 = torch.empty((shape[1], shape[0]), dtype=dt, device="cuda").t()
    B0.copy_((torch.rand(shape, dtype=torch.float32, device="cuda", generator=g) + 1).to(dt))
    if nan_opposite:
        mask = torch.triu(torch.ones(blk, blk, dtype=torch.bool, device="cuda"), 1)
        for c0 in range(0, n, blk):
            c1 = min(n, c0 + blk)
            if uplo == "L":      # rows above the diagonal
                A[:c0, c0:c1] = float("nan")
                A[c0:c1, c0:c1].masked_fill_(mask[:c1 - c0, :c1 - c0], float("nan"))
            else:
                A[c1:, c0:c1] = float("nan")
                A[c0:c1, c0:c1].masked_fill_(mask[:c1 - c0, :c1 - c0].t(), float("nan"))
    return A, B0


def _backward_error(torch, side, uplo, trans, alpha, func, A, B0, X, blk=4096):
    """FP64 backward error with the triangle of A extracted panel by panel (the opposite triangle holds NaN) and the product done
    by cuBLAS DGEMM in row/column blocks, so that nothing larger than one panel of FP64 temporaries is alive at n = 32768."""
    n = A.shape[0]
    Xd, Bd = X.double(), B0.double()
    R = (alpha * Bd) if func == "S" else torch.zeros_like(Bd)
    V = Xd if func == "S" else Bd                       # the operand op(A) multiplies
    nA2 = 0.0
    for c0 in range(0, n, blk):
        c1 = min(n, c0 + blk)
        P = A[:, c0:c1].double()
        P = torch.tril(P, -c0) if uplo == "L" else torch.triu(P, -c0)     # NaNs of the opposite triangle are dropped here
        P = torch.nan_to_num(P, nan=0.0) if torch.isnan(P).any() else P
        nA2 += float((P * P).sum())
        # T = tri(A); op(T) = T or T^T.  Column panel P = T[:, c0:c1].
        if side == "L":
            if trans == "N":
                R -= (P @ V[c0:c1, :]) if func == "S" else -(P @ V[c0:c1, :]) * alpha
            else:   # op(T) = T^T: rows c0:c1 of op(T) are P^T
                upd = P.t() @ V
                R[c0:c1, :] -= upd if func == "S" else -upd * alpha
        else:
            if trans == "N":   # X * T: columns c0:c1 of the product are X @ P
                upd = V @ P
                R[:, c0:c1] -= upd if func == "S" else -upd * alpha
            else:              # X * T^T = sum over panels X[:, c0:c1] @ P^T
                R -= (V[:, c0:c1] @ P.t()) if func == "S" else -(V[:, c0:c1] @ P.t()) * alpha
        del P
    nA = nA2 ** 0.5
    if func == "S":   # R = alpha*B0 - op(A) X
        return (torch.linalg.norm(R) / (nA * torch.linalg.norm(Xd) + abs(alpha) * torch.linalg.norm(Bd))).item()
    # R = alpha * op(A) B0
    return (torch.linalg.norm(Xd - R) / (abs(alpha) * nA * torch.linalg.norm(Bd))).item()


CONFIGS = [
    # id, dtype, side, uplo, trans, func, n, m
    ("C2_fp64_LLN_trsm_n16384", "float64", "L", "L", "N", "S", 16384, 16384),
    ("C3_fp32_LUT_trmm_n16384", "float32", "L", "U", "T", "M", 16384, 16384),
    ("C4slice_fp16_RLN_trsm_n32768_m16384", "float16", "R", "L", "N", "S", 32768, 16384),
    ("C5slice_fp64_LLN_trmm_n32768_m8192", "float64", "L", "L", "N", "M", 32768, 8192),
    ("fp32_LLN_trsm_n16384", "float32", "L", "L", "N", "S", 16384, 16384),
    ("fp16_LLN_trsm_n16384", "float16", "L", "L", "N", "S", 16384, 16384),
]


@pytest.mark.parametrize("cfg", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_baseline_configs_full_size(nla, gpu, cfg):
    """BASELINE.json configs[1..4] at their per-GPU sizes (C4 / C5: one GPU's share at 8 GPUs), default options, through nla_rectrxm."""
    import torch

    name, dts, side, uplo, trans, func, n, m = cfg
    dt = getattr(torch, dts)
    A, B0 = _make(torch, n, m, side, uplo, dt, seed=1234 + len(name))
    X = B0.clone(memory_format=torch.preserve_format)
    assert X.stride(0) == 1
    alpha = 1.0
    nla.unified_rectrxm(side, uplo, trans, alpha, func, A, X)
    torch.cuda.synchronize()
    assert torch.isfinite(X).all(), "a NaN of the unreferenced triangle reached the result"
    err = _backward_error(torch, side, uplo, trans, alpha, func, A, B0, X)
    print(f"{name}: backward error {err:.3e} (tolerance {TOL[dts]:g})")
    assert err < TOL[dts], (name, err)
    # exact linearity in alpha: alpha = 2 must give exactly twice the alpha = 1 result (Float16: barring subnormal intermediates)
    X2 = B0.clone(memory_format=torch.preserve_format)
    nla.unified_rectrxm(side, uplo, trans, 2.0, func, A, X2)
    torch.cuda.synchronize()
    if dts != "float16":
        assert torch.equal(X2, 2 * X), name
    else:
        assert (X2 != 2 * X).float().mean().item() < 1e-3, name
    del X2, X, A, B0
    torch.cuda.empty_cache()
