"""CPU restatement (NumPy) of NextLA.jl's `unified_rectrxm!` path.  TEST INFRASTRUCTURE ONLY.

This module is the *oracle*: a literal, slow restatement of the reference algorithm used to check the
CUDA library.  Nothing under `nextla.jl_b200/` may import it; only `tests/`, `__graft_entry__.smoke()`
and the `cpu_baseline` / `--impl reference` legs of `bench.py` do.

Each function cites the reference file:line it follows (paths relative to /root/reference):

  unified_rectrxm   <- src/rectrxm.jl:43-76
  unified_rec       <- src/rectrxm.jl:101-198
  gemm_add/gemm_sub <- src/matmul.jl:69-81 (wrappers), matmul_kernel <- src/matmul.jl:5-66
  *_trsm leaves     <- src/trsm.jl:5-150
  *_trmm leaves     <- src/trmm.jl:43-389
  laswp             <- src/lu.jl:470-530
  getrf2            <- src/lu.jl:185-299 (the recursive LU whose TRSM / GEMM steps sit on this path)

Arithmetic is carried out in the element type T of the arrays (float64 / float32 / float16), exactly as
the Julia kernels do (`eltype(output)` accumulators, src/matmul.jl:18-19,50-54; the final GEMM update is
done in Float64 because `alpha::Float64`, src/matmul.jl:6,64).

Parity status: the reference ships NO golden vectors for this path and cannot be executed here (no
Julia in the image).  The oracle is pinned by the reference's *own test criterion* instead: on the
reference's test grid (test/unified_rectrxm.jl:10-44, test/trsm.jl:10-64) it agrees with OpenBLAS
trsm/trmm (the routine the reference's tests compare against) within the reference's tolerances
(1e-14 FP64, 1e-5 FP32).  Everything the reference's tests do not cover (n > 256 for TRSM, alpha != 1,
Float16, non power-of-two n) is "parity unpinned" by the reference and pinned here by BLAS equivalence
and backward error only.  See tests/test_oracle.py and DESIGN.md.
"""
from __future__ import annotations

import math

import numpy as np

TILE_DIM = 32  # src/matmul.jl:3
TRMM_THRESHOLD = 16  # src/rectrxm.jl:52
TRSM_THRESHOLD = 256  # src/rectrxm.jl:63


# --------------------------------------------------------------------------------------------
# GEMM update  (src/matmul.jl)
# --------------------------------------------------------------------------------------------
def matmul_kernel(output: np.ndarray, in1: np.ndarray, in2: np.ndarray, alpha: float) -> None:
    """output[I,J] += alpha * sum_k in1[I,k]*in2[k,J]   (src/matmul.jl:5-66).

    Per output element the reference forms, for every K-tile of 32, a partial sum `tmp` in T over the
    zero-padded tile (src/matmul.jl:50-53), adds it to the running `outval` in T (:54) and finally does
    `output += alpha*outval` in Float64 (alpha is a Float64, :6,:64), rounding once to T on the store.
    Vectorised over (I,J); the k order inside a tile is the sequential 1..32 order.
    """
    T = output.dtype.type
    N, R = in1.shape
    M = in2.shape[1]
    assert in2.shape[0] == R and output.shape == (N, M)
    outval = np.zeros((N, M), dtype=T)
    for t in range(math.ceil(R / TILE_DIM)):
        k0, k1 = t * TILE_DIM, min((t + 1) * TILE_DIM, R)
        tmp = np.zeros((N, M), dtype=T)
        for k in range(k0, k1):  # padded entries are zeros and add nothing
            tmp += in1[:, k : k + 1] * in2[k : k + 1, :]
        outval += tmp
    res = output.astype(np.float64) + np.float64(alpha) * outval.astype(np.float64)
    output[...] = res.astype(T)


def gemm_add(A: np.ndarray, B: np.ndarray, C: np.ndarray) -> None:
    """GEMM_ADD!(A,B,C): C += A*B   (src/matmul.jl:69-74)."""
    matmul_kernel(C, A, B, 1.0)


def gemm_sub(A: np.ndarray, B: np.ndarray, C: np.ndarray) -> None:
    """GEMM_SUB!(A,B,C): A -= B*C   (src/matmul.jl:76-81)."""
    matmul_kernel(A, B, C, -1.0)


# --------------------------------------------------------------------------------------------
# TRSM leaves  (src/trsm.jl)
# --------------------------------------------------------------------------------------------
def left_lower_trsm(A: np.ndarray, B: np.ndarray) -> None:
    """LeftLowerTRSM! (src/trsm.jl:128-132) -> lower_left_kernel (src/trsm.jl:5-33).

    The launcher passes Transpose(A) and the kernel indexes it as A[i,row], i.e. the parent's A[row,i].
    x_r <- b_r/d_r (:15-18); then for pivots i = 1..n, rows r > i: x_r -= (a_ri/d_r) * x_i (:21-27).
    """
    n = B.shape[0]
    d = np.diagonal(A).copy()
    X = B / d[:, None]
    for i in range(n):
        if i + 1 < n:
            acol = A[i + 1 :, i] / d[i + 1 :]
            X[i + 1 :, :] -= acol[:, None] * X[i : i + 1, :]
    B[...] = X


def left_upper_trsm(A: np.ndarray, B: np.ndarray) -> None:
    """LeftUpperTRSM! (src/trsm.jl:134-138) -> upper_left_kernel (src/trsm.jl:36-64): pivots n..1, rows r < i."""
    n = B.shape[0]
    d = np.diagonal(A).copy()
    X = B / d[:, None]
    for i in range(n - 1, 0, -1):
        acol = A[:i, i] / d[:i]
        X[:i, :] -= acol[:, None] * X[i : i + 1, :]
    B[...] = X


def right_lower_trsm(A: np.ndarray, B: np.ndarray) -> None:
    """RightLowerTRSM! (src/trsm.jl:140-144) -> right_lower_kernel (src/trsm.jl:67-95).

    One RHS per row of B: x_c <- b_c/d_c; pivots i = n..1, columns c < i: x_c -= x_i * (a_ic/d_c).
    """
    n = B.shape[1]
    d = np.diagonal(A).copy()
    X = B / d[None, :]
    for i in range(n - 1, 0, -1):
        arow = A[i, :i] / d[:i]
        X[:, :i] -= X[:, i : i + 1] * arow[None, :]
    B[...] = X


def right_upper_trsm(A: np.ndarray, B: np.ndarray) -> None:
    """RightUpperTRSM! (src/trsm.jl:146-150) -> right_upper_kernel (src/trsm.jl:98-126).

    Launcher passes Transpose(A), kernel reads A[col,i] of it = parent's A[i,col]; pivots 1..n, cols c > i.
    """
    n = B.shape[1]
    d = np.diagonal(A).copy()
    X = B / d[None, :]
    for i in range(n - 1):
        arow = A[i, i + 1 :] / d[i + 1 :]
        X[:, i + 1 :] -= X[:, i : i + 1] * arow[None, :]
    B[...] = X


# --------------------------------------------------------------------------------------------
# TRMM leaves  (src/trmm.jl) -- valid for n <= 16 (one 16x16 tile of A in shared memory)
# --------------------------------------------------------------------------------------------
def left_lower_trmm(A: np.ndarray, B: np.ndarray) -> None:
    """LeftLowerTRMM! (src/trmm.jl:332-338) -> kernel :43-109: out[i,j] = sum_{k=1..i} A[i,k]*B[k,j] (:91-93)."""
    n = A.shape[0]
    assert n <= 16
    T = B.dtype.type
    out = np.zeros_like(B)
    for i in range(n):
        acc = np.zeros(B.shape[1], dtype=T)
        for k in range(i + 1):
            acc += A[i, k] * B[k, :]
        out[i, :] = acc
    B[...] = out


def left_upper_trmm(A: np.ndarray, B: np.ndarray) -> None:
    """LeftUpperTRMM! (src/trmm.jl:352-356) -> kernel :116-180: out[i,j] = sum_{k=i..N} A[i,k]*B[k,j] (:164-166)."""
    n = A.shape[0]
    assert n <= 16
    T = B.dtype.type
    out = np.zeros_like(B)
    for i in range(n):
        acc = np.zeros(B.shape[1], dtype=T)
        for k in range(i, n):
            acc += A[i, k] * B[k, :]
        out[i, :] = acc
    B[...] = out


def right_lower_trmm(A: np.ndarray, B: np.ndarray) -> None:
    """RightLowerTRMM! (src/trmm.jl:367-372) -> kernel :189-250: out[i,j] = sum_{k=j..N} B[i,k]*A[k,j] (:233-235)."""
    n = A.shape[0]
    assert n <= 16
    T = B.dtype.type
    out = np.zeros_like(B)
    for j in range(n):
        acc = np.zeros(B.shape[0], dtype=T)
        for k in range(j, n):
            acc += B[:, k] * A[k, j]
        out[:, j] = acc
    B[...] = out


def right_upper_trmm(A: np.ndarray, B: np.ndarray) -> None:
    """RightUpperTRMM! (src/trmm.jl:384-389) -> kernel :252-312: out[i,j] = sum_{k=1..j} B[i,k]*A[k,j] (:296-298)."""
    n = A.shape[0]
    assert n <= 16
    T = B.dtype.type
    out = np.zeros_like(B)
    for j in range(n):
        acc = np.zeros(B.shape[0], dtype=T)
        for k in range(j + 1):
            acc += B[:, k] * A[k, j]
        out[:, j] = acc
    B[...] = out


# --------------------------------------------------------------------------------------------
# Recursive splitter  (src/rectrxm.jl)
# --------------------------------------------------------------------------------------------
def unified_rec(func: str, side: str, uplo: str, A: np.ndarray, n: int, B: np.ndarray, threshold: int) -> None:
    """src/rectrxm.jl:101-198.  A and B are NumPy views (the Julia code uses `view`/Transpose wrappers)."""
    if n <= threshold:  # :103-126  (unknown chars fall into the else branches exactly as in Julia)
        if func == "S":
            if side == "L" and uplo == "L":
                left_lower_trsm(A, B)
            elif side == "L" and uplo == "U":
                left_upper_trsm(A, B)
            elif side == "R" and uplo == "L":
                right_lower_trsm(A, B)
            else:
                right_upper_trsm(A, B)
        else:
            if side == "L" and uplo == "L":
                left_lower_trmm(A, B)
            elif side == "L" and uplo == "U":
                left_upper_trmm(A, B)
            elif side == "R" and uplo == "L":
                right_lower_trmm(A, B)
            else:
                right_upper_trmm(A, B)
        return

    # :129-134  power of two -> halve; otherwise largest power of two below n
    if n & (n - 1) == 0:
        mid = n // 2
    else:
        mid = 2 ** int(math.floor(math.log2(n)))
    rem = n - mid

    A11, A22 = A[:mid, :mid], A[mid:n, mid:n]  # :137-140
    A21, A12 = A[mid:n, :mid], A[:mid, mid:n]
    if side == "L":  # :143-149
        B1, B2 = B[:mid, :], B[mid:n, :]
    else:
        B1, B2 = B[:, :mid], B[:, mid:n]

    forward = (
        (side == "L" and uplo == "L" and func == "S")
        or (side == "R" and uplo == "U" and func == "S")
        or (side == "L" and uplo == "U" and func == "M")
        or (side == "R" and uplo == "L" and func == "M")
    )  # :153-156
    if forward:
        unified_rec(func, side, uplo, A11, mid, B1, threshold)  # :159
        if side == "L":
            if func == "S":
                gemm_sub(B2, A21, B1)  # :164  B2 -= A21*B1
            else:
                gemm_add(A12, B2, B1)  # :166  B1 += A12*B2
        else:
            if func == "S":
                gemm_sub(B2, B1, A12)  # :170  B2 -= B1*A12
            else:
                gemm_add(B2, A21, B1)  # :172  B1 += B2*A21
        unified_rec(func, side, uplo, A22, rem, B2, threshold)  # :176
    else:
        unified_rec(func, side, uplo, A22, rem, B2, threshold)  # :179
        if side == "L":
            if func == "S":
                gemm_sub(B1, A12, B2)  # :184  B1 -= A12*B2
            else:
                gemm_add(A21, B1, B2)  # :186  B2 += A21*B1
        else:
            if func == "S":
                gemm_sub(B1, B2, A21)  # :190  B1 -= B2*A21
            else:
                gemm_add(B1, A12, B2)  # :192  B2 += B1*A12
        unified_rec(func, side, uplo, A11, mid, B1, threshold)  # :196


def unified_rectrxm(side: str, uplo: str, transpose: str, alpha: float, func: str, A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """unified_rectrxm!(side, uplo, transpose, alpha, func, A, B)  (src/rectrxm.jl:43-76).  In place on B."""
    T = B.dtype.type
    threshold = TRMM_THRESHOLD  # :52
    n = A.shape[0]  # :53
    if transpose in ("T", "C"):  # :56-59 (real eltypes: Adjoint == Transpose)
        A = A.T
        uplo = "U" if uplo == "L" else "L"
    if func == "S":  # :62-65
        threshold = TRSM_THRESHOLD
        B[...] = (np.float64(alpha) * B.astype(np.float64)).astype(T)  # alpha is a Float64 -> promoted product
    unified_rec(func, side, uplo, A, n, B, threshold)  # :68
    if func == "M":  # :71-73
        B[...] = (np.float64(alpha) * B.astype(np.float64)).astype(T)
    return B


# --------------------------------------------------------------------------------------------
# Input recipes and error metrics shared by tests / bench
# --------------------------------------------------------------------------------------------
def make_inputs(n: int, m: int, side: str, uplo: str, dtype, seed: int, recipe: str = "reference"):
    """`reference` recipe = test/unified_rectrxm.jl:20-27: A = tri(U(0,1)+1) + 10 I, B = U(0,1)+1.
    `scaled` recipe: strictly-triangular part U(-1,1)/sqrt(n), diagonal U(1,2), B = U(0,1)+1 (well conditioned at any n)."""
    rng = np.random.RandomState(seed)
    if recipe == "reference":
        A = rng.rand(n, n) + 1.0
        A = np.tril(A) if uplo == "L" else np.triu(A)
        A = A + 10.0 * np.eye(n)
    else:
        A = (2.0 * rng.rand(n, n) - 1.0) / math.sqrt(n)
        A = np.tril(A, -1) if uplo == "L" else np.triu(A, 1)
        A = A + np.diag(1.0 + rng.rand(n))
    B = rng.rand(n, m) + 1.0 if side == "L" else rng.rand(m, n) + 1.0
    return np.asfortranarray(A.astype(dtype)), np.asfortranarray(B.astype(dtype))


def op_matrix(A: np.ndarray, uplo: str, trans: str) -> np.ndarray:
    """op(tri(A)) in float64: the matrix the BLAS definition of the call uses."""
    A64 = A.astype(np.float64)
    Tm = np.tril(A64) if uplo == "L" else np.triu(A64)
    return Tm.T if trans in ("T", "C") else Tm


def blas_reference(side, uplo, trans, alpha, func, A, B0) -> np.ndarray:
    """The reference's own test oracle (test/unified_rectrxm.jl:36-40): OpenBLAS trsm!/trmm! via SciPy
    for float64/float32; for float16 (no BLAS routine) the float64 result on the float16-rounded inputs."""
    from scipy.linalg import blas

    lower = 1 if uplo == "L" else 0
    ta = 0 if trans == "N" else 1
    sd = 0 if side == "L" else 1
    if A.dtype == np.float64:
        f = blas.dtrsm if func == "S" else blas.dtrmm
        return f(alpha, A, B0, side=sd, lower=lower, trans_a=ta, diag=0)
    if A.dtype == np.float32:
        f = blas.strsm if func == "S" else blas.strmm
        return f(np.float32(alpha), A, B0, side=sd, lower=lower, trans_a=ta, diag=0)
    f = blas.dtrsm if func == "S" else blas.dtrmm
    return f(alpha, A.astype(np.float64), B0.astype(np.float64), side=sd, lower=lower, trans_a=ta, diag=0)


def error_metric(side, uplo, trans, alpha, func, A, B0, X) -> float:
    """Normwise relative error in float64:
    TRSM: backward error ||op(A) X - alpha B0||_F / (||A||_F ||X||_F + |alpha| ||B0||_F)   (side R: X op(A))
    TRMM: ||X - alpha op(A) B0||_F / (|alpha| ||A||_F ||B0||_F)."""
    Tm = op_matrix(A, uplo, trans)
    X64, B64 = X.astype(np.float64), B0.astype(np.float64)
    nA = np.linalg.norm(Tm)
    if func == "S":
        R = (Tm @ X64 if side == "L" else X64 @ Tm) - alpha * B64
        return float(np.linalg.norm(R) / (nA * np.linalg.norm(X64) + abs(alpha) * np.linalg.norm(B64)))
    P = alpha * (Tm @ B64 if side == "L" else B64 @ Tm)
    return float(np.linalg.norm(X64 - P) / (abs(alpha) * nA * np.linalg.norm(B64) + 1e-300))


# --------------------------------------------------------------------------------------------
# Recursive LU  (src/lu.jl) -- SURVEY.md 8(f2)
# --------------------------------------------------------------------------------------------
def laswp(A: np.ndarray, first: int, last: int, ipiv: np.ndarray, incx: int) -> None:
    """src/lu.jl:470-530: rows i and ipiv[i] (1-based) of A are exchanged for i = first..last (incx > 0) or last..first (incx < 0).
    The reference walks the columns in blocks of 32 (:487-504) and then the remainder (:505-519); the result does not depend on it."""
    if incx == 0:
        return
    rows = range(first, last + 1) if incx > 0 else range(last, first - 1, -1)
    for i in rows:
        ip = int(ipiv[i - 1])
        if ip != i:
            A[[i - 1, ip - 1], :] = A[[ip - 1, i - 1], :]


def getrf2(A: np.ndarray, ipiv: np.ndarray) -> int:
    """src/lu.jl:185-299, in place on A (m x n, element type kept) and ipiv (1-based, view-relative like the reference's views);
    returns info (0, or the 1-based index of the first exactly-zero pivot).

    The reference carries `info` as a by-value Int argument and writes `info[] = 1` in the base cases (:219, :247), which cannot update the
    caller's value; the intent is LAPACK's dgetrf2 (the code is a transcription of it: iinfo propagated as `iinfo + n1`, :260-262, :289-291)
    and that is what is restated here.  The TRSM / GEMM steps (:277, :280) are BLAS calls in the reference; here they are NumPy in the
    element type of A."""
    m, n = A.shape
    T = A.dtype.type
    if m == 0 or n == 0:                       # :206
        return 0
    if m == 1:                                 # :216-222
        ipiv[0] = 1
        return 1 if A[0, 0] == 0 else 0
    if n == 1:                                 # :224-251
        sfmin = np.finfo(A.dtype).tiny         # lamch('S')
        col = A[:, 0]
        idamax = int(np.argmax(np.abs(col)))   # first index of the largest |.| (np.argmax returns the first maximum)
        ipiv[0] = idamax + 1
        if col[idamax] == 0:
            return 1
        if idamax != 0:
            col[0], col[idamax] = col[idamax], col[0]
        if abs(col[0]) >= sfmin:
            col[1:] *= T(1) / col[0]           # BLAS.scal! by the reciprocal (:240)
        else:
            col[1:] /= col[0]                  # :242
        return 0
    n1 = min(m, n) // 2                        # :254
    info = getrf2(A[:, :n1], ipiv[:n1])        # :258
    laswp(A[:, n1:], 1, n1, ipiv, 1)           # :274
    L11 = np.tril(A[:n1, :n1], -1).astype(A.dtype) + np.eye(n1, dtype=A.dtype)
    A12 = A[:n1, n1:]
    for i in range(1, n1):                     # trsm!('L','L','N','U') (:277): forward substitution with a unit diagonal
        A12[i, :] -= (L11[i, :i] @ A12[:i, :]).astype(A.dtype)
    A[n1:, n1:] -= (A[n1:, :n1] @ A12).astype(A.dtype)      # gemm!('N','N',-1,...,1) (:280)
    k = min(m, n)
    iinfo = getrf2(A[n1:, n1:], ipiv[n1:k])    # :284
    if info == 0 and iinfo > 0:                # :289-291
        info = iinfo + n1
    ipiv[n1:k] += n1                           # :293-295
    laswp(A[:, :n1], n1 + 1, k, ipiv, 1)       # :298
    return info
