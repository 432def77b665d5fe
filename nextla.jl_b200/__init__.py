"""nextla.jl_b200 -- host-side mirror (Python/ctypes) of NextLA.jl's `unified_rectrxm!` interface on top of
the C-ABI CUDA library `libnextla_b200.so` (include/nextla_b200.h).

The reference is a Julia package; Julia is not installed in this image, so the drop-in Julia method lives in
`julia/NextLAB200.jl` (unexecuted here) and this module drives the *same* C ABI for tests and benchmarks.
Function names and argument order follow the reference:

  unified_rectrxm(side, uplo, transpose, alpha, func, A, B)   <- unified_rectrxm!  src/rectrxm.jl:43
  LeftLowerTRSM(A, B) ... RightUpperTRSM(A, B)                 <- src/trsm.jl:128-150
  LeftLowerTRMM(A, B) ... RightUpperTRMM(A, B)                 <- src/trmm.jl:332-389
  GEMM_ADD(A, B, C)  (C += A*B),  GEMM_SUB(A, B, C)  (A -= B*C) <- src/matmul.jl:69-81
  trsm(side, uplo, transa, diag, A, B, alpha), trmm(...)       <- src/trsm.jl:186-205, src/trmm.jl:430-448
  laswp(A, first, last, ipiv, incx), getrf2_update(A, n1, ipiv) <- src/lu.jl:470-530, :274-280 (the device steps of the recursive LU)
  getrf2(A)                                                     <- src/lu.jl:185-299 (the whole recursive LU on the device, nla_getrf2)
  lauum(uplo, A, ib)                                            <- src/lauum.jl:52-186 (block loop; O(n^3) steps on this library's kernels)

Matrices are column-major device arrays: 2-D torch CUDA tensors with stride (1, ld) (use `colmajor()` /
`to_numpy()`).  Everything runs on the GPU through the library; there is NO CPU fallback -- importing is fine
without a GPU, but any compute call raises if the library or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnextla_b200.so")
HEADER_PATH = os.path.join(_HERE, "..", "include", "nextla_b200.h")

NLA_F64, NLA_F32, NLA_F16, NLA_C64, NLA_C128 = 0, 1, 2, 3, 4
_lib = None
_handles = {}


class NextLAError(RuntimeError):
    pass


def load_library():
    """dlopen the CUDA library and declare the prototypes of every symbol in include/nextla_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NextLAError(f"{LIB_PATH} is missing: build it with `python nextla.jl_b200/build.py` (no CPU fallback exists)")
    lib = ctypes.CDLL(LIB_PATH)
    c = ctypes
    H, I, L, D, P, CH = c.c_void_p, c.c_int, c.c_int64, c.c_double, c.c_void_p, c.c_char
    protos = {
        "nla_create": (I, [c.POINTER(H), I]),
        "nla_destroy": (I, [H]),
        "nla_status_string": (c.c_char_p, [I]),
        "nla_last_cuda_error": (I, [H]),
        "nla_version": (I, []),
        "nla_probe_fp64_peak": (I, [H, c.POINTER(c.c_double)]),
        "nla_lauum": (I, [H, CH, I, L, P, L, L, P]),
        "nla_rectrxm_complex": (I, [H, CH, CH, CH, CH, CH, I, L, L, D, D, P, L, P, L, P]),
        "nla_mg_create": (I, [c.POINTER(H), I, c.POINTER(c.c_int)]),
        "nla_mg_destroy": (I, [H]),
        "nla_mg_device_count": (I, [H]),
        "nla_mg_handle": (H, [H, I]),
        "nla_mg_stream": (P, [H, I]),
        "nla_mg_last_nccl_error": (I, [H]),
        "nla_mg_sync": (I, [H]),
        "nla_mg_rectrxm": (I, [H, CH, CH, CH, CH, I, L, D, I, P, L, c.POINTER(c.c_void_p), c.POINTER(c.c_int64), c.POINTER(c.c_int64)]),
        "nla_mg_rectrxm_host": (I, [H, CH, CH, CH, CH, I, L, D, P, L, c.POINTER(c.c_void_p), c.POINTER(c.c_int64), c.POINTER(c.c_int64)]),
        "nla_workspace_bytes": (L, [H, CH, CH, I, L, L]),
        "nla_reserve": (I, [H, CH, CH, I, L, L]),
        "nla_set_workspace": (I, [H, P, L]),
        "nla_rectrxm": (I, [H, CH, CH, CH, CH, I, L, L, D, P, L, P, L, P]),
        "nla_trxm": (I, [H, CH, CH, CH, CH, CH, I, L, L, D, P, L, P, L, P]),
        "nla_rectrxm_host": (I, [H, CH, CH, CH, CH, I, L, L, D, P, L, P, L]),
        "nla_rectrxm_gated": (I, [H, CH, CH, CH, CH, I, L, L, D, P, L, P, L, P, L, L, c.POINTER(c.c_void_p)]),
        "nla_rectrxm_hostb_gated": (I, [H, CH, CH, CH, CH, I, L, L, D, P, L, P, L, L, L, c.POINTER(c.c_void_p)]),
        "nla_memcpy2d_async": (I, [H, P, L, P, L, L, L, I, P]),
        "nla_laswp": (I, [H, I, L, L, P, L, L, L, P, I, P]),
        "nla_getrf2": (I, [H, I, L, L, P, L, P, P, P]),
        "nla_panel_order": (L, [CH, CH, CH, CH, L, L, c.POINTER(c.c_int64), L]),
        "nla_trsm_leaf": (I, [H, CH, CH, I, L, L, P, L, P, L, P]),
        "nla_trmm_leaf": (I, [H, CH, CH, I, L, L, P, L, P, L, P]),
        "nla_leaf_max": (L, [I]),
        "nla_gemm_update": (I, [H, I, CH, CH, L, L, L, I, P, L, P, L, P, L, P]),
        "nla_set_option": (I, [H, c.c_char_p, L]),
        "nla_get_option": (L, [H, c.c_char_p]),
        "nla_launch_count": (L, [H, I]),
        "nla_profile_read": (L, [H, c.POINTER(c.c_double), L]),
        "nla_plan": (L, [CH, CH, CH, CH, L, L, c.POINTER(c.c_int64), L]),
        "nla_host_plan": (L, [CH, CH, CH, CH, L, L, L, I, c.POINTER(c.c_int64), L, c.POINTER(c.c_int64), L, c.POINTER(c.c_int64),
                              c.POINTER(c.c_int64), c.POINTER(c.c_int64)]),
    }
    for name, (res, args) in protos.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def exported_symbols():
    return ["nla_create", "nla_destroy", "nla_status_string", "nla_last_cuda_error", "nla_version", "nla_rectrxm", "nla_rectrxm_host",
            "nla_workspace_bytes", "nla_reserve", "nla_set_workspace", "nla_probe_fp64_peak", "nla_lauum", "nla_getrf2", "nla_rectrxm_complex",
            "nla_mg_create", "nla_mg_destroy", "nla_mg_device_count", "nla_mg_handle", "nla_mg_stream", "nla_mg_last_nccl_error", "nla_mg_sync",
            "nla_mg_rectrxm", "nla_mg_rectrxm_host",
            "nla_rectrxm_gated", "nla_rectrxm_hostb_gated", "nla_panel_order", "nla_trxm", "nla_memcpy2d_async", "nla_laswp", "nla_host_plan",
            "nla_trsm_leaf", "nla_trmm_leaf", "nla_leaf_max", "nla_gemm_update", "nla_set_option", "nla_get_option", "nla_launch_count", "nla_plan", "nla_profile_read"]


def _check(rc: int, h=None):
    if rc != 0:
        lib = load_library()
        msg = lib.nla_status_string(rc).decode()
        if rc == 5 and h is not None:
            msg += f" [cudaError {lib.nla_last_cuda_error(h)}]"
        raise NextLAError(f"nextla_b200 status {rc}: {msg}")


class Handle:
    """Owns one nla_handle_t (one per device)."""

    def __init__(self, device: int = 0):
        lib = load_library()
        self._h = ctypes.c_void_p()
        _check(lib.nla_create(ctypes.byref(self._h), device))
        self.device = device

    def set_option(self, key: str, value: int):
        _check(load_library().nla_set_option(self._h, key.encode(), int(value)), self._h)

    def get_option(self, key: str) -> int:
        return int(load_library().nla_get_option(self._h, key.encode()))

    def workspace_bytes(self, side: str, func: str, dtype_code: int, n: int, m: int) -> int:
        r = int(load_library().nla_workspace_bytes(self._h, side.encode(), func.encode(), dtype_code, n, m))
        if r < 0:
            _check(-r, self._h)
        return r

    def reserve(self, side: str, func: str, dtype_code: int, n: int, m: int):
        _check(load_library().nla_reserve(self._h, side.encode(), func.encode(), dtype_code, n, m), self._h)

    def set_workspace(self, buf):
        """Caller-owned arena: a 1-D torch CUDA uint8 tensor (256-byte aligned), or None to return to library-owned workspaces.
        The tensor is kept alive by the handle."""
        self._ws_keepalive = buf
        if buf is None:
            _check(load_library().nla_set_workspace(self._h, None, 0), self._h)
        else:
            _check(load_library().nla_set_workspace(self._h, buf.data_ptr(), buf.numel() * buf.element_size()), self._h)

    def probe_fp64_peak(self) -> float:
        """Measured DMMA.8x8x4 issue-rate peak of this GPU in TFLOP/s (a few milliseconds, synchronous)."""
        v = ctypes.c_double(0.0)
        _check(load_library().nla_probe_fp64_peak(self._h, ctypes.byref(v)), self._h)
        return float(v.value)

    def launch_count(self, reset: bool = False) -> int:
        return int(load_library().nla_launch_count(self._h, 1 if reset else 0))

    def profile_read(self, max_records: int = 1 << 16):
        """[(kind, flops, ms)] for every launch since option "profile" was set / the last read (kind 0 = leaf, 1 = GEMM update)."""
        buf = (ctypes.c_double * (3 * max_records))()
        n = load_library().nla_profile_read(self._h, buf, max_records)
        if n < 0:
            _check(-n, self._h)
        n = min(n, max_records)
        return [(int(buf[3 * i]), buf[3 * i + 1], buf[3 * i + 2]) for i in range(n)]

    def close(self):
        if self._h:
            load_library().nla_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def default_handle(device: Optional[int] = None, stream=None) -> Handle:
    """The cached handle of (device, stream).  A handle owns device workspaces, so calls through one handle must be ordered on one
    stream (include/nextla_b200.h): the cache is keyed by the stream as well -- two threads / streams on one GPU get two handles."""
    import torch

    if device is None:
        device = torch.cuda.current_device()
    s = torch.cuda.current_stream(device) if stream is None else stream
    key = (device, int(s.cuda_stream))
    if key not in _handles:
        _handles[key] = Handle(device)
    return _handles[key]


def _handle_for(handle: Optional["Handle"], t, stream=None) -> "Handle":
    """Explicit handle (checked against the tensor's device) or the cached one of (tensor device, stream)."""
    if handle is None:
        return default_handle(t.device.index, stream)
    if handle.device != t.device.index:
        raise NextLAError(f"handle of device {handle.device} used with a matrix on device {t.device.index}")
    return handle


# ---- column-major device matrices ---------------------------------------------------------------
_NP2DT = {np.dtype(np.float64): NLA_F64, np.dtype(np.float32): NLA_F32, np.dtype(np.float16): NLA_F16}


def colmajor(a: np.ndarray, device: str = "cuda"):
    """Host ndarray -> column-major torch CUDA tensor of the same logical shape (stride (1, rows))."""
    import torch

    a = np.asarray(a)
    t = torch.from_numpy(np.ascontiguousarray(a.T)).to(device)
    return t.t()


def empty_colmajor(rows: int, cols: int, dtype, device: str = "cuda", ld: Optional[int] = None):
    import torch

    ld = rows if ld is None else ld
    return torch.empty((cols, ld), dtype=dtype, device=device).t()[:rows, :]


def to_numpy(t) -> np.ndarray:
    return np.asfortranarray(t.detach().cpu().numpy())


def _desc(t):
    """(ptr, rows, cols, ld, dtype code) of a column-major 2-D torch CUDA tensor."""
    import torch

    if not t.is_cuda:
        raise NextLAError("matrices must be CUDA tensors (no CPU fallback)")
    if t.dim() != 2:
        raise NextLAError("matrices must be 2-D")
    rows, cols = t.shape
    if rows > 1 and t.stride(0) != 1:
        raise NextLAError("matrices must be column-major (stride(0) == 1); use nextla colmajor()")
    ld = t.stride(1) if (cols > 1 and rows > 0) else max(1, rows)
    if ld < max(1, rows):
        raise NextLAError("invalid leading dimension")
    code = {torch.float64: NLA_F64, torch.float32: NLA_F32, torch.float16: NLA_F16, torch.complex64: NLA_C64, torch.complex128: NLA_C128}.get(t.dtype)
    if code is None:
        raise NextLAError(f"unsupported dtype {t.dtype}")
    return t.data_ptr(), rows, cols, ld, code


def _stream_ptr(stream, device=None):
    import torch

    s = torch.cuda.current_stream(device) if stream is None else stream
    return ctypes.c_void_p(s.cuda_stream)


def _ch(c: str) -> bytes:
    if not isinstance(c, str) or len(c) != 1:
        raise NextLAError(f"expected a single character, got {c!r}")
    return c.encode()


def unified_rectrxm(side: str, uplo: str, transpose: str, alpha: float, func: str, A, B, stream=None, handle: Optional[Handle] = None):
    """unified_rectrxm!(side, uplo, transpose, alpha, func, A, B) -- src/rectrxm.jl:43-76.  In place on B; returns B.
    Asynchronous on the current torch stream, like the reference (which does not synchronise, :75)."""
    h = _handle_for(handle, A, stream)
    pa, ar, ac, lda, dta = _desc(A)
    pb, br, bc, ldb, dtb = _desc(B)
    if dta != dtb:
        raise NextLAError("A and B must have the same element type (src/rectrxm.jl:101)")
    if ar != ac:
        raise NextLAError("A must be square")
    n = ar
    m = bc if side == "L" else br
    if (side == "L" and br != n) or (side == "R" and bc != n):
        raise NextLAError("dimension mismatch between A and B")
    if dta in (NLA_C64, NLA_C128):   # complex element types: 'C' != 'T', complex alpha (nla_rectrxm_complex)
        a = complex(alpha)
        rc = load_library().nla_rectrxm_complex(h._h, _ch(side), _ch(uplo), _ch(transpose), b"N", _ch(func), dta, n, m, a.real, a.imag, pa, lda,
                                                pb, ldb, _stream_ptr(stream, h.device))
        _check(rc, h._h)
        return B
    rc = load_library().nla_rectrxm(h._h, _ch(side), _ch(uplo), _ch(transpose), _ch(func), dta, n, m, float(alpha), pa, lda, pb, ldb,
                                    _stream_ptr(stream, h.device))
    _check(rc, h._h)
    return B


def panel_order(side: str, uplo: str, transpose: str, func: str, n: int, panel_cols: int):
    """Host-only: the order in which the schedule first reads the column panels of A (panel p = columns [p*panel_cols, ...)),
    i.e. the order in which a pipelined broadcast should deliver them."""
    lib = load_library()
    cnt = lib.nla_panel_order(_ch(side), _ch(uplo), _ch(transpose), _ch(func), n, panel_cols, None, 0)
    if cnt < 0:
        _check(-cnt)
    buf = (ctypes.c_int64 * max(cnt, 1))()
    lib.nla_panel_order(_ch(side), _ch(uplo), _ch(transpose), _ch(func), n, panel_cols, buf, cnt)
    return [int(buf[i]) for i in range(cnt)]


def unified_rectrxm_gated(side: str, uplo: str, transpose: str, alpha: float, func: str, A, B, panel_cols: int, panel_events, stream=None,
                          handle: Optional[Handle] = None):
    """unified_rectrxm with A arriving in column panels: `panel_events[p]` is a recorded torch.cuda.Event marking panel p
    (columns [p*panel_cols, (p+1)*panel_cols)) valid; the schedule waits for a panel right before the first launch reading it."""
    h = _handle_for(handle, A, stream)
    pa, ar, ac, lda, dta = _desc(A)
    pb, br, bc, ldb, dtb = _desc(B)
    if dta != dtb or ar != ac:
        raise NextLAError("A must be square and share B's element type")
    n = ar
    m = bc if side == "L" else br
    if (side == "L" and br != n) or (side == "R" and bc != n):
        raise NextLAError("dimension mismatch between A and B")
    evs = (ctypes.c_void_p * len(panel_events))(*[ctypes.c_void_p(e.cuda_event) for e in panel_events])
    rc = load_library().nla_rectrxm_gated(h._h, _ch(side), _ch(uplo), _ch(transpose), _ch(func), dta, n, m, float(alpha), pa, lda, pb, ldb,
                                          _stream_ptr(stream, h.device), int(panel_cols), len(panel_events), evs)
    _check(rc, h._h)
    return B


def unified_rectrxm_host(side, uplo, transpose, alpha, func, A: np.ndarray, B: np.ndarray, handle: Optional[Handle] = None, device: int = 0):
    """Same operation on HOST column-major ndarrays (Fortran order); B is overwritten.  Synchronous."""
    h = handle or default_handle(device)
    if not (A.flags.f_contiguous and B.flags.f_contiguous) or A.dtype != B.dtype:
        raise NextLAError("host matrices must be Fortran-ordered ndarrays of the same dtype")
    n = A.shape[0]
    m = B.shape[1] if side == "L" else B.shape[0]
    rc = load_library().nla_rectrxm_host(h._h, _ch(side), _ch(uplo), _ch(transpose), _ch(func), _NP2DT[A.dtype], n, m, float(alpha),
                                         A.ctypes.data, max(1, A.shape[0]), B.ctypes.data, max(1, B.shape[0]))
    _check(rc, h._h)
    return B


def _leaf(solve: bool, side: str, uplo: str, A, B, stream=None, handle=None):
    h = _handle_for(handle, A, stream)
    pa, ar, ac, lda, dta = _desc(A)
    pb, br, bc, ldb, dtb = _desc(B)
    if dta != dtb or ar != ac:
        raise NextLAError("bad leaf arguments")
    n = ar
    m = bc if side == "L" else br
    fn = load_library().nla_trsm_leaf if solve else load_library().nla_trmm_leaf
    _check(fn(h._h, _ch(side), _ch(uplo), dta, n, m, pa, lda, pb, ldb, _stream_ptr(stream, h.device)), h._h)
    return B


def LeftLowerTRSM(A, B, **kw): return _leaf(True, "L", "L", A, B, **kw)      # src/trsm.jl:128
def LeftUpperTRSM(A, B, **kw): return _leaf(True, "L", "U", A, B, **kw)      # src/trsm.jl:134
def RightLowerTRSM(A, B, **kw): return _leaf(True, "R", "L", A, B, **kw)     # src/trsm.jl:140
def RightUpperTRSM(A, B, **kw): return _leaf(True, "R", "U", A, B, **kw)     # src/trsm.jl:146
def LeftLowerTRMM(A, B, **kw): return _leaf(False, "L", "L", A, B, **kw)     # src/trmm.jl:332
def LeftUpperTRMM(A, B, **kw): return _leaf(False, "L", "U", A, B, **kw)     # src/trmm.jl:352
def RightLowerTRMM(A, B, **kw): return _leaf(False, "R", "L", A, B, **kw)    # src/trmm.jl:367
def RightUpperTRMM(A, B, **kw): return _leaf(False, "R", "U", A, B, **kw)    # src/trmm.jl:384


def _gemm(C, A, B, sign: int, transa="N", transb="N", stream=None, handle=None):
    h = _handle_for(handle, C, stream)
    pa, ar, ac, lda, dta = _desc(A)
    pb, br, bc, ldb, dtb = _desc(B)
    pc, M, N, ldc, dtc = _desc(C)
    K = ac if transa == "N" else ar
    if not (dta == dtb == dtc):
        raise NextLAError("GEMM operands must share one element type")
    _check(load_library().nla_gemm_update(h._h, dtc, _ch(transa), _ch(transb), M, N, K, sign, pa, lda, pb, ldb, pc, ldc, _stream_ptr(stream, h.device)), h._h)
    return C


def GEMM_ADD(A, B, C, **kw):
    """GEMM_ADD!(A, B, C): C += A*B -- src/matmul.jl:69-74."""
    return _gemm(C, A, B, +1, **kw)


def GEMM_SUB(A, B, C, **kw):
    """GEMM_SUB!(A, B, C): A -= B*C -- src/matmul.jl:76-81."""
    return _gemm(A, B, C, -1, **kw)


def unified_trxm(side: str, uplo: str, transpose: str, diag: str, alpha: float, func: str, A, B, stream=None, handle: Optional[Handle] = None):
    """unified_rectrxm with the BLAS `diag` flag (nla_trxm): diag = 'U' treats the diagonal of A as ones without reading it."""
    h = _handle_for(handle, A, stream)
    pa, ar, ac, lda, dta = _desc(A)
    pb, br, bc, ldb, dtb = _desc(B)
    if dta != dtb or ar != ac:
        raise NextLAError("A must be square and share B's element type")
    n = ar
    m = bc if side == "L" else br
    if (side == "L" and br != n) or (side == "R" and bc != n):
        raise NextLAError("dimension mismatch between A and B")
    if dta in (NLA_C64, NLA_C128):
        a = complex(alpha)
        rc = load_library().nla_rectrxm_complex(h._h, _ch(side), _ch(uplo), _ch(transpose), _ch(diag), _ch(func), dta, n, m, a.real, a.imag, pa, lda,
                                                pb, ldb, _stream_ptr(stream, h.device))
        _check(rc, h._h)
        return B
    rc = load_library().nla_trxm(h._h, _ch(side), _ch(uplo), _ch(transpose), _ch(diag), _ch(func), dta, n, m, float(alpha), pa, lda, pb, ldb,
                                 _stream_ptr(stream, h.device))
    _check(rc, h._h)
    return B


def trsm(side, uplo, transa, diag, A, B, alpha=1.0, **kw):
    """trsm(side, uplo, transa, diag, A, B, alpha) -- src/trsm.jl:186-205.  Unlike the reference (which ignores `transa` and `diag`
    and is limited to one leaf) this honours both and recurses."""
    return unified_trxm(side, uplo, transa, diag, alpha, "S", A, B, **kw)


def trmm(side, uplo, transa, diag, A, B, alpha=1.0, **kw):
    """trmm(side, uplo, transa, diag, A, B, alpha) -- src/trmm.jl:430-448."""
    return unified_trxm(side, uplo, transa, diag, alpha, "M", A, B, **kw)


def laswp(A, first: int, last: int, ipiv, incx: int = 1, stream=None, handle: Optional[Handle] = None):
    """laswp(A, first, last, ipiv, incx) -- src/lu.jl:470-530 on a device matrix: swap rows i and ipiv[i] (1-based) for i = first..last.
    `ipiv` is a CUDA int64 vector."""
    import torch

    h = _handle_for(handle, A, stream)
    pa, rows, cols, lda, dta = _desc(A)
    if not (ipiv.is_cuda and ipiv.dtype == torch.int64 and ipiv.is_contiguous()):
        raise NextLAError("ipiv must be a contiguous CUDA int64 vector")
    if last > ipiv.numel():
        raise NextLAError("ipiv is shorter than `last`")
    _check(load_library().nla_laswp(h._h, dta, rows, cols, pa, lda, int(first), int(last), ipiv.data_ptr(), int(incx), _stream_ptr(stream, h.device)), h._h)
    return A


def getrf2_update(A, n1: int, ipiv, **kw):
    """The device steps of one level of the reference's recursive LU (getrf2!, src/lu.jl:274-280) after the left panel
    A[:, :n1] has been factored with pivots ipiv[:n1]:  laswp on the right part, A12 <- L11^-1 A12 (unit lower), A22 <- A22 - A21 A12."""
    m, n = A.shape
    laswp(A[:, n1:], 1, n1, ipiv, 1, **kw)                              # src/lu.jl:274
    trsm("L", "L", "N", "U", A[:n1, :n1], A[:n1, n1:], 1.0, **kw)        # src/lu.jl:277
    if m > n1:
        GEMM_SUB(A[n1:, n1:], A[n1:, :n1], A[:n1, n1:], **kw)            # src/lu.jl:280
    return A


def getrf2(A, ipiv=None, info=None, stream=None, handle: Optional[Handle] = None):
    """getrf2!(A, ipiv, info) -- src/lu.jl:185-299: recursive LU with partial pivoting of the device matrix A (Float64 / Float32,
    column-major view), in place and entirely on the device (nla_getrf2).  Returns (A, ipiv, info): `ipiv` a CUDA int64 vector of
    min(m, n) 1-based pivots, `info` a CUDA int32 scalar tensor (0, or the first i with U[i, i] == 0) -- the call is asynchronous, so
    read `info.item()` when the value is needed."""
    import torch

    h = _handle_for(handle, A, stream)
    pa, rows, cols, lda, dta = _desc(A)
    k = min(rows, cols)
    if ipiv is None:
        ipiv = torch.empty(k, dtype=torch.int64, device=A.device)
    if info is None:
        info = torch.zeros((), dtype=torch.int32, device=A.device)
    if not (ipiv.is_cuda and ipiv.dtype == torch.int64 and ipiv.is_contiguous() and ipiv.numel() >= k):
        raise NextLAError("ipiv must be a contiguous CUDA int64 vector of at least min(m, n) entries")
    if not (info.is_cuda and info.dtype == torch.int32 and info.numel() == 1):
        raise NextLAError("info must be a CUDA int32 scalar")
    _check(load_library().nla_getrf2(h._h, dta, rows, cols, pa, lda, ipiv.data_ptr(), info.data_ptr(), _stream_ptr(stream, h.device)), h._h)
    return A, ipiv, info


def lauum(uplo: str, A, ib: int = 1024, stream=None, handle: Optional[Handle] = None):
    """lauum!(uplo, n, A, ib) -- src/lauum.jl:52-186: A := L^H * L (uplo 'L') or U * U^H (uplo 'U'), the triangular factor stored in
    the `uplo` triangle of the device matrix A, result in the same triangle; the opposite triangle is neither read nor written
    (the reference multiplies the full diagonal blocks, i.e. assumes it is zero).  One call into the library (nla_lauum): the block
    loop, the trmm / GEMM steps and the triangle-masked diagonal-block products all run there."""
    h = _handle_for(handle, A, stream)
    pa, rows, cols, lda, dta = _desc(A)
    if rows != cols:
        raise NextLAError("A must be square")
    _check(load_library().nla_lauum(h._h, _ch(uplo), dta, rows, pa, lda, int(ib), _stream_ptr(stream, h.device)), h._h)
    return A


class MultiGPU:
    """Single-process multi-GPU driver (nla_mg_* in the C ABI): one host thread, every GPU of the box; the library owns the NCCL
    communicators, streams, events and the replicas of A.  The torch.distributed path (sharded.py, one process per GPU) is what
    bench.py runs under torchrun; this is the same pipeline for a host that has no process group -- e.g. a Julia session."""

    def __init__(self, devices=None, ngpu: Optional[int] = None):
        import torch

        lib = load_library()
        if devices is None:
            devices = list(range(ngpu if ngpu is not None else torch.cuda.device_count()))
        self.devices = list(devices)
        arr = (ctypes.c_int * len(self.devices))(*self.devices)
        self._mg = ctypes.c_void_p()
        _check(lib.nla_mg_create(ctypes.byref(self._mg), len(self.devices), arr))

    def handle_option(self, i: int, key: str, value: int):
        h = load_library().nla_mg_handle(self._mg, i)
        _check(load_library().nla_set_option(h, key.encode(), int(value)))

    def _shards(self, side, shards, host: bool):
        n_sh = len(self.devices)
        if len(shards) != n_sh:
            raise NextLAError(f"expected {n_sh} shards of B, got {len(shards)}")
        ptrs = (ctypes.c_void_p * n_sh)()
        ms = (ctypes.c_int64 * n_sh)()
        lds = (ctypes.c_int64 * n_sh)()
        for i, b in enumerate(shards):
            if host:
                if not b.flags.f_contiguous:
                    raise NextLAError("host shards must be Fortran-ordered ndarrays")
                ptrs[i], rows, cols, lds[i] = b.ctypes.data, b.shape[0], b.shape[1], max(1, b.shape[0])
            else:
                p, rows, cols, ld, _ = _desc(b)
                if b.device.index != self.devices[i]:
                    raise NextLAError(f"shard {i} must live on device {self.devices[i]}")
                ptrs[i], lds[i] = p, ld
            ms[i] = cols if side == "L" else rows
        return ptrs, ms, lds

    def rectrxm(self, side, uplo, transpose, alpha, func, A, root: int, B_shards):
        """A: column-major matrix on GPU self.devices[root]; B_shards[i]: column-major shard on GPU self.devices[i].  Asynchronous."""
        pa, ar, ac, lda, dta = _desc(A)
        if A.device.index != self.devices[root] or ar != ac:
            raise NextLAError("A must be square and live on the root device")
        ptrs, ms, lds = self._shards(side, B_shards, False)
        _check(load_library().nla_mg_rectrxm(self._mg, _ch(side), _ch(uplo), _ch(transpose), _ch(func), dta, ar, float(alpha), root, pa, lda, ptrs, ms, lds))
        return B_shards

    def rectrxm_host(self, side, uplo, transpose, alpha, func, A: np.ndarray, B_shards):
        """A and the shards of B are HOST arrays (Fortran order; page-locked memory for full speed).  Synchronous; shards are overwritten."""
        if not A.flags.f_contiguous:
            raise NextLAError("A must be a Fortran-ordered ndarray")
        ptrs, ms, lds = self._shards(side, B_shards, True)
        _check(load_library().nla_mg_rectrxm_host(self._mg, _ch(side), _ch(uplo), _ch(transpose), _ch(func), _NP2DT[A.dtype], A.shape[0], float(alpha),
                                                  A.ctypes.data, max(1, A.shape[0]), ptrs, ms, lds))
        return B_shards

    def sync(self):
        _check(load_library().nla_mg_sync(self._mg))

    def close(self):
        if self._mg:
            load_library().nla_mg_destroy(self._mg)
            self._mg = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def host_plan(side: str, uplo: str, transpose: str, func: str, n: int, cutoff: int = 1024, slabs: int = 1, a_resident: bool = False):
    """Host-only: the host-buffer pipeline's plan -- (ops, xfers, need, last): ops as in plan() with the large updates cut into 1024-wide
    pieces, xfers = [(kind, i, j)] host->device copies in issue order (kind 0: tile (i, j) of A, 1: chunk i of B for slab j),
    need[op][slab] = index of the last transfer the op waits for, last[chunk] = the op after which the chunk is downloaded."""
    lib = load_library()
    nx = ctypes.c_int64(0)
    args = (_ch(side), _ch(uplo), _ch(transpose), _ch(func), n, cutoff, slabs, 1 if a_resident else 0)
    nops = lib.nla_host_plan(*args, None, 0, None, 0, ctypes.byref(nx), None, None)
    if nops < 0:
        _check(-nops)
    nt = -(-n // 1024)
    ops = (ctypes.c_int64 * (6 * max(nops, 1)))()
    xf = (ctypes.c_int64 * (3 * max(nx.value, 1)))()
    need = (ctypes.c_int64 * (max(nops, 1) * slabs))()
    last = (ctypes.c_int64 * max(nt, 1))()
    lib.nla_host_plan(*args, ops, nops, xf, nx.value, ctypes.byref(nx), need, last)
    return ([tuple(ops[6 * i + j] for j in range(6)) for i in range(nops)], [tuple(xf[3 * i + j] for j in range(3)) for i in range(nx.value)],
            [[need[i * slabs + q] for q in range(slabs)] for i in range(nops)], [last[c] for c in range(nt)])


def plan(side: str, uplo: str, transpose: str, func: str, n: int, leaf: int = 0):
    """Host-only: the launch schedule as a list of (kind, c0, cn, k0, kn, carries_alpha) tuples (kind 0 = leaf, 1 = update)."""
    lib = load_library()
    cnt = lib.nla_plan(_ch(side), _ch(uplo), _ch(transpose), _ch(func), n, leaf, None, 0)
    if cnt < 0:
        _check(-cnt)
    buf = (ctypes.c_int64 * (6 * max(cnt, 1)))()
    lib.nla_plan(_ch(side), _ch(uplo), _ch(transpose), _ch(func), n, leaf, buf, cnt)
    return [tuple(buf[6 * i + j] for j in range(6)) for i in range(cnt)]
