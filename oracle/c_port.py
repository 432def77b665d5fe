"""ctypes loader for the C/OpenMP restatement (oracle/nla_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libnla_oracle.so")
_DTYPES = {np.dtype(np.float64): 0, np.dtype(np.float32): 1, np.dtype(np.float16): 2}
_lib = None


def build() -> str:
    src = [os.path.join(_HERE, f) for f in ("nla_oracle.c", "nla_oracle_impl.h", "Makefile")]
    if not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.nla_oracle_rectrxm.restype = ctypes.c_int
        _lib.nla_oracle_rectrxm.argtypes = [ctypes.c_char] * 4 + [ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_double,
                                            ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64]
        _lib.nla_oracle_num_threads.restype = ctypes.c_int
    return _lib


def num_threads() -> int:
    return lib().nla_oracle_num_threads()


def unified_rectrxm(side: str, uplo: str, transpose: str, alpha: float, func: str, A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """Same call shape as the reference's unified_rectrxm! (src/rectrxm.jl:43).  A, B column-major (Fortran order); B in place."""
    assert A.flags.f_contiguous and B.flags.f_contiguous and A.dtype == B.dtype
    n = A.shape[0]
    m = B.shape[1] if side == "L" else B.shape[0]
    rc = lib().nla_oracle_rectrxm(side.encode(), uplo.encode(), transpose.encode(), func.encode(), _DTYPES[A.dtype], n, m,
                                  float(alpha), A.ctypes.data, A.strides[1] // A.itemsize if A.ndim == 2 and A.shape[1] > 1 else max(1, A.shape[0]),
                                  B.ctypes.data, B.strides[1] // B.itemsize if B.shape[1] > 1 else max(1, B.shape[0]))
    if rc != 0:
        raise RuntimeError(f"nla_oracle_rectrxm failed: {rc}")
    return B
