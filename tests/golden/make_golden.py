"""Generates tests/golden/rectrxm_golden.npz.  Run from the repo root: python tests/golden/make_golden.py

The reference (Julia) cannot run here and ships no golden vectors, so the fixtures hold, for a fixed list of
seeded cases: the inputs, the output of the oracle (oracle/reference_port.py, literal restatement of the reference)
and the output of OpenBLAS trsm/trmm via SciPy (the routine the reference's own tests compare against,
test/unified_rectrxm.jl:36-40).  The committed file pins the oracle against accidental edits and lets the GPU box
(where /root/reference is absent) check the CUDA library against frozen numbers."""
import itertools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_port as rp  # noqa: E402

CASES = []
for dt in ("float64", "float32", "float16"):
    for (n, m) in [(16, 3), (40, 5), (48, 8)]:
        for side, uplo, trans, func in itertools.product("LR", "LU", "NT", "SM"):
            CASES.append((dt, n, m, side, uplo, trans, func, 1.0 if (n + m) % 2 else -0.5))
# one case above the reference's TRSM threshold (256) so that its GEMM_SUB! path is frozen too
CASES.append(("float64", 300, 4, "L", "L", "N", "S", 1.0))
CASES.append(("float64", 300, 4, "R", "U", "T", "S", 2.0))


def main():
    out = {}
    for i, (dt, n, m, side, uplo, trans, func, alpha) in enumerate(CASES):
        dtype = np.dtype(dt)
        A, B0 = rp.make_inputs(n, m, side, uplo, dtype, seed=4242 + i, recipe="scaled" if dt == "float16" else "reference")
        X = rp.unified_rectrxm(side, uplo, trans, alpha, func, A, B0.copy(order="F"))
        blas = rp.blas_reference(side, uplo, trans, alpha, func, A, B0)
        key = f"c{i:03d}"
        out[key + "_meta"] = np.array([dt, str(n), str(m), side, uplo, trans, func, repr(alpha)])
        out[key + "_A"] = A
        out[key + "_B0"] = B0
        out[key + "_oracle"] = X
        out[key + "_blas"] = blas.astype(np.float64)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "rectrxm_golden.npz"), **out)
    print(len(CASES), "cases written")


if __name__ == "__main__":
    main()
