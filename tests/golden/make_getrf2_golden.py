"""Generates tests/golden/getrf2_golden.npz.  Run from the repo root: python tests/golden/make_getrf2_golden.py

The reference (Julia) cannot run here and its LU tests hold no vectors (test/lu.jl builds random matrices and checks L*U ~ A[p, :]), so the
fixture holds, for a fixed list of seeded cases: the input, the output of the oracle (oracle/reference_port.getrf2, restatement of
src/lu.jl:185-299: factors, pivots, info) and of LAPACK getrf via SciPy.  It pins the oracle against accidental edits and lets the GPU box
check nla_getrf2 against frozen numbers."""
import os
import sys

import numpy as np
from scipy.linalg import lapack

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_port as rp  # noqa: E402

# (dtype, m, n, zero column or -1)
CASES = [("float64", 1, 1, -1), ("float64", 7, 1, -1), ("float64", 1, 7, -1), ("float64", 33, 33, -1), ("float64", 64, 40, -1),
         ("float64", 40, 64, -1), ("float64", 100, 110, -1), ("float64", 130, 97, -1), ("float64", 48, 48, 17), ("float64", 48, 48, 0),
         ("float32", 33, 33, -1), ("float32", 100, 90, -1), ("float32", 40, 64, -1)]


def main():
    out = {}
    for i, (dt, m, n, zc) in enumerate(CASES):
        rng = np.random.RandomState(9000 + i)
        A0 = np.asfortranarray((rng.rand(m, n) - 0.5).astype(dt))
        if zc >= 0:
            A0[:, zc] = 0
        LU = A0.copy(order="F")
        ipiv = np.zeros(min(m, n), dtype=np.int64)
        info = rp.getrf2(LU, ipiv)
        getrf = lapack.dgetrf if dt == "float64" else lapack.sgetrf
        lu_ref, piv_ref, info_ref = getrf(A0)
        assert info == info_ref and np.array_equal(ipiv - 1, piv_ref), (dt, m, n)
        key = f"c{i:02d}"
        out[key + "_meta"] = np.array([dt, str(m), str(n), str(zc)])
        out[key + "_A"] = A0
        out[key + "_oracle_lu"] = LU
        out[key + "_oracle_ipiv"] = ipiv
        out[key + "_info"] = np.array([info])
        out[key + "_lapack_lu"] = lu_ref
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "getrf2_golden.npz"), **out)
    print("wrote", len(CASES), "cases")


if __name__ == "__main__":
    main()
