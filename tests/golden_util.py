import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rectrxm_golden.npz")


def load_cases():
    z = np.load(GOLDEN)
    keys = sorted({k.split("_")[0] for k in z.files})
    for k in keys:
        dt, n, m, side, uplo, trans, func, alpha = z[k + "_meta"]
        yield dict(key=k, dtype=np.dtype(str(dt)), n=int(n), m=int(m), side=str(side), uplo=str(uplo), trans=str(trans), func=str(func),
                   alpha=float(alpha), A=np.asfortranarray(z[k + "_A"]), B0=np.asfortranarray(z[k + "_B0"]),
                   oracle=np.asfortranarray(z[k + "_oracle"]), blas=z[k + "_blas"])
