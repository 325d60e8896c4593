"""Debug helper (test infrastructure): Gram-Schmidt variants of the b200 BV type against the numpy oracle on a nearly dependent basis.
    python tests/dbg_orthog.py        (needs a GPU)"""
import sys, ctypes, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import slepc_oracle as O
from slepc_b200 import slepc as SL
from slepc_b200.slepc import S, c_dbl, c_int
SL.initialize(0)
for fuse in ("0", "1", "2"):
    os.environ["B2K_BV_FUSE"] = fuse
    n, k = 3001, 12
    rng = np.random.default_rng(11)
    A = rng.standard_normal((n, k))
    A[:, 5] = A[:, 4] + 1e-6 * rng.standard_normal(n)
    A[:, 9] = A[:, :3] @ np.array([1.0, -2.0, 0.5]) + 1e-9 * rng.standard_normal(n)
    X = SL.BV.create(n, k, bvtype="b200"); X.from_numpy(A)
    Xo = O.BV(n, k); Xo.V[:] = A
    nrm, lin = c_dbl(), c_int()
    H = np.zeros(k + 1)
    for j in range(k):
        p0 = X.counters()[0]
        S.BVOrthogonalizeColumn(X.h, j, H.ctypes.data_as(ctypes.c_void_p), ctypes.byref(nrm), ctypes.byref(lin))
        np0 = Xo.npasses
        ho, no, lo = Xo.orthogonalize_column(j)
        d = np.abs(H[:j] - ho)
        print(fuse, j, "passes", X.counters()[0] - p0, Xo.npasses - np0, "nrm", nrm.value, no, "maxdiff", d.max() if j else 0, "at", d.argmax() if j else -1, H[:j][d.argmax()] if j else 0, ho[d.argmax()] if j else 0)
        S.BVScaleColumn(X.h, j, 1.0 / nrm.value); Xo.scale_column(j, 1.0 / no)
