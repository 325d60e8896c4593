#!/usr/bin/env python
"""Block orthogonalisation timings on one B200 through the host API (BVOrthogonalize, bvorthog.c:682-765): CHOL / SVQB (Gram matrix +
restart GEMM), TSQRCHOL (Householder tree for R + restart GEMM), TSQR (Householder tree forward + backward), column-by-column GS,
on n = 4096^2 rows.  Wall clock around the call with a device synchronisation on both sides (host LAPACK of the k x k and of the
stacked CTA triangles included).     python tools/kbench3.py > gpurun_out/kbench3.jsonl"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from slepc_b200 import slepc as SL              # noqa: E402
from slepc_b200.slepc import S                  # noqa: E402


def main():
    SL.initialize(0)
    n = 4096 * 4096
    names = {SL.BV_ORTHOG_BLOCK_GS: "gs", SL.BV_ORTHOG_BLOCK_CHOL: "chol", SL.BV_ORTHOG_BLOCK_SVQB: "svqb", SL.BV_ORTHOG_BLOCK_TSQR: "tsqr",
             SL.BV_ORTHOG_BLOCK_TSQRCHOL: "tsqrchol"}
    for k in (16, 32, 64):
        X = SL.BV.create(n, k, bvtype="b200")
        Y = SL.BV.create(n, k, bvtype="b200")
        for j in range(k):
            S.BVSetRandomColumn(X.h, j)
        M = SL.Mat.seqdense(np.zeros((k, k)))
        for block, name in names.items():
            if name == "gs" and k > 32:
                continue
            best = None
            for rep in range(3):
                S.BVCopy(X.h, Y.h)
                S.BVSetOrthogonalization(Y.h, SL.BV_ORTHOG_CGS, SL.BV_ORTHOG_REFINE_IFNEEDED, 0.7071, block)
                S.B2KDeviceSynchronize()
                t0 = time.perf_counter()
                S.BVOrthogonalize(Y.h, None)
                S.B2KDeviceSynchronize()
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            S.BVDot(Y.h, Y.h, M.h)
            orth = float(np.linalg.norm(M.dense_array() - np.eye(k)))
            print(json.dumps(dict(op="BVOrthogonalize", block=name, n=n, k=k, ms=round(best * 1e3, 3), basis_sweeps_equiv=round(best / (8.0 * n * k / 6.45e12), 2),
                                  orth=orth)), flush=True)
        for o in (X, Y, M):
            o.destroy()


if __name__ == "__main__":
    main()
