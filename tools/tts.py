#!/usr/bin/env python
"""Time-to-solution of a full EPSSolve / SVDSolve on the B200 path (BASELINE.json metric).

    python tools/tts.py --case c2 [--scale 0.25] [--maxits N]            (1 GPU)
    python -m torch.distributed.run --nproc-per-node N … tools/tts.py --case c3      (N GPUs, rows split)

Prints one JSON line: wall seconds of the solve (first start-vector op to convergence; matrix build excluded and reported
separately), iterations, converged pairs, max relative residual ||Ax-kx||/|k| as the reference computes it
(epssolve.c:742-815), and the distance of every returned value to the analytic spectrum where one exists.
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="c2", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--scale", type=float, default=1.0, help="linear grid scale (c1-c3), m scale (c4), row scale (c5)")
    ap.add_argument("--maxits", type=int, default=0)
    ap.add_argument("--ncv", type=int, default=0)
    ap.add_argument("--tol", type=float, default=1e-8)
    ap.add_argument("--oneside", action="store_true", help="c5: one-sided thick-restart Lanczos (trlanczos.c:357-448)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    from slepc_b200 import _b2k, matgen
    from slepc_b200 import dist as D
    from slepc_b200 import slepc as SL
    from slepc_b200.slepc import S
    lib = _b2k.load()
    D.init()

    def split(N):
        base, rem = divmod(N, world)
        r0 = rank * base + min(rank, rem)
        return r0, r0 + base + (1 if rank < rem else 0)

    t0 = time.perf_counter()
    analytic = None
    info = {}
    if args.case in ("c1", "c2", "c3"):
        dim = {"c1": 1, "c2": 2, "c3": 3}[args.case]
        g = {"c1": 1000000, "c2": 4096, "c3": 512}[args.case]
        g = max(4, int(round(g * args.scale)))
        nev = {"c1": 10, "c2": 20, "c3": 10}[args.case]
        ncv = args.ncv or {"c1": 25, "c2": 64, "c3": 25}[args.case]
        M = SL.Mat.laplacian(dim, g, g if dim >= 2 else 1, g if dim >= 3 else 1)
        th = 2 - 2 * np.cos(np.arange(max(1, g - 40), g + 1) * np.pi / (g + 1))       # the top 40 1-D values are enough
        if dim == 1:
            analytic = th
        elif dim == 2:
            analytic = (th[:, None] + th[None, :]).ravel()
        else:
            analytic = (th[:, None, None] + th[None, :, None] + th[None, None, :]).ravel()
        solver = SL.EPS(M, hermitian=True)
        S.EPSSetDimensions(solver.h, nev, ncv, SL.PETSC_DETERMINE)
        info = dict(matrix=f"{dim}-D Laplacian grid {g}^{dim}", rows=g ** dim, nev=nev, ncv=ncv)
    elif args.case == "c4":
        m = max(10, int(round(6324 * args.scale)))
        N = matgen.markov_size(m)
        r0, r1 = split(N)
        rp, ci, v = matgen.markov_rows(m, r0, r1)
        M = SL.Mat()
        pp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        S.MatCreateB200CSR(N, N, r0, r1, pp(rp), pp(ci), pp(v), r0, r1, M.ref)
        D.setup_halo(M, N)
        solver = SL.EPS(M, hermitian=False)
        S.EPSSetWhichEigenpairs(solver.h, SL.EPS_LARGEST_REAL)
        S.EPSSetDimensions(solver.h, 8, args.ncv or SL.PETSC_DETERMINE, SL.PETSC_DETERMINE)
        x0, _ = M.create_vecs()
        v0 = np.zeros(r1 - r0)
        for gidx in (0, 1, 2):                                     # ex5.c:74-83: e0+e1+e2
            if r0 <= gidx < r1:
                v0[gidx - r0] = 1.0
        x0.set_values(v0)
        S.EPSSetInitialSpace(solver.h, 1, (ctypes.c_void_p * 1)(x0.h))
        info = dict(matrix=f"Markov model m={m}", rows=N, nev=8, ncv=args.ncv or "default")
    else:
        Mr = int(round(5e7 * args.scale)); Nc = int(round(1e7 * args.scale))
        r0, r1 = split(Mr)
        c0, c1 = split(Nc)
        rp, ci, v = matgen.random_sparse_rows(Mr, Nc, 20, r0=r0, r1=r1)
        M = SL.Mat()
        pp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        S.MatCreateB200CSR(Mr, Nc, r0, r1, pp(rp), pp(ci), pp(v), c0, c1, M.ref)
        D.setup_halo(M, Nc)
        del rp, ci, v
        solver = SL.SVD(M)                                         # A^T through MatMultTranspose (implicit) on > 1 rank
        S.SVDSetDimensions(solver.h, 10, args.ncv or 20, SL.PETSC_DETERMINE)
        if args.oneside:
            S.SVDTRLanczosSetOneSide(solver.h, 1)
        info = dict(matrix=f"random sparse {Mr}x{Nc}, 20 draws/row", rows=Mr, cols=Nc, nsv=10, ncv=args.ncv or 20,
                    variant="one-sided" if args.oneside else "two-sided")
    is_svd = args.case == "c5"
    if is_svd:
        S.SVDSetTolerances(solver.h, args.tol, args.maxits or SL.PETSC_CURRENT)
    else:
        S.EPSSetTolerances(solver.h, args.tol, args.maxits or SL.PETSC_CURRENT)
    S.B2KDeviceSynchronize()
    t_build = time.perf_counter() - t0
    ctx = S.B2KGetContext()
    n0 = ctypes.c_uint64(); lib.b2k_ctx_launches(ctx, ctypes.byref(n0))
    t1 = time.perf_counter()
    solver.solve()
    S.B2KDeviceSynchronize()
    t_solve = time.perf_counter() - t1
    n1 = ctypes.c_uint64(); lib.b2k_ctx_launches(ctx, ctypes.byref(n1))
    nconv = solver.nconv
    vals, errs = [], []
    for i in range(nconv):
        vals.append(solver.triplet(i) if is_svd else solver.eigenvalue(i)[0])
        errs.append(solver.error(i))
    out = dict(case=args.case, n_gpus=world, reductions=("peer-memory one-shot" if D.P2P else ("nccl" if world > 1 else "local")), **info, tol=args.tol, seconds_solve=t_solve, seconds_build=t_build, its=solver.its,
               reason=solver.reason, nconv=nconv, values=vals[:12], max_rel_residual=max(errs) if errs else None,
               kernel_launches=n1.value - n0.value)
    if not is_svd:
        bv = solver.bv()
        out["gs_passes"], out["matmults"] = bv.counters()
    if analytic is not None and vals:
        out["max_rel_dist_to_analytic"] = float(max(np.min(np.abs(analytic - x)) / abs(x) for x in vals))
    if rank == 0:
        print(json.dumps(out), flush=True)
    D.finalize()


if __name__ == "__main__":
    main()
