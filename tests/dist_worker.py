"""Worker of tests/test_dist_cpu.py: one of WORLD_SIZE processes (gloo, CPU).  Each rank owns a contiguous row block of
the operator and of every basis vector (the reference's PetscLayout split, bvbasic.c:129-134); the C host driver
(libb2kslepc) runs unchanged on every rank, its collectives go through the callback communicator
(B2KCommInitCallbacks → torch.distributed/gloo), the arithmetic is the oracle's host BV/Mat plug-in.
Rank 0 writes the results as JSON to argv[2]."""
import ctypes
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cpu_plugin as CP                      # noqa: E402
from oracle import slepc_oracle as O                      # noqa: E402
from slepc_b200 import slepc as SL                        # noqa: E402
from slepc_b200.slepc import S                            # noqa: E402

AR = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.c_int, ctypes.c_void_p)
SR = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double),
                      ctypes.c_int, ctypes.c_int, ctypes.c_void_p)


def np_view(ptr, n):
    return np.ctypeslib.as_array(ptr, shape=(n,)) if n > 0 else np.empty(0)


def make_callbacks(rank):
    def allreduce(buf, count, op, user):
        try:
            a = np_view(buf, count)
            t = torch.from_numpy(a.copy())
            dist.all_reduce(t, op=dist.ReduceOp.SUM if op == 0 else dist.ReduceOp.MAX)
            a[:] = t.numpy()
            return 0
        except Exception as e:                                # noqa: BLE001
            print("allreduce callback:", e, file=sys.stderr)
            return 1

    def sendrecv(sbuf, nsend, dest, rbuf, nrecv, src, user):
        try:
            reqs = []
            if nsend > 0:
                reqs.append(dist.isend(torch.from_numpy(np_view(sbuf, nsend).copy()), dest))
            r = None
            if nrecv > 0:
                r = torch.empty(nrecv, dtype=torch.float64)
                reqs.append(dist.irecv(r, src))
            for q in reqs:
                q.wait()
            if r is not None:
                np_view(rbuf, nrecv)[:] = r.numpy()
            return 0
        except Exception as e:                                # noqa: BLE001
            print("sendrecv callback:", e, file=sys.stderr)
            return 1

    return AR(allreduce), SR(sendrecv)


def gather_obj(obj):
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def gather_vec(vec):
    """global numpy vector from the row blocks"""
    parts = gather_obj(vec.get_values())
    return np.concatenate(parts)


def main():
    case, outpath = sys.argv[1], sys.argv[2]
    dist.init_process_group("gloo")
    rank, size = dist.get_rank(), dist.get_world_size()
    CP.load()
    ar, sr = make_callbacks(rank)
    S.B2KCommInitCallbacks(rank, size, ar, sr, None)
    res = {"case": case, "size": size}

    def eps_solve(A, nev, hermitian, which=None, v0=None, ncv=None):
        M = CP.mat_csr(A, rank, size, exchange=gather_obj)
        eps = SL.EPS(M, hermitian=hermitian)
        CP.use_cpu_bv(eps)
        S.EPSSetDimensions(eps.h, nev, ncv if ncv else SL.PETSC_DETERMINE, SL.PETSC_DETERMINE)
        if which:
            S.EPSSetWhichEigenpairs(eps.h, which)
        keep = []
        if v0 is not None:
            x, _ = M.create_vecs()
            rs, re = CP.split_rows(A.shape[0], size)[rank]
            x.set_values(v0[rs:re])
            S.EPSSetInitialSpace(eps.h, 1, (ctypes.c_void_p * 1)(x.h))
            keep.append(x)
        eps.solve()
        lam = [eps.eigenvalue(i) for i in range(eps.nconv)]
        errs = [eps.error(i) for i in range(eps.nconv)]
        xr, _ = M.create_vecs()
        X = []
        for i in range(min(eps.nconv, nev)):
            eps.eigenpair(i, xr)
            X.append(gather_vec(xr))
        return dict(nconv=eps.nconv, its=eps.its, lam=lam, errs=errs, X=np.array(X).T)

    if case == "bv":
        # BV known answers on a split basis: dot/dotvec/norm must be globally reduced, multinplace local
        n, k = 41, 6
        rs, re = CP.split_rows(n, size)[rank]
        rng = np.random.default_rng(5)
        Aglob = rng.standard_normal((n, k))
        X = CP.bv_cpu(re - rs, k, N=n, rstart=rs)
        X.from_numpy(Aglob[rs:re])
        nrm, lin = ctypes.c_double(), ctypes.c_int()
        norms = []
        for j in range(k):
            S.BVOrthonormalizeColumn(X.h, j, 0, ctypes.byref(nrm), ctypes.byref(lin))
            norms.append(nrm.value)
        Q = np.concatenate(gather_obj(X.to_numpy()), axis=0)
        Xo = O.BV(n, k)
        Xo.V[:] = Aglob
        norms_o = [Xo.orthonormalize_column(j)[0] for j in range(k)]
        res.update(orth=float(np.linalg.norm(Q.T @ Q - np.eye(k))), dq=float(np.abs(Q - Xo.V).max()),
                   dn=float(np.abs(np.array(norms) - np.array(norms_o)).max()))
    elif case == "bvsplit":
        # bv/tests/test10.c (its test list runs it on 2 ranks): split-phase operations posted together give exactly what the
        # regular, individually reduced ones give; the merged reduction of the BV-level pairs is ONE callback all-reduce
        n, k = 10, 5
        rs, re = CP.split_rows(n, size)[rank]
        Aglob = np.zeros((n, k))
        for j in range(k):
            for i in range(4):
                if i + j < n:
                    Aglob[i + j, j] = 3 * i + j - 2
        X = CP.bv_cpu(re - rs, k, N=n, rstart=rs)
        X.from_numpy(Aglob[rs:re])
        v = ctypes.c_void_p()
        S.BVCreateVec(X.h, ctypes.byref(v))
        S.VecSet(v, 1.0)
        z, zs = np.zeros(k + 6), np.zeros(k + 6)
        dptr = lambda a, off: ctypes.cast(a.ctypes.data + 8 * off, ctypes.POINTER(ctypes.c_double))
        nrm, w = ctypes.c_double(), ctypes.c_void_p()
        S.BVGetColumn(X.h, 0, ctypes.byref(w)); S.VecDot(w, v, dptr(z, 0)); S.BVRestoreColumn(X.h, 0, ctypes.byref(w))
        S.BVDotVec(X.h, v, dptr(z, 1))
        S.BVDotColumn(X.h, 2, dptr(z, 1 + k))
        S.BVGetColumn(X.h, 1, ctypes.byref(w)); S.VecNorm(w, SL.NORM_2, ctypes.byref(nrm)); S.BVRestoreColumn(X.h, 1, ctypes.byref(w))
        z[k + 3] = nrm.value
        S.BVNormVec(X.h, v, SL.NORM_2, ctypes.byref(nrm)); z[k + 4] = nrm.value
        S.BVNormColumn(X.h, 0, SL.NORM_2, ctypes.byref(nrm)); z[k + 5] = nrm.value
        S.BVGetColumn(X.h, 0, ctypes.byref(w))
        S.VecDotBegin(w, v, dptr(zs, 0))
        S.BVDotVecBegin(X.h, v, dptr(zs, 1))
        S.BVDotColumnBegin(X.h, 2, dptr(zs, 1 + k))
        S.VecDotEnd(w, v, dptr(zs, 0))
        S.BVRestoreColumn(X.h, 0, ctypes.byref(w))
        S.BVDotVecEnd(X.h, v, dptr(zs, 1))
        S.BVDotColumnEnd(X.h, 2, dptr(zs, 1 + k))
        S.BVGetColumn(X.h, 1, ctypes.byref(w))
        S.VecNormBegin(w, SL.NORM_2, ctypes.byref(nrm))
        S.BVNormVecBegin(X.h, v, SL.NORM_2, ctypes.byref(nrm))
        S.BVNormColumnBegin(X.h, 0, SL.NORM_2, ctypes.byref(nrm))
        S.VecNormEnd(w, SL.NORM_2, ctypes.byref(nrm)); zs[k + 3] = nrm.value
        S.BVRestoreColumn(X.h, 1, ctypes.byref(w))
        S.BVNormVecEnd(X.h, v, SL.NORM_2, ctypes.byref(nrm)); zs[k + 4] = nrm.value
        S.BVNormColumnEnd(X.h, 0, SL.NORM_2, ctypes.byref(nrm)); zs[k + 5] = nrm.value
        one = np.ones(n)
        ref = np.concatenate([[Aglob[:, 0] @ one], Aglob.T @ one, Aglob[:, :2].T @ Aglob[:, 2],
                              [np.linalg.norm(Aglob[:, 1]), np.sqrt(n), np.linalg.norm(Aglob[:, 0])]])
        res.update(diff=float(np.abs(z - zs).sum()), err=float(np.abs(z - ref).max()))
        S.VecDestroy(ctypes.byref(v))
    elif case in ("bvchol", "bvsvqb", "bvtsqr", "bvtsqrchol"):
        # BVOrthogonalize CHOL / SVQB (bvorthog.c:586-675) on a split basis with 2 leading columns: the Gram matrix is
        # globally reduced by BVDot, the k x k factorisation is replicated, BVMult/BVMultInPlace are local
        n, k, l = 57, 7, 2
        rs, re = CP.split_rows(n, size)[rank]
        rng = np.random.default_rng(7)
        Aglob = rng.standard_normal((n, k))
        name = case[2:]
        block = {"chol": SL.BV_ORTHOG_BLOCK_CHOL, "svqb": SL.BV_ORTHOG_BLOCK_SVQB, "tsqr": SL.BV_ORTHOG_BLOCK_TSQR,
                 "tsqrchol": SL.BV_ORTHOG_BLOCK_TSQRCHOL}[name]    # TSQR: the ranks' triangles are gathered and factored once more (bv.c)
        X = CP.bv_cpu(re - rs, k, N=n, rstart=rs)
        X.from_numpy(Aglob[rs:re])
        S.BVSetOrthogonalization(X.h, SL.BV_ORTHOG_CGS, SL.BV_ORTHOG_REFINE_IFNEEDED, 0.7071, block)
        R = SL.Mat.seqdense(np.zeros((k, k)))
        X.set_active(0, l)
        S.BVOrthogonalize(X.h, R.h)
        X.set_active(l, k)
        S.BVOrthogonalize(X.h, R.h)
        Q = np.concatenate(gather_obj(X.to_numpy()), axis=0)
        Xo = O.BV(n, k)
        Xo.V[:] = Aglob
        Xo.set_active(0, l)
        Xo.orthogonalize_block(name)
        Xo.set_active(l, k)
        Xo.orthogonalize_block(name)
        res.update(orth=float(np.linalg.norm(Q.T @ Q - np.eye(k))), dq=float(np.abs(Q - Xo.V).max()),
                   resid=float(np.linalg.norm(Aglob - Q @ R.dense_array())))
    elif case == "hep":
        nx = 24
        A = O.laplacian_2d(nx)
        r = eps_solve(A, 4, True)
        ref = O.eps_krylovschur(A, nx * nx, nev=4)
        X = r.pop("X")
        res.update(r, ref=list(ref.eigr[:ref.nconv]), ref_nconv=ref.nconv,
                   resid=[float(np.linalg.norm(A @ X[:, i] - r["lam"][i][0] * X[:, i])) for i in range(X.shape[1])])
    elif case == "nhep":
        m = 15
        A = O.markov_model(m)
        v0 = np.zeros(A.shape[0]); v0[:3] = 1.0
        r = eps_solve(A, 4, False, which=SL.EPS_LARGEST_REAL, v0=v0)
        r.pop("X")
        ref = O.eps_krylovschur(A, A.shape[0], nev=4, which="largest_real", hermitian=False, v0=v0)
        res.update(r, ref=list(ref.eigr[:ref.nconv]), ref_nconv=ref.nconv)
    elif case in ("svd", "svd_oneside"):
        Mr, Nc = 35, 30
        A = O.grcar_rect(Mr, Nc)
        AT = A.T.tocsr()
        Ma = CP.mat_csr(A, rank, size, exchange=gather_obj)
        Mt = CP.mat_csr(AT, rank, size, exchange=gather_obj)
        svd = SL.SVD(Ma, Mt)
        CP.use_cpu_bv(svd)
        S.SVDSetDimensions(svd.h, 4, SL.PETSC_DETERMINE, SL.PETSC_DETERMINE)
        S.SVDTRLanczosSetOneSide(svd.h, 1 if case == "svd_oneside" else 0)
        svd.solve()
        ref = O.svd_trlanczos(A, AT, Mr, Nc, nsv=4, oneside=(case == "svd_oneside"))
        res.update(nconv=svd.nconv, sigma=[svd.triplet(i) for i in range(svd.nconv)], errs=[svd.error(i) for i in range(svd.nconv)],
                   ref=list(ref.sigma[:ref.nconv]), ref_nconv=ref.nconv)
    else:
        raise SystemExit(f"unknown case {case}")
    dist.barrier()
    if rank == 0:
        json.dump(res, open(outpath, "w"))
    S.B2KCommReset()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
