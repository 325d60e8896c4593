"""Worker of tests/test_multi_gpu.py: one process per GPU (NCCL), the product's BV type "b200" and Mat type "b200csr"
with rows split across the ranks.  Rank 0 writes the results as JSON to argv[2]."""
import ctypes
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from slepc_b200 import dist as D                          # noqa: E402
from slepc_b200 import matgen                             # noqa: E402
from slepc_b200 import slepc as SL                        # noqa: E402
from slepc_b200.slepc import S                            # noqa: E402

pp = lambda a: a.ctypes.data_as(ctypes.c_void_p)


def gather(x):
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, x)
    return out


def csr_mat(rows_fn, M, N, rank, world):
    r0, r1 = D.split(M, world, rank)
    c0, c1 = D.split(N, world, rank)
    rp, ci, v = rows_fn(r0, r1)
    A = SL.Mat()
    S.MatCreateB200CSR(M, N, r0, r1, pp(rp), pp(ci), pp(v), c0, c1, A.ref)
    D.setup_halo(A, N)
    return A


def main():
    case, outpath = sys.argv[1], sys.argv[2]
    rank, world = D.init()
    res = {"case": case, "world": world}
    if case == "bv":
        n, k = 100003, 9
        r0, r1 = D.split(n, world, rank)
        rng = np.random.default_rng(11)
        Ag = rng.standard_normal((n, k))
        X = SL.BV.create(r1 - r0, k, N=n, rstart=r0)
        X.from_numpy(Ag[r0:r1])
        nrm, lin = ctypes.c_double(), ctypes.c_int()
        norms = []
        for j in range(k):
            S.BVOrthonormalizeColumn(X.h, j, 0, ctypes.byref(nrm), ctypes.byref(lin))
            norms.append(nrm.value)
        Q = np.concatenate(gather(X.to_numpy()), axis=0)
        Qr, R = np.linalg.qr(Ag)
        res.update(orth=float(np.linalg.norm(Q.T @ Q - np.eye(k))), span=float(np.linalg.norm(Q - Qr @ (Qr.T @ Q))),
                   dn=float(np.abs(np.array(norms) - np.abs(np.diag(R))).max() / np.abs(np.diag(R)).max()),
                   norms=norms, q_checksum=[float(x) for x in Q[::997].ravel()])
    elif case == "lap":
        nx, ny = 96, 64                                      # nx (slowest) is split over the ranks
        M = SL.Mat.laplacian(2, nx, ny)
        eps = SL.EPS(M, hermitian=True)
        S.EPSSetDimensions(eps.h, 6, 24, SL.PETSC_DETERMINE)
        eps.solve()
        lam = [eps.eigenvalue(i)[0] for i in range(eps.nconv)]
        errs = [eps.error(i) for i in range(eps.nconv)]
        tx = 2 - 2 * np.cos(np.arange(1, nx + 1) * np.pi / (nx + 1)); ty = 2 - 2 * np.cos(np.arange(1, ny + 1) * np.pi / (ny + 1))
        an = np.sort((tx[:, None] + ty[None, :]).ravel())[::-1]
        res.update(nconv=eps.nconv, lam=lam, errs=errs, analytic=list(an[:8]), its=eps.its)
    elif case == "markov":
        m = 40
        N = matgen.markov_size(m)
        A = csr_mat(lambda a, b: matgen.markov_rows(m, a, b), N, N, rank, world)
        eps = SL.EPS(A, hermitian=False)
        S.EPSSetWhichEigenpairs(eps.h, SL.EPS_LARGEST_REAL)
        S.EPSSetDimensions(eps.h, 4, SL.PETSC_DETERMINE, SL.PETSC_DETERMINE)
        r0, r1 = D.split(N, world, rank)
        v0 = np.zeros(N); v0[:3] = 1.0
        x0, _ = A.create_vecs()
        x0.set_values(v0[r0:r1])
        S.EPSSetInitialSpace(eps.h, 1, (ctypes.c_void_p * 1)(x0.h))
        eps.solve()
        import scipy.sparse as sp
        rp, ci, v = matgen.markov_rows(m)
        Ad = sp.csr_matrix((v, ci, rp), shape=(N, N)).toarray()
        ev = np.linalg.eigvals(Ad)
        ev = np.sort(ev.real[np.abs(ev.imag) < 1e-12])[::-1]
        res.update(nconv=eps.nconv, lam=[eps.eigenvalue(i)[0] for i in range(eps.nconv)], errs=[eps.error(i) for i in range(eps.nconv)],
                   ref=list(ev[:4]))
    elif case == "svd":
        Mr, Nc = 3000, 1100
        A = csr_mat(lambda a, b: matgen.random_sparse_rows(Mr, Nc, 6, seed=3, r0=a, r1=b), Mr, Nc, rank, world)
        import scipy.sparse as sp
        rp, ci, v = matgen.random_sparse_rows(Mr, Nc, 6, seed=3)
        Ag = sp.csr_matrix((v, ci, rp), shape=(Mr, Nc))
        AT = Ag.T.tocsr(); AT.sort_indices()
        def at_rows(a, b):
            loc = AT[a:b]
            return loc.indptr.astype(np.int32), loc.indices.astype(np.int32), loc.data.astype(np.float64)
        At = csr_mat(at_rows, Nc, Mr, rank, world)
        svd = SL.SVD(A, At)
        S.SVDSetDimensions(svd.h, 5, SL.PETSC_DETERMINE, SL.PETSC_DETERMINE)
        svd.solve()
        sref = np.linalg.svd(Ag.toarray(), compute_uv=False)
        res.update(nconv=svd.nconv, sigma=[svd.triplet(i) for i in range(svd.nconv)], errs=[svd.error(i) for i in range(svd.nconv)],
                   ref=list(sref[:5]))
        # the same solve without an explicit A^T: MatMultTranspose = local transpose products + reverse halo (scatter-add)
        svd2 = SL.SVD(A)
        S.SVDSetDimensions(svd2.h, 5, SL.PETSC_DETERMINE, SL.PETSC_DETERMINE)
        svd2.solve()
        res.update(nconv_impl=svd2.nconv, sigma_impl=[svd2.triplet(i) for i in range(svd2.nconv)],
                   errs_impl=[svd2.error(i) for i in range(svd2.nconv)])
    else:
        raise SystemExit(f"unknown case {case}")
    res["p2p"] = bool(D.P2P)
    if rank == 0:
        json.dump(res, open(outpath, "w"))
    D.finalize()


if __name__ == "__main__":
    main()
