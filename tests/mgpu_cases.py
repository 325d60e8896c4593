"""Row-partitioned correctness cases on N GPUs (one process per GPU, NCCL already initialised through slepc_b200.dist):
the product's BV type "b200" and Mat type "b200csr" against numpy/scipy answers computed on the spot.  Shared by
tests/mgpu_worker.py (pytest, `gpurun --gpus N`) and by bench.py, which runs them at every N > 1 and prints
`multi_gpu_parity: {cases, passed}` in its JSON line so that the driver's scaling run carries correctness at 2/4/8 ranks.
Every function is collective; rank 0's return value is the one that is checked."""
import ctypes

import numpy as np
import torch.distributed as dist

from slepc_b200 import dist as D
from slepc_b200 import matgen
from slepc_b200 import slepc as SL
from slepc_b200.slepc import S

pp = lambda a: a.ctypes.data_as(ctypes.c_void_p)


def stage(msg):
    """progress line per rank (stderr, flushed): tells where a hung run stopped"""
    import sys
    import time
    print(f"[mgpu rank {dist.get_rank()} t={time.time() % 1000:.2f}] {msg}", file=sys.stderr, flush=True)


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


def case_bv(rank, world, n=100003, k=9):
    """DGKS orthonormalisation of k random columns split over the ranks: every k-vector reduction crosses the GPUs"""
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
    # block orthogonalisation by the Householder tree (TSQR): tiles inside the CTAs, CTAs on the host, ranks in bv.c
    X.from_numpy(Ag[r0:r1])
    S.BVSetOrthogonalization(X.h, SL.BV_ORTHOG_CGS, SL.BV_ORTHOG_REFINE_IFNEEDED, 0.7071, SL.BV_ORTHOG_BLOCK_TSQR)
    Rt = SL.Mat.seqdense(np.zeros((k, k)))
    S.BVOrthogonalize(X.h, Rt.h)
    Qt = np.concatenate(gather(X.to_numpy()), axis=0)
    Rm = Rt.dense_array()
    sg = np.where(np.diag(R) < 0, -1.0, 1.0)
    tsqr = dict(orth=float(np.linalg.norm(Qt.T @ Qt - np.eye(k))), resid=float(np.linalg.norm(Ag - Qt @ Rm) / np.linalg.norm(Ag)),
                dr=float(np.linalg.norm(Rm - R * sg[:, None]) / np.linalg.norm(R)))
    X.destroy(); Rt.destroy()
    return dict(orth=float(np.linalg.norm(Q.T @ Q - np.eye(k))), span=float(np.linalg.norm(Q - Qr @ (Qr.T @ Q))),
                dn=float(np.abs(np.array(norms) - np.abs(np.diag(R))).max() / np.abs(np.diag(R)).max()),
                norms=norms, q_checksum=[float(x) for x in Q[::997].ravel()], tsqr=tsqr)


def case_bv_transports(rank, world):
    """the same orthonormalisation through the NVLink mailboxes (k_reduce_partials_xg) and through ncclAllReduce"""
    if not D.P2P:
        a = case_bv(rank, world)
        return dict(p2p=False, a=a, b=a)
    a = case_bv(rank, world)
    S.B2KCommDisableP2P()
    b = case_bv(rank, world)
    S.B2KCommEnableP2P()
    return dict(p2p=True, a=a, b=b)


def case_lap(rank, world, nx=96, ny=64):
    M = SL.Mat.laplacian(2, nx, ny)                       # nx (slowest) is split over the ranks
    eps = SL.EPS(M, hermitian=True)
    S.EPSSetDimensions(eps.h, 6, 24, SL.PETSC_DETERMINE)
    eps.solve()
    lam = [eps.eigenvalue(i)[0] for i in range(eps.nconv)]
    errs = [eps.error(i) for i in range(eps.nconv)]
    tx = 2 - 2 * np.cos(np.arange(1, nx + 1) * np.pi / (nx + 1)); ty = 2 - 2 * np.cos(np.arange(1, ny + 1) * np.pi / (ny + 1))
    an = np.sort((tx[:, None] + ty[None, :]).ravel())[::-1]
    res = dict(nconv=eps.nconv, lam=lam, errs=errs, analytic=list(an[:8]), its=eps.its)
    eps.destroy(); M.destroy()
    return res


def case_markov(rank, world, m=40):
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
    res = dict(nconv=eps.nconv, lam=[eps.eigenvalue(i)[0] for i in range(eps.nconv)], errs=[eps.error(i) for i in range(eps.nconv)],
               ref=list(ev[:4]))
    eps.destroy(); A.destroy()
    return res


def case_svd(rank, world, Mr=3000, Nc=1100):
    stage("svd: build A")
    A = csr_mat(lambda a, b: matgen.random_sparse_rows(Mr, Nc, 6, seed=3, r0=a, r1=b), Mr, Nc, rank, world)
    import scipy.sparse as sp
    rp, ci, v = matgen.random_sparse_rows(Mr, Nc, 6, seed=3)
    Ag = sp.csr_matrix((v, ci, rp), shape=(Mr, Nc))
    AT = Ag.T.tocsr(); AT.sort_indices()

    def at_rows(a, b):
        loc = AT[a:b]
        return loc.indptr.astype(np.int32), loc.indices.astype(np.int32), loc.data.astype(np.float64)
    stage("svd: build At")
    At = csr_mat(at_rows, Nc, Mr, rank, world)
    stage("svd: two-sided solve, explicit transpose")
    svd = SL.SVD(A, At)
    S.SVDSetDimensions(svd.h, 5, SL.PETSC_DETERMINE, SL.PETSC_DETERMINE)
    svd.solve()
    sref = np.linalg.svd(Ag.toarray(), compute_uv=False)
    res = dict(nconv=svd.nconv, sigma=[svd.triplet(i) for i in range(svd.nconv)], errs=[svd.error(i) for i in range(svd.nconv)],
               ref=list(sref[:5]))
    # the same solve without an explicit A^T: MatMultTranspose = local transpose products + reverse halo (scatter-add)
    stage("svd: two-sided solve, implicit transpose")
    svd2 = SL.SVD(A)
    S.SVDSetDimensions(svd2.h, 5, SL.PETSC_DETERMINE, SL.PETSC_DETERMINE)
    svd2.solve()
    res.update(nconv_impl=svd2.nconv, sigma_impl=[svd2.triplet(i) for i in range(svd2.nconv)],
               errs_impl=[svd2.error(i) for i in range(svd2.nconv)])
    # one-sided recurrence (fused split reduction) on the same matrix
    stage("svd: one-sided solve")
    svd3 = SL.SVD(A, At)
    S.SVDSetDimensions(svd3.h, 5, SL.PETSC_DETERMINE, SL.PETSC_DETERMINE)
    S.SVDTRLanczosSetOneSide(svd3.h, 1)
    svd3.solve()
    res.update(nconv_one=svd3.nconv, sigma_one=[svd3.triplet(i) for i in range(svd3.nconv)], errs_one=[svd3.error(i) for i in range(svd3.nconv)])
    stage("svd: destroy")
    for o in (svd, svd2, svd3, At, A):
        o.destroy()
    stage("svd: done")
    return res


def case_spmv_slab(rank, world, g=192):
    """MatMult of the slab-partitioned 3-D Laplacian (ghost planes over the halo) against the closed-form stencil on a
    hash-random x: element-wise, every rank checks its own rows"""
    M = SL.Mat.laplacian(3, g, g, g)
    x, y = M.create_vecs()
    plane = g * g
    p0, p1 = D.split(g, world, rank)
    idx = np.arange(max(p0 - 1, 0) * plane, min(p1 + 1, g) * plane, dtype=np.uint64)
    xa = matgen.hash_uniform(idx, 7)
    off = (p0 - max(p0 - 1, 0)) * plane
    x.set_values(xa[off:off + (p1 - p0) * plane])
    S.MatMult(M.h, x.h, y.h)
    yg = y.get_values()
    X = xa.reshape(-1, g, g)
    a, b = off // plane, off // plane + (p1 - p0)
    Y = 6.0 * X[a:b]
    Y -= X[a - 1:b - 1] if a > 0 else np.concatenate([np.zeros((1, g, g)), X[a:b - 1]])
    Y -= X[a + 1:b + 1] if b < X.shape[0] else np.concatenate([X[a + 1:b], np.zeros((1, g, g))])
    Y[:, 1:, :] -= X[a:b, :-1, :]; Y[:, :-1, :] -= X[a:b, 1:, :]
    Y[:, :, 1:] -= X[a:b, :, :-1]; Y[:, :, :-1] -= X[a:b, :, 1:]
    err = float(np.abs(yg - Y.ravel()).max())
    errs = gather(err)
    for o in (x, y, M):
        o.destroy()
    return dict(max_err=max(errs), per_rank=errs)


def verify(case, r):
    """the assertions of tests/test_multi_gpu.py as a function: returns (ok, message)"""
    try:
        if case == "bv":
            assert r["orth"] < 1e-13 and r["span"] < 1e-10 and r["dn"] < 1e-12, (r["orth"], r["span"], r["dn"])
            t = r["tsqr"]
            assert t["orth"] < 1e-13 and t["resid"] < 1e-14 and t["dr"] < 1e-12, t
        elif case == "bv_transports":
            a, b = r["a"], r["b"]
            assert np.allclose(a["norms"], b["norms"], rtol=1e-14, atol=0)
            assert np.allclose(a["q_checksum"], b["q_checksum"], rtol=0, atol=1e-14)
            assert a["orth"] < 1e-13 and b["orth"] < 1e-13
        elif case == "lap":
            assert r["nconv"] >= 6
            an = np.array(r["analytic"])
            for x in r["lam"][:6]:
                assert np.min(np.abs(an - x)) < 1e-10 * abs(x), x
            assert max(r["errs"][:6]) < 5e-8
        elif case == "markov":
            assert r["nconv"] >= 4
            assert np.allclose(r["lam"][:4], r["ref"][:4], rtol=1e-9, atol=0), (r["lam"][:4], r["ref"][:4])
            assert max(r["errs"][:4]) < 5e-8
        elif case == "svd":
            for sfx in ("", "_impl", "_one"):
                assert r["nconv" + sfx] >= 5, sfx
                assert np.allclose(r["sigma" + sfx][:5], r["ref"][:5], rtol=1e-10, atol=0), (sfx, r["sigma" + sfx][:5], r["ref"][:5])
                assert max(r["errs" + sfx][:5]) < 5e-8, sfx
        elif case == "spmv_slab":
            assert r["max_err"] <= 1e-13 * 7 * 6, r
        else:
            raise AssertionError(f"unknown case {case}")
    except AssertionError as e:                       # noqa: BLE001
        return False, f"{case}: {e}"
    return True, ""


CASES = {"bv": case_bv, "bv_transports": case_bv_transports, "lap": case_lap, "markov": case_markov, "svd": case_svd,
         "spmv_slab": case_spmv_slab}


def run_all(rank, world, names=None):
    """collective; returns {"cases": n, "passed": m, "failed": [...]} on every rank (rank 0's verdict is broadcast)"""
    names = list(names or CASES)
    failed = []
    for name in names:
        try:
            r = CASES[name](rank, world)
            ok, msg = verify(name, r) if rank == 0 else (True, "")
        except Exception as e:                        # noqa: BLE001
            ok, msg = False, f"{name}: raised {type(e).__name__}: {e}"
        oks = gather((ok, msg))
        bad = [m for o, m in oks if not o]
        if bad:
            failed.append(bad[0][:300])
    return dict(cases=len(names), passed=len(names) - len(failed), failed=failed, names=names, ranks=world)
