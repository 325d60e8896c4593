"""Host logic of the product (libb2kslepc: BV front-end, DGKS driver, Arnoldi/Lanczos, DS, Krylov-Schur,
TRLanczos) exercised on CPU through the oracle's host-memory BV/Mat plug-in, against
  (a) the reference's golden outputs (same files as tests/test_oracle_golden.py), and
  (b) the numpy oracle on the same seeded inputs.
The product's own BV type needs a GPU and is covered by tests/test_slepc_gpu.py (-m gpu); here the C host
driver is what is under test, the arithmetic comes from oracle/oracle_cpu.c.
"""
import ctypes
import os

import numpy as np
import pytest

from oracle import cpu_plugin as CP
from oracle import slepc_oracle as O
from slepc_b200 import slepc as SL
from slepc_b200.slepc import S, c_dbl, c_int

import bv_scenarios as SC

EPS = np.finfo(float).eps


@pytest.fixture(scope="module", autouse=True)
def _plugin():
    CP.load()


def make_bv(n, m):
    return CP.bv_cpu(n, m)


# ---- BV known-answer tests of the reference (bv/tests/test1,2,4,13) through the C front-end ------------
def test_bv_test1():
    SC.scenario_test1(make_bv)


def test_bv_test2():
    SC.scenario_test2(make_bv)


def test_bv_test4():
    SC.scenario_test4(make_bv)


def test_bv_test4_trans():
    SC.scenario_test4(make_bv, trans=True)


@pytest.mark.parametrize("block", [SL.BV_ORTHOG_BLOCK_GS, SL.BV_ORTHOG_BLOCK_CHOL, SL.BV_ORTHOG_BLOCK_SVQB, SL.BV_ORTHOG_BLOCK_TSQR,
                                   SL.BV_ORTHOG_BLOCK_TSQRCHOL])
@pytest.mark.parametrize("resid", [False, True])
def test_bv_test11_block_orthogonalize(block, resid):
    """bv/tests/test11.c with every -bv_orthog_block of its test list (:255-268: gs, chol, tsqr, tsqrchol, svqb)"""
    SC.scenario_test11(make_bv, block, resid=resid)


@pytest.mark.parametrize("block", [SL.BV_ORTHOG_BLOCK_TSQR, SL.BV_ORTHOG_BLOCK_TSQRCHOL])
def test_bv_block_orthogonalize_tsqr_other_shapes(block):
    SC.scenario_test11(make_bv, block, n=180, l=0, k=7, resid=True)
    SC.scenario_test11(make_bv, block, n=4099, l=3, k=17, resid=True)


def test_bv_norm_types():
    SC.scenario_norms(make_bv)


def test_bv_test13():
    SC.scenario_test13(make_bv)


def test_bv_test10_split_reductions():
    SC.scenario_test10(make_bv)


def test_bv_test3_nonstandard_inner_product():
    SC.scenario_test3(make_bv, CP.mat_csr)


def test_bv_test7_matmult():
    SC.scenario_test7(make_bv, CP.mat_csr)


def test_bv_test12_rank_deficient_block_gs():
    SC.scenario_test12(make_bv)


def test_bv_test18_normalize():
    SC.scenario_test18(make_bv, CP.mat_csr)


def test_bv_errors():
    SC.scenario_errors(make_bv)


@pytest.mark.parametrize("otype,refine", [(SL.BV_ORTHOG_CGS, SL.BV_ORTHOG_REFINE_IFNEEDED), (SL.BV_ORTHOG_CGS, SL.BV_ORTHOG_REFINE_ALWAYS),
                                          (SL.BV_ORTHOG_CGS, SL.BV_ORTHOG_REFINE_NEVER), (SL.BV_ORTHOG_MGS, SL.BV_ORTHOG_REFINE_IFNEEDED)])
def test_bv_orthog_variants_match_oracle(otype, refine):
    SC.scenario_orthog_vs_oracle(make_bv, otype, refine, n=300, k=12, tol=1e-12)


# ---- solvers ------------------------------------------------------------------------------------------
def fmt5(v):
    return [f"{x:.5f}" for x in v]


def solve_eps(A, nev, hermitian=True, ncv=None, tol=None, which=None, v0=None, lock=True, mpd=None):
    M = CP.mat_csr(A)
    eps = SL.EPS(M, hermitian=hermitian)
    CP.use_cpu_bv(eps)
    S.EPSSetDimensions(eps.h, nev, ncv if ncv else SL.PETSC_DETERMINE, mpd if mpd else SL.PETSC_DETERMINE)
    if tol:
        S.EPSSetTolerances(eps.h, tol, SL.PETSC_CURRENT)
    if which:
        S.EPSSetWhichEigenpairs(eps.h, which)
    if not lock:
        S.EPSKrylovSchurSetLocking(eps.h, 0)
    keep = [M]
    if v0 is not None:
        r, _ = M.create_vecs()
        r.set_values(v0)
        arr = (ctypes.c_void_p * 1)(r.h)
        S.EPSSetInitialSpace(eps.h, 1, arr)
        keep.append(r)
    eps.solve()
    eps._keep = keep
    return eps


def test_eps_test4_golden():
    A = O.laplacian_1d(30)
    eps = solve_eps(A, 4, tol=1000 * EPS)
    assert eps.reason > 0 and eps.nconv >= 4
    lam = [eps.eigenvalue(i)[0] for i in range(4)]
    assert fmt5(lam) == ["3.98974", "3.95906", "3.90828", "3.83792"]
    for i in range(4):
        assert eps.error(i) < 5 * 1000 * EPS


def test_eps_ex2_golden_and_oracle_parity():
    nx = 72
    A = O.laplacian_2d(nx)
    eps = solve_eps(A, 4, ncv=20)
    ref = O.eps_krylovschur(A, nx * nx, nev=4, ncv=20)
    assert eps.reason > 0 and eps.nconv >= 4
    # same deterministic start vector and same algorithm ⇒ same path as the numpy oracle
    assert eps.nconv == ref.nconv and eps.its == ref.its
    lam = np.array([eps.eigenvalue(i)[0] for i in range(eps.nconv)])
    assert np.allclose(lam, ref.eigr[ref.perm], rtol=1e-10, atol=0)
    gold = ["7.99630", "7.99074", "7.98519", "7.98150"]
    got = fmt5(lam[:4])
    assert got[0] == gold[0] and set(got) <= set(gold)
    for i in range(eps.nconv):
        assert eps.error(i) < 5e-8


@pytest.mark.parametrize("lock", [True, False])
def test_eps_ex5_markov_golden(lock):
    m = 15
    A = O.markov_model(m)
    N = m * (m + 1) // 2
    v0 = np.zeros(N)
    v0[:3] = 1.0
    eps = solve_eps(A, 4, hermitian=False, which=SL.EPS_LARGEST_REAL, v0=v0, lock=lock)
    assert eps.reason > 0 and eps.nconv >= 4
    lam = [eps.eigenvalue(i) for i in range(4)]
    assert fmt5([l[0] for l in lam]) == ["1.00000", "0.97137", "0.90423", "0.85714"]
    assert all(l[1] == 0.0 for l in lam)
    for i in range(4):
        assert eps.error(i) < 5e-8
    ref = O.eps_krylovschur(A, N, nev=4, which="largest_real", hermitian=False, v0=v0, lock=lock)
    assert eps.its == ref.its and eps.nconv == ref.nconv


def test_eps_nonsymmetric_complex_pairs():
    """random nonsymmetric matrix: complex conjugate pairs must stay together (DSSort_NHEP_Total, final sort)"""
    import scipy.sparse as sp
    rng = np.random.default_rng(3)
    n = 200
    A = (sp.random(n, n, density=0.05, random_state=5, format="csr") + sp.diags(rng.standard_normal(n))).tocsr()
    eps = solve_eps(A, 6, hermitian=False, tol=1e-9)
    assert eps.reason > 0
    ev = np.linalg.eigvals(A.toarray())
    ev = ev[np.argsort(-np.abs(ev))]
    got = np.array([complex(*eps.eigenvalue(i)) for i in range(eps.nconv)])
    for g in got[:6]:
        assert np.min(np.abs(ev - g)) < 1e-7 * max(1.0, abs(g))
    r, _ = eps._keep[0].create_vecs()
    i_ = eps._keep[0].create_vecs()[0]
    for i in range(min(6, eps.nconv)):
        assert eps.error(i) < 1e-7
        re, im = eps.eigenpair(i, r, i_)
        x = r.get_values() + 1j * i_.get_values()
        assert np.linalg.norm(A @ x - complex(re, im) * x) < 1e-6 * abs(complex(re, im)) * np.linalg.norm(x)


def test_eps_shift_and_which():
    n = 100
    A = O.laplacian_1d(n)
    M = CP.mat_csr(A)
    eps = SL.EPS(M, hermitian=True)
    CP.use_cpu_bv(eps)
    st = ctypes.c_void_p()
    S.EPSGetST(eps.h, ctypes.byref(st))
    S.STSetShift(st, 1.5)
    S.EPSSetDimensions(eps.h, 3, SL.PETSC_DETERMINE, SL.PETSC_DETERMINE)
    S.EPSSetWhichEigenpairs(eps.h, SL.EPS_SMALLEST_REAL)
    S.EPSSetTolerances(eps.h, 1e-10, 2000)
    eps.solve()
    exact = np.sort(2 - 2 * np.cos(np.arange(1, n + 1) * np.pi / (n + 1)))
    assert eps.nconv >= 3
    lam = [eps.eigenvalue(i)[0] for i in range(3)]
    assert np.allclose(lam, exact[:3], rtol=1e-8)


def test_eps_argument_errors():
    A = O.laplacian_1d(30)
    M = CP.mat_csr(A)
    eps = SL.EPS(M)
    CP.use_cpu_bv(eps)
    S.EPSSetDimensions(eps.h, 4, 4, SL.PETSC_DETERMINE)
    with pytest.raises(SL.SlepcError, match="ncv must be at least nev\\+1"):
        eps.solve()
    eps2 = SL.EPS(M)
    with pytest.raises(SL.SlepcError, match="Must call EPSSolve"):
        eps2.nconv
    with pytest.raises(SL.SlepcError, match="Illegal value of nev"):
        S.EPSSetDimensions(eps2.h, 0, SL.PETSC_DETERMINE, SL.PETSC_DETERMINE)


@pytest.mark.parametrize("lock", [True, False])
def test_svd_test3_golden(lock):
    Mr, N = 35, 30
    A = O.grcar_rect(Mr, N)
    MA, MT = CP.mat_csr(A), CP.mat_csr(A.T.tocsr())
    svd = SL.SVD(MA, MT)
    CP.use_cpu_bv(svd)
    S.SVDSetDimensions(svd.h, 4, SL.PETSC_DETERMINE, SL.PETSC_DETERMINE)
    if not lock:
        S.SVDTRLanczosSetLocking(svd.h, 0)
    svd.solve()
    assert svd.reason > 0 and svd.nconv >= 4
    sig = [svd.triplet(i) for i in range(4)]
    assert fmt5(sig) == ["3.22175", "3.21797", "3.16825", "3.15128"]
    ref = O.svd_trlanczos(A, A.T.tocsr(), Mr, N, nsv=4, lock=lock)
    assert svd.its == ref.its and svd.nconv == ref.nconv
    for i in range(4):
        assert svd.error(i) < 5e-8


ONESIDE = [("cgs", SL.BV_ORTHOG_CGS, SL.BV_ORTHOG_REFINE_IFNEEDED), ("mgs", SL.BV_ORTHOG_MGS, SL.BV_ORTHOG_REFINE_IFNEEDED),
           ("always", SL.BV_ORTHOG_CGS, SL.BV_ORTHOG_REFINE_ALWAYS), ("never", SL.BV_ORTHOG_CGS, SL.BV_ORTHOG_REFINE_NEVER)]


@pytest.mark.parametrize("name,otype,oref", ONESIDE)
def test_svd_test3_oneside_golden(name, otype, oref):
    """test3.c 1_trlanczos_one / _one_mgs / _one_always (output/test3_1.out): SVDOneSideTRLanczosCGS/MGS, trlanczos.c:264-448"""
    Mr, N = 35, 30
    A = O.grcar_rect(Mr, N)
    MA, MT = CP.mat_csr(A), CP.mat_csr(A.T.tocsr())
    svd = SL.SVD(MA, MT)
    CP.use_cpu_bv(svd)
    S.SVDSetDimensions(svd.h, 4, SL.PETSC_DETERMINE, SL.PETSC_DETERMINE)
    S.SVDTRLanczosSetOneSide(svd.h, 1)
    for bv in svd.bvs():
        S.BVSetOrthogonalization(bv.h, otype, oref, 0.7071, 0)
    svd.solve()
    assert svd.reason > 0 and svd.nconv >= 4
    sig = [svd.triplet(i) for i in range(4)]
    assert fmt5(sig) == ["3.22175", "3.21797", "3.16825", "3.15128"]
    ref = O.svd_trlanczos(A, A.T.tocsr(), Mr, N, nsv=4, oneside=True, orthog=(otype, oref, 0.7071))
    assert svd.its == ref.its and svd.nconv == ref.nconv
    assert np.allclose(sig, ref.sigma[:4], rtol=1e-12)
    for i in range(4):
        assert svd.error(i) < 5e-8


def test_svd_wide_matrix_swaps():
    """M < N: SVDSetUp works with the transpose and swaps U/V (svdsetup.c:301-343)"""
    A = O.grcar_rect(35, 30).T.tocsr()          # 30 x 35
    MA, MT = CP.mat_csr(A), CP.mat_csr(A.T.tocsr())
    svd = SL.SVD(MA, MT)
    CP.use_cpu_bv(svd)
    S.SVDSetDimensions(svd.h, 3, SL.PETSC_DETERMINE, SL.PETSC_DETERMINE)
    svd.solve()
    assert svd.nconv >= 3
    sv = np.linalg.svd(A.toarray(), compute_uv=False)
    v, u = MA.create_vecs()
    for i in range(3):
        s = svd.triplet(i, u, v)
        assert abs(s - sv[i]) < 1e-8 * sv[i]
        uu, vv = u.get_values(), v.get_values()
        assert np.linalg.norm(A @ vv - s * uu) < 1e-7 * s
        assert svd.error(i) < 5e-8


def test_bv_split_phase_reductions():
    """BVNormColumnBegin/End, BVDotVecBegin/End, BVDotColumnBegin/End (bvglobal.c:188-260,350-440,703-800): queued local parts,
    ONE reduction at the first End, results equal to the blocking calls; Ends out of order are an error (the PetscSplitReduction
    fallback of a BV type without the *_begin/_end slots, which is what the CPU plug-in is)."""
    rng = np.random.default_rng(21)
    X, U = make_bv(60, 6), make_bv(35, 3)
    X.from_numpy(rng.standard_normal((60, 6)))
    U.from_numpy(rng.standard_normal((35, 3)))
    Xh, Uh = X.to_numpy(), U.to_numpy()
    a, b = c_dbl(), c_dbl()
    m1, m2 = np.zeros(5), np.zeros(4)
    X.set_active(0, 5)
    S.BVNormColumnBegin(U.h, 1, SL.NORM_2, ctypes.byref(a))
    v5 = SC.with_column(X, 5, lambda v: S.BVDotVecBegin(X.h, v, m1.ctypes.data_as(ctypes.c_void_p)))
    S.BVDotColumnBegin(X.h, 4, m2.ctypes.data_as(ctypes.c_void_p))
    S.BVNormColumnBegin(U.h, 2, SL.NORM_2, ctypes.byref(b))
    S.BVNormColumnEnd(U.h, 1, SL.NORM_2, ctypes.byref(a))
    S.BVDotVecEnd(X.h, None, m1.ctypes.data_as(ctypes.c_void_p))
    S.BVDotColumnEnd(X.h, 4, m2.ctypes.data_as(ctypes.c_void_p))
    S.BVNormColumnEnd(U.h, 2, SL.NORM_2, ctypes.byref(b))
    assert np.isclose(a.value, np.linalg.norm(Uh[:, 1]), rtol=1e-14) and np.isclose(b.value, np.linalg.norm(Uh[:, 2]), rtol=1e-14)
    assert np.allclose(m1, Xh[:, :5].T @ Xh[:, 5], rtol=1e-13, atol=1e-13)
    assert np.allclose(m2, Xh[:, :4].T @ Xh[:, 4], rtol=1e-13, atol=1e-13)
    # a second batch works after the first was fully collected, and a wrong order is refused
    S.BVNormColumnBegin(U.h, 0, SL.NORM_2, ctypes.byref(a))
    S.BVDotColumnBegin(X.h, 3, m2.ctypes.data_as(ctypes.c_void_p))
    with pytest.raises(SL.SlepcError):
        S.BVDotColumnEnd(X.h, 3, m2.ctypes.data_as(ctypes.c_void_p))
    SL.S.B2KClearError() if hasattr(SL.S, "B2KClearError") else None
    S.BVNormColumnEnd(U.h, 0, SL.NORM_2, ctypes.byref(a))
    S.BVDotColumnEnd(X.h, 3, m2.ctypes.data_as(ctypes.c_void_p))
    assert np.isclose(a.value, np.linalg.norm(Uh[:, 0]), rtol=1e-14)
    assert np.allclose(m2[:3], Xh[:, :3].T @ Xh[:, 3], rtol=1e-13, atol=1e-13)
    X.destroy(); U.destroy()


# ---- generalized symmetric-definite problems, B-inner product, shift-and-invert ---------------------------
def _ex13(n, m=None):
    import scipy.sparse as sp
    m = n if m is None else m
    A = O.laplacian_2d(m, n)
    B = sp.identity(n * m, format="csr") * 4.0
    return A, B


def test_eps_ex13_generalized_sinvert_golden_and_oracle_parity():
    """ex13 -eps_nev 4 -eps_ncv 22 -eps_tol 1e-5 -st_type sinvert through the C host driver (EPS_GHEP, BVSetMatrix, STSINVERT with the
    CG linear solves of ksp.c) on the CPU plug-in: the reference's golden 0.04051, 0.09963, 0.09963, 0.15875 and the numpy oracle
    (which solves with a sparse LU like the reference) to 1e-9."""
    A, B = _ex13(10)
    Am, Bm = CP.mat_csr(A), CP.mat_csr(B)
    eps = SL.EPS(Am, hermitian=True, B=Bm)
    CP.use_cpu_bv(eps)
    S.EPSSetDimensions(eps.h, 4, 22, SL.PETSC_DETERMINE)
    S.EPSSetTolerances(eps.h, 1e-5, SL.PETSC_CURRENT)
    eps.st_sinvert(0.0)
    eps.solve()
    assert eps.reason > 0 and eps.nconv >= 4
    lam = [eps.eigenvalue(i)[0] for i in range(eps.nconv)]
    assert [f"{x:.5f}" for x in lam[:4]] == ["0.04051", "0.09963", "0.09963", "0.15875"]
    ref = O.eps_krylovschur(A, 100, nev=4, ncv=22, tol=1e-5, B=B, sigma=0.0, sinvert=True)
    assert eps.nconv == ref.nconv and eps.its == ref.its
    assert np.allclose(lam, ref.eigr[ref.perm], rtol=1e-9, atol=0)
    assert max(eps.error(i) for i in range(4)) < 5e-5
    assert eps.ksp_iterations() > 0
    # eigenvectors are B-normalised
    x, _ = Am.create_vecs()
    bx, _ = Am.create_vecs()
    S.EPSGetEigenpair(eps.h, 0, None, None, x.h, None)
    S.MatMult(Bm.h, x.h, bx.h)
    d = c_dbl()
    S.VecDot(x.h, bx.h, ctypes.byref(d))
    assert abs(d.value - 1.0) < 1e-10


@pytest.mark.parametrize("sigma", [0.0, -0.3])
def test_eps_generalized_nontrivial_mass_matrix(sigma):
    """A x = k B x with a tridiagonal mass-like B (not a multiple of I): smallest eigenvalues by shift-and-invert against
    scipy's generalized dense solver; B-orthonormality of the basis"""
    import scipy.linalg as sla
    import scipy.sparse as sp
    n = 60
    A = O.laplacian_1d(n) * (n + 1.0)
    B = sp.diags([np.full(n - 1, 1.0), np.full(n, 4.0), np.full(n - 1, 1.0)], [-1, 0, 1], format="csr") / (6.0 * (n + 1.0))
    w = sla.eigh(A.toarray(), B.toarray(), eigvals_only=True)
    Am, Bm = CP.mat_csr(A), CP.mat_csr(B)
    eps = SL.EPS(Am, hermitian=True, B=Bm)
    CP.use_cpu_bv(eps)
    S.EPSSetDimensions(eps.h, 5, 20, SL.PETSC_DETERMINE)
    eps.st_sinvert(sigma)
    eps.solve()
    assert eps.nconv >= 5
    lam = np.array([eps.eigenvalue(i)[0] for i in range(5)])
    assert np.allclose(lam, w[:5], rtol=1e-9, atol=0)
    assert max(eps.error(i) for i in range(5)) < 5e-8


def test_eps_generalized_shift_largest():
    """EPS_GHEP with the default ST (shift 0): Op = B^{-1} A, largest eigenvalues, B-inner product"""
    import scipy.linalg as sla
    A, B = _ex13(9, 7)
    w = sla.eigh(A.toarray(), B.toarray(), eigvals_only=True)[::-1]
    Am, Bm = CP.mat_csr(A), CP.mat_csr(B)
    eps = SL.EPS(Am, hermitian=True, B=Bm)
    CP.use_cpu_bv(eps)
    S.EPSSetDimensions(eps.h, 3, 16, SL.PETSC_DETERMINE)
    eps.solve()
    assert eps.nconv >= 3
    lam = np.array([eps.eigenvalue(i)[0] for i in range(3)])
    assert np.allclose(lam, w[:3], rtol=1e-9, atol=0)
    assert max(eps.error(i) for i in range(3)) < 5e-8


def test_bv_set_matrix_inner_product():
    """BVSetMatrix (bvbasic.c:497): BVDotVec / BVDot / BVNormColumn / BVOrthogonalize* use <x,y> = y^T B x"""
    import scipy.sparse as sp
    rng = np.random.default_rng(5)
    n, k = 40, 5
    B = sp.diags([np.full(n - 1, -1.0), np.full(n, 3.0), np.full(n - 1, -1.0)], [-1, 0, 1], format="csr")
    Bm = CP.mat_csr(B)
    X = make_bv(n, k)
    Xh = rng.standard_normal((n, k))
    X.from_numpy(Xh)
    S.BVSetMatrix(X.h, Bm.h, 0)
    nrm = c_dbl()
    S.BVNormColumn(X.h, 2, SL.NORM_2, ctypes.byref(nrm))
    assert np.isclose(nrm.value, np.sqrt(Xh[:, 2] @ (B @ Xh[:, 2])), rtol=1e-13)
    M = SL.Mat.seqdense(np.zeros((k, k)))
    S.BVDot(X.h, X.h, M.h)
    assert np.allclose(M.dense_array(), Xh.T @ (B @ Xh), rtol=1e-12, atol=1e-12)
    for otype in (SL.BV_ORTHOG_CGS, SL.BV_ORTHOG_MGS):
        X.from_numpy(Xh)
        S.BVSetOrthogonalization(X.h, otype, SL.BV_ORTHOG_REFINE_IFNEEDED, 0.7071, SL.BV_ORTHOG_BLOCK_GS)
        lin = c_int()
        for j in range(k):
            S.BVOrthonormalizeColumn(X.h, j, 0, ctypes.byref(nrm), ctypes.byref(lin))
        Q = X.to_numpy()
        assert np.linalg.norm(Q.T @ (B @ Q) - np.eye(k)) < 1e-13
    X.destroy()


# ---- BVKrylovLoop_Private with a type that runs stretches of the recurrence on its own (ops.krylov_steps) ---------------------------
def _solve_with_krylov_steps(pattern, kind):
    """the CPU plug-in's test double of the krylov_steps slot (oracle_cpu.c BVKrylovSteps_CPU): completes pattern[i] steps on its
    i-th call, declines on 0, and on a negative entry leaves one more step half done (what a device-side breakdown looks like)"""
    import scipy.sparse as sp
    old = os.environ.get("ORACLE_KRYLOV_STEPS")
    if pattern is None:
        os.environ.pop("ORACLE_KRYLOV_STEPS", None)
    else:
        os.environ["ORACLE_KRYLOV_STEPS"] = pattern
    try:
        x0 = None
        if kind == "lanczos":
            A, herm, nev, ncv, which = O.laplacian_2d(30), True, 4, 16, None
        elif kind == "arnoldi":
            A, herm, nev, ncv, which = O.markov_model(25), False, 4, 14, SL.EPS_LARGEST_REAL
            x0 = np.zeros(A.shape[0]); x0[:3] = 1.0
        else:                                           # exact breakdown at the very first step: start vector = eigenvector
            n = 400
            A, herm, nev, ncv, which = sp.diags(np.repeat([1.0, 2.0, 3.5, 5.0, 9.0], n // 5)).tocsr(), True, 2, 12, None
            x0 = np.zeros(n); x0[0] = 1.0
        M = CP.mat_csr(A)
        eps = SL.EPS(M, hermitian=herm)
        CP.use_cpu_bv(eps)
        if which is not None:
            S.EPSSetWhichEigenpairs(eps.h, which)
        S.EPSSetDimensions(eps.h, nev, ncv, SL.PETSC_DETERMINE)
        keep = None
        if x0 is not None:
            keep, _ = M.create_vecs()
            keep.set_values(x0)
            S.EPSSetInitialSpace(eps.h, 1, (ctypes.c_void_p * 1)(keep.h))
        eps.solve()
        out = dict(its=eps.its, nconv=eps.nconv, reason=eps.reason, lam=[eps.eigenvalue(i) for i in range(eps.nconv)],
                   errs=[eps.error(i) for i in range(eps.nconv)], passes=eps.bv().counters()[0])
        for o in (eps, M) + ((keep,) if keep is not None else ()):
            o.destroy()
        return out
    finally:
        if old is None:
            os.environ.pop("ORACLE_KRYLOV_STEPS", None)
        else:
            os.environ["ORACLE_KRYLOV_STEPS"] = old


@pytest.mark.parametrize("kind", ["lanczos", "arnoldi", "breakdown"])
@pytest.mark.parametrize("pattern", ["1000", "3", "1,0,2", "-2,5", "0,0,7", "-1"])
def test_krylov_loop_resumes_after_a_type_run_stretch(kind, pattern):
    """whatever part of a cycle the BV type completes on its own (all of it, a few steps, nothing, or a stretch that ends in a step
    it disowns), BVMatLanczos / BVMatArnoldi finish the cycle step by step and ask again: iteration counts, converged pairs and
    eigenvalues are bit-identical to the plain loop of bvkrylov.c:56-226"""
    ref = _solve_with_krylov_steps(None, kind)
    got = _solve_with_krylov_steps(pattern, kind)
    assert ref["reason"] > 0 and ref["nconv"] >= 2
    assert (got["its"], got["nconv"], got["reason"]) == (ref["its"], ref["nconv"], ref["reason"])
    assert got["lam"] == ref["lam"] and got["errs"] == ref["errs"]


def test_eps_test1_ghep_b_orthonormality_host():
    """eps/tests/test1.c through the C host driver on the CPU plug-in: golden 21.89996, 21.65898, 21.28794, 20.82229
    (output/test1_1.out) and B-orthonormal eigenvectors to 10 tol"""
    import scipy.sparse as sp
    n = 18
    N = n * n
    A = O.laplacian_2d(n).tocsr()
    B = sp.diags(2.0 / np.log(np.arange(N) + 2.0)).tocsr()
    Am, Bm = CP.mat_csr(A), CP.mat_csr(B)
    eps = SL.EPS(Am, hermitian=True, B=Bm)
    CP.use_cpu_bv(eps)
    S.EPSSetDimensions(eps.h, 4, SL.PETSC_DETERMINE, SL.PETSC_DETERMINE)
    S.EPSSetTolerances(eps.h, 1e-10, SL.PETSC_CURRENT)
    eps.solve()
    assert eps.reason > 0 and eps.nconv >= 4
    assert [f"{eps.eigenvalue(i)[0]:.5f}" for i in range(4)] == ["21.89996", "21.65898", "21.28794", "20.82229"]
    ref = O.eps_krylovschur(A, N, nev=4, tol=1e-10, B=B)
    assert (eps.nconv, eps.its) == (ref.nconv, ref.its)
    x, _ = Am.create_vecs()
    X = []
    for i in range(eps.nconv):
        S.EPSGetEigenpair(eps.h, i, None, None, x.h, None)
        X.append(x.get_values().copy())
    X = np.array(X).T
    assert np.abs(X.T @ (B @ X) - np.eye(eps.nconv)).max() < 10 * 1e-10
    for o in (eps, x, Am, Bm):
        o.destroy()


def test_eps_test6_diagonal_host():
    import scipy.sparse as sp
    Am = CP.mat_csr(sp.diags(np.arange(1.0, 31.0)).tocsr())
    eps = SL.EPS(Am, hermitian=True)
    CP.use_cpu_bv(eps)
    S.EPSSetDimensions(eps.h, 4, SL.PETSC_DETERMINE, SL.PETSC_DETERMINE)
    S.EPSSetTolerances(eps.h, 1e-10, SL.PETSC_CURRENT)
    eps.solve()
    assert [f"{eps.eigenvalue(i)[0]:.5f}" for i in range(4)] == ["30.00000", "29.00000", "28.00000", "27.00000"]   # output/test6_1.out
    for o in (eps, Am):
        o.destroy()


def test_svd_test4_more_columns_than_rows_host():
    """svd/tests/test4.c (1_trlanczos) through the C host driver: M < N, so SVDSetUp works on the transposed pair and hands the
    vectors back in the caller's roles (u of length 20, v of length 22); golden 2.99254, same restarts as the oracle"""
    import scipy.sparse as sp
    m, n = 20, 22
    A = sp.lil_matrix((m, n))
    for i in range(m):
        A[i, i], A[i, i + 1] = 1.0, 2.0
    A = A.tocsr()
    Am, ATm = CP.mat_csr(A), CP.mat_csr(A.T.tocsr())
    svd = SL.SVD(Am, ATm)
    CP.use_cpu_bv(svd)
    S.SVDSetDimensions(svd.h, 1, 12, SL.PETSC_DETERMINE)
    S.SVDTRLanczosSetRestart(svd.h, 0.6)
    svd.solve()
    ref = O.svd_trlanczos(A.T.tocsr(), A, n, m, nsv=1, ncv=12, keep=0.6)
    assert (svd.nconv, svd.its) == (ref.nconv, ref.its) and svd.nconv >= 1
    assert f"{svd.triplet(0):.5f}" == "2.99254"                    # output/test4_1.out
    assert svd.error(0) < 5e-8
    v, u = Am.create_vecs()                                        # v: column space (22), u: row space (20)
    sigma = c_dbl()
    S.SVDGetSingularTriplet(svd.h, 0, ctypes.byref(sigma), u.h, v.h)
    uu, vv = u.get_values(), v.get_values()
    assert len(uu) == m and len(vv) == n
    assert np.linalg.norm(A @ vv - sigma.value * uu) < 1e-8 and np.linalg.norm(A.T @ uu - sigma.value * vv) < 1e-8
    for o in (svd, u, v, Am, ATm):
        o.destroy()


@pytest.mark.parametrize("locking", [0, 1])
def test_eps_test2_multiple_solves_same_object(locking):
    """eps/tests/test2.c (suffix 1_ks: -eps_type krylovschur -eps_krylovschur_locking {{0 1}}): several EPSSolve calls on ONE solver
    object with the same matrix, changing the wanted part of the spectrum in between; goldens output/test2_1.out rows 1 and 2
    (the interior part uses harmonic extraction, outside this path)"""
    Am = CP.mat_csr(O.laplacian_1d(30))
    eps = SL.EPS(Am, hermitian=True)
    CP.use_cpu_bv(eps)
    S.EPSSetDimensions(eps.h, 4, SL.PETSC_DETERMINE, SL.PETSC_DETERMINE)
    S.EPSKrylovSchurSetLocking(eps.h, locking)
    S.EPSSetWhichEigenpairs(eps.h, SL.EPS_LARGEST_REAL)
    eps.solve()
    assert [f"{eps.eigenvalue(i)[0]:.5f}" for i in range(4)] == ["3.98974", "3.95906", "3.90828", "3.83792"]
    S.EPSSetWhichEigenpairs(eps.h, SL.EPS_SMALLEST_REAL)
    eps.solve()
    assert [f"{eps.eigenvalue(i)[0]:.5f}" for i in range(4)] == ["0.01026", "0.04094", "0.09172", "0.16208"]
    assert max(eps.error(i) for i in range(4)) < 5e-8
    S.EPSSetWhichEigenpairs(eps.h, SL.EPS_LARGEST_REAL)            # and back: nothing of the previous solve may leak
    eps.solve()
    assert [f"{eps.eigenvalue(i)[0]:.5f}" for i in range(4)] == ["3.98974", "3.95906", "3.90828", "3.83792"]
    for o in (eps, Am):
        o.destroy()


def test_eps_test3_multiple_solves_different_matrices():
    """eps/tests/test3.c: EPSSetOperators with a second matrix on the SAME solver object, then EPSSolve again (tridiagonal -1 / random
    diagonal / -1, n = 30, nev = 4).  The reference's printed values depend on PetscRandom, so the pin is numpy's eigvalsh of the
    two matrices (1e-10 relative) and the -terse residual criterion"""
    import scipy.sparse as sp
    n = 30
    rng = np.random.default_rng(3)
    mats = [sp.diags([np.full(n - 1, -1.0), rng.uniform(0, 1, n), np.full(n - 1, -1.0)], [-1, 0, 1], format="csr") for _ in range(2)]
    Ms = [CP.mat_csr(A) for A in mats]
    eps = SL.EPS(Ms[0], hermitian=True)
    CP.use_cpu_bv(eps)
    S.EPSSetDimensions(eps.h, 4, SL.PETSC_DETERMINE, SL.PETSC_DETERMINE)
    S.EPSSetTolerances(eps.h, 1e-10, SL.PETSC_CURRENT)
    for A, M in zip(mats, Ms):
        S.EPSSetOperators(eps.h, M.h, None)
        eps.solve()
        assert eps.reason > 0 and eps.nconv >= 4
        w = np.linalg.eigvalsh(A.toarray())
        ref = w[np.argsort(-np.abs(w))][:4]                        # default: largest magnitude
        lam = np.array([eps.eigenvalue(i)[0] for i in range(4)])
        assert np.allclose(lam, ref, rtol=1e-10, atol=0), (lam, ref)
        assert max(eps.error(i) for i in range(4)) < 5e-10
    for o in [eps] + Ms:
        o.destroy()


def _svd_values(svd, nsv, ncv=None):
    S.SVDSetDimensions(svd.h, nsv, ncv or SL.PETSC_DETERMINE, SL.PETSC_DETERMINE)
    svd.solve()
    assert svd.nconv >= nsv and max(svd.error(i) for i in range(nsv)) < 5e-6
    return [f"{svd.triplet(i):.5f}" for i in range(nsv)]


def test_svd_test8_resolve_with_larger_subspace():
    """svd/tests/test8.c: SVDSolve, then SVDSetDimensions(nsv, ncv+2) and SVDSolve again on the same object (output/test8_1.out)"""
    A = O.grcar_rect(30, 30)
    Am, ATm = CP.mat_csr(A), CP.mat_csr(A.T.tocsr())
    svd = SL.SVD(Am, ATm)
    CP.use_cpu_bv(svd)
    S.SVDSetTolerances(svd.h, 1e-6, 1000)
    assert _svd_values(svd, 3, 12) == ["3.22149", "3.21754", "3.16696"]
    assert _svd_values(svd, 3, 14) == ["3.22149", "3.21754", "3.16696"]
    for o in (svd, Am, ATm):
        o.destroy()


def test_svd_test9_test14_new_operator_on_the_same_object():
    """svd/tests/test9.c (a matrix of ANOTHER size: SVDSetOperators resets the bases, svdsetup.c:107-116) and test14.c (same size),
    goldens output/test9_1.out and test14_1.out; then a tall matrix after a wide one (the M < N swap state must not leak)"""
    import scipy.sparse as sp
    mats = {}
    for key, A in (("g30", O.grcar_rect(30, 30)), ("g60", O.grcar_rect(60, 60)), ("tall", O.grcar_rect(35, 30))):
        mats[key] = (A, CP.mat_csr(A), CP.mat_csr(A.T.tocsr()))
    m, n = 20, 22
    A1, B1 = sp.lil_matrix((m, n)), sp.lil_matrix((m, n))
    for i in range(m):
        A1[i, i], A1[i, i + 1] = 1.0, 2.0
        if i == 0:
            B1[i, i] = 1.0
        else:
            B1[i, i - 1], B1[i, i] = 2.0, 1.0
    for key, A in (("bidA", A1.tocsr()), ("bidB", B1.tocsr())):
        mats[key] = (A, CP.mat_csr(A), CP.mat_csr(A.T.tocsr()))
    svd = SL.SVD(mats["g30"][1], mats["g30"][2])
    CP.use_cpu_bv(svd)
    S.SVDSetTolerances(svd.h, 1e-6, 1000)

    def switch(key):
        S.SVDSetOperators(svd.h, mats[key][1].h, None)
        S.SVDSetTransposeMatrix(svd.h, mats[key][2].h)
    assert _svd_values(svd, 3) == ["3.22149", "3.21754", "3.16696"]
    switch("g60")
    assert _svd_values(svd, 3) == ["3.23585", "3.23548", "3.21932"]            # test9_1.out
    switch("bidA")
    assert _svd_values(svd, 3) == ["2.99254", "2.97023", "2.93324"]            # test14_1.out
    switch("bidB")
    assert _svd_values(svd, 3) == ["2.99205", "2.96825", "2.92879"]
    switch("tall")
    ref = np.linalg.svd(mats["tall"][0].toarray(), compute_uv=False)[:3]
    assert _svd_values(svd, 3) == [f"{x:.5f}" for x in ref]
    svd.destroy()
    for _, a, b in mats.values():
        a.destroy(); b.destroy()


def test_eps_new_operator_of_another_size_on_the_same_object():
    """EPSSetOperators with a matrix of another dimension (EPSReset, epssetup.c:441-458): the basis is rebuilt at the next set-up"""
    M1, M2 = CP.mat_csr(O.laplacian_1d(30)), CP.mat_csr(O.laplacian_2d(12))
    eps = SL.EPS(M1, hermitian=True)
    CP.use_cpu_bv(eps)
    S.EPSSetDimensions(eps.h, 4, SL.PETSC_DETERMINE, SL.PETSC_DETERMINE)
    eps.solve()
    assert [f"{eps.eigenvalue(i)[0]:.5f}" for i in range(4)] == ["3.98974", "3.95906", "3.90828", "3.83792"]
    S.EPSSetOperators(eps.h, M2.h, None)
    eps.solve()
    ref = O.eps_krylovschur(O.laplacian_2d(12), 144, nev=4)
    assert (eps.nconv, eps.its) == (ref.nconv, ref.its)
    assert np.allclose([eps.eigenvalue(i)[0] for i in range(4)], ref.eigr[:4], rtol=1e-12, atol=0)
    S.EPSSetOperators(eps.h, M1.h, None)
    eps.solve()
    assert [f"{eps.eigenvalue(i)[0]:.5f}" for i in range(4)] == ["3.98974", "3.95906", "3.90828", "3.83792"]
    for o in (eps, M1, M2):
        o.destroy()


def test_svd_ex15_lauchli_breakdown_in_the_first_cycle():
    """svd/tutorials/ex15.c (output/ex15_1.out: 101 x 100 Lauchli matrix, mu = 1e-7, nsv = 1): the matrix has two distinct singular
    values, so the bidiagonalisation breaks down inside the first cycle and TRLanczos accepts everything it has: `Number of
    iterations of the method: 1`, `Number of converged approximate singular triplets: 10`, sigma_1 = 10.000000, the rest = mu"""
    import scipy.sparse as sp
    n, mu = 100, 1e-7
    A = sp.lil_matrix((n + 1, n))
    A[0, :] = 1.0
    for i in range(1, n + 1):
        A[i, i - 1] = mu                                          # ex15.c:43-48
    A = A.tocsr()
    Am, ATm = CP.mat_csr(A), CP.mat_csr(A.T.tocsr())
    svd = SL.SVD(Am, ATm)
    CP.use_cpu_bv(svd)
    S.SVDSetDimensions(svd.h, 1, SL.PETSC_DETERMINE, SL.PETSC_DETERMINE)
    svd.solve()
    ref = O.svd_trlanczos(A, A.T.tocsr(), n + 1, n, nsv=1)
    assert (svd.nconv, svd.its) == (10, 1) == (ref.nconv, ref.its)
    assert f"{svd.triplet(0):.6f}" == "10.000000"
    assert all(f"{svd.triplet(i):.6f}" == "0.000000" and abs(svd.triplet(i) - mu) < 1e-12 for i in range(1, 10))
    for o in (svd, Am, ATm):
        o.destroy()


def test_eps_ex19_3d_laplacian_smallest_with_multiplicities():
    """eps/tutorials/ex19.c (output/ex19_1.out, -eps_nev 8 -eps_type krylovschur -eps_ncv 64): the 7-point Laplacian on a 10^3 grid — the
    stencil of BASELINE configs[2] — smallest eigenvalues, the triple ones found with their multiplicity; the tutorial itself
    compares with the analytic spectrum (ex19.c:25-45), so does this test"""
    n = 10
    A = O.laplacian_3d(n).tocsr()
    Am = CP.mat_csr(A)
    eps = SL.EPS(Am, hermitian=True)
    CP.use_cpu_bv(eps)
    S.EPSSetWhichEigenpairs(eps.h, SL.EPS_SMALLEST_REAL)
    S.EPSSetDimensions(eps.h, 8, 64, SL.PETSC_DETERMINE)
    eps.solve()
    assert eps.reason > 0 and eps.nconv >= 8
    lam = [eps.eigenvalue(i)[0] for i in range(8)]
    assert [f"{x:.5f}" for x in lam] == ["0.24304", "0.47952", "0.47952", "0.47952", "0.71600", "0.71600", "0.71600", "0.85231"]
    th = 2 - 2 * np.cos(np.arange(1, n + 1) * np.pi / (n + 1))
    exact = np.sort((th[:, None, None] + th[None, :, None] + th[None, None, :]).ravel())[:8]
    assert np.allclose(lam, exact, rtol=1e-10, atol=0)
    ref = O.eps_krylovschur(A, n ** 3, nev=8, ncv=64, which="smallest_real")
    assert (eps.nconv, eps.its) == (ref.nconv, ref.its)
    assert np.allclose(lam, ref.eigr[ref.perm][:8], rtol=1e-12, atol=0)
    for o in (eps, Am):
        o.destroy()


def test_eps_ex9_brusselator_complex_pairs():
    """eps/tutorials/ex9.c (output/ex9_1.out, -n 50 -eps_nev 4, krylovschur, EPS_LARGEST_REAL): the Brusselator wave model, a real
    non-symmetric 100 x 100 operator [tau1 T + (beta-1) I, alpha^2 I; -beta I, tau2 T - alpha^2 I] (ex9.c:191-225, assembled here instead
    of the tutorial's MatShell) whose rightmost eigenvalues are complex conjugate pairs: 0.00007+-2.13946i, -0.67386+-2.52812i"""
    import scipy.sparse as sp
    N, alpha, beta, d1, d2, L = 50, 2.0, 5.45, 0.008, 0.004, 0.51302
    h = 1.0 / (N + 1)
    tau1, tau2 = d1 / (h * L) ** 2, d2 / (h * L) ** 2
    T = sp.diags([np.ones(N - 1), -2 * np.ones(N), np.ones(N - 1)], [-1, 0, 1])
    I = sp.identity(N)
    A = sp.bmat([[tau1 * T + (beta - 1) * I, alpha ** 2 * I], [-beta * I, tau2 * T - alpha ** 2 * I]]).tocsr()
    Am = CP.mat_csr(A)
    eps = SL.EPS(Am, hermitian=False)
    CP.use_cpu_bv(eps)
    S.EPSSetWhichEigenpairs(eps.h, SL.EPS_LARGEST_REAL)
    S.EPSSetDimensions(eps.h, 4, SL.PETSC_DETERMINE, SL.PETSC_DETERMINE)
    eps.solve()
    assert eps.reason > 0 and eps.nconv >= 4
    got = ["%.5f%+.5fi" % eps.eigenvalue(i) for i in range(4)]
    assert got == ["0.00007+2.13946i", "0.00007-2.13946i", "-0.67386+2.52812i", "-0.67386-2.52812i"]
    assert max(eps.error(i) for i in range(4)) < 5e-8
    ref = O.eps_krylovschur(A, 2 * N, nev=4, which="largest_real", hermitian=False)
    assert (eps.nconv, eps.its) == (ref.nconv, ref.its)
    for o in (eps, Am):
        o.destroy()
