"""The reference's BV known-answer programs (src/sys/classes/bv/tests/test1.c, test2.c, test4.c,
test13.c) restated against the C front-end (include/b2kslepc.h).  The reference validates a new BV type
by adding its name to the `-bv_type {{…}}` loops of exactly these programs; here each
scenario takes a `make_bv(n, m)` factory so the same code runs with the oracle's CPU type
(tests/test_host_cpu.py) and with the product's "b200" type on the GPU (tests/test_slepc_gpu.py).
Golden numbers: bv/tests/output/test1_1_bv_type-svec.out, test2_1.out, test4_1.out, test13_1.out.
"""
import ctypes

import numpy as np
import pytest

from oracle import slepc_oracle as O
from slepc_b200 import slepc as SL
from slepc_b200.slepc import S, c_dbl, c_int

EPS = np.finfo(float).eps


def g6(x):
    return float(f"{x:.6g}")


def fill_test1(X, ncols, n):
    for j in range(ncols):
        c = np.zeros(n)
        for i in range(4):
            if i + j < n:
                c[i + j] = 3 * i + j - 2
        X.set_column(j, c)


def qmat(k, l):
    return SL.Mat.seqdense(np.array([[2.0 if i < j else -0.5 for j in range(l)] for i in range(k)]))


def norm_column(X, j, t=SL.NORM_2):
    v = c_dbl()
    S.BVNormColumn(X.h, j, t, ctypes.byref(v))
    return v.value


def norm(X, t=SL.NORM_FROBENIUS):
    v = c_dbl()
    S.BVNorm(X.h, t, ctypes.byref(v))
    return v.value


def with_column(X, j, fn):
    v = ctypes.c_void_p()
    S.BVGetColumn(X.h, j, ctypes.byref(v))
    try:
        return fn(v)
    finally:
        S.BVRestoreColumn(X.h, j, ctypes.byref(v))


def scenario_test1(make_bv):
    n, k, l = 10, 5, 3
    X, Y = make_bv(n, k), make_bv(n, l)
    fill_test1(X, k, n)
    for j in range(l):
        Y.set_column(j, np.full(n, (j + 1) / 4.0))
    Q = qmat(k, l)
    S.BVMult(Y.h, 2.0, 1.0, X.h, Q.h)
    assert np.allclose(Y.get_column(0), [2.25, 0.25, -5.75, -15.75, -19.75, -20.75, -17.75, -10.75, 0.25, 0.25], atol=1e-13)
    assert np.allclose(Y.get_column(2), [-7.25, 0.75, 24.75, 44.75, 20.75, -20.25, -17.25, -10.25, 0.75, 0.75], atol=1e-13)
    z = np.array([2.0 * (-0.5) ** i for i in range(k)])
    with_column(Y, 0, lambda v: S.BVMultVec(X.h, -1.0, 1.0, v, z.ctypes.data_as(ctypes.c_void_p)))
    assert np.allclose(Y.get_column(0), [6.25, -2.75, -11.75, -26.0, -14.0, -24.125, -16.25, -12.125, 0.25, 0.25], atol=1e-13)
    M = SL.Mat.seqdense(np.zeros((l, k)))
    S.BVDot(X.h, Y.h, M.h)
    gold_M = np.array([[-244.25, -262.75, -379.125, -413.375, -412.0], [215.0, -35.0, -243.0, -377.0, -397.0],
                       [427.5, 438.5, 76.5, -186.5, -310.5]])
    assert np.allclose(M.dense_array(), gold_M, atol=1e-11)
    zz = np.zeros(k)
    with_column(Y, 0, lambda v: S.BVDotVec(X.h, v, zz.ctypes.data_as(ctypes.c_void_p)))
    assert np.allclose(zz, [-244.25, -262.75, -379.125, -413.375, -412.0], atol=1e-11)
    S.BVMultInPlace(X.h, Q.h, 1, l)
    S.BVScale(X.h, 2.0)
    assert np.allclose(X.get_column(1), [-8, 5, 14, 19, -20, -21, -18, -11, 0, 0], atol=1e-12)
    assert np.allclose(X.get_column(2), [-8, 0, 24, 44, 20, -21, -18, -11, 0, 0], atol=1e-12)
    assert g6(norm_column(X, 0)) == 16.7332
    assert g6(norm(X)) == 87.1436
    assert list(X.to_numpy()[0, :k]) == [-4.0, -8.0, -8.0, 0.0, 0.0]
    for o in (X, Y, Q, M):
        o.destroy()


def scenario_test2(make_bv):
    n, k = 20, 8
    X = make_bv(n, k)
    X0 = np.zeros((n, k), order="F")
    for j in range(k):
        for i in range(n // 2 + 1):
            if i + j < n:
                X0[i + j, j] = (3.0 * i + j - 2) / (2 * (i + j + 1))
    X.from_numpy(X0)
    Y = SL.BV()
    S.BVDuplicate(X.h, Y.ref)
    S.BVCopy(X.h, Y.h)
    Z = SL.BV()
    S.BVDuplicate(X.h, Z.ref)
    S.BVCopy(X.h, Z.h)
    nrm = c_dbl()
    for j in range(k):
        S.BVOrthogonalizeColumn(X.h, j, None, ctypes.byref(nrm), None)
        S.BVScaleColumn(X.h, j, 1.0 / nrm.value)
    M = SL.Mat.seqdense(np.zeros((k, k)))
    S.BVDot(X.h, X.h, M.h)
    assert np.linalg.norm(M.dense_array() - np.eye(k), 1) < 100 * EPS
    R = SL.Mat.seqdense(np.zeros((k, k)))
    S.BVOrthogonalize(Y.h, R.h)
    S.BVDot(Y.h, Y.h, M.h)
    assert np.linalg.norm(M.dense_array() - np.eye(k), 1) < 100 * EPS
    S.BVMult(Z.h, -1.0, 1.0, Y.h, R.h)
    assert norm(Z) < 100 * EPS
    Rm = R.dense_array()
    assert np.allclose(np.tril(Rm, -1), 0.0)
    e = SL.Vec()
    S.BVCreateVec(X.h, e.ref)
    S.VecSet(e.h, 1.0)
    S.BVOrthogonalizeVec(X.h, e.h, None, ctypes.byref(nrm), None)
    assert g6(nrm.value) == 2.50931
    for o in (X, Y, Z, M, R, e):
        o.destroy()


def scenario_test4(make_bv, trans=False):
    n, kx, lx, ky, ly = 18, 12, 3, 8, 2
    X = make_bv(n, kx + 2)
    X.set_active(lx, kx)
    fill_test1(X, kx + 2, n)
    Y = make_bv(n, ky + 1)
    Y.set_active(ly, ky)
    for j in range(ky + 1):
        Y.set_column(j, np.full(n, (j + 1) / 4.0))
    Qn = np.array([[2.0 if i < j else -0.5 for j in range(ky)] for i in range(kx)])
    Q = SL.Mat.seqdense(Qn)
    S.BVResize(X.h, kx + 4, 1)
    S.BVMult(Y.h, 2.0, 0.5, X.h, Q.h)
    z = np.array([2.0 * (-0.5) ** i for i in range(kx - lx)])
    with_column(Y, 0, lambda v: S.BVMultVec(X.h, -1.0, 1.0, v, z.ctypes.data_as(ctypes.c_void_p)))
    M = SL.Mat.seqdense(np.zeros((ky, kx)))
    S.BVDot(X.h, Y.h, M.h)
    zz = np.zeros(kx - lx)
    with_column(Y, 0, lambda v: S.BVDotVec(X.h, v, zz.ctypes.data_as(ctypes.c_void_p)))
    # cross-check the windowed products with the numpy oracle on the same data
    Xo, Yo = O.BV(n, kx + 4), O.BV(n, ky + 1)
    Xo.set_active(lx, kx)
    Yo.set_active(ly, ky)
    for j in range(kx + 2):
        for i in range(4):
            if i + j < n:
                Xo.col(j)[i + j] = 3 * i + j - 2
    for j in range(ky + 1):
        Yo.col(j)[:] = (j + 1) / 4.0
    Yo.mult(2.0, 0.5, Xo, Qn)
    Xo.multvec(-1.0, 1.0, Yo.col(0), z)
    assert np.allclose(Y.to_numpy(), Yo.V, atol=1e-12)
    assert np.allclose(M.dense_array()[ly:ky, lx:kx], Xo.dot(Yo), atol=1e-10)
    assert np.allclose(zz, Xo.dotvec(Yo.col(0)), atol=1e-10)
    if trans:
        Qt = SL.Mat.seqdense(Qn.T.copy())
        S.BVMultInPlaceHermitianTranspose(X.h, Qt.h, lx + 1, ky)
        Qt.destroy()
    else:
        S.BVMultInPlace(X.h, Q.h, lx + 1, ky)
    S.BVScale(X.h, 2.0)
    assert g6(norm_column(X, lx)) == 25.7682
    assert g6(norm(X)) == 328.469
    for o in (X, Y, Q, M):
        o.destroy()


def scenario_test11(make_bv, block, n=20, l=2, k=8, resid=True):
    """bv/tests/test11.c:37-215 (output/test11_1.out, the `-resid` variants of :255): block orthogonalisation of the active
    columns after orthogonalising the leading ones.  At the reference's sizes (n=20, l=2, k=8 and n=180, l=0, k=7 with GS)
    every check is the reference's `< 100*eps`; other shapes scale the bound with the size and the conditioning of the
    test matrix (Gram-matrix methods lose orthogonality like eps*cond^2).  Also compared with the numpy restatement of
    the same method (bvorthog.c:586-675, bvlapack.c:136-345)."""
    X = make_bv(n, k)
    X0 = np.zeros((n, k))
    for j in range(k):
        for i in range(n // 2 + 1):
            if i + j < n:
                X0[i + j, j] = (3.0 * i + j - 2) / (2 * (i + j + 1))
    gram = block in (SL.BV_ORTHOG_BLOCK_CHOL, SL.BV_ORTHOG_BLOCK_SVQB)        # TSQR is Householder: orthogonal whatever the conditioning
    if block == SL.BV_ORTHOG_BLOCK_TSQRCHOL and (n, l, k) != (20, 2, 8):
        gram = None                                                           # Q = V inv(R) with a backward-stable R: eps*cond
    if (n, l, k) == (20, 2, 8) or ((n, l, k) == (180, 0, 7) and block == SL.BV_ORTHOG_BLOCK_GS):
        tol_orth = tol_res = 100 * EPS
    else:
        c = np.linalg.cond(X0)
        tol_orth = 100 * EPS * k * (c * c if gram else c)
        if block == SL.BV_ORTHOG_BLOCK_TSQR:
            tol_orth = 100 * EPS * k
        tol_res = 100 * EPS * np.linalg.norm(X0)
    X.from_numpy(X0)
    Y = SL.BV()
    S.BVDuplicate(X.h, Y.ref)
    S.BVCopy(X.h, Y.h)
    S.BVSetOrthogonalization(Y.h, SL.BV_ORTHOG_CGS, SL.BV_ORTHOG_REFINE_IFNEEDED, 0.7071, block)
    M = SL.Mat.seqdense(np.zeros((k, k)))
    R = SL.Mat.seqdense(np.zeros((k, k))) if resid else None
    Rh = R.h if resid else None
    name = {SL.BV_ORTHOG_BLOCK_GS: None, SL.BV_ORTHOG_BLOCK_CHOL: "chol", SL.BV_ORTHOG_BLOCK_SVQB: "svqb",
            SL.BV_ORTHOG_BLOCK_TSQR: "tsqr", SL.BV_ORTHOG_BLOCK_TSQRCHOL: "tsqrchol"}[block]
    Yo = O.BV(n, k)
    Yo.V[:, :] = X0

    def offdiag(lo, hi):
        S.BVDot(Y.h, Y.h, M.h)
        A = M.dense_array()[lo:hi, lo:hi]
        return np.linalg.norm(A - np.eye(hi - lo))

    if l > 0:
        Y.set_active(0, l)
        S.BVOrthogonalize(Y.h, Rh)
        assert offdiag(0, l) < tol_orth                       # Level of orthogonality of Q1
        if resid:
            assert np.linalg.norm(X0[:, :l] - Y.to_numpy()[:, :l] @ R.dense_array()[:l, :l]) < tol_res
        if name:
            Yo.set_active(0, l)
            Yo.orthogonalize_block(name)
    Y.set_active(l, k)
    S.BVOrthogonalize(Y.h, Rh)
    if l > 0:
        assert offdiag(l, k) < tol_orth                       # Level of orthogonality of Q2
    Y.set_active(0, k)
    assert offdiag(0, k) < tol_orth                           # Level of orthogonality of Q
    Q = Y.to_numpy()
    if resid:
        Rm = R.dense_array()
        assert np.linalg.norm(X0 - Q @ Rm) < tol_res          # Residual ||X-Q*R||
        if block != SL.BV_ORTHOG_BLOCK_SVQB:
            assert np.allclose(np.tril(Rm, -1), 0.0)
    if name:
        Yo.set_active(l, k)
        Ro = Yo.orthogonalize_block(name, want_R=True)
        Qo = Yo.V[:, :k]
        if name in ("chol", "tsqr", "tsqrchol"):              # unique factorisation (R with a positive diagonal): column for column
            assert np.linalg.norm(Q - Qo) < 1e3 * tol_orth
            if resid:
                assert np.linalg.norm(R.dense_array()[:, l:k] - Ro[:, l:k]) < 1e3 * tol_orth * np.linalg.norm(Ro)
        else:
            # SVQB post-multiplies by eigenvectors of the scaled Gram matrix: each column is defined up to its sign and is
            # as sensitive as the eigenvector (eps / spectral gap), so compare the spans of the leading and of the active block
            for a, b in ((0, l), (l, k)):
                if b > a:
                    P, Po = Q[:, a:b] @ Q[:, a:b].T, Qo[:, a:b] @ Qo[:, a:b].T
                    assert np.linalg.norm(P @ X0 - Po @ X0) < 1e3 * tol_orth * np.linalg.norm(X0)
    for o in (X, Y, M) + ((R,) if resid else ()):
        o.destroy()


def scenario_norms(make_bv, n=1237, k=6, l=1):
    """BVNorm / BVNormColumn with every NormType of bvglobal.c:455-560 (2, Frobenius, 1, infinity) on the active block
    and on single columns, against numpy"""
    X = make_bv(n, k)
    A = np.random.default_rng(3).standard_normal((n, k))
    X.from_numpy(A)
    X.set_active(l, k)
    B = A[:, l:k]
    assert np.isclose(norm(X, SL.NORM_FROBENIUS), np.linalg.norm(B), rtol=1e-13)
    assert np.isclose(norm(X, SL.NORM_1), np.abs(B).sum(0).max(), rtol=1e-13)
    assert np.isclose(norm(X, SL.NORM_INFINITY), np.abs(B).sum(1).max(), rtol=1e-13)
    for j in (0, k - 1):
        assert np.isclose(norm_column(X, j, SL.NORM_2), np.linalg.norm(A[:, j]), rtol=1e-13)
        assert np.isclose(norm_column(X, j, SL.NORM_1), np.abs(A[:, j]).sum(), rtol=1e-13)
        assert norm_column(X, j, SL.NORM_INFINITY) == np.abs(A[:, j]).max()
    X.destroy()


def scenario_test13(make_bv):
    n, k = 10, 5
    X = make_bv(n, k)
    fill_test1(X, k, n)
    S.BVDotColumn(X.h, 2, None)          # q == NULL ⇒ the internal buffer (test13.c:60)
    S.BVMultColumn(X.h, -1.0, 1.0, 2, None)
    assert g6(norm(X)) == 711.996
    X.destroy()


def scenario_test10(make_bv, n=10, k=5):
    """bv/tests/test10.c:51-111 (output/test10_1.out: `0.`): the split-phase operations — VecDotBegin/End, BVDotVecBegin/End,
    BVDotColumnBegin/End, VecNormBegin/End, BVNormVecBegin/End, BVNormColumnBegin/End, posted together and collected in the
    reference's order — give exactly what the regular operations give (the test prints the 1-norm of the difference: 0)"""
    X = make_bv(n, k)
    fill_test1(X, k, n)                                   # the same fill as test1 (test10.c:44-54)
    v = ctypes.c_void_p()
    S.BVCreateVec(X.h, ctypes.byref(v))
    S.VecSet(v, 1.0)
    z, zs = np.zeros(k + 6), np.zeros(k + 6)

    def dptr(a, off):
        return ctypes.cast(a.ctypes.data + 8 * off, ctypes.POINTER(ctypes.c_double))
    nrm = c_dbl()
    # regular operations
    w = ctypes.c_void_p()
    S.BVGetColumn(X.h, 0, ctypes.byref(w)); S.VecDot(w, v, dptr(z, 0)); S.BVRestoreColumn(X.h, 0, ctypes.byref(w))
    S.BVDotVec(X.h, v, dptr(z, 1))
    S.BVDotColumn(X.h, 2, dptr(z, 1 + k))
    S.BVGetColumn(X.h, 1, ctypes.byref(w)); S.VecNorm(w, SL.NORM_2, ctypes.byref(nrm)); S.BVRestoreColumn(X.h, 1, ctypes.byref(w))
    z[k + 3] = nrm.value
    S.BVNormVec(X.h, v, SL.NORM_2, ctypes.byref(nrm)); z[k + 4] = nrm.value
    S.BVNormColumn(X.h, 0, SL.NORM_2, ctypes.byref(nrm)); z[k + 5] = nrm.value
    # split operations, posted and collected as test10.c:82-104 does
    S.BVGetColumn(X.h, 0, ctypes.byref(w))
    S.VecDotBegin(w, v, dptr(zs, 0))
    S.BVDotVecBegin(X.h, v, dptr(zs, 1))
    S.BVDotColumnBegin(X.h, 2, dptr(zs, 1 + k))
    S.VecDotEnd(w, v, dptr(zs, 0))
    S.BVRestoreColumn(X.h, 0, ctypes.byref(w))
    S.BVDotVecEnd(X.h, v, dptr(zs, 1))
    S.BVDotColumnEnd(X.h, 2, dptr(zs, 1 + k))
    y = ctypes.c_void_p()
    S.BVGetColumn(X.h, 1, ctypes.byref(y))
    S.VecNormBegin(y, SL.NORM_2, ctypes.byref(nrm))
    S.BVNormVecBegin(X.h, v, SL.NORM_2, ctypes.byref(nrm))
    S.BVNormColumnBegin(X.h, 0, SL.NORM_2, ctypes.byref(nrm))
    S.VecNormEnd(y, SL.NORM_2, ctypes.byref(nrm)); zs[k + 3] = nrm.value
    S.BVRestoreColumn(X.h, 1, ctypes.byref(y))
    S.BVNormVecEnd(X.h, v, SL.NORM_2, ctypes.byref(nrm)); zs[k + 4] = nrm.value
    S.BVNormColumnEnd(X.h, 0, SL.NORM_2, ctypes.byref(nrm)); zs[k + 5] = nrm.value
    assert np.abs(z - zs).sum() == 0.0, (z, zs)           # the reference prints "0."
    # and the numbers themselves (numpy on the same fill)
    A = X.to_numpy()
    one = np.ones(n)
    ref = np.concatenate([[A[:, 0] @ one], A.T @ one, A[:, :2].T @ A[:, 2], [np.linalg.norm(A[:, 1]), np.sqrt(n), np.linalg.norm(A[:, 0])]])
    assert np.allclose(z, ref, rtol=1e-14, atol=1e-14), (z, ref)
    S.VecDestroy(ctypes.byref(v))
    X.destroy()


def scenario_test3(make_bv, make_mat, n=10, k=5):
    """bv/tests/test3.c (output/test3_1.out): BV with the non-standard inner product of the tridiagonal B — B-norm of the first
    column 8.94427, B-orthonormalisation column by column (level of B-orthogonality < 100 eps), then B-norm 1 through the
    split-phase pair BVNormVecBegin / End"""
    import scipy.sparse as sp
    B = make_mat(sp.diags([np.full(n - 1, -1.0), np.full(n, 2.0), np.full(n - 1, -1.0)], [-1, 0, 1], format="csr"))
    X = make_bv(n, k)
    S.BVSetMatrix(X.h, B.h, 0)
    fill_test1(X, k, n)                                   # test3.c:57-66, the fill of test1
    assert g6(norm_column(X, 0)) == 8.94427
    nrm = c_dbl()
    for j in range(k):
        S.BVOrthogonalizeColumn(X.h, j, None, ctypes.byref(nrm), None)
        S.BVScaleColumn(X.h, j, 1.0 / nrm.value)
    M = SL.Mat.seqdense(np.zeros((k, k)))
    S.BVDot(X.h, X.h, M.h)
    assert np.abs(M.dense_array() - np.eye(k)).sum(axis=0).max() < 100 * EPS
    v = ctypes.c_void_p()
    S.BVGetColumn(X.h, 0, ctypes.byref(v))
    S.BVNormVecBegin(X.h, v, SL.NORM_1, ctypes.byref(nrm))     # the type is ignored with a matrix (bvglobal.c:569-572)
    S.BVNormVecEnd(X.h, v, SL.NORM_1, ctypes.byref(nrm))
    S.BVRestoreColumn(X.h, 0, ctypes.byref(v))
    assert abs(nrm.value - 1.0) < 100 * EPS                    # the reference prints "1."
    for o in (X, M, B):
        o.destroy()


def scenario_test7(make_bv, make_mat, m=10, k=5):
    """bv/tests/test7.c (output/test7_1.out, `Norm of error: 0.` twice; the BVGetMat view in between is a viewer feature): Y(:,2:k+2) =
    B X with the 1-D Laplacian B and X(:,j) = ones in the first j+1 rows, against the known answer (e_0 + e_j - e_{j+1}); then the
    same product one column at a time (BVMatMultColumn on a shifted copy)"""
    import scipy.sparse as sp
    n = m
    Bs = sp.diags([np.full(n - 1, -1.0), np.full(n, 2.0), np.full(n - 1, -1.0)], [-1, 0, 1], format="csr")
    B = make_mat(Bs)
    X0 = np.zeros((n, k))
    for j in range(k):
        X0[:j + 1, j] = 1.0
    X, Y = make_bv(n, k), make_bv(m, k + 4)
    X.from_numpy(X0)
    Y.from_numpy(np.zeros((m, k + 4)))
    Y.set_active(2, k + 2)
    S.BVMatMult(X.h, B.h, Y.h)
    Z0 = np.zeros((m, k))
    for j in range(k):
        Z0[0, j] += 1.0
        if j < n:
            Z0[j, j] += 1.0
        if j + 1 < n:
            Z0[j + 1, j] += -1.0
    Yn = Y.to_numpy()
    assert np.abs(Yn[:, 2:k + 2] - Z0).max() == 0.0           # "Norm of error: 0."
    assert np.abs(Yn[:, :2]).max() == 0.0 and np.abs(Yn[:, k + 2:]).max() == 0.0      # the columns outside the active window are untouched
    assert np.array_equal(Z0, Bs @ X0)
    # column by column: W(:,j+1) = B W(:,j)
    W = make_bv(n, 3)
    W0 = np.zeros((n, 3)); W0[:, 0] = np.arange(1.0, n + 1)
    W.from_numpy(W0)
    S.BVMatMultColumn(W.h, B.h, 0)
    S.BVMatMultColumn(W.h, B.h, 1)
    Wn = W.to_numpy()
    assert np.array_equal(Wn[:, 1], Bs @ W0[:, 0]) and np.array_equal(Wn[:, 2], Bs @ (Bs @ W0[:, 0]))
    for o in (X, Y, W, B):
        o.destroy()


def scenario_test12(make_bv, n=20, k=8):
    """bv/tests/test12.c (output/test12_1.out, -bv_orthog_block gs): block orthogonalisation of a RANK-DEFICIENT basis — column k/2
    depends on columns 0 and 1, column k-1 on columns 1 and k/2+1.  BVOrthogonalize_GS (bvorthog.c:506-550) does not ask for
    the linear-dependence flag: the dependent column is reduced to rounding noise, re-orthogonalised by the DGKS passes and
    normalised, R(j,j) is tiny, and both checks hold at the reference's 100 eps"""
    def col(j):
        c = np.zeros(n)
        for i in range(n // 2 + 1):
            if i + j < n:
                c[i + j] = (3.0 * i + j - 2) / (2 * (i + j + 1))
        return c
    X0 = np.zeros((n, k))
    for j in range(k // 2):
        X0[:, j] = col(j)
    j = k // 2
    X0[:, j] = X0[:, 0] + 0.5 * X0[:, 1]
    for j in range(k // 2 + 1, k - 1):
        X0[:, j] = col(j)
    X0[:, k - 1] = X0[:, 1] - 1.2 * X0[:, k // 2 + 1]
    X = make_bv(n, k)
    X.from_numpy(X0)
    S.BVSetOrthogonalization(X.h, SL.BV_ORTHOG_CGS, SL.BV_ORTHOG_REFINE_IFNEEDED, 0.7071, SL.BV_ORTHOG_BLOCK_GS)
    R = SL.Mat.seqdense(np.zeros((k, k)))
    S.BVOrthogonalize(X.h, R.h)
    M = SL.Mat.seqdense(np.zeros((k, k)))
    S.BVDot(X.h, X.h, M.h)
    assert np.abs(M.dense_array() - np.eye(k)).sum(axis=0).max() < 100 * EPS          # MatNorm(M - I, NORM_1): level of orthogonality
    assert np.linalg.norm(X0 - X.to_numpy() @ R.dense_array()) < 100 * EPS             # residual ||X - QR||_F
    Rm = R.dense_array()
    assert abs(Rm[k // 2, k // 2]) < 1e-13 and abs(Rm[k - 1, k - 1]) < 1e-13          # the two dependent columns
    for o in (X, R, M):
        o.destroy()


def scenario_test18(make_bv, make_mat, n=250, l=6, k=15):
    """bv/tests/test18.c (output/test18_1.out, args -n 250 -l 6 -k 15): BVNormalize of the active columns — plain 2-norms, B-norms
    with the tridiagonal inner-product matrix of :98-109, and conjugate pairs (eigi != 0: the two columns of a pair are scaled
    by the norm of the complex vector, :128-166); each deviation from 1 below the reference's 100 eps"""
    import scipy.sparse as sp
    X0 = np.zeros((n, k))
    for j in range(k):
        for i in range(n // 2 + 1):
            if i + j < n:
                X0[i + j, j] = (3.0 * i + j - 2) / (2 * (i + j + 1))
    X, Y, Z = make_bv(n, k), make_bv(n, k), make_bv(n, k)
    for W in (X, Y, Z):
        W.from_numpy(X0)
        W.set_active(l, k)
    S.BVNormalize(X.h, None)
    assert max(abs(norm_column(X, j) - 1.0) for j in range(l, k)) < 100 * EPS
    assert np.array_equal(X.to_numpy()[:, :l], X0[:, :l])                   # the leading columns are not touched
    B = make_mat(sp.diags([np.full(n - 1, -1.0), np.full(n, 2.0), np.full(n - 1, -1.0)], [-1, 0, 1], format="csr"))
    S.BVSetMatrix(Y.h, B.h, 0)
    S.BVNormalize(Y.h, None)
    assert max(abs(norm_column(Y, j) - 1.0) for j in range(l, k)) < 100 * EPS          # B-norms (BVNormColumn with a matrix)
    Yn = Y.to_numpy()
    Bd = sp.diags([np.full(n - 1, -1.0), np.full(n, 2.0), np.full(n - 1, -1.0)], [-1, 0, 1]).toarray()
    assert max(abs(np.sqrt(Yn[:, j] @ Bd @ Yn[:, j]) - 1.0) for j in range(l, k)) < 100 * EPS
    eigi = np.zeros(k)
    rng = np.random.default_rng(18)
    for j in range(l + 1, k - 1, 5):
        a = rng.uniform(0.1, 1.0)
        eigi[j], eigi[j + 1] = a, -a
    S.BVNormalize(Z.h, eigi.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    Zn = Z.to_numpy()
    j, err = l, 0.0
    while j < k:
        if eigi[j] != 0.0:
            nr = np.hypot(np.linalg.norm(Zn[:, j]), np.linalg.norm(Zn[:, j + 1]))
            nr0 = np.hypot(np.linalg.norm(X0[:, j]), np.linalg.norm(X0[:, j + 1]))          # both columns of a pair by the same factor
            assert np.allclose(Zn[:, j], X0[:, j] / nr0, rtol=1e-14, atol=0) and np.allclose(Zn[:, j + 1], X0[:, j + 1] / nr0, rtol=1e-14, atol=0)
            j += 1
        else:
            nr = np.linalg.norm(Zn[:, j])
        err = max(err, abs(nr - 1.0))
        j += 1
    assert err < 100 * EPS
    for o in (X, Y, Z, B):
        o.destroy()


def scenario_errors(make_bv):
    """error behaviour of the front-end (same checks and messages as bvops.c / bvbasic.c)"""
    X, Y = make_bv(10, 5), make_bv(10, 3)
    v1, v2, v3 = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    S.BVGetColumn(X.h, 0, ctypes.byref(v1))
    with pytest.raises(SL.SlepcError, match="already fetched"):
        S.BVGetColumn(X.h, 0, ctypes.byref(v2))
    S.BVGetColumn(X.h, 1, ctypes.byref(v2))
    with pytest.raises(SL.SlepcError, match="Too many requested columns"):
        S.BVGetColumn(X.h, 2, ctypes.byref(v3))
    with pytest.raises(SL.SlepcError, match="has not been fetched"):
        S.BVRestoreColumn(X.h, 3, ctypes.byref(v1))
    S.BVRestoreColumn(X.h, 0, ctypes.byref(v1))
    S.BVRestoreColumn(X.h, 1, ctypes.byref(v2))
    with pytest.raises(SL.SlepcError, match="only 5 are available"):
        S.BVGetColumn(X.h, 5, ctypes.byref(v1))
    Q = SL.Mat.seqdense(np.zeros((4, 3)))
    with pytest.raises(SL.SlepcError, match="Mat argument has 4 rows, should have at least 5"):
        S.BVMult(Y.h, 1.0, 0.0, X.h, Q.h)
    with pytest.raises(SL.SlepcError, match="X and Y arguments must be different"):
        S.BVMult(X.h, 1.0, 0.0, X.h, None)
    Q5 = SL.Mat.seqdense(np.eye(5))
    with pytest.raises(SL.SlepcError, match="Argument s has wrong value"):
        S.BVMultInPlace(X.h, Q5.h, 7, 8)
    with pytest.raises(SL.SlepcError, match="requested value of e is larger"):
        S.BVSetActiveColumns(X.h, 0, 5) or S.BVMultInPlace(X.h, SL.Mat.seqdense(np.eye(5)[:, :3]).h, 0, 4)
    with pytest.raises(SL.SlepcError, match="Requested norm not available"):
        S.BVNorm(X.h, SL.NORM_2, ctypes.byref(c_dbl()))
    with pytest.raises(SL.SlepcError, match="Illegal value of k"):
        S.BVSetActiveColumns(X.h, 0, 9)
    with pytest.raises(SL.SlepcError, match="Index j=7 but BV only has 5 columns"):
        S.BVOrthogonalizeColumn(X.h, 7, None, None, None)
    with pytest.raises(SL.SlepcError, match="Unable to find requested BV type"):
        b = SL.BV()
        S.BVCreate(b.ref)
        S.BVSetType(b.h, b"nonexistent")
    for o in (X, Y, Q, Q5):
        o.destroy()


def scenario_orthog_vs_oracle(make_bv, otype, refine, n, k, tol):
    """column-by-column orthonormalisation of an ill-conditioned block: coefficients, norms and the basis must
    match the numpy oracle (same algorithm, different reduction order ⇒ tolerance, not bit equality)."""
    rng = np.random.default_rng(11)
    A = rng.standard_normal((n, k))
    if refine != SL.BV_ORTHOG_REFINE_NEVER:                 # without refinement such columns are rounding noise
        A[:, 5] = A[:, 4] + 1e-6 * rng.standard_normal(n)   # forces DGKS refinement
        A[:, 9] = A[:, :3] @ np.array([1.0, -2.0, 0.5]) + 1e-9 * rng.standard_normal(n)
    X = make_bv(n, k)
    X.from_numpy(A)
    S.BVSetOrthogonalization(X.h, otype, refine, 0.7071, 0)
    Xo = O.BV(n, k)
    Xo.V[:] = A
    Xo.orthog_type, Xo.orthog_ref = otype, refine
    nrm, lin = c_dbl(), c_int()
    H = np.zeros(k + 1)
    scale = float(np.linalg.norm(A, axis=0).max())
    norms = []
    for j in range(k):
        S.BVOrthogonalizeColumn(X.h, j, H.ctypes.data_as(ctypes.c_void_p), ctypes.byref(nrm), ctypes.byref(lin))
        ho, no, lo = Xo.orthogonalize_column(j)
        assert bool(lin.value) == bool(lo)
        assert abs(nrm.value - no) <= 1e-7 * max(1.0, abs(no))
        # a coefficient against q_i carries the rounding of column i amplified by ||a||/||r_i|| (r_i = column i before
        # normalisation): near-dependent columns (5 and 9) are amplified noise directions in BOTH implementations
        # and every later basis vector inherits that perturbation
        atol = 1e-13 * scale * scale / min(norms) if norms else 0.0
        assert np.all(np.abs(H[:j] - ho) <= atol + 1e-12 * np.abs(ho)), (j, np.abs(H[:j] - ho), atol)
        norms.append(no)
        S.BVScaleColumn(X.h, j, 1.0 / nrm.value)
        Xo.scale_column(j, 1.0 / no)
    Q = X.to_numpy()
    assert np.linalg.norm(Q.T @ Q - np.eye(k), 1) < (1e-6 if refine == SL.BV_ORTHOG_REFINE_NEVER else 100 * k * EPS)
    # same subspace as the oracle's basis
    assert np.linalg.norm(Q - Xo.V @ (Xo.V.T @ Q)) < 1e-6
    X.destroy()
