"""The dense projected-problem solvers of the host layer (slepc_b200/host/ds.c: DS HEP dense and compact-arrow, NHEP, SVD over LAPACK — they
stay on the host by design) driven exactly as the reference's own DS test programs drive them, against the reference's
printed outputs (src/sys/classes/ds/tests/output/*.out, numbers copied below with their file).  CPU only: DS never touches the GPU."""
import ctypes

import numpy as np
import pytest

from slepc_b200 import slepc as SL
from slepc_b200.slepc import S

DS_STATE_RAW, DS_STATE_INTERMEDIATE = 0, 1


def _fn(name):
    return ctypes.cast(getattr(SL.load(), name), ctypes.c_void_p)


def _ds(kind, n, l=0, k=0, compact=False, extrarow=False):
    ds = ctypes.c_void_p()
    S.DSCreate(ctypes.byref(ds))
    S.DSSetType(ds, kind.encode())
    ld = n + 2                                              # the reference tests a leading dimension larger than n
    S.DSAllocate(ds, ld)
    S.DSSetDimensions(ds, n, l, k)
    if compact:
        S.DSSetCompact(ds, 1)
    S.DSSetExtraRow(ds, 1 if extrarow else 0)
    return ds, ld


def _array(ds, which, ld, ncols, real=False):
    p = ctypes.POINTER(ctypes.c_double)()
    (S.DSGetArrayReal if real else S.DSGetArray)(ds, which, ctypes.byref(p))
    a = np.ctypeslib.as_array(p, shape=(ncols, ld)).T       # column-major ld x ncols view
    return a, p


def _release(ds, which, p, real=False):
    (S.DSRestoreArrayReal if real else S.DSRestoreArray)(ds, which, ctypes.byref(p))


def _solve_sorted(ds, n, cmp="SlepcCompareLargestMagnitude", complex_pairs=False):
    S.DSSetEigenvalueComparison(ds, _fn(cmp), None)
    wr, wi = np.zeros(n), np.zeros(n)
    pr = wr.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    pi = wi.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if complex_pairs else None
    S.DSSolve(ds, pr, pi)
    S.DSSort(ds, pr, pi, None, None, None)
    return wr, wi


def test_ds_hep_dense_storage_is_refused():
    """DS HEP exists in compact storage only here (what Krylov-Schur uses, krylovschur.c:160-167; ds/tests/test2.c is the dense variant)"""
    ds, ld = _ds("hep", 6)
    S.DSSetState(ds, DS_STATE_RAW)
    with pytest.raises(SL.SlepcError, match="compact storage only"):
        _solve_sorted(ds, 6)
    S.DSDestroy(ctypes.byref(ds))


@pytest.mark.parametrize("extrarow", [False, True])
def test_ds_test3_hep_compact_arrow(extrarow):
    """ds/tests/test3.c (-n 9, l = 2, k = 5; output/test3_1.out / test3_2.out): compact storage, the arrow-plus-tridiagonal shape the
    Krylov-Schur restart produces (dshep.c:221-262 DSArrowTridiag): two locked leading values stay, the trailing block is solved and sorted"""
    n, l, k = 9, 2, 5
    ds, ld = _ds("hep", n, l, k, compact=True, extrarow=extrarow)
    T, pT = _array(ds, SL.DS_MAT_T, ld, 3, real=True)
    for i in range(n):
        T[i, 0] = i + 1.0
    for i in range(l, n - 1):
        T[i, 1] = 1.0
    if extrarow:
        T[n - 1, 1] = 1.0
    _release(ds, SL.DS_MAT_T, pT, real=True)
    S.DSSetState(ds, DS_STATE_RAW)
    wr, _ = _solve_sorted(ds, n)
    if extrarow:
        S.DSUpdateExtraRow(ds)
    gold = [1.00000, 2.00000, 9.74775, 8.26575, 7.21468, 5.92893, 4.66166, 3.64745, 2.53379]
    assert [f"{x:.5f}" for x in wr] == [f"{x:.5f}" for x in gold]
    # independent check of the same matrix with numpy
    M = np.diag(np.arange(1.0, n + 1))
    for i in range(l, k):
        M[i, k] = M[k, i] = 1.0                            # the arrow: rows l..k-1 couple with row k
    for i in range(k, n - 1):
        M[i, i + 1] = M[i + 1, i] = 1.0
    w = np.linalg.eigvalsh(M[l:, l:])
    assert np.allclose(np.sort(wr[l:]), w, rtol=1e-13, atol=1e-13)
    if extrarow:
        T, pT = _array(ds, SL.DS_MAT_T, ld, 3, real=True)
        Q, pQ = _array(ds, SL.DS_MAT_Q, ld, ld)
        assert abs(sum(T[i, 1] - Q[n - 1, i] for i in range(l, n))) <= 10 * np.finfo(float).eps                 # test3.c:80-84
        _release(ds, SL.DS_MAT_T, pT, real=True); _release(ds, SL.DS_MAT_Q, pQ)
    S.DSDestroy(ctypes.byref(ds))


@pytest.mark.parametrize("extrarow", [False, True])
def test_ds_test1_nhep_grcar(extrarow):
    """ds/tests/test1.c (-n 10 default; output/test1_1.out / test1_2.out): Grcar matrix in upper-Hessenberg form (state INTERMEDIATE),
    DSSolve (hseqr) + DSSort by largest magnitude: the five conjugate pairs in the reference's order, `Value of rnorm for 3rd vector =
    0.122`, `Norm of 1st vector = 1.000`; with the extra row DSUpdateExtraRow leaves row n = -e_{n-1}^T Q"""
    n = 10
    ds, ld = _ds("nhep", n, extrarow=extrarow)
    A, pA = _array(ds, SL.DS_MAT_A, ld, ld)
    for i in range(1, n):
        A[i, i - 1] = -1.0
    for j in range(4):
        for i in range(n - j):
            A[i, i + j] = 1.0
    if extrarow:
        A[n, n - 1] = -1.0
    _release(ds, SL.DS_MAT_A, pA)
    S.DSSetState(ds, DS_STATE_INTERMEDIATE)
    wr, wi = _solve_sorted(ds, n, complex_pairs=True)
    if extrarow:
        S.DSUpdateExtraRow(ds)
    st = ctypes.c_int()
    S.DSGetState(ds, ctypes.byref(st))
    assert st.value == 2                                    # CONDENSED
    gold = [(0.19797, 2.12926), (1.58254, 1.01946), (0.56478, 1.75988), (1.12811, 1.27812), (1.52659, 0.40240)]
    got = [(f"{wr[2 * i]:.5f}", f"{abs(wi[2 * i]):.5f}") for i in range(5)]
    assert got == [(f"{a:.5f}", f"{b:.5f}") for a, b in gold]
    for i in range(5):                                      # conjugate pairs, positive imaginary part first (as LAPACK returns them)
        assert wr[2 * i] == wr[2 * i + 1] and wi[2 * i] == -wi[2 * i + 1] and wi[2 * i] > 0
    if extrarow:
        A, pA = _array(ds, SL.DS_MAT_A, ld, ld)
        Q, pQ = _array(ds, SL.DS_MAT_Q, ld, ld)
        assert abs(sum(A[n, i] + Q[n - 1, i] for i in range(n))) <= 10 * np.finfo(float).eps        # test1.c:96-100
        _release(ds, SL.DS_MAT_A, pA); _release(ds, SL.DS_MAT_Q, pQ)
    j, rnorm = ctypes.c_int(2), ctypes.c_double()
    S.DSVectors(ds, SL.DS_MAT_X, ctypes.byref(j), ctypes.byref(rnorm))
    if not extrarow:
        assert f"{rnorm.value:.3f}" == "0.122"
    S.DSVectors(ds, SL.DS_MAT_X, None, None)
    X, pX = _array(ds, SL.DS_MAT_X, ld, ld)
    assert f"{np.sqrt((X[:n, 0] ** 2 + X[:n, 1] ** 2).sum()):.3f}" == "1.000"       # the first vector is complex: columns 0 and 1
    _release(ds, SL.DS_MAT_X, pX)
    S.DSDestroy(ctypes.byref(ds))


@pytest.mark.parametrize("extrarow", [False, True])
def test_ds_test8_svd_compact_arrow(extrarow):
    """ds/tests/test8.c (n = m = 10, l = 2, k = 5; output/test8_1.out): compact upper arrow-bidiagonal matrix, the shape the thick-restart
    Lanczos bidiagonalisation produces (dssvd.c:277 DSArrowBidiag + bdsqr), sorted by largest value after the two locked ones; with the
    extra row DSUpdateExtraRow leaves it equal to the last row of U (test8.c:78-82)"""
    n, l, k = 10, 2, 5
    ds, ld = _ds("svd", n, l, k)
    S.DSSVDSetDimensions(ds, n)
    S.DSSetCompact(ds, 1)
    S.DSSetExtraRow(ds, 1 if extrarow else 0)
    T, pT = _array(ds, SL.DS_MAT_T, ld, 3, real=True)
    for i in range(n):
        T[i, 0] = i + 1.0
    for i in range(l, n - 1):
        T[i, 1] = 1.0
    if extrarow:
        T[n - 1, 1] = 1.0
    _release(ds, SL.DS_MAT_T, pT, real=True)
    S.DSSetState(ds, DS_STATE_RAW)
    w, _ = _solve_sorted(ds, n, cmp="SlepcCompareLargestReal")
    if extrarow:
        S.DSUpdateExtraRow(ds)
    gold = [1.00000, 2.00000, 10.23838, 9.05119, 8.03408, 7.07377, 6.18432, 4.84902, 3.90608, 2.94118]
    assert [f"{x:.5f}" for x in w] == [f"{x:.5f}" for x in gold]
    M = np.diag(np.arange(1.0, n + 1))                      # numpy on the same matrix: arrow column k over rows l..k-1, then bidiagonal
    for i in range(l, k):
        M[i, k] = 1.0
    for i in range(k, n - 1):
        M[i, i + 1] = 1.0
    assert np.allclose(np.sort(w[l:]), np.sort(np.linalg.svd(M[l:, l:], compute_uv=False)), rtol=1e-13, atol=1e-13)
    if extrarow:
        T, pT = _array(ds, SL.DS_MAT_T, ld, 3, real=True)
        U, pU = _array(ds, SL.DS_MAT_U, ld, ld)
        assert abs(sum(T[i, 1] - U[n - 1, i] for i in range(l, n))) <= 10 * np.finfo(float).eps
        _release(ds, SL.DS_MAT_T, pT, real=True); _release(ds, SL.DS_MAT_U, pU)
    S.DSDestroy(ctypes.byref(ds))
