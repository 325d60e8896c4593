"""ctypes bindings to the host LAPACK/BLAS used by the ORACLE (test infrastructure only).

The reference calls LAPACK through PETSc's `LAPACKxxx_` macros (e.g. `LAPACKsteqr_`
at /root/reference/src/sys/classes/ds/impls/hep/dshep.c:408).  In this image the only host
LAPACK is the OpenBLAS bundled with scipy (`scipy.libs/libscipy_openblas-*.so`, LP64 symbols
with a `scipy_` prefix).  scipy.linalg.lapack does not wrap dsteqr/dhseqr/dtrevc/dbdsqr, so we
bind the Fortran symbols directly.  Everything is column-major (order='F') float64.
"""
import ctypes
import glob
import os

import numpy as np

_c_int = ctypes.c_int
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def find_openblas():
    import scipy
    pat = os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas*.so")
    hits = sorted(glob.glob(pat))
    if not hits:
        raise RuntimeError("scipy-bundled OpenBLAS not found (needed for host LAPACK)")
    return os.path.realpath(hits[0])


_lib = ctypes.CDLL(find_openblas(), mode=ctypes.RTLD_GLOBAL)


def _f(name):
    return getattr(_lib, "scipy_" + name + "_")


def _d(a):
    assert a.dtype == np.float64
    return a.ctypes.data_as(_dp)


def _i(v):
    return ctypes.byref(_c_int(int(v)))


def steqr(compz, d, e, Z, ldz, n, zoff=0):
    """dsteqr on d[0:n], e[0:n-1]; Z is a flat F-ordered buffer, zoff an element offset."""
    work = np.zeros(max(1, 2 * n - 2))
    info = _c_int(0)
    zp = ctypes.cast(ctypes.addressof(_d(Z).contents) + 8 * zoff, _dp)
    _f("dsteqr")(ctypes.c_char_p(compz.encode()), _i(n), _d(d), _d(e), zp, _i(ldz), _d(work),
                 ctypes.byref(info), ctypes.c_size_t(1))
    if info.value:
        raise RuntimeError(f"dsteqr info={info.value}")


def lartg(f, g):
    c = ctypes.c_double()
    s = ctypes.c_double()
    r = ctypes.c_double()
    _f("dlartg")(ctypes.byref(ctypes.c_double(f)), ctypes.byref(ctypes.c_double(g)), ctypes.byref(c),
                 ctypes.byref(s), ctypes.byref(r))
    return c.value, s.value, r.value


def gehrd(n, ilo, ihi, A, lda):
    tau = np.zeros(max(1, n))
    lwork = 6 * lda
    work = np.zeros(lwork)
    info = _c_int(0)
    _f("dgehrd")(_i(n), _i(ilo), _i(ihi), _d(A), _i(lda), _d(tau), _d(work), _i(lwork), ctypes.byref(info))
    if info.value:
        raise RuntimeError(f"dgehrd info={info.value}")
    return tau


def orghr(n, ilo, ihi, Q, ldq, tau):
    lwork = 6 * ldq
    work = np.zeros(lwork)
    info = _c_int(0)
    _f("dorghr")(_i(n), _i(ilo), _i(ihi), _d(Q), _i(ldq), _d(tau), _d(work), _i(lwork), ctypes.byref(info))
    if info.value:
        raise RuntimeError(f"dorghr info={info.value}")


def hseqr(n, ilo, ihi, H, ldh, wr, wi, Z, ldz):
    lwork = 6 * ldh
    work = np.zeros(lwork)
    info = _c_int(0)
    _f("dhseqr")(ctypes.c_char_p(b"S"), ctypes.c_char_p(b"V"), _i(n), _i(ilo), _i(ihi), _d(H), _i(ldh),
                 _d(wr), _d(wi), _d(Z), _i(ldz), _d(work), _i(lwork), ctypes.byref(info),
                 ctypes.c_size_t(1), ctypes.c_size_t(1))
    if info.value:
        raise RuntimeError(f"dhseqr info={info.value}")


def trexc(n, T, ldt, Q, ldq, ifst, ilst):
    work = np.zeros(max(1, n))
    info = _c_int(0)
    a = _c_int(ifst)
    b = _c_int(ilst)
    _f("dtrexc")(ctypes.c_char_p(b"V"), _i(n), _d(T), _i(ldt), _d(Q), _i(ldq), ctypes.byref(a),
                 ctypes.byref(b), _d(work), ctypes.byref(info), ctypes.c_size_t(1))
    if info.value:
        raise RuntimeError(f"dtrexc info={info.value}")


def trevc_right_some(n, T, ldt, select, VR, ldvr, mm, vroff=0):
    """dtrevc('R','S',select,...) writing mm columns at VR+vroff. Returns mout."""
    sel = np.ascontiguousarray(select, dtype=np.int32)
    work = np.zeros(3 * max(1, ldt))
    info = _c_int(0)
    mout = _c_int(0)
    vp = ctypes.cast(ctypes.addressof(_d(VR).contents) + 8 * vroff, _dp)
    _f("dtrevc")(ctypes.c_char_p(b"R"), ctypes.c_char_p(b"S"), sel.ctypes.data_as(_ip), _i(n), _d(T), _i(ldt),
                 vp, _i(ldvr), vp, _i(ldvr), _i(mm), ctypes.byref(mout), _d(work), ctypes.byref(info),
                 ctypes.c_size_t(1), ctypes.c_size_t(1))
    if info.value:
        raise RuntimeError(f"dtrevc info={info.value}")
    return mout.value


def trevc_right_all(n, T, ldt, VR, ldvr, back=True):
    """dtrevc('R','B' or 'A',NULL,...): all right eigenvectors; 'B' back-transforms with the matrix in VR."""
    work = np.zeros(3 * max(1, ldt))
    info = _c_int(0)
    mout = _c_int(0)
    _f("dtrevc")(ctypes.c_char_p(b"R"), ctypes.c_char_p(b"B" if back else b"A"), None, _i(n), _d(T), _i(ldt),
                 _d(VR), _i(ldvr), _d(VR), _i(ldvr), _i(n), ctypes.byref(mout), _d(work), ctypes.byref(info),
                 ctypes.c_size_t(1), ctypes.c_size_t(1))
    if info.value:
        raise RuntimeError(f"dtrevc info={info.value}")
    return mout.value


def bdsqr(uplo, n, ncvt, nru, d, e, VT, ldvt, U, ldu, off=0):
    work = np.zeros(4 * max(1, n))
    info = _c_int(0)
    vtp = ctypes.cast(ctypes.addressof(_d(VT).contents) + 8 * off, _dp)
    up = ctypes.cast(ctypes.addressof(_d(U).contents) + 8 * off, _dp)
    _f("dbdsqr")(ctypes.c_char_p(uplo.encode()), _i(n), _i(ncvt), _i(nru), _i(0), _d(d), _d(e), vtp, _i(ldvt),
                 up, _i(ldu), None, _i(max(1, ldu)), _d(work), ctypes.byref(info), ctypes.c_size_t(1))
    if info.value:
        raise RuntimeError(f"dbdsqr info={info.value}")
