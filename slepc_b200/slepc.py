"""ctypes binding of libb2kslepc.so (include/b2kslepc.h): the C host side that mirrors SLEPc's
BV / DS / ST / EPS / SVD interface for the Krylov hot path.  Function names and argument order are the
C ones (which are SLEPc's); this module only adds error translation and a few numpy conveniences.

No fallback: the library must have been built (`make`), and any object that touches basis vectors needs
`initialize()` to have found a GPU.  The oracle under oracle/ is never imported from here.
"""
import ctypes
import os
import re

import numpy as np

from . import _b2k

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "lib", "libb2kslepc.so")
HEADER = os.path.join(ROOT, "include", "b2kslepc.h")

c_int, c_dbl, c_vp = ctypes.c_int, ctypes.c_double, ctypes.c_void_p

_SCALAR_TYPES = {
    "int": c_int, "PetscInt": c_int, "PetscBool": c_int, "PetscErrorCode": c_int,
    "double": c_dbl, "PetscScalar": c_dbl, "PetscReal": c_dbl,
    "int64_t": ctypes.c_int64, "uint64_t": ctypes.c_uint64, "size_t": ctypes.c_size_t,
}
_ENUMS = ("B2KMemType", "NormType", "PetscViewerFormat", "BVOrthogType", "BVOrthogRefineType", "BVOrthogBlockType", "DSStateType", "DSMatType",
          "EPSProblemType", "EPSWhich", "EPSConvergedReason", "EPSErrorType", "EPSConv", "SVDWhich", "SVDConvergedReason",
          "SVDErrorType")


def parse_header(path):
    """{function name: (restype, [argtypes])} for every prototype of a C header of this repo."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    src = re.sub(r"^\s*#.*$", " ", src, flags=re.M)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b([A-Za-z_]\w*)\s*\(([^;{}()]*(?:\([^()]*\)[^;{}()]*)*)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if ret.startswith("typedef") or not ret:
            continue
        argtypes = []
        if args and args != "void":
            depth, cur, parts = 0, "", []
            for ch in args:
                if ch == "(":
                    depth += 1
                elif ch == ")":
                    depth -= 1
                if ch == "," and depth == 0:
                    parts.append(cur)
                    cur = ""
                else:
                    cur += ch
            parts.append(cur)
            for p in parts:
                p = p.strip()
                if "*" in p or "[" in p or "(" in p:
                    argtypes.append(c_vp)
                    continue
                toks = [t for t in re.split(r"\s+", p) if t not in ("const", "unsigned", "struct")]
                t = toks[0]
                if t in _SCALAR_TYPES:
                    argtypes.append(_SCALAR_TYPES[t])
                elif t in _ENUMS:
                    argtypes.append(c_int)
                else:
                    argtypes.append(c_vp)          # opaque handles and function-pointer typedefs
        if "*" in ret:
            restype = ctypes.c_char_p if "char" in ret else c_vp
        else:
            rt = [t for t in re.split(r"\s+", ret) if t not in ("const", "extern", "static", "inline")]
            restype = _SCALAR_TYPES.get(rt[-1], c_vp if rt[-1] != "void" else None)
        protos[name] = (restype, argtypes)
    return protos


_lib = None
PROTOTYPES = None


class SlepcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


def load():
    global _lib, PROTOTYPES
    if _lib is not None:
        return _lib
    _b2k.load()                     # libb200krylov.so first (RTLD_GLOBAL) so that the dependency resolves in-tree
    if not os.path.exists(LIB_PATH):
        raise _b2k.B2KError(f"{LIB_PATH} not found: run `make` (or __graft_entry__.build()); there is no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    PROTOTYPES = parse_header(HEADER)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def chk(ierr):
    if ierr:
        lib = load()
        msg = lib.B2KGetLastErrorMessage().decode(errors="replace")
        lib.B2KClearError()
        raise SlepcError(ierr, msg)


class _Lib:
    """attribute access returns checked callables: S.BVCreate(byref(h)) raises SlepcError on a non-zero code"""

    def __getattr__(self, name):
        fn = getattr(load(), name)
        if PROTOTYPES.get(name, (c_int,))[0] is not c_int:
            return fn

        def call(*a):
            chk(fn(*a))
        call.__name__ = name
        setattr(self, name, call)
        return call


S = _Lib()

# enum values of include/b2kslepc.h
MEM_HOST, MEM_DEVICE = 0, 1
NORM_1, NORM_2, NORM_FROBENIUS, NORM_INFINITY = 0, 1, 2, 3
BV_ORTHOG_CGS, BV_ORTHOG_MGS = 0, 1
BV_ORTHOG_REFINE_IFNEEDED, BV_ORTHOG_REFINE_NEVER, BV_ORTHOG_REFINE_ALWAYS = 0, 1, 2
BV_ORTHOG_BLOCK_GS, BV_ORTHOG_BLOCK_CHOL, BV_ORTHOG_BLOCK_TSQR, BV_ORTHOG_BLOCK_TSQRCHOL, BV_ORTHOG_BLOCK_SVQB = 0, 1, 2, 3, 4
EPS_HEP, EPS_GHEP, EPS_NHEP = 1, 2, 3
EPS_LARGEST_MAGNITUDE, EPS_SMALLEST_MAGNITUDE, EPS_LARGEST_REAL, EPS_SMALLEST_REAL = 1, 2, 3, 4
EPS_LARGEST_IMAGINARY, EPS_SMALLEST_IMAGINARY, EPS_TARGET_MAGNITUDE, EPS_TARGET_REAL = 5, 6, 7, 8
EPS_ERROR_ABSOLUTE, EPS_ERROR_RELATIVE = 0, 1
EPS_CONVERGED_TOL, EPS_CONVERGED_USER, EPS_DIVERGED_ITS, EPS_DIVERGED_BREAKDOWN, EPS_CONVERGED_ITERATING = 1, 2, -1, -2, 0
SVD_LARGEST, SVD_SMALLEST = 0, 1
SVD_ERROR_ABSOLUTE, SVD_ERROR_RELATIVE = 0, 1
DS_MAT_A, DS_MAT_T, DS_MAT_Q, DS_MAT_X, DS_MAT_U, DS_MAT_V = 0, 3, 5, 7, 9, 10
PETSC_DETERMINE, PETSC_DEFAULT, PETSC_CURRENT = -1, -2, -3


def initialize(device=0):
    S.B2KInitialize(device)


def finalize():
    S.B2KFinalize()


def _ptr(a):
    return a.ctypes.data_as(c_vp) if a is not None else None


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Handle:
    _destroy = None

    def __init__(self):
        self.h = c_vp()

    def destroy(self):
        if self.h and self._destroy:
            getattr(S, self._destroy)(ctypes.byref(self.h))
        self.h = c_vp()

    @property
    def ref(self):
        return ctypes.byref(self.h)


class Vec(Handle):
    _destroy = "VecDestroy"

    @classmethod
    def wrap(cls, h, own=False):
        v = cls()
        v.h = h if isinstance(h, c_vp) else c_vp(h)
        if not own:
            v._destroy = None
        return v

    def size(self):
        n = c_int()
        S.VecGetLocalSize(self.h, ctypes.byref(n))
        return n.value

    def set_values(self, a):
        a = _f64(a)
        assert a.size == self.size()
        S.VecSetValuesHost(self.h, _ptr(a))

    def get_values(self):
        out = np.empty(self.size())
        S.VecGetValuesHost(self.h, _ptr(out))
        return out

    def norm(self):
        v = c_dbl()
        S.VecNorm(self.h, NORM_2, ctypes.byref(v))
        return v.value


class Mat(Handle):
    _destroy = "MatDestroy"

    @classmethod
    def seqdense(cls, a):
        """sequential dense host Mat holding a COPY of the 2-D array a (column-major)"""
        a = np.asfortranarray(a, dtype=np.float64)
        m = cls()
        S.MatCreateSeqDense(a.shape[0], a.shape[1], None, m.ref)
        p = c_vp()
        S.MatDenseGetArray(m.h, ctypes.byref(p))
        ctypes.memmove(p.value, a.ctypes.data, a.nbytes)
        m.shape = a.shape
        return m

    def dense_array(self):
        p, lda, M, N = c_vp(), c_int(), c_int(), c_int()
        S.MatDenseGetArray(self.h, ctypes.byref(p))
        S.MatDenseGetLDA(self.h, ctypes.byref(lda))
        S.MatGetSize(self.h, ctypes.byref(M), ctypes.byref(N))
        buf = (ctypes.c_double * (lda.value * N.value)).from_address(p.value)
        return np.frombuffer(buf, dtype=np.float64).reshape((lda.value, N.value), order="F")[:M.value, :].copy()

    @classmethod
    def b200csr(cls, A, rstart=0, rend=None, cstart=0, cend=None):
        """rows [rstart,rend) of the scipy CSR matrix A (global column indices) → HBM"""
        A = A.tocsr()
        A.sort_indices()
        M, N = A.shape
        rend = M if rend is None else rend
        cend = N if cend is None else cend
        loc = A[rstart:rend]
        rp, ci, v = _i32(loc.indptr), _i32(loc.indices), _f64(loc.data)
        m = cls()
        S.MatCreateB200CSR(M, N, rstart, rend, _ptr(rp), _ptr(ci), _ptr(v), cstart, cend, m.ref)
        return m

    @classmethod
    def laplacian(cls, dim, nx, ny=1, nz=1):
        m = cls()
        S.MatCreateB200Laplacian(dim, nx, ny, nz, m.ref)
        return m

    def transpose(self):
        t = Mat()
        S.MatB200CSRTranspose(self.h, t.ref)
        return t

    def create_vecs(self):
        r, l = Vec(), Vec()
        S.MatCreateVecs(self.h, r.ref, l.ref)
        return r, l

    def local_size(self):
        m, n = c_int(), c_int()
        S.MatGetLocalSize(self.h, ctypes.byref(m), ctypes.byref(n))
        return m.value, n.value

    def mult(self, x, y):
        S.MatMult(self.h, x.h, y.h)


class BV(Handle):
    _destroy = "BVDestroy"

    @classmethod
    def create(cls, n, m, N=None, bvtype="b200", mem=MEM_DEVICE, rstart=0):
        bv = cls()
        S.BVCreate(bv.ref)
        t = Vec()
        S.VecCreateWithArray(mem, n, n if N is None else N, None, t.ref)
        _set_vec_rstart(t, rstart)
        S.BVSetSizesFromVec(bv.h, t.h, m)
        t.destroy()
        S.BVSetType(bv.h, bvtype.encode())
        return bv

    @classmethod
    def wrap(cls, h):
        bv = cls()
        bv.h = h
        bv._destroy = None
        return bv

    def sizes(self):
        n, N, m = c_int(), c_int(), c_int()
        S.BVGetSizes(self.h, ctypes.byref(n), ctypes.byref(N), ctypes.byref(m))
        return n.value, N.value, m.value

    def set_column(self, j, a):
        S.BVSetColumnHost(self.h, j, _ptr(_f64(a)))

    def get_column(self, j):
        out = np.empty(self.sizes()[0])
        S.BVGetColumnHost(self.h, j, _ptr(out))
        return out

    def to_numpy(self, ncols=None):
        n, _, m = self.sizes()
        ncols = m if ncols is None else ncols
        out = np.empty((n, ncols), order="F")
        for j in range(ncols):
            out[:, j] = self.get_column(j)
        return out

    def from_numpy(self, a):
        for j in range(a.shape[1]):
            self.set_column(j, a[:, j])

    def set_active(self, l, k):
        S.BVSetActiveColumns(self.h, l, k)

    def counters(self):
        a, b = ctypes.c_int64(), ctypes.c_int64()
        S.BVGetCounters(self.h, ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value


def _set_vec_rstart(vec, rstart):
    """first global row of a row-partitioned template vector (test/driver helper)"""
    if rstart:
        S.B2KVecSetOwnershipStart(vec.h, int(rstart))


class EPS(Handle):
    _destroy = "EPSDestroy"

    def __init__(self, A=None, hermitian=True, B=None):
        super().__init__()
        S.EPSCreate(self.ref)
        if A is not None:
            S.EPSSetOperators(self.h, A.h, B.h if B is not None else None)
            S.EPSSetProblemType(self.h, (EPS_GHEP if B is not None else EPS_HEP) if hermitian else EPS_NHEP)

    def st_sinvert(self, sigma=0.0):
        """-st_type sinvert -eps_target sigma"""
        st = c_vp()
        S.EPSGetST(self.h, ctypes.byref(st))
        S.STSetType(st, b"sinvert")
        S.EPSSetTarget(self.h, sigma)
        S.EPSSetWhichEigenpairs(self.h, EPS_TARGET_MAGNITUDE)

    def ksp_iterations(self):
        st, ksp, n = c_vp(), c_vp(), c_int()
        S.EPSGetST(self.h, ctypes.byref(st))
        S.STGetKSP(st, ctypes.byref(ksp))
        S.KSPGetTotalIterations(ksp, ctypes.byref(n))
        return n.value

    def bv(self):
        h = c_vp()
        S.EPSGetBV(self.h, ctypes.byref(h))
        return BV.wrap(h)

    def solve(self):
        S.EPSSolve(self.h)

    def cycles(self, n):
        done = c_int()
        S.EPSSolveRestartCycles(self.h, n, ctypes.byref(done))
        return done.value

    @property
    def nconv(self):
        n = c_int()
        S.EPSGetConverged(self.h, ctypes.byref(n))
        return n.value

    @property
    def its(self):
        n = c_int()
        S.EPSGetIterationNumber(self.h, ctypes.byref(n))
        return n.value

    @property
    def reason(self):
        n = c_int()
        S.EPSGetConvergedReason(self.h, ctypes.byref(n))
        return n.value

    def eigenvalue(self, i):
        re, im = c_dbl(), c_dbl()
        S.EPSGetEigenvalue(self.h, i, ctypes.byref(re), ctypes.byref(im))
        return re.value, im.value

    def eigenpair(self, i, vr, vi=None):
        re, im = c_dbl(), c_dbl()
        S.EPSGetEigenpair(self.h, i, ctypes.byref(re), ctypes.byref(im), vr.h, vi.h if vi is not None else None)
        return re.value, im.value

    def error(self, i, kind=EPS_ERROR_RELATIVE):
        e = c_dbl()
        S.EPSComputeError(self.h, i, kind, ctypes.byref(e))
        return e.value

    def dims(self):
        a, b, c = c_int(), c_int(), c_int()
        S.EPSGetDimensions(self.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
        return a.value, b.value, c.value


class SVD(Handle):
    _destroy = "SVDDestroy"

    def __init__(self, A=None, AT=None):
        super().__init__()
        S.SVDCreate(self.ref)
        if A is not None:
            S.SVDSetOperators(self.h, A.h, None)
        if AT is not None:
            S.SVDSetTransposeMatrix(self.h, AT.h)

    def solve(self):
        S.SVDSolve(self.h)

    @property
    def nconv(self):
        n = c_int()
        S.SVDGetConverged(self.h, ctypes.byref(n))
        return n.value

    @property
    def its(self):
        n = c_int()
        S.SVDGetIterationNumber(self.h, ctypes.byref(n))
        return n.value

    @property
    def reason(self):
        n = c_int()
        S.SVDGetConvergedReason(self.h, ctypes.byref(n))
        return n.value

    def triplet(self, i, u=None, v=None):
        s = c_dbl()
        S.SVDGetSingularTriplet(self.h, i, ctypes.byref(s), u.h if u is not None else None, v.h if v is not None else None)
        return s.value

    def error(self, i, kind=SVD_ERROR_RELATIVE):
        e = c_dbl()
        S.SVDComputeError(self.h, i, kind, ctypes.byref(e))
        return e.value

    def bvs(self):
        v, u = c_vp(), c_vp()
        S.SVDGetBV(self.h, ctypes.byref(v), ctypes.byref(u))
        return BV.wrap(v), BV.wrap(u)
