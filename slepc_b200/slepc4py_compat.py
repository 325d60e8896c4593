"""slepc4py / petsc4py look-alike over the B200 host layer: the object model and method names of the reference's Python binding
(src/binding/slepc4py/src/slepc4py/SLEPc/{EPS,SVD,BV,ST}.pyx and the petsc4py Mat / Vec it takes), so that a slepc4py script
for the Krylov path — demo/ex1.py, ex2.py, ex4.py style — runs with two changed import lines:

    from slepc_b200.slepc4py_compat import PETSc, SLEPc          # instead of: from petsc4py import PETSc; from slepc4py import SLEPc

Covered: PETSc.Mat (createAIJ from CSR / setValues + assemble for small matrices, createVecs, mult, getSize, getOwnershipRange),
PETSc.Vec (setArray / getArray / norm / dot / axpy / scale / copy), SLEPc.EPS (create, setOperators, setProblemType, setDimensions,
setTolerances, setWhichEigenpairs, setTarget, getST, getBV, setMonitor, solve, getConverged, getEigenvalue, getEigenpair,
computeError, getIterationNumber, getDimensions, getTolerances, getType, getConvergedReason), SLEPc.ST (setType, setShift),
SLEPc.SVD (create, setOperators, setDimensions, setTolerances, setWhichSingularTriplets, setImplicitTranspose, solve,
getConverged, getSingularTriplet / getValue, computeError), SLEPc.BV (the subset EPS.getBV() users touch).
Eigenvalues are returned as Python complex numbers like slepc4py does in real builds (kr + 1j*ki).

Everything runs on the GPU through BV type "b200" / Mat type "b200csr"; there is no CPU fallback (initialise with
PETSc.init(device)).  The options database, viewers and the other solver classes of slepc4py are outside this path."""
import ctypes

import numpy as np

from . import slepc as SL
from .slepc import S, c_dbl, c_int, c_vp


class _Enum:
    pass


class PETSc:
    DECIDE = DETERMINE = SL.PETSC_DETERMINE
    DEFAULT = SL.PETSC_DEFAULT
    CURRENT = SL.PETSC_CURRENT
    _initialized = False

    @staticmethod
    def init(device=0):
        if not PETSc._initialized:
            SL.initialize(device)
            PETSc._initialized = True

    class Sys:
        @staticmethod
        def Print(*args, **kw):
            print(*args, **kw)

    class Vec:
        def __init__(self, impl=None):
            self._v = impl

        def getSize(self):
            return self._v.size()

        getLocalSize = getSize

        def setArray(self, a):
            self._v.set_values(np.asarray(a, dtype=np.float64))

        def getArray(self):
            return self._v.get_values()

        array = property(getArray, setArray)

        def set(self, alpha):
            S.VecSet(self._v.h, float(alpha))

        def norm(self):
            return self._v.norm()

        def dot(self, other):
            d = c_dbl()
            S.VecDot(self._v.h, other._v.h, ctypes.byref(d))
            return d.value

        def axpy(self, alpha, x):
            S.VecAXPY(self._v.h, float(alpha), x._v.h)

        def scale(self, alpha):
            S.VecScale(self._v.h, float(alpha))

        def copy(self, result=None):
            if result is None:
                w = SL.Vec()
                S.VecCreateB200(self._v.size(), self._v.size(), w.ref)
                result = PETSc.Vec(w)
            S.VecCopy(self._v.h, result._v.h)
            return result

        def duplicate(self):
            w = SL.Vec()
            S.VecCreateB200(self._v.size(), self._v.size(), w.ref)
            return PETSc.Vec(w)

        def destroy(self):
            if self._v is not None:
                self._v.destroy()
                self._v = None

    class Mat:
        """sparse operator in HBM (Mat type b200csr).  Build it with createAIJ(size, csr=(indptr, indices, data)) — what petsc4py
        offers for bulk construction — or, for small matrices, with setValues / A[i, j] = v followed by assemble()."""

        def __init__(self):
            self._m = None
            self._shape = None
            self._coo = {}

        def create(self, comm=None):
            return self

        def setSizes(self, size):
            m, n = size if isinstance(size, (list, tuple)) else (size, size)
            self._shape = (int(m), int(n))

        def setFromOptions(self):
            pass

        def setUp(self):
            pass

        def getOwnershipRange(self):
            return 0, self._shape[0]

        def getSize(self):
            return self._shape

        def setValue(self, i, j, v, addv=None):
            self._coo[(int(i), int(j))] = float(v)

        def setValues(self, rows, cols, values, addv=None):
            rows, cols = np.atleast_1d(rows), np.atleast_1d(cols)
            vals = np.asarray(values, dtype=np.float64).reshape(len(rows), len(cols))
            for a, i in enumerate(rows):
                for b, j in enumerate(cols):
                    self._coo[(int(i), int(j))] = float(vals[a, b])

        def __setitem__(self, key, value):
            i, j = key
            n = self._shape[1]
            rows = range(*i.indices(self._shape[0])) if isinstance(i, slice) else [i % self._shape[0]]
            cols = range(*j.indices(n)) if isinstance(j, slice) else [j % n]
            vals = np.broadcast_to(np.asarray(value, dtype=np.float64), (len(rows), len(cols)))
            for a, r in enumerate(rows):
                for b, c in enumerate(cols):
                    self._coo[(r, c)] = float(vals[a, b])

        def assemble(self):
            import scipy.sparse as sp
            if self._m is None:
                keys = np.array(list(self._coo.keys()), dtype=np.int64).reshape(-1, 2)
                vals = np.array(list(self._coo.values()))
                A = sp.csr_matrix((vals, (keys[:, 0], keys[:, 1])), shape=self._shape)
                self._m = SL.Mat.b200csr(A)
                self._coo = {}

        assemblyBegin = assemblyEnd = assemble

        def createAIJ(self, size, csr=None, comm=None, nnz=None):
            import scipy.sparse as sp
            self.setSizes(size)
            if csr is not None:
                indptr, indices, data = csr
                self._m = SL.Mat.b200csr(sp.csr_matrix((np.asarray(data, dtype=np.float64), np.asarray(indices), np.asarray(indptr)), shape=self._shape))
            return self

        @classmethod
        def fromScipy(cls, A):
            m = cls()
            m._shape = A.shape
            m._m = SL.Mat.b200csr(A)
            return m

        def createVecs(self):
            r, l = self._m.create_vecs()
            return PETSc.Vec(r), PETSc.Vec(l)

        getVecs = createVecs

        def mult(self, x, y):
            S.MatMult(self._m.h, x._v.h, y._v.h)

        def multTranspose(self, x, y):
            S.MatMultTranspose(self._m.h, x._v.h, y._v.h)

        def destroy(self):
            if self._m is not None:
                self._m.destroy()
                self._m = None


class SLEPc:
    DECIDE = DETERMINE = SL.PETSC_DETERMINE
    DEFAULT = SL.PETSC_DEFAULT
    CURRENT = SL.PETSC_CURRENT

    class ST:
        class Type(_Enum):
            SHIFT, SINVERT = "shift", "sinvert"

        def __init__(self, h):
            self._h = h

        def setType(self, t):
            S.STSetType(self._h, t.encode())

        def getType(self):
            p = ctypes.c_char_p()
            S.STGetType(self._h, ctypes.byref(p))
            return p.value.decode()

        def setShift(self, sigma):
            S.STSetShift(self._h, float(sigma))

        def getShift(self):
            d = c_dbl()
            S.STGetShift(self._h, ctypes.byref(d))
            return d.value

    class BV:
        class OrthogType(_Enum):
            CGS, MGS = SL.BV_ORTHOG_CGS, SL.BV_ORTHOG_MGS

        class OrthogRefineType(_Enum):
            IFNEEDED, NEVER, ALWAYS = SL.BV_ORTHOG_REFINE_IFNEEDED, SL.BV_ORTHOG_REFINE_NEVER, SL.BV_ORTHOG_REFINE_ALWAYS

        def __init__(self, impl):
            self._bv = impl

        def getSizes(self):
            n, N, m = self._bv.sizes()
            return (n, N), m

        def getType(self):
            p = ctypes.c_char_p()
            S.BVGetType(self._bv.h, ctypes.byref(p))
            return p.value.decode()

        def setOrthogonalization(self, type=None, refine=None, eta=None, block=None):
            t, r, e, b = c_int(), c_int(), c_dbl(), c_int()
            S.BVGetOrthogonalization(self._bv.h, ctypes.byref(t), ctypes.byref(r), ctypes.byref(e), ctypes.byref(b))
            S.BVSetOrthogonalization(self._bv.h, t.value if type is None else type, r.value if refine is None else refine,
                                     e.value if eta is None else float(eta), b.value if block is None else block)

        def getColumn(self, j):
            return self._bv.get_column(j)

    class EPS:
        class ProblemType(_Enum):
            HEP, GHEP, NHEP = SL.EPS_HEP, SL.EPS_GHEP, SL.EPS_NHEP

        class Which(_Enum):
            LARGEST_MAGNITUDE, SMALLEST_MAGNITUDE, LARGEST_REAL, SMALLEST_REAL = 1, 2, 3, 4
            LARGEST_IMAGINARY, SMALLEST_IMAGINARY, TARGET_MAGNITUDE, TARGET_REAL = 5, 6, 7, 8

        class Type(_Enum):
            KRYLOVSCHUR = "krylovschur"

        class ErrorType(_Enum):
            ABSOLUTE, RELATIVE = SL.EPS_ERROR_ABSOLUTE, SL.EPS_ERROR_RELATIVE

        class ConvergedReason(_Enum):
            CONVERGED_TOL, CONVERGED_USER, DIVERGED_ITS, DIVERGED_BREAKDOWN, ITERATING = 1, 2, -1, -2, 0

        def __init__(self):
            self._e = None
            self._cb = None

        def create(self, comm=None):
            PETSc.init()
            self._e = SL.EPS()
            return self

        def setOperators(self, A, B=None):
            self._A, self._B = A, B
            S.EPSSetOperators(self._e.h, A._m.h, B._m.h if B is not None else None)

        def setProblemType(self, t):
            S.EPSSetProblemType(self._e.h, t)

        def setType(self, t):
            S.EPSSetType(self._e.h, t.encode())

        def getType(self):
            return "krylovschur"

        def setFromOptions(self):
            pass

        def setDimensions(self, nev=None, ncv=None, mpd=None):
            S.EPSSetDimensions(self._e.h, SL.PETSC_CURRENT if nev is None else nev, SL.PETSC_CURRENT if ncv is None else ncv, SL.PETSC_CURRENT if mpd is None else mpd)

        def getDimensions(self):
            return self._e.dims()

        def setTolerances(self, tol=None, max_it=None):
            S.EPSSetTolerances(self._e.h, float(SL.PETSC_CURRENT) if tol is None else tol, SL.PETSC_CURRENT if max_it is None else max_it)

        def getTolerances(self):
            t, m = c_dbl(), c_int()
            S.EPSGetTolerances(self._e.h, ctypes.byref(t), ctypes.byref(m))
            return t.value, m.value

        def setWhichEigenpairs(self, which):
            S.EPSSetWhichEigenpairs(self._e.h, which)

        def setTarget(self, target):
            S.EPSSetTarget(self._e.h, float(target))

        def getST(self):
            h = c_vp()
            S.EPSGetST(self._e.h, ctypes.byref(h))
            return SLEPc.ST(h)

        def getBV(self):
            return SLEPc.BV(self._e.bv())

        def setInitialSpace(self, space):
            space = space if isinstance(space, (list, tuple)) else [space]
            arr = (ctypes.c_void_p * len(space))(*[v._v.h for v in space])
            S.EPSSetInitialSpace(self._e.h, len(space), arr)

        def setMonitor(self, monitor, args=None, kargs=None):
            MON = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                   ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.c_void_p)

            def cb(_eps, its, nconv, eigr, eigi, errest, nest, _ctx):
                monitor(self, its, nconv, [complex(eigr[i], eigi[i]) for i in range(nest)], [errest[i] for i in range(nest)], *(args or ()), **(kargs or {}))
                return 0
            self._cb = MON(cb)
            S.EPSMonitorSet(self._e.h, self._cb, None)

        def setUp(self):
            S.EPSSetUp(self._e.h)

        def solve(self):
            self._e.solve()

        def getIterationNumber(self):
            return self._e.its

        def getConverged(self):
            return self._e.nconv

        def getConvergedReason(self):
            return self._e.reason

        def getEigenvalue(self, i):
            re, im = self._e.eigenvalue(i)
            return complex(re, im)

        def getEigenpair(self, i, Vr=None, Vi=None):
            re, im = c_dbl(), c_dbl()
            S.EPSGetEigenpair(self._e.h, i, ctypes.byref(re), ctypes.byref(im), Vr._v.h if Vr is not None else None, Vi._v.h if Vi is not None else None)
            return complex(re.value, im.value)

        def getEigenvector(self, i, Vr, Vi=None):
            self.getEigenpair(i, Vr, Vi)

        def getErrorEstimate(self, i):
            d = c_dbl()
            S.EPSGetErrorEstimate(self._e.h, i, ctypes.byref(d))
            return d.value

        def computeError(self, i, etype=None):
            return self._e.error(i, SL.EPS_ERROR_RELATIVE if etype is None else etype)

        def destroy(self):
            if self._e is not None:
                self._e.destroy()
                self._e = None

    class SVD:
        class Which(_Enum):
            LARGEST, SMALLEST = SL.SVD_LARGEST, SL.SVD_SMALLEST

        class Type(_Enum):
            TRLANCZOS = "trlanczos"

        class ErrorType(_Enum):
            ABSOLUTE, RELATIVE = SL.SVD_ERROR_ABSOLUTE, SL.SVD_ERROR_RELATIVE

        def __init__(self):
            self._s = None

        def create(self, comm=None):
            PETSc.init()
            self._s = SL.SVD()
            return self

        def setOperators(self, A, B=None):
            assert B is None, "generalized SVD is outside this path"
            self._A = A
            S.SVDSetOperators(self._s.h, A._m.h, None)

        setOperator = setOperators

        def setType(self, t):
            assert t == "trlanczos", "this build provides trlanczos"

        def getType(self):
            return "trlanczos"

        def setFromOptions(self):
            pass

        def setDimensions(self, nsv=None, ncv=None, mpd=None):
            S.SVDSetDimensions(self._s.h, SL.PETSC_CURRENT if nsv is None else nsv, SL.PETSC_CURRENT if ncv is None else ncv, SL.PETSC_CURRENT if mpd is None else mpd)

        def setTolerances(self, tol=None, max_it=None):
            S.SVDSetTolerances(self._s.h, float(SL.PETSC_CURRENT) if tol is None else tol, SL.PETSC_CURRENT if max_it is None else max_it)

        def setWhichSingularTriplets(self, which):
            S.SVDSetWhichSingularTriplets(self._s.h, which)

        def setImplicitTranspose(self, flag):
            S.SVDSetImplicitTranspose(self._s.h, 1 if flag else 0)

        def setTRLanczosOneSide(self, flag):
            S.SVDTRLanczosSetOneSide(self._s.h, 1 if flag else 0)

        def solve(self):
            self._s.solve()

        def getIterationNumber(self):
            return self._s.its

        def getConverged(self):
            return self._s.nconv

        def getValue(self, i):
            return self._s.triplet(i)

        def getSingularTriplet(self, i, U=None, V=None):
            sig = c_dbl()
            S.SVDGetSingularTriplet(self._s.h, i, ctypes.byref(sig), U._v.h if U is not None else None, V._v.h if V is not None else None)
            return sig.value

        def computeError(self, i, etype=None):
            return self._s.error(i, SL.SVD_ERROR_RELATIVE if etype is None else etype)

        def destroy(self):
            if self._s is not None:
                self._s.destroy()
                self._s = None
