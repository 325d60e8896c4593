"""ORACLE — CPU restatement (numpy + host LAPACK) of SLEPc 3.22's Krylov hot path.

THIS IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg may import it.  The product path (slepc_b200/) never does.

Parity status: PINNED — checked in tests/test_oracle_golden.py against the reference's own
golden outputs (bv/tests/output/test1_1_bv_type-svec.out, test2_1.out, test4_1.out, test13_1.out,
eps/tests/output/test4_1.out, eps/tutorials/output/ex2_1.out, ex5_1.out, svd/tests/output/test3_1.out)
and against analytic Laplacian spectra.  The reference itself (needs PETSc + MPI) cannot be built
in this image, see DESIGN.md.

Every function cites the reference file:line (relative to /root/reference) it restates.
All dense arrays are column-major float64 (PetscScalar=double, real build).
"""
import math

import numpy as np

from . import _lapack as L

EPS_MACH = np.finfo(np.float64).eps

# --------------------------------------------------------------------------------------------
# deterministic "random" vectors shared by the oracle, the C host code and the CUDA kernels
# (the reference uses PetscRandom, epssolve.c:852 / bvops.c:360-379, which is not reproducible
#  across implementations; SURVEY.md §8(d) prescribes a fixed hash-based start vector)
# --------------------------------------------------------------------------------------------
_M64 = (1 << 64) - 1


def hash_uniform(idx, seed):
    """value in [-1,1) for global row index idx (array of ints) — splitmix64 finaliser."""
    i = np.asarray(idx, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = (i + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed & _M64) * np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    u = (x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return 2.0 * u - 1.0


# --------------------------------------------------------------------------------------------
# comparators: slepcsc.c:152-200 (result>0 ⇔ second argument is "better")
# --------------------------------------------------------------------------------------------
def _absev(re, im):
    return math.hypot(re, im)


def cmp_largest_magnitude(ar, ai, br, bi):
    a, b = _absev(ar, ai), _absev(br, bi)
    return 1 if a < b else (-1 if a > b else 0)


def cmp_smallest_magnitude(ar, ai, br, bi):
    a, b = _absev(ar, ai), _absev(br, bi)
    return 1 if a > b else (-1 if a < b else 0)


def cmp_largest_real(ar, ai, br, bi):
    return 1 if ar < br else (-1 if ar > br else 0)


def cmp_smallest_real(ar, ai, br, bi):
    return 1 if ar > br else (-1 if ar < br else 0)


COMPARATORS = {
    "largest_magnitude": cmp_largest_magnitude,
    "smallest_magnitude": cmp_smallest_magnitude,
    "largest_real": cmp_largest_real,
    "smallest_real": cmp_smallest_real,
}


# --------------------------------------------------------------------------------------------
# BV  (bvimpl.h:25-113 struct; svec.c storage; bvblas.c kernels; bvorthog.c Gram-Schmidt)
# --------------------------------------------------------------------------------------------
class BV:
    CGS, MGS = 0, 1
    REFINE_IFNEEDED, REFINE_NEVER, REFINE_ALWAYS = 0, 1, 2

    def __init__(self, n, m, nc=0):
        self.n, self.m, self.nc = n, m, nc
        self.V = np.zeros((n, nc + m), order="F")       # svec.c:397: one long column-major array
        self.l, self.k = 0, m                           # active window, bvbasic.c BVSetActiveColumns
        self.orthog_type = BV.CGS                       # defaults bvfunc.c:176-179
        self.orthog_ref = BV.REFINE_IFNEEDED
        self.orthog_eta = 0.7071
        self.buffer = np.zeros((nc + m, m), order="F")  # bvbasic.c:757-789 (col 0 scratch, col j = h_j)
        self.rng_seed = 0x5EED
        self.npasses = 0                                # instrumentation only
        self.matrix = None                              # BVSetMatrix (bvbasic.c:497): inner product <x,y> = y^T B x

    def _ip(self, x):
        """BV_IPMatMult (bvimpl.h:147-157): B x, or x itself for the standard inner product."""
        return x if self.matrix is None else self.matrix @ x

    # ---- column access (bvbasic.c:1116-1137: physical column = nc+j) -------------------------
    def col(self, j):
        return self.V[:, self.nc + j]

    def set_active(self, l, k):
        self.l, self.k = l, k

    def set_random_column(self, j, row0=0):
        """deterministic replacement for BVSetRandomColumn (bvops.c:482)."""
        self.col(j)[:] = hash_uniform(np.arange(row0, row0 + self.n), self.rng_seed + j)

    # ---- level-2 ops on the active window, shapes as in svec.c:38-129 -------------------------
    def _act(self):
        return self.V[:, self.nc + self.l:self.nc + self.k]

    def dotvec(self, y):
        """BVDotVec (bvglobal.c:151 → svec.c:109 → bvblas.c:240): q = X(:,l:k)^T y."""
        return self._act().T @ y

    def multvec(self, alpha, beta, y, q):
        """BVMultVec (bvops.c:110 → svec.c:38 → bvblas.c:56): y = beta*y + alpha*X(:,l:k) q."""
        q = np.asarray(q, dtype=np.float64)[: self.k - self.l]
        if beta == 0.0:
            y[:] = alpha * (self._act() @ q)
        else:
            y[:] = beta * y + alpha * (self._act() @ q)

    def mult(self, alpha, beta, X, Q=None):
        """BVMult (bvops.c:22 → svec.c:17 → bvblas.c:24): Y = beta*Y + alpha*X*Q with
        Q sub-block rows X.l:X.k, cols Y.l:Y.k (svec.c:29)."""
        Y = self
        if Q is None:
            Y._act()[:] = beta * Y._act() + alpha * X._act()     # bvblas.c:112 BVAXPY
            return
        Qs = np.asarray(Q)[X.l:X.k, Y.l:Y.k]
        if beta == 0.0:
            Y._act()[:] = alpha * (X._act() @ Qs)
        else:
            Y._act()[:] = beta * Y._act() + alpha * (X._act() @ Qs)

    def mult_in_place(self, Q, s, e, trans=False):
        """BVMultInPlace (bvops.c:220 → svec.c:54 → bvblas.c:74): V(:,s:e) = V(:,l:k) Q(l:k, s:e)."""
        if s >= e:
            return
        Q = np.asarray(Q)
        Qs = Q.T if trans else Q
        B = Qs[self.l:self.k, s:e]
        self.V[:, self.nc + s:self.nc + e] = self._act() @ B

    def dot(self, Y):
        """BVDot(X,Y,M) (bvglobal.c:52 → svec.c:91 → bvblas.c:133): M = Y^T X on the active
        windows; returns the (Y.k-Y.l) x (X.k-X.l) block."""
        return Y._act().T @ self._act()

    def scale(self, alpha):
        self._act()[:] *= alpha

    def scale_column(self, j, alpha):
        """BVScaleColumn (bvops.c:419 → svec.c:150)."""
        self.col(j)[:] *= alpha

    def norm_column(self, j):
        """BVNormColumn (bvglobal.c:523 → svec.c:164 → bvlapack.c:37) NORM_2; sqrt(x^T B x) with BVSetMatrix (bvglobal.c:547)."""
        if self.matrix is not None:
            x = self.col(j)
            return self._safe_sqrt(float(x @ (self.matrix @ x)))
        return float(np.linalg.norm(self.col(j)))

    def norm_fro(self):
        return float(np.linalg.norm(self._act()))

    def copy_column(self, j, i):
        """BVCopyColumn(V,j,i) (bvbasic.c:1871): V[i] <- V[j]."""
        self.col(i)[:] = self.col(j)

    # ---- Gram-Schmidt (bvorthog.c) -------------------------------------------------------------
    @staticmethod
    def _safe_sqrt(alpha):
        """BV_SafeSqrt, bvimpl.h:121-141 (definite inner product)."""
        deftol = 10 * EPS_MACH
        if not alpha > -deftol:
            raise FloatingPointError(f"Invalid inner product: {alpha}")
        return 0.0 if alpha < 0.0 else math.sqrt(alpha)

    def _cgs1(self, j, want_onorm, want_norm):
        """BVOrthogonalizeCGS1, bvorthog.c:91-132, for a column (v==NULL)."""
        nc = self.nc
        c = self.buffer[:, 0]
        h = self.buffer[:, j]
        W = self.V[:, 0:nc + j]
        w = self.col(j)
        self.npasses += 1
        onorm = norm = None
        z = self._ip(w)                               # svec.c:117-120: the dot products use B w
        if want_onorm or want_norm:
            c[0:nc + j] = W.T @ z                     # BVDotColumnInc :32-47 (k=j+1)
            c[nc + j] = float(w @ z)
            beta = self._safe_sqrt(c[nc + j])         # BV_SquareRoot
        else:
            c[0:nc + j] = W.T @ z
        w -= W @ c[0:nc + j]                          # BVMultColumn(-1,1,j,c)
        if want_onorm:
            onorm = beta
        if want_norm:
            s = float(np.sum(c[0:nc + j] ** 2))       # BV_SquareSum
            nr = beta * beta - s
            norm = self.norm_column(j) if nr <= 0.0 else math.sqrt(nr)
        if j > 0 or nc > 0:
            h[0:nc + j] += c[0:nc + j]                # BV_AddCoefficients
        return onorm, norm

    def _mgs1(self, j, want_onorm, want_norm):
        """BVOrthogonalizeMGS1, bvorthog.c:52-85."""
        nc = self.nc
        c = self.buffer[:, 0]
        h = self.buffer[:, j]
        w = self.col(j)
        self.npasses += 1
        onorm = self.norm_column(j) if want_onorm else None
        for i in range(-nc, j):
            vi = self.col(i)
            d = float(self._ip(w) @ vi)               # bvorthog.c:69-74
            c[nc + i] = d
            w -= d * vi
        norm = self.norm_column(j) if want_norm else None
        if j > 0 or nc > 0:
            h[0:nc + j] += c[0:nc + j]
        return onorm, norm

    def orthogonalize_gs(self, j):
        """BVOrthogonalizeGS, bvorthog.c:145-217 (column form). Returns (norm, lindep)."""
        mgs = self.orthog_type == BV.MGS
        gs1 = self._mgs1 if mgs else self._cgs1
        nc = self.nc
        eta = self.orthog_eta
        self.buffer[0:nc + j, j] = 0.0                            # BV_CleanCoefficients
        if self.orthog_ref == BV.REFINE_IFNEEDED:
            onrm, nrm = gs1(j, True, True)
            l = 1
            while l < 3 and nrm and abs(nrm) < eta * abs(onrm):   # :180
                l += 1
                if mgs:
                    onrm = nrm
                    _, nrm = gs1(j, False, True)
                else:
                    onrm, nrm = gs1(j, True, True)
            lindep = not (nrm and abs(nrm) >= eta * abs(onrm))    # :186
        elif self.orthog_ref == BV.REFINE_NEVER:
            gs1(j, False, False)
            nrm = self.norm_column(j)
            lindep = not nrm
        else:
            gs1(j, False, False)
            onrm, nrm = gs1(j, True, True)
            lindep = not (nrm and abs(nrm) >= eta * abs(onrm))
        self.buffer[nc + j, j] = 0.0 if lindep else nrm           # BV_SetValue :212-213
        return nrm, lindep

    def orthogonalize_column(self, j):
        """BVOrthogonalizeColumn, bvorthog.c:315-340. Returns (H[l..j), norm, lindep)."""
        nrm, lindep = self.orthogonalize_gs(j)
        return self.buffer[self.nc + self.l:self.nc + j, j].copy(), nrm, lindep

    def orthonormalize_column(self, j, replace=False):
        """BVOrthonormalizeColumn, bvorthog.c:380-427."""
        nrm, lindep = self.orthogonalize_gs(j)
        if replace and (nrm == 0.0 or lindep):
            for _ in range(2):
                self.set_random_column(j)
                nrm, lindep = self.orthogonalize_gs(j)
                if not (nrm == 0.0 or lindep):
                    break
        if nrm != 1.0 and nrm != 0.0:
            self.scale_column(j, 1.0 / nrm)
        return nrm, lindep

    def orthogonalize_vec(self, v):
        """BVOrthogonalizeVec, bvorthog.c:249-273 with CGS/IFNEEDED: against columns -nc..k-1."""
        nc, k = self.nc, self.k
        W = self.V[:, 0:nc + k]
        eta = self.orthog_eta
        H = np.zeros(nc + k)

        def cgs1():
            c = W.T @ v
            beta = float(np.linalg.norm(v))
            v[:] -= W @ c
            s = float(np.sum(c * c))
            nr = beta * beta - s
            nrm = float(np.linalg.norm(v)) if nr <= 0.0 else math.sqrt(nr)
            H[:] += c
            return beta, nrm

        onrm, nrm = cgs1()
        l = 1
        while l < 3 and nrm and abs(nrm) < eta * abs(onrm):
            l += 1
            onrm, nrm = cgs1()
        lindep = not (nrm and abs(nrm) >= eta * abs(onrm))
        return H[nc + self.l:], nrm, lindep

    def orthogonalize_block_gs(self):
        """BVOrthogonalize with BV_ORTHOG_BLOCK_GS (bvorthog.c:560-594): column-by-column GS on the
        active columns, R upper triangular with the norms on the diagonal."""
        m = self.k
        R = np.zeros((m, m), order="F")
        for j in range(self.l, self.k):
            nrm, lindep = self.orthogonalize_gs(j)
            if nrm != 0.0 and nrm != 1.0:
                self.scale_column(j, 1.0 / nrm)
            R[0:j, j] = self.buffer[self.nc:self.nc + j, j]
            R[j, j] = nrm
        return R

    def orthogonalize_block(self, method="chol", want_R=False):
        """BVOrthogonalize with BV_ORTHOG_BLOCK_CHOL / SVQB (bvorthog.c:586-601, 660-675; bvlapack.c:136-202, 261-345)
        on the active columns l..k-1 (leading columns 0..l-1 assumed orthonormal).  Returns R (k x k; upper triangular
        for chol, full inv(S) for svqb) when want_R, with R[0:l, l:k] = V1^T V2 of the block Gram-Schmidt step."""
        import scipy.linalg as sla
        assert self.nc == 0
        l, k = self.l, self.k
        Rbuf = np.zeros((self.m, self.m), order="F")
        if l:                                                   # BVOrthogonalize_BlockGS bvorthog.c:488-499
            V1 = self.V[:, 0:l]
            R12 = V1.T @ self.V[:, l:k]
            Rbuf[0:l, l:k] = R12
            self.V[:, l:k] -= V1 @ R12
        n = k - l
        if method in ("tsqr", "tsqrchol"):                      # bvorthog.c:611-656: Householder QR (LAPACK geqrf, bvlapack.c:378-396)
            Qh, Rh = np.linalg.qr(self.V[:, l:k])               # R normalised to a non-negative diagonal (unique factor)
            sg = np.where(np.diag(Rh) < 0, -1.0, 1.0)
            Rh = Rh * sg[:, None]
            Rbuf[l:k, l:k] = np.triu(Rh)
            if method == "tsqr":
                self.V[:, l:k] = Qh * sg[None, :]               # orgqr
            else:
                self.V[:, l:k] = self.V[:, l:k] @ np.triu(sla.solve_triangular(Rh, np.eye(n), lower=False))   # BVMatTriInv + BVMultInPlace
            if want_R:
                R = np.zeros((k, k), order="F")
                for j in range(l, k):
                    R[0:j + 1, j] = Rbuf[0:j + 1, j]
                return R
            return None
        G = self.V[:, l:k].T @ self.V[:, l:k]                   # BVDot(V,V,R)
        if method == "chol":
            try:
                C = sla.cholesky(G, lower=False)
            except np.linalg.LinAlgError:                       # bvlapack.c:172-180
                C = sla.cholesky(G + 50.0 * EPS_MACH * np.eye(n), lower=False)
            S = sla.solve_triangular(C, np.eye(n), lower=False)
            S = np.triu(S)
            Rbuf[l:k, l:k] = np.triu(C)
        elif method == "svqb":
            D = 1.0 / np.sqrt(np.diag(G))
            lam, U = np.linalg.eigh((G * D[:, None]) * D[None, :])
            S = (U * D[:, None]) / np.sqrt(lam)[None, :]
            Rbuf[l:k, l:k] = (U.T * np.sqrt(lam)[:, None]) / D[None, :]
        else:
            raise ValueError(method)
        self.V[:, l:k] = self.V[:, l:k] @ S                     # BVMultInPlace(V,S,l,k)
        if want_R:
            R = np.zeros((k, k), order="F")
            if method == "chol":                                # BV_StoreCoeffsBlock_Default tri=TRUE: rows 0..j of column j
                for j in range(l, k):
                    R[0:j + 1, j] = Rbuf[0:j + 1, j]
            else:
                R[0:k, l:k] = Rbuf[0:k, l:k]
            return R
        return None

    # ---- Krylov recurrences (bvkrylov.c) ------------------------------------------------------
    def mat_mult_column(self, A, j):
        """BVMatMultColumn, bvops.c:862-885: V[j+1] = A V[j]."""
        self.col(j + 1)[:] = A @ self.col(j)

    def mat_arnoldi(self, A, H, k, m):
        """BVMatArnoldi, bvkrylov.c:56-113. H is (ldh x >=m) F-array. Returns (m, beta, breakdown)."""
        lindep = False
        beta = 0.0
        nc, mm = self.nc, self.m
        for j in range(k, m):
            self.mat_mult_column(A, j)
            beta, lindep = self.orthonormalize_column(j + 1, False)
            if lindep:
                m = j + 1
                break
        if H is not None:
            a = self.buffer
            for j in range(k, m - 1):
                H[0:j + 2, j] = a[nc:nc + j + 2, j + 1]
            H[0:m, m - 1] = a[nc:nc + m, m]
            if H.shape[0] > m:
                H[m, m - 1] = a[nc + m, m]
        return m, beta, lindep

    def mat_lanczos(self, A, T, k, m):
        """BVMatLanczos, bvkrylov.c:165-226. T is (ldt x 2+) F-array: col0 alpha, col1 beta."""
        lindep = False
        beta = 0.0
        nc = self.nc
        for j in range(k, m):
            self.mat_mult_column(A, j)
            beta, lindep = self.orthonormalize_column(j + 1, False)
            if lindep:
                m = j + 1
                break
        if T is not None:
            a = self.buffer
            for j in range(k, m):
                T[j, 0] = a[nc + j, j + 1]
                T[j, 1] = a[nc + j + 1, j + 1]
        return m, beta, lindep


# --------------------------------------------------------------------------------------------
# DS  (src/sys/classes/ds): HEP compact / NHEP / SVD compact, all with extra row
# --------------------------------------------------------------------------------------------
RAW, INTERMEDIATE, CONDENSED, TRUNCATED = 0, 1, 2, 3


class DS:
    def __init__(self, kind, ld, compare):
        assert kind in ("hep", "nhep", "svd")
        self.kind, self.ld = kind, ld
        self.compare = compare
        self.n = self.l = self.k = self.t = 0
        self.m = 0                                   # DSSVD: number of columns (dssvd.c:14-17)
        self.tsvd = 0
        self.state = RAW
        self.extrarow = True
        self.A = np.zeros((ld, ld), order="F")
        self.Q = np.zeros((ld, ld), order="F")       # also DS_MAT_U for svd
        self.Vm = np.zeros((ld, ld), order="F")      # DS_MAT_V (svd)
        self.X = np.zeros((ld, ld), order="F")
        self.T = np.zeros((ld, 3), order="F")        # compact storage dshep.c:24-74
        self.perm = np.zeros(ld, dtype=np.int64)

    def set_dimensions(self, n, l, k):
        """DSSetDimensions, dsops.c:129-165."""
        self.n, self.l, self.k = n, l, k
        self.t = n

    def set_state(self, s):
        self.state = s

    # ---- sorting helpers dspriv.c:172-270 --------------------------------------------------------
    def _sort_real(self, eig, perm):
        """DSSortEigenvaluesReal_Private, dspriv.c:224-243 (insertion sort on perm[l:t])."""
        n, l = self.t, self.l
        for i in range(l + 1, n):
            re = eig[perm[i]]
            j = i - 1
            result = self.compare(re, 0.0, eig[perm[j]], 0.0)
            while result < 0 and j >= l:
                perm[j], perm[j + 1] = perm[j + 1], perm[j]
                j -= 1
                if j >= l:
                    result = self.compare(re, 0.0, eig[perm[j]], 0.0)

    @staticmethod
    def _permute_columns(M, istart, iend, n, perm):
        """DSPermuteColumns_Private, dspriv.c:248-270 (destroys perm)."""
        for i in range(istart, iend):
            p = perm[i]
            if p != i:
                j = i + 1
                while perm[j] != i:
                    j += 1
                perm[j] = p
                perm[i] = i
                tmp = M[0:n, p].copy()
                M[0:n, p] = M[0:n, i]
                M[0:n, i] = tmp

    # ---- HEP (dshep.c) -----------------------------------------------------------------------------
    @staticmethod
    def arrow_tridiag(n, d, e, Q):
        """DSArrowTridiag, dshep.c:221-262 (Rutishauser). d,e,Q are views starting at (l,l)."""
        if n <= 2:
            return
        for j in range(n - 2):
            temp = e[j + 1]
            c, s, r = L.lartg(temp, e[j])
            e[j + 1] = r
            s = -s
            temp = d[j + 1]
            e[j] = c * s * (temp - d[j])
            d[j + 1] = s * s * d[j] + c * c * temp
            d[j] = c * c * d[j] + s * s * temp
            j2 = j + 2
            x = Q[0:j2, j].copy()
            y = Q[0:j2, j + 1].copy()
            Q[0:j2, j] = c * x + s * y
            Q[0:j2, j + 1] = c * y - s * x
            for i in range(j - 1, -1, -1):
                off = -s * e[i]
                e[i] = c * e[i]
                temp = e[i + 1]
                c, s, r = L.lartg(temp, off)
                e[i + 1] = r
                s = -s
                temp = (d[i] - d[i + 1]) * s - 2.0 * c * e[i]
                p = s * temp
                d[i + 1] += p
                d[i] -= p
                e[i] = -e[i] - c * temp
                x = Q[0:j2, i].copy()
                y = Q[0:j2, i + 1].copy()
                Q[0:j2, i] = c * x + s * y
                Q[0:j2, i + 1] = c * y - s * x

    def _solve_hep(self, wr):
        """DSSolve_HEP_QR (dshep.c:383-426) + DSIntermediate_HEP compact branch (:267-290)."""
        n, l, ld = self.n, self.l, self.ld
        d = self.T[:, 0]
        e = self.T[:, 1]
        n1 = max(0, self.k - l + 1)
        # DSSetIdentity(Q): dspriv.c (zero cols l..n-1 then unit diagonal)
        self.Q[:, l:n] = 0.0
        for i in range(l, n):
            self.Q[i, i] = 1.0
        if self.state < INTERMEDIATE:
            self.arrow_tridiag(n1, d[l:], e[l:], self.Q[l:, l:])
        wr[0:l] = d[0:l]
        L.steqr("V", d[l:n], e[l:n], self.Q[l:, l:], ld, n - l)
        wr[l:n] = d[l:n]
        e[0:n - 1] = 0.0

    def _sort_hep(self, wr):
        """DSSort_HEP, dshep.c:323-347 (rr == NULL)."""
        n, l = self.n, self.l
        d = self.T[:, 0]
        perm = self.perm
        self._sort_real(d, perm)
        for i in range(l, n):
            wr[i] = d[perm[i]]
        self._permute_columns(self.Q, l, n, n, perm)
        d[l:n] = wr[l:n]

    def _update_extra_row_hep(self):
        """DSUpdateExtraRow_HEP, dshep.c:349-381 (compact)."""
        n = self.n
        e = self.T[:, 1]
        beta = e[n - 1]
        for i in range(n):
            e[i] = beta * self.Q[n - 1, i]
        self.k = n

    # ---- NHEP (dsutil.c, dsnhep.c) -----------------------------------------------------------------
    def _solve_nhep(self, wr, wi):
        """DSSolve_NHEP_Private, dsutil.c:21-91."""
        n, l, ld = self.n, self.l, self.ld
        A, Q = self.A, self.Q
        ilo = l + 1
        Q[:, :] = 0.0
        for i in range(n):
            Q[i, i] = 1.0
        if n == 1:
            wr[0] = A[0, 0]
            wi[0] = 0.0
            return
        if self.state < INTERMEDIATE:
            tau = L.gehrd(n, ilo, n, A, ld)
            for j in range(n - 1):
                for i in range(j + 2, n):
                    Q[i, j] = A[i, j]
                    A[i, j] = 0.0
            L.orghr(n, ilo, n, Q, ld, tau)
        L.hseqr(n, ilo, n, A, ld, wr, wi, Q, ld)
        j = 0
        while j < l:
            if j == n - 1 or A[j + 1, j] == 0.0:
                wr[j] = A[j, j]
                wi[j] = 0.0
            else:
                wr[j] = A[j, j]
                wr[j + 1] = A[j, j]
                wi[j] = math.sqrt(abs(A[j + 1, j])) * math.sqrt(abs(A[j, j + 1]))
                wi[j + 1] = -wi[j]
                j += 1
            j += 1

    def _sort_nhep(self, wr, wi):
        """DSSort_NHEP_Total, dsutil.c:93-171 (selection sort with trexc)."""
        n, ld = self.n, self.ld
        T, Q = self.A, self.Q
        i = self.l
        while i < n - 1:
            re, im = wr[i], wi[i]
            pos = 0
            j = i + 1
            if im != 0:
                j = i + 2
            while j < n:
                if self.compare(re, im, wr[j], wi[j]) > 0:
                    re, im = wr[j], wi[j]
                    pos = j
                if wi[j] != 0:
                    j += 1
                j += 1
            if pos:
                L.trexc(n, T, ld, Q, ld, pos + 1, i + 1)
                j = i
                while j < n:
                    wr[j] = T[j, j]
                    if j < n - 1 and T[j + 1, j] != 0.0:
                        wi[j] = math.sqrt(abs(T[j + 1, j])) * math.sqrt(abs(T[j, j + 1]))
                        wr[j + 1] = wr[j]
                        wi[j + 1] = -wi[j]
                        j += 1
                    else:
                        wi[j] = 0.0
                    j += 1
            if wi[i] != 0:
                i += 1
            i += 1

    def _update_extra_row_nhep(self):
        """DSUpdateExtraRow_NHEP, dsnhep.c:318-340."""
        n = self.n
        x = self.A[n, 0:n].copy()
        self.A[n, 0:n] = self.Q[0:n, 0:n].T @ x
        self.k = n

    def vectors_nhep_some(self, k):
        """DSVectors_NHEP_Eigen_Some, dsnhep.c:101-166 (right). Returns (newk, rnorm)."""
        n, ld = self.n, self.ld
        A, X = self.A, self.X
        select = np.zeros(n, dtype=np.int32)
        select[k] = 1
        iscomplex = k < n - 1 and A[k + 1, k] != 0.0
        mm = 2 if iscomplex else 1
        if iscomplex:
            select[k + 1] = 1
        Y = X[:, k:]
        mout = L.trevc_right_some(n, A, ld, select, Y, ld, mm)
        assert mout == mm
        if self.state >= CONDENSED:
            w0 = X[0:n, k].copy()
            X[0:n, k] = self.Q[0:n, 0:n] @ w0
            if iscomplex:
                w1 = X[0:n, k + 1].copy()
                X[0:n, k + 1] = self.Q[0:n, 0:n] @ w1
            norm = float(np.linalg.norm(X[0:n, k]))
            if iscomplex:
                norm = math.hypot(norm, float(np.linalg.norm(X[0:n, k + 1])))
            X[0:n, k:k + mm] /= norm
        newk = k + 1 if iscomplex else k
        rnorm = math.hypot(X[n - 1, k], X[n - 1, k + 1]) if iscomplex else abs(X[n - 1, k])
        return newk, rnorm

    def vectors_nhep_all(self):
        """DSVectors_NHEP_Eigen_All, dsnhep.c:168-232: X = Q * eigvecs(T), column-normalised."""
        n, ld = self.n, self.ld
        A, X = self.A, self.X
        back = self.state >= CONDENSED
        if back:
            X[:, :] = self.Q
        L.trevc_right_all(n, A, ld, X, ld, back)
        i = 0
        while i < n:
            iscomplex = i < n - 1 and A[i + 1, i] != 0.0
            norm = float(np.linalg.norm(X[0:n, i]))
            cols = 1
            if iscomplex:
                norm = math.hypot(norm, float(np.linalg.norm(X[0:n, i + 1])))
                cols = 2
            X[0:n, i:i + cols] /= norm
            i += cols

    def get_truncate_size(self, l, n, k):
        """DSGetTruncateSize_Default, dsops.c:329-344 (NHEP only: do not split a 2x2 block)."""
        if self.kind != "nhep":
            return k
        if self.A[l + k, l + k - 1] != 0.0:
            if l + k < n - 1:
                k += 1
            else:
                k -= 1
        return k

    # ---- SVD (dssvd.c) ------------------------------------------------------------------------------
    @staticmethod
    def arrow_bidiag(n, d, e, Q, P):
        """DSArrowBidiag, dssvd.c:277-339."""
        if n <= 2:
            return

        def rot(M, j2, a, b, c, s):
            x = M[0:j2, a].copy()
            y = M[0:j2, b].copy()
            M[0:j2, a] = c * x + s * y
            M[0:j2, b] = c * y - s * x

        for j in range(n - 2):
            temp0 = e[j + 1]
            c, s, r = L.lartg(temp0, e[j])
            e[j + 1] = r
            s = -s
            j2 = j + 2
            rot(Q, j2, j, j + 1, c, s)
            temp0 = d[j + 1]
            temp1 = c * temp0
            temp2 = -s * d[j]
            ct, st, r = L.lartg(temp1, temp2)
            d[j + 1] = r
            st = -st
            e[j] = -c * st * d[j] + s * ct * temp0
            d[j] = c * ct * d[j] + s * st * temp0
            rot(P, j2, j, j + 1, ct, st)
            for i in range(j - 1, -1, -1):
                off = -st * e[i]
                e[i] = ct * e[i]
                temp0 = e[i + 1]
                c, s, r = L.lartg(temp0, off)
                e[i + 1] = r
                s = -s
                rot(Q, j2, i, i + 1, c, s)
                temp0 = d[i + 1]
                temp1 = -s * e[i] + c * temp0
                temp2 = c * e[i] + s * temp0
                off = -s * d[i]
                ct, st, r = L.lartg(temp1, off)
                d[i + 1] = r
                st = -st
                e[i] = -c * st * d[i] + ct * temp2
                d[i] = c * ct * d[i] + st * temp2
                rot(P, j2, i, i + 1, ct, st)

    def _solve_svd(self, wr):
        """DSSolve_SVD_QR (dssvd.c:400-457) + DSIntermediate_SVD compact branch (:344-398)."""
        n, m, l, ld = self.n, self.m, self.l, self.ld
        d = self.T[:, 0]
        e = self.T[:, 1]
        U, V = self.Q, self.Vm
        n1 = max(0, self.k - l + 1)
        U[:, :] = 0.0
        V[:, :] = 0.0
        for i in range(n):
            U[i, i] = 1.0
        for i in range(m):
            V[i, i] = 1.0
        if self.state < INTERMEDIATE:
            self.arrow_bidiag(n1, d[l:], e[l:], U[l:, l:], V[l:, l:])
        wr[0:l] = d[0:l]
        nn1, m1 = n - l, m - l
        nm = min(nn1, m1)
        Vt = np.zeros((ld, ld), order="F")
        Vt[l:m, l:m] = V[l:m, l:m].T
        L.bdsqr("U" if n >= m else "L", nm, m1, nn1, d[l:], e[l:], Vt[l:, l:], ld, U[l:, l:], ld)
        V[l:m, l:m] = Vt[l:m, l:m].T
        for i in range(l, min(n, m)):
            wr[i] = d[i]
        e[0:n - 1] = 0.0

    def _sort_svd(self, wr):
        """DSSort_SVD, dssvd.c:156-184 + DSPermuteBoth_Private dspriv.c:305-333."""
        l = self.l
        n = min(self.n, self.m)
        d = self.T[:, 0]
        perm = self.perm
        self._sort_real(d, perm)
        for i in range(l, n):
            wr[i] = d[perm[i]]
        for i in range(l, n):
            p = perm[i]
            if p != i:
                j = i + 1
                while perm[j] != i:
                    j += 1
                perm[j] = p
                perm[i] = i
                for M, rows in ((self.Q, self.n), (self.Vm, self.m)):
                    tmp = M[0:rows, p].copy()
                    M[0:rows, p] = M[0:rows, i]
                    M[0:rows, i] = tmp
        d[l:n] = wr[l:n]

    def _update_extra_row_svd(self):
        """DSUpdateExtraRow_SVD, dssvd.c:186-221 (compact)."""
        n, m = self.n, self.m
        e = self.T[:, 1]
        beta = e[m - 1]
        for i in range(n):
            e[i] = beta * self.Q[n - 1, i]
        self.k = m

    # ---- front-ends (dsops.c) ---------------------------------------------------------------------
    def solve(self, wr, wi=None):
        """DSSolve, dsops.c:723-743."""
        if self.state >= CONDENSED:
            return
        if self.kind == "hep":
            self._solve_hep(wr)
            if wi is not None:
                wi[self.l:self.n] = 0.0
        elif self.kind == "nhep":
            self._solve_nhep(wr, wi)
        else:
            self._solve_svd(wr)
        self.state = CONDENSED

    def sort(self, wr, wi=None):
        """DSSort, dsops.c:777-800."""
        self.perm[0:self.n] = np.arange(self.n)
        if self.kind == "hep":
            self._sort_hep(wr)
        elif self.kind == "nhep":
            self._sort_nhep(wr, wi)
        else:
            self._sort_svd(wr)

    def update_extra_row(self):
        getattr(self, "_update_extra_row_" + self.kind)()

    def vectors_rnorm(self, k):
        """DSVectors(ds,DS_MAT_X,&k,&rnorm): dshep.c:137-172 / dsnhep.c:101-166."""
        if self.kind == "hep":
            return k, abs(self.Q[self.n - 1, k])
        return self.vectors_nhep_some(k)

    def truncate(self, n, trim):
        """DSTruncate (dsops.c:232-254) → DSTruncate_HEP/NHEP/SVD (dshep.c:643, dsnhep.c:394, dssvd.c:223)."""
        l = self.l
        dense = self.kind == "nhep"
        if trim:
            if dense and self.extrarow:
                self.A[self.n, l:self.n] = 0.0
            self.l = 0
            self.k = 0
            self.n = n
            self.t = n
            if self.kind == "svd":
                self.m = n
                self.tsvd = n
        else:
            if dense and self.extrarow and self.k == self.n:
                self.A[n, l:n] = self.A[self.n, l:n]
                self.A[self.n, l:self.n] = 0.0
            self.k = n if self.extrarow else 0
            self.t = self.n
            if self.kind == "svd":
                self.tsvd = self.m
                self.m = n
            self.n = n
        self.state = RAW if trim else TRUNCATED

    def mat_rows(self, which):
        """DSMatGetSize rows, dsops.c:276-295 / dssvd.c:607-634."""
        if self.state == TRUNCATED:
            return self.tsvd if (self.kind == "svd" and which == "V") else self.t
        return self.m if (self.kind == "svd" and which == "V") else self.n


# --------------------------------------------------------------------------------------------
# EPS Krylov-Schur (krylovschur.c:93-336, epskrylov.c:207-295, epsdefault.c:224-302)
# --------------------------------------------------------------------------------------------
def eps_default_dims(n, nev, ncv=None, mpd=None):
    """EPSSetDimensions_Default, epssetup.c:648-676."""
    if ncv is not None:
        if not (ncv >= nev + 1 or (ncv == nev and ncv == n)):
            raise ValueError("The value of ncv must be at least nev+1")
    elif mpd is not None:
        ncv = min(n, nev + mpd)
    else:
        if nev < 500:
            ncv = min(n, max(2 * nev, nev + 15))
        else:
            mpd = 500
            ncv = min(n, nev + mpd)
    if mpd is None:
        mpd = ncv
    return ncv, mpd


class EPSResult:
    pass


class STOperator:
    """The transformed operator of ST as `Op @ x` (STApply_Generic, stsolve.c:16-25: y = P^{-1} M x):
         shift,   standard      Op = A - sigma I
         shift,   generalized   Op = B^{-1} (A - sigma B)          (shift.c:60-114)
         sinvert                Op = (A - sigma B)^{-1} [B]        (sinvert.c:79-150)
    The linear solves use a sparse LU (scipy splu), the reference's default KSPPREONLY + PCLU (stsles.c:104-131)."""

    def __init__(self, A, B=None, sigma=0.0, sinvert=False):
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla
        self.A, self.B, self.sigma, self.sinvert = A, B, sigma, sinvert
        n = A.shape[0]
        Bm = sp.identity(n, format="csc") if B is None else sp.csc_matrix(B)
        self.T = (sp.csc_matrix(A) - sigma * Bm).tocsc()
        self.lu = None
        if sinvert:
            self.lu = spla.splu(self.T)
        elif B is not None:
            self.lu = spla.splu(Bm)

    def __matmul__(self, x):
        if self.sinvert:
            return self.lu.solve(x if self.B is None else self.B @ x)
        y = self.T @ x
        return y if self.B is None else self.lu.solve(y)

    def back(self, re, im=0.0):
        """STBackTransform: shift.c:49-58, sinvert.c:51-77 (real eigenvalues)."""
        if not self.sinvert:
            return re + self.sigma, im
        if im == 0.0:
            return 1.0 / re + self.sigma, 0.0
        z = 1.0 / complex(re, im)
        return z.real + self.sigma, z.imag


def eps_krylovschur(A, n, nev, ncv=None, mpd=None, tol=1e-8, max_it=None, which="largest_magnitude",
                    hermitian=True, v0=None, keep=0.5, lock=True, orthog=None, seed=0x5EED, monitor=None,
                    B=None, sigma=0.0, sinvert=False, target=None, purify=True):
    """EPSSolve with -eps_type krylovschur.  Standard problem with ST=shift(0) by default; with B (symmetric positive
    definite, EPS_GHEP) and / or sinvert the operator is the ST's (STOperator), the basis is B-orthonormal (BVSetMatrix,
    epssetup.c:372-381), the projected values are compared after STBackTransform (SlepcMap_ST), convergence of sinvert is tested
    on the transformed value (epskrylov.c:246), the eigenvectors are purified and B-normalised (epsdefault.c:28-50).
    A: anything supporting A @ x. Follows EPSSetUp_KrylovSchur (krylovschur.c:93-194) and
    EPSSolve_KrylovSchur_Default (:227-336)."""
    st = None
    if B is not None or sinvert or sigma != 0.0:
        st = STOperator(A, B, sigma, sinvert)
        if target is None:
            target = sigma
        if sinvert and which == "largest_magnitude":
            which = "target_magnitude"                      # EPSSetWhichEigenpairs_Default epsdefault.c:210-220
    ncv, mpd = eps_default_dims(n, nev, ncv, mpd)
    if ncv > nev + mpd:
        raise ValueError("The value of ncv must not be larger than nev+mpd")
    if max_it is None:
        max_it = max(100, 2 * n // ncv)                     # krylovschur.c:113
    if which in ("target_magnitude", "target_real"):
        tg = 0.0 if target is None else target
        base = (lambda ar, ai, br, bi: cmp_smallest_magnitude(ar - tg, ai, br - tg, bi)) if which == "target_magnitude" else \
               (lambda ar, ai, br, bi: (abs(ar - tg) > abs(br - tg)) - (abs(ar - tg) < abs(br - tg)))      # slepcsc.c:233-260
    else:
        base = COMPARATORS[which]
    if st is not None:
        compare = lambda ar, ai, br, bi: base(*st.back(ar, ai), *st.back(br, bi))     # SlepcMap_ST, slepcsc.c:40-63
    else:
        compare = base
    Op = st if st is not None else A
    V = BV(n, ncv + 1)                                      # EPSAllocateSolution(eps,1) epssetup.c:692
    V.rng_seed = seed
    if B is not None and hermitian:
        V.matrix = B                                        # epssetup.c:372-381
    if orthog:
        V.orthog_type, V.orthog_ref, V.orthog_eta = orthog
    ld = ncv + 1
    ds = DS("hep" if hermitian else "nhep", ld, compare)   # krylovschur.c:153-168
    eigr = np.zeros(ncv + 1)
    eigi = np.zeros(ncv + 1)
    errest = np.zeros(ncv + 1)
    nconv = 0
    its = 0
    reason = 0
    nmatvec = 0

    # EPSGetStartVector(eps,0) epssolve.c:841-871
    if v0 is not None:
        V.col(0)[:] = v0
    else:
        V.set_random_column(0)
    if B is not None and hermitian:
        V.col(0)[:] = Op @ V.col(0).copy()                  # into the range of OP for definite generalized problems, epssolve.c:855-862
    _, norm, lindep = V.orthogonalize_column(0)
    if lindep or norm == 0.0:
        raise RuntimeError("Initial vector is zero or belongs to the deflation space")
    V.scale_column(0, 1.0 / norm)

    l = 0
    while reason == 0:
        its += 1
        nv = min(nconv + mpd, ncv)
        ds.set_dimensions(nv, nconv, nconv + l)
        if hermitian:
            nv, beta, breakdown = V.mat_lanczos(Op, ds.T, nconv + l, nv)
        else:
            nv, beta, breakdown = V.mat_arnoldi(Op, ds.A, nconv + l, nv)
        nmatvec += nv - (nconv + l)
        ds.set_dimensions(nv, nconv, nconv + l)
        ds.set_state(RAW if l else INTERMEDIATE)
        V.set_active(nconv, nv)

        ds.solve(eigr, eigi)
        ds.sort(eigr, eigi)
        ds.update_extra_row()

        # EPSKrylovConvergence(eps,FALSE,nconv,nv-nconv,beta,0,1,&k) epskrylov.c:207-295
        marker = -1
        k = nconv
        while k < nv:
            re, im = eigr[k], eigi[k]
            if st is not None and not sinvert:
                re, im = st.back(re, im)                    # only for STSHIFT, epskrylov.c:246
            newk, resnorm = ds.vectors_rnorm(k)
            resnorm *= beta
            w = math.hypot(re, im)
            errest[k] = resnorm / w if w != 0.0 else np.finfo(np.float64).max   # epsdefault.c:224
            if marker == -1 and errest[k] >= tol:
                marker = k
            if newk == k + 1:
                errest[k + 1] = errest[k]
                k += 1
            if marker != -1:
                break
            k += 1
        if marker != -1:
            k = marker
        # EPSStoppingBasic epsdefault.c:290-302
        if k >= nev:
            reason = 2          # EPS_CONVERGED_TOL
        elif its >= max_it:
            reason = -3         # EPS_DIVERGED_ITS
        if reason != 0 or breakdown or k == nv:
            l = 0
        else:
            l = max(1, int((nv - k) * keep))
            if not hermitian:
                l = ds.get_truncate_size(k, nv, l)
        if not lock and l > 0:
            l += k
            k = 0
        if reason == 0:
            if breakdown or k == nv:
                if k < nev:
                    # EPSGetStartVector(eps,k,&breakdown)
                    V.set_random_column(k)
                    if B is not None and hermitian:
                        V.col(k)[:] = Op @ V.col(k).copy()
                    V.set_active(0, k)  # not in the reference; GS below uses its own window
                    _, norm, lindep = V.orthogonalize_column(k)
                    V.set_active(nconv, nv)
                    if lindep:
                        reason = -1     # EPS_DIVERGED_BREAKDOWN
                    else:
                        V.scale_column(k, 1.0 / norm)
            else:
                ds.truncate(k + l, False)
        Q = ds.Q
        V.mult_in_place(Q, nconv, k + l)
        if reason == 0 and not breakdown:
            V.copy_column(nv, k + l)
        nconv = k
        if monitor:
            monitor(its, nconv, eigr, errest, nv)
    ds.truncate(nconv, True)

    res = EPSResult()
    res.nconv, res.its, res.reason, res.nmatvec = nconv, its, reason, nmatvec
    res.ncv, res.mpd, res.max_it = ncv, mpd, max_it
    if st is not None:                                      # EPSComputeValues: STBackTransform, epssolve.c:160
        for i in range(nconv):
            eigr[i], eigi[i] = st.back(eigr[i], eigi[i])
        compare = base                                      # the final ordering compares the back-transformed values (eps->sc has no map)
    res.eigr, res.eigi, res.errest = eigr[:nconv].copy(), eigi[:nconv].copy(), errest[:nconv].copy()
    res.V, res.ds = V, ds
    # EPSComputeVectors: Hermitian = Lanczos vectors as they are (epsdefault.c:28); NHEP = V*Z (:105-125)
    if hermitian:
        res.X = V.V[:, :nconv].copy()
        if B is not None and purify:                        # EPSComputeVectors_Hermitian epsdefault.c:28-50: EPS_Purify + BVNormalize (B-norm)
            for i in range(nconv):
                x = Op @ res.X[:, i]
                res.X[:, i] = x / math.sqrt(float(x @ (B @ x)))
    else:
        if nconv > 0:
            ds.vectors_nhep_all()      # state RAW after the trim ⇒ eigenvectors of T, no back-transform
            res.X = V.V[:, :nconv] @ ds.X[:nconv, :nconv]
        else:
            res.X = np.zeros((n, 0))
    # final ordering: SlepcSortEigenvalues(eps->sc,...) epssolve.c:178 — stable insertion by comparator
    perm = list(range(nconv))
    for i in range(1, nconv):
        j = i
        while j > 0 and compare(eigr[perm[j - 1]], eigi[perm[j - 1]], eigr[perm[j]], eigi[perm[j]]) > 0:
            perm[j - 1], perm[j] = perm[j], perm[j - 1]
            j -= 1
    res.perm = np.array(perm, dtype=np.int64)
    return res


def eps_relative_error(A, lam_r, lam_i, xr, xi=None):
    """EPSComputeError(EPS_ERROR_RELATIVE) → EPSComputeResidualNorm_Private, epssolve.c:666-722,
    742-815: ||A x - lambda x||_2 / |lambda| for real arithmetic (complex pair handled jointly)."""
    if lam_i == 0.0:
        r = A @ xr - lam_r * xr
        nr = float(np.linalg.norm(r))
        nx = float(np.linalg.norm(xr))
    else:
        u = A @ xr - lam_r * xr + lam_i * xi
        v = A @ xi - lam_r * xi - lam_i * xr
        nr = math.hypot(float(np.linalg.norm(u)), float(np.linalg.norm(v)))
        nx = math.hypot(float(np.linalg.norm(xr)), float(np.linalg.norm(xi)))
    return nr / (math.hypot(lam_r, lam_i) * nx)


# --------------------------------------------------------------------------------------------
# SVD thick-restart Lanczos (trlanczos.c:203-551, gklanczos.c:58-216)
# --------------------------------------------------------------------------------------------
def svd_default_dims(N, nsv, ncv=None, mpd=None):
    """SVDSetDimensions_Default, svdsetup.c:454-478."""
    if ncv is not None:
        if ncv < nsv:
            raise ValueError("The value of ncv must be at least nsv")
    elif mpd is not None:
        ncv = min(N, nsv + mpd)
    else:
        if nsv < 500:
            ncv = min(N, max(2 * nsv, 10))
        else:
            mpd = 500
            ncv = min(N, nsv + mpd)
    if mpd is None:
        mpd = ncv
    return ncv, mpd


class SVDResult:
    pass


def _svd_orthogonalize_cgs(V, i, h, a, norm_fn):
    """SVDOrthogonalizeCGS, trlanczos.c:319-355: the post-processing after the first (merged) CGS pass of the
    one-sided recurrence.  h[0:i] are the already scaled coefficients, h[i] = v_i^T v_i before the update."""
    refine, eta = V.orthog_ref, V.orthog_eta
    if refine == BV.REFINE_NEVER:
        return norm_fn(i)
    if refine == BV.REFINE_ALWAYS:
        V.set_active(0, i)
        c = V.dotvec(V.col(i))                                # BVDotColumn
        h[0:i] = c
        V.multvec(-1.0, 1.0, V.col(i), c)                     # BVMultColumn
        return norm_fn(i)
    dot = h[i]
    onorm = math.sqrt(dot) / a
    s = float(np.sum(h[0:i] ** 2))
    nrm = dot / (a * a) - s
    nrm = math.sqrt(nrm) if nrm > 0.0 else norm_fn(i)
    if nrm < eta * onorm:
        V.set_active(0, i)
        c = V.dotvec(V.col(i))
        h[0:i] = c
        V.multvec(-1.0, 1.0, V.col(i), c)
        nrm = norm_fn(i)
    return nrm


def _svd_oneside_cgs(A, AT, alpha, beta, V, U, nconv, l, n, work):
    """SVDOneSideTRLanczosCGS, trlanczos.c:357-448.  Returns the number of mat-vecs."""
    k = nconv + l
    nmv = 1
    U.col(k)[:] = A @ V.col(k)
    if l > 0:
        U.set_active(nconv, k)                                # BVMultColumn(U,-1,1,k,work) uses columns nconv..k-1
        work[0:l] = beta[nconv:nconv + l]
        U.multvec(-1.0, 1.0, U.col(k), work[0:l])
        U.set_active(nconv, n)
    refine = V.orthog_ref

    def step(i):
        """norm of u_{i-1} and the merged first CGS pass of v_i (the Begin/End split reductions of :385-400)"""
        a = U.norm_column(i - 1)
        if refine == BV.REFINE_IFNEEDED:
            V.set_active(0, i + 1)
            work[0:i + 1] = V.dotvec(V.col(i))                # includes v_i^T v_i at work[i]
            V.set_active(0, i)
        else:
            V.set_active(0, i)
            work[0:i] = V.dotvec(V.col(i))
        U.scale_column(i - 1, 1.0 / a)
        work[0:i] /= a
        V.multvec(-1.0, 1.0 / a, V.col(i), work[0:i])
        b = _svd_orthogonalize_cgs(V, i, work, a, V.norm_column)
        return a, b

    for i in range(k + 1, n):
        V.col(i)[:] = AT @ U.col(i - 1)
        a, b = step(i)
        V.scale_column(i, 1.0 / b)
        if not abs(b) > 10 * EPS_MACH:
            raise FloatingPointError("Recurrence generated a zero vector; use a two-sided variant")
        U.col(i)[:] = A @ V.col(i) - b * U.col(i - 1)
        nmv += 2
        alpha[i - 1] = a
        beta[i - 1] = b
    V.col(n)[:] = AT @ U.col(n - 1)
    nmv += 1
    a, b = step(n)
    V.set_active(nconv, n)
    alpha[n - 1] = a
    beta[n - 1] = b
    return nmv


def _svd_oneside_mgs(A, AT, alpha, beta, V, U, nconv, l, n, work):
    """SVDOneSideTRLanczosMGS, trlanczos.c:264-314."""
    k = nconv + l
    nmv = 1
    U.col(k)[:] = A @ V.col(k)
    if l > 0:
        U.set_active(nconv, k)
        U.multvec(-1.0, 1.0, U.col(k), beta[nconv:nconv + l].copy())
        U.set_active(nconv, n)
    a = U.norm_column(k)
    U.scale_column(k, 1.0 / a)
    alpha[k] = a
    for i in range(k + 1, n):
        V.col(i)[:] = AT @ U.col(i - 1)
        b, _ = V.orthonormalize_column(i, False)
        beta[i - 1] = b
        U.col(i)[:] = A @ V.col(i) - b * U.col(i - 1)
        nmv += 2
        a = U.norm_column(i)
        U.scale_column(i, 1.0 / a)
        alpha[i] = a
    V.col(n)[:] = AT @ U.col(n - 1)
    nmv += 1
    _, b, _ = V.orthogonalize_column(n)
    beta[n - 1] = b
    return nmv


def svd_trlanczos(A, AT, M, N, nsv, ncv=None, mpd=None, tol=1e-8, max_it=None, keep=0.5, lock=True,
                  v0=None, seed=0x5EED, orthog=None, oneside=False):
    """SVDSolve with -svd_type trlanczos (two-sided, or one-sided with `oneside`), largest singular values.
    Requires M>=N (svdsetup.c:301-306 swaps A/AT otherwise; callers do the swap)."""
    assert M >= N
    ncv, mpd = svd_default_dims(N, nsv, ncv, mpd)
    if max_it is None:
        max_it = max(N // ncv, 100)                          # trlanczos.c:213
    ld = ncv + 1
    U = BV(M, ncv + 1)
    V = BV(N, ncv + 1)
    V.rng_seed = seed
    U.rng_seed = seed + 1000
    if orthog:
        for B in (U, V):
            B.orthog_type, B.orthog_ref, B.orthog_eta = orthog
    ds = DS("svd", ld, cmp_largest_real)
    sigma = np.zeros(ncv + 1)
    errest = np.zeros(ncv + 1)
    w = np.zeros(ld)
    swork = np.zeros(ncv + 1)
    nconv = 0
    its = 0
    reason = 0
    nmatvec = 0
    # start vector trlanczos.c:468-472
    if v0 is not None:
        V.col(0)[:] = v0
    else:
        V.set_random_column(0)
    V.orthonormalize_column(0, True)
    l = 0
    while reason == 0:
        its += 1
        nv = min(nconv + mpd, ncv)
        alpha = ds.T[:, 0]
        beta = ds.T[:, 1]
        if oneside:                                          # trlanczos.c:480-483
            n_ = nv
            breakdown = False
            fn = _svd_oneside_mgs if V.orthog_type == BV.MGS else _svd_oneside_cgs
            nmatvec += fn(A, AT, alpha, beta, V, U, nconv, l, nv, swork)
        else:
            # SVDTwoSideLanczos gklanczos.c:58-113
            k0 = nconv + l
            n_ = nv
            breakdown = False
            U.col(k0)[:] = A @ V.col(k0)
            nmatvec += 1
            alpha[k0], lindep = U.orthonormalize_column(k0, False)
            if lindep:
                n_ = k0
                breakdown = True
            else:
                for i in range(k0 + 1, nv):
                    V.col(i)[:] = AT @ U.col(i - 1)
                    nmatvec += 1
                    beta[i - 1], lindep = V.orthonormalize_column(i, False)
                    if lindep:
                        n_ = i
                        break
                    U.col(i)[:] = A @ V.col(i)
                    nmatvec += 1
                    alpha[i], lindep = U.orthonormalize_column(i, False)
                    if lindep:
                        n_ = i
                        break
                if not lindep:
                    V.col(n_)[:] = AT @ U.col(n_ - 1)
                    nmatvec += 1
                    _, beta[n_ - 1], lindep = V.orthogonalize_column(n_)
                breakdown = lindep
        nv = n_
        V.scale_column(nv, 1.0 / beta[nv - 1])               # trlanczos.c:487
        V.set_active(nconv, nv)
        U.set_active(nconv, nv)

        ds.set_dimensions(nv, nconv, nconv + l)
        ds.m = nv                                            # DSSVDSetDimensions
        ds.set_state(RAW if l else INTERMEDIATE)
        ds.solve(w)
        ds.sort(w)
        ds.update_extra_row()
        sigma[nconv:nv] = w[nconv:nv]

        # SVDKrylovConvergence gklanczos.c:189-216 with SVDConvergedRelative svddefault.c:29-37
        marker = -1
        k = nconv
        while k < nv:
            resnorm = abs(beta[k])
            errest[k] = resnorm / sigma[k] if sigma[k] != 0.0 else np.finfo(np.float64).max  # svddefault.c:29-34
            if marker == -1 and errest[k] >= tol:
                marker = k
            if marker != -1:
                break
            k += 1
        if marker != -1:
            k = marker
        if k >= nsv:
            reason = 2
        elif its >= max_it:
            reason = -3
        if reason != 0 or breakdown or k == nv:
            l = 0
        else:
            l = max(1, int((nv - k) * keep))
        if not lock and l > 0:
            l += k
            k = 0
        if reason == 0:
            if breakdown or k == nv:
                if k < nsv:
                    V.set_random_column(k)
                    _, lindep = V.orthonormalize_column(k, False)
                    if lindep:
                        reason = -1
            else:
                ds.truncate(k + l, False)
        V.mult_in_place(ds.Vm, nconv, k + l)
        U.mult_in_place(ds.Q, nconv, k + l)
        if reason == 0 and not breakdown:
            V.copy_column(nv, k + l)
        nconv = k
    if oneside:                                              # trlanczos.c:540-542
        for i in range(nconv):
            U.orthonormalize_column(i, False)
    ds.truncate(nconv, True)
    res = SVDResult()
    res.nconv, res.its, res.reason, res.nmatvec = nconv, its, reason, nmatvec
    res.ncv, res.mpd, res.max_it = ncv, mpd, max_it
    res.sigma = sigma[:nconv].copy()
    res.errest = errest[:nconv].copy()
    res.U = U.V[:, :nconv].copy()
    res.V = V.V[:, :nconv].copy()
    return res


def svd_relative_error(A, AT, sigma, u, v):
    """SVDComputeResidualNorms_Standard (svdsolve.c:349-373) + SVDComputeError relative
    (svdsolve.c:486): hypot(||Av - sigma u||, ||A^T u - sigma v||) / sigma."""
    n1 = float(np.linalg.norm(A @ v - sigma * u))
    n2 = float(np.linalg.norm(AT @ u - sigma * v))
    return math.hypot(n1, n2) / sigma


# --------------------------------------------------------------------------------------------
# matrix generators (scipy CSR) — behaviour of the reference examples
# --------------------------------------------------------------------------------------------
def laplacian_1d(n):
    """src/eps/tutorials/ex1.c:37-48: tridiag(-1,2,-1)."""
    import scipy.sparse as sp
    return sp.diags([-np.ones(n - 1), 2 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1], format="csr")


def laplacian_2d(nx, ny=None):
    """src/eps/tutorials/ex2.c:39-54: 5-point stencil, II = i*ny + j, i in [0,nx)."""
    import scipy.sparse as sp
    ny = nx if ny is None else ny
    Ix = sp.identity(nx, format="csr")
    Iy = sp.identity(ny, format="csr")
    Tx = sp.diags([-np.ones(nx - 1), 2 * np.ones(nx), -np.ones(nx - 1)], [-1, 0, 1], format="csr")
    Ty = sp.diags([-np.ones(ny - 1), 2 * np.ones(ny), -np.ones(ny - 1)], [-1, 0, 1], format="csr")
    A = (sp.kron(Tx, Iy) + sp.kron(Ix, Ty)).tocsr()
    A.sort_indices()
    return A


def laplacian_3d(nx, ny=None, nz=None):
    """7-point stencil (6,-1) on nx*ny*nz, natural ordering (slowest index first) — SURVEY §8(d) C3."""
    import scipy.sparse as sp
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz

    def T(k):
        return sp.diags([-np.ones(k - 1), 2 * np.ones(k), -np.ones(k - 1)], [-1, 0, 1], format="csr")

    def I(k):
        return sp.identity(k, format="csr")

    A = (sp.kron(sp.kron(T(nx), I(ny)), I(nz)) + sp.kron(sp.kron(I(nx), T(ny)), I(nz))
         + sp.kron(sp.kron(I(nx), I(ny)), T(nz))).tocsr()
    A.sort_indices()
    return A


def markov_model(m):
    """MatMarkovModel, src/eps/tutorials/ex5.c:138-170 (N = m(m+1)/2)."""
    import scipy.sparse as sp
    N = m * (m + 1) // 2
    cst = 0.5 / (m - 1)
    rows, cols, vals = [], [], []
    ix = 0
    for i in range(1, m + 1):
        jmax = m - i + 1
        for j in range(1, jmax + 1):
            ix += 1
            if j != jmax:
                pd = cst * (i + j - 1)
                rows.append(ix - 1); cols.append(ix); vals.append(2 * pd if i == 1 else pd)
                rows.append(ix - 1); cols.append(ix + jmax - 1); vals.append(2 * pd if j == 1 else pd)
            pu = 0.5 - cst * (i + j - 3)
            if j > 1:
                rows.append(ix - 1); cols.append(ix - 2); vals.append(pu)
            if i > 1:
                rows.append(ix - 1); cols.append(ix - jmax - 2); vals.append(pu)
    A = sp.csr_matrix((vals, (rows, cols)), shape=(N, N))
    A.sort_indices()
    return A


def grcar_rect(M, N):
    """src/svd/tests/test3.c: rectangular Grcar-like matrix: A(i,i-1)=-1, A(i,i..i+3)=1."""
    import scipy.sparse as sp
    rows, cols, vals = [], [], []
    for i in range(M):
        if 0 < i <= N:
            rows.append(i); cols.append(i - 1); vals.append(-1.0)
        for j in range(4):
            if i + j < N:
                rows.append(i); cols.append(i + j); vals.append(1.0)
    A = sp.csr_matrix((vals, (rows, cols)), shape=(M, N))
    A.sort_indices()
    return A
