/*
 * ds.c — the small dense projected problem (order <= ncv+1) of the Krylov solvers, solved on the
 * host with LAPACK and replicated on every rank (DS_PARALLEL_REDUNDANT, dsbasic.c:126), exactly
 * where the north star keeps it.  Three types, as used by Krylov-Schur and TRLanczos:
 *   "hep"  compact arrow/tridiagonal + extra row   src/sys/classes/ds/impls/hep/dshep.c
 *   "nhep" dense Hessenberg/Schur + extra row      src/sys/classes/ds/impls/nhep/dsnhep.c, impls/dsutil.c
 *   "svd"  compact arrow/bidiagonal + extra column src/sys/classes/ds/impls/svd/dssvd.c
 * Front-end semantics follow src/sys/classes/ds/interface/{dsbasic,dsops,dspriv}.c.
 */
#include "b2kimpl.h"

#define DSCheckAlloc(ds) PetscCheck((ds)->allocated, PETSC_ERR_ORDER, "Must call DSAllocate() first")
#define IS(ds, t) (!strcmp((ds)->type, t))

PetscErrorCode DSCreate(DS *out)
{
  DS ds = (DS)calloc(1, sizeof(*ds));
  PetscCheck(ds, PETSC_ERR_MEM, "out of memory");
  ds->state = DS_STATE_RAW;
  ds->cmp = SlepcCompareLargestMagnitude;
  *out = ds;
  return PETSC_SUCCESS;
}

PetscErrorCode DSDestroy(DS *pds)
{
  if (!pds || !*pds) return PETSC_SUCCESS;
  DS ds = *pds;
  for (int i = 0; i < DS_NUM_MAT; i++) { free(ds->mat[i]); if (ds->lent[i]) { ds->lent[i]->dense = NULL; MatDestroy(&ds->lent[i]); } }
  free(ds->perm); free(ds->work); free(ds->iwork);
  free(ds);
  *pds = NULL;
  return PETSC_SUCCESS;
}

PetscErrorCode DSSetType(DS ds, const char *type)
{
  PetscCheck(!strcmp(type, DSHEP) || !strcmp(type, DSNHEP) || !strcmp(type, DSSVD), PETSC_ERR_ARG_UNKNOWN_TYPE, "Unable to find requested DS type %s (hep, nhep and svd serve the Krylov path)", type);
  strcpy(ds->type, type);
  return PETSC_SUCCESS;
}
PetscErrorCode DSSetCompact(DS ds, PetscBool comp) { ds->compact = comp; return PETSC_SUCCESS; }
PetscErrorCode DSSetExtraRow(DS ds, PetscBool ext)
{
  PetscCheck(!(ds->n > 0 && ds->n == ds->ld), PETSC_ERR_ORDER, "Cannot set extra row after setting n=ld");
  ds->extrarow = ext;
  return PETSC_SUCCESS;
}

PetscErrorCode DSAllocate(DS ds, PetscInt ld)
{
  PetscCheck(ds->type[0], PETSC_ERR_ORDER, "Must call DSSetType() first");
  PetscCheck(ld > 0, PETSC_ERR_ARG_OUTOFRANGE, "Leading dimension should be at least one");
  for (int i = 0; i < DS_NUM_MAT; i++) { free(ds->mat[i]); ds->mat[i] = NULL; }
  free(ds->perm); free(ds->work); free(ds->iwork);
  ds->ld = ld;
  const size_t sq = (size_t)ld * (size_t)ld;
  const int need[] = {DS_MAT_A, DS_MAT_T, DS_MAT_Q, DS_MAT_X, DS_MAT_U, DS_MAT_V, DS_MAT_W};
  for (size_t i = 0; i < sizeof(need) / sizeof(need[0]); i++) {
    ds->mat[need[i]] = (PetscScalar *)calloc(need[i] == DS_MAT_T ? 3 * (size_t)ld : sq, sizeof(PetscScalar));
    PetscCheck(ds->mat[need[i]], PETSC_ERR_MEM, "out of memory");
  }
  ds->perm = (PetscInt *)calloc((size_t)ld, sizeof(PetscInt));
  ds->work = (PetscScalar *)calloc(sq + 16 * (size_t)ld + 64, sizeof(PetscScalar));
  ds->iwork = (PetscInt *)calloc(2 * (size_t)ld + 8, sizeof(PetscInt));
  PetscCheck(ds->perm && ds->work && ds->iwork, PETSC_ERR_MEM, "out of memory");
  ds->allocated = PETSC_TRUE;
  ds->n = ds->l = ds->k = ds->t = ds->m = ds->tm = 0;
  ds->state = DS_STATE_RAW;
  return PETSC_SUCCESS;
}
PetscErrorCode DSGetLeadingDimension(DS ds, PetscInt *ld) { *ld = ds->ld; return PETSC_SUCCESS; }

/* dsops.c:129-165 */
PetscErrorCode DSSetDimensions(DS ds, PetscInt n, PetscInt l, PetscInt k)
{
  DSCheckAlloc(ds);
  const PetscInt on = ds->n, ol = ds->l, ok = ds->k;
  if (n == PETSC_DECIDE || n == PETSC_DEFAULT) ds->n = ds->extrarow ? ds->ld - 1 : ds->ld;
  else {
    PetscCheck(n >= 0 && n <= ds->ld, PETSC_ERR_ARG_OUTOFRANGE, "Illegal value of n. Must be between 0 and ld");
    PetscCheck(!ds->extrarow || n < ds->ld, PETSC_ERR_ARG_OUTOFRANGE, "Illegal value of n. Must be less than ld if extra row is used");
    ds->n = n;
  }
  ds->t = ds->n;
  if (l == PETSC_DECIDE || l == PETSC_DEFAULT) ds->l = 0;
  else { PetscCheck(l >= 0 && l <= ds->n, PETSC_ERR_ARG_OUTOFRANGE, "Illegal value of l. Must be between 0 and n"); ds->l = l; }
  if (k == PETSC_DECIDE || k == PETSC_DEFAULT) ds->k = ds->n / 2;
  else { PetscCheck(k >= 0 || k <= ds->n, PETSC_ERR_ARG_OUTOFRANGE, "Illegal value of k. Must be between 0 and n"); ds->k = k; }
  (void)on; (void)ol; (void)ok;
  return PETSC_SUCCESS;
}
PetscErrorCode DSGetDimensions(DS ds, PetscInt *n, PetscInt *l, PetscInt *k, PetscInt *t)
{
  if (n) *n = ds->n;
  if (l) *l = ds->l;
  if (k) *k = ds->k;
  if (t) *t = ds->t;
  return PETSC_SUCCESS;
}
/* dssvd.c:636-668 */
PetscErrorCode DSSVDSetDimensions(DS ds, PetscInt m)
{
  DSCheckAlloc(ds);
  PetscCheck(IS(ds, DSSVD), PETSC_ERR_ARG_WRONG, "DSSVDSetDimensions() needs a DS of type svd");
  if (m == PETSC_DECIDE || m == PETSC_DEFAULT) ds->m = ds->ld;
  else { PetscCheck(m > 0 && m <= ds->ld, PETSC_ERR_ARG_OUTOFRANGE, "Illegal value of m. Must be between 1 and ld"); ds->m = m; }
  return PETSC_SUCCESS;
}
PetscErrorCode DSSetState(DS ds, DSStateType state)
{
  PetscCheck(state >= DS_STATE_RAW && state <= DS_STATE_TRUNCATED, PETSC_ERR_ARG_WRONG, "Wrong state");
  ds->state = state;
  return PETSC_SUCCESS;
}
PetscErrorCode DSGetState(DS ds, DSStateType *state) { *state = ds->state; return PETSC_SUCCESS; }
PetscErrorCode DSSetEigenvalueComparison(DS ds, SlepcEigenvalueComparisonFn fn, void *ctx) { ds->cmp = fn; ds->cmpctx = ctx; return PETSC_SUCCESS; }

PetscErrorCode DSGetArray(DS ds, DSMatType m, PetscScalar **a)
{
  DSCheckAlloc(ds);
  PetscCheck(m >= 0 && m < DS_NUM_MAT && ds->mat[m], PETSC_ERR_ARG_WRONGSTATE, "Requested matrix was not created in this DS");
  *a = ds->mat[m];
  return PETSC_SUCCESS;
}
PetscErrorCode DSRestoreArray(DS ds, DSMatType m, PetscScalar **a) { (void)ds; (void)m; *a = NULL; return PETSC_SUCCESS; }
PetscErrorCode DSGetArrayReal(DS ds, DSMatType m, PetscReal **a) { return DSGetArray(ds, m, a); }
PetscErrorCode DSRestoreArrayReal(DS ds, DSMatType m, PetscReal **a) { (void)ds; (void)m; *a = NULL; return PETSC_SUCCESS; }

/* DSMatGetSize dsops.c:276-295, DSMatGetSize_SVD dssvd.c:607-634 */
static PetscErrorCode DSMatGetSize_Private(DS ds, DSMatType t, PetscInt *rows, PetscInt *cols)
{
  if (IS(ds, DSSVD)) {
    PetscCheck(ds->m, PETSC_ERR_ORDER, "You should set the number of columns with DSSVDSetDimensions()");
    switch (t) {
    case DS_MAT_A: *rows = ds->n; *cols = ds->extrarow ? ds->m + 1 : ds->m; break;
    case DS_MAT_T: *rows = ds->n; *cols = 3; break;
    case DS_MAT_U: *rows = ds->state == DS_STATE_TRUNCATED ? ds->t : ds->n; *cols = ds->n; break;
    case DS_MAT_V: *rows = ds->state == DS_STATE_TRUNCATED ? ds->tm : ds->m; *cols = ds->m; break;
    default: SETERRQ(PETSC_ERR_ARG_OUTOFRANGE, "Invalid t parameter");
    }
  } else {
    if (ds->state == DS_STATE_TRUNCATED && t >= DS_MAT_Q) *rows = ds->t;
    else *rows = (t == DS_MAT_A && ds->extrarow) ? ds->n + 1 : ds->n;
    *cols = (t == DS_MAT_T) ? 3 : ds->n;
  }
  return PETSC_SUCCESS;
}

/* dsops.c:417-470: a rows x cols view with LDA = ld of the DS storage */
PetscErrorCode DSGetMat(DS ds, DSMatType m, Mat *A)
{
  DSCheckAlloc(ds);
  PetscCheck(m >= 0 && m < DS_NUM_MAT && ds->mat[m], PETSC_ERR_ARG_WRONGSTATE, "Requested matrix was not created in this DS");
  PetscCheck(!ds->lent[m], PETSC_ERR_ARG_WRONGSTATE, "DSRestoreMat() must be called before a second DSGetMat() of the same matrix");
  PetscInt rows = 0, cols = 0;
  PetscCall(DSMatGetSize_Private(ds, m, &rows, &cols));
  PetscCheck(rows && cols, PETSC_ERR_ORDER, "Must call DSSetDimensions() first");
  PetscCall(MatCreateSeqDense(rows, cols, ds->mat[m], A));
  (*A)->lda = ds->ld;
  ds->lent[m] = *A;
  return PETSC_SUCCESS;
}
PetscErrorCode DSRestoreMat(DS ds, DSMatType m, Mat *A)
{
  DSCheckAlloc(ds);
  PetscCheck(A && *A && ds->lent[m] == *A, PETSC_ERR_ARG_WRONG, "The Mat was not obtained with DSGetMat()");
  ds->lent[m] = NULL;
  PetscCall(MatDestroy(A));
  return PETSC_SUCCESS;
}

/* ---- sorting helpers, dspriv.c:172-333 ----------------------------------------------------------------- */
static PetscErrorCode DSSortEigenvaluesReal_Private(DS ds, const PetscReal *eig, PetscInt *perm, PetscInt l, PetscInt n)
{
  PetscInt result, tmp;
  for (PetscInt i = l + 1; i < n; i++) {
    const PetscScalar re = eig[perm[i]];
    PetscInt j = i - 1;
    PetscCall(ds->cmp(re, 0.0, eig[perm[j]], 0.0, &result, ds->cmpctx));
    while (result < 0 && j >= l) {
      tmp = perm[j]; perm[j] = perm[j + 1]; perm[j + 1] = tmp; j--;
      if (j >= l) PetscCall(ds->cmp(re, 0.0, eig[perm[j]], 0.0, &result, ds->cmpctx));
    }
  }
  return PETSC_SUCCESS;
}

static void swap_cols(PetscScalar *M, PetscInt ld, PetscInt n, PetscInt a, PetscInt b)
{
  for (PetscInt r = 0; r < n; r++) { const PetscScalar t = M[r + (size_t)a * ld]; M[r + (size_t)a * ld] = M[r + (size_t)b * ld]; M[r + (size_t)b * ld] = t; }
}

/* DSPermuteColumns_Private / DSPermuteBoth_Private (M2 may be NULL); destroys perm */
static void DSPermute_Private(DS ds, PetscInt istart, PetscInt iend, PetscScalar *M1, PetscInt n1, PetscScalar *M2, PetscInt n2, PetscInt *perm)
{
  for (PetscInt i = istart; i < iend; i++) {
    const PetscInt p = perm[i];
    if (p != i) {
      PetscInt j = i + 1;
      while (perm[j] != i) j++;
      perm[j] = p; perm[i] = i;
      swap_cols(M1, ds->ld, n1, i, p);
      if (M2) swap_cols(M2, ds->ld, n2, i, p);
    }
  }
}

/* ---- HEP: dshep.c ----------------------------------------------------------------------------------------- */
static void rot_cols(PetscScalar *Q, PetscInt ld, PetscInt rows, PetscInt a, PetscInt b, PetscReal c, PetscReal s)
{
  for (PetscInt r = 0; r < rows; r++) {
    const PetscScalar x = Q[r + (size_t)a * ld], y = Q[r + (size_t)b * ld];
    Q[r + (size_t)a * ld] = c * x + s * y;
    Q[r + (size_t)b * ld] = c * y - s * x;
  }
}

/* arrowhead → tridiagonal by plane rotations with bulge chasing (Rutishauser), dshep.c:221-262 */
static void DSArrowTridiag(PetscInt n, PetscReal *d, PetscReal *e, PetscScalar *Q, PetscInt ld)
{
  PetscReal c, s, r, temp, off, p;
  if (n <= 2) return;
  for (PetscInt j = 0; j < n - 2; j++) {
    temp = e[j + 1];
    LAPACK(dlartg)(&temp, &e[j], &c, &s, &r);
    e[j + 1] = r; s = -s;
    temp = d[j + 1];
    e[j] = c * s * (temp - d[j]);
    d[j + 1] = s * s * d[j] + c * c * temp;
    d[j] = c * c * d[j] + s * s * temp;
    rot_cols(Q, ld, j + 2, j, j + 1, c, s);
    for (PetscInt i = j - 1; i >= 0; i--) {
      off = -s * e[i];
      e[i] = c * e[i];
      temp = e[i + 1];
      LAPACK(dlartg)(&temp, &off, &c, &s, &r);
      e[i + 1] = r; s = -s;
      temp = (d[i] - d[i + 1]) * s - 2.0 * c * e[i];
      p = s * temp;
      d[i + 1] += p;
      d[i] -= p;
      e[i] = -e[i] - c * temp;
      rot_cols(Q, ld, j + 2, i, i + 1, c, s);
    }
  }
}

/* DSSolve_HEP_QR dshep.c:383-426 with the compact branch of DSIntermediate_HEP (:267-290) */
static PetscErrorCode DSSolve_HEP(DS ds, PetscScalar *wr, PetscScalar *wi)
{
  PetscCheck(ds->compact, PETSC_ERR_SUP, "DS type hep is implemented in compact storage only (what Krylov-Schur uses, krylovschur.c:160-167)");
  const PetscInt n = ds->n, l = ds->l, ld = ds->ld;
  PetscReal *d = ds->mat[DS_MAT_T], *e = d + ld;
  PetscScalar *Q = ds->mat[DS_MAT_Q];
  const int n1 = PetscMax(0, ds->k - l + 1), n3 = n - l;
  int info = 0, ldi = ld;
  for (PetscInt j = l; j < n; j++) { memset(Q + (size_t)j * ld, 0, sizeof(PetscScalar) * (size_t)ld); Q[j + (size_t)j * ld] = 1.0; }   /* DSSetIdentity */
  if (ds->state < DS_STATE_INTERMEDIATE) DSArrowTridiag(n1, d + l, e + l, Q + l + (size_t)l * ld, ld);
  for (PetscInt i = 0; i < l; i++) wr[i] = d[i];
  if (n3 > 0) {
    LAPACK(dsteqr)("V", &n3, d + l, e + l, Q + l + (size_t)l * ld, &ldi, ds->work, &info, 1);
    PetscCheck(!info, PETSC_ERR_LIB, "Error in LAPACK subroutine steqr: info=%d", info);
  }
  for (PetscInt i = l; i < n; i++) wr[i] = d[i];
  for (PetscInt i = 0; i < n - 1; i++) e[i] = 0.0;
  if (wi) for (PetscInt i = l; i < n; i++) wi[i] = 0.0;
  return PETSC_SUCCESS;
}

/* DSSort_HEP dshep.c:323-347 (rr == NULL) */
static PetscErrorCode DSSort_HEP(DS ds, PetscScalar *wr)
{
  const PetscInt n = ds->n, l = ds->l;
  PetscReal *d = ds->mat[DS_MAT_T];
  PetscCall(DSSortEigenvaluesReal_Private(ds, d, ds->perm, l, ds->t));
  for (PetscInt i = l; i < n; i++) wr[i] = d[ds->perm[i]];
  DSPermute_Private(ds, l, n, ds->mat[DS_MAT_Q], n, NULL, 0, ds->perm);
  for (PetscInt i = l; i < n; i++) d[i] = wr[i];
  return PETSC_SUCCESS;
}

/* DSUpdateExtraRow_HEP dshep.c:349-381 (compact) */
static PetscErrorCode DSUpdateExtraRow_HEP(DS ds)
{
  const PetscInt n = ds->n, ld = ds->ld;
  PetscReal *e = ds->mat[DS_MAT_T] + ld;
  const PetscScalar *Q = ds->mat[DS_MAT_Q];
  const PetscReal beta = e[n - 1];                 /* all entries assumed zero except the last one */
  for (PetscInt i = 0; i < n; i++) e[i] = beta * Q[n - 1 + (size_t)i * ld];
  ds->k = n;
  return PETSC_SUCCESS;
}

/* ---- NHEP: dsutil.c, dsnhep.c --------------------------------------------------------------------------------- */
static PetscErrorCode DSSolve_NHEP(DS ds, PetscScalar *wr, PetscScalar *wi)
{
  PetscCheck(wi, PETSC_ERR_ARG_NULL, "wi is required in real arithmetic");
  const PetscInt n = ds->n, l = ds->l, ld = ds->ld;
  PetscScalar *A = ds->mat[DS_MAT_A], *Q = ds->mat[DS_MAT_Q];
  int ni = n, ilo = l + 1, ldi = ld, info = 0, lwork = ld * ld + 16 * ld;
  PetscScalar *tau = ds->work, *work = ds->work + ld;
  lwork -= ld;
  memset(Q, 0, sizeof(PetscScalar) * (size_t)ld * (size_t)ld);
  for (PetscInt i = 0; i < n; i++) Q[i + (size_t)i * ld] = 1.0;
  if (n == 1) { wr[0] = A[0]; wi[0] = 0.0; return PETSC_SUCCESS; }   /* dsutil.c:44-49 */
  if (ds->state < DS_STATE_INTERMEDIATE) {         /* reduce to upper Hessenberg form */
    LAPACK(dgehrd)(&ni, &ilo, &ni, A, &ldi, tau, work, &lwork, &info);
    PetscCheck(!info, PETSC_ERR_LIB, "Error in LAPACK subroutine gehrd: info=%d", info);
    for (PetscInt j = 0; j < n - 1; j++)
      for (PetscInt i = j + 2; i < n; i++) { Q[i + (size_t)j * ld] = A[i + (size_t)j * ld]; A[i + (size_t)j * ld] = 0.0; }
    LAPACK(dorghr)(&ni, &ilo, &ni, Q, &ldi, tau, work, &lwork, &info);
    PetscCheck(!info, PETSC_ERR_LIB, "Error in LAPACK subroutine orghr: info=%d", info);
  }
  LAPACK(dhseqr)("S", "V", &ni, &ilo, &ni, A, &ldi, wr, wi, Q, &ldi, work, &lwork, &info, 1, 1);
  PetscCheck(!info, PETSC_ERR_LIB, "Error in LAPACK subroutine hseqr: info=%d", info);
  for (PetscInt j = 0; j < l; j++) {               /* eigenvalues of the locked part, dsutil.c:75-88 */
    if (j == n - 1 || A[j + 1 + (size_t)j * ld] == 0.0) { wr[j] = A[j + (size_t)j * ld]; wi[j] = 0.0; }
    else {
      wr[j] = A[j + (size_t)j * ld]; wr[j + 1] = A[j + (size_t)j * ld];
      wi[j] = sqrt(fabs(A[j + 1 + (size_t)j * ld])) * sqrt(fabs(A[j + (size_t)(j + 1) * ld]));
      wi[j + 1] = -wi[j];
      j++;
    }
  }
  return PETSC_SUCCESS;
}

/* DSSort_NHEP_Total dsutil.c:93-171: selection sort, moving blocks with trexc */
static PetscErrorCode DSSort_NHEP(DS ds, PetscScalar *wr, PetscScalar *wi)
{
  const PetscInt n = ds->n, ld = ds->ld;
  PetscScalar *T = ds->mat[DS_MAT_A], *Q = ds->mat[DS_MAT_Q];
  int ni = n, ldi = ld, info = 0, ifst, ilst;
  PetscInt result, pos;
  for (PetscInt i = ds->l; i < n - 1; i++) {
    PetscScalar re = wr[i], im = wi[i];
    pos = 0;
    PetscInt j = i + 1;
    if (im != 0) j = i + 2;                        /* complex eigenvalue: skip its conjugate */
    for (; j < n; j++) {
      PetscCall(ds->cmp(re, im, wr[j], wi[j], &result, ds->cmpctx));
      if (result > 0) { re = wr[j]; im = wi[j]; pos = j; }
      if (wi[j] != 0) j++;
    }
    if (pos) {
      ifst = pos + 1; ilst = i + 1;
      LAPACK(dtrexc)("V", &ni, T, &ldi, Q, &ldi, &ifst, &ilst, ds->work, &info, 1);
      PetscCheck(!info, PETSC_ERR_LIB, "Error in LAPACK subroutine trexc: info=%d", info);
      for (j = i; j < n; j++) {                    /* recover the eigenvalues from T */
        wr[j] = T[j + (size_t)j * ld];
        if (j < n - 1 && T[j + 1 + (size_t)j * ld] != 0.0) {
          wi[j] = sqrt(fabs(T[j + 1 + (size_t)j * ld])) * sqrt(fabs(T[j + (size_t)(j + 1) * ld]));
          wr[j + 1] = wr[j]; wi[j + 1] = -wi[j];
          j++;
        } else wi[j] = 0.0;
      }
    }
    if (wi[i] != 0) i++;
  }
  return PETSC_SUCCESS;
}

/* DSUpdateExtraRow_NHEP dsnhep.c:318-340: row n of A <- Q^T * row */
static PetscErrorCode DSUpdateExtraRow_NHEP(DS ds)
{
  const PetscInt n = ds->n, ld = ds->ld;
  PetscScalar *A = ds->mat[DS_MAT_A], *x = ds->work, *y = ds->work + ld;
  const PetscScalar *Q = ds->mat[DS_MAT_Q];
  const double one = 1.0, zero = 0.0;
  const int ni = n, ldi = ld, inc = 1;
  for (PetscInt i = 0; i < n; i++) x[i] = A[n + (size_t)i * ld];
  LAPACK(dgemv)("C", &ni, &ni, &one, Q, &ldi, x, &inc, &zero, y, &inc, 1);
  for (PetscInt i = 0; i < n; i++) A[n + (size_t)i * ld] = y[i];
  ds->k = n;
  return PETSC_SUCCESS;
}

static PetscReal nrm2(PetscInt n, const PetscScalar *x) { const int ni = n, inc = 1; return LAPACK(dnrm2)(&ni, x, &inc); }

/* DSVectors_NHEP_Eigen_Some dsnhep.c:101-166 (right vectors) */
static PetscErrorCode DSVectors_NHEP_Some(DS ds, PetscInt *k, PetscReal *rnorm)
{
  const PetscInt n = ds->n, ld = ds->ld;
  PetscScalar *A = ds->mat[DS_MAT_A], *X = ds->mat[DS_MAT_X];
  const PetscScalar *Q = ds->mat[DS_MAT_Q];
  int *select = (int *)ds->iwork;
  int ni = n, ldi = ld, mm = 1, mout = 0, info = 0;
  const int inc = 1;
  const double one = 1.0, zero = 0.0;
  PetscScalar *work = ds->work;                    /* 3*ld for trevc, then ld for the back-transform */
  for (PetscInt i = 0; i < n; i++) select[i] = 0;
  const PetscBool iscomplex = (*k < n - 1 && A[*k + 1 + (size_t)(*k) * ld] != 0.0) ? PETSC_TRUE : PETSC_FALSE;
  select[*k] = 1;
  if (iscomplex) { mm = 2; select[*k + 1] = 1; }
  PetscScalar *Y = X + (size_t)(*k) * ld;
  LAPACK(dtrevc)("R", "S", select, &ni, A, &ldi, Y, &ldi, Y, &ldi, &mm, &mout, work, &info, 1, 1);
  PetscCheck(!info, PETSC_ERR_LIB, "Error in LAPACK subroutine trevc: info=%d", info);
  PetscCheck(mout == mm, PETSC_ERR_ARG_WRONG, "Inconsistent arguments");
  if (ds->state >= DS_STATE_CONDENSED) {           /* back-transform with Q and normalise */
    PetscScalar *w = work + 3 * ld;
    for (int c = 0; c < mm; c++) {
      PetscScalar *col = X + (size_t)(*k + c) * ld;
      memcpy(w, col, sizeof(PetscScalar) * (size_t)n);
      LAPACK(dgemv)("N", &ni, &ni, &one, Q, &ldi, w, &inc, &zero, col, &inc, 1);
    }
    PetscReal norm = nrm2(n, X + (size_t)(*k) * ld);
    if (iscomplex) norm = hypot(norm, nrm2(n, X + (size_t)(*k + 1) * ld));
    for (int c = 0; c < mm; c++) for (PetscInt i = 0; i < n; i++) X[i + (size_t)(*k + c) * ld] /= norm;
  }
  if (rnorm) {                                     /* residual estimate: |last component| (times beta in the caller) */
    if (iscomplex) *rnorm = hypot(X[n - 1 + (size_t)(*k) * ld], X[n - 1 + (size_t)(*k + 1) * ld]);
    else *rnorm = fabs(X[n - 1 + (size_t)(*k) * ld]);
  }
  if (iscomplex) (*k)++;
  return PETSC_SUCCESS;
}

/* DSVectors_NHEP_Eigen_All dsnhep.c:168-232 (right vectors), column-normalised */
static PetscErrorCode DSVectors_NHEP_All(DS ds)
{
  const PetscInt n = ds->n, ld = ds->ld;
  PetscScalar *A = ds->mat[DS_MAT_A], *X = ds->mat[DS_MAT_X];
  int ni = n, ldi = ld, mout = 0, info = 0;
  const PetscBool back = ds->state >= DS_STATE_CONDENSED ? PETSC_TRUE : PETSC_FALSE;
  if (back) memcpy(X, ds->mat[DS_MAT_Q], sizeof(PetscScalar) * (size_t)ld * (size_t)ld);
  LAPACK(dtrevc)("R", back ? "B" : "A", NULL, &ni, A, &ldi, X, &ldi, X, &ldi, &ni, &mout, ds->work, &info, 1, 1);
  PetscCheck(!info, PETSC_ERR_LIB, "Error in LAPACK subroutine trevc: info=%d", info);
  for (PetscInt i = 0; i < n; i++) {
    const PetscBool iscomplex = (i < n - 1 && A[i + 1 + (size_t)i * ld] != 0.0) ? PETSC_TRUE : PETSC_FALSE;
    const PetscInt cols = iscomplex ? 2 : 1;
    PetscReal norm = nrm2(n, X + (size_t)i * ld);
    if (iscomplex) norm = hypot(norm, nrm2(n, X + (size_t)(i + 1) * ld));
    for (PetscInt c = 0; c < cols; c++) for (PetscInt r = 0; r < n; r++) X[r + (size_t)(i + c) * ld] /= norm;
    i += cols - 1;
  }
  return PETSC_SUCCESS;
}

/* ---- SVD: dssvd.c --------------------------------------------------------------------------------------------- */
/* arrowhead → upper bidiagonal, dssvd.c:277-339 */
static void DSArrowBidiag(PetscInt n, PetscReal *d, PetscReal *e, PetscScalar *Q, PetscInt ldq, PetscScalar *P, PetscInt ldp)
{
  PetscReal c, s, ct, st, r, off, temp0, temp1, temp2;
  if (n <= 2) return;
  for (PetscInt j = 0; j < n - 2; j++) {
    temp0 = e[j + 1];
    LAPACK(dlartg)(&temp0, &e[j], &c, &s, &r);
    e[j + 1] = r; s = -s;
    rot_cols(Q, ldq, j + 2, j, j + 1, c, s);
    temp0 = d[j + 1];
    temp1 = c * temp0;
    temp2 = -s * d[j];
    LAPACK(dlartg)(&temp1, &temp2, &ct, &st, &r);
    d[j + 1] = r; st = -st;
    e[j] = -c * st * d[j] + s * ct * temp0;
    d[j] = c * ct * d[j] + s * st * temp0;
    rot_cols(P, ldp, j + 2, j, j + 1, ct, st);
    for (PetscInt i = j - 1; i >= 0; i--) {
      off = -st * e[i];
      e[i] = ct * e[i];
      temp0 = e[i + 1];
      LAPACK(dlartg)(&temp0, &off, &c, &s, &r);
      e[i + 1] = r; s = -s;
      rot_cols(Q, ldq, j + 2, i, i + 1, c, s);
      temp0 = d[i + 1];
      temp1 = -s * e[i] + c * temp0;
      temp2 = c * e[i] + s * temp0;
      off = -s * d[i];
      LAPACK(dlartg)(&temp1, &off, &ct, &st, &r);
      d[i + 1] = r; st = -st;
      e[i] = -c * st * d[i] + ct * temp2;
      d[i] = c * ct * d[i] + st * temp2;
      rot_cols(P, ldp, j + 2, i, i + 1, ct, st);
    }
  }
}

/* DSSolve_SVD_QR dssvd.c:400-457 + compact branch of DSIntermediate_SVD (:344-398) */
static PetscErrorCode DSSolve_SVD(DS ds, PetscScalar *wr)
{
  PetscCheck(ds->compact, PETSC_ERR_SUP, "DS type svd is implemented in compact storage only (what TRLanczos uses, trlanczos.c:257-259)");
  PetscCheck(ds->m, PETSC_ERR_ORDER, "You should set the number of columns with DSSVDSetDimensions()");
  const PetscInt n = ds->n, m = ds->m, l = ds->l, ld = ds->ld;
  PetscReal *d = ds->mat[DS_MAT_T], *e = d + ld;
  PetscScalar *U = ds->mat[DS_MAT_U], *V = ds->mat[DS_MAT_V], *Vt = ds->mat[DS_MAT_W];
  const size_t off = (size_t)l + (size_t)l * ld;
  int n1 = n - l, m1 = m - l, nm = PetscMin(n1, m1), ldi = ld, zero = 0, info = 0;
  const PetscInt nlead = PetscMax(0, ds->k - l + 1);
  memset(U, 0, sizeof(PetscScalar) * (size_t)ld * ld);
  memset(V, 0, sizeof(PetscScalar) * (size_t)ld * ld);
  for (PetscInt i = 0; i < n; i++) U[i + (size_t)i * ld] = 1.0;
  for (PetscInt i = 0; i < m; i++) V[i + (size_t)i * ld] = 1.0;
  if (ds->state < DS_STATE_INTERMEDIATE) DSArrowBidiag(nlead, d + l, e + l, U + off, ld, V + off, ld);
  for (PetscInt i = 0; i < l; i++) wr[i] = d[i];
  for (PetscInt i = l; i < m; i++) for (PetscInt j = l; j < m; j++) Vt[i + (size_t)j * ld] = V[j + (size_t)i * ld];   /* LAPACK wants V^T */
  LAPACK(dbdsqr)(n >= m ? "U" : "L", &nm, &m1, &n1, &zero, d + l, e + l, Vt + off, &ldi, U + off, &ldi, NULL, &ldi, ds->work, &info, 1);
  PetscCheck(!info, PETSC_ERR_LIB, "Error in LAPACK subroutine bdsqr: info=%d", info);
  for (PetscInt i = l; i < m; i++) for (PetscInt j = l; j < m; j++) V[i + (size_t)j * ld] = Vt[j + (size_t)i * ld];
  for (PetscInt i = l; i < PetscMin(n, m); i++) wr[i] = d[i];
  for (PetscInt i = 0; i < n - 1; i++) e[i] = 0.0;
  return PETSC_SUCCESS;
}

/* DSSort_SVD dssvd.c:156-184 */
static PetscErrorCode DSSort_SVD(DS ds, PetscScalar *wr)
{
  const PetscInt l = ds->l, n = PetscMin(ds->n, ds->m);
  PetscReal *d = ds->mat[DS_MAT_T];
  PetscCall(DSSortEigenvaluesReal_Private(ds, d, ds->perm, l, n));
  for (PetscInt i = l; i < n; i++) wr[i] = d[ds->perm[i]];
  DSPermute_Private(ds, l, n, ds->mat[DS_MAT_U], ds->n, ds->mat[DS_MAT_V], ds->m, ds->perm);
  for (PetscInt i = l; i < n; i++) d[i] = wr[i];
  return PETSC_SUCCESS;
}

/* DSUpdateExtraRow_SVD dssvd.c:186-221 (compact) */
static PetscErrorCode DSUpdateExtraRow_SVD(DS ds)
{
  const PetscInt n = ds->n, m = ds->m, ld = ds->ld;
  PetscReal *e = ds->mat[DS_MAT_T] + ld;
  const PetscScalar *U = ds->mat[DS_MAT_U];
  const PetscReal beta = e[m - 1];
  for (PetscInt i = 0; i < n; i++) e[i] = beta * U[n - 1 + (size_t)i * ld];
  ds->k = m;
  return PETSC_SUCCESS;
}

/* ---- front-ends: dsops.c ------------------------------------------------------------------------------------------- */
PetscErrorCode DSSolve(DS ds, PetscScalar eigr[], PetscScalar eigi[])
{
  DSCheckAlloc(ds);
  PetscCheck(eigr, PETSC_ERR_ARG_NULL, "null eigr");
  if (ds->state >= DS_STATE_CONDENSED) return PETSC_SUCCESS;
  if (IS(ds, DSHEP)) PetscCall(DSSolve_HEP(ds, eigr, eigi));
  else if (IS(ds, DSNHEP)) PetscCall(DSSolve_NHEP(ds, eigr, eigi));
  else PetscCall(DSSolve_SVD(ds, eigr));
  ds->state = DS_STATE_CONDENSED;
  return PETSC_SUCCESS;
}

PetscErrorCode DSSort(DS ds, PetscScalar eigr[], PetscScalar eigi[], PetscScalar rr[], PetscScalar ri[], PetscInt *k)
{
  DSCheckAlloc(ds);
  PetscCheck(ds->state >= DS_STATE_CONDENSED, PETSC_ERR_ORDER, "Must call DSSolve() first");
  PetscCheck(ds->state != DS_STATE_TRUNCATED, PETSC_ERR_ORDER, "Cannot sort a truncated DS");
  PetscCheck(!rr && !ri && !k, PETSC_ERR_SUP, "sorting by user-provided values is outside the Krylov path");
  PetscCheck(ds->cmp, PETSC_ERR_ORDER, "Must provide a sorting criterion first");
  for (PetscInt i = 0; i < ds->n; i++) ds->perm[i] = i;   /* dsops.c:795 */
  if (IS(ds, DSHEP)) PetscCall(DSSort_HEP(ds, eigr));
  else if (IS(ds, DSNHEP)) PetscCall(DSSort_NHEP(ds, eigr, eigi));
  else PetscCall(DSSort_SVD(ds, eigr));
  return PETSC_SUCCESS;
}

PetscErrorCode DSUpdateExtraRow(DS ds)
{
  DSCheckAlloc(ds);
  PetscCheck(ds->extrarow, PETSC_ERR_ARG_WRONGSTATE, "Should have called DSSetExtraRow");
  if (IS(ds, DSHEP)) PetscCall(DSUpdateExtraRow_HEP(ds));
  else if (IS(ds, DSNHEP)) PetscCall(DSUpdateExtraRow_NHEP(ds));
  else PetscCall(DSUpdateExtraRow_SVD(ds));
  return PETSC_SUCCESS;
}

PetscErrorCode DSSynchronize(DS ds, PetscScalar eigr[], PetscScalar eigi[]) { (void)ds; (void)eigr; (void)eigi; return PETSC_SUCCESS; }

/* DSVectors(ds,DS_MAT_X,&j,&rnorm): dshep.c:137-172 / dsnhep.c:234-258 */
PetscErrorCode DSVectors(DS ds, DSMatType mat, PetscInt *j, PetscReal *rnorm)
{
  DSCheckAlloc(ds);
  PetscCheck(ds->state >= DS_STATE_RAW, PETSC_ERR_ORDER, "bad state");
  PetscCheck(mat == DS_MAT_X, PETSC_ERR_SUP, "only right eigenvectors (DS_MAT_X) are on the Krylov path");
  if (IS(ds, DSHEP)) {
    const PetscInt ld = ds->ld;
    const PetscScalar *Q = ds->mat[DS_MAT_Q];
    PetscCheck(!rnorm || j, PETSC_ERR_ARG_NULL, "rnorm needs a column index");
    if (j) {
      if (ds->state >= DS_STATE_CONDENSED) memcpy(ds->mat[DS_MAT_X] + (size_t)(*j) * ld, Q + (size_t)(*j) * ld, sizeof(PetscScalar) * (size_t)ld);
      else { memset(ds->mat[DS_MAT_X] + (size_t)(*j) * ld, 0, sizeof(PetscScalar) * (size_t)ld); ds->mat[DS_MAT_X][*j + (size_t)(*j) * ld] = 1.0; }
      if (rnorm) *rnorm = fabs(Q[ds->n - 1 + (size_t)(*j) * ld]);
    } else {
      if (ds->state >= DS_STATE_CONDENSED) memcpy(ds->mat[DS_MAT_X], Q, sizeof(PetscScalar) * (size_t)ld * ld);
      else { memset(ds->mat[DS_MAT_X], 0, sizeof(PetscScalar) * (size_t)ld * ld); for (PetscInt i = 0; i < ds->n; i++) ds->mat[DS_MAT_X][i + (size_t)i * ld] = 1.0; }
    }
  } else if (IS(ds, DSNHEP)) {
    if (j) PetscCall(DSVectors_NHEP_Some(ds, j, rnorm));
    else PetscCall(DSVectors_NHEP_All(ds));
  } else SETERRQ(PETSC_ERR_SUP, "DSVectors is not defined for DS type svd");
  return PETSC_SUCCESS;
}

/* DSTruncate dsops.c:232-254 → DSTruncate_HEP dshep.c:643, _NHEP dsnhep.c:394, _SVD dssvd.c:223 */
PetscErrorCode DSTruncate(DS ds, PetscInt n, PetscBool trim)
{
  DSCheckAlloc(ds);
  PetscCheck(n >= ds->l && n <= ds->n, PETSC_ERR_ARG_OUTOFRANGE, "Illegal value of n (%d). Must be between l (%d) and n (%d)", n, ds->l, ds->n);
  const PetscInt l = ds->l, ld = ds->ld;
  const PetscBool dense = IS(ds, DSNHEP) ? PETSC_TRUE : PETSC_FALSE;
  PetscScalar *A = ds->mat[DS_MAT_A];
  if (dense && ds->state == DS_STATE_CONDENSED && n < ds->n && n > 0)   /* dsnhep.c:400-404 */
    PetscCheck(A[n + (size_t)(n - 1) * ld] == 0.0, PETSC_ERR_ARG_WRONG, "The given size would break a 2x2 block, call DSGetTruncateSize() first");
  if (trim) {
    if (dense && ds->extrarow) for (PetscInt i = l; i < ds->n; i++) A[ds->n + (size_t)i * ld] = 0.0;   /* clean extra row */
    ds->l = 0; ds->k = 0; ds->n = n; ds->t = n;
    if (IS(ds, DSSVD)) { ds->m = n; ds->tm = n; }
  } else {
    if (dense && ds->extrarow && ds->k == ds->n) {
      for (PetscInt i = l; i < n; i++) A[n + (size_t)i * ld] = A[ds->n + (size_t)i * ld];   /* move the extra row up */
      for (PetscInt i = l; i < ds->n; i++) A[ds->n + (size_t)i * ld] = 0.0;
    }
    ds->k = ds->extrarow ? n : 0;
    ds->t = ds->n;
    if (IS(ds, DSSVD)) { ds->tm = ds->m; ds->m = n; }
    ds->n = n;
  }
  ds->state = trim ? DS_STATE_RAW : DS_STATE_TRUNCATED;
  return PETSC_SUCCESS;
}

/* dsops.c:329-400: do not split a 2x2 diagonal block (NHEP only) */
PetscErrorCode DSGetTruncateSize(DS ds, PetscInt l, PetscInt n, PetscInt *k)
{
  DSCheckAlloc(ds);
  PetscCheck(k, PETSC_ERR_ARG_NULL, "null k");
  if (!IS(ds, DSNHEP)) return PETSC_SUCCESS;
  if (!l) l = ds->l;
  if (!n) n = ds->n;
  if (ds->mat[DS_MAT_A][l + *k + (size_t)(l + *k - 1) * ds->ld] != 0.0) {
    if (l + *k < n - 1) (*k)++;
    else (*k)--;
  }
  return PETSC_SUCCESS;
}
