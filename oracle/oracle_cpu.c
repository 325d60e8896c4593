/*
 * ORACLE / CPU BASELINE — TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * A host-memory BV type ("oraclecpu") and a host CSR Mat type ("oraclecsr") that plug into the same
 * C host driver as the product (libb2kslepc: bv.c / ds.c / eps.c / svd.c) through BVRegister() and the
 * Mat ops table.  They restate the reference's CPU kernels:
 *   BVMult_BLAS_Private / BVMultVec / BVMultInPlace / BVDot / BVDotVec / BVScale / BVNorm
 *        src/sys/classes/bv/interface/bvblas.c:24-278, bvlapack.c:37-83  (host BLAS: dgemm/dgemv)
 *   BVSVEC storage and ops table   src/sys/classes/bv/impls/svec/svec.c:17-563
 *   MatMult (PETSc AIJ)            reached from bvops.c:879 — CSR SpMV, here with OpenMP
 * No `gramschmidt` hook is installed, so the front-end's BVOrthogonalizeCGS1 (bvorthog.c:91-132) runs
 * exactly as in the reference: gemv^T, gemv, estimated norm.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load this
 * library (oracle/_build/liboraclecpu.so).  Nothing under slepc_b200/ links or dlopens it.
 * Parity status: PINNED through tests/test_host_cpu.py (reference goldens, see oracle/slepc_oracle.py).
 */
#include <omp.h>
#include "b2kimpl.h"

void LAPACK(dgemm)(const char *, const char *, const int *, const int *, const int *, const double *, const double *, const int *,
                   const double *, const int *, const double *, double *, const int *, size_t, size_t);

typedef struct { double *V; double *ts_fact; PetscInt ts_nk, ts_rows; } BV_CPU;   /* ts_*: LAPACK-factored copy kept between tsqr_factor and tsqr_formq */
#define COL(bv, d, j) ((d)->V + (size_t)((bv)->nc + (j)) * (size_t)(bv)->ld)

/* ---- host Vec arithmetic registered with libb2kslepc (B2KVecRegisterHostOps, sys.c) ---------------- */
typedef struct {
  PetscErrorCode (*set)(Vec, PetscScalar);
  PetscErrorCode (*sumsq_local)(Vec, PetscReal *);
  PetscErrorCode (*dot_local)(Vec, Vec, PetscScalar *);
  PetscErrorCode (*axpy)(Vec, PetscScalar, Vec);
  PetscErrorCode (*scale)(Vec, PetscScalar);
  PetscErrorCode (*pointwisemult)(Vec, Vec, Vec);        /* w = x .* y */
} B2KVecHostOps;
PetscErrorCode B2KVecRegisterHostOps(const B2KVecHostOps *ops);

static PetscErrorCode hv_set(Vec v, PetscScalar a) { for (PetscInt i = 0; i < v->n; i++) v->array[i] = a; return PETSC_SUCCESS; }
static PetscErrorCode hv_sumsq(Vec v, PetscReal *s)
{
  double t = 0.0;
  const double *x = v->array;
  const PetscInt n = v->n;
#pragma omp parallel for reduction(+ : t) if (n > 100000)
  for (PetscInt i = 0; i < n; i++) t += x[i] * x[i];
  *s = t;
  return PETSC_SUCCESS;
}
static PetscErrorCode hv_dot(Vec x, Vec y, PetscScalar *s)
{
  double t = 0.0;
  const double *a = x->array, *b = y->array;
  const PetscInt n = x->n;
#pragma omp parallel for reduction(+ : t) if (n > 100000)
  for (PetscInt i = 0; i < n; i++) t += a[i] * b[i];
  *s = t;
  return PETSC_SUCCESS;
}
static PetscErrorCode hv_axpy(Vec y, PetscScalar a, Vec x)
{
  double *yy = y->array;
  const double *xx = x->array;
  const PetscInt n = y->n;
#pragma omp parallel for if (n > 100000)
  for (PetscInt i = 0; i < n; i++) yy[i] += a * xx[i];
  return PETSC_SUCCESS;
}
static PetscErrorCode hv_scale(Vec v, PetscScalar a)
{
  double *x = v->array;
  const PetscInt n = v->n;
#pragma omp parallel for if (n > 100000)
  for (PetscInt i = 0; i < n; i++) x[i] *= a;
  return PETSC_SUCCESS;
}

static PetscErrorCode hv_pointwisemult(Vec w, Vec x, Vec y)
{
  double *ww = w->array;
  const double *xx = x->array, *yy = y->array;
  const PetscInt n = w->n;
#pragma omp parallel for if (n > 100000)
  for (PetscInt i = 0; i < n; i++) ww[i] = xx[i] * yy[i];
  return PETSC_SUCCESS;
}

/* ---- BV ops, shapes of bvblas.c -------------------------------------------------------------------- */
/* gemv^T in row chunks so that every thread streams its own slab of V (bvblas.c:240-261 + MPIU_Allreduce) */
static void gemvT_chunked(const double *V, PetscInt ld, PetscInt n, PetscInt k, const double *y, double *q)
{
  const int nth = omp_get_max_threads();
  if (n < 200000 || nth == 1) {
    const int ni = n, ki = k, ldi = ld, inc = 1;
    const double one = 1.0, zero = 0.0;
    LAPACK(dgemv)("C", &ni, &ki, &one, V, &ldi, y, &inc, &zero, q, &inc, 1);
    return;
  }
  double *part = (double *)calloc((size_t)nth * (size_t)k, sizeof(double));
#pragma omp parallel
  {
    const int t = omp_get_thread_num(), T = omp_get_num_threads();
    const PetscInt r0 = (PetscInt)((int64_t)n * t / T), r1 = (PetscInt)((int64_t)n * (t + 1) / T);
    const int ni = r1 - r0, ki = k, ldi = ld, inc = 1;
    const double one = 1.0, zero = 0.0;
    if (ni > 0) LAPACK(dgemv)("C", &ni, &ki, &one, V + r0, &ldi, y + r0, &inc, &zero, part + (size_t)t * k, &inc, 1);
  }
  for (PetscInt c = 0; c < k; c++) { double s = 0.0; for (int t = 0; t < nth; t++) s += part[(size_t)t * k + c]; q[c] = s; }
  free(part);
}

static void gemvN_chunked(const double *V, PetscInt ld, PetscInt n, PetscInt k, double alpha, const double *q, double beta, double *y)
{
  if (n < 200000 || omp_get_max_threads() == 1) {
    const int ni = n, ki = k, ldi = ld, inc = 1;
    LAPACK(dgemv)("N", &ni, &ki, &alpha, V, &ldi, q, &inc, &beta, y, &inc, 1);
    return;
  }
#pragma omp parallel
  {
    const int t = omp_get_thread_num(), T = omp_get_num_threads();
    const PetscInt r0 = (PetscInt)((int64_t)n * t / T), r1 = (PetscInt)((int64_t)n * (t + 1) / T);
    const int ni = r1 - r0, ki = k, ldi = ld, inc = 1;
    if (ni > 0) LAPACK(dgemv)("N", &ni, &ki, &alpha, V + r0, &ldi, q, &inc, &beta, y + r0, &inc, 1);
  }
}

static PetscErrorCode BVDotVec_CPU_Private(BV X, Vec y, PetscScalar *q, PetscBool reduce)
{
  BV_CPU *d = (BV_CPU *)X->data;
  const PetscInt k = X->k - X->l;
  PetscScalar *qq = q ? q : X->buffer;
  if (k <= 0) return PETSC_SUCCESS;
  if (X->n > 0) gemvT_chunked(COL(X, d, X->l), X->ld, X->n, k, y->array, qq);
  else memset(qq, 0, sizeof(double) * (size_t)k);
  if (reduce) PetscCall(B2KCommAllreduce(X->comm, qq, k, 0, B2K_MEM_HOST));
  return PETSC_SUCCESS;
}
static PetscErrorCode BVDotVec_CPU(BV X, Vec y, PetscScalar *q) { return BVDotVec_CPU_Private(X, y, q, PETSC_TRUE); }
static PetscErrorCode BVDotVec_Local_CPU(BV X, Vec y, PetscScalar *q) { return BVDotVec_CPU_Private(X, y, q, PETSC_FALSE); }

static PetscErrorCode BVMultVec_CPU(BV X, PetscScalar alpha, PetscScalar beta, Vec y, PetscScalar *q)
{
  BV_CPU *d = (BV_CPU *)X->data;
  const PetscInt k = X->k - X->l;
  const PetscScalar *qq = q ? q : X->buffer;
  if (X->n == 0) return PETSC_SUCCESS;
  if (k > 0) gemvN_chunked(COL(X, d, X->l), X->ld, X->n, k, alpha, qq, beta, y->array);
  else if (beta != 1.0) for (PetscInt i = 0; i < X->n; i++) y->array[i] *= beta;
  return PETSC_SUCCESS;
}

/* C(n x nc) = beta*C + alpha * A(n x k) * B(k x nc), row-chunked dgemm */
static void gemm_rows(double *C, PetscInt ldc, const double *A, PetscInt lda, PetscInt n, PetscInt k, PetscInt nc, double alpha, double beta,
                      const double *B, PetscInt ldb, int transb)
{
#pragma omp parallel
  {
    const int t = omp_get_thread_num(), T = omp_get_num_threads();
    const PetscInt r0 = (PetscInt)((int64_t)n * t / T), r1 = (PetscInt)((int64_t)n * (t + 1) / T);
    const int ni = r1 - r0, ki = k, nci = nc, ldai = lda, ldbi = ldb, ldci = ldc;
    if (ni > 0 && nci > 0) LAPACK(dgemm)("N", transb ? "T" : "N", &ni, &nci, &ki, &alpha, A + r0, &ldai, B, &ldbi, &beta, C + r0, &ldci, 1, 1);
  }
}

static PetscErrorCode BVMult_CPU(BV Y, PetscScalar alpha, PetscScalar beta, BV X, Mat Q)
{
  BV_CPU *y = (BV_CPU *)Y->data, *x = (BV_CPU *)X->data;
  const PetscInt ky = Y->k - Y->l, kx = X->k - X->l;
  if (ky <= 0 || Y->n == 0) return PETSC_SUCCESS;
  if (!Q) {
    for (PetscInt j = 0; j < ky; j++) {
      double *yc = COL(Y, y, Y->l + j);
      const double *xc = COL(X, x, X->l + j);
      for (PetscInt i = 0; i < Y->n; i++) yc[i] = alpha * xc[i] + beta * yc[i];
    }
    return PETSC_SUCCESS;
  }
  gemm_rows(COL(Y, y, Y->l), Y->ld, COL(X, x, X->l), X->ld, Y->n, kx, ky, alpha, beta, Q->dense + (size_t)Y->l * Q->lda + X->l, Q->lda, 0);
  return PETSC_SUCCESS;
}

/* bvblas.c:74-106: out-of-place product through a workspace, then copy back */
static PetscErrorCode BVMultInPlace_CPU_Private(BV V, Mat Q, PetscInt s, PetscInt e, int trans)
{
  BV_CPU *d = (BV_CPU *)V->data;
  const PetscInt l = V->l, k = V->k - V->l, ncol = e - s, n = V->n;
  if (ncol <= 0 || n == 0) return PETSC_SUCCESS;
  const size_t need = (size_t)n * (size_t)ncol;
  if (V->lwork < need) { free(V->work); V->work = (double *)malloc(sizeof(double) * need); V->lwork = need; PetscCheck(V->work, PETSC_ERR_MEM, "out of memory"); }
  const double *B = trans ? Q->dense + (size_t)l * Q->lda + s : Q->dense + (size_t)s * Q->lda + l;
  gemm_rows(V->work, n, COL(V, d, l), V->ld, n, k, ncol, 1.0, 0.0, B, Q->lda, trans);
#pragma omp parallel for
  for (PetscInt j = 0; j < ncol; j++) memcpy(COL(V, d, s + j), V->work + (size_t)j * n, sizeof(double) * (size_t)n);
  return PETSC_SUCCESS;
}
static PetscErrorCode BVMultInPlace_CPU(BV V, Mat Q, PetscInt s, PetscInt e) { return BVMultInPlace_CPU_Private(V, Q, s, e, 0); }
static PetscErrorCode BVMultInPlaceTrans_CPU(BV V, Mat Q, PetscInt s, PetscInt e) { return BVMultInPlace_CPU_Private(V, Q, s, e, 1); }

static PetscErrorCode BVDot_CPU(BV X, BV Y, Mat M)
{
  BV_CPU *x = (BV_CPU *)X->data, *y = (BV_CPU *)Y->data;
  const PetscInt ky = Y->k - Y->l, kx = X->k - X->l;
  double *tmp = (double *)calloc((size_t)ky * (size_t)kx + 1, sizeof(double));
  PetscCheck(tmp, PETSC_ERR_MEM, "out of memory");
  if (X->n > 0) {
    const int kyi = ky, kxi = kx, ni = X->n, ldy = Y->ld, ldx = X->ld;
    const double one = 1.0, zero = 0.0;
    LAPACK(dgemm)("T", "N", &kyi, &kxi, &ni, &one, COL(Y, y, Y->l), &ldy, COL(X, x, X->l), &ldx, &zero, tmp, &kyi, 1, 1);
  }
  PetscErrorCode ierr = B2KCommAllreduce(X->comm, tmp, ky * kx, 0, B2K_MEM_HOST);
  for (PetscInt j = 0; j < kx; j++) memcpy(M->dense + (size_t)(X->l + j) * M->lda + Y->l, tmp + (size_t)j * ky, sizeof(double) * (size_t)ky);
  free(tmp);
  PetscCall(ierr);
  return PETSC_SUCCESS;
}

static PetscErrorCode BVScale_CPU(BV bv, PetscInt j, PetscScalar alpha)
{
  BV_CPU *d = (BV_CPU *)bv->data;
  const PetscInt j0 = (j < 0) ? bv->l : j, j1 = (j < 0) ? bv->k : j + 1;
  for (PetscInt c = j0; c < j1; c++) {
    double *x = COL(bv, d, c);
    const PetscInt n = bv->n;
#pragma omp parallel for if (n > 100000)
    for (PetscInt i = 0; i < n; i++) x[i] *= alpha;
  }
  return PETSC_SUCCESS;
}

static PetscErrorCode BVNorm_CPU_Private(BV bv, PetscInt j, NormType type, PetscReal *val, PetscBool reduce)
{
  BV_CPU *d = (BV_CPU *)bv->data;
  const PetscInt j0 = (j < 0) ? bv->l : j, j1 = (j < 0) ? bv->k : j + 1, n = bv->n;
  if (type == NORM_2 || type == NORM_FROBENIUS) {
    double s = 0.0;
    for (PetscInt c = j0; c < j1; c++) {
      const double *x = COL(bv, d, c);
      double t = 0.0;
#pragma omp parallel for reduction(+ : t) if (n > 100000)
      for (PetscInt i = 0; i < n; i++) t += x[i] * x[i];
      s += t;
    }
    if (reduce) PetscCall(B2KCommAllreduce(bv->comm, &s, 1, 0, B2K_MEM_HOST));
    *val = sqrt(s);
  } else if (type == NORM_1) {
    double mx = 0.0;
    for (PetscInt c = j0; c < j1; c++) {
      const double *x = COL(bv, d, c);
      double t = 0.0;
      for (PetscInt i = 0; i < n; i++) t += fabs(x[i]);
      if (reduce) PetscCall(B2KCommAllreduce(bv->comm, &t, 1, 0, B2K_MEM_HOST));
      mx = PetscMax(mx, t);
    }
    *val = mx;
  } else {
    double mx = 0.0;
    for (PetscInt i = 0; i < n; i++) { double t = 0.0; for (PetscInt c = j0; c < j1; c++) t += fabs(COL(bv, d, c)[i]); mx = PetscMax(mx, t); }
    if (reduce) PetscCall(B2KCommAllreduce(bv->comm, &mx, 1, 1, B2K_MEM_HOST));
    *val = mx;
  }
  return PETSC_SUCCESS;
}
static PetscErrorCode BVNorm_CPU(BV bv, PetscInt j, NormType t, PetscReal *v) { return BVNorm_CPU_Private(bv, j, t, v, PETSC_TRUE); }
static PetscErrorCode BVNorm_Local_CPU(BV bv, PetscInt j, NormType t, PetscReal *v) { return BVNorm_CPU_Private(bv, j, t, v, PETSC_FALSE); }

static PetscErrorCode BVNormalize_CPU(BV bv, PetscScalar *eigi)
{
  for (PetscInt i = bv->l; i < bv->k; i++) {
    const PetscInt cols = (eigi && eigi[i] != 0.0 && i + 1 < bv->k) ? 2 : 1;
    const PetscInt lsave = bv->l, ksave = bv->k;
    PetscReal nrm = 0.0;
    bv->l = i; bv->k = i + cols;
    PetscErrorCode ierr = BVNorm_CPU(bv, -1, NORM_FROBENIUS, &nrm);
    if (!ierr && nrm != 0.0 && nrm != 1.0) ierr = BVScale_CPU(bv, -1, 1.0 / nrm);
    bv->l = lsave; bv->k = ksave;
    PetscCall(ierr);
    i += cols - 1;
  }
  return PETSC_SUCCESS;
}

static PetscErrorCode BVMatMult_CPU(BV V, Mat A, BV W)
{
  BV_CPU *v = (BV_CPU *)V->data, *w = (BV_CPU *)W->data;
  Vec x, y;
  PetscCall(VecCreateWithArray(B2K_MEM_HOST, V->n, V->N, NULL, &x));
  PetscCall(VecCreateWithArray(B2K_MEM_HOST, W->n, W->N, NULL, &y));
  PetscErrorCode ierr = PETSC_SUCCESS;
  for (PetscInt j = 0; j < V->k - V->l && !ierr; j++) { x->array = COL(V, v, V->l + j); y->array = COL(W, w, W->l + j); ierr = MatMult(A, x, y); }
  PetscCall(VecDestroy(&x));
  PetscCall(VecDestroy(&y));
  PetscCall(ierr);
  return PETSC_SUCCESS;
}

static PetscErrorCode BVCopy_CPU(BV V, BV W)
{
  BV_CPU *v = (BV_CPU *)V->data, *w = (BV_CPU *)W->data;
  for (PetscInt j = 0; j < V->k - V->l; j++) memcpy(COL(W, w, W->l + j), COL(V, v, V->l + j), sizeof(double) * (size_t)V->n);
  return PETSC_SUCCESS;
}
static PetscErrorCode BVCopyColumn_CPU(BV V, PetscInt j, PetscInt i)
{
  BV_CPU *d = (BV_CPU *)V->data;
  memcpy(COL(V, d, i), COL(V, d, j), sizeof(double) * (size_t)V->n);
  return PETSC_SUCCESS;
}
static PetscErrorCode BVResize_CPU(BV bv, PetscInt m, PetscBool copy)
{
  BV_CPU *d = (BV_CPU *)bv->data;
  double *Vn = (double *)calloc((size_t)(bv->nc + m) * (size_t)bv->ld + 2, sizeof(double));
  PetscCheck(Vn, PETSC_ERR_MEM, "out of memory");
  if (copy) memcpy(Vn, d->V, sizeof(double) * (size_t)(bv->nc + PetscMin(m, bv->m)) * (size_t)bv->ld);
  free(d->V);
  d->V = Vn;
  return PETSC_SUCCESS;
}
static PetscErrorCode BVGetColumn_CPU(BV bv, PetscInt j, Vec *v)
{
  (void)v;
  PetscCall(VecPlaceArray(bv->cv[(bv->ci[0] == j) ? 0 : 1], COL(bv, (BV_CPU *)bv->data, j)));
  return PETSC_SUCCESS;
}
static PetscErrorCode BVRestoreColumn_CPU(BV bv, PetscInt j, Vec *v)
{
  (void)v;
  PetscCall(VecResetArray(bv->cv[(bv->ci[0] == j) ? 0 : 1]));
  return PETSC_SUCCESS;
}
static PetscErrorCode BVGetArray_CPU(BV bv, PetscScalar **a) { *a = ((BV_CPU *)bv->data)->V; return PETSC_SUCCESS; }
static PetscErrorCode BVSetRandomColumn_CPU(BV bv, PetscInt j)
{
  double *x = COL(bv, (BV_CPU *)bv->data, j);
  for (PetscInt i = 0; i < bv->n; i++) x[i] = B2KHashUniform((uint64_t)(bv->row0 + i), bv->rng_seed + (uint64_t)j);
  return PETSC_SUCCESS;
}
/* TEST DOUBLE of the krylov_steps slot (the device type enqueues a whole restart cycle and may stop early at a step it cannot
   finish on its own: tests/test_host_cpu.py drives the front-end's resume logic of BVKrylovLoop_Private with it).  Active only
   when ORACLE_KRYLOV_STEPS is set to a comma-separated pattern: the i-th call completes min(pattern[i mod len], m-k) steps with
   the step-by-step arithmetic and reports where it stopped; 0 = declines; a negative entry -s completes s steps and then runs
   ONE MORE step that it disowns (its column is left modified, as after a device-side breakdown: the loop must redo it). */
static PetscErrorCode BVKrylovSteps_CPU(BV V, Mat A, PetscInt k, PetscInt m, PetscInt *jnext, PetscReal *beta)
{
  static int ncall = 0;
  const char *pat = getenv("ORACLE_KRYLOV_STEPS");
  *jnext = k;
  if (!pat || !*pat) return PETSC_SUCCESS;
  int vals[64], nv = 0;
  for (const char *p = pat; *p && nv < 64;) { vals[nv++] = atoi(p); while (*p && *p != ',') p++; if (*p == ',') p++; }
  if (!nv) return PETSC_SUCCESS;
  const int want = vals[ncall++ % nv];
  PetscInt todo = want < 0 ? -want : want;
  if (todo > m - k) todo = m - k;
  PetscErrorCode (*self)(BV, Mat, PetscInt, PetscInt, PetscInt *, PetscReal *) = V->ops.krylov_steps;
  V->ops.krylov_steps = NULL;                      /* the public calls below must not come back here */
  PetscErrorCode ierr = PETSC_SUCCESS;
  PetscInt j = k;
  for (; !ierr && j < k + todo; j++) {
    PetscBool lindep = PETSC_FALSE;
    PetscReal b = 0.0;
    ierr = BVMatMultColumn(V, A, j);
    if (!ierr) ierr = BVOrthonormalizeColumn(V, j + 1, PETSC_FALSE, &b, &lindep);
    if (!ierr && lindep) break;                    /* a breakdown is the loop's business: this step is not reported as done */
    if (!ierr) *beta = b;
  }
  if (!ierr && want < 0 && j == k + todo && j < m) {   /* one disowned step: column j+1 and its coefficients are garbage afterwards */
    ierr = BVMatMultColumn(V, A, j);
    if (!ierr) ierr = V->ops.scale(V, j + 1, 3.0);
  }
  V->ops.krylov_steps = self;
  PetscCall(ierr);
  *jnext = j;
  return PETSC_SUCCESS;
}

/* local Householder QR of the active columns with LAPACK, as BVOrthogonalize_LAPACK_TSQR does on the raw array (bvlapack.c:378-396):
   geqrf -> R; with wantq the factored copy is kept and tsqr_formq applies it to [W ; 0] (ormqr) */
static PetscErrorCode BVTSQRFactor_CPU(BV bv, PetscBool wantq, PetscScalar *R)
{
  BV_CPU *d = (BV_CPU *)bv->data;
  const PetscInt nk = bv->k - bv->l, rows = bv->n > nk ? bv->n : nk;      /* padded with zero rows when this rank has fewer rows than columns */
  free(d->ts_fact); d->ts_fact = NULL;
  if (nk <= 0) return PETSC_SUCCESS;
  int rows_ = rows, nk_ = nk, lwork = 64 * (int)nk + 64, info = 0;
  double *F = (double *)calloc((size_t)rows * (size_t)nk + (size_t)nk + (size_t)lwork, sizeof(double));
  PetscCheck(F, PETSC_ERR_MEM, "out of memory");
  for (PetscInt c = 0; c < nk; c++) memcpy(F + (size_t)c * rows, COL(bv, d, bv->l + c), sizeof(double) * (size_t)bv->n);
  LAPACK(dgeqrf)(&rows_, &nk_, F, &rows_, F + (size_t)rows * nk, F + (size_t)rows * nk + nk, &lwork, &info);
  if (info) { free(F); SETERRQ(PETSC_ERR_LIB, "Error in LAPACK subroutine geqrf: info=%d", info); }
  for (PetscInt c = 0; c < nk; c++) for (PetscInt i = 0; i < nk; i++) R[i + (size_t)c * nk] = (i <= c) ? F[i + (size_t)c * rows] : 0.0;
  if (wantq) { d->ts_fact = F; d->ts_nk = nk; d->ts_rows = rows; }
  else free(F);
  return PETSC_SUCCESS;
}
static PetscErrorCode BVTSQRFormQ_CPU(BV bv, const PetscScalar *W)
{
  BV_CPU *d = (BV_CPU *)bv->data;
  const PetscInt nk = bv->k - bv->l;
  if (nk <= 0) return PETSC_SUCCESS;
  PetscCheck(d->ts_fact && d->ts_nk == nk, PETSC_ERR_ORDER, "tsqr_formq without a matching tsqr_factor(wantq)");
  const PetscInt rows = d->ts_rows;
  int rows_ = rows, nk_ = nk, lwork = 64 * (int)nk + 64, info = 0;
  double *C = (double *)calloc((size_t)rows * (size_t)nk + (size_t)lwork, sizeof(double));
  PetscCheck(C, PETSC_ERR_MEM, "out of memory");
  for (PetscInt c = 0; c < nk; c++) memcpy(C + (size_t)c * rows, W + (size_t)c * nk, sizeof(double) * (size_t)nk);
  LAPACK(dormqr)("L", "N", &rows_, &nk_, &nk_, d->ts_fact, &rows_, d->ts_fact + (size_t)rows * nk, C, &rows_, C + (size_t)rows * nk, &lwork, &info, 1, 1);
  if (!info) for (PetscInt c = 0; c < nk; c++) memcpy(COL(bv, d, bv->l + c), C + (size_t)c * rows, sizeof(double) * (size_t)bv->n);
  free(C);
  free(d->ts_fact); d->ts_fact = NULL;
  PetscCheck(!info, PETSC_ERR_LIB, "Error in LAPACK subroutine ormqr: info=%d", info);
  return PETSC_SUCCESS;
}

static PetscErrorCode BVDestroy_CPU(BV bv)
{
  BV_CPU *d = (BV_CPU *)bv->data;
  if (d) { free(d->ts_fact); free(d->V); free(d); }
  bv->data = NULL;
  return PETSC_SUCCESS;
}

PetscErrorCode BVCreate_OracleCPU(BV bv)
{
  BV_CPU *d = (BV_CPU *)calloc(1, sizeof(*d));
  PetscCheck(d, PETSC_ERR_MEM, "out of memory");
  d->V = (double *)calloc((size_t)(bv->nc + bv->m) * (size_t)bv->ld + 2, sizeof(double));
  PetscCheck(d->V, PETSC_ERR_MEM, "out of memory");
  bv->data = d;
  bv->mem = B2K_MEM_HOST;
  bv->ops.mult = BVMult_CPU;
  bv->ops.multvec = BVMultVec_CPU;
  bv->ops.multinplace = BVMultInPlace_CPU;
  bv->ops.multinplacetrans = BVMultInPlaceTrans_CPU;
  bv->ops.dot = BVDot_CPU;
  bv->ops.dotvec = BVDotVec_CPU;
  bv->ops.dotvec_local = BVDotVec_Local_CPU;
  bv->ops.scale = BVScale_CPU;
  bv->ops.norm = BVNorm_CPU;
  bv->ops.norm_local = BVNorm_Local_CPU;
  bv->ops.normalize = BVNormalize_CPU;
  bv->ops.matmult = BVMatMult_CPU;
  bv->ops.copy = BVCopy_CPU;
  bv->ops.copycolumn = BVCopyColumn_CPU;
  bv->ops.resize = BVResize_CPU;
  bv->ops.getcolumn = BVGetColumn_CPU;
  bv->ops.restorecolumn = BVRestoreColumn_CPU;
  bv->ops.getarray = BVGetArray_CPU;
  bv->ops.destroy = BVDestroy_CPU;
  bv->ops.setrandomcolumn = BVSetRandomColumn_CPU;
  bv->ops.tsqr_factor = BVTSQRFactor_CPU;
  bv->ops.tsqr_formq = BVTSQRFormQ_CPU;
  if (getenv("ORACLE_KRYLOV_STEPS")) bv->ops.krylov_steps = BVKrylovSteps_CPU;     /* test double, see above */
  return PETSC_SUCCESS;
}

/* ---- host CSR Mat ("oraclecsr"): local column numbering [owned | ghosts], halo through the callback communicator ---- */
typedef struct {
  PetscInt  nghost, nrecv, nsend, nsendtot;
  PetscInt *rowptr, *colidx;
  double   *val, *xfull, *sendbuf;
  PetscInt *recvrank, *recvcount, *sendrank, *sendcount, *sendidx;
} Mat_CPUCSR;

static PetscErrorCode MatMult_CPUCSR(Mat A, Vec x, Vec y)
{
  Mat_CPUCSR *a = (Mat_CPUCSR *)A->data;
  const double *xx = x->array;
  if (a->nghost || a->nsend) {                      /* a rank that only SENDS still takes part in the exchange */
    memcpy(a->xfull, x->array, sizeof(double) * (size_t)A->n);
    for (PetscInt i = 0; i < a->nsendtot; i++) a->sendbuf[i] = x->array[a->sendidx[i]];
    /* pairwise exchange in rank order (the callback is a blocking sendrecv) */
    PetscInt soff = 0, roff = 0, q = 0, p = 0;
    while (q < a->nsend || p < a->nrecv) {
      const PetscInt sr = q < a->nsend ? a->sendrank[q] : 1 << 30, rr = p < a->nrecv ? a->recvrank[p] : 1 << 30;
      const PetscInt peer = PetscMin(sr, rr);
      const PetscInt ns = (sr == peer) ? a->sendcount[q] : 0, nr = (rr == peer) ? a->recvcount[p] : 0;
      PetscCall(B2KCommSendrecv(B2KCommWorld(), a->sendbuf + soff, ns, peer, a->xfull + A->n + roff, nr, peer, B2K_MEM_HOST));
      if (sr == peer) { soff += ns; q++; }
      if (rr == peer) { roff += nr; p++; }
    }
    xx = a->xfull;
  }
  const PetscInt m = A->m;
  const PetscInt *rp = a->rowptr, *ci = a->colidx;
  const double *v = a->val;
  double *yy = y->array;
#pragma omp parallel for schedule(static)
  for (PetscInt r = 0; r < m; r++) {
    double s = 0.0;
    for (PetscInt k = rp[r]; k < rp[r + 1]; k++) s += v[k] * xx[ci[k]];
    yy[r] = s;
  }
  return PETSC_SUCCESS;
}

/* diagonal of the owned block (square matrices whose owned columns are the owned rows: local column r is global row r) */
static PetscErrorCode MatGetDiagonal_CPUCSR(Mat A, Vec d)
{
  Mat_CPUCSR *a = (Mat_CPUCSR *)A->data;
  PetscCheck(A->M == A->N && A->rstart == A->cstart && A->m == A->n, PETSC_ERR_SUP, "MatGetDiagonal needs a square matrix whose owned columns are its owned rows");
  for (PetscInt r = 0; r < A->m; r++) {
    double v = 0.0;
    for (PetscInt k = a->rowptr[r]; k < a->rowptr[r + 1]; k++) if (a->colidx[k] == r) v = a->val[k];
    d->array[r] = v;
  }
  return PETSC_SUCCESS;
}

static PetscErrorCode MatDestroy_CPUCSR(Mat A)
{
  Mat_CPUCSR *a = (Mat_CPUCSR *)A->data;
  if (!a) return PETSC_SUCCESS;
  free(a->rowptr); free(a->colidx); free(a->val); free(a->xfull); free(a->sendbuf);
  free(a->recvrank); free(a->recvcount); free(a->sendrank); free(a->sendcount); free(a->sendidx);
  free(a);
  A->data = NULL;
  return PETSC_SUCCESS;
}

static PetscInt *dupi(const PetscInt *s, PetscInt n) { PetscInt *d = (PetscInt *)malloc(sizeof(PetscInt) * (size_t)(n > 0 ? n : 1)); if (n > 0) memcpy(d, s, sizeof(PetscInt) * (size_t)n); return d; }

/* colidx are LOCAL: < (cend-cstart) owned, otherwise ncl + ghost slot.  Peers must be listed in increasing rank order. */
PetscErrorCode MatCreateOracleCSR(PetscInt M, PetscInt N, PetscInt rstart, PetscInt rend, PetscInt cstart, PetscInt cend, const PetscInt *rowptr,
                                  const PetscInt *colidx_local, const PetscScalar *val, PetscInt nghost, PetscInt nrecv, const PetscInt *recvrank,
                                  const PetscInt *recvcount, PetscInt nsend, const PetscInt *sendrank, const PetscInt *sendcount,
                                  const PetscInt *sendidx, Mat *out)
{
  Mat A;
  PetscCall(MatCreate_Private(&A));
  Mat_CPUCSR *a = (Mat_CPUCSR *)calloc(1, sizeof(*a));
  PetscCheck(a, PETSC_ERR_MEM, "out of memory");
  strcpy(A->type, "oraclecsr");
  A->M = M; A->N = N; A->m = rend - rstart; A->n = cend - cstart;
  A->rstart = rstart; A->rend = rend; A->cstart = cstart; A->cend = cend;
  A->mem = B2K_MEM_HOST;
  A->data = a;
  A->ops.mult = MatMult_CPUCSR;
  A->ops.getdiagonal = MatGetDiagonal_CPUCSR;
  A->ops.destroy = MatDestroy_CPUCSR;
  const PetscInt m = A->m, nnz = m ? rowptr[m] : 0;
  a->rowptr = dupi(rowptr, m + 1);
  a->colidx = dupi(colidx_local, nnz);
  a->val = (double *)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
  if (nnz) memcpy(a->val, val, sizeof(double) * (size_t)nnz);
  a->nghost = nghost;
  if (nghost) {
    PetscInt stot = 0;
    for (PetscInt q = 0; q < nsend; q++) stot += sendcount[q];
    a->nrecv = nrecv; a->nsend = nsend; a->nsendtot = stot;
    a->recvrank = dupi(recvrank, nrecv); a->recvcount = dupi(recvcount, nrecv);
    a->sendrank = dupi(sendrank, nsend); a->sendcount = dupi(sendcount, nsend); a->sendidx = dupi(sendidx, stot);
    a->xfull = (double *)calloc((size_t)A->n + (size_t)nghost + 1, sizeof(double));
    a->sendbuf = (double *)calloc((size_t)stot + 1, sizeof(double));
  } else if (nsend) {
    PetscInt stot = 0;
    for (PetscInt q = 0; q < nsend; q++) stot += sendcount[q];
    a->nsend = nsend; a->nsendtot = stot;
    a->sendrank = dupi(sendrank, nsend); a->sendcount = dupi(sendcount, nsend); a->sendidx = dupi(sendidx, stot);
    a->sendbuf = (double *)calloc((size_t)stot + 1, sizeof(double));
    a->xfull = (double *)calloc((size_t)A->n + 1, sizeof(double));
    a->nghost = 0;
  }
  *out = A;
  return PETSC_SUCCESS;
}

/* d-dimensional Laplacian stencil, single rank, natural ordering (ex1.c:37-48, ex2.c:39-54; 7-point analogue) */
PetscErrorCode MatCreateOracleLaplacian(PetscInt dim, PetscInt nx, PetscInt ny, PetscInt nz, Mat *out)
{
  if (dim < 3) nz = 1;
  if (dim < 2) ny = 1;
  const int64_t plane = (int64_t)ny * nz, N = (int64_t)nx * plane;
  PetscCheck(N < 2147483647LL / 8, PETSC_ERR_ARG_OUTOFRANGE, "grid too large");
  PetscInt *rp = (PetscInt *)malloc(sizeof(PetscInt) * (size_t)(N + 1));
  PetscCheck(rp, PETSC_ERR_MEM, "out of memory");
  rp[0] = 0;
  for (int64_t g = 0; g < N; g++) {
    const int64_t i = g / plane, rem = g - i * plane, j = rem / nz, k = rem - j * nz;
    int c = 1 + (i > 0) + (i < nx - 1);
    if (dim >= 2) c += (j > 0) + (j < ny - 1);
    if (dim >= 3) c += (k > 0) + (k < nz - 1);
    rp[g + 1] = rp[g] + c;
  }
  const int64_t nnz = rp[N];
  PetscInt *ci = (PetscInt *)malloc(sizeof(PetscInt) * (size_t)nnz);
  double *v = (double *)malloc(sizeof(double) * (size_t)nnz);
  PetscCheck(ci && v, PETSC_ERR_MEM, "out of memory");
#pragma omp parallel for schedule(static)
  for (int64_t g = 0; g < N; g++) {
    const int64_t i = g / plane, rem = g - i * plane, j = rem / nz, k = rem - j * nz;
    int64_t p = rp[g];
    if (i > 0) { ci[p] = (PetscInt)(g - plane); v[p++] = -1.0; }
    if (dim >= 2 && j > 0) { ci[p] = (PetscInt)(g - nz); v[p++] = -1.0; }
    if (dim >= 3 && k > 0) { ci[p] = (PetscInt)(g - 1); v[p++] = -1.0; }
    ci[p] = (PetscInt)g; v[p++] = 2.0 * dim;
    if (dim >= 3 && k < nz - 1) { ci[p] = (PetscInt)(g + 1); v[p++] = -1.0; }
    if (dim >= 2 && j < ny - 1) { ci[p] = (PetscInt)(g + nz); v[p++] = -1.0; }
    if (i < nx - 1) { ci[p] = (PetscInt)(g + plane); v[p++] = -1.0; }
  }
  PetscErrorCode ierr = MatCreateOracleCSR((PetscInt)N, (PetscInt)N, 0, (PetscInt)N, 0, (PetscInt)N, rp, ci, v, 0, 0, NULL, NULL, 0, NULL, NULL, NULL, out);
  free(rp); free(ci); free(v);
  PetscCall(ierr);
  return PETSC_SUCCESS;
}

/* a rank that only SENDS halo values still has to take part in the exchange */
PetscErrorCode MatOracleCSRHasHalo(Mat A, PetscBool *flg) { Mat_CPUCSR *a = (Mat_CPUCSR *)A->data; *flg = (a->nghost || a->nsend) ? PETSC_TRUE : PETSC_FALSE; return PETSC_SUCCESS; }

void scipy_openblas_set_num_threads(int);

PetscErrorCode OracleCPURegister(void)
{
  scipy_openblas_set_num_threads(1);   /* parallelism comes from the OpenMP row chunks above; BLAS runs serial inside them */
  static const B2KVecHostOps ops = {hv_set, hv_sumsq, hv_dot, hv_axpy, hv_scale, hv_pointwisemult};
  PetscCall(B2KVecRegisterHostOps(&ops));
  PetscCall(BVRegister("oraclecpu", BVCreate_OracleCPU));
  return PETSC_SUCCESS;
}

int OracleCPUGetMaxThreads(void) { return omp_get_max_threads(); }
void OracleCPUSetThreads(int n) { omp_set_num_threads(n); }

/* STREAM-triad bandwidth of the host (GB/s), reported next to the CPU baseline */
double OracleCPUStreamTriad(int64_t n, int reps)
{
  double *a = (double *)malloc(sizeof(double) * (size_t)n), *b = (double *)malloc(sizeof(double) * (size_t)n), *c = (double *)malloc(sizeof(double) * (size_t)n);
  if (!a || !b || !c) { free(a); free(b); free(c); return -1.0; }
#pragma omp parallel for
  for (int64_t i = 0; i < n; i++) { a[i] = 0.0; b[i] = 1.0; c[i] = 2.0; }
  double best = 1e30;
  for (int r = 0; r < reps; r++) {
    const double t0 = omp_get_wtime();
#pragma omp parallel for
    for (int64_t i = 0; i < n; i++) a[i] = b[i] + 3.0 * c[i];
    const double t = omp_get_wtime() - t0;
    if (t < best) best = t;
  }
  const double gbs = 24.0 * (double)n / best / 1e9;
  free(a); free(b); free(c);
  return gbs;
}
