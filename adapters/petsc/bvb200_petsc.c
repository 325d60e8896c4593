/*
 * bvb200_petsc.c — BV type "b200" for a REAL SLEPc 3.22 / PETSc 3.22 installation: the constructor a maintainer registers with
 *
 *     BVRegister("b200", BVCreate_B200);                 include/slepcbv.h:259, src/sys/classes/bv/interface/bvfunc.c:646
 *
 * (from PetscDLLibraryRegister_b200() below, loaded with -dll_append libb200petsc.so) and selects with -bv_type b200 or
 * BVSetType(bv,"b200").  It is the PETSc-typed twin of slepc_b200/host/bvb200.c: the same sm_100a kernels behind the same C ABI
 * (include/b2k.h), with Vec / Mat / MPI_Comm where the stand-alone host layer has its own look-alikes.
 *
 * Every one of the 35 slots of struct _BVOps (include/slepc/private/bvimpl.h:24-61) is decided at the end of this file:
 * implemented here, or NULL where the reference's front-end has a generic path that is correct for this type.
 *
 * This file cannot be linked in this repository (no PETSc in the image): it is type-checked against the reference's OWN headers
 * with the stand-in PETSc declarations of adapters/petsc/stub/ (tests/test_adapter_syntax.py, gcc -fsyntax-only).
 * Layout facts it relies on: the basis is ONE device array of (nc+m) columns x ld doubles, column c at V + (nc+c) ld
 * (the BVSVEC layout, svec.c:397-563, bvimpl.h:471-484); bv->cuda stays PETSC_FALSE, so the coefficient buffer bv->buffer is a
 * HOST Vec handled by the *_Default helpers (bvimpl.h:289-415) while the column Vecs lent by BVGetColumn are CUDA Vecs
 * (bv->vtype) that wrap device columns with VecCUDAPlaceArray.
 */
#include <slepc/private/bvimpl.h>
#include "b2k.h"
#include "b2k_petsc.h"

typedef struct {
  double   *V;             /* device: (nc+m) x ld                                                    */
  double   *dco, *hco;     /* 4 coefficient slots in HBM + pinned host mirror                         */
  PetscInt  slot;
  double   *dQ, *hQ;       /* staging of small dense blocks                                           */
  size_t    dQ_elems;
  /* second pass prepared (and, gated on the device, already applied) behind the first: see slepc_b200/host/bvb200.c */
  PetscBool pend_valid, pend2_valid, pend2_ran, expect_refine;
  PetscInt  pend_j, last_j;
  PetscObjectState pend_state, last_state;
  PetscReal pend_nrm2, pend2_nrm2;
  double   *pend_c;
  b2k_comm  comm;          /* row-partition communicator of PetscObjectComm(bv)                        */
  PetscBool owns_V;
} BV_B200;

#define B2KCall(call) do { int rc_ = (call); PetscCheck(!rc_, PETSC_COMM_SELF, PETSC_ERR_GPU, "%s failed (%d): %s", #call, rc_, b2k_last_error()); } while (0)
#define COL(bv, d, j) ((d)->V + (size_t)((bv)->nc + (j)) * (size_t)(bv)->ld)
#define SLOT(d, i) ((d)->dco + (size_t)(i) * (size_t)(d)->slot)

static PetscErrorCode BVState_B200(BV bv, PetscObjectState *s) { return PetscObjectStateGet((PetscObject)bv, s); }

/* all-reduce (unless the kernels already summed over the GPUs through the NVLink mailboxes), one copy to pinned memory, one wait */
static PetscErrorCode BVFetch_B200(BV bv, BV_B200 *d, double *dptr, PetscInt count, PetscBool reduce, int fused, double **hptr)
{
  b2k_ctx ctx = B2KPetscContext();

  PetscFunctionBegin;
  if (fused) B2KCall(b2k_comm_reduce_scope(d->comm, 0, NULL));
  else if (reduce) B2KCall(b2k_comm_allreduce_sum(d->comm, dptr, (int)count));
  *hptr = d->hco + (dptr - d->dco);
  B2KCall(b2k_d2h_async(ctx, *hptr, dptr, sizeof(double) * (size_t)count));
  B2KCall(b2k_ctx_sync(ctx));
  (void)bv;
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* ---- level-2: svec.c:38-52,109-129 / sveccuda.cu:45-69,126-143 ------------------------------------------------------ */
static PetscErrorCode BVDotVec_B200_Private(BV X, Vec y, PetscScalar *q, PetscBool reduce)
{
  BV_B200           *d = (BV_B200 *)X->data;
  const PetscInt     k = X->k - X->l;
  const PetscScalar *py;
  PetscScalar       *qq = q;
  double            *hp;
  int                fused = 0;

  PetscFunctionBegin;
  if (k <= 0) PetscFunctionReturn(PETSC_SUCCESS);
  if (reduce) B2KCall(b2k_comm_reduce_scope(d->comm, 1, &fused));
  PetscCall(VecCUDAGetArrayRead(y, &py));
  B2KCall(b2k_dotvec(B2KPetscContext(), COL(X, d, X->l), X->ld, X->n, (int)k, py, SLOT(d, 0)));
  PetscCall(VecCUDARestoreArrayRead(y, &py));
  PetscCall(BVFetch_B200(X, d, SLOT(d, 0), k, reduce, fused, &hp));
  if (!q) PetscCall(VecGetArray(X->buffer, &qq));
  PetscCall(PetscArraycpy(qq, hp, k));
  if (!q) PetscCall(VecRestoreArray(X->buffer, &qq));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode BVDotVec_B200(BV X, Vec y, PetscScalar *q) { return BVDotVec_B200_Private(X, y, q, PETSC_TRUE); }
static PetscErrorCode BVDotVec_Local_B200(BV X, Vec y, PetscScalar *q) { return BVDotVec_B200_Private(X, y, q, PETSC_FALSE); }

static PetscErrorCode BVMultVec_B200(BV X, PetscScalar alpha, PetscScalar beta, Vec y, PetscScalar *q)
{
  BV_B200       *d = (BV_B200 *)X->data;
  b2k_ctx        ctx = B2KPetscContext();
  const PetscInt k = X->k - X->l;
  PetscScalar   *py, *qq = q;

  PetscFunctionBegin;
  if (!q) PetscCall(VecGetArray(X->buffer, &qq));
  if (k > 0) B2KCall(b2k_h2d_async(ctx, SLOT(d, 2), qq, sizeof(double) * (size_t)k));
  if (!q) PetscCall(VecRestoreArray(X->buffer, &qq));
  PetscCall(VecCUDAGetArray(y, &py));
  B2KCall(b2k_multvec(ctx, COL(X, d, X->l), X->ld, X->n, k > 0 ? (int)k : 0, alpha, beta, py, SLOT(d, 2)));
  PetscCall(VecCUDARestoreArray(y, &py));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* ---- the fused classical Gram-Schmidt pass: replaces BVOrthogonalizeCGS1 (bvorthog.c:91-132) through the gramschmidt slot
        (bvimpl.h:53, taken at bvorthog.c:134).  Sweep 1: c = V^T w and w^T w in one reduction; sweep 2: w -= V c with the next
        pass' V^T w and ||w||^2 from the same read of V; sweep 3 (refinement) launched speculatively and gated on the device by
        the criterion of bvorthog.c:180: three sweeps instead of four and ONE host synchronisation per column. ------------------ */
static PetscErrorCode BVGramSchmidt_B200(BV bv, PetscInt j, Vec v, PetscBool *which, PetscScalar *h, PetscScalar *c, PetscReal *onorm, PetscReal *norm)
{
  BV_B200         *d = (BV_B200 *)bv->data;
  b2k_ctx          ctx = B2KPetscContext();
  const PetscInt   kk = bv->nc + j;
  PetscScalar     *cc = c, *w;
  double          *hp;
  int              fused = 0;
  PetscObjectState state;

  PetscFunctionBegin;
  (void)which;
  PetscCall(BVState_B200(bv, &state));
  if (v) PetscCall(VecCUDAGetArray(v, &w));
  else w = COL(bv, d, j);
  bv->k = j;                                                    /* bvorthog.c:99 */
  if (!c) PetscCall(VecGetArray(bv->buffer, &cc));

  if (kk == 0) {
    if (onorm || norm) {
      B2KCall(b2k_comm_reduce_scope(d->comm, 1, &fused));
      B2KCall(b2k_sumsq(ctx, w, bv->ld, bv->n, 1, SLOT(d, 0)));
      PetscCall(BVFetch_B200(bv, d, SLOT(d, 0), 1, PETSC_TRUE, fused, &hp));
      if (onorm) *onorm = PetscSqrtReal(hp[0]);
      if (norm) *norm = PetscSqrtReal(hp[0]);
    }
  } else if (!v && d->pend_valid && d->pend_j == j && d->pend_state == state) {
    /* refinement pass: its coefficients came out of the previous fused sweep, its update already ran on the device */
    d->pend_valid = PETSC_FALSE;
    d->expect_refine = PETSC_TRUE;
    PetscCall(PetscArraycpy(cc, d->pend_c, kk));
    if (d->pend2_valid && d->pend2_ran) {
      if (norm) *norm = PetscSqrtReal(PetscMax(d->pend2_nrm2, 0.0));
    } else {
      B2KCall(b2k_comm_reduce_scope(d->comm, 1, &fused));
      B2KCall(b2k_gs_update_norm(ctx, d->V, bv->ld, bv->n, (int)kk, w, SLOT(d, 1), SLOT(d, 3)));
      PetscCall(BVFetch_B200(bv, d, SLOT(d, 3), 1, PETSC_TRUE, fused, &hp));
      if (norm) *norm = PetscSqrtReal(hp[0]);
    }
    d->pend2_valid = PETSC_FALSE;
    if (onorm) *onorm = PetscSqrtReal(PetscMax(d->pend_nrm2, 0.0));
    PetscCall(BV_AddCoefficients_Default(bv, j, h, c));
  } else {
    const PetscBool repeat = (!v && d->last_j == j && d->last_state == state) ? PETSC_TRUE : PETSC_FALSE;
    PetscCheck(!(d->pend_valid && d->pend2_valid && d->pend2_ran), PetscObjectComm((PetscObject)bv), PETSC_ERR_PLIB, "a column was refined on the device but the pass was never collected");
    if (d->pend_valid) d->expect_refine = PETSC_FALSE;
    if (repeat) d->expect_refine = PETSC_TRUE;
    d->pend_valid = PETSC_FALSE;
    d->last_j = v ? -1 : j; d->last_state = state;
    B2KCall(b2k_comm_reduce_scope(d->comm, 1, &fused));
    B2KCall(b2k_gs_dot(ctx, d->V, bv->ld, bv->n, (int)kk, w, SLOT(d, 0)));
    if (!fused) B2KCall(b2k_comm_allreduce_sum(d->comm, SLOT(d, 0), (int)kk + 1));
    if (!v && !repeat && (d->expect_refine || bv->orthog_ref == BV_ORTHOG_REFINE_ALWAYS)) {
      const PetscBool ifneeded = (bv->orthog_ref == BV_ORTHOG_REFINE_IFNEEDED && onorm && norm) ? PETSC_TRUE : PETSC_FALSE;
      const PetscBool spec = (ifneeded || bv->orthog_ref == BV_ORTHOG_REFINE_ALWAYS) ? PETSC_TRUE : PETSC_FALSE;
      B2KCall(b2k_gs_update_dot(ctx, d->V, bv->ld, bv->n, (int)kk, w, SLOT(d, 0), SLOT(d, 1)));
      if (!fused) B2KCall(b2k_comm_allreduce_sum(d->comm, SLOT(d, 1), (int)kk + 1));
      if (ifneeded) B2KCall(b2k_gs_update_norm_gated(ctx, d->V, bv->ld, bv->n, (int)kk, w, SLOT(d, 1), SLOT(d, 3), SLOT(d, 0) + kk, SLOT(d, 1) + kk, bv->orthog_eta));
      else if (spec) B2KCall(b2k_gs_update_norm(ctx, d->V, bv->ld, bv->n, (int)kk, w, SLOT(d, 1), SLOT(d, 3)));
      if (spec && !fused) B2KCall(b2k_comm_allreduce_sum(d->comm, SLOT(d, 3), 1));
      PetscCall(BVFetch_B200(bv, d, SLOT(d, 0), spec ? 3 * d->slot + 1 : 2 * d->slot, PETSC_FALSE, fused, &hp));
      PetscCall(PetscArraycpy(d->pend_c, hp + d->slot, kk + 1));
      d->pend_nrm2 = hp[d->slot + kk];
      d->pend_j = j; d->pend_state = state; d->pend_valid = PETSC_TRUE; d->pend2_valid = spec;
      if (spec) {
        const PetscReal on = PetscSqrtReal(PetscMax(hp[kk], 0.0)), nr = PetscSqrtReal(PetscMax(d->pend_nrm2, 0.0));
        d->pend2_ran = (!ifneeded || (nr != 0.0 && PetscAbsReal(nr) < bv->orthog_eta * PetscAbsReal(on))) ? PETSC_TRUE : PETSC_FALSE;
        d->pend2_nrm2 = hp[3 * d->slot];
      }
    } else {
      B2KCall(b2k_gs_update_norm(ctx, d->V, bv->ld, bv->n, (int)kk, w, SLOT(d, 0), SLOT(d, 0) + kk + 1));
      if (!fused) B2KCall(b2k_comm_allreduce_sum(d->comm, SLOT(d, 0) + kk + 1, 1));
      PetscCall(BVFetch_B200(bv, d, SLOT(d, 0), kk + 2, PETSC_FALSE, fused, &hp));
      d->pend_nrm2 = hp[kk + 1];
    }
    PetscCall(PetscArraycpy(cc, hp, kk));
    if (!c) cc[kk] = hp[kk];                                    /* the buffer keeps (w,w) next to the coefficients, bvorthog.c:40 */
    if (onorm || norm) {
      PetscCheck(hp[kk] > -bv->deftol, PetscObjectComm((PetscObject)bv), PETSC_ERR_FP, "The inner product is not well defined: indefinite matrix %g", (double)hp[kk]);
      if (onorm) *onorm = PetscSqrtReal(PetscMax(hp[kk], 0.0));
      if (norm) *norm = PetscSqrtReal(PetscMax(d->pend_nrm2, 0.0));
    }
    PetscCall(BV_AddCoefficients_Default(bv, j, h, c));
  }
  if (!c) PetscCall(VecRestoreArray(bv->buffer, &cc));
  if (v) PetscCall(VecCUDARestoreArray(v, &w));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* ---- scale / norms: svec.c:150-175, sveccuda.cu:164-214 ---------------------------------------------------------------- */
static PetscErrorCode BVScale_B200(BV bv, PetscInt j, PetscScalar alpha)
{
  BV_B200 *d = (BV_B200 *)bv->data;

  PetscFunctionBegin;
  if (j < 0) B2KCall(b2k_scale(B2KPetscContext(), COL(bv, d, bv->l), bv->ld, bv->n, (int)(bv->k - bv->l), alpha));
  else B2KCall(b2k_scale(B2KPetscContext(), COL(bv, d, j), bv->ld, bv->n, 1, alpha));
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode BVNorm_B200_Private(BV bv, PetscInt j, NormType type, PetscReal *val, PetscBool reduce)
{
  BV_B200       *d = (BV_B200 *)bv->data;
  b2k_ctx        ctx = B2KPetscContext();
  const double  *X = (j < 0) ? COL(bv, d, bv->l) : COL(bv, d, j);
  const PetscInt k = (j < 0) ? bv->k - bv->l : 1;
  double        *hp;
  int            fused = 0;

  PetscFunctionBegin;
  *val = 0.0;
  if (k <= 0) PetscFunctionReturn(PETSC_SUCCESS);
  if (type == NORM_2 || type == NORM_FROBENIUS) {
    if (reduce) B2KCall(b2k_comm_reduce_scope(d->comm, 1, &fused));
    B2KCall(b2k_sumsq(ctx, X, bv->ld, bv->n, (int)k, SLOT(d, 3)));
    PetscCall(BVFetch_B200(bv, d, SLOT(d, 3), 1, reduce, fused, &hp));
    *val = PetscSqrtReal(hp[0]);
  } else if (type == NORM_1) {
    if (reduce) B2KCall(b2k_comm_reduce_scope(d->comm, 1, &fused));
    B2KCall(b2k_colabssum(ctx, X, bv->ld, bv->n, (int)k, SLOT(d, 3)));
    PetscCall(BVFetch_B200(bv, d, SLOT(d, 3), k, reduce, fused, &hp));
    for (PetscInt i = 0; i < k; i++) *val = PetscMax(*val, hp[i]);
  } else if (type == NORM_INFINITY) {
    B2KCall(b2k_rowabssum_max(ctx, X, bv->ld, bv->n, (int)k, SLOT(d, 3)));
    if (reduce) B2KCall(b2k_comm_allreduce_max(d->comm, SLOT(d, 3), 1));
    PetscCall(BVFetch_B200(bv, d, SLOT(d, 3), 1, PETSC_FALSE, 0, &hp));
    *val = hp[0];
  } else SETERRQ(PetscObjectComm((PetscObject)bv), PETSC_ERR_SUP, "Requested norm not available");
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode BVNorm_B200(BV bv, PetscInt j, NormType t, PetscReal *v) { return BVNorm_B200_Private(bv, j, t, v, PETSC_TRUE); }
static PetscErrorCode BVNorm_Local_B200(BV bv, PetscInt j, NormType t, PetscReal *v) { return BVNorm_B200_Private(bv, j, t, v, PETSC_FALSE); }

static PetscErrorCode BVNormalize_B200(BV bv, PetscScalar *eigi)
{
  BV_B200 *d = (BV_B200 *)bv->data;
  b2k_ctx  ctx = B2KPetscContext();
  double  *hp;

  PetscFunctionBegin;
  for (PetscInt i = bv->l; i < bv->k; i++) {
    const PetscInt cols = (eigi && eigi[i] != 0.0 && i + 1 < bv->k) ? 2 : 1;       /* conjugate pair = two real columns */
    int            fused = 0;
    B2KCall(b2k_comm_reduce_scope(d->comm, 1, &fused));
    B2KCall(b2k_sumsq(ctx, COL(bv, d, i), bv->ld, bv->n, (int)cols, SLOT(d, 3)));
    PetscCall(BVFetch_B200(bv, d, SLOT(d, 3), 1, PETSC_TRUE, fused, &hp));
    if (hp[0] != 0.0 && hp[0] != 1.0) B2KCall(b2k_scale(ctx, COL(bv, d, i), bv->ld, bv->n, (int)cols, 1.0 / PetscSqrtReal(hp[0])));
    i += cols - 1;
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* ---- level-3: svec.c:17-36,54-107 / sveccuda.cu:18-124.  Q and M are sequential dense HOST Mats with their own LDA ------ */
static PetscErrorCode BVUploadBlock_B200(BV_B200 *d, Mat Q, PetscInt r0, PetscInt c0, PetscInt nr, PetscInt ncol)
{
  const PetscScalar *q;
  PetscInt           ldq;

  PetscFunctionBegin;
  PetscCheck((size_t)nr * (size_t)ncol <= d->dQ_elems, PETSC_COMM_SELF, PETSC_ERR_ARG_SIZ, "dense block larger than the scratch area");
  PetscCall(MatDenseGetLDA(Q, &ldq));
  PetscCall(MatDenseGetArrayRead(Q, &q));
  for (PetscInt cidx = 0; cidx < ncol; cidx++) PetscCall(PetscArraycpy(d->hQ + (size_t)cidx * (size_t)nr, q + (size_t)(c0 + cidx) * (size_t)ldq + r0, nr));
  PetscCall(MatDenseRestoreArrayRead(Q, &q));
  if (nr * ncol > 0) B2KCall(b2k_h2d_async(B2KPetscContext(), d->dQ, d->hQ, sizeof(double) * (size_t)nr * (size_t)ncol));
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode BVMult_B200(BV Y, PetscScalar alpha, PetscScalar beta, BV X, Mat Q)
{
  BV_B200       *y = (BV_B200 *)Y->data, *x = (BV_B200 *)X->data;
  const PetscInt ky = Y->k - Y->l, kx = X->k - X->l;

  PetscFunctionBegin;
  if (ky <= 0) PetscFunctionReturn(PETSC_SUCCESS);
  if (!Q) B2KCall(b2k_axpby(B2KPetscContext(), COL(Y, y, Y->l), Y->ld, COL(X, x, X->l), X->ld, Y->n, (int)ky, alpha, beta));
  else {
    PetscCall(BVUploadBlock_B200(y, Q, X->l, Y->l, kx, ky));
    B2KCall(b2k_mult(B2KPetscContext(), COL(Y, y, Y->l), Y->ld, COL(X, x, X->l), X->ld, Y->n, (int)ky, (int)kx, alpha, beta, y->dQ, kx > 0 ? (int)kx : 1));
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode BVMultInPlace_B200(BV V, Mat Q, PetscInt s, PetscInt e)
{
  BV_B200       *d = (BV_B200 *)V->data;
  const PetscInt l = V->l, k = V->k - V->l;

  PetscFunctionBegin;
  if (s >= e) PetscFunctionReturn(PETSC_SUCCESS);
  PetscCall(BVUploadBlock_B200(d, Q, l, l, k, e - l));
  B2KCall(b2k_mult_inplace(B2KPetscContext(), COL(V, d, l), V->ld, V->n, (int)k, (int)(s - l), (int)(e - l), d->dQ, k > 0 ? (int)k : 1, 0));   /* in place: no lda x (e-s) workspace */
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode BVMultInPlaceHermitianTranspose_B200(BV V, Mat Q, PetscInt s, PetscInt e)
{
  BV_B200       *d = (BV_B200 *)V->data;
  const PetscInt l = V->l, k = V->k - V->l;

  PetscFunctionBegin;
  if (s >= e) PetscFunctionReturn(PETSC_SUCCESS);
  PetscCall(BVUploadBlock_B200(d, Q, l, l, e - l, k));
  B2KCall(b2k_mult_inplace(B2KPetscContext(), COL(V, d, l), V->ld, V->n, (int)k, (int)(s - l), (int)(e - l), d->dQ, (int)(e - l), 1));
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode BVDot_B200(BV X, BV Y, Mat M)
{
  BV_B200       *x = (BV_B200 *)X->data, *y = (BV_B200 *)Y->data;
  b2k_ctx        ctx = B2KPetscContext();
  const PetscInt ky = Y->k - Y->l, kx = X->k - X->l;
  PetscScalar   *m;
  PetscInt       ldm;

  PetscFunctionBegin;
  PetscCheck((size_t)ky * (size_t)kx <= x->dQ_elems, PETSC_COMM_SELF, PETSC_ERR_ARG_SIZ, "result block larger than the scratch area");
  B2KCall(b2k_dot(ctx, COL(Y, y, Y->l), Y->ld, COL(X, x, X->l), X->ld, X->n, (int)ky, (int)kx, x->dQ, (int)ky));   /* one sweep, FP64 tensor cores */
  B2KCall(b2k_comm_allreduce_sum(x->comm, x->dQ, (int)(ky * kx)));
  B2KCall(b2k_d2h(ctx, x->hQ, x->dQ, sizeof(double) * (size_t)ky * (size_t)kx));
  PetscCall(MatDenseGetLDA(M, &ldm));
  PetscCall(MatDenseGetArray(M, &m));
  for (PetscInt jx = 0; jx < kx; jx++) PetscCall(PetscArraycpy(m + (size_t)(X->l + jx) * (size_t)ldm + Y->l, x->hQ + (size_t)jx * (size_t)ky, ky));   /* svec.c:101 */
  PetscCall(MatDenseRestoreArray(M, &m));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* column loop of sveccuda.cu:269-303 through the lent column Vecs (the operator may be any Mat, including the b200 MatShell) */
static PetscErrorCode BVMatMult_B200(BV V, Mat A, BV W)
{
  Vec vv, ww;

  PetscFunctionBegin;
  for (PetscInt j = 0; j < V->k - V->l; j++) {
    PetscCall(BVGetColumn(V, V->l + j, &vv));
    PetscCall(BVGetColumn(W, W->l + j, &ww));
    PetscCall(MatMult(A, vv, ww));
    PetscCall(BVRestoreColumn(V, V->l + j, &vv));
    PetscCall(BVRestoreColumn(W, W->l + j, &ww));
  }
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode BVCopy_B200(BV V, BV W)
{
  BV_B200 *v = (BV_B200 *)V->data, *w = (BV_B200 *)W->data;

  PetscFunctionBegin;
  B2KCall(b2k_copy(B2KPetscContext(), COL(W, w, W->l), W->ld, COL(V, v, V->l), V->ld, V->n, (int)(V->k - V->l)));
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode BVCopyColumn_B200(BV V, PetscInt j, PetscInt i)
{
  BV_B200 *d = (BV_B200 *)V->data;

  PetscFunctionBegin;
  B2KCall(b2k_copy(B2KPetscContext(), COL(V, d, i), V->ld, COL(V, d, j), V->ld, V->n, 1));
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode BVAllocScratch_B200(BV bv, BV_B200 *d, PetscInt m)
{
  b2k_ctx        ctx = B2KPetscContext();
  const PetscInt cols = bv->nc + m;

  PetscFunctionBegin;
  PetscCheck(cols <= B2K_MAX_COLUMNS, PetscObjectComm((PetscObject)bv), PETSC_ERR_SUP, "BV type b200 holds at most %d columns, %" PetscInt_FMT " requested", B2K_MAX_COLUMNS, cols);
  d->slot = (cols + 9) & ~1;
  B2KCall(b2k_malloc(ctx, (void **)&d->dco, sizeof(double) * 4 * (size_t)d->slot));
  B2KCall(b2k_memset0(ctx, d->dco, sizeof(double) * 4 * (size_t)d->slot));
  B2KCall(b2k_host_alloc((void **)&d->hco, sizeof(double) * 4 * (size_t)d->slot));
  d->dQ_elems = (size_t)cols * (size_t)cols + 16;
  B2KCall(b2k_malloc(ctx, (void **)&d->dQ, sizeof(double) * d->dQ_elems));
  PetscCall(PetscMalloc1(d->dQ_elems, &d->hQ));
  PetscCall(PetscMalloc1(cols + 2, &d->pend_c));
  d->pend_valid = d->pend2_valid = PETSC_FALSE;
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode BVFreeScratch_B200(BV_B200 *d)
{
  b2k_ctx ctx = B2KPetscContext();

  PetscFunctionBegin;
  B2KCall(b2k_free(ctx, d->dco));
  B2KCall(b2k_free(ctx, d->dQ));
  B2KCall(b2k_host_free(d->hco));
  PetscCall(PetscFree(d->hQ));
  PetscCall(PetscFree(d->pend_c));
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode BVResize_B200(BV bv, PetscInt m, PetscBool copy)
{
  BV_B200     *d = (BV_B200 *)bv->data;
  b2k_ctx      ctx = B2KPetscContext();
  double      *Vnew = NULL;
  const size_t bytes = sizeof(double) * (size_t)(bv->nc + m) * (size_t)bv->ld;

  PetscFunctionBegin;
  B2KCall(b2k_malloc(ctx, (void **)&Vnew, bytes));
  B2KCall(b2k_memset0(ctx, Vnew, bytes));
  if (copy) B2KCall(b2k_copy(ctx, Vnew, bv->ld, d->V, bv->ld, bv->n, (int)(bv->nc + PetscMin(m, bv->m))));
  B2KCall(b2k_free(ctx, d->V));
  d->V = Vnew;
  PetscCall(BVFreeScratch_B200(d));
  PetscCall(BVAllocScratch_B200(bv, d, m));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* the front-end has chosen the slot l (BVAvailableVec, bvimpl.h:235) and recorded bv->ci[l] = j: svec.c:292-315 */
static PetscErrorCode BVGetColumn_B200(BV bv, PetscInt j, Vec *v)
{
  BV_B200  *d = (BV_B200 *)bv->data;
  const int l = (bv->ci[0] == bv->nc + j) ? 0 : 1;

  PetscFunctionBegin;
  (void)v;
  PetscCall(VecCUDAPlaceArray(bv->cv[l], COL(bv, d, j)));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode BVRestoreColumn_B200(BV bv, PetscInt j, Vec *v)
{
  const int l = (bv->ci[0] == bv->nc + j) ? 0 : 1;

  PetscFunctionBegin;
  (void)v;
  PetscCall(VecCUDAResetArray(bv->cv[l]));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* BVGetArray hands out a HOST pointer in the reference's contract (svec.c:317-351: VecGetArray): a host mirror is filled on
   Get and written back on Restore — correct for the callers that use it (BVOrthogonalize TSQR panels, user code), never on the
   Krylov path */
static PetscErrorCode BVGetArray_B200(BV bv, PetscScalar **a)
{
  BV_B200     *d = (BV_B200 *)bv->data;
  const size_t n = (size_t)(bv->nc + bv->m) * (size_t)bv->ld;

  PetscFunctionBegin;
  PetscCall(PetscMalloc1(n, a));
  B2KCall(b2k_d2h(B2KPetscContext(), *a, d->V, sizeof(double) * n));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode BVRestoreArray_B200(BV bv, PetscScalar **a)
{
  BV_B200     *d = (BV_B200 *)bv->data;
  const size_t n = (size_t)(bv->nc + bv->m) * (size_t)bv->ld;

  PetscFunctionBegin;
  B2KCall(b2k_h2d(B2KPetscContext(), d->V, *a, sizeof(double) * n));
  PetscCall(PetscFree(*a));
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode BVGetArrayRead_B200(BV bv, const PetscScalar **a) { return BVGetArray_B200(bv, (PetscScalar **)a); }
static PetscErrorCode BVRestoreArrayRead_B200(BV bv, const PetscScalar **a)
{
  PetscScalar *p = (PetscScalar *)*a;

  PetscFunctionBegin;
  (void)bv;
  PetscCall(PetscFree(p));
  *a = NULL;
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* BVGetMat (bvbasic.c:1234-1290 calls the slot when it exists): a MATDENSECUDA that WRAPS the active columns in place */
static PetscErrorCode BVGetMat_B200(BV bv, Mat *A)
{
  BV_B200       *d = (BV_B200 *)bv->data;
  const PetscInt m = bv->k - bv->l;

  PetscFunctionBegin;
  PetscCall(MatCreateDenseCUDA(PetscObjectComm((PetscObject)bv), bv->n, PETSC_DECIDE, bv->N, m, COL(bv, d, bv->l), &bv->Aget));
  *A = bv->Aget;
  PetscFunctionReturn(PETSC_SUCCESS);
}
static PetscErrorCode BVRestoreMat_B200(BV bv, Mat *A)
{
  PetscFunctionBegin;
  (void)A;
  PetscCall(MatDestroy(&bv->Aget));                              /* the wrapped columns were modified in place: nothing to copy */
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode BVView_B200(BV bv, PetscViewer viewer)
{
  PetscScalar *a;

  PetscFunctionBegin;
  PetscCall(BVGetArray_B200(bv, &a));
  for (PetscInt j = 0; j < bv->nc + bv->m; j++) {
    PetscCall(PetscViewerASCIIPrintf(viewer, "column %" PetscInt_FMT "\n", j - bv->nc));
    for (PetscInt i = 0; i < bv->n; i++) PetscCall(PetscViewerASCIIPrintf(viewer, "%18.16e\n", (double)a[(size_t)j * (size_t)bv->ld + i]));
  }
  PetscCall(PetscFree(a));
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode BVDestroy_B200(BV bv)
{
  BV_B200 *d = (BV_B200 *)bv->data;

  PetscFunctionBegin;
  if (d->owns_V) B2KCall(b2k_free(B2KPetscContext(), d->V));
  PetscCall(BVFreeScratch_B200(d));
  PetscCall(VecDestroy(&bv->cv[0]));
  PetscCall(VecDestroy(&bv->cv[1]));
  PetscCall(PetscFree(bv->data));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* BVDuplicate: the new BV gets its own storage from its own constructor call; only the orthogonalisation state is fresh */
static PetscErrorCode BVDuplicate_B200(BV V, BV W)
{
  PetscFunctionBegin;
  (void)V;
  ((BV_B200 *)W->data)->expect_refine = PETSC_TRUE;
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* the constructor: what BVCreate_Svec does at svec.c:397-563 */
PETSC_EXTERN PetscErrorCode BVCreate_B200(BV bv)
{
  BV_B200  *d;
  b2k_ctx   ctx;
  PetscBool iscuda;
  PetscInt  tglobal = 0;

  PetscFunctionBegin;
  PetscCall(B2KPetscInitialize(PetscObjectComm((PetscObject)bv)));
  ctx = B2KPetscContext();
  PetscCall(PetscNew(&d));
  bv->data = (void *)d;
  PetscCall(PetscStrcmp(bv->vtype, VECCUDA, &iscuda));
  if (!iscuda) PetscCall(PetscStrcmp(bv->vtype, VECMPICUDA, &iscuda));
  if (!iscuda) PetscCall(PetscStrcmp(bv->vtype, VECSEQCUDA, &iscuda));
  PetscCheck(iscuda, PetscObjectComm((PetscObject)bv), PETSC_ERR_SUP, "BV type b200 needs CUDA vectors (-vec_type cuda / MatShellSetVecType): got %s", bv->vtype);
  bv->cuda = PETSC_FALSE;                                        /* coefficients live in a host buffer (the *_Default helpers); the kernels are ours */
  PetscCall(BV_SetDefaultLD(bv, bv->n));
  if (bv->ld & 1) bv->ld++;                                      /* 16-byte aligned columns */
  PetscCall(PetscIntMultError(bv->m, bv->N, &tglobal));          /* svec.c:425: only the overflow check; offsets are 64-bit here */
  PetscCall(B2KPetscCommGet(PetscObjectComm((PetscObject)bv), &d->comm));
  if (PetscUnlikely(bv->issplit)) {                              /* BVGetSplit: alias the parent's columns, svec.c:427-461 */
    BV_B200 *p = (BV_B200 *)bv->splitparent->data;
    PetscCheck(bv->issplit > 0, PetscObjectComm((PetscObject)bv), PETSC_ERR_SUP, "BVGetSplitRows() is not available for type b200");
    d->V = (bv->issplit == 1) ? p->V : p->V + (size_t)bv->splitparent->lsplit * (size_t)bv->ld;
    d->owns_V = PETSC_FALSE;
  } else {
    const size_t bytes = sizeof(double) * (size_t)(bv->nc + bv->m) * (size_t)bv->ld;
    B2KCall(b2k_malloc(ctx, (void **)&d->V, bytes));
    B2KCall(b2k_memset0(ctx, d->V, bytes));
    d->owns_V = PETSC_TRUE;
  }
  if (PetscUnlikely(bv->Acreate)) {                              /* BVCreateFromMat: svec.c:473-484 */
    const PetscScalar *aa;
    PetscInt           lda;
    PetscCall(MatDenseGetArrayRead(bv->Acreate, &aa));
    PetscCall(MatDenseGetLDA(bv->Acreate, &lda));
    for (PetscInt j = 0; j < bv->m; j++) B2KCall(b2k_h2d(ctx, COL(bv, d, j - bv->nc), aa + (size_t)j * (size_t)lda, sizeof(double) * (size_t)bv->n));
    PetscCall(MatDenseRestoreArrayRead(bv->Acreate, &aa));
    PetscCall(MatDestroy(&bv->Acreate));
  }
  PetscCall(BVAllocScratch_B200(bv, d, bv->m));
  d->expect_refine = PETSC_TRUE;
  d->last_j = -1;
  PetscCall(BVCreateVecEmpty(bv, &bv->cv[0]));                  /* svec.c:486-487 */
  PetscCall(BVCreateVecEmpty(bv, &bv->cv[1]));

  /* all 35 slots of struct _BVOps (bvimpl.h:24-61) */
  bv->ops->mult             = BVMult_B200;
  bv->ops->multvec          = BVMultVec_B200;
  bv->ops->multinplace      = BVMultInPlace_B200;
  bv->ops->multinplacetrans = BVMultInPlaceHermitianTranspose_B200;
  bv->ops->dot              = BVDot_B200;
  bv->ops->dotvec           = BVDotVec_B200;
  bv->ops->dotvec_local     = BVDotVec_Local_B200;
  bv->ops->dotvec_begin     = NULL;   /* the front-end's PetscSplitReduction path over dotvec_local is used (bvglobal.c:205-222); the   */
  bv->ops->dotvec_end       = NULL;   /* stand-alone host layer shows the fused device queue these slots can carry (bvb200.c)          */
  bv->ops->scale            = BVScale_B200;
  bv->ops->norm             = BVNorm_B200;
  bv->ops->norm_local       = BVNorm_Local_B200;
  bv->ops->norm_begin       = NULL;   /* as dotvec_begin/end: bvglobal.c:727-741                                                       */
  bv->ops->norm_end         = NULL;
  bv->ops->normalize        = BVNormalize_B200;
  bv->ops->matmult          = BVMatMult_B200;
  bv->ops->copy             = BVCopy_B200;
  bv->ops->copycolumn       = BVCopyColumn_B200;
  bv->ops->resize           = BVResize_B200;
  bv->ops->getcolumn        = BVGetColumn_B200;
  bv->ops->restorecolumn    = BVRestoreColumn_B200;
  bv->ops->getarray         = BVGetArray_B200;
  bv->ops->restorearray     = BVRestoreArray_B200;
  bv->ops->getarrayread     = BVGetArrayRead_B200;
  bv->ops->restorearrayread = BVRestoreArrayRead_B200;
  bv->ops->restoresplit     = NULL;   /* split children alias the parent's device columns: nothing to copy back (svec.c has none either)  */
  bv->ops->restoresplitrows = NULL;   /* BVGetSplitRows is refused in the constructor                                                     */
  bv->ops->gramschmidt      = BVGramSchmidt_B200;
  bv->ops->getmat           = BVGetMat_B200;
  bv->ops->restoremat       = BVRestoreMat_B200;
  bv->ops->duplicate        = BVDuplicate_B200;
  bv->ops->create           = NULL;   /* set by BVSetType to this function's address (bvbasic.c:57)                                        */
  bv->ops->setfromoptions   = NULL;   /* no type-specific options: the B2K_* environment switches of include/b2k.h apply                    */
  bv->ops->view             = BVView_B200;
  bv->ops->destroy          = BVDestroy_B200;
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* what `-dll_append libb200petsc.so` calls (PetscDLLibraryRegister_<name>, PETSc dynamic library loading) */
PETSC_EXTERN PetscErrorCode PetscDLLibraryRegister_b200petsc(void)
{
  PetscFunctionBegin;
  PetscCall(BVRegister("b200", BVCreate_B200));
  PetscCall(MatRegister("b200csr", MatCreate_B200CSR));
  PetscFunctionReturn(PETSC_SUCCESS);
}
