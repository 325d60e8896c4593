/*
 * bvb200.c — BV type "b200": the basis lives in HBM as one column-major n x (nc+m) array
 * (the layout of BVSVEC, src/sys/classes/bv/impls/svec/svec.c:397-563) and every ops-table slot is
 * a short shim over the sm_100a kernels of libb200krylov (include/b2k.h), the way
 * impls/svec/sveccuda/sveccuda.cu:18-499 shims cuBLAS.  Differences from the reference's CUDA BV:
 *   - the `gramschmidt` slot (bvimpl.h:53) is filled: one classical GS pass = ONE reduction sweep
 *     (V^T w and w^T w together) + ONE update sweep that already emits ||w_new||^2 and, when a DGKS
 *     refinement is expected, the next pass' V^T w_new, so a refined orthogonalisation reads V three
 *     times instead of four and needs one host synchronisation per pass instead of 3 blocking 8-byte
 *     copies + 5 micro-kernels (bvcuda.cu:345-548);
 *   - coefficients are reduced across GPUs on the device (NCCL) and land in pinned host memory with a
 *     single copy; no per-call cudaMalloc (bvcuda.cu:217,231,90-92);
 *   - BVMultInPlace runs in place (no lda x (e-s) workspace + cudaMemcpy2D, bvcuda.cu:90-94).
 */
#include "b2kimpl.h"

typedef struct {
  double  *V;            /* device: (nc+m) columns of ld doubles                                  */
  double  *dco;          /* device coefficient staging: 4 slots of `slot` doubles                 */
  double  *hco;          /* pinned host mirror of dco                                             */
  PetscInt slot;
  double  *dQ;           /* device scratch for small dense blocks (Q, M)                          */
  size_t   dQ_elems;
  double  *hQ;           /* host staging for compacted Q blocks                                   */
  /* second-pass cache filled by the fused update sweep */
  PetscBool pend_valid;
  PetscInt  pend_j;
  uint64_t  pend_state;
  PetscReal pend_nrm2;
  double   *pend_c;      /* host copy of V^T w_new                                                */
  int       fuse_mode;   /* 0 never, 1 always, 2 adaptive (env B2K_BV_FUSE)                       */
  PetscBool expect_refine;
  PetscInt  last_j;      /* column and state of the last uncached pass: a repeat means DGKS refined  */
  uint64_t  last_state;
  /* one synchronisation per column: the update sweep of the refinement pass is launched speculatively behind the first pass,
     gated on the device by the DGKS criterion (b2k_gs_update_norm_gated); the host takes the same decision from the same
     numbers, so it knows whether the pass ran */
  int       onesync;     /* env B2K_BV_ONESYNC, default 1                                         */
  PetscBool pend2_valid; /* the second pass was issued with the first                              */
  PetscBool pend2_ran;   /* ... and the criterion let it run                                       */
  PetscReal pend2_nrm2;  /* ||w||^2 after it                                                       */
  /* TSQR between tsqr_factor(wantq) and tsqr_formq: reflector scalars in HBM, the factored stack of the CTAs' triangles on the host */
  double   *ts_coef;     /* device */
  double   *ts_stack;    /* host: (nblk*nk) x nk factored by dgeqrf, then nk tau                    */
  PetscInt  ts_nk, ts_nblk;
  /* asynchronous Krylov cycle (krylov_steps): 4 coefficient slots PER STEP in HBM and their pinned mirror */
  int       async;       /* env B2K_BV_ASYNC, default 1                                           */
  double   *kr_d, *kr_h;
  PetscInt  kr_steps;
  PetscBool kr_probed;
} BV_B200;

#define CTX() B2KGetContext()
#define COL(bv, d, j) ((d)->V + (size_t)((bv)->nc + (j)) * (size_t)(bv)->ld)
#define SLOT(d, i) ((d)->dco + (size_t)(i) * (size_t)(d)->slot)
#define HSLOT(d, i) ((d)->hco + (size_t)(i) * (size_t)(d)->slot)

static PetscErrorCode BVAllocScratch_B200(BV bv, BV_B200 *d)
{
  b2k_ctx ctx = CTX();
  const PetscInt cols = bv->nc + bv->m;
  d->slot = (cols + 9) & ~1;
  B2KCall(b2k_malloc(ctx, (void **)&d->dco, sizeof(double) * 4 * (size_t)d->slot));
  B2KCall(b2k_memset0(ctx, d->dco, sizeof(double) * 4 * (size_t)d->slot));
  B2KCall(b2k_host_alloc((void **)&d->hco, sizeof(double) * 4 * (size_t)d->slot));
  d->dQ_elems = (size_t)cols * (size_t)cols + 16;
  B2KCall(b2k_malloc(ctx, (void **)&d->dQ, sizeof(double) * d->dQ_elems));
  d->hQ = (double *)malloc(sizeof(double) * d->dQ_elems);
  d->pend_c = (double *)calloc((size_t)cols + 2, sizeof(double));
  PetscCheck(d->hQ && d->pend_c, PETSC_ERR_MEM, "out of memory");
  d->pend_valid = PETSC_FALSE;
  return PETSC_SUCCESS;
}

static PetscErrorCode BVFreeScratch_B200(BV_B200 *d)
{
  b2k_ctx ctx = CTX();
  if (ctx) { b2k_free(ctx, d->dco); b2k_free(ctx, d->dQ); }
  b2k_host_free(d->hco);
  free(d->hQ); free(d->pend_c);
  d->dco = d->dQ = d->hco = d->hQ = d->pend_c = NULL;
  return PETSC_SUCCESS;
}

/* Reductions over the ranks: when the communicator has peer-memory mailboxes (B2KCommP2POpen) the second stage of the
   kernels' own reduction already sums over the GPUs (k_reduce_partials_xg) and the all-reduce below is skipped.
   BVScope_B200(bv, reduce, &fused) opens the scope before the kernels are launched; BVFetch_B200 closes it. */
static PetscErrorCode BVScope_B200(BV bv, PetscBool reduce, PetscBool *fused)
{
  *fused = PETSC_FALSE;
  if (reduce) PetscCall(B2KCommReduceScope(bv->comm, PETSC_TRUE, fused));
  return PETSC_SUCCESS;
}

/* sum over ranks of `count` doubles in slot memory (unless the kernels did it), then bring them to the pinned mirror and wait */
static PetscErrorCode BVFetch_B200(BV bv, BV_B200 *d, double *dptr, PetscInt count, PetscBool reduce, PetscBool fused, double **hptr)
{
  b2k_ctx ctx = CTX();
  if (fused) PetscCall(B2KCommReduceScope(bv->comm, PETSC_FALSE, NULL));
  else if (reduce) PetscCall(B2KCommAllreduce(bv->comm, dptr, count, 0, B2K_MEM_DEVICE));
  double *hp = d->hco + (dptr - d->dco);
  B2KCall(b2k_d2h_async(ctx, hp, dptr, sizeof(double) * (size_t)count));
  B2KCall(b2k_ctx_sync(ctx));
  if (fused) {
    int bad = 0;
    B2KCall(b2k_comm_p2p_error(bv->comm->nccl, &bad));
    PetscCheck(!bad, PETSC_ERR_LIB, "peer-memory reduction timed out: a rank did not take part in the collective");
  }
  *hptr = hp;
  return PETSC_SUCCESS;
}

/* ---- dotvec / multvec: svec.c:38-52,109-129, sveccuda.cu:45-69,126-143 --------------------------- */
static PetscErrorCode BVDotVec_B200_Private(BV X, Vec y, PetscScalar *q, PetscBool reduce)
{
  BV_B200 *d = (BV_B200 *)X->data;
  const PetscInt k = X->k - X->l;
  PetscScalar *qq = q ? q : X->buffer;
  if (k <= 0) return PETSC_SUCCESS;
  PetscCheck(y->mem == B2K_MEM_DEVICE, PETSC_ERR_ARG_INCOMP, "BV type b200 needs device vectors");
  PetscBool fused;
  PetscCall(BVScope_B200(X, reduce, &fused));
  B2KCall(b2k_dotvec(CTX(), COL(X, d, X->l), X->ld, X->n, k, y->array, SLOT(d, 0)));
  double *hp;
  PetscCall(BVFetch_B200(X, d, SLOT(d, 0), k, reduce, fused, &hp));
  memcpy(qq, hp, sizeof(double) * (size_t)k);
  return PETSC_SUCCESS;
}
static PetscErrorCode BVDotVec_B200(BV X, Vec y, PetscScalar *q) { return BVDotVec_B200_Private(X, y, q, PETSC_TRUE); }
static PetscErrorCode BVDotVec_Local_B200(BV X, Vec y, PetscScalar *q) { return BVDotVec_B200_Private(X, y, q, PETSC_FALSE); }

static PetscErrorCode BVMultVec_B200(BV X, PetscScalar alpha, PetscScalar beta, Vec y, PetscScalar *q)
{
  BV_B200 *d = (BV_B200 *)X->data;
  b2k_ctx ctx = CTX();
  const PetscInt k = X->k - X->l;
  const PetscScalar *qq = q ? q : X->buffer;
  PetscCheck(y->mem == B2K_MEM_DEVICE, PETSC_ERR_ARG_INCOMP, "BV type b200 needs device vectors");
  if (k > 0) B2KCall(b2k_h2d_async(ctx, SLOT(d, 2), qq, sizeof(double) * (size_t)k));   /* pageable source: staged before return */
  B2KCall(b2k_multvec(ctx, COL(X, d, X->l), X->ld, X->n, k > 0 ? k : 0, alpha, beta, y->array, SLOT(d, 2)));
  return PETSC_SUCCESS;
}

/* ---- the fused Gram-Schmidt pass: replaces BVOrthogonalizeCGS1 (bvorthog.c:91-132) ------------------ */
static PetscErrorCode BVGramSchmidt_B200(BV bv, PetscInt j, Vec v, PetscBool *which, PetscScalar *h, PetscScalar *c, PetscReal *onorm,
                                         PetscReal *norm)
{
  (void)which;
  BV_B200 *d = (BV_B200 *)bv->data;
  b2k_ctx ctx = CTX();
  const PetscInt kk = bv->nc + j;                 /* columns to project against: physical 0..kk-1 (l = -nc) */
  PetscScalar *cc = c ? c : bv->buffer;
  double *w, *hp;
  PetscBool fused = PETSC_FALSE;
  if (v) { PetscCheck(v->mem == B2K_MEM_DEVICE, PETSC_ERR_ARG_INCOMP, "BV type b200 needs device vectors"); w = v->array; }
  else w = COL(bv, d, j);
  bv->k = j;                                      /* as bvorthog.c:99 */

  if (kk == 0) {                                  /* nothing to project: only the norm */
    if (onorm || norm) {
      PetscReal beta;
      PetscCall(BVScope_B200(bv, PETSC_TRUE, &fused));
      B2KCall(b2k_sumsq(ctx, w, bv->ld, bv->n, 1, SLOT(d, 0)));
      PetscCall(BVFetch_B200(bv, d, SLOT(d, 0), 1, PETSC_TRUE, fused, &hp));
      beta = sqrt(hp[0]);
      if (onorm) *onorm = beta;
      if (norm) *norm = beta;
    }
    return PETSC_SUCCESS;
  }

  if (!v && d->pend_valid && d->pend_j == j && d->pend_state == bv->state) {
    /* refinement pass: V^T w and ||w||^2 of the CURRENT w came out of the previous fused sweep */
    d->pend_valid = PETSC_FALSE;
    d->expect_refine = PETSC_TRUE;
    memcpy(cc, d->pend_c, sizeof(double) * (size_t)kk);
    const PetscReal beta2 = d->pend_nrm2;
    const PetscBool need_norm = norm ? PETSC_TRUE : PETSC_FALSE;
    if (d->pend2_valid && d->pend2_ran) {
      /* the update sweep of this pass already ran behind the first one (gated on the device): nothing to launch, nothing to wait for */
      d->pend2_valid = PETSC_FALSE;
      if (norm) *norm = sqrt(d->pend2_nrm2 > 0.0 ? d->pend2_nrm2 : 0.0);
    } else if (need_norm) {
      d->pend2_valid = PETSC_FALSE;
      PetscCall(BVScope_B200(bv, PETSC_TRUE, &fused));
      B2KCall(b2k_gs_update_norm(ctx, d->V, bv->ld, bv->n, kk, w, SLOT(d, 1), SLOT(d, 3)));   /* c2 still sits in slot 1 */
      PetscCall(BVFetch_B200(bv, d, SLOT(d, 3), 1, PETSC_TRUE, fused, &hp));
      *norm = sqrt(hp[0]);
    } else {
      d->pend2_valid = PETSC_FALSE;
      B2KCall(b2k_multvec(ctx, d->V, bv->ld, bv->n, kk, -1.0, 1.0, w, SLOT(d, 1)));
    }
    if (onorm) *onorm = sqrt(beta2 > 0.0 ? beta2 : 0.0);
    BV_AddCoefficients(bv, j, h, c);
    return PETSC_SUCCESS;
  }
  PetscCheck(!(d->pend_valid && d->pend2_valid && d->pend2_ran), PETSC_ERR_PLIB, "column %d was refined on the device but the host never collected the pass", (int)d->pend_j);
  if (d->pend_valid) d->expect_refine = PETSC_FALSE;   /* the cached second pass was never asked for */
  d->pend_valid = PETSC_FALSE;
  const PetscBool repeat = (!v && d->last_j == j && d->last_state == bv->state) ? PETSC_TRUE : PETSC_FALSE;
  if (repeat) d->expect_refine = PETSC_TRUE;           /* refinement happened although we did not prepare for it */
  d->last_j = v ? -1 : j; d->last_state = bv->state;

  /* sweep 1: c = V^T w and w^T w in one reduction (BVDotColumnInc bvorthog.c:32-47) */
  PetscCall(BVScope_B200(bv, PETSC_TRUE, &fused));
  B2KCall(b2k_gs_dot(ctx, d->V, bv->ld, bv->n, kk, w, SLOT(d, 0)));
  if (!fused) PetscCall(B2KCommAllreduce(bv->comm, SLOT(d, 0), kk + 1, 0, B2K_MEM_DEVICE));
  const PetscBool fuse = (!v && !repeat && (d->fuse_mode == 1 || (d->fuse_mode == 2 && (d->expect_refine || bv->orthog_ref == BV_ORTHOG_REFINE_ALWAYS)))) ? PETSC_TRUE : PETSC_FALSE;
  if (fuse) {
    /* sweep 2: w -= V c, and from the same read of V the next pass' V^T w_new and ||w_new||^2 */
    B2KCall(b2k_gs_update_dot(ctx, d->V, bv->ld, bv->n, kk, w, SLOT(d, 0), SLOT(d, 1)));
    if (!fused) PetscCall(B2KCommAllreduce(bv->comm, SLOT(d, 1), kk + 1, 0, B2K_MEM_DEVICE));
    /* sweep 3, speculative: the refinement's update w -= V c2 (+ ||w||^2).  REFINE_IFNEEDED: gated on the device by the
       criterion of bvorthog.c:180 evaluated from (w^T w, ||w_1||^2) = (slot 0, slot 1)[kk]; REFINE_ALWAYS: unconditional */
    const PetscBool ifneeded = (bv->orthog_ref == BV_ORTHOG_REFINE_IFNEEDED && onorm && norm) ? PETSC_TRUE : PETSC_FALSE;
    const PetscBool spec = (d->onesync && (ifneeded || bv->orthog_ref == BV_ORTHOG_REFINE_ALWAYS)) ? PETSC_TRUE : PETSC_FALSE;
    if (spec) {
      if (ifneeded) B2KCall(b2k_gs_update_norm_gated(ctx, d->V, bv->ld, bv->n, kk, w, SLOT(d, 1), SLOT(d, 3), SLOT(d, 0) + kk, SLOT(d, 1) + kk, bv->orthog_eta));
      else B2KCall(b2k_gs_update_norm(ctx, d->V, bv->ld, bv->n, kk, w, SLOT(d, 1), SLOT(d, 3)));
      if (!fused) PetscCall(B2KCommAllreduce(bv->comm, SLOT(d, 3), 1, 0, B2K_MEM_DEVICE));
    }
    PetscCall(BVFetch_B200(bv, d, SLOT(d, 0), spec ? 3 * d->slot + 1 : 2 * d->slot, PETSC_FALSE, fused, &hp));
    memcpy(d->pend_c, hp + d->slot, sizeof(double) * (size_t)(kk + 1));
    d->pend_nrm2 = hp[d->slot + kk];
    d->pend_j = j; d->pend_state = bv->state; d->pend_valid = PETSC_TRUE;
    d->pend2_valid = spec;
    if (spec) {
      const PetscReal on = sqrt(hp[kk] > 0.0 ? hp[kk] : 0.0), nr = sqrt(d->pend_nrm2 > 0.0 ? d->pend_nrm2 : 0.0);
      d->pend2_ran = (!ifneeded || (nr != 0.0 && fabs(nr) < bv->orthog_eta * fabs(on))) ? PETSC_TRUE : PETSC_FALSE;   /* the gate, bit for bit */
      d->pend2_nrm2 = hp[3 * d->slot];
    }
  } else {
    /* sweep 2: w -= V c with the explicit ||w_new||^2 folded in */
    B2KCall(b2k_gs_update_norm(ctx, d->V, bv->ld, bv->n, kk, w, SLOT(d, 0), SLOT(d, 0) + kk + 1));
    if (!fused) PetscCall(B2KCommAllreduce(bv->comm, SLOT(d, 0) + kk + 1, 1, 0, B2K_MEM_DEVICE));
    PetscCall(BVFetch_B200(bv, d, SLOT(d, 0), kk + 2, PETSC_FALSE, fused, &hp));
    d->pend_nrm2 = hp[kk + 1];
  }
  memcpy(cc, hp, sizeof(double) * (size_t)kk);
  if (!c) cc[kk] = hp[kk];                        /* the buffer keeps (w,w) next to the coefficients, bvorthog.c:40 */
  if (onorm || norm) {
    const PetscReal deftol = 10 * PETSC_MACHINE_EPSILON;
    PetscCheck(hp[kk] > -deftol, PETSC_ERR_FP, "The inner product is not well defined: indefinite matrix %g", hp[kk]);   /* BV_SafeSqrt */
    if (onorm) *onorm = sqrt(hp[kk] > 0.0 ? hp[kk] : 0.0);
    if (norm) *norm = sqrt(d->pend_nrm2 > 0.0 ? d->pend_nrm2 : 0.0);   /* explicit, not the beta^2 - sum c^2 estimate of :124 */
  }
  BV_AddCoefficients(bv, j, h, c);
  return PETSC_SUCCESS;
}

/* ---- a whole Krylov cycle without host round trips --------------------------------------------------------------------------
   BVMatLanczos / BVMatArnoldi (bvkrylov.c:56-226) call, per column, MatMult, then BVOrthonormalizeColumn whose DGKS loop
   (bvorthog.c:174-203) and final scaling (:417-422) need beta, the norms and the refinement decision ON THE HOST: one blocking
   round trip per column at best.  Here the steps k..m-1 are enqueued back to back: SpMV, dot sweep, update+dot sweep, the
   refinement's update sweep gated on the device by the same criterion, and the normalisation by 1/sqrt of whichever norm the gate
   selected (b2k_scale_rsqrt_gated) — every step writes its coefficients into its own slots, the host copies all of them once and
   rebuilds the coefficient buffer (h = c1 [+ c2], norm, passes) with the reference's arithmetic.  The launch queue never drains,
   so the per-kernel launch latency and the synchronisation gap disappear from the latency-bound regime (small local row counts,
   strong scaling).  A step that needs what the device did not do (third pass, breakdown, zero norm) ends the asynchronous part:
   everything before it is final, the step-by-step loop of bv.c redoes it. */
static PetscErrorCode BVKrylovSteps_B200(BV V, Mat A, PetscInt k, PetscInt m, PetscInt *jnext, PetscReal *beta)
{
  BV_B200 *d = (BV_B200 *)V->data;
  b2k_ctx ctx = CTX();
  const PetscInt steps = m - k;
  *jnext = k;
  if (!d->async || !d->onesync || d->fuse_mode == 0 || steps < 2 || V->orthog_type != BV_ORTHOG_CGS || V->orthog_ref != BV_ORTHOG_REFINE_IFNEEDED ||
      V->matrix || V->nc != 0)
    return PETSC_SUCCESS;                          /* every condition is the same on all ranks: the choice is collective */
  if (d->kr_steps < steps) {
    if (d->kr_d) B2KCall(b2k_free(ctx, d->kr_d));
    if (d->kr_h) b2k_host_free(d->kr_h);
    d->kr_d = d->kr_h = NULL; d->kr_steps = 0;
    const size_t bytes = sizeof(double) * 4 * (size_t)d->slot * (size_t)V->m;
    B2KCall(b2k_malloc(ctx, (void **)&d->kr_d, bytes));
    B2KCall(b2k_host_alloc((void **)&d->kr_h, bytes));
    d->kr_steps = V->m;
  }
  d->pend_valid = PETSC_FALSE; d->pend2_valid = PETSC_FALSE; d->last_j = -1;
  PetscBool fused = PETSC_FALSE, anyfused = PETSC_FALSE;
  /* two schedules, chosen like the step-by-step path chooses (expect_refine follows what DGKS did last): `spec` prepares the
     refinement pass speculatively (3 sweeps enqueued, the third gated); otherwise 2 sweeps (dot, update + norm) and a step that
     turns out to need a refinement ends the asynchronous part there (Arnoldi on the Markov matrix refines once in 3 500 steps) */
  const PetscBool spec = (d->fuse_mode == 1 || d->expect_refine) ? PETSC_TRUE : PETSC_FALSE;
  if (!spec && !d->kr_probed && m > k + 3) m = k + 3;        /* nothing is known about this recurrence yet: a short probe, not a whole cycle */
  d->kr_probed = PETSC_TRUE;
  for (PetscInt j = k; j < m; j++) {
    PetscCall(BVMatMultColumn(V, A, j));
    const PetscInt kk = j + 1;
    double *w = COL(V, d, kk), *S0 = d->kr_d + 4 * (size_t)d->slot * (size_t)(j - k), *S1 = S0 + d->slot, *S3 = S0 + 3 * d->slot;
    PetscCall(BVScope_B200(V, PETSC_TRUE, &fused));
    anyfused = fused ? PETSC_TRUE : anyfused;
    B2KCall(b2k_gs_dot(ctx, d->V, V->ld, V->n, (int)kk, w, S0));
    if (!fused) PetscCall(B2KCommAllreduce(V->comm, S0, (int)kk + 1, 0, B2K_MEM_DEVICE));
    if (spec) {
      B2KCall(b2k_gs_update_dot(ctx, d->V, V->ld, V->n, (int)kk, w, S0, S1));
      if (!fused) PetscCall(B2KCommAllreduce(V->comm, S1, (int)kk + 1, 0, B2K_MEM_DEVICE));
      B2KCall(b2k_gs_update_norm_gated(ctx, d->V, V->ld, V->n, (int)kk, w, S1, S3, S0 + kk, S1 + kk, V->orthog_eta));
      if (!fused) PetscCall(B2KCommAllreduce(V->comm, S3, 1, 0, B2K_MEM_DEVICE));
    } else {
      B2KCall(b2k_gs_update_norm(ctx, d->V, V->ld, V->n, (int)kk, w, S0, S1 + kk));      /* ||w_1||^2 where the other schedule leaves it */
      if (!fused) PetscCall(B2KCommAllreduce(V->comm, S1 + kk, 1, 0, B2K_MEM_DEVICE));
    }
    if (fused) PetscCall(B2KCommReduceScope(V->comm, PETSC_FALSE, NULL));
    if (spec) B2KCall(b2k_scale_rsqrt_gated(ctx, w, V->n, S0 + kk, S1 + kk, S3, V->orthog_eta));
    else B2KCall(b2k_scale_rsqrt(ctx, w, V->n, S1 + kk));
  }
  B2KCall(b2k_d2h_async(ctx, d->kr_h, d->kr_d, sizeof(double) * 4 * (size_t)d->slot * (size_t)(m - k)));
  B2KCall(b2k_ctx_sync(ctx));
  if (anyfused) {
    int bad = 0;
    B2KCall(b2k_comm_p2p_error(V->comm->nccl, &bad));
    PetscCheck(!bad, PETSC_ERR_LIB, "peer-memory reduction timed out: a rank did not take part in the collective");
  }
  const size_t ldb = (size_t)(V->nc + V->m);
  PetscInt j, nrefined = 0;
  for (j = k; j < m; j++) {
    const PetscInt kk = j + 1;
    const double *H0 = d->kr_h + 4 * (size_t)d->slot * (size_t)(j - k), *H1 = H0 + d->slot, *H3 = H0 + 3 * d->slot;
    const PetscReal deftol = 10 * PETSC_MACHINE_EPSILON;
    if (!(H0[kk] > -deftol) || !(H1[kk] > -deftol)) break;                                        /* BV_SafeSqrt: let the synchronous path raise it */
    const PetscReal onrm = sqrt(H0[kk] > 0.0 ? H0[kk] : 0.0), nrm1 = sqrt(H1[kk] > 0.0 ? H1[kk] : 0.0);
    const PetscBool refine = (nrm1 != 0.0 && fabs(nrm1) < V->orthog_eta * fabs(onrm)) ? PETSC_TRUE : PETSC_FALSE;   /* bvorthog.c:180, = the gate */
    PetscReal nrm = nrm1, on = onrm;
    nrefined += refine ? 1 : 0;
    if (refine && !spec) { d->expect_refine = PETSC_TRUE; break; }                               /* not prepared: the step-by-step loop refines it */
    if (refine) {
      if (!(H3[0] > -deftol)) break;
      nrm = sqrt(H3[0] > 0.0 ? H3[0] : 0.0); on = nrm1;
      if (nrm != 0.0 && fabs(nrm) < V->orthog_eta * fabs(on)) break;                              /* a third pass is due (l < 3) */
    }
    if (!(nrm != 0.0 && fabs(nrm) >= V->orthog_eta * fabs(on))) break;                            /* lindep / breakdown */
    PetscScalar *hcol = V->buffer + (size_t)kk * ldb;
    for (PetscInt i = 0; i < kk; i++) hcol[i] = refine ? (0.0 + H0[i]) + H1[i] : 0.0 + H0[i];    /* BV_CleanCoefficients, then h += c per pass */
    memcpy(V->buffer, refine ? H1 : H0, sizeof(double) * (size_t)(kk + 1));                        /* column 0 = c of the last pass (+ its (w,w)) */
    hcol[kk] = nrm;                                                                               /* BV_SetValue(bv,k,k,h,nrm) bvorthog.c:212 */
    V->n_gs_passes += refine ? 2 : 1;
    V->state += 2;
    *beta = nrm;
  }
  if (spec && j == m && nrefined == 0 && d->fuse_mode == 2) d->expect_refine = PETSC_FALSE;   /* a whole cycle without refinement: 2 sweeps next time */
  *jnext = j;
  return PETSC_SUCCESS;
}

/* ---- split-phase reductions: the dotvec_begin/end, norm_begin/end slots (bvimpl.h:33-34,38-39) -------------------------
   What the reference's one-sided Lanczos merges with BVNormColumnBegin + BVDotVecBegin … End (trlanczos.c:383-396): every Begin
   launches its sweep with a LOCAL reduction into one shared device queue (any BV of type b200 of this process: U and V of
   the SVD share it); the first End sums the whole queue over the GPUs in ONE short all-reduce (NVLink mailboxes, else
   ncclAllReduce), copies it to pinned memory once and synchronises once; the other Ends only pick their numbers up. */
#define SRQ_MAX 2048
static struct {
  double *d, *h;
  int     n, nent, next;
  struct { int off, cnt; void *owner; } e[64];
  PetscBool reduced;
} g_srq;

static PetscErrorCode BVSplitQueuePush_B200(BV bv, PetscInt cnt, double **slot)
{
  if (!g_srq.d) {
    B2KCall(b2k_malloc(CTX(), (void **)&g_srq.d, sizeof(double) * SRQ_MAX));
    B2KCall(b2k_host_alloc((void **)&g_srq.h, sizeof(double) * SRQ_MAX));
  }
  PetscCheck(!g_srq.reduced, PETSC_ERR_ORDER, "Called before all BVxxxEnd() called");
  PetscCheck(g_srq.nent < 64 && g_srq.n + cnt <= SRQ_MAX, PETSC_ERR_ARG_SIZ, "too many outstanding split reductions");
  g_srq.e[g_srq.nent].off = g_srq.n; g_srq.e[g_srq.nent].cnt = (int)cnt; g_srq.e[g_srq.nent].owner = (void *)bv;
  *slot = g_srq.d + g_srq.n;
  g_srq.n += (int)((cnt + 1) & ~1);               /* even offsets: 16-byte aligned outputs */
  g_srq.nent++;
  return PETSC_SUCCESS;
}

static PetscErrorCode BVSplitQueuePop_B200(BV bv, PetscInt cnt, const double **vals)
{
  b2k_ctx ctx = CTX();
  if (!g_srq.reduced) {
    PetscCall(B2KCommAllreduce(bv->comm, g_srq.d, g_srq.n, 0, B2K_MEM_DEVICE));
    B2KCall(b2k_d2h_async(ctx, g_srq.h, g_srq.d, sizeof(double) * (size_t)g_srq.n));
    B2KCall(b2k_ctx_sync(ctx));
    g_srq.reduced = PETSC_TRUE;
    g_srq.next = 0;
  }
  PetscCheck(g_srq.next < g_srq.nent, PETSC_ERR_ARG_WRONGSTATE, "Called BVxxxEnd() more times than BVxxxBegin()");
  PetscCheck(g_srq.e[g_srq.next].owner == (void *)bv && g_srq.e[g_srq.next].cnt == (int)cnt, PETSC_ERR_ARG_WRONGSTATE,
             "Called BVxxxEnd() in a different order or with a different BV than BVxxxBegin()");
  *vals = g_srq.h + g_srq.e[g_srq.next].off;
  if (++g_srq.next == g_srq.nent) { g_srq.n = g_srq.nent = g_srq.next = 0; g_srq.reduced = PETSC_FALSE; }
  return PETSC_SUCCESS;
}

static PetscErrorCode BVDotVecBegin_B200(BV X, Vec y, PetscScalar *m)
{
  (void)m;
  BV_B200 *d = (BV_B200 *)X->data;
  const PetscInt k = X->k - X->l;
  PetscCheck(y->mem == B2K_MEM_DEVICE, PETSC_ERR_ARG_INCOMP, "BV type b200 needs device vectors");
  double *slot;
  PetscCall(BVSplitQueuePush_B200(X, k, &slot));
  if (k > 0) B2KCall(b2k_dotvec(CTX(), COL(X, d, X->l), X->ld, X->n, k, y->array, slot));     /* reduce scope closed: local sums */
  return PETSC_SUCCESS;
}
static PetscErrorCode BVDotVecEnd_B200(BV X, Vec y, PetscScalar *m)
{
  (void)y;
  const PetscInt k = X->k - X->l;
  const double *v;
  PetscCall(BVSplitQueuePop_B200(X, k, &v));
  PetscScalar *mm = m ? m : X->buffer;
  if (k > 0) memcpy(mm, v, sizeof(double) * (size_t)k);
  return PETSC_SUCCESS;
}
static PetscErrorCode BVNormBegin_B200(BV bv, PetscInt j, NormType type, PetscReal *val)
{
  (void)val;
  BV_B200 *d = (BV_B200 *)bv->data;
  PetscCheck(type == NORM_2 || type == NORM_FROBENIUS, PETSC_ERR_SUP, "split-phase norms of type b200 are 2-norms");
  const double *X = (j < 0) ? COL(bv, d, bv->l) : COL(bv, d, j);
  const PetscInt k = (j < 0) ? bv->k - bv->l : 1;
  double *slot;
  PetscCall(BVSplitQueuePush_B200(bv, 1, &slot));
  B2KCall(b2k_sumsq(CTX(), X, bv->ld, bv->n, k > 0 ? k : 1, slot));
  return PETSC_SUCCESS;
}
static PetscErrorCode BVNormEnd_B200(BV bv, PetscInt j, NormType type, PetscReal *val)
{
  (void)j; (void)type;
  const double *v;
  PetscCall(BVSplitQueuePop_B200(bv, 1, &v));
  *val = sqrt(v[0]);
  return PETSC_SUCCESS;
}

/* ---- scale / norm / normalize: svec.c:150-175, sveccuda.cu:164-214, bvglobal.c:836 ----------------- */
static PetscErrorCode BVScale_B200(BV bv, PetscInt j, PetscScalar alpha)
{
  BV_B200 *d = (BV_B200 *)bv->data;
  if (j < 0) B2KCall(b2k_scale(CTX(), COL(bv, d, bv->l), bv->ld, bv->n, bv->k - bv->l, alpha));
  else B2KCall(b2k_scale(CTX(), COL(bv, d, j), bv->ld, bv->n, 1, alpha));
  return PETSC_SUCCESS;
}

static PetscErrorCode BVNorm_B200_Private(BV bv, PetscInt j, NormType type, PetscReal *val, PetscBool reduce)
{
  BV_B200 *d = (BV_B200 *)bv->data;
  b2k_ctx ctx = CTX();
  const double *X = (j < 0) ? COL(bv, d, bv->l) : COL(bv, d, j);
  const PetscInt k = (j < 0) ? bv->k - bv->l : 1;
  double *hp;
  if (k <= 0) { *val = 0.0; return PETSC_SUCCESS; }
  PetscBool fused = PETSC_FALSE;
  if (type == NORM_2 || type == NORM_FROBENIUS) {
    PetscCall(BVScope_B200(bv, reduce, &fused));
    B2KCall(b2k_sumsq(ctx, X, bv->ld, bv->n, k, SLOT(d, 3)));
    PetscCall(BVFetch_B200(bv, d, SLOT(d, 3), 1, reduce, fused, &hp));
    *val = sqrt(hp[0]);
  } else if (type == NORM_1) {
    PetscCall(BVScope_B200(bv, reduce, &fused));
    B2KCall(b2k_colabssum(ctx, X, bv->ld, bv->n, k, SLOT(d, 3)));
    PetscCall(BVFetch_B200(bv, d, SLOT(d, 3), k, reduce, fused, &hp));
    PetscReal mx = 0.0;
    for (PetscInt i = 0; i < k; i++) mx = PetscMax(mx, hp[i]);
    *val = mx;
  } else if (type == NORM_INFINITY) {             /* rows are split over the ranks: local max row sum, then a MAX reduction */
    B2KCall(b2k_rowabssum_max(ctx, X, bv->ld, bv->n, k, SLOT(d, 3)));
    if (reduce) PetscCall(B2KCommAllreduce(bv->comm, SLOT(d, 3), 1, 1, B2K_MEM_DEVICE));
    PetscCall(BVFetch_B200(bv, d, SLOT(d, 3), 1, PETSC_FALSE, PETSC_FALSE, &hp));
    *val = hp[0];
  } else SETERRQ(PETSC_ERR_ARG_WRONG, "unknown norm type %d", (int)type);
  return PETSC_SUCCESS;
}
static PetscErrorCode BVNorm_B200(BV bv, PetscInt j, NormType type, PetscReal *val) { return BVNorm_B200_Private(bv, j, type, val, PETSC_TRUE); }
static PetscErrorCode BVNorm_Local_B200(BV bv, PetscInt j, NormType type, PetscReal *val) { return BVNorm_B200_Private(bv, j, type, val, PETSC_FALSE); }

static PetscErrorCode BVNormalize_B200(BV bv, PetscScalar *eigi)
{
  BV_B200 *d = (BV_B200 *)bv->data;
  b2k_ctx ctx = CTX();
  double *hp;
  for (PetscInt i = bv->l; i < bv->k; i++) {
    const PetscInt cols = (eigi && eigi[i] != 0.0 && i + 1 < bv->k) ? 2 : 1;   /* complex conjugate pair stored as two columns */
    PetscBool fused;
    PetscCall(BVScope_B200(bv, PETSC_TRUE, &fused));
    B2KCall(b2k_sumsq(ctx, COL(bv, d, i), bv->ld, bv->n, cols, SLOT(d, 3)));
    PetscCall(BVFetch_B200(bv, d, SLOT(d, 3), 1, PETSC_TRUE, fused, &hp));
    const PetscReal nrm = sqrt(hp[0]);
    if (nrm != 0.0 && nrm != 1.0) B2KCall(b2k_scale(ctx, COL(bv, d, i), bv->ld, bv->n, cols, 1.0 / nrm));
    i += cols - 1;
  }
  return PETSC_SUCCESS;
}

/* ---- level-3: svec.c:17-36,54-107, sveccuda.cu:18-124 ---------------------------------------------- */
static PetscErrorCode BVUploadBlock_B200(BV_B200 *d, const PetscScalar *Q, PetscInt ldq, PetscInt r0, PetscInt c0, PetscInt nr, PetscInt ncol)
{
  PetscCheck((size_t)nr * (size_t)ncol <= d->dQ_elems, PETSC_ERR_ARG_SIZ, "dense block %d x %d larger than the scratch area", nr, ncol);
  for (PetscInt cidx = 0; cidx < ncol; cidx++)
    memcpy(d->hQ + (size_t)cidx * (size_t)nr, Q + (size_t)(c0 + cidx) * (size_t)ldq + r0, sizeof(double) * (size_t)nr);
  if (nr * ncol > 0) B2KCall(b2k_h2d_async(CTX(), d->dQ, d->hQ, sizeof(double) * (size_t)nr * (size_t)ncol));
  return PETSC_SUCCESS;
}

static PetscErrorCode BVMult_B200(BV Y, PetscScalar alpha, PetscScalar beta, BV X, Mat Q)
{
  BV_B200 *y = (BV_B200 *)Y->data, *x = (BV_B200 *)X->data;
  const PetscInt ky = Y->k - Y->l, kx = X->k - X->l;
  if (ky <= 0) return PETSC_SUCCESS;
  if (!Q) {                                       /* BVAXPY bvblas.c:112 */
    B2KCall(b2k_axpby(CTX(), COL(Y, y, Y->l), Y->ld, COL(X, x, X->l), X->ld, Y->n, ky, alpha, beta));
    return PETSC_SUCCESS;
  }
  PetscCall(BVUploadBlock_B200(y, Q->dense, Q->lda, X->l, Y->l, kx, ky));   /* rows from X->l, columns from Y->l: svec.c:29 */
  B2KCall(b2k_mult(CTX(), COL(Y, y, Y->l), Y->ld, COL(X, x, X->l), X->ld, Y->n, ky, kx, alpha, beta, y->dQ, kx > 0 ? kx : 1));
  return PETSC_SUCCESS;
}

static PetscErrorCode BVMultInPlace_B200(BV V, Mat Q, PetscInt s, PetscInt e)
{
  BV_B200 *d = (BV_B200 *)V->data;
  const PetscInt l = V->l, k = V->k - V->l;
  if (s >= e) return PETSC_SUCCESS;
  PetscCall(BVUploadBlock_B200(d, Q->dense, Q->lda, l, l, k, e - l));        /* Q(l:k, l:e) compacted, ld = k */
  B2KCall(b2k_mult_inplace(CTX(), COL(V, d, l), V->ld, V->n, k, s - l, e - l, d->dQ, k > 0 ? k : 1, 0));
  return PETSC_SUCCESS;
}

static PetscErrorCode BVMultInPlaceHermitianTranspose_B200(BV V, Mat Q, PetscInt s, PetscInt e)
{
  BV_B200 *d = (BV_B200 *)V->data;
  const PetscInt l = V->l, k = V->k - V->l;
  if (s >= e) return PETSC_SUCCESS;
  PetscCall(BVUploadBlock_B200(d, Q->dense, Q->lda, l, l, e - l, k));        /* Q(l:e, l:k) compacted, ld = e-l */
  B2KCall(b2k_mult_inplace(CTX(), COL(V, d, l), V->ld, V->n, k, s - l, e - l, d->dQ, e - l, 1));
  return PETSC_SUCCESS;
}

static PetscErrorCode BVDot_B200(BV X, BV Y, Mat M)
{
  BV_B200 *x = (BV_B200 *)X->data, *y = (BV_B200 *)Y->data;
  b2k_ctx ctx = CTX();
  const PetscInt ky = Y->k - Y->l, kx = X->k - X->l;
  PetscCheck((size_t)ky * (size_t)kx <= x->dQ_elems, PETSC_ERR_ARG_SIZ, "result block larger than the scratch area");
  B2KCall(b2k_dot(ctx, COL(Y, y, Y->l), Y->ld, COL(X, x, X->l), X->ld, X->n, ky, kx, x->dQ, ky));
  PetscCall(B2KCommAllreduce(X->comm, x->dQ, ky * kx, 0, B2K_MEM_DEVICE));
  B2KCall(b2k_d2h(ctx, x->hQ, x->dQ, sizeof(double) * (size_t)ky * (size_t)kx));
  for (PetscInt jx = 0; jx < kx; jx++)             /* M(ly:ky, lx:kx), svec.c:101 */
    memcpy(M->dense + (size_t)(X->l + jx) * (size_t)M->lda + Y->l, x->hQ + (size_t)jx * (size_t)ky, sizeof(double) * (size_t)ky);
  return PETSC_SUCCESS;
}

/* BVMatMult: the whole active block at once when the Mat can (BV_MATMULT_MAT, svec.c:203-231: the matrix is read once per 16
   columns by b2k_csr_spmm), the column loop of sveccuda.cu:269-303 otherwise (MatShell operators) */
static PetscErrorCode BVMatMult_B200(BV V, Mat A, BV W)
{
  BV_B200 *v = (BV_B200 *)V->data, *w = (BV_B200 *)W->data;
  const PetscInt k = V->k - V->l;
  if (A->ops.multblock && A->mem == B2K_MEM_DEVICE && k > 1) {
    PetscCall(A->ops.multblock(A, COL(V, v, V->l), V->ld, COL(W, w, W->l), W->ld, k));
    return PETSC_SUCCESS;
  }
  Vec x, y;
  PetscCall(VecCreateWithArray(B2K_MEM_DEVICE, V->n, V->N, NULL, &x));
  PetscCall(VecCreateWithArray(B2K_MEM_DEVICE, W->n, W->N, NULL, &y));
  x->rstart = V->row0; y->rstart = W->row0;
  PetscErrorCode ierr = PETSC_SUCCESS;
  for (PetscInt j = 0; j < k && !ierr; j++) {
    x->array = COL(V, v, V->l + j);
    y->array = COL(W, w, W->l + j);
    ierr = MatMult(A, x, y);
  }
  PetscCall(VecDestroy(&x));
  PetscCall(VecDestroy(&y));
  PetscCall(ierr);
  return PETSC_SUCCESS;
}

static PetscErrorCode BVCopy_B200(BV V, BV W)
{
  BV_B200 *v = (BV_B200 *)V->data, *w = (BV_B200 *)W->data;
  B2KCall(b2k_copy(CTX(), COL(W, w, W->l), W->ld, COL(V, v, V->l), V->ld, V->n, V->k - V->l));
  return PETSC_SUCCESS;
}

static PetscErrorCode BVCopyColumn_B200(BV V, PetscInt j, PetscInt i)
{
  BV_B200 *d = (BV_B200 *)V->data;
  B2KCall(b2k_copy(CTX(), COL(V, d, i), V->ld, COL(V, d, j), V->ld, V->n, 1));
  return PETSC_SUCCESS;
}

static PetscErrorCode BVResize_B200(BV bv, PetscInt m, PetscBool copy)
{
  BV_B200 *d = (BV_B200 *)bv->data;
  b2k_ctx ctx = CTX();
  double *Vnew = NULL;
  const size_t bytes = sizeof(double) * (size_t)(bv->nc + m) * (size_t)bv->ld;
  B2KCall(b2k_malloc(ctx, (void **)&Vnew, bytes));
  B2KCall(b2k_memset0(ctx, Vnew, bytes));
  if (copy) B2KCall(b2k_copy(ctx, Vnew, bv->ld, d->V, bv->ld, bv->n, bv->nc + PetscMin(m, bv->m)));
  B2KCall(b2k_free(ctx, d->V));
  d->V = Vnew;
  const PetscInt msave = bv->m;
  PetscCall(BVFreeScratch_B200(d));
  bv->m = m;
  PetscErrorCode ierr = BVAllocScratch_B200(bv, d);
  bv->m = msave;                                  /* the front-end updates m after we return */
  PetscCall(ierr);
  return PETSC_SUCCESS;
}

static PetscErrorCode BVGetColumn_B200(BV bv, PetscInt j, Vec *v)
{
  BV_B200 *d = (BV_B200 *)bv->data;
  const int l = (bv->ci[0] == j) ? 0 : 1;         /* the front-end recorded the slot (svec.c:292-303) */
  (void)v;
  PetscCall(VecPlaceArray(bv->cv[l], COL(bv, d, j)));
  return PETSC_SUCCESS;
}
static PetscErrorCode BVRestoreColumn_B200(BV bv, PetscInt j, Vec *v)
{
  const int l = (bv->ci[0] == j) ? 0 : 1;
  (void)v;
  PetscCall(VecResetArray(bv->cv[l]));
  return PETSC_SUCCESS;
}
static PetscErrorCode BVGetArray_B200(BV bv, PetscScalar **a) { *a = ((BV_B200 *)bv->data)->V; return PETSC_SUCCESS; }

static PetscErrorCode BVSetRandomColumn_B200(BV bv, PetscInt j)
{
  BV_B200 *d = (BV_B200 *)bv->data;
  B2KCall(b2k_set_random(CTX(), COL(bv, d, j), bv->n, bv->row0, bv->rng_seed + (uint64_t)j));
  return PETSC_SUCCESS;
}

/* ---- tall-skinny QR of the local rows of the active columns (BVOrthogonalize TSQR / TSQRCHOL; b2k_tsqr.cu) --------------------
   level 1 on the device (flat Householder tree over 128-row tiles inside every CTA), level 2 = one LAPACK QR of the CTAs'
   stacked triangles on the host.  The reference does LAPACK geqrf/orgqr on the host copy of the rows (bvlapack.c:378-396). */
static void BVTSQRReset_B200(BV_B200 *d)
{
  if (d->ts_coef && CTX()) b2k_free(CTX(), d->ts_coef);
  free(d->ts_stack);
  d->ts_coef = NULL; d->ts_stack = NULL; d->ts_nk = d->ts_nblk = 0;
}

static PetscErrorCode BVTSQRFactor_B200(BV bv, PetscBool wantq, PetscScalar *R)
{
  BV_B200 *d = (BV_B200 *)bv->data;
  b2k_ctx ctx = CTX();
  const PetscInt nk = bv->k - bv->l;
  int nblk = 0;
  int64_t rpc = 0, ncoef = 0;
  BVTSQRReset_B200(d);
  if (nk <= 0) return PETSC_SUCCESS;
  PetscCheck(nk <= B2K_TSQR_MAX_K, PETSC_ERR_SUP, "TSQR on BV type b200 takes at most %d active columns (%d given); use chol or svqb", B2K_TSQR_MAX_K, nk);
  B2KCall(b2k_tsqr_plan(ctx, bv->n, (int)nk, &nblk, &rpc, &ncoef));
  double *dR = NULL;
  const size_t relems = (size_t)nblk * (size_t)nk * (size_t)nk;
  B2KCall(b2k_malloc(ctx, (void **)&dR, sizeof(double) * relems));
  if (wantq) B2KCall(b2k_malloc(ctx, (void **)&d->ts_coef, sizeof(double) * (size_t)ncoef));
  int rc = b2k_tsqr_forward(ctx, COL(bv, d, bv->l), bv->ld, bv->n, (int)nk, wantq ? 1 : 0, dR, d->ts_coef);
  const PetscInt rows = (PetscInt)nblk * nk;
  int rows_ = rows, nk_ = nk, lwork = 64 * (int)nk + 64, info = 0;
  double *blk = (double *)malloc(sizeof(double) * relems);
  double *stack = (double *)malloc(sizeof(double) * ((size_t)rows * (size_t)nk + (size_t)nk + (size_t)lwork));
  if (!rc && (!blk || !stack)) rc = B2K_ERR_MEM;
  if (!rc) rc = b2k_d2h(ctx, blk, dR, sizeof(double) * relems);
  b2k_free(ctx, dR);
  if (rc) { free(blk); free(stack); BVTSQRReset_B200(d); SETERRQ(PETSC_ERR_GPU, "TSQR forward sweep failed (%d): %s", rc, b2k_last_error()); }
  for (int b = 0; b < nblk; b++)
    for (PetscInt c = 0; c < nk; c++) memcpy(stack + (size_t)b * nk + (size_t)c * rows, blk + ((size_t)b * nk + c) * nk, sizeof(double) * (size_t)nk);
  free(blk);
  double *tau = stack + (size_t)rows * nk, *work = tau + nk;
  LAPACK(dgeqrf)(&rows_, &nk_, stack, &rows_, tau, work, &lwork, &info);
  if (info) { free(stack); BVTSQRReset_B200(d); SETERRQ(PETSC_ERR_LIB, "Error in LAPACK subroutine geqrf: info=%d", info); }
  for (PetscInt c = 0; c < nk; c++) for (PetscInt i = 0; i < nk; i++) R[i + (size_t)c * nk] = (i <= c) ? stack[i + (size_t)c * rows] : 0.0;
  if (wantq) { d->ts_stack = stack; d->ts_nk = nk; d->ts_nblk = nblk; }
  else free(stack);
  return PETSC_SUCCESS;
}

/* columns l..k-1 <- Q_local W: the blocks of (orthogonal factor of the stack) * W start the backward sweep of every CTA */
static PetscErrorCode BVTSQRFormQ_B200(BV bv, const PetscScalar *W)
{
  BV_B200 *d = (BV_B200 *)bv->data;
  b2k_ctx ctx = CTX();
  const PetscInt nk = bv->k - bv->l;
  if (nk <= 0) return PETSC_SUCCESS;
  PetscCheck(d->ts_stack && d->ts_nk == nk, PETSC_ERR_ORDER, "tsqr_formq without a matching tsqr_factor(wantq)");
  const PetscInt nblk = d->ts_nblk, rows = nblk * nk;
  int rows_ = rows, nk_ = nk, lwork = 64 * (int)nk + 64, info = 0;
  double *C = (double *)calloc((size_t)rows * (size_t)nk + (size_t)lwork, sizeof(double));
  double *blk = (double *)malloc(sizeof(double) * (size_t)rows * (size_t)nk);
  PetscCheck(C && blk, PETSC_ERR_MEM, "out of memory");
  for (PetscInt c = 0; c < nk; c++) memcpy(C + (size_t)c * rows, W + (size_t)c * nk, sizeof(double) * (size_t)nk);     /* [W ; 0] */
  LAPACK(dormqr)("L", "N", &rows_, &nk_, &nk_, d->ts_stack, &rows_, d->ts_stack + (size_t)rows * nk, C, &rows_, C + (size_t)rows * nk, &lwork, &info, 1, 1);
  if (info) { free(C); free(blk); SETERRQ(PETSC_ERR_LIB, "Error in LAPACK subroutine ormqr: info=%d", info); }
  for (PetscInt b = 0; b < nblk; b++)
    for (PetscInt c = 0; c < nk; c++) memcpy(blk + ((size_t)b * nk + c) * nk, C + (size_t)b * nk + (size_t)c * rows, sizeof(double) * (size_t)nk);
  free(C);
  double *dW = NULL;
  int rc = b2k_malloc(ctx, (void **)&dW, sizeof(double) * (size_t)rows * (size_t)nk);
  if (!rc) rc = b2k_h2d(ctx, dW, blk, sizeof(double) * (size_t)rows * (size_t)nk);
  if (!rc) rc = b2k_tsqr_backward(ctx, COL(bv, d, bv->l), bv->ld, bv->n, (int)nk, dW, d->ts_coef);
  if (!rc) rc = b2k_ctx_sync(ctx);
  if (dW) b2k_free(ctx, dW);
  free(blk);
  BVTSQRReset_B200(d);
  PetscCheck(!rc, PETSC_ERR_GPU, "TSQR backward sweep failed (%d): %s", rc, b2k_last_error());
  return PETSC_SUCCESS;
}

static PetscErrorCode BVDestroy_B200(BV bv)
{
  BV_B200 *d = (BV_B200 *)bv->data;
  if (!d) return PETSC_SUCCESS;
  BVTSQRReset_B200(d);
  if (CTX() && d->kr_d) b2k_free(CTX(), d->kr_d);
  if (d->kr_h) b2k_host_free(d->kr_h);
  if (CTX()) b2k_free(CTX(), d->V);
  PetscCall(BVFreeScratch_B200(d));
  free(d);
  bv->data = NULL;
  return PETSC_SUCCESS;
}

/* constructor registered with BVRegister("b200", …): what BVCreate_Svec does at svec.c:397-563 */
PetscErrorCode BVCreate_B200(BV bv)
{
  b2k_ctx ctx = CTX();
  PetscCheck(ctx, PETSC_ERR_ORDER, "BV type b200 needs a GPU context: call B2KInitialize() first (there is no CPU fallback)");
  PetscCheck(bv->nc + bv->m <= B2K_MAX_COLUMNS, PETSC_ERR_SUP, "BV type b200 holds at most %d columns (constraints included), %d requested: "
             "reduce ncv/mpd (EPS defaults to ncv = nev+500 for nev >= 500)", B2K_MAX_COLUMNS, bv->nc + bv->m);
  BV_B200 *d = (BV_B200 *)calloc(1, sizeof(*d));
  PetscCheck(d, PETSC_ERR_MEM, "out of memory");
  bv->data = d;
  bv->mem = B2K_MEM_DEVICE;
  const size_t bytes = sizeof(double) * (size_t)(bv->nc + bv->m) * (size_t)bv->ld;   /* 64-bit: m*ld overflows int32, svec.c:425 */
  B2KCall(b2k_malloc(ctx, (void **)&d->V, bytes));
  B2KCall(b2k_memset0(ctx, d->V, bytes));
  PetscCall(BVAllocScratch_B200(bv, d));
  const char *e = getenv("B2K_BV_FUSE");
  d->fuse_mode = e ? atoi(e) : 2;
  e = getenv("B2K_BV_ONESYNC");
  d->onesync = e ? atoi(e) : 1;
  e = getenv("B2K_BV_ASYNC");
  d->async = e ? atoi(e) : 1;
  d->expect_refine = PETSC_TRUE;
  d->last_j = -1;

  bv->ops.mult = BVMult_B200;
  bv->ops.multvec = BVMultVec_B200;
  bv->ops.multinplace = BVMultInPlace_B200;
  bv->ops.multinplacetrans = BVMultInPlaceHermitianTranspose_B200;
  bv->ops.dot = BVDot_B200;
  bv->ops.dotvec = BVDotVec_B200;
  bv->ops.dotvec_local = BVDotVec_Local_B200;
  bv->ops.dotvec_begin = BVDotVecBegin_B200;
  bv->ops.dotvec_end = BVDotVecEnd_B200;
  bv->ops.scale = BVScale_B200;
  bv->ops.norm = BVNorm_B200;
  bv->ops.norm_local = BVNorm_Local_B200;
  bv->ops.norm_begin = BVNormBegin_B200;
  bv->ops.norm_end = BVNormEnd_B200;
  bv->ops.normalize = BVNormalize_B200;
  bv->ops.matmult = BVMatMult_B200;
  bv->ops.copy = BVCopy_B200;
  bv->ops.copycolumn = BVCopyColumn_B200;
  bv->ops.resize = BVResize_B200;
  bv->ops.getcolumn = BVGetColumn_B200;
  bv->ops.restorecolumn = BVRestoreColumn_B200;
  bv->ops.getarray = BVGetArray_B200;
  bv->ops.gramschmidt = BVGramSchmidt_B200;
  bv->ops.destroy = BVDestroy_B200;
  bv->ops.setrandomcolumn = BVSetRandomColumn_B200;
  bv->ops.krylov_steps = BVKrylovSteps_B200;
  bv->ops.tsqr_factor = BVTSQRFactor_B200;
  bv->ops.tsqr_formq = BVTSQRFormQ_B200;
  return PETSC_SUCCESS;
}
