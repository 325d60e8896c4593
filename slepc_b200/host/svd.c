/*
 * svd.c — SVD with the thick-restart Lanczos bidiagonalisation solver (two-sided), the second caller
 * of the hot path.  Restates (host control flow only)
 *   SVDSetUp (transpose handling, M<N swap)       src/svd/interface/svdsetup.c:226-350, 454-478
 *   SVDSetUp_TRLanczos / SVDSolve_TRLanczos       src/svd/impls/trlanczos/trlanczos.c:203-262, 450-551
 *   SVDTwoSideLanczos, SVDKrylovConvergence       src/svd/impls/lanczos/gklanczos.c:58-113, 189-216
 *   SVDSolve, SVDGetSingularTriplet, SVDComputeError   src/svd/interface/svdsolve.c:116, 313, 486
 *   SVDOneSideTRLanczosCGS/MGS, SVDOrthogonalizeCGS    src/svd/impls/trlanczos/trlanczos.c:264-448 (SVDTRLanczosSetOneSide)
 * The generalized / hyperbolic branches (trlanczos.c:553-1620) are outside the named path.
 */
#include "b2kimpl.h"

PetscErrorCode SVDCreate(SVD *out)
{
  SVD svd = (SVD)calloc(1, sizeof(*svd));
  PetscCheck(svd, PETSC_ERR_MEM, "out of memory");
  svd->nsv = 1; svd->ncv = PETSC_DETERMINE; svd->mpd = PETSC_DETERMINE; svd->max_it = PETSC_DETERMINE;
  svd->tol = (PetscReal)PETSC_DETERMINE;
  svd->which = SVD_LARGEST;
  svd->lock = PETSC_TRUE;
  PetscCall(BVCreate(&svd->U));
  PetscCall(BVCreate(&svd->V));
  PetscCall(DSCreate(&svd->ds));
  *out = svd;
  return PETSC_SUCCESS;
}

PetscErrorCode SVDDestroy(SVD *psvd)
{
  if (!psvd || !*psvd) return PETSC_SUCCESS;
  SVD svd = *psvd;
  if (svd->owns_AT) { if (svd->swapped) PetscCall(MatDestroy(&svd->A)); else PetscCall(MatDestroy(&svd->AT)); }
  PetscCall(BVDestroy(&svd->U));
  PetscCall(BVDestroy(&svd->V));
  PetscCall(DSDestroy(&svd->ds));
  PetscCall(VecDestroy(&svd->iniV));
  PetscCall(VecDestroy(&svd->iniU));
  for (int i = 0; i < 4; i++) PetscCall(VecDestroy(&svd->work[i]));
  free(svd->sigma); free(svd->errest); free(svd->perm); free(svd->swork);
  free(svd);
  *psvd = NULL;
  return PETSC_SUCCESS;
}

PetscErrorCode SVDSetOperators(SVD svd, Mat A, Mat B)
{
  PetscCheck(A, PETSC_ERR_ARG_NULL, "null matrix");
  PetscCheck(!B, PETSC_ERR_SUP, "the generalized SVD is outside the Krylov hot path");
  if (svd->OP && (svd->OP->M != A->M || svd->OP->N != A->N || svd->OP->m != A->m || svd->OP->n != A->n)) {
    /* SVDReset, svdsetup.c:107-116: another size ⇒ the bases, the work vectors and the initial vectors go; a transpose given for the
       old matrix cannot belong to the new one */
    if (svd->owns_AT) { if (svd->swapped) PetscCall(MatDestroy(&svd->A)); else PetscCall(MatDestroy(&svd->AT)); svd->owns_AT = PETSC_FALSE; }
    svd->A = svd->AT = NULL;
    if (svd->swapped) { BV bv = svd->V; svd->V = svd->U; svd->U = bv; svd->swapped = PETSC_FALSE; }
    if (svd->V->sizes_set) PetscCall(BVForgetSizes_Private(svd->V));
    if (svd->U->sizes_set) PetscCall(BVForgetSizes_Private(svd->U));
    for (int i = 0; i < 4; i++) PetscCall(VecDestroy(&svd->work[i]));
    PetscCall(VecDestroy(&svd->iniV));
    PetscCall(VecDestroy(&svd->iniU));
    svd->userAT = NULL;
  }
  svd->OP = A;
  svd->setup_done = PETSC_FALSE; svd->solved = PETSC_FALSE;
  return PETSC_SUCCESS;
}
PetscErrorCode SVDSetTransposeMatrix(SVD svd, Mat AT) { svd->userAT = AT; svd->setup_done = PETSC_FALSE; return PETSC_SUCCESS; }
PetscErrorCode SVDSetImplicitTranspose(SVD svd, PetscBool impl) { svd->impltrans = impl; svd->setup_done = PETSC_FALSE; return PETSC_SUCCESS; }
PetscErrorCode SVDSetType(SVD svd, const char *type)
{
  (void)svd;
  PetscCheck(!strcmp(type, SVDTRLANCZOS), PETSC_ERR_ARG_UNKNOWN_TYPE, "Unknown SVD type given: %s (this build provides trlanczos)", type);
  return PETSC_SUCCESS;
}
PetscErrorCode SVDSetDimensions(SVD svd, PetscInt nsv, PetscInt ncv, PetscInt mpd)
{
  if (nsv != PETSC_CURRENT) { PetscCheck(nsv > 0, PETSC_ERR_ARG_OUTOFRANGE, "Illegal value of nsv. Must be > 0"); svd->nsv = nsv; }
  if (ncv == PETSC_DETERMINE || ncv == PETSC_DEFAULT) svd->ncv = PETSC_DETERMINE;
  else if (ncv != PETSC_CURRENT) { PetscCheck(ncv > 0, PETSC_ERR_ARG_OUTOFRANGE, "Illegal value of ncv. Must be > 0"); svd->ncv = ncv; }
  if (mpd == PETSC_DETERMINE || mpd == PETSC_DEFAULT) svd->mpd = PETSC_DETERMINE;
  else if (mpd != PETSC_CURRENT) { PetscCheck(mpd > 0, PETSC_ERR_ARG_OUTOFRANGE, "Illegal value of mpd. Must be > 0"); svd->mpd = mpd; }
  svd->setup_done = PETSC_FALSE;
  return PETSC_SUCCESS;
}
PetscErrorCode SVDGetDimensions(SVD svd, PetscInt *nsv, PetscInt *ncv, PetscInt *mpd)
{
  if (nsv) *nsv = svd->nsv;
  if (ncv) *ncv = svd->ncv;
  if (mpd) *mpd = svd->mpd;
  return PETSC_SUCCESS;
}
PetscErrorCode SVDSetTolerances(SVD svd, PetscReal tol, PetscInt maxits)
{
  if (tol == (PetscReal)PETSC_DETERMINE || tol == (PetscReal)PETSC_DEFAULT) svd->tol = (PetscReal)PETSC_DETERMINE;
  else if (tol != (PetscReal)PETSC_CURRENT) { PetscCheck(tol > 0.0, PETSC_ERR_ARG_OUTOFRANGE, "Illegal value of tol. Must be > 0"); svd->tol = tol; }
  if (maxits == PETSC_DETERMINE || maxits == PETSC_DEFAULT) svd->max_it = PETSC_DETERMINE;
  else if (maxits != PETSC_CURRENT) { PetscCheck(maxits > 0, PETSC_ERR_ARG_OUTOFRANGE, "Illegal value of maxits. Must be > 0"); svd->max_it = maxits; }
  svd->setup_done = PETSC_FALSE;
  return PETSC_SUCCESS;
}
PetscErrorCode SVDSetWhichSingularTriplets(SVD svd, SVDWhich which)
{
  PetscCheck(which == SVD_LARGEST || which == SVD_SMALLEST, PETSC_ERR_ARG_OUTOFRANGE, "Invalid 'which' parameter");
  svd->which = which;
  svd->setup_done = PETSC_FALSE;
  return PETSC_SUCCESS;
}
PetscErrorCode SVDTRLanczosSetRestart(SVD svd, PetscReal keep)
{
  if (keep == (PetscReal)PETSC_DEFAULT || keep == (PetscReal)PETSC_DECIDE) svd->keep = 0.5;
  else { PetscCheck(keep >= 0.1 && keep <= 0.9, PETSC_ERR_ARG_OUTOFRANGE, "The keep argument %g must be in the range [0.1,0.9]", keep); svd->keep = keep; }
  return PETSC_SUCCESS;
}
PetscErrorCode SVDTRLanczosSetLocking(SVD svd, PetscBool lock) { svd->lock = lock; return PETSC_SUCCESS; }
PetscErrorCode SVDTRLanczosSetOneSide(SVD svd, PetscBool oneside)
{
  svd->oneside = oneside;
  return PETSC_SUCCESS;
}
PetscErrorCode SVDGetBV(SVD svd, BV *V, BV *U) { if (V) *V = svd->V; if (U) *U = svd->U; return PETSC_SUCCESS; }

PetscErrorCode SVDSetInitialSpaces(SVD svd, PetscInt nr, Vec isr[], PetscInt nl, Vec isl[])
{
  PetscCheck(nr >= 0 && nl >= 0, PETSC_ERR_ARG_OUTOFRANGE, "Argument cannot be negative");
  PetscCall(VecDestroy(&svd->iniV));
  PetscCall(VecDestroy(&svd->iniU));
  if (nr > 0) {
    if (isr[0]->mem == B2K_MEM_DEVICE) PetscCall(VecCreateB200(isr[0]->n, isr[0]->N, &svd->iniV)); else PetscCall(VecCreateHost(isr[0]->n, isr[0]->N, &svd->iniV));
    PetscCall(VecCopy(isr[0], svd->iniV));
  }
  if (nl > 0) {
    if (isl[0]->mem == B2K_MEM_DEVICE) PetscCall(VecCreateB200(isl[0]->n, isl[0]->N, &svd->iniU)); else PetscCall(VecCreateHost(isl[0]->n, isl[0]->N, &svd->iniU));
    PetscCall(VecCopy(isl[0], svd->iniU));
  }
  svd->setup_done = PETSC_FALSE;
  return PETSC_SUCCESS;
}

PetscErrorCode SVDSetUp(SVD svd)
{
  if (svd->setup_done) return PETSC_SUCCESS;
  PetscCheck(svd->OP, PETSC_ERR_ARG_WRONGSTATE, "SVDSetOperators() must be called first");
  svd->reason = SVD_CONVERGED_ITERATING;
  if (svd->tol == (PetscReal)PETSC_DETERMINE) svd->tol = SLEPC_DEFAULT_TOL;
  const PetscInt M = svd->OP->M, N = svd->OP->N;
  /* transpose handling svdsetup.c:271-315: explicit A^T unless told otherwise; a user-supplied A^T
     (SVDSetTransposeMatrix) replaces MatHermitianTranspose on more than one rank */
  if (svd->owns_AT) { if (svd->swapped) PetscCall(MatDestroy(&svd->A)); else PetscCall(MatDestroy(&svd->AT)); svd->owns_AT = PETSC_FALSE; }
  Mat T = svd->userAT;
  if (!T) {
    /* svdsetup.c:271-315: explicit transpose unless asked otherwise or the Mat type cannot build one (b200csr on more than
       one rank has no MATOP_TRANSPOSE): then a virtual transpose over MatMultTranspose */
    int csize = 1;
    PetscCall(B2KCommGetRank(B2KCommWorld(), NULL, &csize));
    const PetscBool is_csr = !strcmp(svd->OP->type, "b200csr") ? PETSC_TRUE : PETSC_FALSE;
    if (svd->impltrans || !is_csr || csize > 1) PetscCall(MatCreateHermitianTranspose(svd->OP, &T));
    else PetscCall(MatB200CSRTranspose(svd->OP, &T));
    svd->owns_AT = PETSC_TRUE;
  } else PetscCheck(T->M == N && T->N == M, PETSC_ERR_ARG_SIZ, "the transpose matrix is %d x %d, expected %d x %d", T->M, T->N, N, M);
  const PetscBool swap = (M < N) ? PETSC_TRUE : PETSC_FALSE;
  if (!swap) { svd->A = svd->OP; svd->AT = T; }
  else { svd->A = T; svd->AT = svd->OP; }        /* work with A^T so that the working A has at least as many rows as columns */
  if (swap && !svd->swapped) { BV bv = svd->V; svd->V = svd->U; svd->U = bv; Vec t = svd->iniV; svd->iniV = svd->iniU; svd->iniU = t; }
  if (!swap && svd->swapped) { BV bv = svd->V; svd->V = svd->U; svd->U = bv; Vec t = svd->iniV; svd->iniV = svd->iniU; svd->iniU = t; }
  svd->swapped = swap;

  const PetscInt maxnsol = PetscMin(M, N);
  if (svd->ncv != PETSC_DETERMINE) svd->ncv = PetscMin(svd->ncv, maxnsol);
  svd->nsv = PetscMin(svd->nsv, maxnsol);
  PetscCheck(svd->ncv == PETSC_DETERMINE || svd->nsv <= svd->ncv, PETSC_ERR_ARG_OUTOFRANGE, "nsv bigger than ncv");
  /* SVDSetDimensions_Default svdsetup.c:454-478 */
  if (svd->ncv != PETSC_DETERMINE) PetscCheck(svd->ncv >= svd->nsv, PETSC_ERR_USER_INPUT, "The value of ncv must be at least nsv");
  else if (svd->mpd != PETSC_DETERMINE) svd->ncv = PetscMin(maxnsol, svd->nsv + svd->mpd);
  else {
    if (svd->nsv < 500) svd->ncv = PetscMin(maxnsol, PetscMax(2 * svd->nsv, 10));
    else { svd->mpd = 500; svd->ncv = PetscMin(maxnsol, svd->nsv + svd->mpd); }
  }
  if (svd->mpd == PETSC_DETERMINE) svd->mpd = svd->ncv;
  /* SVDSetUp_TRLanczos trlanczos.c:211-216 */
  PetscCheck(svd->ncv <= svd->nsv + svd->mpd, PETSC_ERR_USER_INPUT, "The value of ncv must not be larger than nsv+mpd");
  PetscCheck(svd->lock || svd->mpd >= svd->ncv, PETSC_ERR_SUP, "Should not use mpd parameter in non-locking variant");
  if (svd->max_it == PETSC_DETERMINE) svd->max_it = PetscMax(svd->A->N / svd->ncv, 100);
  if (!svd->keep) svd->keep = 0.5;

  /* SVDAllocateSolution(svd,1) svdsetup.c */
  const PetscInt requested = svd->ncv + 1;
  if (requested != svd->allocated) {
    free(svd->sigma); free(svd->errest); free(svd->perm); free(svd->swork);
    svd->swork = (PetscScalar *)calloc((size_t)requested + 1, sizeof(PetscScalar));
    svd->sigma = (PetscReal *)calloc((size_t)requested, sizeof(PetscReal));
    svd->errest = (PetscReal *)calloc((size_t)requested, sizeof(PetscReal));
    svd->perm = (PetscInt *)calloc((size_t)requested, sizeof(PetscInt));
    PetscCheck(svd->sigma && svd->errest && svd->perm && svd->swork, PETSC_ERR_MEM, "out of memory");
    svd->allocated = requested;
  }
  BV bvs[2] = {svd->V, svd->U};
  for (int b = 0; b < 2; b++) {
    BV bv = bvs[b];
    const PetscInt old = bv->sizes_set ? bv->m : 0;
    if (!old) {
      if (!bv->ctor) PetscCall(BVSetType(bv, BVB200));
      Vec t;
      if (b == 0) PetscCall(MatCreateVecs(svd->A, &t, NULL)); else PetscCall(MatCreateVecs(svd->A, NULL, &t));
      PetscErrorCode ierr = BVSetSizesFromVec(bv, t, requested);
      PetscCall(VecDestroy(&t));
      PetscCall(ierr);
    } else if (old != requested) PetscCall(BVResize(bv, requested, PETSC_FALSE));
  }
  PetscCall(BVSetRandomSeed(svd->U, svd->V->rng_seed + 1000));
  PetscCall(DSSetType(svd->ds, DSSVD));
  PetscCall(DSSetCompact(svd->ds, PETSC_TRUE));
  PetscCall(DSSetExtraRow(svd->ds, PETSC_TRUE));
  PetscCall(DSAllocate(svd->ds, svd->ncv + 1));
  PetscCall(DSSetEigenvalueComparison(svd->ds, svd->which == SVD_LARGEST ? SlepcCompareLargestReal : SlepcCompareSmallestReal, NULL));
  if (svd->iniV) {
    PetscCheck(svd->iniV->n == svd->V->n, PETSC_ERR_ARG_INCOMP, "the initial vector has local size %d, expected %d", svd->iniV->n, svd->V->n);
    PetscCall(BVInsertVec(svd->V, 0, svd->iniV));
    PetscCall(BVOrthonormalizeColumn(svd->V, 0, PETSC_TRUE, NULL, NULL));
  }
  svd->setup_done = PETSC_TRUE; svd->solved = PETSC_FALSE; svd->started = PETSC_FALSE;
  return PETSC_SUCCESS;
}

static PetscErrorCode SVDMatMultColumns_Private(Mat A, BV X, PetscInt jx, BV Y, PetscInt jy)
{
  Vec x, y;
  PetscCall(BVGetColumn(X, jx, &x));
  PetscCall(BVGetColumn(Y, jy, &y));
  PetscErrorCode ierr = MatMult(A, x, y);
  PetscCall(BVRestoreColumn(X, jx, &x));
  PetscCall(BVRestoreColumn(Y, jy, &y));
  PetscCall(ierr);
  Y->n_matmult++;
  return PETSC_SUCCESS;
}

/* gklanczos.c:58-113 */
static PetscErrorCode SVDTwoSideLanczos(SVD svd, PetscReal *alpha, PetscReal *beta, PetscInt k, PetscInt *n, PetscBool *breakdown)
{
  PetscBool lindep = PETSC_FALSE;
  PetscCall(SVDMatMultColumns_Private(svd->A, svd->V, k, svd->U, k));
  PetscCall(BVOrthonormalizeColumn(svd->U, k, PETSC_FALSE, alpha + k, &lindep));
  if (lindep) { *n = k; if (breakdown) *breakdown = lindep; return PETSC_SUCCESS; }
  for (PetscInt i = k + 1; i < *n; i++) {
    PetscCall(SVDMatMultColumns_Private(svd->AT, svd->U, i - 1, svd->V, i));
    PetscCall(BVOrthonormalizeColumn(svd->V, i, PETSC_FALSE, beta + i - 1, &lindep));
    if (lindep) { *n = i; break; }
    PetscCall(SVDMatMultColumns_Private(svd->A, svd->V, i, svd->U, i));
    PetscCall(BVOrthonormalizeColumn(svd->U, i, PETSC_FALSE, alpha + i, &lindep));
    if (lindep) { *n = i; break; }
  }
  if (!lindep) {
    PetscCall(SVDMatMultColumns_Private(svd->AT, svd->U, *n - 1, svd->V, *n));
    PetscCall(BVOrthogonalizeColumn(svd->V, *n, NULL, beta + *n - 1, &lindep));
  }
  if (breakdown) *breakdown = lindep;
  return PETSC_SUCCESS;
}

/* SVDOrthogonalizeCGS trlanczos.c:319-355: post-processing of the merged first CGS pass of the one-sided recurrence;
   h[0:i] = scaled coefficients, h[i] = v_i^T v_i taken before the update, a = ||u_{i-1}|| */
static PetscErrorCode SVDOrthogonalizeCGS_Private(BV V, PetscInt i, PetscScalar *h, PetscReal a, BVOrthogRefineType refine, PetscReal eta, PetscReal *norm)
{
  switch (refine) {
  case BV_ORTHOG_REFINE_NEVER:
    PetscCall(BVNormColumn(V, i, NORM_2, norm));
    break;
  case BV_ORTHOG_REFINE_ALWAYS:
    PetscCall(BVSetActiveColumns(V, 0, i));
    PetscCall(BVDotColumn(V, i, h));
    PetscCall(BVMultColumn(V, -1.0, 1.0, i, h));
    PetscCall(BVNormColumn(V, i, NORM_2, norm));
    break;
  case BV_ORTHOG_REFINE_IFNEEDED: {
    const PetscScalar dot = h[i];
    const PetscReal onorm = sqrt(dot) / a;
    PetscReal sum = 0.0;
    for (PetscInt j = 0; j < i; j++) sum += h[j] * h[j];
    *norm = dot / (a * a) - sum;
    if (*norm > 0.0) *norm = sqrt(*norm);
    else PetscCall(BVNormColumn(V, i, NORM_2, norm));
    if (*norm < eta * onorm) {
      PetscCall(BVSetActiveColumns(V, 0, i));
      PetscCall(BVDotColumn(V, i, h));
      PetscCall(BVMultColumn(V, -1.0, 1.0, i, h));
      PetscCall(BVNormColumn(V, i, NORM_2, norm));
    }
  } break;
  }
  return PETSC_SUCCESS;
}

/* ||u_{i-1}|| together with the first CGS pass of v_i, merged into ONE reduction with the split-phase calls exactly as
   trlanczos.c:383-396 does (BVNormColumnBegin + BVDotVecBegin … End: one all-reduce, one host synchronisation), then the
   scaled update v_i <- v_i/a - V (h/a) and the refinement decision */
static PetscErrorCode SVDOneSideStep_Private(SVD svd, PetscInt i, PetscScalar *work, BVOrthogRefineType refine, PetscReal eta, PetscReal *pa, PetscReal *pb)
{
  BV V = svd->V, U = svd->U;
  PetscReal a, b;
  Vec vi = NULL;
  PetscCall(BVNormColumnBegin(U, i - 1, NORM_2, &a));
  if (refine == BV_ORTHOG_REFINE_IFNEEDED) {
    PetscCall(BVSetActiveColumns(V, 0, i + 1));
    PetscCall(BVGetColumn(V, i, &vi));
    PetscCall(BVDotVecBegin(V, vi, work));               /* work[i] = v_i^T v_i comes with the same sweep */
  } else {
    PetscCall(BVSetActiveColumns(V, 0, i));
    PetscCall(BVDotColumnBegin(V, i, work));
  }
  PetscCall(BVNormColumnEnd(U, i - 1, NORM_2, &a));
  if (refine == BV_ORTHOG_REFINE_IFNEEDED) {
    PetscCall(BVDotVecEnd(V, vi, work));
    PetscCall(BVRestoreColumn(V, i, &vi));
    PetscCall(BVSetActiveColumns(V, 0, i));
  } else PetscCall(BVDotColumnEnd(V, i, work));
  PetscCall(BVScaleColumn(U, i - 1, 1.0 / a));
  for (PetscInt j = 0; j < i; j++) work[j] = work[j] / a;
  PetscCall(BVMultColumn(V, -1.0, 1.0 / a, i, work));
  PetscCall(SVDOrthogonalizeCGS_Private(V, i, work, a, refine, eta, &b));
  *pa = a; *pb = b;
  return PETSC_SUCCESS;
}

/* u_i = A v_i - b u_{i-1}  (trlanczos.c:408-416) */
static PetscErrorCode SVDOneSideNextU_Private(SVD svd, PetscInt i, PetscReal b)
{
  Vec ui, ui1;
  PetscCall(SVDMatMultColumns_Private(svd->A, svd->V, i, svd->U, i));
  PetscCall(BVGetColumn(svd->U, i, &ui));
  PetscCall(BVGetColumn(svd->U, i - 1, &ui1));
  PetscErrorCode ierr = VecAXPY(ui, -b, ui1);
  PetscCall(BVRestoreColumn(svd->U, i, &ui));
  PetscCall(BVRestoreColumn(svd->U, i - 1, &ui1));
  PetscCall(ierr);
  svd->U->state++;
  return PETSC_SUCCESS;
}

/* u_k = A v_k - U(:,nconv:k) beta(nconv:k)  (trlanczos.c:365-374, :272-281) */
static PetscErrorCode SVDOneSideFirstU_Private(SVD svd, PetscReal *beta, PetscInt nconv, PetscInt l, PetscInt n, PetscScalar *work)
{
  const PetscInt k = nconv + l;
  PetscCall(SVDMatMultColumns_Private(svd->A, svd->V, k, svd->U, k));
  if (l > 0) {
    PetscCall(BVSetActiveColumns(svd->U, nconv, n));
    for (PetscInt i = 0; i < l; i++) work[i] = beta[i + nconv];
    PetscCall(BVMultColumn(svd->U, -1.0, 1.0, k, work));
  }
  return PETSC_SUCCESS;
}

/* SVDOneSideTRLanczosCGS trlanczos.c:357-448: only V is orthogonalised (one merged CGS pass + refinement if needed);
   U comes from the three-term recurrence and is re-orthonormalised once, after convergence */
static PetscErrorCode SVDOneSideTRLanczosCGS(SVD svd, PetscReal *alpha, PetscReal *beta, PetscInt nconv, PetscInt l, PetscInt n, PetscScalar *work)
{
  PetscReal a, b, eta;
  const PetscInt k = nconv + l;
  BVOrthogRefineType refine;
  PetscCall(SVDOneSideFirstU_Private(svd, beta, nconv, l, n, work));
  PetscCall(BVGetOrthogonalization(svd->V, NULL, &refine, &eta, NULL));
  for (PetscInt i = k + 1; i < n; i++) {
    PetscCall(SVDMatMultColumns_Private(svd->AT, svd->U, i - 1, svd->V, i));
    PetscCall(SVDOneSideStep_Private(svd, i, work, refine, eta, &a, &b));
    PetscCall(BVScaleColumn(svd->V, i, 1.0 / b));
    PetscCheck(fabs(b) > 10 * PETSC_MACHINE_EPSILON, PETSC_ERR_PLIB, "Recurrence generated a zero vector; use a two-sided variant");
    PetscCall(SVDOneSideNextU_Private(svd, i, b));
    alpha[i - 1] = a;
    beta[i - 1] = b;
  }
  PetscCall(SVDMatMultColumns_Private(svd->AT, svd->U, n - 1, svd->V, n));
  PetscCall(SVDOneSideStep_Private(svd, n, work, refine, eta, &a, &b));
  PetscCall(BVSetActiveColumns(svd->V, nconv, n));
  alpha[n - 1] = a;
  beta[n - 1] = b;
  return PETSC_SUCCESS;
}

/* SVDOneSideTRLanczosMGS trlanczos.c:264-314 (taken when the BV orthogonalisation type is MGS, :481) */
static PetscErrorCode SVDOneSideTRLanczosMGS(SVD svd, PetscReal *alpha, PetscReal *beta, PetscInt nconv, PetscInt l, PetscInt n, PetscScalar *work)
{
  PetscReal a, b;
  const PetscInt k = nconv + l;
  PetscCall(SVDOneSideFirstU_Private(svd, beta, nconv, l, n, work));
  PetscCall(BVNormColumn(svd->U, k, NORM_2, &a));
  PetscCall(BVScaleColumn(svd->U, k, 1.0 / a));
  alpha[k] = a;
  for (PetscInt i = k + 1; i < n; i++) {
    PetscCall(SVDMatMultColumns_Private(svd->AT, svd->U, i - 1, svd->V, i));
    PetscCall(BVOrthonormalizeColumn(svd->V, i, PETSC_FALSE, &b, NULL));
    beta[i - 1] = b;
    PetscCall(SVDOneSideNextU_Private(svd, i, b));
    PetscCall(BVNormColumn(svd->U, i, NORM_2, &a));
    PetscCall(BVScaleColumn(svd->U, i, 1.0 / a));
    alpha[i] = a;
  }
  PetscCall(SVDMatMultColumns_Private(svd->AT, svd->U, n - 1, svd->V, n));
  PetscCall(BVOrthogonalizeColumn(svd->V, n, NULL, &b, NULL));
  beta[n - 1] = b;
  return PETSC_SUCCESS;
}

/* one pass of the restart loop of SVDSolve_TRLanczos, trlanczos.c:475-537 */
static PetscErrorCode SVDTRLanczosCycle_Private(SVD svd)
{
  PetscInt k, l = svd->l, nv;
  PetscBool breakdown = PETSC_FALSE;
  const PetscInt ld = svd->ds->ld;
  PetscReal *alpha, *beta;
  PetscScalar *w = svd->ds->work + (size_t)ld * ld + 8 * (size_t)ld;   /* ld scalars at the tail of the DS workspace */
  Mat U, V;

  svd->its++;
  nv = PetscMin(svd->nconv + svd->mpd, svd->ncv);
  PetscCall(DSGetArrayReal(svd->ds, DS_MAT_T, &alpha));
  beta = alpha + ld;
  if (svd->oneside) {                             /* trlanczos.c:480-483 */
    PetscScalar *swork = svd->swork;
    if (svd->V->orthog_type == BV_ORTHOG_MGS) PetscCall(SVDOneSideTRLanczosMGS(svd, alpha, beta, svd->nconv, l, nv, swork));
    else PetscCall(SVDOneSideTRLanczosCGS(svd, alpha, beta, svd->nconv, l, nv, swork));
  } else PetscCall(SVDTwoSideLanczos(svd, alpha, beta, svd->nconv + l, &nv, &breakdown));
  PetscCall(BVScaleColumn(svd->V, nv, 1.0 / beta[nv - 1]));
  PetscCall(BVSetActiveColumns(svd->V, svd->nconv, nv));
  PetscCall(BVSetActiveColumns(svd->U, svd->nconv, nv));

  PetscCall(DSSetDimensions(svd->ds, nv, svd->nconv, svd->nconv + l));
  PetscCall(DSSVDSetDimensions(svd->ds, nv));
  PetscCall(DSSetState(svd->ds, l ? DS_STATE_RAW : DS_STATE_INTERMEDIATE));
  PetscCall(DSSolve(svd->ds, w, NULL));
  PetscCall(DSSort(svd->ds, w, NULL, NULL, NULL, NULL));
  PetscCall(DSUpdateExtraRow(svd->ds));
  PetscCall(DSSynchronize(svd->ds, w, NULL));
  for (PetscInt i = svd->nconv; i < nv; i++) svd->sigma[i] = w[i];

  /* SVDKrylovConvergence gklanczos.c:189-216 with SVDConvergedRelative svddefault.c:29-34 */
  {
    PetscInt marker = -1;
    for (k = svd->nconv; k < nv; k++) {
      const PetscReal resnorm = fabs(beta[k]);
      svd->errest[k] = (svd->sigma[k] != 0.0) ? resnorm / svd->sigma[k] : PETSC_MAX_REAL;
      if (marker == -1 && svd->errest[k] >= svd->tol) marker = k;
      if (marker != -1) break;
    }
    if (marker != -1) k = marker;
  }
  svd->reason = SVD_CONVERGED_ITERATING;           /* SVDStoppingBasic */
  if (k >= svd->nsv) svd->reason = SVD_CONVERGED_TOL;
  else if (svd->its >= svd->max_it) svd->reason = SVD_DIVERGED_ITS;

  if (svd->reason != SVD_CONVERGED_ITERATING || breakdown || k == nv) l = 0;
  else l = PetscMax(1, (PetscInt)((nv - k) * svd->keep));
  if (!svd->lock && l > 0) { l += k; k = 0; }

  if (svd->reason == SVD_CONVERGED_ITERATING) {
    if (breakdown || k == nv) {                   /* start a new bidiagonalization */
      if (k < svd->nsv) {
        PetscCall(BVSetRandomColumn(svd->V, k));
        PetscCall(BVOrthonormalizeColumn(svd->V, k, PETSC_FALSE, NULL, &breakdown));
        if (breakdown) svd->reason = SVD_DIVERGED_BREAKDOWN;
      }
    } else PetscCall(DSTruncate(svd->ds, k + l, PETSC_FALSE));
  }
  /* converged singular vectors and restart vectors */
  PetscCall(DSGetMat(svd->ds, DS_MAT_V, &V));
  PetscErrorCode ierr = BVMultInPlace(svd->V, V, svd->nconv, k + l);
  PetscCall(DSRestoreMat(svd->ds, DS_MAT_V, &V));
  PetscCall(ierr);
  PetscCall(DSGetMat(svd->ds, DS_MAT_U, &U));
  ierr = BVMultInPlace(svd->U, U, svd->nconv, k + l);
  PetscCall(DSRestoreMat(svd->ds, DS_MAT_U, &U));
  PetscCall(ierr);
  if (svd->reason == SVD_CONVERGED_ITERATING && !breakdown) PetscCall(BVCopyColumn(svd->V, nv, k + l));
  svd->nconv = k;
  svd->l = l;
  return PETSC_SUCCESS;
}

PetscErrorCode SVDSolve(SVD svd)
{
  if (svd->solved && svd->setup_done) return PETSC_SUCCESS;
  PetscCall(SVDSetUp(svd));
  svd->its = 0; svd->nconv = 0;
  for (PetscInt i = 0; i < svd->ncv; i++) { svd->sigma[i] = 0.0; svd->errest[i] = 0.0; svd->perm[i] = i; }
  svd->reason = SVD_CONVERGED_ITERATING;
  if (!svd->iniV) {                               /* normalize start vector, trlanczos.c:468-472 */
    PetscCall(BVSetRandomColumn(svd->V, 0));
    PetscCall(BVOrthonormalizeColumn(svd->V, 0, PETSC_TRUE, NULL, NULL));
  }
  svd->l = 0;
  while (svd->reason == SVD_CONVERGED_ITERATING) PetscCall(SVDTRLanczosCycle_Private(svd));
  if (svd->oneside)                               /* orthonormalize U columns in one side method, trlanczos.c:540-542 */
    for (PetscInt i = 0; i < svd->nconv; i++) PetscCall(BVOrthonormalizeColumn(svd->U, i, PETSC_FALSE, NULL, NULL));
  PetscCall(DSTruncate(svd->ds, svd->nconv, PETSC_TRUE));
  svd->solved = PETSC_TRUE;
  /* sort singular triplets, svdsolve.c:149-157 */
  for (PetscInt i = 0; i < svd->nconv; i++) svd->perm[i] = i;
  for (PetscInt i = 1; i < svd->nconv; i++) {      /* ascending by sigma (PetscSortRealWithPermutation), reversed for SVD_LARGEST */
    const PetscInt p = svd->perm[i];
    PetscInt j = i - 1;
    while (j >= 0 && svd->sigma[svd->perm[j]] > svd->sigma[p]) { svd->perm[j + 1] = svd->perm[j]; j--; }
    svd->perm[j + 1] = p;
  }
  if (svd->which == SVD_LARGEST)
    for (PetscInt i = 0; i < svd->nconv / 2; i++) { const PetscInt t = svd->perm[i]; svd->perm[i] = svd->perm[svd->nconv - 1 - i]; svd->perm[svd->nconv - 1 - i] = t; }
  return PETSC_SUCCESS;
}

#define SVDCheckSolved(svd) PetscCheck((svd)->solved, PETSC_ERR_ARG_WRONGSTATE, "Must call SVDSolve() first")
PetscErrorCode SVDGetIterationNumber(SVD svd, PetscInt *its) { *its = svd->its; return PETSC_SUCCESS; }
PetscErrorCode SVDGetConverged(SVD svd, PetscInt *nconv) { SVDCheckSolved(svd); *nconv = svd->nconv; return PETSC_SUCCESS; }
PetscErrorCode SVDGetConvergedReason(SVD svd, SVDConvergedReason *reason) { SVDCheckSolved(svd); *reason = svd->reason; return PETSC_SUCCESS; }

/* svdsolve.c:313-340 */
PetscErrorCode SVDGetSingularTriplet(SVD svd, PetscInt i, PetscReal *sigma, Vec u, Vec v)
{
  SVDCheckSolved(svd);
  PetscCheck(i >= 0, PETSC_ERR_ARG_OUTOFRANGE, "The index cannot be negative");
  PetscCheck(i < svd->nconv, PETSC_ERR_ARG_OUTOFRANGE, "The index can be nconv-1 at most, see SVDGetConverged()");
  if (sigma) *sigma = svd->sigma[svd->perm[i]];
  if (u || v) {
    if (svd->OP->M < svd->OP->N) { Vec w = u; u = v; v = w; }
    if (u) PetscCall(BVCopyVec(svd->U, svd->perm[i], u));
    if (v) PetscCall(BVCopyVec(svd->V, svd->perm[i], v));
  }
  return PETSC_SUCCESS;
}

/* svdsolve.c:486-580 (standard problem): hypot(||A v - s u||, ||A^T u - s v||), relative: / (s*sqrt(2)) */
PetscErrorCode SVDComputeError(SVD svd, PetscInt i, SVDErrorType type, PetscReal *error)
{
  SVDCheckSolved(svd);
  PetscCheck(error, PETSC_ERR_ARG_NULL, "null error");
  if (!svd->work[0]) {
    PetscCall(MatCreateVecs(svd->OP, &svd->work[1], &svd->work[0]));   /* u: rows, v: cols of OP */
    PetscCall(MatCreateVecs(svd->OP, &svd->work[3], &svd->work[2]));
  }
  Vec u = svd->work[0], v = svd->work[1], x = svd->work[2], y = svd->work[3];
  PetscReal sigma, norm1, norm2;
  PetscCall(SVDGetSingularTriplet(svd, i, &sigma, u, v));
  PetscCall(MatMult(svd->OP, v, x));
  PetscCall(VecAXPY(x, -sigma, u));
  PetscCall(VecNorm(x, NORM_2, &norm1));
  PetscCall(MatMult(svd->OP->M < svd->OP->N ? svd->A : svd->AT, u, y));
  PetscCall(VecAXPY(y, -sigma, v));
  PetscCall(VecNorm(y, NORM_2, &norm2));
  *error = hypot(norm1, norm2);
  if (type == SVD_ERROR_RELATIVE) *error /= sigma * 1.4142135623730951;
  return PETSC_SUCCESS;
}
