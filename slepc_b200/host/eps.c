/*
 * eps.c — EPS for standard problems with the Krylov-Schur solver: the caller of the hot path.
 * Restates (host control flow only; every n-long operation goes through BV / MatMult)
 *   EPSSetUp_KrylovSchur / EPSSolve_KrylovSchur_Default   src/eps/impls/krylov/krylovschur/krylovschur.c:93-194, 227-336
 *   EPSKrylovConvergence                                   src/eps/impls/krylov/epskrylov.c:207-295
 *   EPSSetDimensions_Default, EPSAllocateSolution          src/eps/interface/epssetup.c:648-676, 692-740
 *   EPSSolve, EPSGetStartVector, EPSGetEigenpair, EPSComputeError   src/eps/interface/epssolve.c:119, 841, 406, 742
 *   EPSConvergedRelative/Absolute, EPSStoppingBasic, EPSComputeVectors_Schur   src/eps/interface/epsdefault.c
 * Spectrum slicing, indefinite, two-sided, BSE, harmonic extraction, arbitrary selection, regions and
 * balancing are outside the named path.
 */
#include "b2kimpl.h"

PetscErrorCode EPSCreate(EPS *out)
{
  EPS eps = (EPS)calloc(1, sizeof(*eps));
  PetscCheck(eps, PETSC_ERR_MEM, "out of memory");
  eps->nev = 1; eps->ncv = PETSC_DETERMINE; eps->mpd = PETSC_DETERMINE; eps->max_it = PETSC_DETERMINE;
  eps->tol = (PetscReal)PETSC_DETERMINE;
  eps->which = (EPSWhich)0;
  eps->problem_type = (EPSProblemType)0;
  eps->conv = EPS_CONV_REL;
  eps->keep = 0.0; eps->lock = PETSC_TRUE;
  eps->purify = PETSC_TRUE;                       /* epsbasic.c:80 */
  PetscCall(STCreate(&eps->st));
  PetscCall(BVCreate(&eps->V));
  PetscCall(DSCreate(&eps->ds));
  *out = eps;
  return PETSC_SUCCESS;
}

PetscErrorCode EPSDestroy(EPS *peps)
{
  if (!peps || !*peps) return PETSC_SUCCESS;
  EPS eps = *peps;
  PetscCall(STDestroy(&eps->st));
  PetscCall(BVDestroy(&eps->V));
  PetscCall(DSDestroy(&eps->ds));
  PetscCall(VecDestroy(&eps->inivec));
  for (int i = 0; i < 5; i++) PetscCall(VecDestroy(&eps->work[i]));
  free(eps->eigr); free(eps->eigi); free(eps->errest); free(eps->perm);
  free(eps);
  *peps = NULL;
  return PETSC_SUCCESS;
}

PetscErrorCode EPSSetOperators(EPS eps, Mat A, Mat B)
{
  PetscCheck(A, PETSC_ERR_ARG_NULL, "null matrix");
  PetscCheck(A->M == A->N, PETSC_ERR_ARG_WRONG, "A is a non-square matrix (%d rows, %d cols)", A->M, A->N);
  if (B) PetscCheck(B->M == B->N && B->M == A->M, PETSC_ERR_ARG_WRONG, "Dimensions of A and B do not match (%d, %d)", A->M, B->M);   /* epssetup.c:446-452 */
  if (eps->V && eps->V->sizes_set && (eps->V->N != A->N || eps->V->n != A->n)) {    /* EPSReset: epssetup.c:441-458 (different dimension) */
    PetscCall(BVForgetSizes_Private(eps->V));
    for (int i = 0; i < 5; i++) PetscCall(VecDestroy(&eps->work[i]));
    PetscCall(VecDestroy(&eps->inivec));
    eps->nini = 0;
  }
  Mat mats[2] = {A, B};
  PetscCall(STSetMatrices(eps->st, B ? 2 : 1, mats));
  eps->B = B;
  eps->n = A->N; eps->nloc = A->n;
  eps->setup_done = PETSC_FALSE; eps->solved = PETSC_FALSE; eps->started = PETSC_FALSE;
  return PETSC_SUCCESS;
}

PetscErrorCode EPSSetProblemType(EPS eps, EPSProblemType type)
{
  PetscCheck(type == EPS_HEP || type == EPS_NHEP || type == EPS_GHEP, PETSC_ERR_SUP, "EPS_HEP, EPS_GHEP (B symmetric positive definite) and EPS_NHEP are available");
  eps->problem_type = type;
  eps->ishermitian = (type == EPS_HEP || type == EPS_GHEP) ? PETSC_TRUE : PETSC_FALSE;
  eps->isgeneralized = (type == EPS_GHEP) ? PETSC_TRUE : PETSC_FALSE;
  eps->setup_done = PETSC_FALSE;
  return PETSC_SUCCESS;
}

PetscErrorCode EPSSetType(EPS eps, const char *type)
{
  (void)eps;
  PetscCheck(!strcmp(type, EPSKRYLOVSCHUR), PETSC_ERR_ARG_UNKNOWN_TYPE, "Unknown EPS type given: %s (this build provides krylovschur)", type);
  return PETSC_SUCCESS;
}

PetscErrorCode EPSSetDimensions(EPS eps, PetscInt nev, PetscInt ncv, PetscInt mpd)
{
  if (nev != PETSC_CURRENT) { PetscCheck(nev > 0, PETSC_ERR_ARG_OUTOFRANGE, "Illegal value of nev. Must be > 0"); eps->nev = nev; }
  if (ncv == PETSC_DETERMINE || ncv == PETSC_DEFAULT) eps->ncv = PETSC_DETERMINE;
  else if (ncv != PETSC_CURRENT) { PetscCheck(ncv > 0, PETSC_ERR_ARG_OUTOFRANGE, "Illegal value of ncv. Must be > 0"); eps->ncv = ncv; }
  if (mpd == PETSC_DETERMINE || mpd == PETSC_DEFAULT) eps->mpd = PETSC_DETERMINE;
  else if (mpd != PETSC_CURRENT) { PetscCheck(mpd > 0, PETSC_ERR_ARG_OUTOFRANGE, "Illegal value of mpd. Must be > 0"); eps->mpd = mpd; }
  eps->setup_done = PETSC_FALSE;
  return PETSC_SUCCESS;
}
PetscErrorCode EPSGetDimensions(EPS eps, PetscInt *nev, PetscInt *ncv, PetscInt *mpd)
{
  if (nev) *nev = eps->nev;
  if (ncv) *ncv = eps->ncv;
  if (mpd) *mpd = eps->mpd;
  return PETSC_SUCCESS;
}

PetscErrorCode EPSSetTolerances(EPS eps, PetscReal tol, PetscInt maxits)
{
  if (tol == (PetscReal)PETSC_DETERMINE || tol == (PetscReal)PETSC_DEFAULT) eps->tol = (PetscReal)PETSC_DETERMINE;
  else if (tol != (PetscReal)PETSC_CURRENT) { PetscCheck(tol > 0.0, PETSC_ERR_ARG_OUTOFRANGE, "Illegal value of tol. Must be > 0"); eps->tol = tol; }
  if (maxits == PETSC_DETERMINE || maxits == PETSC_DEFAULT) eps->max_it = PETSC_DETERMINE;
  else if (maxits != PETSC_CURRENT) { PetscCheck(maxits > 0, PETSC_ERR_ARG_OUTOFRANGE, "Illegal value of maxits. Must be > 0"); eps->max_it = maxits; }
  eps->setup_done = PETSC_FALSE;
  return PETSC_SUCCESS;
}
PetscErrorCode EPSGetTolerances(EPS eps, PetscReal *tol, PetscInt *maxits) { if (tol) *tol = eps->tol; if (maxits) *maxits = eps->max_it; return PETSC_SUCCESS; }

PetscErrorCode EPSSetWhichEigenpairs(EPS eps, EPSWhich which)
{
  PetscCheck(which >= EPS_LARGEST_MAGNITUDE && which <= EPS_TARGET_REAL, PETSC_ERR_ARG_OUTOFRANGE, "Invalid 'which' value");
  eps->which = which;
  eps->setup_done = PETSC_FALSE;
  return PETSC_SUCCESS;
}
/* epsopts.c:882-903: the target is also the default shift of the ST (STSetDefaultShift) */
PetscErrorCode EPSSetTarget(EPS eps, PetscScalar target)
{
  eps->target = target;
  PetscCall(STSetShift(eps->st, target));
  eps->setup_done = PETSC_FALSE;
  return PETSC_SUCCESS;
}
PetscErrorCode EPSSetPurify(EPS eps, PetscBool purify) { eps->purify = purify; eps->setup_done = PETSC_FALSE; return PETSC_SUCCESS; }
PetscErrorCode EPSSetConvergenceTest(EPS eps, EPSConv conv) { eps->conv = conv; return PETSC_SUCCESS; }
PetscErrorCode EPSKrylovSchurSetRestart(EPS eps, PetscReal keep)
{
  if (keep == (PetscReal)PETSC_DEFAULT || keep == (PetscReal)PETSC_DECIDE) eps->keep = 0.5;
  else { PetscCheck(keep >= 0.1 && keep <= 0.9, PETSC_ERR_ARG_OUTOFRANGE, "The keep argument %g must be in the range [0.1,0.9]", keep); eps->keep = keep; }
  return PETSC_SUCCESS;
}
PetscErrorCode EPSKrylovSchurSetLocking(EPS eps, PetscBool lock) { eps->lock = lock; return PETSC_SUCCESS; }
PetscErrorCode EPSMonitorSet(EPS eps, EPSMonitorFn monitor, void *ctx) { eps->monitor = monitor; eps->monitorctx = ctx; return PETSC_SUCCESS; }
PetscErrorCode EPSGetBV(EPS eps, BV *V) { *V = eps->V; return PETSC_SUCCESS; }
PetscErrorCode EPSGetDS(EPS eps, DS *ds) { *ds = eps->ds; return PETSC_SUCCESS; }
PetscErrorCode EPSGetST(EPS eps, ST *st) { *st = eps->st; return PETSC_SUCCESS; }

/* epssetup.c:592: Krylov solvers use the first vector of the initial space as the start vector */
PetscErrorCode EPSSetInitialSpace(EPS eps, PetscInt n, Vec is[])
{
  PetscCheck(n >= 0, PETSC_ERR_ARG_OUTOFRANGE, "Argument n cannot be negative");
  PetscCall(VecDestroy(&eps->inivec));
  eps->nini = 0;
  if (n > 0) {
    PetscCheck(is && is[0], PETSC_ERR_ARG_NULL, "null vector");
    if (is[0]->mem == B2K_MEM_DEVICE) PetscCall(VecCreateB200(is[0]->n, is[0]->N, &eps->inivec));
    else PetscCall(VecCreateHost(is[0]->n, is[0]->N, &eps->inivec));
    PetscCall(VecCopy(is[0], eps->inivec));
    eps->nini = 1;
  }
  eps->setup_done = PETSC_FALSE;
  return PETSC_SUCCESS;
}

/* comparator seen by DS: SlepcSCCompare with map = STBackTransform (slepcsc.c:40-63) */
static PetscErrorCode EPSCompare_Private(PetscScalar ar, PetscScalar ai, PetscScalar br, PetscScalar bi, PetscInt *res, void *ctx)
{
  EPS eps = (EPS)ctx;                              /* SlepcMap_ST: the values are compared after STBackTransform, with the live shift */
  PetscCall(STBackTransform(eps->st, 1, &ar, &ai));
  PetscCall(STBackTransform(eps->st, 1, &br, &bi));
  return eps->sc.fn(ar, ai, br, bi, res, eps->sc.ctx);
}

PetscErrorCode EPSSetUp(EPS eps)
{
  if (eps->setup_done) return PETSC_SUCCESS;
  PetscCheck(eps->st->A, PETSC_ERR_ARG_WRONGSTATE, "EPSSetOperators() must be called first");
  if (!eps->problem_type) PetscCall(EPSSetProblemType(eps, eps->B ? EPS_GHEP : EPS_NHEP));   /* epssetup.c:311-318 (GNHEP is outside this path) */
  PetscCheck(!eps->B || eps->isgeneralized, PETSC_ERR_ARG_INCOMP, "Inconsistent EPS state: the problem type does not match the number of matrices");   /* epssetup.c:319 */
  PetscCheck(eps->B || !eps->isgeneralized, PETSC_ERR_ARG_INCOMP, "Inconsistent EPS state: the problem type does not match the number of matrices");
  if (eps->tol == (PetscReal)PETSC_DETERMINE) eps->tol = SLEPC_DEFAULT_TOL;
  PetscCall(STSetUp(eps->st));
  /* EPSSetUp_KrylovSchur krylovschur.c:108-116 */
  PetscInt ncv = eps->ncv, mpd = eps->mpd;
  if (ncv != PETSC_DETERMINE) {
    PetscCheck(ncv >= eps->nev + 1 || (ncv == eps->nev && ncv == eps->n), PETSC_ERR_USER_INPUT, "The value of ncv must be at least nev+1");
  } else if (mpd != PETSC_DETERMINE) ncv = PetscMin(eps->n, eps->nev + mpd);
  else {
    if (eps->nev < 500) ncv = PetscMin(eps->n, PetscMax(2 * eps->nev, eps->nev + 15));
    else { mpd = 500; ncv = PetscMin(eps->n, eps->nev + mpd); }
  }
  if (mpd == PETSC_DETERMINE) mpd = ncv;
  eps->ncv = ncv; eps->mpd = mpd;
  PetscCheck(eps->ncv <= eps->nev + eps->mpd, PETSC_ERR_USER_INPUT, "The value of ncv must not be larger than nev+mpd");
  if (eps->max_it == PETSC_DETERMINE) eps->max_it = PetscMax(100, 2 * eps->n / eps->ncv);
  if (!eps->which) eps->which = eps->st->sinvert ? EPS_TARGET_MAGNITUDE : EPS_LARGEST_MAGNITUDE;   /* EPSSetWhichEigenpairs_Default epsdefault.c:210-220 */
  PetscCheck(eps->lock || eps->mpd >= eps->ncv, PETSC_ERR_SUP, "Should not use mpd parameter in non-locking variant");
  if (!eps->keep) eps->keep = 0.5;

  switch (eps->which) {
  case EPS_LARGEST_MAGNITUDE: eps->sc.fn = SlepcCompareLargestMagnitude; eps->sc.ctx = NULL; break;
  case EPS_SMALLEST_MAGNITUDE: eps->sc.fn = SlepcCompareSmallestMagnitude; eps->sc.ctx = NULL; break;
  case EPS_LARGEST_REAL: eps->sc.fn = SlepcCompareLargestReal; eps->sc.ctx = NULL; break;
  case EPS_SMALLEST_REAL: eps->sc.fn = SlepcCompareSmallestReal; eps->sc.ctx = NULL; break;
  case EPS_LARGEST_IMAGINARY: eps->sc.fn = SlepcCompareLargestImaginary; eps->sc.ctx = NULL; break;
  case EPS_SMALLEST_IMAGINARY: eps->sc.fn = SlepcCompareSmallestImaginary; eps->sc.ctx = NULL; break;
  case EPS_TARGET_MAGNITUDE: eps->sc.fn = SlepcCompareTargetMagnitude; eps->sc.ctx = &eps->target; break;
  case EPS_TARGET_REAL: eps->sc.fn = SlepcCompareTargetReal; eps->sc.ctx = &eps->target; break;
  }
  PetscCall(STGetShift(eps->st, &eps->sc.sigma));
  PetscCall(DSSetEigenvalueComparison(eps->ds, EPSCompare_Private, eps));

  /* EPSAllocateSolution(eps,1) epssetup.c:692-740 */
  const PetscInt requested = eps->ncv + 1;
  if (requested != eps->allocated) {
    free(eps->eigr); free(eps->eigi); free(eps->errest); free(eps->perm);
    eps->eigr = (PetscScalar *)calloc((size_t)requested, sizeof(PetscScalar));
    eps->eigi = (PetscScalar *)calloc((size_t)requested, sizeof(PetscScalar));
    eps->errest = (PetscReal *)calloc((size_t)requested, sizeof(PetscReal));
    eps->perm = (PetscInt *)calloc((size_t)requested, sizeof(PetscInt));
    PetscCheck(eps->eigr && eps->eigi && eps->errest && eps->perm, PETSC_ERR_MEM, "out of memory");
    eps->allocated = requested;
  }
  PetscInt oldsize = 0;
  if (eps->V->sizes_set) oldsize = eps->V->m;
  if (!oldsize) {
    if (!eps->V->ctor) PetscCall(BVSetType(eps->V, BVB200));                      /* reference default is BVMAT (epssetup.c:718); this build has one type */
    Vec t;
    PetscCall(MatCreateVecs(eps->st->A, &t, NULL));
    PetscErrorCode ierr = BVSetSizesFromVec(eps->V, t, requested);
    PetscCall(VecDestroy(&t));
    PetscCall(ierr);
  } else if (oldsize != requested) PetscCall(BVResize(eps->V, requested, PETSC_FALSE));

  /* generalized symmetric-definite problem: B-inner product in the basis (epssetup.c:372-381); purification only makes sense there */
  if (eps->isgeneralized && eps->ishermitian) { Mat B; PetscCall(STGetBilinearForm(eps->st, &B)); PetscCall(BVSetMatrix(eps->V, B, PETSC_FALSE)); }
  else PetscCall(BVSetMatrix(eps->V, NULL, PETSC_FALSE));
  if (!eps->isgeneralized) eps->purify = PETSC_FALSE;                            /* epssetup.c:359-369 */

  /* DS: krylovschur.c:153-168 */
  if (eps->ishermitian) {
    PetscCall(DSSetType(eps->ds, DSHEP));
    PetscCall(DSSetCompact(eps->ds, PETSC_TRUE));
  } else {
    PetscCall(DSSetType(eps->ds, DSNHEP));
    PetscCall(DSSetCompact(eps->ds, PETSC_FALSE));
  }
  PetscCall(DSSetExtraRow(eps->ds, PETSC_TRUE));
  PetscCall(DSAllocate(eps->ds, eps->ncv + 1));
  if (eps->nini) {
    PetscCheck(eps->inivec->n == eps->V->n, PETSC_ERR_ARG_INCOMP, "the initial vector has local size %d, the problem %d", eps->inivec->n, eps->V->n);
    PetscCall(BVInsertVec(eps->V, 0, eps->inivec));
  }
  eps->setup_done = PETSC_TRUE;
  eps->solved = PETSC_FALSE; eps->started = PETSC_FALSE; eps->vectors_done = PETSC_FALSE;
  return PETSC_SUCCESS;
}

/* epssolve.c:841-871 */
static PetscErrorCode EPSGetStartVector(EPS eps, PetscInt i, PetscBool *breakdown)
{
  PetscReal norm;
  PetscBool lindep;
  if (i > 0 || eps->nini == 0) PetscCall(BVSetRandomColumn(eps->V, i));
  if (eps->isgeneralized && eps->ishermitian) {   /* force the vector into the range of OP for definite generalized problems, epssolve.c:855-862 */
    Vec w, z;
    PetscCall(BVCreateVec(eps->V, &w));
    PetscErrorCode ierr = BVCopyVec(eps->V, i, w);
    if (!ierr) {
      PetscCall(BVGetColumn(eps->V, i, &z));
      ierr = STApply(eps->st, w, z);
      PetscCall(BVRestoreColumn(eps->V, i, &z));
      eps->V->state++;
    }
    PetscCall(VecDestroy(&w));
    PetscCall(ierr);
  }
  PetscCall(BVOrthogonalizeColumn(eps->V, i, NULL, &norm, &lindep));
  if (breakdown) *breakdown = lindep;
  else if (lindep || norm == 0.0) {
    PetscCheck(i, PETSC_ERR_PLIB, "Initial vector is zero or belongs to the deflation space");
    PetscCheck(!i, PETSC_ERR_CONV_FAILED, "Unable to generate more start vectors");
  }
  PetscCall(BVScaleColumn(eps->V, i, 1.0 / norm));
  return PETSC_SUCCESS;
}

/* epskrylov.c:207-295 (trivial region, no true residual, single-sided) */
static PetscErrorCode EPSKrylovConvergence(EPS eps, PetscInt kini, PetscInt nits, PetscReal beta, PetscInt *kout)
{
  PetscInt k, newk, marker = -1;
  PetscReal resnorm;
  for (k = kini; k < kini + nits; k++) {
    PetscScalar re = eps->eigr[k], im = eps->eigi[k];
    if (!eps->st->sinvert) PetscCall(STBackTransform(eps->st, 1, &re, &im));   /* only for STSHIFT, epskrylov.c:246; sinvert tests the transformed value */
    newk = k;
    PetscCall(DSVectors(eps->ds, DS_MAT_X, &newk, &resnorm));
    resnorm *= beta;
    if (eps->conv == EPS_CONV_REL) {                    /* EPSConvergedRelative epsdefault.c:224 */
      const PetscReal w = SlepcAbsEigenvalue(re, im);
      eps->errest[k] = (w != 0.0) ? resnorm / w : PETSC_MAX_REAL;
    } else eps->errest[k] = resnorm;                    /* EPSConvergedAbsolute */
    if (marker == -1 && eps->errest[k] >= eps->tol) marker = k;
    if (newk == k + 1) { eps->errest[k + 1] = eps->errest[k]; k++; }
    if (marker != -1) break;
  }
  if (marker != -1) k = marker;
  *kout = k;
  return PETSC_SUCCESS;
}

static PetscErrorCode EPSSolveBegin_Private(EPS eps)
{
  PetscCall(EPSSetUp(eps));
  eps->nconv = 0; eps->its = 0;
  for (PetscInt i = 0; i < eps->ncv; i++) { eps->eigr[i] = 0.0; eps->eigi[i] = 0.0; eps->errest[i] = 0.0; eps->perm[i] = i; }
  eps->reason = EPS_CONVERGED_ITERATING;
  PetscCall(EPSGetStartVector(eps, 0, NULL));
  eps->l = 0;
  eps->started = PETSC_TRUE; eps->solved = PETSC_FALSE; eps->vectors_done = PETSC_FALSE;
  return PETSC_SUCCESS;
}

/* one pass of the restart loop of EPSSolve_KrylovSchur_Default, krylovschur.c:248-332 */
static PetscErrorCode EPSKrylovSchurCycle_Private(EPS eps)
{
  PetscInt k = eps->nconv, l = eps->l, nv, nconv;
  PetscReal beta = 0.0;
  PetscBool breakdown;
  Mat Op, T, U;
  const PetscBool hermitian = eps->ishermitian;

  eps->its++;
  nv = PetscMin(eps->nconv + eps->mpd, eps->ncv);
  PetscCall(DSSetDimensions(eps->ds, nv, eps->nconv, eps->nconv + l));
  PetscCall(STGetOperator(eps->st, &Op));
  PetscCall(DSGetMat(eps->ds, hermitian ? DS_MAT_T : DS_MAT_A, &T));
  PetscErrorCode ierr = hermitian ? BVMatLanczos(eps->V, Op, T, eps->nconv + l, &nv, &beta, &breakdown)
                                  : BVMatArnoldi(eps->V, Op, T, eps->nconv + l, &nv, &beta, &breakdown);
  PetscCall(DSRestoreMat(eps->ds, hermitian ? DS_MAT_T : DS_MAT_A, &T));
  PetscCall(ierr);
  PetscCall(STRestoreOperator(eps->st, &Op));
  PetscCall(DSSetDimensions(eps->ds, nv, eps->nconv, eps->nconv + l));
  PetscCall(DSSetState(eps->ds, l ? DS_STATE_RAW : DS_STATE_INTERMEDIATE));
  PetscCall(BVSetActiveColumns(eps->V, eps->nconv, nv));

  PetscCall(DSSolve(eps->ds, eps->eigr, eps->eigi));
  PetscCall(DSSort(eps->ds, eps->eigr, eps->eigi, NULL, NULL, NULL));
  PetscCall(DSUpdateExtraRow(eps->ds));
  PetscCall(DSSynchronize(eps->ds, eps->eigr, eps->eigi));
  eps->n_dssolve++;

  PetscCall(EPSKrylovConvergence(eps, eps->nconv, nv - eps->nconv, beta, &k));
  /* EPSStoppingBasic epsdefault.c:290-302 */
  eps->reason = EPS_CONVERGED_ITERATING;
  if (k >= eps->nev) eps->reason = EPS_CONVERGED_TOL;
  else if (eps->its >= eps->max_it) eps->reason = EPS_DIVERGED_ITS;
  nconv = k;

  if (eps->reason != EPS_CONVERGED_ITERATING || breakdown || k == nv) l = 0;
  else {
    l = PetscMax(1, (PetscInt)((nv - k) * eps->keep));
    if (!hermitian) PetscCall(DSGetTruncateSize(eps->ds, k, nv, &l));
  }
  if (!eps->lock && l > 0) { l += k; k = 0; }     /* non-locking variant: reset no. of converged pairs */

  if (eps->reason == EPS_CONVERGED_ITERATING) {
    if (breakdown || k == nv) {                   /* start a new factorization */
      if (k < eps->nev) {
        PetscCall(EPSGetStartVector(eps, k, &breakdown));
        if (breakdown) eps->reason = EPS_DIVERGED_BREAKDOWN;
      }
    } else PetscCall(DSTruncate(eps->ds, k + l, PETSC_FALSE));   /* prepare the Rayleigh quotient for restart */
  }
  /* update the corresponding vectors V(:,idx) = V*Q(:,idx) */
  PetscCall(DSGetMat(eps->ds, DS_MAT_Q, &U));
  ierr = BVMultInPlace(eps->V, U, eps->nconv, k + l);
  PetscCall(DSRestoreMat(eps->ds, DS_MAT_Q, &U));
  PetscCall(ierr);
  if (eps->reason == EPS_CONVERGED_ITERATING && !breakdown) PetscCall(BVCopyColumn(eps->V, nv, k + l));
  eps->nconv = k;
  eps->l = l;
  if (eps->monitor) PetscCall(eps->monitor(eps, eps->its, nconv, eps->eigr, eps->eigi, eps->errest, nv, eps->monitorctx));
  return PETSC_SUCCESS;
}

/* SlepcSortEigenvalues slepcsc.c:89-140 (real arithmetic: conjugate pairs stay together) */
static PetscErrorCode EPSSortEigenvalues_Private(EPS eps, PetscInt n, PetscScalar *eigr, PetscScalar *eigi, PetscInt *perm)
{
  PetscScalar re, im;
  PetscInt i, j, result, tmp;
  for (i = n - 1; i >= 0; i--) {
    re = eigr[perm[i]]; im = eigi[perm[i]];
    j = i + 1;
    if (im != 0) { i--; im = eigi[perm[i]]; }      /* complex eigenvalue */
    while (j < n) {
      PetscCall(eps->sc.fn(re, im, eigr[perm[j]], eigi[perm[j]], &result, eps->sc.ctx));   /* eps->sc has no map: the values are already back-transformed, epssolve.c:178 */
      if (result <= 0) break;
      if (!im) {
        if (eigi[perm[j]] == 0.0) { tmp = perm[j - 1]; perm[j - 1] = perm[j]; perm[j] = tmp; j++; }
        else { tmp = perm[j - 1]; perm[j - 1] = perm[j]; perm[j] = perm[j + 1]; perm[j + 1] = tmp; j += 2; }
      } else {
        if (eigi[perm[j]] == 0.0) { tmp = perm[j - 2]; perm[j - 2] = perm[j]; perm[j] = perm[j - 1]; perm[j - 1] = tmp; j++; }
        else {
          tmp = perm[j - 2]; perm[j - 2] = perm[j]; perm[j] = tmp;
          tmp = perm[j - 1]; perm[j - 1] = perm[j + 1]; perm[j + 1] = tmp;
          j += 2;
        }
      }
    }
  }
  return PETSC_SUCCESS;
}

/* EPSComputeVectors: Hermitian = the Lanczos/Schur vectors as they are (epsdefault.c:28);
   non-Hermitian = EPSComputeVectors_Schur (epsdefault.c:105-125): X = V * Z with T Z = Z D */
static PetscErrorCode EPSComputeVectors(EPS eps)
{
  if (eps->vectors_done) return PETSC_SUCCESS;
  eps->vectors_done = PETSC_TRUE;
  if (eps->nconv == 0) return PETSC_SUCCESS;
  if (eps->ishermitian) {                         /* EPSComputeVectors_Hermitian epsdefault.c:28-50 */
    if (eps->purify) {                            /* EPS_Purify epsimpl.h:297-312: x <- OP x, then B-normalise */
      Vec v, z;
      PetscCall(BVCreateVec(eps->V, &v));
      PetscErrorCode ierr = PETSC_SUCCESS;
      for (PetscInt i = 0; i < eps->nconv && !ierr; i++) {
        ierr = BVCopyVec(eps->V, i, v);
        if (ierr) break;
        PetscCall(BVGetColumn(eps->V, i, &z));
        ierr = STApply(eps->st, v, z);
        PetscCall(BVRestoreColumn(eps->V, i, &z));
        eps->V->state++;
      }
      PetscCall(VecDestroy(&v));
      PetscCall(ierr);
      PetscCall(BVSetActiveColumns(eps->V, 0, eps->nconv));
      PetscCall(BVNormalize(eps->V, NULL));
    }
    return PETSC_SUCCESS;
  }
  Mat Z;
  PetscCall(DSVectors(eps->ds, DS_MAT_X, NULL, NULL));
  PetscCall(DSGetMat(eps->ds, DS_MAT_X, &Z));
  PetscErrorCode ierr = BVMultInPlace(eps->V, Z, 0, eps->nconv);
  PetscCall(DSRestoreMat(eps->ds, DS_MAT_X, &Z));
  PetscCall(ierr);
  return PETSC_SUCCESS;
}

static PetscErrorCode EPSSolveEnd_Private(EPS eps)
{
  PetscCall(DSTruncate(eps->ds, eps->nconv, PETSC_TRUE));               /* krylovschur.c:335 */
  PetscCheck(eps->reason, PETSC_ERR_PLIB, "Internal error, solver returned without setting converged reason");
  eps->solved = PETSC_TRUE; eps->started = PETSC_FALSE;
  PetscCall(BVSetActiveColumns(eps->V, 0, eps->nconv));                  /* epssolve.c:151 */
  PetscCall(STBackTransform(eps->st, eps->nconv, eps->eigr, eps->eigi)); /* EPSComputeValues */
  for (PetscInt i = 0; i < eps->nconv - 1; i++) {                        /* conjugate pairs: positive imaginary part first, :162-175 */
    if (eps->eigi[i] != 0) {
      if (eps->eigi[i] < 0) {
        eps->eigi[i] = -eps->eigi[i];
        eps->eigi[i + 1] = -eps->eigi[i + 1];
        PetscCall(EPSComputeVectors(eps));
        PetscCall(BVScaleColumn(eps->V, i + 1, -1.0));
      }
      i++;
    }
  }
  for (PetscInt i = 0; i < eps->nconv; i++) eps->perm[i] = i;
  PetscCall(EPSSortEigenvalues_Private(eps, eps->nconv, eps->eigr, eps->eigi, eps->perm));
  return PETSC_SUCCESS;
}

PetscErrorCode EPSSolve(EPS eps)
{
  if (eps->solved && eps->setup_done) return PETSC_SUCCESS;
  if (!eps->started || !eps->setup_done) PetscCall(EPSSolveBegin_Private(eps));
  while (eps->reason == EPS_CONVERGED_ITERATING) PetscCall(EPSKrylovSchurCycle_Private(eps));
  PetscCall(EPSSolveEnd_Private(eps));
  return PETSC_SUCCESS;
}

/* bench/profiling entry: advance the solve by at most `nrestarts` restart cycles (a "step" of bench.py) */
PetscErrorCode EPSSolveRestartCycles(EPS eps, PetscInt nrestarts, PetscInt *done)
{
  PetscInt n = 0;
  if (eps->solved && eps->setup_done) { if (done) *done = 0; return PETSC_SUCCESS; }
  if (!eps->started || !eps->setup_done) PetscCall(EPSSolveBegin_Private(eps));
  while (eps->reason == EPS_CONVERGED_ITERATING && n < nrestarts) { PetscCall(EPSKrylovSchurCycle_Private(eps)); n++; }
  if (eps->reason != EPS_CONVERGED_ITERATING) PetscCall(EPSSolveEnd_Private(eps));
  if (done) *done = n;
  return PETSC_SUCCESS;
}

#define EPSCheckSolved(eps) PetscCheck((eps)->solved, PETSC_ERR_ARG_WRONGSTATE, "Must call EPSSolve() first")

PetscErrorCode EPSGetIterationNumber(EPS eps, PetscInt *its) { *its = eps->its; return PETSC_SUCCESS; }
PetscErrorCode EPSGetConverged(EPS eps, PetscInt *nconv) { EPSCheckSolved(eps); *nconv = eps->nconv; return PETSC_SUCCESS; }
PetscErrorCode EPSGetConvergedReason(EPS eps, EPSConvergedReason *reason) { EPSCheckSolved(eps); *reason = eps->reason; return PETSC_SUCCESS; }

PetscErrorCode EPSGetEigenvalue(EPS eps, PetscInt i, PetscScalar *eigr, PetscScalar *eigi)
{
  EPSCheckSolved(eps);
  PetscCheck(i >= 0, PETSC_ERR_ARG_OUTOFRANGE, "The index cannot be negative");
  PetscCheck(i < eps->nconv, PETSC_ERR_ARG_OUTOFRANGE, "The index can be nconv-1 at most, see EPSGetConverged()");
  const PetscInt k = eps->perm[i];
  if (eigr) *eigr = eps->eigr[k];
  if (eigi) *eigi = eps->eigi[k];
  return PETSC_SUCCESS;
}

PetscErrorCode EPSGetEigenpair(EPS eps, PetscInt i, PetscScalar *eigr, PetscScalar *eigi, Vec Vr, Vec Vi)
{
  PetscCall(EPSGetEigenvalue(eps, i, eigr, eigi));
  if (Vr || Vi) {
    PetscCall(EPSComputeVectors(eps));
    const PetscInt k = eps->perm[i];
    const PetscScalar im = eps->eigi[k];            /* BV_GetEigenvector bvimpl.h:421-444 */
    if (im > 0.0) {
      if (Vr) PetscCall(BVCopyVec(eps->V, k, Vr));
      if (Vi) PetscCall(BVCopyVec(eps->V, k + 1, Vi));
    } else if (im < 0.0) {
      if (Vr) PetscCall(BVCopyVec(eps->V, k - 1, Vr));
      if (Vi) { PetscCall(BVCopyVec(eps->V, k, Vi)); PetscCall(VecScale(Vi, -1.0)); }
    } else {
      if (Vr) PetscCall(BVCopyVec(eps->V, k, Vr));
      if (Vi) PetscCall(VecSet(Vi, 0.0));
    }
  }
  return PETSC_SUCCESS;
}

PetscErrorCode EPSGetErrorEstimate(EPS eps, PetscInt i, PetscReal *errest)
{
  EPSCheckSolved(eps);
  PetscCheck(i >= 0 && i < eps->nconv, PETSC_ERR_ARG_OUTOFRANGE, "The index can be nconv-1 at most, see EPSGetConverged()");
  *errest = eps->errest[eps->perm[i]];
  return PETSC_SUCCESS;
}

/* EPSComputeResidualNorm_Private epssolve.c:666-722 (standard problem) */
static PetscErrorCode EPSComputeResidualNorm_Private(EPS eps, PetscScalar kr, PetscScalar ki, Vec xr, Vec xi, Vec *z, PetscReal *norm)
{
  Mat A = eps->st->A, B = eps->st->B;
  Vec u = z[0], v = z[1], w = z[2];
  PetscReal ni, nr;
  if (ki == 0 || fabs(ki) < fabs(kr * PETSC_MACHINE_EPSILON)) {
    PetscCall(MatMult(A, xr, u));                 /* u = A x */
    if (fabs(kr) > PETSC_MACHINE_EPSILON) {
      if (B) { PetscCall(MatMult(B, xr, w)); PetscCall(VecAXPY(u, -kr, w)); }   /* u = A x - k B x */
      else PetscCall(VecAXPY(u, -kr, xr));
    }
    PetscCall(VecNorm(u, NORM_2, norm));
  } else {
    PetscCheck(!B, PETSC_ERR_SUP, "complex eigenvalues of a generalized problem are outside this path");
    PetscCall(MatMult(A, xr, u));
    if (SlepcAbsEigenvalue(kr, ki) > PETSC_MACHINE_EPSILON) {
      PetscCall(VecCopy(xr, v));
      PetscCall(VecAXPY(u, -kr, v));
      PetscCall(VecCopy(xi, w));
      PetscCall(VecAXPY(u, ki, w));
    }
    PetscCall(VecNorm(u, NORM_2, &nr));
    PetscCall(MatMult(A, xi, u));
    if (SlepcAbsEigenvalue(kr, ki) > PETSC_MACHINE_EPSILON) {
      PetscCall(VecAXPY(u, -kr, w));
      PetscCall(VecAXPY(u, -ki, v));
    }
    PetscCall(VecNorm(u, NORM_2, &ni));
    *norm = SlepcAbsEigenvalue(nr, ni);
  }
  return PETSC_SUCCESS;
}

/* epssolve.c:742-815 */
PetscErrorCode EPSComputeError(EPS eps, PetscInt i, EPSErrorType type, PetscReal *error)
{
  EPSCheckSolved(eps);
  PetscCheck(error, PETSC_ERR_ARG_NULL, "null error");
  for (int q = 0; q < 5; q++) if (!eps->work[q]) PetscCall(BVCreateVec(eps->V, &eps->work[q]));
  Vec xr = eps->work[0], xi = eps->work[3], w[3] = {eps->work[1], eps->work[4], eps->work[2]};
  PetscScalar kr, ki;
  PetscCall(EPSGetEigenpair(eps, i, &kr, &ki, xr, xi));
  PetscCall(EPSComputeResidualNorm_Private(eps, kr, ki, xr, xi, w, error));
  PetscReal vecnorm = 1.0;
  if (eps->problem_type == EPS_GHEP) PetscCall(VecNorm(xr, NORM_2, &vecnorm));   /* eigenvectors are B-normalised: epssolve.c:776 */
  switch (type) {
  case EPS_ERROR_ABSOLUTE: break;
  case EPS_ERROR_RELATIVE: *error /= SlepcAbsEigenvalue(kr, ki) * vecnorm; break;
  default: SETERRQ(PETSC_ERR_SUP, "backward errors need a matrix norm operation, which this Mat type does not provide");
  }
  return PETSC_SUCCESS;
}
