/*
 * ksp.c — the linear solves of the spectral transformation: y = B^{-1} w (generalized problem with STSHIFT) and
 * y = (A - sigma B)^{-1} w (STSINVERT), reached from STMatSolve (src/sys/classes/st/interface/stsles.c:166-200) inside
 * STApply_Generic (stsolve.c:16-25).  The reference's default is a direct solve (KSPPREONLY + PCLU, stsles.c:104-131); a sparse
 * factorisation on the GPU is outside this path, so this is Jacobi-preconditioned conjugate gradients on the vectors of the
 * problem's own memory space (device vectors for Mat type b200csr: SpMV k_spmv_sell_pipe, dot/axpy kernels of b2k_bv.cu).
 * Requires a symmetric positive definite coefficient matrix: B of a GHEP, or A - sigma B with sigma below the spectrum.
 */
#include "b2kimpl.h"

PetscErrorCode KSPCreate(KSP *out)
{
  KSP k = (KSP)calloc(1, sizeof(*k));
  PetscCheck(k, PETSC_ERR_MEM, "out of memory");
  k->rtol = 1e-10;                     /* SLEPC_DEFAULT_TOL * 1e-2, stsles.c:118 */
  k->abstol = 1e-50;
  k->max_it = 10000;
  *out = k;
  return PETSC_SUCCESS;
}

static PetscErrorCode KSPReset_Private(KSP k)
{
  PetscCall(VecDestroy(&k->dinv)); PetscCall(VecDestroy(&k->r)); PetscCall(VecDestroy(&k->z));
  PetscCall(VecDestroy(&k->p)); PetscCall(VecDestroy(&k->q));
  k->setup = PETSC_FALSE;
  return PETSC_SUCCESS;
}

PetscErrorCode KSPDestroy(KSP *pk)
{
  if (!pk || !*pk) return PETSC_SUCCESS;
  PetscCall(KSPReset_Private(*pk));
  free(*pk);
  *pk = NULL;
  return PETSC_SUCCESS;
}

PetscErrorCode KSPSetOperators(KSP k, Mat A, Mat P)
{
  (void)P;
  PetscCheck(A && A->M == A->N, PETSC_ERR_ARG_WRONG, "KSP needs a square matrix");
  if (k->A != A) PetscCall(KSPReset_Private(k));
  k->A = A;
  k->setup = PETSC_FALSE;
  return PETSC_SUCCESS;
}

PetscErrorCode KSPSetTolerances(KSP k, PetscReal rtol, PetscReal abstol, PetscReal dtol, PetscInt maxits)
{
  (void)dtol;
  if (rtol != (PetscReal)PETSC_DEFAULT && rtol != (PetscReal)PETSC_CURRENT) { PetscCheck(rtol >= 0.0 && rtol < 1.0, PETSC_ERR_ARG_OUTOFRANGE, "Relative tolerance %g must be non-negative and less than 1.0", rtol); k->rtol = rtol; }
  if (abstol != (PetscReal)PETSC_DEFAULT && abstol != (PetscReal)PETSC_CURRENT) { PetscCheck(abstol >= 0.0, PETSC_ERR_ARG_OUTOFRANGE, "Absolute tolerance %g must be non-negative", abstol); k->abstol = abstol; }
  if (maxits != PETSC_DEFAULT && maxits != PETSC_CURRENT) { PetscCheck(maxits >= 0, PETSC_ERR_ARG_OUTOFRANGE, "Maximum number of iterations %d must be non-negative", maxits); k->max_it = maxits; }
  return PETSC_SUCCESS;
}

PetscErrorCode KSPSetUp(KSP k)
{
  PetscCheck(k->A, PETSC_ERR_ORDER, "KSPSetOperators() must be called first");
  if (k->setup) return PETSC_SUCCESS;
  PetscCall(KSPReset_Private(k));
  PetscCall(MatCreateVecs(k->A, &k->r, &k->dinv));
  PetscCall(MatCreateVecs(k->A, &k->z, &k->p));
  PetscCall(MatCreateVecs(k->A, &k->q, NULL));
  PetscCall(MatGetDiagonal(k->A, k->dinv));       /* PCJACOBI */
  PetscCall(VecReciprocal(k->dinv));
  k->setup = PETSC_TRUE;
  return PETSC_SUCCESS;
}

/* preconditioned conjugate gradients, zero initial guess */
PetscErrorCode KSPSolve(KSP k, Vec b, Vec x)
{
  PetscReal bnorm, rnorm;
  PetscScalar rz, rznew, pq;
  PetscCall(KSPSetUp(k));
  k->its = 0;
  PetscCall(VecSet(x, 0.0));
  PetscCall(VecCopy(b, k->r));
  PetscCall(VecNorm(b, NORM_2, &bnorm));
  if (bnorm == 0.0) return PETSC_SUCCESS;
  const PetscReal target = PetscMax(k->rtol * bnorm, k->abstol);
  PetscCall(VecPointwiseMult(k->z, k->dinv, k->r));
  PetscCall(VecCopy(k->z, k->p));
  PetscCall(VecDot(k->r, k->z, &rz));
  rnorm = bnorm;
  while (rnorm > target && k->its < k->max_it) {
    PetscCall(MatMult(k->A, k->p, k->q));
    PetscCall(VecDot(k->p, k->q, &pq));
    PetscCheck(pq > 0.0, PETSC_ERR_CONV_FAILED, "conjugate gradients met a direction of non-positive curvature (p^T A p = %g): the coefficient matrix of the "
               "ST's linear solve is not positive definite (choose the shift below the spectrum, or B positive definite)", pq);
    const PetscScalar alpha = rz / pq;
    PetscCall(VecAXPY(x, alpha, k->p));
    PetscCall(VecAXPY(k->r, -alpha, k->q));
    PetscCall(VecNorm(k->r, NORM_2, &rnorm));
    k->its++;
    if (rnorm <= target) break;
    PetscCall(VecPointwiseMult(k->z, k->dinv, k->r));
    PetscCall(VecDot(k->r, k->z, &rznew));
    PetscCall(VecAYPX(k->p, rznew / rz, k->z));
    rz = rznew;
  }
  k->total_its += k->its;
  PetscCheck(rnorm <= target, PETSC_ERR_CONV_FAILED, "the linear solve of the spectral transformation did not reach ||r|| <= %g in %d iterations (||r|| = %g)", target, k->max_it, rnorm);
  return PETSC_SUCCESS;
}

PetscErrorCode KSPGetIterationNumber(KSP k, PetscInt *its) { *its = k->its; return PETSC_SUCCESS; }
PetscErrorCode KSPGetTotalIterations(KSP k, PetscInt *its) { *its = k->total_its; return PETSC_SUCCESS; }
