/*
 * st.c — spectral transformation: STSHIFT (src/sys/classes/st/impls/shift/shift.c) and STSINVERT
 * (impls/sinvert/sinvert.c) for standard and generalized problems.
 *
 *   STSHIFT,   A x = k x      : Op = A - sigma I                  (a bare MatMult for sigma = 0: STMatMAXPY_Private returns A
 *                                                                   itself, stsolve.c:607-610; STApply_Generic stsolve.c:16-25)
 *   STSHIFT,   A x = k B x    : Op = B^{-1} (A - sigma B)         (T[0] = A - sigma B multiplied, KSP on B: shift.c:60-114)
 *   STSINVERT, A x = k [B] x  : Op = (A - sigma B)^{-1} [B]       (T[0] = B multiplied, KSP on A - sigma B: sinvert.c:79-150)
 * STApply = multiply by T[0] then solve with the KSP matrix (STApply_Generic); back-transformation k = theta + sigma
 * (shift.c:49-58) or k = sigma + 1/theta (sinvert.c:51-77).  The linear solves are Jacobi-preconditioned CG (ksp.c).
 */
#include "b2kimpl.h"

PetscErrorCode STCreate(ST *out)
{
  ST st = (ST)calloc(1, sizeof(*st));
  PetscCheck(st, PETSC_ERR_MEM, "out of memory");
  *out = st;
  return PETSC_SUCCESS;
}

static PetscErrorCode STReset_Private(ST st)
{
  if (st->Op) PetscCall(MatDestroy(&st->Op));
  if (st->T) PetscCall(MatDestroy(&st->T));
  PetscCall(VecDestroy(&st->w));
  st->setup = PETSC_FALSE;
  return PETSC_SUCCESS;
}

PetscErrorCode STDestroy(ST *pst)
{
  if (!pst || !*pst) return PETSC_SUCCESS;
  ST st = *pst;
  PetscCall(STReset_Private(st));
  PetscCall(KSPDestroy(&st->ksp));
  free(st);
  *pst = NULL;
  return PETSC_SUCCESS;
}

PetscErrorCode STSetType(ST st, const char *type)
{
  PetscCheck(type, PETSC_ERR_ARG_NULL, "null type");
  if (!strcmp(type, STSHIFT)) st->sinvert = PETSC_FALSE;
  else if (!strcmp(type, STSINVERT)) st->sinvert = PETSC_TRUE;
  else SETERRQ(PETSC_ERR_ARG_UNKNOWN_TYPE, "Unable to find requested ST type %s (this build provides shift and sinvert)", type);
  st->setup = PETSC_FALSE;
  return PETSC_SUCCESS;
}
PetscErrorCode STGetType(ST st, const char **type) { *type = st->sinvert ? STSINVERT : STSHIFT; return PETSC_SUCCESS; }

PetscErrorCode STGetKSP(ST st, KSP *ksp)
{
  if (!st->ksp) PetscCall(KSPCreate(&st->ksp));
  *ksp = st->ksp;
  return PETSC_SUCCESS;
}

PetscErrorCode STGetBilinearForm(ST st, Mat *B) { *B = st->nmat > 1 ? st->B : NULL; return PETSC_SUCCESS; }

PetscErrorCode STSetMatrices(ST st, PetscInt n, Mat A[])
{
  PetscCheck(n == 1 || n == 2, PETSC_ERR_ARG_OUTOFRANGE, "Must have one or two matrices (polynomial eigenproblems are outside this path), you have %d", n);
  PetscCheck(A && A[0] && (n == 1 || A[1]), PETSC_ERR_ARG_NULL, "null matrix");
  if (n == 2) PetscCheck(A[1]->M == A[0]->M && A[1]->N == A[0]->N && A[1]->m == A[0]->m, PETSC_ERR_ARG_INCOMP, "A and B have different dimensions or row distributions");
  st->A = A[0]; st->B = (n == 2) ? A[1] : NULL; st->nmat = n;
  st->setup = PETSC_FALSE;
  return PETSC_SUCCESS;
}

PetscErrorCode STSetShift(ST st, PetscScalar shift) { if (st->sigma != shift) { st->sigma = shift; st->setup = PETSC_FALSE; } return PETSC_SUCCESS; }
PetscErrorCode STGetShift(ST st, PetscScalar *shift) { *shift = st->sigma; return PETSC_SUCCESS; }

/* T = A - sigma B (B = I for a standard problem): what STMatMAXPY_Private builds explicitly (stsolve.c:560-640), kept
   matrix-free here — two products and an AXPY */
static PetscErrorCode MatMult_STT(Mat T, Vec x, Vec y)
{
  ST st = (ST)T->data;
  PetscCall(MatMult(st->A, x, y));
  if (st->sigma != 0.0) {
    if (st->B) { PetscCall(MatMult(st->B, x, st->w)); PetscCall(VecAXPY(y, -st->sigma, st->w)); }
    else PetscCall(VecAXPY(y, -st->sigma, x));
  }
  return PETSC_SUCCESS;
}
static PetscErrorCode MatGetDiagonal_STT(Mat T, Vec d)
{
  ST st = (ST)T->data;
  PetscCall(MatGetDiagonal(st->A, d));
  if (st->sigma != 0.0) {
    if (st->B) { PetscCall(MatGetDiagonal(st->B, st->w)); PetscCall(VecAXPY(d, -st->sigma, st->w)); }
    else {
      Vec one;
      PetscCall(MatCreateVecs(st->A, &one, NULL));
      PetscErrorCode ierr = VecSet(one, 1.0);
      if (!ierr) ierr = VecAXPY(d, -st->sigma, one);
      PetscCall(VecDestroy(&one));
      PetscCall(ierr);
    }
  }
  return PETSC_SUCCESS;
}

/* the transformed operator as a Mat (STGetOperator, stsolve.c:324-378): MatMult = STApply */
static PetscErrorCode MatMult_STOperator(Mat Op, Vec x, Vec y) { return STApply((ST)Op->data, x, y); }

static PetscErrorCode STShell_Private(ST st, MatMultFn mult, Mat *out)
{
  Mat A = st->A;
  PetscCall(MatCreateShell(A->m, A->n, A->M, A->N, A->mem, st, out));
  (*out)->rstart = A->rstart; (*out)->rend = A->rend; (*out)->cstart = A->cstart; (*out)->cend = A->cend;
  PetscCall(MatShellSetMult(*out, mult));
  return PETSC_SUCCESS;
}

PetscErrorCode STSetUp(ST st)
{
  PetscCheck(st->A, PETSC_ERR_ARG_WRONGSTATE, "STSetMatrices() must be called first");
  if (st->setup) return PETSC_SUCCESS;
  PetscCall(STReset_Private(st));
  const PetscBool solve = (st->sinvert || st->B) ? PETSC_TRUE : PETSC_FALSE;
  if (st->sigma != 0.0 || solve) {
    PetscCall(MatCreateVecs(st->A, &st->w, NULL));
    PetscCall(STShell_Private(st, MatMult_STT, &st->T));
    PetscCall(MatShellSetGetDiagonal(st->T, MatGetDiagonal_STT));
  }
  if (solve) {
    KSP ksp;
    PetscCall(STGetKSP(st, &ksp));
    /* KSP matrix P: sinvert.c:93-97 (A - sigma B), shift.c:96-101 (B) */
    PetscCall(KSPSetOperators(ksp, st->sinvert ? st->T : st->B, NULL));
    PetscCall(KSPSetUp(ksp));
    PetscCall(STShell_Private(st, MatMult_STOperator, &st->Op));
  } else if (st->sigma != 0.0) {
    PetscCall(STShell_Private(st, MatMult_STT, &st->Op));   /* standard problem, shift: the operator is T itself */
  }
  st->setup = PETSC_TRUE;
  return PETSC_SUCCESS;
}

PetscErrorCode STGetOperator(ST st, Mat *Op)
{
  PetscCall(STSetUp(st));
  *Op = st->Op ? st->Op : st->A;
  return PETSC_SUCCESS;
}
PetscErrorCode STRestoreOperator(ST st, Mat *Op) { (void)st; *Op = NULL; return PETSC_SUCCESS; }

/* STApply_Generic stsolve.c:16-25: y = P^{-1} M x */
PetscErrorCode STApply(ST st, Vec x, Vec y)
{
  PetscCall(STSetUp(st));
  if (!st->sinvert && !st->B) {                   /* standard problem, shift */
    if (st->sigma == 0.0) PetscCall(MatMult(st->A, x, y));
    else PetscCall(MatMult_STT(st->T, x, y));
    return PETSC_SUCCESS;
  }
  Vec w;
  PetscCall(MatCreateVecs(st->A, &w, NULL));      /* st->w is the work vector of T's product */
  PetscErrorCode ierr;
  if (st->sinvert) {                              /* y = (A - sigma B)^{-1} B x */
    if (st->B) ierr = MatMult(st->B, x, w); else ierr = VecCopy(x, w);
  } else ierr = MatMult_STT(st->T, x, w);         /* y = B^{-1} (A - sigma B) x */
  if (!ierr) ierr = KSPSolve(st->ksp, w, y);
  PetscCall(VecDestroy(&w));
  PetscCall(ierr);
  return PETSC_SUCCESS;
}

/* STBackTransform_Shift shift.c:49-58; STBackTransform_Sinvert sinvert.c:51-77 (real arithmetic: conjugate pairs) */
PetscErrorCode STBackTransform(ST st, PetscInt n, PetscScalar *eigr, PetscScalar *eigi)
{
  if (!st->sinvert) { for (PetscInt j = 0; j < n; j++) eigr[j] += st->sigma; return PETSC_SUCCESS; }
  for (PetscInt j = 0; j < n; j++) {
    if (!eigi || eigi[j] == 0.0) eigr[j] = 1.0 / eigr[j] + st->sigma;
    else {
      const PetscScalar r = eigr[j], i = eigi[j];
      PetscScalar t;
      if (fabs(i) <= fabs(r)) { t = i / r; const PetscScalar d = r + i * t; eigr[j] = 1.0 / d + st->sigma; eigi[j] = -t / d; }   /* 1/(r+ij) without overflow */
      else { t = r / i; const PetscScalar d = r * t + i; eigr[j] = t / d + st->sigma; eigi[j] = -1.0 / d; }
    }
  }
  return PETSC_SUCCESS;
}
