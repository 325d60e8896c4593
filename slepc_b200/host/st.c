/*
 * st.c — spectral transformation, shift only (src/sys/classes/st/impls/shift/shift.c).  On the
 * Krylov path ST is a pass-through: for sigma = 0 the operator IS the user's matrix
 * (STMatMAXPY_Private returns A itself, stsolve.c:607-610; STApply_Generic is a bare MatMult,
 * stsolve.c:16-25); for sigma != 0 it is the shell y = A x - sigma x (stsolve.c:244-258, 324-358).
 * sinvert / cayley / filter need linear solves and are out of scope.
 */
#include "b2kimpl.h"

PetscErrorCode STCreate(ST *out)
{
  ST st = (ST)calloc(1, sizeof(*st));
  PetscCheck(st, PETSC_ERR_MEM, "out of memory");
  *out = st;
  return PETSC_SUCCESS;
}

PetscErrorCode STDestroy(ST *pst)
{
  if (!pst || !*pst) return PETSC_SUCCESS;
  ST st = *pst;
  if (st->Op) PetscCall(MatDestroy(&st->Op));
  free(st);
  *pst = NULL;
  return PETSC_SUCCESS;
}

PetscErrorCode STSetMatrices(ST st, PetscInt n, Mat A[])
{
  PetscCheck(n == 1, PETSC_ERR_SUP, "only standard problems (one matrix) are on the Krylov hot path");
  PetscCheck(A && A[0], PETSC_ERR_ARG_NULL, "null matrix");
  st->A = A[0];
  st->setup = PETSC_FALSE;
  return PETSC_SUCCESS;
}

PetscErrorCode STSetShift(ST st, PetscScalar shift) { if (st->sigma != shift) { st->sigma = shift; st->setup = PETSC_FALSE; } return PETSC_SUCCESS; }
PetscErrorCode STGetShift(ST st, PetscScalar *shift) { *shift = st->sigma; return PETSC_SUCCESS; }

/* y = A x - sigma x : MatMult_STOperator for STSHIFT */
static PetscErrorCode MatMult_STShift(Mat Op, Vec x, Vec y)
{
  ST st = (ST)Op->data;
  PetscCall(MatMult(st->A, x, y));
  PetscCall(VecAXPY(y, -st->sigma, x));
  return PETSC_SUCCESS;
}

PetscErrorCode STSetUp(ST st)
{
  PetscCheck(st->A, PETSC_ERR_ARG_WRONGSTATE, "STSetMatrices() must be called first");
  if (st->setup) return PETSC_SUCCESS;
  if (st->Op) PetscCall(MatDestroy(&st->Op));
  if (st->sigma != 0.0) {
    Mat A = st->A;
    PetscCall(MatCreateShell(A->m, A->n, A->M, A->N, A->mem, st, &st->Op));
    st->Op->rstart = A->rstart; st->Op->rend = A->rend; st->Op->cstart = A->cstart; st->Op->cend = A->cend;
    PetscCall(MatShellSetMult(st->Op, MatMult_STShift));
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

PetscErrorCode STApply(ST st, Vec x, Vec y)
{
  Mat Op;
  PetscCall(STGetOperator(st, &Op));
  PetscCall(MatMult(Op, x, y));
  return PETSC_SUCCESS;
}

/* STBackTransform_Shift shift.c:49-58 */
PetscErrorCode STBackTransform(ST st, PetscInt n, PetscScalar *eigr, PetscScalar *eigi)
{
  (void)eigi;
  for (PetscInt j = 0; j < n; j++) eigr[j] += st->sigma;
  return PETSC_SUCCESS;
}
