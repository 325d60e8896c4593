/* eps_test4.c — "Test the solution of a HEP without calling EPSSetFromOptions (based on ex1.c)", src/eps/tests/test4.c, against
   include/b2kslepc.h with -type krylovschur: 1-D Laplacian n = 30, 4 largest eigenvalues, tol = 1000 eps.
       ./eps_test4        reproduces src/eps/tests/output/test4_1.out */
#include "exutil.h"

int main(int argc, char **argv)
{
  const PetscInt n = opt_int(argc, argv, "-n", 30);
  const PetscReal tol = 1000 * PETSC_MACHINE_EPSILON;
  Mat A;
  EPS eps;
  PetscInt nev;

  CHK(B2KInitialize(opt_int(argc, argv, "-device", 0)));
  printf("\n1-D Laplacian Eigenproblem, n=%d\n\n", n);
  CsrBuilder b;
  csr_init(&b, n);
  for (PetscInt i = 0; i < n; i++) {
    if (i > 0) csr_add(&b, i - 1, -1.0);
    if (i < n - 1) csr_add(&b, i + 1, -1.0);
    csr_add(&b, i, 2.0);
    csr_end_row(&b);
  }
  CHK(MatCreateB200CSR(n, n, 0, n, b.rowptr, b.col, b.val, 0, n, &A));
  csr_free(&b);

  CHK(EPSCreate(&eps));
  CHK(EPSSetOperators(eps, A, NULL));
  CHK(EPSSetProblemType(eps, EPS_HEP));
  CHK(EPSSetDimensions(eps, 4, PETSC_DETERMINE, PETSC_DETERMINE));
  CHK(EPSSetTolerances(eps, tol, PETSC_CURRENT));
  CHK(EPSSetType(eps, EPSKRYLOVSCHUR));
  CHK(EPSSolve(eps));
  CHK(EPSGetDimensions(eps, &nev, NULL, NULL));
  printf(" Number of requested eigenvalues: %d\n", nev);
  eps_error_view_terse(eps);
  CHK(EPSDestroy(&eps));
  CHK(MatDestroy(&A));
  CHK(B2KFinalize());
  return 0;
}
