/* ex2.c — "Standard symmetric eigenproblem corresponding to the Laplacian operator in 2 dimensions", the program of
   src/eps/tutorials/ex2.c written against include/b2kslepc.h: the operator is a Mat of type b200csr resident in HBM and the
   basis a BV of type b200, so EPSSolve runs on the B200 path (no CPU fallback: without a GPU B2KInitialize fails).
       ./ex2 -n 72 -eps_nev 4 -eps_ncv 20 -terse        reproduces src/eps/tutorials/output/ex2_1.out */
#include "exutil.h"

int main(int argc, char **argv)
{
  const PetscInt n = opt_int(argc, argv, "-n", 10), m = opt_int(argc, argv, "-m", n), N = n * m;
  const PetscInt nev_req = opt_int(argc, argv, "-eps_nev", 1), ncv = opt_int(argc, argv, "-eps_ncv", PETSC_DETERMINE);
  Mat A;
  EPS eps;
  PetscInt nev;

  CHK(B2KInitialize(opt_int(argc, argv, "-device", 0)));
  printf("\n2-D Laplacian Eigenproblem, N=%d (%dx%d grid)\n\n", N, n, m);

  /* the matrix of ex2.c:39-54 (grid point (i,j) is row i*m+j; 4 on the diagonal, -1 for the four neighbours), row by row */
  CsrBuilder b;
  csr_init(&b, N);
  for (PetscInt II = 0; II < N; II++) {
    const PetscInt i = II / m, j = II - i * m;
    if (i > 0) csr_add(&b, II - m, -1.0);
    if (i < n - 1) csr_add(&b, II + m, -1.0);
    if (j > 0) csr_add(&b, II - 1, -1.0);
    if (j < m - 1) csr_add(&b, II + 1, -1.0);
    csr_add(&b, II, 4.0);
    csr_end_row(&b);
  }
  CHK(MatCreateB200CSR(N, N, 0, N, b.rowptr, b.col, b.val, 0, N, &A));
  csr_free(&b);

  CHK(EPSCreate(&eps));
  CHK(EPSSetOperators(eps, A, NULL));
  CHK(EPSSetProblemType(eps, EPS_HEP));
  CHK(EPSSetDimensions(eps, nev_req, ncv, PETSC_DETERMINE));
  CHK(EPSSolve(eps));

  printf(" Solution method: %s\n\n", EPSKRYLOVSCHUR);
  CHK(EPSGetDimensions(eps, &nev, NULL, NULL));
  printf(" Number of requested eigenvalues: %d\n", nev);
  if (opt_has(argc, argv, "-terse")) eps_error_view_terse(eps);
  else {
    PetscInt nconv, its;
    CHK(EPSGetConverged(eps, &nconv));
    CHK(EPSGetIterationNumber(eps, &its));
    printf(" Number of iterations of the method: %d\n Number of converged eigenpairs: %d\n\n", its, nconv);
    printf("           k          ||Ax-kx||/||kx||\n   ----------------- ------------------\n");
    for (PetscInt i = 0; i < nconv; i++) {
      PetscScalar kr, ki;
      PetscReal err;
      CHK(EPSGetEigenvalue(eps, i, &kr, &ki));
      CHK(EPSComputeError(eps, i, EPS_ERROR_RELATIVE, &err));
      printf("   %12f       %12g\n", kr, err);
    }
    printf("\n");
  }
  CHK(EPSDestroy(&eps));
  CHK(MatDestroy(&A));
  CHK(B2KFinalize());
  return 0;
}
