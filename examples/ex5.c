/* ex5.c — "Eigenvalue problem associated with a Markov model of a random walk on a triangular grid", the program of
   src/eps/tutorials/ex5.c against include/b2kslepc.h (non-symmetric Krylov-Schur = Arnoldi on the B200 path).
       ./ex5 -m 15 -eps_nev 4 -eps_largest_real -terse      reproduces src/eps/tutorials/output/ex5_1.out (hence 1.00000 first) */
#include "exutil.h"

/* The transposed transition matrix of the walk (ex5.c:138-170).  States are the grid points (i,j), i,j >= 1, i+j <= m+1,
   numbered block by block in i; a step to a neighbour with a larger index has probability pd = (i+j-1)/(2(m-1)) (doubled on the
   boundary it cannot leave the other way), a step back pu = 1/2 - (i+j-3)/(2(m-1)). */
static void markov_rows(PetscInt m, CsrBuilder *b)
{
  const double cst = 0.5 / (double)(m - 1);
  PetscInt row = 0;
  for (PetscInt i = 1; i <= m; i++) {
    const PetscInt jmax = m - i + 1;
    for (PetscInt j = 1; j <= jmax; j++, row++) {
      const double pd = cst * (double)(i + j - 1), pu = 0.5 - cst * (double)(i + j - 3);
      if (j != jmax) {
        csr_add(b, row + 1, i == 1 ? 2.0 * pd : pd);           /* north */
        csr_add(b, row + jmax, j == 1 ? 2.0 * pd : pd);        /* east  */
      }
      if (j > 1) csr_add(b, row - 1, pu);                      /* south */
      if (i > 1) csr_add(b, row - jmax - 1, pu);               /* west  */
      csr_end_row(b);
    }
  }
}

int main(int argc, char **argv)
{
  const PetscInt m = opt_int(argc, argv, "-m", 15), N = m * (m + 1) / 2;
  const PetscInt nev_req = opt_int(argc, argv, "-eps_nev", 1), ncv = opt_int(argc, argv, "-eps_ncv", PETSC_DETERMINE);
  Mat A;
  EPS eps;
  Vec v0;
  PetscInt nev;

  CHK(B2KInitialize(opt_int(argc, argv, "-device", 0)));
  printf("\nMarkov Model, N=%d (m=%d)\n\n", N, m);
  CsrBuilder b;
  csr_init(&b, N);
  markov_rows(m, &b);
  CHK(MatCreateB200CSR(N, N, 0, N, b.rowptr, b.col, b.val, 0, N, &A));
  csr_free(&b);

  CHK(EPSCreate(&eps));
  CHK(EPSSetOperators(eps, A, NULL));
  CHK(EPSSetProblemType(eps, EPS_NHEP));
  CHK(EPSSetDimensions(eps, nev_req, ncv, PETSC_DETERMINE));
  if (opt_has(argc, argv, "-eps_largest_real")) CHK(EPSSetWhichEigenpairs(eps, EPS_LARGEST_REAL));

  /* initial vector e0+e1+e2 (ex5.c:74-83) */
  CHK(MatCreateVecs(A, &v0, NULL));
  {
    double *h = (double *)calloc((size_t)N, sizeof(double));
    h[0] = h[1] = h[2] = 1.0;
    CHK(VecSetValuesHost(v0, h));
    free(h);
  }
  CHK(EPSSetInitialSpace(eps, 1, &v0));
  CHK(EPSSolve(eps));

  printf(" Solution method: %s\n\n", EPSKRYLOVSCHUR);
  CHK(EPSGetDimensions(eps, &nev, NULL, NULL));
  printf(" Number of requested eigenvalues: %d\n", nev);
  eps_error_view_terse(eps);
  CHK(EPSDestroy(&eps));
  CHK(MatDestroy(&A));
  CHK(VecDestroy(&v0));
  CHK(B2KFinalize());
  return 0;
}
