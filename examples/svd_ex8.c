/* svd_ex8.c — "Estimates the 2-norm condition number of a matrix A, that is, the ratio of the largest to the smallest singular
   values.  The matrix is a Grcar matrix." (src/svd/tutorials/ex8.c) against include/b2kslepc.h: two thick-restart Lanczos solves
   on the same SVD object, SVD_LARGEST then SVD_SMALLEST.
       ./svd_ex8        reproduces src/svd/tutorials/output/ex8_1.out */
#include "exutil.h"

int main(int argc, char **argv)
{
  const PetscInt N = opt_int(argc, argv, "-n", 30);
  Mat A;
  SVD svd;
  PetscInt nconv1, nconv2;
  PetscReal sigma_1 = 0.0, sigma_n = 0.0;

  CHK(B2KInitialize(opt_int(argc, argv, "-device", 0)));
  printf("\nEstimate the condition number of a Grcar matrix, n=%d\n\n", N);
  CsrBuilder b;
  csr_init(&b, N);
  for (PetscInt i = 0; i < N; i++) {               /* -1 on the subdiagonal, ones on the diagonal and the three superdiagonals */
    if (i > 0) csr_add(&b, i - 1, -1.0);
    for (PetscInt j = i; j < i + 4 && j < N; j++) csr_add(&b, j, 1.0);
    csr_end_row(&b);
  }
  CHK(MatCreateB200CSR(N, N, 0, N, b.rowptr, b.col, b.val, 0, N, &A));
  csr_free(&b);

  CHK(SVDCreate(&svd));
  CHK(SVDSetOperators(svd, A, NULL));
  CHK(SVDSetDimensions(svd, 1, PETSC_DETERMINE, PETSC_DETERMINE));

  CHK(SVDSetWhichSingularTriplets(svd, SVD_LARGEST));
  CHK(SVDSolve(svd));
  CHK(SVDGetConverged(svd, &nconv1));
  if (nconv1 > 0) CHK(SVDGetSingularTriplet(svd, 0, &sigma_1, NULL, NULL));
  else printf(" Unable to compute large singular value!\n\n");

  CHK(SVDSetWhichSingularTriplets(svd, SVD_SMALLEST));
  CHK(SVDSolve(svd));
  CHK(SVDGetConverged(svd, &nconv2));
  if (nconv2 > 0) CHK(SVDGetSingularTriplet(svd, 0, &sigma_n, NULL, NULL));
  else printf(" Unable to compute small singular value!\n\n");

  if (nconv1 > 0 && nconv2 > 0) {
    printf(" Computed singular values: sigma_1=%.4f, sigma_n=%.4f\n", sigma_1, sigma_n);
    printf(" Estimated condition number: sigma_1/sigma_n=%.4f\n\n", sigma_1 / sigma_n);
  }
  CHK(SVDDestroy(&svd));
  CHK(MatDestroy(&A));
  CHK(B2KFinalize());
  return 0;
}
