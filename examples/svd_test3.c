/* svd_test3.c — "SVD of a rectangular Grcar matrix" (src/svd/tests/test3.c) against include/b2kslepc.h: thick-restart Lanczos
   bidiagonalisation on the B200 path, two-sided by default.
       ./svd_test3 -svd_nsv 4 [-svd_trlanczos_oneside] [-bv_orthog_type mgs] [-bv_orthog_refine always] [-svd_trlanczos_locking 0]
   all reproduce src/svd/tests/output/test3_1.out */
#include "exutil.h"

static const char *opt_str(int argc, char **argv, const char *name, const char *dflt)
{
  for (int i = 1; i + 1 < argc; i++) if (!strcmp(argv[i], name)) return argv[i + 1];
  return dflt;
}

int main(int argc, char **argv)
{
  const PetscInt N = opt_int(argc, argv, "-n", 35), M = opt_int(argc, argv, "-m", 30);   /* N rows, M columns as in test3.c:37 */
  Mat A;
  SVD svd;
  Vec v0, w0;
  BV V, U;

  CHK(B2KInitialize(opt_int(argc, argv, "-device", 0)));
  printf("\nSVD of a rectangular Grcar matrix, %dx%d\n\n", N, M);
  CsrBuilder b;
  csr_init(&b, N);
  for (PetscInt i = 0; i < N; i++) {               /* -1 below the diagonal, ones on the diagonal and the next three columns */
    if (i > 0 && i - 1 < M) csr_add(&b, i - 1, -1.0);
    for (PetscInt j = i; j < i + 4 && j < M; j++) csr_add(&b, j, 1.0);
    csr_end_row(&b);
  }
  CHK(MatCreateB200CSR(N, M, 0, N, b.rowptr, b.col, b.val, 0, M, &A));
  csr_free(&b);

  CHK(SVDCreate(&svd));
  CHK(SVDSetOperators(svd, A, NULL));
  CHK(SVDSetDimensions(svd, opt_int(argc, argv, "-svd_nsv", 1), opt_int(argc, argv, "-svd_ncv", PETSC_DETERMINE), PETSC_DETERMINE));
  CHK(SVDTRLanczosSetOneSide(svd, opt_has(argc, argv, "-svd_trlanczos_oneside")));
  CHK(SVDTRLanczosSetLocking(svd, opt_int(argc, argv, "-svd_trlanczos_locking", 1)));
  CHK(SVDGetBV(svd, &V, &U));
  {
    const char *ot = opt_str(argc, argv, "-bv_orthog_type", "cgs"), *orf = opt_str(argc, argv, "-bv_orthog_refine", "ifneeded");
    const BVOrthogType type = !strcmp(ot, "mgs") ? BV_ORTHOG_MGS : BV_ORTHOG_CGS;
    const BVOrthogRefineType ref = !strcmp(orf, "always") ? BV_ORTHOG_REFINE_ALWAYS : (!strcmp(orf, "never") ? BV_ORTHOG_REFINE_NEVER : BV_ORTHOG_REFINE_IFNEEDED);
    CHK(BVSetOrthogonalization(V, type, ref, PETSC_DEFAULT, BV_ORTHOG_BLOCK_GS));
    CHK(BVSetOrthogonalization(U, type, ref, PETSC_DEFAULT, BV_ORTHOG_BLOCK_GS));
  }
  /* initial vectors of all ones (test3.c:75-79) */
  CHK(MatCreateVecs(A, &v0, &w0));
  CHK(VecSet(v0, 1.0));
  CHK(VecSet(w0, 1.0));
  CHK(SVDSetInitialSpaces(svd, 1, &v0, 1, &w0));
  CHK(SVDSolve(svd));
  svd_error_view(svd);

  /* level of orthogonality of the computed singular vectors (test3.c:88-102) */
  {
    PetscInt nconv;
    CHK(SVDGetConverged(svd, &nconv));
    if (nconv > 1) {
      Vec *Us = (Vec *)malloc(sizeof(Vec) * (size_t)nconv), *Vs = (Vec *)malloc(sizeof(Vec) * (size_t)nconv);
      PetscReal lev1 = 0.0, lev2 = 0.0;
      for (PetscInt i = 0; i < nconv; i++) {
        CHK(MatCreateVecs(A, &Vs[i], &Us[i]));
        CHK(SVDGetSingularTriplet(svd, i, NULL, Us[i], Vs[i]));
      }
      for (PetscInt i = 0; i < nconv; i++)
        for (PetscInt j = 0; j <= i; j++) {
          PetscScalar d1, d2;
          CHK(VecDot(Us[i], Us[j], &d1));
          CHK(VecDot(Vs[i], Vs[j], &d2));
          if (i == j) { d1 -= 1.0; d2 -= 1.0; }
          if (d1 < 0) d1 = -d1;
          if (d2 < 0) d2 = -d2;
          if (d1 > lev1) lev1 = d1;
          if (d2 > lev2) lev2 = d2;
        }
      if (lev1 + lev2 < 20 * 1e-8) printf("Level of orthogonality below the tolerance\n");
      else printf("Level of orthogonality: %g (U) %g (V)\n", lev1, lev2);
      for (PetscInt i = 0; i < nconv; i++) { CHK(VecDestroy(&Us[i])); CHK(VecDestroy(&Vs[i])); }
      free(Us); free(Vs);
    }
  }
  CHK(VecDestroy(&v0));
  CHK(VecDestroy(&w0));
  CHK(SVDDestroy(&svd));
  CHK(MatDestroy(&A));
  CHK(B2KFinalize());
  return 0;
}
