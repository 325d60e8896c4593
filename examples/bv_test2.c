/* bv_test2.c — "Test BV orthogonalization functions" (src/sys/classes/bv/tests/test2.c) against include/b2kslepc.h:
   BVOrthogonalizeColumn, BVOrthogonalize (QR with R), BVOrthogonalizeVec; [-bv_orthog_type cgs|mgs] as in the reference's test
   loop.     ./bv_test2        reproduces src/sys/classes/bv/tests/output/test2_1.out */
#include <math.h>
#include "exutil.h"

static double shifted_norm1(Mat M, PetscInt k)     /* || M - I ||_1 (MatShift(M,-1) + MatNorm(M,NORM_1)) */
{
  PetscScalar *a;
  PetscInt lda;
  double mx = 0.0;
  CHK(MatDenseGetArray(M, &a));
  CHK(MatDenseGetLDA(M, &lda));
  for (PetscInt j = 0; j < k; j++) {
    double s = 0.0;
    for (PetscInt i = 0; i < k; i++) s += fabs(a[i + j * lda] - (i == j ? 1.0 : 0.0));
    if (s > mx) mx = s;
  }
  CHK(MatDenseRestoreArray(M, &a));
  return mx;
}

int main(int argc, char **argv)
{
  const PetscInt n = opt_int(argc, argv, "-n", 20), k = opt_int(argc, argv, "-k", 8);
  Vec t, e;
  BV X, Y, Z;
  Mat M, R;
  PetscReal norm;
  char buf[32];
  const char *ot = "cgs";
  for (int i = 1; i + 1 < argc; i++) if (!strcmp(argv[i], "-bv_orthog_type")) ot = argv[i + 1];

  CHK(B2KInitialize(opt_int(argc, argv, "-device", 0)));
  printf("Test BV orthogonalization with %d columns of length %d.\n", k, n);
  CHK(VecCreateB200(n, n, &t));
  CHK(BVCreate(&X));
  CHK(PetscObjectSetName((PetscObject)X, "X"));
  CHK(BVSetType(X, BVB200));
  CHK(BVSetSizesFromVec(X, t, k));
  CHK(BVSetOrthogonalization(X, !strcmp(ot, "mgs") ? BV_ORTHOG_MGS : BV_ORTHOG_CGS, BV_ORTHOG_REFINE_IFNEEDED, PETSC_DEFAULT, BV_ORTHOG_BLOCK_GS));

  PetscScalar *col = (PetscScalar *)malloc(sizeof(PetscScalar) * (size_t)n);
  for (PetscInt j = 0; j < k; j++) {
    for (PetscInt i = 0; i < n; i++) col[i] = 0.0;
    for (PetscInt i = 0; i <= n / 2; i++) if (i + j < n) col[i + j] = (3.0 * i + j - 2) / (2 * (i + j + 1));
    CHK(BVSetColumnHost(X, j, col));
  }
  free(col);

  CHK(BVDuplicate(X, &Y));
  CHK(PetscObjectSetName((PetscObject)Y, "Y"));
  CHK(BVCopy(X, Y));
  CHK(BVDuplicate(X, &Z));
  CHK(PetscObjectSetName((PetscObject)Z, "Z"));
  CHK(BVCopy(X, Z));

  for (PetscInt j = 0; j < k; j++) {               /* BVOrthogonalizeColumn */
    CHK(BVOrthogonalizeColumn(X, j, NULL, &norm, NULL));
    CHK(BVScaleColumn(X, j, 1.0 / norm));
  }
  CHK(MatCreateSeqDense(k, k, NULL, &M));
  CHK(BVDot(X, X, M));
  norm = shifted_norm1(M, k);
  if (norm < 100 * PETSC_MACHINE_EPSILON) printf("Level of orthogonality < 100*eps\n");
  else printf("Level of orthogonality: %g\n", norm);

  CHK(MatCreateSeqDense(k, k, NULL, &R));          /* BVOrthogonalize */
  CHK(PetscObjectSetName((PetscObject)R, "R"));
  CHK(BVOrthogonalize(Y, R));
  CHK(BVDot(Y, Y, M));
  norm = shifted_norm1(M, k);
  if (norm < 100 * PETSC_MACHINE_EPSILON) printf("Level of orthogonality < 100*eps\n");
  else printf("Level of orthogonality: %g\n", norm);

  CHK(BVMult(Z, -1.0, 1.0, Y, R));                 /* residual */
  CHK(BVNorm(Z, NORM_FROBENIUS, &norm));
  if (norm < 100 * PETSC_MACHINE_EPSILON) printf("Residual ||X-QR|| < 100*eps\n");
  else printf("Residual ||X-QR||: %g\n", norm);

  CHK(VecCreateB200(n, n, &e));                    /* BVOrthogonalizeVec */
  CHK(VecSet(e, 1.0));
  CHK(BVOrthogonalizeVec(X, e, NULL, &norm, NULL));
  B2KFormatReal(norm, buf);
  printf("Norm of ones(n,1) after orthogonalizing against X: %s\n", buf);

  CHK(MatDestroy(&M));
  CHK(MatDestroy(&R));
  CHK(BVDestroy(&X));
  CHK(BVDestroy(&Y));
  CHK(BVDestroy(&Z));
  CHK(VecDestroy(&e));
  CHK(VecDestroy(&t));
  CHK(B2KFinalize());
  return 0;
}
