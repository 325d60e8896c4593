/* bv_test1.c — "Test BV operations" (src/sys/classes/bv/tests/test1.c) against include/b2kslepc.h with the BV type of this build:
   BVMult, BVMultVec, BVDot, BVDotVec, BVMultInPlace, BVScale, BVNorm and the viewers.  With -verbose the output is the
   reference's output/test1_1_bv_type-svec.out line for line, except the type name and the BVMatMult mode line.
       ./bv_test1 -verbose */
#include "exutil.h"

static void print_real(double x) { char b[32]; B2KFormatReal(x, b); printf("%s", b); }

int main(int argc, char **argv)
{
  const PetscInt n = opt_int(argc, argv, "-n", 10), k = opt_int(argc, argv, "-k", 5), l = opt_int(argc, argv, "-l", 3);
  const int verbose = opt_has(argc, argv, "-verbose");
  Vec t, v;
  BV X, Y;
  Mat Q, M;
  PetscScalar *q, *z, *col;
  PetscInt lda;
  PetscReal nrm;
  PetscViewer view;

  CHK(B2KInitialize(opt_int(argc, argv, "-device", 0)));
  printf("Test BV with %d columns of dimension %d.\n", k, n);

  CHK(VecCreateB200(n, n, &t));                    /* template vector */
  CHK(BVCreate(&X));
  CHK(PetscObjectSetName((PetscObject)X, "X"));
  CHK(BVSetType(X, BVB200));
  CHK(BVSetSizesFromVec(X, t, k));

  CHK(PetscViewerASCIIGetStdout(&view));
  CHK(PetscViewerPushFormat(view, PETSC_VIEWER_ASCII_INFO_DETAIL));
  CHK(BVView(X, view));
  CHK(PetscViewerPopFormat(view));

  /* fill X: column j holds 3i+j-2 at rows i+j, i = 0..3 */
  col = (PetscScalar *)malloc(sizeof(PetscScalar) * (size_t)n);
  for (PetscInt j = 0; j < k; j++) {
    for (PetscInt i = 0; i < n; i++) col[i] = 0.0;
    for (PetscInt i = 0; i < 4; i++) if (i + j < n) col[i + j] = (PetscScalar)(3 * i + j - 2);
    CHK(BVSetColumnHost(X, j, col));
  }
  if (verbose) CHK(BVView(X, view));

  CHK(BVCreate(&Y));
  CHK(PetscObjectSetName((PetscObject)Y, "Y"));
  CHK(BVSetType(Y, BVB200));
  CHK(BVSetSizesFromVec(Y, t, l));
  for (PetscInt j = 0; j < l; j++) {
    CHK(BVGetColumn(Y, j, &v));
    CHK(VecSet(v, (PetscScalar)(j + 1) / 4.0));
    CHK(BVRestoreColumn(Y, j, &v));
  }
  if (verbose) CHK(BVView(Y, view));

  CHK(MatCreateSeqDense(k, l, NULL, &Q));
  CHK(PetscObjectSetName((PetscObject)Q, "Q"));
  CHK(MatDenseGetArray(Q, &q));
  CHK(MatDenseGetLDA(Q, &lda));
  for (PetscInt i = 0; i < k; i++)
    for (PetscInt j = 0; j < l; j++) q[i + j * lda] = (i < j) ? 2.0 : -0.5;
  CHK(MatDenseRestoreArray(Q, &q));
  if (verbose) CHK(MatView(Q, NULL));

  CHK(BVMult(Y, 2.0, 1.0, X, Q));
  if (verbose) { printf("After BVMult - - - - - - - - -\n"); CHK(BVView(Y, view)); }

  CHK(BVGetColumn(Y, 0, &v));
  z = (PetscScalar *)malloc(sizeof(PetscScalar) * (size_t)k);
  z[0] = 2.0;
  for (PetscInt i = 1; i < k; i++) z[i] = -0.5 * z[i - 1];
  CHK(BVMultVec(X, -1.0, 1.0, v, z));
  CHK(BVRestoreColumn(Y, 0, &v));
  if (verbose) { printf("After BVMultVec - - - - - - -\n"); CHK(BVView(Y, view)); }

  CHK(MatCreateSeqDense(l, k, NULL, &M));
  CHK(PetscObjectSetName((PetscObject)M, "M"));
  CHK(BVDot(X, Y, M));
  if (verbose) { printf("After BVDot - - - - - - - - -\n"); CHK(MatView(M, NULL)); }

  CHK(BVGetColumn(Y, 0, &v));
  CHK(BVDotVec(X, v, z));
  CHK(BVRestoreColumn(Y, 0, &v));
  if (verbose) {
    Vec zv;
    printf("After BVDotVec - - - - - - -\n");
    CHK(VecCreateWithArray(B2K_MEM_HOST, k, k, z, &zv));
    CHK(PetscObjectSetName((PetscObject)zv, "z"));
    CHK(VecView(zv, view));
    CHK(VecDestroy(&zv));
  }
  free(z);

  CHK(BVMultInPlace(X, Q, 1, l));
  CHK(BVScale(X, 2.0));
  if (verbose) { printf("After BVMultInPlace - - - - -\n"); CHK(BVView(X, view)); }

  CHK(BVNormColumn(X, 0, NORM_2, &nrm));
  printf("2-Norm of X[0] = "); print_real(nrm); printf("\n");
  CHK(BVNorm(X, NORM_FROBENIUS, &nrm));
  printf("Frobenius Norm of X = "); print_real(nrm); printf("\n");

  printf("First row of X =\n");
  for (PetscInt i = 0; i < k; i++) { CHK(BVGetColumnHost(X, i, col)); print_real(col[0]); printf(" "); }
  printf("\n");
  free(col);

  CHK(BVDestroy(&X));
  CHK(BVDestroy(&Y));
  CHK(MatDestroy(&Q));
  CHK(MatDestroy(&M));
  CHK(VecDestroy(&t));
  CHK(B2KFinalize());
  return 0;
}
