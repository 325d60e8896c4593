/* exutil.h — the little a tutorial program needs around include/b2kslepc.h: option lookup in argv (the -n / -eps_nev /
   -terse style of the reference's programs), a row-wise CSR assembler standing in for MatSetValues + MatAssembly, and the
   `-terse` result line of EPSErrorView / SVDErrorView (epsview.c:296-339, svdview.c) so that the programs print what the
   reference's output files hold. */
#ifndef EXUTIL_H
#define EXUTIL_H
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "b2kslepc.h"

#define CHK(call)                                                                                   \
  do {                                                                                              \
    PetscErrorCode ierr_ = (call);                                                                  \
    if (ierr_) { fprintf(stderr, "error %d: %s\n", ierr_, B2KGetLastErrorMessage()); exit(ierr_); } \
  } while (0)

static int opt_has(int argc, char **argv, const char *name)
{
  for (int i = 1; i < argc; i++) if (!strcmp(argv[i], name)) return 1;
  return 0;
}
static int opt_int(int argc, char **argv, const char *name, int dflt)
{
  for (int i = 1; i + 1 < argc; i++) if (!strcmp(argv[i], name)) return atoi(argv[i + 1]);
  return dflt;
}

/* rows are appended in order; the entries of a row may come in any order (kept sorted by column) */
typedef struct { PetscInt nrows, nnz, cap, *rowptr, *col; PetscScalar *val; } CsrBuilder;
static void csr_init(CsrBuilder *b, PetscInt nrows_max)
{
  b->nrows = 0; b->nnz = 0; b->cap = 8 * (nrows_max + 1);
  b->rowptr = (PetscInt *)calloc((size_t)nrows_max + 1, sizeof(PetscInt));
  b->col = (PetscInt *)malloc(sizeof(PetscInt) * (size_t)b->cap);
  b->val = (PetscScalar *)malloc(sizeof(PetscScalar) * (size_t)b->cap);
}
static void csr_add(CsrBuilder *b, PetscInt col, PetscScalar v)
{
  if (b->nnz >= b->cap) {
    b->cap *= 2;
    b->col = (PetscInt *)realloc(b->col, sizeof(PetscInt) * (size_t)b->cap);
    b->val = (PetscScalar *)realloc(b->val, sizeof(PetscScalar) * (size_t)b->cap);
  }
  PetscInt q = b->nnz++;
  while (q > b->rowptr[b->nrows] && b->col[q - 1] > col) { b->col[q] = b->col[q - 1]; b->val[q] = b->val[q - 1]; q--; }
  b->col[q] = col; b->val[q] = v;
}
static void csr_end_row(CsrBuilder *b) { b->rowptr[++b->nrows] = b->nnz; }
static void csr_free(CsrBuilder *b) { free(b->rowptr); free(b->col); free(b->val); }

static void print_value(PetscScalar re, PetscScalar im)
{
  if (im != 0.0) printf("%.5f%+.5fi", re, im);
  else printf("%.5f", re);
}

/* EPSErrorView(eps,EPS_ERROR_RELATIVE,NULL) with -terse: epsview.c:296-339 */
static void eps_error_view_terse(EPS eps)
{
  PetscInt nev, nconv;
  PetscReal tol, err;
  CHK(EPSGetDimensions(eps, &nev, NULL, NULL));
  CHK(EPSGetConverged(eps, &nconv));
  CHK(EPSGetTolerances(eps, &tol, NULL));
  if (nconv < nev) { printf(" Problem: less than %d eigenvalues converged\n\n", nev); return; }
  for (PetscInt i = 0; i < nev; i++) {
    CHK(EPSComputeError(eps, i, EPS_ERROR_RELATIVE, &err));
    if (err >= 5.0 * tol) { printf(" Problem: some of the first %d relative errors are higher than the tolerance\n\n", nev); return; }
  }
  printf(" All requested eigenvalues computed up to the required tolerance:");
  for (PetscInt i = 0; i <= (nev - 1) / 8; i++) {
    printf("\n     ");
    for (PetscInt j = 0; j < (nev - 8 * i < 8 ? nev - 8 * i : 8); j++) {
      PetscScalar kr, ki;
      CHK(EPSGetEigenvalue(eps, 8 * i + j, &kr, &ki));
      print_value(kr, ki);
      if (8 * i + j + 1 < nev) printf(", ");
    }
  }
  printf("\n\n");
}

/* SVDErrorView(svd,SVD_ERROR_RELATIVE,NULL), the default ASCII format: svdview.c:208-243 */
static void svd_error_view(SVD svd)
{
  PetscInt nsv, nconv;
  PetscReal err, sigma;
  const PetscReal tol = 1e-8;
  CHK(SVDGetDimensions(svd, &nsv, NULL, NULL));
  CHK(SVDGetConverged(svd, &nconv));
  if (nconv < nsv) { printf(" Problem: less than %d singular values converged\n\n", nsv); return; }
  for (PetscInt i = 0; i < nsv; i++) {
    CHK(SVDComputeError(svd, i, SVD_ERROR_RELATIVE, &err));
    if (err >= 5.0 * tol) { printf(" Problem: some of the first %d relative errors are higher than the tolerance\n\n", nsv); return; }
  }
  printf(" All requested singular values computed up to the required tolerance:");
  for (PetscInt i = 0; i <= (nsv - 1) / 8; i++) {
    printf("\n     ");
    for (PetscInt j = 0; j < (nsv - 8 * i < 8 ? nsv - 8 * i : 8); j++) {
      CHK(SVDGetSingularTriplet(svd, 8 * i + j, &sigma, NULL, NULL));
      printf("%.5f", sigma);
      if (8 * i + j + 1 < nsv) printf(", ");
    }
  }
  printf("\n\n");
}
#endif
