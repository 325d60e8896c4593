/* ex3.c — "Solves the same eigenproblem as in example ex2, but using a shell matrix" (src/eps/tutorials/ex3.c) against
   include/b2kslepc.h: EPS only ever sees a MatShell whose MATOP_MULT is a user callback (ex3.c:46-49,140-168) — the plug-in
   route for operators that are not a registered Mat type.  Here the callback forwards to a b200csr matrix held in the shell's
   context (a user with a matrix-free stencil kernel would launch it on B2KGetContext()'s stream instead); the vectors it is
   handed live in HBM.
       ./ex3 -n 72 -eps_nev 4 -eps_ncv 20 -terse        reproduces src/eps/tutorials/output/ex3_1.out */
#include "exutil.h"

typedef struct { Mat L; PetscInt calls; } ShellCtx;

static PetscErrorCode MatMult_Laplacian2D(Mat A, Vec x, Vec y)
{
  ShellCtx *c;
  PetscErrorCode ierr = MatShellGetContext(A, (void **)&c);
  if (ierr) return ierr;
  c->calls++;
  return MatMult(c->L, x, y);
}

int main(int argc, char **argv)
{
  const PetscInt n = opt_int(argc, argv, "-n", 10), N = n * n;
  Mat A;
  EPS eps;
  Vec t;
  B2KMemType mem;
  ShellCtx ctx = {NULL, 0};
  PetscInt nev;

  CHK(B2KInitialize(opt_int(argc, argv, "-device", 0)));
  printf("\n2-D Laplacian Eigenproblem (matrix-free version), N=%d (%dx%d grid)\n\n", N, n, n);

  CsrBuilder b;
  csr_init(&b, N);
  for (PetscInt II = 0; II < N; II++) {
    const PetscInt i = II / n, j = II - i * n;
    if (i > 0) csr_add(&b, II - n, -1.0);
    if (i < n - 1) csr_add(&b, II + n, -1.0);
    if (j > 0) csr_add(&b, II - 1, -1.0);
    if (j < n - 1) csr_add(&b, II + 1, -1.0);
    csr_add(&b, II, 4.0);
    csr_end_row(&b);
  }
  CHK(MatCreateB200CSR(N, N, 0, N, b.rowptr, b.col, b.val, 0, N, &ctx.L));
  csr_free(&b);

  /* the shell hands out vectors of the memory kind of the operator behind it (MatShellSetVecType in the reference) */
  CHK(MatCreateVecs(ctx.L, &t, NULL));
  CHK(VecGetMemType(t, &mem));
  CHK(VecDestroy(&t));
  CHK(MatCreateShell(N, N, N, N, mem, &ctx, &A));
  CHK(MatShellSetMult(A, MatMult_Laplacian2D));

  CHK(EPSCreate(&eps));
  CHK(EPSSetOperators(eps, A, NULL));
  CHK(EPSSetProblemType(eps, EPS_HEP));
  CHK(EPSSetDimensions(eps, opt_int(argc, argv, "-eps_nev", 1), opt_int(argc, argv, "-eps_ncv", PETSC_DETERMINE), PETSC_DETERMINE));
  CHK(EPSSolve(eps));

  printf(" Solution method: %s\n\n", EPSKRYLOVSCHUR);
  CHK(EPSGetDimensions(eps, &nev, NULL, NULL));
  printf(" Number of requested eigenvalues: %d\n", nev);
  eps_error_view_terse(eps);
  if (opt_has(argc, argv, "-calls")) printf(" shell MatMult calls: %d\n", ctx.calls);
  CHK(EPSDestroy(&eps));
  CHK(MatDestroy(&A));
  CHK(MatDestroy(&ctx.L));
  CHK(B2KFinalize());
  return 0;
}
