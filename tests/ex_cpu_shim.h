/* TEST INFRASTRUCTURE (force-included by tests/test_z_examples.py with `gcc -include`): compiles the programs of examples/
   against the CPU oracle plug-in (oracle/oracle_cpu.c) instead of the GPU types, so that their host logic — matrix assembly,
   solver set-up, output format — is checked against the reference's .out files on a box without a GPU.  The sources under
   examples/ never see this header in the product build. */
#ifndef EX_CPU_SHIM_H
#define EX_CPU_SHIM_H
#include "b2kslepc.h"
PetscErrorCode OracleCPURegister(void);
PetscErrorCode MatCreateOracleCSR(PetscInt M, PetscInt N, PetscInt rstart, PetscInt rend, PetscInt cstart, PetscInt cend, const PetscInt *rowptr,
                                  const PetscInt *colidx_local, const PetscScalar *val, PetscInt nghost, PetscInt nrecv, const PetscInt *recvrank,
                                  const PetscInt *recvcount, PetscInt nsend, const PetscInt *sendrank, const PetscInt *sendcount,
                                  const PetscInt *sendidx, Mat *out);
static PetscErrorCode ShimInit(int dev) { (void)dev; return OracleCPURegister(); }
#include <stdlib.h>
static Mat shim_last_A, shim_last_AT;             /* the oracle CSR type has no MatMultTranspose: keep an explicit A^T for the SVD */
static PetscErrorCode ShimMatCSR(PetscInt M, PetscInt N, PetscInt rs, PetscInt re, const PetscInt *rp, const PetscInt *ci, const PetscScalar *v,
                                 PetscInt cs, PetscInt ce, Mat *A)
{
  PetscErrorCode ierr = MatCreateOracleCSR(M, N, rs, re, cs, ce, rp, ci, v, 0, 0, NULL, NULL, 0, NULL, NULL, NULL, A);
  if (ierr) return ierr;
  const PetscInt nnz = rp[M];
  PetscInt *trp = (PetscInt *)calloc((size_t)N + 2, sizeof(PetscInt)), *tci = (PetscInt *)malloc(sizeof(PetscInt) * (size_t)(nnz + 1));
  PetscScalar *tv = (PetscScalar *)malloc(sizeof(PetscScalar) * (size_t)(nnz + 1));
  for (PetscInt k = 0; k < nnz; k++) trp[ci[k] + 2]++;
  for (PetscInt c = 0; c < N; c++) trp[c + 2] += trp[c + 1];
  for (PetscInt r = 0; r < M; r++) for (PetscInt k = rp[r]; k < rp[r + 1]; k++) { const PetscInt q = trp[ci[k] + 1]++; tci[q] = r; tv[q] = v[k]; }
  ierr = MatCreateOracleCSR(N, M, 0, N, 0, M, trp, tci, tv, 0, 0, NULL, NULL, 0, NULL, NULL, NULL, &shim_last_AT);
  free(trp); free(tci); free(tv);
  shim_last_A = *A;
  return ierr;
}
static PetscErrorCode ShimSVDSetOperators(SVD svd, Mat A, Mat B)
{
  PetscErrorCode ierr = SVDSetOperators(svd, A, B);
  if (!ierr && A == shim_last_A) ierr = SVDSetTransposeMatrix(svd, shim_last_AT);
  return ierr;
}
static PetscErrorCode ShimEPSCreate(EPS *eps)
{
  BV bv;
  PetscErrorCode ierr = EPSCreate(eps);
  if (!ierr) ierr = EPSGetBV(*eps, &bv);
  if (!ierr) ierr = BVSetType(bv, "oraclecpu");
  return ierr;
}
static PetscErrorCode ShimSVDCreate(SVD *svd)
{
  BV v, u;
  PetscErrorCode ierr = SVDCreate(svd);
  if (!ierr) ierr = SVDGetBV(*svd, &v, &u);
  if (!ierr) ierr = BVSetType(v, "oraclecpu");
  if (!ierr) ierr = BVSetType(u, "oraclecpu");
  return ierr;
}
#define B2KInitialize    ShimInit
#define MatCreateB200CSR ShimMatCSR
#define EPSCreate        ShimEPSCreate
#define SVDCreate        ShimSVDCreate
#define VecCreateB200    VecCreateHost
#undef  BVB200
#define BVB200           "oraclecpu"
#define SVDSetOperators  ShimSVDSetOperators
#endif
