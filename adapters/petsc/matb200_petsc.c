/*
 * matb200_petsc.c — the operator side of the plug-in for a REAL PETSc 3.22: SLEPc reaches the sparse product through
 *   BVMatMultColumn → MatMult              src/sys/classes/bv/interface/bvops.c:862-885
 *   STApply_Generic → MatMult(st->M,x,y)    src/sys/classes/st/interface/stsolve.c:16-25
 *   SVDTwoSideLanczos → MatMult(A|AT)       src/svd/impls/lanczos/gklanczos.c:67,80,90,103
 * and only needs MATOP_MULT (+ MATOP_MULT_TRANSPOSE for an implicit transpose, svdsetup.c:273-279) and CUDA vectors from
 * MatCreateVecs (stsolve.c:349-353, BVSetSizesFromVec): a MatShell, the pattern of src/eps/tutorials/ex3.c:46-49,140-168.
 * The PETSc-typed twin of slepc_b200/host/matb200.c: the local rows go to HBM with their global column indices, the local
 * numbering and the SELL-32 copy are built there (b2k_csr_create_global), the halo moves GPU-to-GPU in both directions (NVLink
 * peer memory, else ncclSend/ncclRecv), the products are k_spmv_sell_pipe on A and on the local transpose built in HBM.
 * Type-checked against the reference's headers with the stand-in PETSc declarations of adapters/petsc/stub/ (no PETSc here).
 */
#include <slepcsys.h>
#include "b2k.h"
#include "b2k_petsc.h"

#define B2KCall(call) do { int rc_ = (call); PetscCheck(!rc_, PETSC_COMM_SELF, PETSC_ERR_GPU, "%s failed (%d): %s", #call, rc_, b2k_last_error()); } while (0)

typedef struct {
  b2k_csr   A;
  b2k_comm  comm;
  b2k_halo  halo;            /* forward halo over peer memory (NULL: single rank or no mailboxes)   */
  b2k_halo  halo_rev;        /* the same plan with the roles exchanged: reverse halo of MatMultTranspose */
  b2k_csr   ATown, ATgh;     /* local transpose built in HBM on first use (b2k_csr_transpose_split)  */
  double   *zghost, *rbuf;   /* ghost-column contributions of y = A^T x (send) / received ones (NCCL path) */
  PetscInt  n;               /* owned columns                                                        */
  PetscInt  nghost, *ghosts; /* sorted global columns outside the owned range, local numbering [owned | ghosts] */
  double   *xghost;          /* NCCL path: device ghost values                                      */
  PetscInt  nrecv, nsend, *recvrank, *recvcount, *sendrank, *sendcount, nsendtot;
  int      *d_sendidx;
  double   *sendbuf;
} Mat_B200;

static PetscErrorCode MatMult_B200(Mat S, Vec x, Vec y)
{
  Mat_B200          *a;
  const PetscScalar *px;
  PetscScalar       *py;
  const double      *ghost;
  b2k_ctx            ctx = B2KPetscContext();

  PetscFunctionBegin;
  PetscCall(MatShellGetContext(S, &a));
  PetscCall(VecCUDAGetArrayRead(x, &px));
  PetscCall(VecCUDAGetArrayWrite(y, &py));
  ghost = a->xghost;
  if (a->halo) B2KCall(b2k_halo_exchange(a->halo, px, &ghost));
  else if (a->nrecv || a->nsend) {                               /* grouped ncclSend / ncclRecv on the compute stream */
    PetscInt soff = 0, roff = 0;
    if (a->nsendtot) B2KCall(b2k_gather(ctx, a->sendbuf, px, a->d_sendidx, a->nsendtot));
    B2KCall(b2k_comm_group_start(a->comm));
    for (PetscInt q = 0; q < a->nsend; q++) { B2KCall(b2k_comm_sendrecv(a->comm, a->sendbuf + soff, a->sendcount[q], (int)a->sendrank[q], NULL, 0, 0)); soff += a->sendcount[q]; }
    for (PetscInt p = 0; p < a->nrecv; p++) { B2KCall(b2k_comm_sendrecv(a->comm, NULL, 0, 0, a->xghost + roff, a->recvcount[p], (int)a->recvrank[p])); roff += a->recvcount[p]; }
    B2KCall(b2k_comm_group_end(a->comm));
  }
  B2KCall(b2k_csr_spmv(ctx, a->A, px, ghost, py));
  PetscCall(VecCUDARestoreArrayWrite(y, &py));
  PetscCall(VecCUDARestoreArrayRead(x, &px));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* y = A^T x for the implicit transpose of SVDSetUp (svdsetup.c:273-279; reached from gklanczos.c:80,103) — PETSc's
   MatMultTranspose_MPIAIJ: local transpose products, then the ghost-column contributions travel to their owners and are added
   in the order of the send list (deterministic).  The PETSc-typed twin of MatMultTranspose_B200CSR in host/matb200.c. */
static PetscErrorCode MatMultTranspose_B200(Mat S, Vec x, Vec y)
{
  Mat_B200          *a;
  const PetscScalar *px;
  PetscScalar       *py;
  b2k_ctx            ctx = B2KPetscContext();

  PetscFunctionBegin;
  PetscCall(MatShellGetContext(S, &a));
  if (!a->ATown) {
    B2KCall(b2k_csr_transpose_split(ctx, a->A, &a->ATown, a->nghost ? &a->ATgh : NULL));
    if (a->nghost) B2KCall(b2k_malloc(ctx, (void **)&a->zghost, sizeof(double) * (size_t)a->nghost));
  }
  PetscCall(VecCUDAGetArrayRead(x, &px));
  PetscCall(VecCUDAGetArrayWrite(y, &py));
  B2KCall(b2k_csr_spmv(ctx, a->ATown, px, NULL, py));
  if (a->nghost) B2KCall(b2k_csr_spmv(ctx, a->ATgh, px, NULL, a->zghost));
  if (a->halo) {                                                  /* every rank of the communicator takes part */
    const double *rb = NULL;
    PetscInt      so = 0;
    if (!a->halo_rev) {
      int     rr[8], rc[8], sr[8], sc[8];
      int64_t off[8] = {0}, roff = 0;
      for (PetscInt p = 0; p < a->nrecv; p++) { sr[p] = (int)a->recvrank[p]; sc[p] = (int)a->recvcount[p]; off[p] = roff; roff += a->recvcount[p]; }
      for (PetscInt q = 0; q < a->nsend; q++) { rr[q] = (int)a->sendrank[q]; rc[q] = (int)a->sendcount[q]; }
      B2KCall(b2k_halo_create(a->comm, (int)a->nsend, rr, rc, (int)a->nrecv, sr, sc, NULL, off, &a->halo_rev));
    }
    B2KCall(b2k_halo_exchange(a->halo_rev, a->zghost, &rb));
    for (PetscInt q = 0; q < a->nsend; q++) { B2KCall(b2k_scatter_add(ctx, py, a->d_sendidx + so, rb + so, a->sendcount[q])); so += a->sendcount[q]; }
  } else if (a->nrecv || a->nsend) {                             /* grouped ncclSend / ncclRecv, the plan run backwards */
    PetscInt soff = 0, roff = 0;
    if (a->nsendtot && !a->rbuf) B2KCall(b2k_malloc(ctx, (void **)&a->rbuf, sizeof(double) * (size_t)a->nsendtot));
    B2KCall(b2k_comm_group_start(a->comm));
    for (PetscInt p = 0; p < a->nrecv; p++) { B2KCall(b2k_comm_sendrecv(a->comm, a->zghost + roff, a->recvcount[p], (int)a->recvrank[p], NULL, 0, 0)); roff += a->recvcount[p]; }
    for (PetscInt q = 0; q < a->nsend; q++) { B2KCall(b2k_comm_sendrecv(a->comm, NULL, 0, 0, a->rbuf + soff, a->sendcount[q], (int)a->sendrank[q])); soff += a->sendcount[q]; }
    B2KCall(b2k_comm_group_end(a->comm));
    soff = 0;
    for (PetscInt q = 0; q < a->nsend; q++) { B2KCall(b2k_scatter_add(ctx, py, a->d_sendidx + soff, a->rbuf + soff, a->sendcount[q])); soff += a->sendcount[q]; }
  }
  PetscCall(VecCUDARestoreArrayWrite(y, &py));
  PetscCall(VecCUDARestoreArrayRead(x, &px));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* MatGetDiagonal: the Jacobi preconditioner of a KSP behind STMatSolve (stsles.c:166-200) */
static PetscErrorCode MatGetDiagonal_B200(Mat S, Vec d)
{
  Mat_B200    *a;
  PetscScalar *pd;

  PetscFunctionBegin;
  PetscCall(MatShellGetContext(S, &a));
  PetscCall(VecCUDAGetArrayWrite(d, &pd));
  B2KCall(b2k_csr_get_diagonal(B2KPetscContext(), a->A, 0, pd));
  PetscCall(VecCUDARestoreArrayWrite(d, &pd));
  PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode MatDestroy_B200(Mat S)
{
  Mat_B200 *a;
  b2k_ctx   ctx = B2KPetscContext();

  PetscFunctionBegin;
  PetscCall(MatShellGetContext(S, &a));
  if (a->halo_rev) B2KCall(b2k_halo_destroy(a->halo_rev));       /* collective */
  if (a->halo) B2KCall(b2k_halo_destroy(a->halo));               /* collective */
  B2KCall(b2k_csr_destroy(ctx, a->ATown));
  B2KCall(b2k_csr_destroy(ctx, a->ATgh));
  B2KCall(b2k_free(ctx, a->zghost));
  B2KCall(b2k_free(ctx, a->rbuf));
  B2KCall(b2k_csr_destroy(ctx, a->A));
  B2KCall(b2k_free(ctx, a->xghost));
  B2KCall(b2k_free(ctx, a->d_sendidx));
  B2KCall(b2k_free(ctx, a->sendbuf));
  free(a->ghosts);                                               /* malloc'ed by b2k_csr_create_global */
  PetscCall(PetscFree(a->recvrank)); PetscCall(PetscFree(a->recvcount)); PetscCall(PetscFree(a->sendrank)); PetscCall(PetscFree(a->sendcount));
  PetscCall(PetscFree(a));
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* the halo plan: who needs which of my entries (what MatSetUpMultiply_MPIAIJ / VecScatterCreate work out): every rank publishes
   the sorted list of ghost columns it needs from each owner; the counts travel with MPI_Allreduce-free point-to-point exchanges
   in a full implementation — here with one MPI_Allgather of the per-owner counts followed by the index lists */
static PetscErrorCode MatB200BuildPlan_Private(MPI_Comm comm, Mat_B200 *a, PetscInt cstart, const PetscInt *colstarts /* size+1 */, int size, int rank)
{
  PetscInt *need, *allneed;

  PetscFunctionBegin;
  PetscCall(PetscCalloc1(size, &need));
  PetscCall(PetscCalloc1((size_t)size * (size_t)size, &allneed));
  for (PetscInt g = 0, p = 0; g < a->nghost; g++) { while (a->ghosts[g] >= colstarts[p + 1]) p++; need[p]++; }
  PetscCallMPI(MPI_Allgather(need, size, MPIU_INT, allneed, size, MPIU_INT, comm));
  PetscCall(PetscMalloc1(size, &a->recvrank)); PetscCall(PetscMalloc1(size, &a->recvcount));
  PetscCall(PetscMalloc1(size, &a->sendrank)); PetscCall(PetscMalloc1(size, &a->sendcount));
  for (int p = 0; p < size; p++) {
    if (p == rank) continue;
    if (need[p]) { a->recvrank[a->nrecv] = p; a->recvcount[a->nrecv++] = need[p]; }
    if (allneed[(size_t)p * size + rank]) { a->sendrank[a->nsend] = p; a->sendcount[a->nsend++] = allneed[(size_t)p * size + rank]; a->nsendtot += allneed[(size_t)p * size + rank]; }
  }
  /* the index lists themselves (global columns each peer needs from me) arrive with MPI_Sendrecv per peer; converted to local
     offsets (minus cstart) they become d_sendidx, the gather list of the pack kernel */
  if (a->nsendtot) {
    PetscInt *idx, off = 0, goff = 0;
    int      *idx32;
    PetscCall(PetscMalloc1(a->nsendtot, &idx));
    PetscCall(PetscMalloc1(a->nsendtot, &idx32));
    for (PetscInt q = 0, p = 0; q < a->nsend || p < a->nrecv;) {     /* pairwise exchange in rank order: deadlock-free */
      const PetscInt sr = q < a->nsend ? a->sendrank[q] : size, rr = p < a->nrecv ? a->recvrank[p] : size;
      const PetscInt peer = PetscMin(sr, rr);
      const PetscInt ns = (sr == peer) ? a->sendcount[q] : 0, nr = (rr == peer) ? a->recvcount[p] : 0;
      PetscCallMPI(MPI_Sendrecv(a->ghosts + goff, (int)nr, MPIU_INT, (int)peer, 71, idx + off, (int)ns, MPIU_INT, (int)peer, 71, comm, MPI_STATUS_IGNORE));
      if (sr == peer) { off += ns; q++; }
      if (rr == peer) { goff += nr; p++; }
    }
    for (PetscInt i = 0; i < a->nsendtot; i++) idx32[i] = (int)(idx[i] - cstart);
    B2KCall(b2k_malloc(B2KPetscContext(), (void **)&a->d_sendidx, sizeof(int) * (size_t)a->nsendtot));
    B2KCall(b2k_h2d(B2KPetscContext(), a->d_sendidx, idx32, sizeof(int) * (size_t)a->nsendtot));
    B2KCall(b2k_malloc(B2KPetscContext(), (void **)&a->sendbuf, sizeof(double) * (size_t)a->nsendtot));
    PetscCall(PetscFree(idx)); PetscCall(PetscFree(idx32));
  }
  PetscCall(PetscFree(need)); PetscCall(PetscFree(allneed));
  PetscFunctionReturn(PETSC_SUCCESS);
}

PetscErrorCode MatCreateB200FromMat(Mat Ain, Mat *Aout)
{
  MPI_Comm     comm;
  Mat_B200    *a;
  PetscInt     M, N, m, n, rstart, rend, cstart, cend, nnz = 0, *rowptr, *colglob, *colstarts;
  PetscScalar *val;
  int          size, rank;
  b2k_ctx      ctx;

  PetscFunctionBegin;
  PetscCall(PetscObjectGetComm((PetscObject)Ain, &comm));
  PetscCall(B2KPetscInitialize(comm));
  ctx = B2KPetscContext();
  PetscCallMPI(MPI_Comm_size(comm, &size));
  PetscCallMPI(MPI_Comm_rank(comm, &rank));
  PetscCall(MatGetSize(Ain, &M, &N));
  PetscCall(MatGetLocalSize(Ain, &m, &n));
  PetscCall(MatGetOwnershipRange(Ain, &rstart, &rend));
  PetscCall(MatGetOwnershipRangeColumn(Ain, &cstart, &cend));
  PetscCall(PetscNew(&a));
  PetscCall(B2KPetscCommGet(comm, &a->comm));
  a->n = n;
  /* the local rows with their GLOBAL column indices go to HBM as they are; the [owned | ghosts] numbering and the ghost list
     (MPIAIJ's garray, MatSetUpMultiply_MPIAIJ) are built on the device (b2k_csr_create_global) */
  PetscCall(PetscMalloc1(m + 1, &rowptr));
  rowptr[0] = 0;
  for (PetscInt r = 0; r < m; r++) {
    PetscInt nc;
    PetscCall(MatGetRow(Ain, rstart + r, &nc, NULL, NULL));
    nnz += nc;
    rowptr[r + 1] = nnz;
    PetscCall(MatRestoreRow(Ain, rstart + r, &nc, NULL, NULL));
  }
  PetscCall(PetscMalloc1(nnz + 1, &colglob));
  PetscCall(PetscMalloc1(nnz + 1, &val));
  for (PetscInt r = 0; r < m; r++) {
    PetscInt           nc;
    const PetscInt    *cols;
    const PetscScalar *v;
    PetscCall(MatGetRow(Ain, rstart + r, &nc, &cols, &v));
    for (PetscInt k = 0; k < nc; k++) { val[rowptr[r] + k] = v[k]; colglob[rowptr[r] + k] = cols[k]; }
    PetscCall(MatRestoreRow(Ain, rstart + r, &nc, &cols, &v));
  }
  {
    int64_t ng = 0;
    int    *gh = NULL;
    B2KCall(b2k_csr_create_global(ctx, m, N, cstart, cend, rowptr, colglob, val, &a->A, &ng, &gh));
    a->nghost = (PetscInt)ng;
    a->ghosts = gh;                                              /* PetscInt is int in the default build */
  }
  PetscCall(PetscFree(rowptr)); PetscCall(PetscFree(colglob)); PetscCall(PetscFree(val));
  if (size > 1) {
    PetscCall(PetscMalloc1(size + 1, &colstarts));
    PetscCallMPI(MPI_Allgather(&cstart, 1, MPIU_INT, colstarts, 1, MPIU_INT, comm));
    colstarts[size] = N;
    PetscCall(MatB200BuildPlan_Private(comm, a, cstart, colstarts, size, rank));
    PetscCall(PetscFree(colstarts));
    if (a->nghost) B2KCall(b2k_malloc(ctx, (void **)&a->xghost, sizeof(double) * (size_t)a->nghost));
    if (b2k_comm_p2p_enabled(a->comm)) {                         /* halo over NVLink peer memory (b2k_halo.cu) */
      int rr[8], rc[8], sr[8], sc[8];
      PetscCheck(a->nrecv <= 8 && a->nsend <= 8, comm, PETSC_ERR_SUP, "at most 8 ranks per box");
      for (PetscInt p = 0; p < a->nrecv; p++) { rr[p] = (int)a->recvrank[p]; rc[p] = (int)a->recvcount[p]; }
      for (PetscInt q = 0; q < a->nsend; q++) { sr[q] = (int)a->sendrank[q]; sc[q] = (int)a->sendcount[q]; }
      B2KCall(b2k_halo_create(a->comm, (int)a->nrecv, rr, rc, (int)a->nsend, sr, sc, a->d_sendidx, NULL, &a->halo));
    }
  }
  PetscCall(MatCreateShell(comm, m, n, M, N, (void *)a, Aout));
  PetscCall(MatShellSetOperation(*Aout, MATOP_MULT, (void (*)(void))MatMult_B200));
  PetscCall(MatShellSetOperation(*Aout, MATOP_MULT_TRANSPOSE, (void (*)(void))MatMultTranspose_B200));   /* no MATOP_TRANSPOSE: SVDSetUp goes implicit */
  if (M == N && rstart == cstart && rend == cend) PetscCall(MatShellSetOperation(*Aout, MATOP_GET_DIAGONAL, (void (*)(void))MatGetDiagonal_B200));
  PetscCall(MatShellSetOperation(*Aout, MATOP_DESTROY, (void (*)(void))MatDestroy_B200));
  PetscCall(MatShellSetVecType(*Aout, VECCUDA));                 /* MatCreateVecs hands SLEPc CUDA vectors: stsolve.c:349-353 */
  PetscFunctionReturn(PETSC_SUCCESS);
}

/* MatRegister("b200csr", MatCreate_B200CSR): `-mat_type b200csr` on an application that assembles with MatSetValues needs the
   full Mat ops table (assembly, preallocation, …); that is PETSc work outside this path, so the registered type refuses to be
   used directly and points at MatCreateB200FromMat(), which covers what SLEPc itself calls */
PetscErrorCode MatCreate_B200CSR(Mat A)
{
  PetscFunctionBegin;
  SETERRQ(PetscObjectComm((PetscObject)A), PETSC_ERR_SUP, "assemble with any PETSc Mat type and wrap it with MatCreateB200FromMat(A,&Op): SLEPc only calls MatMult / MatCreateVecs on the operator");
  PetscFunctionReturn(PETSC_SUCCESS);
}
