/*
 * matb200.c — Mat type "b200csr": the operator of the eigen/singular-value problem as FP64 CSR rows
 * resident in HBM, row-partitioned over the GPUs of one box.  It supplies what SLEPc reaches through
 *   BVMatMultColumn → MatMult            src/sys/classes/bv/interface/bvops.c:862-885
 *   STApply_Generic → MatMult(st->M,x,y) src/sys/classes/st/interface/stsolve.c:16-25
 *   SVDTwoSideLanczos → MatMult(A|AT)    src/svd/impls/lanczos/gklanczos.c:67,80,90,103
 * i.e. PETSc's MatMult_MPIAIJ(CUSPARSE): VecScatter halo exchange + local SpMV.  Here the halo moves
 * GPU-to-GPU over NVLink (NCCL send/recv on the compute stream) and the SpMV is b2k_csr_spmv.
 * With a real PETSc the same object is a MatShell (MATOP_MULT) or a registered Mat type, INTEGRATION.md.
 */
#include "b2kimpl.h"

typedef struct {
  b2k_csr   A;
  PetscInt  nghost;
  PetscInt *ghosts;          /* sorted global column indices outside [cstart,cend)            */
  double   *xghost;          /* device, nghost doubles: halo values in ghost order             */
  /* halo plan */
  PetscBool halo_set;
  PetscInt  nrecv, nsend;
  PetscInt *recvrank, *recvcount, *sendrank, *sendcount;
  PetscInt *sendoff;         /* contiguous sends: offset into x (sendidx == NULL)              */
  int      *d_sendidx;       /* device: local indices to pack (general plan)                   */
  double   *sendbuf;         /* device pack buffer                                             */
  PetscInt  nsendtot;
  int64_t   nnz;
  /* implicit transpose (MatMultTranspose): local A_loc^T split into the owned-column rows and the ghost-column rows,
     built on first use; zghost/rbuf are the reverse-halo buffers */
  b2k_csr   ATown, ATgh;
  double   *zghost, *rbuf;
  /* default on one NVSwitch box (B2K_HALO_P2P=0 turns it off): the forward halo pushed over NVLink peer memory instead of ncclSend/ncclRecv */
  b2k_halo  halo;
  b2k_halo  halo_rev;        /* the same plan run backwards (MatMultTranspose): ghost-column contributions pushed to their owners     */
  double   *gblock;          /* device: ghost values of a block of vectors, nghost x gblock_cols (MatMultBlock) */
  PetscInt  gblock_cols;
} Mat_B200CSR;

#define CTX() B2KGetContext()

/* collective: turn the installed halo plan into a peer-memory halo object whenever the NVLink mailboxes are up (default;
   B2K_HALO_P2P=0 keeps the grouped ncclSend/ncclRecv) */
static PetscErrorCode MatHaloSetUpP2P_B200CSR(Mat A)
{
  Mat_B200CSR *a = (Mat_B200CSR *)A->data;
  B2KComm comm = B2KCommWorld();
  const char *e = getenv("B2K_HALO_P2P");
  if ((e && e[0] == '0') || !comm || comm->kind != 1 || comm->size < 2 || comm->size > 8 || !b2k_comm_p2p_enabled(comm->nccl)) return PETSC_SUCCESS;
  if (a->halo_rev) { B2KCall(b2k_halo_destroy(a->halo_rev)); a->halo_rev = NULL; }
  if (a->halo) { B2KCall(b2k_halo_destroy(a->halo)); a->halo = NULL; }
  int64_t off[8] = {0};
  for (PetscInt q = 0; q < a->nsend && q < 8; q++) off[q] = a->sendoff ? a->sendoff[q] : 0;
  B2KCall(b2k_halo_create(comm->nccl, a->nrecv, a->recvrank, a->recvcount, a->nsend, a->sendrank, a->sendcount, a->d_sendidx,
                          a->d_sendidx ? NULL : off, &a->halo));
  return PETSC_SUCCESS;
}

/* on return *ghost points at the ghost values the SpMV must read */
static PetscErrorCode MatHaloExchange_B200CSR(Mat A, const double *x, const double **ghost)
{
  Mat_B200CSR *a = (Mat_B200CSR *)A->data;
  B2KComm comm = B2KCommWorld();
  int size = 1;
  *ghost = a->xghost;
  PetscCall(B2KCommGetRank(comm, NULL, &size));
  if (a->halo) {                                  /* collective: every rank of the communicator owns one */
    int bad = 0;
    B2KCall(b2k_halo_error(a->halo, &bad));       /* raised by an earlier exchange whose peer did not show up in time */
    PetscCheck(!bad, PETSC_ERR_LIB, "peer-memory halo exchange timed out (code %d): a rank did not take part in MatMult", bad);
    B2KCall(b2k_halo_exchange(a->halo, x, ghost));
    return PETSC_SUCCESS;
  }
  if (size == 1 || (a->nghost == 0 && a->nsend == 0)) {       /* a rank that only SENDS halo values still takes part */
    PetscCheck(a->nghost == 0, PETSC_ERR_ARG_WRONGSTATE, "matrix has %d ghost columns but the communicator has a single rank", a->nghost);
    return PETSC_SUCCESS;
  }
  PetscCheck(a->halo_set, PETSC_ERR_ORDER, "MatB200CSRSetHalo() must be called before MatMult() on more than one rank");
  PetscCheck(comm->kind == 1, PETSC_ERR_SUP, "b200csr needs the NCCL communicator (B2KCommInitNCCL)");
  if (a->d_sendidx && a->nsendtot > 0) B2KCall(b2k_gather(CTX(), a->sendbuf, x, a->d_sendidx, a->nsendtot));
  B2KCall(b2k_comm_group_start(comm->nccl));
  PetscInt soff = 0, roff = 0;
  for (PetscInt q = 0; q < a->nsend; q++) {
    const double *sb = a->d_sendidx ? a->sendbuf + soff : x + a->sendoff[q];
    B2KCall(b2k_comm_sendrecv(comm->nccl, sb, a->sendcount[q], a->sendrank[q], NULL, 0, 0));
    soff += a->sendcount[q];
  }
  for (PetscInt p = 0; p < a->nrecv; p++) {
    B2KCall(b2k_comm_sendrecv(comm->nccl, NULL, 0, 0, a->xghost + roff, a->recvcount[p], a->recvrank[p]));
    roff += a->recvcount[p];
  }
  B2KCall(b2k_comm_group_end(comm->nccl));
  return PETSC_SUCCESS;
}

static PetscErrorCode MatMult_B200CSR(Mat A, Vec x, Vec y)
{
  Mat_B200CSR *a = (Mat_B200CSR *)A->data;
  PetscCheck(x->mem == B2K_MEM_DEVICE && y->mem == B2K_MEM_DEVICE, PETSC_ERR_ARG_INCOMP, "b200csr needs device vectors");
  const double *ghost;
  PetscCall(MatHaloExchange_B200CSR(A, x->array, &ghost));
  B2KCall(b2k_csr_spmv(CTX(), a->A, x->array, ghost, y->array));
  return PETSC_SUCCESS;
}

/* Y(:,0:k) = A X(:,0:k): one halo exchange per column into a ghost block, then ONE block product that reads the matrix once
   per 16 columns (b2k_csr_spmm) — BVMatMult in BV_MATMULT_MAT mode (svec.c:203-231) */
static PetscErrorCode MatMultBlock_B200CSR(Mat A, const PetscScalar *X, PetscInt ldx, PetscScalar *Y, PetscInt ldy, PetscInt k)
{
  Mat_B200CSR *a = (Mat_B200CSR *)A->data;
  b2k_ctx ctx = CTX();
  int size = 1;
  PetscCall(B2KCommGetRank(B2KCommWorld(), NULL, &size));
  if (size == 1 || (a->nghost == 0 && a->nsend == 0)) {
    B2KCall(b2k_csr_spmm(ctx, a->A, X, ldx, NULL, 0, Y, ldy, k));
    return PETSC_SUCCESS;
  }
  if (a->nghost && a->gblock_cols < k) {
    if (a->gblock) B2KCall(b2k_free(ctx, a->gblock));
    a->gblock = NULL;
    B2KCall(b2k_malloc(ctx, (void **)&a->gblock, sizeof(double) * (size_t)a->nghost * (size_t)k));
    a->gblock_cols = k;
  }
  for (PetscInt j = 0; j < k; j++) {              /* collective per column, also on a rank that only sends */
    const double *ghost;
    PetscCall(MatHaloExchange_B200CSR(A, X + (size_t)j * (size_t)ldx, &ghost));
    if (a->nghost) B2KCall(b2k_d2d(ctx, a->gblock + (size_t)j * (size_t)a->nghost, ghost, sizeof(double) * (size_t)a->nghost));
  }
  B2KCall(b2k_csr_spmm(ctx, a->A, X, ldx, a->gblock, a->nghost, Y, ldy, k));
  return PETSC_SUCCESS;
}

/* A_loc^T built in HBM (b2k_csr_transpose_split: stable radix sort of the column indices + gather): rows = local columns
   [owned | ghosts], columns = local rows.  Set-up cost, once; the owned and the ghost block become two matrices so that the
   owned part of y = A^T x is written in place and only the ghost part travels. */
static PetscErrorCode MatBuildLocalTranspose_B200CSR(Mat A)
{
  Mat_B200CSR *a = (Mat_B200CSR *)A->data;
  b2k_ctx ctx = CTX();
  int rc = b2k_csr_transpose_split(ctx, a->A, &a->ATown, a->nghost ? &a->ATgh : NULL);
  if (!rc && a->nghost) rc = b2k_malloc(ctx, (void **)&a->zghost, sizeof(double) * (size_t)a->nghost);
  if (!rc && a->nsendtot && !a->rbuf) rc = b2k_malloc(ctx, (void **)&a->rbuf, sizeof(double) * (size_t)a->nsendtot);
  PetscCheck(!rc, PETSC_ERR_GPU, "building the local transpose failed (%d): %s", rc, b2k_last_error());
  return PETSC_SUCCESS;
}

/* collective, on the first MatMultTranspose: the peer-memory halo with the roles exchanged — what I receive in MatMult (ghost
   segment p, contiguous in zghost) I now push to its owner, and from every rank I send to in MatMult I receive the contributions
   to the entries I sent (in the order of my send list) */
static PetscErrorCode MatHaloReverseSetUp_B200CSR(Mat A)
{
  Mat_B200CSR *a = (Mat_B200CSR *)A->data;
  B2KComm comm = B2KCommWorld();
  int64_t off[8] = {0};
  PetscInt roff = 0;
  for (PetscInt p = 0; p < a->nrecv && p < 8; p++) { off[p] = roff; roff += a->recvcount[p]; }
  B2KCall(b2k_halo_create(comm->nccl, a->nsend, a->sendrank, a->sendcount, a->nrecv, a->recvrank, a->recvcount, NULL, off, &a->halo_rev));
  return PETSC_SUCCESS;
}

/* y = A^T x with A row-partitioned: local y_own = A_own^T x, ghost-column contributions z = A_gh^T x are sent to the owners
   of those columns (the halo plan run backwards) and accumulated in peer order — deterministic.  This is PETSc's
   MatMultTranspose_MPIAIJ (local transpose products + VecScatter SCATTER_REVERSE/ADD_VALUES), reached from
   SVDTwoSideLanczos (gklanczos.c:80,103) when the transpose is implicit (svdsetup.c:273-279,309-315). */
static PetscErrorCode MatMultTranspose_B200CSR(Mat A, Vec x, Vec y)
{
  Mat_B200CSR *a = (Mat_B200CSR *)A->data;
  b2k_ctx ctx = CTX();
  B2KComm comm = B2KCommWorld();
  int size = 1;
  PetscCheck(x->mem == B2K_MEM_DEVICE && y->mem == B2K_MEM_DEVICE, PETSC_ERR_ARG_INCOMP, "b200csr needs device vectors");
  PetscCall(B2KCommGetRank(comm, NULL, &size));
  if (!a->ATown) PetscCall(MatBuildLocalTranspose_B200CSR(A));
  B2KCall(b2k_csr_spmv(ctx, a->ATown, x->array, NULL, y->array));
  if (size == 1 || (!a->halo && a->nghost == 0 && a->nsend == 0)) return PETSC_SUCCESS;    /* with a peer-memory halo every rank takes part */
  PetscCheck(a->halo_set || a->halo, PETSC_ERR_ORDER, "MatB200CSRSetHalo() must be called before MatMultTranspose() on more than one rank");
  PetscCheck(comm->kind == 1, PETSC_ERR_SUP, "b200csr needs the NCCL communicator (B2KCommInitNCCL)");
  if (a->nghost) B2KCall(b2k_csr_spmv(ctx, a->ATgh, x->array, NULL, a->zghost));
  if (a->halo) {
    /* reverse halo over NVLink peer memory (default on one box): push + flags instead of grouped ncclSend/ncclRecv, then the
       contributions are added in the order of the send list — the same fixed order as below */
    int bad = 0;
    if (!a->halo_rev) PetscCall(MatHaloReverseSetUp_B200CSR(A));
    B2KCall(b2k_halo_error(a->halo_rev, &bad));
    PetscCheck(!bad, PETSC_ERR_LIB, "peer-memory reverse halo exchange timed out (code %d): a rank did not take part in MatMultTranspose", bad);
    const double *rb = NULL;
    B2KCall(b2k_halo_exchange(a->halo_rev, a->zghost, &rb));
    PetscInt so = 0;
    for (PetscInt q = 0; q < a->nsend; q++) {
      if (a->d_sendidx) B2KCall(b2k_scatter_add(ctx, y->array, a->d_sendidx + so, rb + so, a->sendcount[q]));
      else B2KCall(b2k_axpby(ctx, y->array + a->sendoff[q], A->n, rb + so, a->sendcount[q], a->sendcount[q], 1, 1.0, 1.0));
      so += a->sendcount[q];
    }
    return PETSC_SUCCESS;
  }
  if (a->nsendtot && !a->rbuf) B2KCall(b2k_malloc(ctx, (void **)&a->rbuf, sizeof(double) * (size_t)a->nsendtot));
  B2KCall(b2k_comm_group_start(comm->nccl));
  PetscInt roff = 0, soff = 0;
  for (PetscInt p = 0; p < a->nrecv; p++) {        /* what I receive in MatMult I now send back */
    B2KCall(b2k_comm_sendrecv(comm->nccl, a->zghost + roff, a->recvcount[p], a->recvrank[p], NULL, 0, 0));
    roff += a->recvcount[p];
  }
  for (PetscInt q = 0; q < a->nsend; q++) {
    B2KCall(b2k_comm_sendrecv(comm->nccl, NULL, 0, 0, a->rbuf + soff, a->sendcount[q], a->sendrank[q]));
    soff += a->sendcount[q];
  }
  B2KCall(b2k_comm_group_end(comm->nccl));
  soff = 0;
  for (PetscInt q = 0; q < a->nsend; q++) {
    if (a->d_sendidx) B2KCall(b2k_scatter_add(ctx, y->array, a->d_sendidx + soff, a->rbuf + soff, a->sendcount[q]));
    else B2KCall(b2k_axpby(ctx, y->array + a->sendoff[q], A->n, a->rbuf + soff, a->sendcount[q], a->sendcount[q], 1, 1.0, 1.0));
    soff += a->sendcount[q];
  }
  return PETSC_SUCCESS;
}

static PetscErrorCode MatGetDiagonal_B200CSR(Mat A, Vec d)
{
  Mat_B200CSR *a = (Mat_B200CSR *)A->data;
  PetscCheck(d->mem == B2K_MEM_DEVICE, PETSC_ERR_ARG_INCOMP, "b200csr needs device vectors");
  PetscCheck(A->M == A->N && A->rstart == A->cstart && A->m == A->n, PETSC_ERR_SUP, "MatGetDiagonal needs a square matrix whose owned columns are its owned rows");
  B2KCall(b2k_csr_get_diagonal(CTX(), a->A, 0, d->array));
  return PETSC_SUCCESS;
}

static PetscErrorCode MatDestroy_B200CSR(Mat A)
{
  Mat_B200CSR *a = (Mat_B200CSR *)A->data;
  if (!a) return PETSC_SUCCESS;
  b2k_ctx ctx = CTX();
  if (ctx) {
    if (a->halo_rev) b2k_halo_destroy(a->halo_rev); /* collective */
    if (a->halo) b2k_halo_destroy(a->halo);         /* collective */
    b2k_csr_destroy(ctx, a->ATown); b2k_csr_destroy(ctx, a->ATgh);
    b2k_free(ctx, a->zghost); b2k_free(ctx, a->rbuf); b2k_free(ctx, a->gblock);
    b2k_csr_destroy(ctx, a->A);
    b2k_free(ctx, a->xghost); b2k_free(ctx, a->d_sendidx); b2k_free(ctx, a->sendbuf);
  }
  free(a->ghosts); free(a->recvrank); free(a->recvcount); free(a->sendrank); free(a->sendcount); free(a->sendoff);
  free(a);
  A->data = NULL;
  return PETSC_SUCCESS;
}

static PetscErrorCode MatSetUp_B200CSR(Mat A, PetscInt M, PetscInt N, PetscInt rstart, PetscInt rend, PetscInt cstart, PetscInt cend, Mat_B200CSR **out)
{
  Mat_B200CSR *a = (Mat_B200CSR *)calloc(1, sizeof(*a));
  PetscCheck(a, PETSC_ERR_MEM, "out of memory");
  strcpy(A->type, "b200csr");
  A->M = M; A->N = N; A->m = rend - rstart; A->n = cend - cstart;
  A->rstart = rstart; A->rend = rend; A->cstart = cstart; A->cend = cend;
  A->mem = B2K_MEM_DEVICE;
  A->data = a;
  A->ops.mult = MatMult_B200CSR;
  A->ops.multtranspose = MatMultTranspose_B200CSR;
  A->ops.multblock = MatMultBlock_B200CSR;
  A->ops.getdiagonal = MatGetDiagonal_B200CSR;
  A->ops.destroy = MatDestroy_B200CSR;
  *out = a;
  return PETSC_SUCCESS;
}

PetscErrorCode MatCreateB200CSR(PetscInt M, PetscInt N, PetscInt rstart, PetscInt rend, const PetscInt *rowptr, const PetscInt *colidx,
                                const PetscScalar *val, PetscInt cstart, PetscInt cend, Mat *out)
{
  b2k_ctx ctx = CTX();
  PetscCheck(ctx, PETSC_ERR_ORDER, "Mat type b200csr needs a GPU context: call B2KInitialize() first (there is no CPU fallback)");
  PetscCheck(rstart >= 0 && rend >= rstart && rend <= M, PETSC_ERR_ARG_OUTOFRANGE, "row range [%d,%d) outside [0,%d)", rstart, rend, M);
  PetscCheck(cstart >= 0 && cend >= cstart && cend <= N, PETSC_ERR_ARG_OUTOFRANGE, "column range [%d,%d) outside [0,%d)", cstart, cend, N);
  PetscCheck(rowptr && (rend == rstart || rowptr[0] == 0), PETSC_ERR_ARG_WRONG, "rowptr[0] must be 0");
  const PetscInt m = rend - rstart;
  const PetscInt nnz = m ? rowptr[m] : 0;
  Mat A;
  Mat_B200CSR *a;
  PetscCall(MatCreate_Private(&A));
  PetscCall(MatSetUp_B200CSR(A, M, N, rstart, rend, cstart, cend, &a));
  a->nnz = nnz;
  /* ghosts = sorted unique off-range columns; local numbering = [owned | ghosts] like MatMPIAIJ's garray (MatSetUpMultiply_MPIAIJ).
     Built in HBM from the arrays as they are: the host does not walk the entries (b2k_csr_create_global) */
  int64_t ng = 0;
  int *gh = NULL;
  const int rc = b2k_csr_create_global(ctx, m, N, cstart, cend, rowptr, colidx, val, &a->A, &ng, &gh);
  if (rc) {
    MatDestroy(&A);
    SETERRQ(rc == B2K_ERR_ARG ? PETSC_ERR_ARG_OUTOFRANGE : PETSC_ERR_GPU, "b2k_csr_create_global failed (%d): %s", rc, b2k_last_error());
  }
  a->nghost = (PetscInt)ng;
  a->ghosts = gh;
  if (a->nghost) B2KCall(b2k_malloc(ctx, (void **)&a->xghost, sizeof(double) * (size_t)a->nghost));
  *out = A;
  return PETSC_SUCCESS;
}

PetscErrorCode MatB200CSRGetGhosts(Mat A, PetscInt *nghost, const PetscInt **ghosts)
{
  PetscCheck(!strcmp(A->type, "b200csr"), PETSC_ERR_ARG_WRONG, "not a b200csr matrix");
  Mat_B200CSR *a = (Mat_B200CSR *)A->data;
  if (nghost) *nghost = a->nghost;
  if (ghosts) *ghosts = a->ghosts;
  return PETSC_SUCCESS;
}

static PetscErrorCode dup_ints(const PetscInt *src, PetscInt n, PetscInt **dst)
{
  *dst = (PetscInt *)malloc(sizeof(PetscInt) * (size_t)(n > 0 ? n : 1));
  PetscCheck(*dst, PETSC_ERR_MEM, "out of memory");
  if (n > 0) memcpy(*dst, src, sizeof(PetscInt) * (size_t)n);
  return PETSC_SUCCESS;
}

PetscErrorCode MatB200CSRSetHalo(Mat A, PetscInt nrecv, const PetscInt *recvrank, const PetscInt *recvcount, PetscInt nsend,
                                 const PetscInt *sendrank, const PetscInt *sendcount, const PetscInt *sendidx)
{
  PetscCheck(!strcmp(A->type, "b200csr"), PETSC_ERR_ARG_WRONG, "not a b200csr matrix");
  Mat_B200CSR *a = (Mat_B200CSR *)A->data;
  b2k_ctx ctx = CTX();
  PetscInt tot = 0, stot = 0;
  for (PetscInt p = 0; p < nrecv; p++) tot += recvcount[p];
  PetscCheck(tot == a->nghost, PETSC_ERR_ARG_SIZ, "halo plan receives %d values but the matrix has %d ghost columns", tot, a->nghost);
  for (PetscInt q = 0; q < nsend; q++) stot += sendcount[q];
  for (PetscInt i = 0; i < stot; i++) PetscCheck(sendidx[i] >= 0 && sendidx[i] < A->n, PETSC_ERR_ARG_OUTOFRANGE, "send index %d outside the local columns", sendidx[i]);
  free(a->recvrank); free(a->recvcount); free(a->sendrank); free(a->sendcount);
  PetscCall(dup_ints(recvrank, nrecv, &a->recvrank));
  PetscCall(dup_ints(recvcount, nrecv, &a->recvcount));
  PetscCall(dup_ints(sendrank, nsend, &a->sendrank));
  PetscCall(dup_ints(sendcount, nsend, &a->sendcount));
  a->nrecv = nrecv; a->nsend = nsend; a->nsendtot = stot;
  if (a->d_sendidx) { B2KCall(b2k_free(ctx, a->d_sendidx)); a->d_sendidx = NULL; }
  if (a->sendbuf) { B2KCall(b2k_free(ctx, a->sendbuf)); a->sendbuf = NULL; }
  if (a->rbuf) { B2KCall(b2k_free(ctx, a->rbuf)); a->rbuf = NULL; }
  if (stot) {
    B2KCall(b2k_malloc(ctx, (void **)&a->d_sendidx, sizeof(int) * (size_t)stot));
    B2KCall(b2k_h2d(ctx, a->d_sendidx, sendidx, sizeof(int) * (size_t)stot));
    B2KCall(b2k_malloc(ctx, (void **)&a->sendbuf, sizeof(double) * (size_t)stot));
  }
  a->halo_set = PETSC_TRUE;
  PetscCall(MatHaloSetUpP2P_B200CSR(A));
  return PETSC_SUCCESS;
}

PetscErrorCode MatB200CSRGetInfo(Mat A, int64_t *nnz, int64_t *nghost)
{
  PetscCheck(!strcmp(A->type, "b200csr"), PETSC_ERR_ARG_WRONG, "not a b200csr matrix");
  Mat_B200CSR *a = (Mat_B200CSR *)A->data;
  if (nnz) *nnz = a->nnz;
  if (nghost) *nghost = a->nghost;
  return PETSC_SUCCESS;
}

/* d-dimensional Laplacian generated on the device, slab-partitioned along the slowest index
   (ex1.c:37-48 tridiag(-1,2,-1); ex2.c:39-54 5-point; 7-point analogue for C3) */
PetscErrorCode MatCreateB200Laplacian(PetscInt dim, PetscInt nx, PetscInt ny, PetscInt nz, Mat *out)
{
  b2k_ctx ctx = CTX();
  PetscCheck(ctx, PETSC_ERR_ORDER, "Mat type b200csr needs a GPU context: call B2KInitialize() first (there is no CPU fallback)");
  PetscCheck(dim >= 1 && dim <= 3, PETSC_ERR_ARG_OUTOFRANGE, "dim must be 1, 2 or 3");
  if (dim < 3) nz = 1;
  if (dim < 2) ny = 1;
  PetscCheck(nx > 0 && ny > 0 && nz > 0, PETSC_ERR_ARG_OUTOFRANGE, "grid sizes must be positive");
  const int64_t plane = (int64_t)ny * nz, Ntot = (int64_t)nx * plane;
  PetscCheck(Ntot < 2147483647LL, PETSC_ERR_ARG_OUTOFRANGE, "global size %lld exceeds PetscInt", (long long)Ntot);
  int rank = 0, size = 1;
  PetscCall(B2KCommGetRank(B2KCommWorld(), &rank, &size));
  PetscCheck(nx >= size, PETSC_ERR_ARG_SIZ, "cannot split %d planes over %d ranks", nx, size);
  /* PetscLayout-style split of the nx planes */
  const PetscInt base = nx / size, rem = nx % size;
  const PetscInt p0 = rank * base + PetscMin(rank, rem), np = base + (rank < rem ? 1 : 0);
  const int64_t row0 = (int64_t)p0 * plane, nrows = (int64_t)np * plane;
  Mat A;
  Mat_B200CSR *a;
  PetscCall(MatCreate_Private(&A));
  PetscCall(MatSetUp_B200CSR(A, (PetscInt)Ntot, (PetscInt)Ntot, (PetscInt)row0, (PetscInt)(row0 + nrows), (PetscInt)row0, (PetscInt)(row0 + nrows), &a));
  int64_t glo = 0, ghi = 0;
  int rc = b2k_csr_laplacian(ctx, dim, nx, ny, nz, row0, nrows, &a->A, &glo, &ghi);
  if (rc) { MatDestroy(&A); SETERRQ(PETSC_ERR_GPU, "b2k_csr_laplacian failed (%d): %s", rc, b2k_last_error()); }
  B2KCall(b2k_csr_info(a->A, NULL, NULL, NULL, &a->nnz));
  a->nghost = (PetscInt)(glo + ghi);
  if (a->nghost) {
    B2KCall(b2k_malloc(ctx, (void **)&a->xghost, sizeof(double) * (size_t)a->nghost));
    /* ghost order [lower plane | upper plane]; the planes sent are contiguous pieces of x: no packing */
    a->recvrank = (PetscInt *)malloc(2 * sizeof(PetscInt)); a->recvcount = (PetscInt *)malloc(2 * sizeof(PetscInt));
    a->sendrank = (PetscInt *)malloc(2 * sizeof(PetscInt)); a->sendcount = (PetscInt *)malloc(2 * sizeof(PetscInt));
    a->sendoff = (PetscInt *)malloc(2 * sizeof(PetscInt));
    PetscCheck(a->recvrank && a->recvcount && a->sendrank && a->sendcount && a->sendoff, PETSC_ERR_MEM, "out of memory");
    if (glo) {
      a->recvrank[a->nrecv] = rank - 1; a->recvcount[a->nrecv++] = (PetscInt)plane;
      a->sendrank[a->nsend] = rank - 1; a->sendcount[a->nsend] = (PetscInt)plane; a->sendoff[a->nsend++] = 0;
    }
    if (ghi) {
      a->recvrank[a->nrecv] = rank + 1; a->recvcount[a->nrecv++] = (PetscInt)plane;
      a->sendrank[a->nsend] = rank + 1; a->sendcount[a->nsend] = (PetscInt)plane; a->sendoff[a->nsend++] = (PetscInt)(nrows - plane);
    }
    for (PetscInt q = 0; q < a->nsend; q++) a->nsendtot += a->sendcount[q];
    a->halo_set = PETSC_TRUE;
  }
  if (size > 1) PetscCall(MatHaloSetUpP2P_B200CSR(A));   /* collective, also on a rank without neighbours */
  *out = A;
  return PETSC_SUCCESS;
}

/* explicit transpose (single rank): what SVDSetUp builds by default, svdsetup.c:300-306 — on the device, nothing crosses PCIe */
PetscErrorCode MatB200CSRTranspose(Mat A, Mat *At)
{
  PetscCheck(!strcmp(A->type, "b200csr"), PETSC_ERR_ARG_WRONG, "not a b200csr matrix");
  Mat_B200CSR *a = (Mat_B200CSR *)A->data, *t;
  b2k_ctx ctx = CTX();
  PetscCheck(a->nghost == 0 && A->m == A->M && A->n == A->N, PETSC_ERR_SUP, "explicit transpose is implemented for a single rank; pass A^T with SVDSetTransposeMatrix() otherwise");
  Mat T;
  PetscCall(MatCreate_Private(&T));
  PetscCall(MatSetUp_B200CSR(T, A->N, A->M, 0, A->N, 0, A->M, &t));
  t->nnz = a->nnz;
  const int rc = b2k_csr_transpose_split(ctx, a->A, &t->A, NULL);
  if (rc) { MatDestroy(&T); SETERRQ(PETSC_ERR_GPU, "b2k_csr_transpose_split failed (%d): %s", rc, b2k_last_error()); }
  *At = T;
  return PETSC_SUCCESS;
}
