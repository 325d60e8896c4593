/*
 * sys.c — the small slice of PETSc/SLEPc "sys" the Krylov path needs: error reporting
 * (PetscCall/SETERRQ behaviour), the device context, the row-partition communicator
 * (MPI_Comm stand-in), Vec, sequential dense Mat, MatShell, eigenvalue comparators
 * (src/sys/slepcsc.c:152-260) and the deterministic hash used instead of PetscRandom.
 */
#include <stdarg.h>
#include "b2kimpl.h"

/* ---------------------------------------------------------------------------------------------- */
static __thread char g_msg[2048];
static __thread size_t g_len;

const char *B2KGetLastErrorMessage(void) { return g_msg; }
void B2KClearError(void) { g_msg[0] = 0; g_len = 0; }

PetscErrorCode B2KErrorSet(PetscErrorCode code, const char *func, const char *file, int line, const char *fmt, ...)
{
  va_list ap;
  g_len = 0;
  va_start(ap, fmt);
  int w = vsnprintf(g_msg, sizeof(g_msg), fmt, ap);
  va_end(ap);
  if (w < 0) w = 0;
  g_len = (size_t)w < sizeof(g_msg) ? (size_t)w : sizeof(g_msg) - 1;
  return B2KErrorTrace(code, func, file, line);
}

PetscErrorCode B2KErrorTrace(PetscErrorCode code, const char *func, const char *file, int line)
{
  if (g_len < sizeof(g_msg) - 1) {
    const char *base = strrchr(file, '/');
    int w = snprintf(g_msg + g_len, sizeof(g_msg) - g_len, "\n  #%d %s() at %s:%d", code, func, base ? base + 1 : file, line);
    if (w > 0) g_len += (size_t)w < sizeof(g_msg) - g_len ? (size_t)w : sizeof(g_msg) - g_len - 1;
  }
  return code;
}

/* ---------------------------------------------------------------------------------------------- */
static b2k_ctx g_ctx = NULL;
static struct _p_B2KComm g_world = {0, 1, 0, NULL, NULL, NULL, NULL};

PetscErrorCode B2KInitialize(int device)
{
  if (g_ctx) return PETSC_SUCCESS;
  B2KCall(b2k_ctx_create(device, &g_ctx));
  return PETSC_SUCCESS;
}

PetscErrorCode B2KFinalize(void)
{
  PetscCall(B2KCommReset());
  if (g_ctx) { b2k_ctx_destroy(g_ctx); g_ctx = NULL; }
  return PETSC_SUCCESS;
}

b2k_ctx B2KGetContext(void) { return g_ctx; }

PetscErrorCode B2KDeviceSynchronize(void)
{
  if (g_ctx) B2KCall(b2k_ctx_sync(g_ctx));
  return PETSC_SUCCESS;
}

B2KComm B2KCommWorld(void) { return &g_world; }

PetscErrorCode B2KCommReset(void)
{
  if (g_world.kind == 1 && g_world.nccl) b2k_comm_destroy(g_world.nccl);
  memset(&g_world, 0, sizeof(g_world));
  g_world.size = 1;
  return PETSC_SUCCESS;
}

PetscErrorCode B2KCommInitNCCL(int rank, int size, const void *id)
{
  PetscCheck(g_ctx, PETSC_ERR_ORDER, "B2KInitialize() must be called before B2KCommInitNCCL()");
  PetscCall(B2KCommReset());
  B2KCall(b2k_comm_create(g_ctx, rank, size, id, &g_world.nccl));
  g_world.rank = rank; g_world.size = size; g_world.kind = 1;
  return PETSC_SUCCESS;
}

PetscErrorCode B2KCommP2PHandle(void *handle_out)
{
  PetscCheck(g_world.kind == 1 && g_world.nccl, PETSC_ERR_ORDER, "B2KCommInitNCCL() must be called before B2KCommP2PHandle()");
  B2KCall(b2k_comm_p2p_handle(g_world.nccl, handle_out));
  return PETSC_SUCCESS;
}

PetscErrorCode B2KCommP2POpen(const void *all_handles)
{
  PetscCheck(g_world.kind == 1 && g_world.nccl, PETSC_ERR_ORDER, "B2KCommInitNCCL() must be called before B2KCommP2POpen()");
  B2KCall(b2k_comm_p2p_open(g_world.nccl, all_handles));
  return PETSC_SUCCESS;
}

PetscErrorCode B2KCommDisableP2P(void)
{
  if (g_world.kind == 1 && g_world.nccl) B2KCall(b2k_comm_p2p_close(g_world.nccl));
  return PETSC_SUCCESS;
}

PetscErrorCode B2KCommEnableP2P(void)
{
  PetscCheck(g_world.kind == 1 && g_world.nccl, PETSC_ERR_ORDER, "B2KCommInitNCCL() and B2KCommP2POpen() come first");
  B2KCall(b2k_comm_p2p_resume(g_world.nccl));
  return PETSC_SUCCESS;
}

PetscErrorCode B2KCommReduceScope(B2KComm comm, PetscBool global, PetscBool *fused)
{
  int f = 0;
  if (comm && comm->kind == 1 && comm->size > 1 && comm->nccl) B2KCall(b2k_comm_reduce_scope(comm->nccl, global ? 1 : 0, &f));
  if (fused) *fused = f ? PETSC_TRUE : PETSC_FALSE;
  return PETSC_SUCCESS;
}

PetscErrorCode B2KCommInitCallbacks(int rank, int size, B2KAllreduceFn ar, B2KSendrecvFn sr, void *user)
{
  PetscCall(B2KCommReset());
  g_world.rank = rank; g_world.size = size; g_world.kind = 2;
  g_world.ar = ar; g_world.sr = sr; g_world.user = user;
  return PETSC_SUCCESS;
}

PetscErrorCode B2KCommGetRank(B2KComm comm, int *rank, int *size)
{
  if (rank) *rank = comm ? comm->rank : 0;
  if (size) *size = comm ? comm->size : 1;
  return PETSC_SUCCESS;
}

PetscErrorCode B2KCommAllreduce(B2KComm comm, double *buf, int count, int op, B2KMemType where)
{
  if (!comm || comm->size == 1 || count == 0) return PETSC_SUCCESS;
  if (comm->kind == 1) {
    PetscCheck(where == B2K_MEM_DEVICE, PETSC_ERR_ARG_INCOMP, "the NCCL communicator reduces device buffers only");
    if (op == 0) B2KCall(b2k_comm_allreduce_sum(comm->nccl, buf, count));
    else B2KCall(b2k_comm_allreduce_max(comm->nccl, buf, count));
  } else if (comm->kind == 2) {
    PetscCheck(where == B2K_MEM_HOST, PETSC_ERR_ARG_INCOMP, "the callback communicator reduces host buffers only");
    PetscCheck(!comm->ar(buf, count, op, comm->user), PETSC_ERR_LIB, "allreduce callback failed");
  } else SETERRQ(PETSC_ERR_PLIB, "communicator of size %d has no transport", comm->size);
  return PETSC_SUCCESS;
}

PetscErrorCode B2KCommSendrecv(B2KComm comm, const double *sbuf, PetscInt nsend, int dest, double *rbuf, PetscInt nrecv, int src,
                               B2KMemType where)
{
  if (!comm || comm->size == 1) return PETSC_SUCCESS;
  if (comm->kind == 1) {
    PetscCheck(where == B2K_MEM_DEVICE, PETSC_ERR_ARG_INCOMP, "the NCCL communicator moves device buffers only");
    B2KCall(b2k_comm_sendrecv(comm->nccl, sbuf, nsend, dest, rbuf, nrecv, src));
  } else if (comm->kind == 2) {
    PetscCheck(where == B2K_MEM_HOST, PETSC_ERR_ARG_INCOMP, "the callback communicator moves host buffers only");
    PetscCheck(!comm->sr(sbuf, nsend, dest, rbuf, nrecv, src, comm->user), PETSC_ERR_LIB, "sendrecv callback failed");
  } else SETERRQ(PETSC_ERR_PLIB, "communicator of size %d has no transport", comm->size);
  return PETSC_SUCCESS;
}

/* ---------------------------------------------------------------------------------------------- */
uint64_t B2KHashMix(uint64_t idx, uint64_t seed)
{
  uint64_t x = (idx + 1ull) * 0x9E3779B97F4A7C15ull + seed * 0xBF58476D1CE4E5B9ull;
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}
double B2KHashUniform(uint64_t idx, uint64_t seed)
{
  return 2.0 * ((double)(B2KHashMix(idx, seed) >> 11) * (1.0 / 9007199254740992.0)) - 1.0;
}

/* ---- comparators, slepcsc.c:152-260: result > 0 iff the second value is "better" --------------- */
#define CMP_BODY(a, b, gt)                 \
  if (gt) { *r = ((a) < (b)) ? 1 : (((a) > (b)) ? -1 : 0); } else { *r = ((a) > (b)) ? 1 : (((a) < (b)) ? -1 : 0); } \
  return PETSC_SUCCESS
PetscErrorCode SlepcCompareLargestMagnitude(PetscScalar ar, PetscScalar ai, PetscScalar br, PetscScalar bi, PetscInt *r, void *ctx)
{ (void)ctx; PetscReal a = SlepcAbsEigenvalue(ar, ai), b = SlepcAbsEigenvalue(br, bi); CMP_BODY(a, b, 1); }
PetscErrorCode SlepcCompareSmallestMagnitude(PetscScalar ar, PetscScalar ai, PetscScalar br, PetscScalar bi, PetscInt *r, void *ctx)
{ (void)ctx; PetscReal a = SlepcAbsEigenvalue(ar, ai), b = SlepcAbsEigenvalue(br, bi); CMP_BODY(a, b, 0); }
PetscErrorCode SlepcCompareLargestReal(PetscScalar ar, PetscScalar ai, PetscScalar br, PetscScalar bi, PetscInt *r, void *ctx)
{ (void)ctx; (void)ai; (void)bi; CMP_BODY(ar, br, 1); }
PetscErrorCode SlepcCompareSmallestReal(PetscScalar ar, PetscScalar ai, PetscScalar br, PetscScalar bi, PetscInt *r, void *ctx)
{ (void)ctx; (void)ai; (void)bi; CMP_BODY(ar, br, 0); }
PetscErrorCode SlepcCompareLargestImaginary(PetscScalar ar, PetscScalar ai, PetscScalar br, PetscScalar bi, PetscInt *r, void *ctx)
{ (void)ctx; (void)ar; (void)br; PetscReal a = fabs(ai), b = fabs(bi); CMP_BODY(a, b, 1); }
PetscErrorCode SlepcCompareSmallestImaginary(PetscScalar ar, PetscScalar ai, PetscScalar br, PetscScalar bi, PetscInt *r, void *ctx)
{ (void)ctx; (void)ar; (void)br; PetscReal a = fabs(ai), b = fabs(bi); CMP_BODY(a, b, 0); }
PetscErrorCode SlepcCompareTargetMagnitude(PetscScalar ar, PetscScalar ai, PetscScalar br, PetscScalar bi, PetscInt *r, void *ctx)
{ PetscScalar t = *(PetscScalar *)ctx; PetscReal a = SlepcAbsEigenvalue(ar - t, ai), b = SlepcAbsEigenvalue(br - t, bi); CMP_BODY(a, b, 0); }
PetscErrorCode SlepcCompareTargetReal(PetscScalar ar, PetscScalar ai, PetscScalar br, PetscScalar bi, PetscInt *r, void *ctx)
{ (void)ai; (void)bi; PetscScalar t = *(PetscScalar *)ctx; PetscReal a = fabs(ar - t), b = fabs(br - t); CMP_BODY(a, b, 0); }

/* ---- Vec ------------------------------------------------------------------------------------------ */
/* host vectors exist for plug-in BV types living in host memory (the test oracle); their arithmetic
   is supplied by that plug-in through B2KVecRegisterHostOps — this library ships device code only */
typedef struct {
  PetscErrorCode (*set)(Vec, PetscScalar);
  PetscErrorCode (*sumsq_local)(Vec, PetscReal *);
  PetscErrorCode (*dot_local)(Vec, Vec, PetscScalar *);
  PetscErrorCode (*axpy)(Vec, PetscScalar, Vec);
  PetscErrorCode (*scale)(Vec, PetscScalar);
  PetscErrorCode (*pointwisemult)(Vec, Vec, Vec);        /* w = x .* y */
} B2KVecHostOps;
static B2KVecHostOps g_hostvec;
static PetscBool g_hostvec_set = PETSC_FALSE;
PetscErrorCode B2KVecRegisterHostOps(const B2KVecHostOps *ops) { g_hostvec = *ops; g_hostvec_set = PETSC_TRUE; return PETSC_SUCCESS; }
#define NEED_HOSTVEC() PetscCheck(g_hostvec_set, PETSC_ERR_SUP, "host vectors need a plug-in that registers host Vec ops; libb2kslepc itself has no CPU arithmetic")
#define NEED_CTX() PetscCheck(g_ctx, PETSC_ERR_ORDER, "B2KInitialize() has not been called (no GPU context)")

PetscErrorCode VecCreateWithArray(B2KMemType mem, PetscInt n, PetscInt N, PetscScalar *array, Vec *v)
{
  Vec x = (Vec)calloc(1, sizeof(*x));
  if (x) x->hdr.classid = B2K_VEC_CLASSID;
  PetscCheck(x, PETSC_ERR_MEM, "out of memory");
  x->n = n; x->N = (N < 0) ? n : N; x->mem = mem; x->array = array; x->owns = PETSC_FALSE;
  *v = x;
  return PETSC_SUCCESS;
}

PetscErrorCode VecCreateB200(PetscInt n, PetscInt N, Vec *v)
{
  NEED_CTX();
  void *p = NULL;
  B2KCall(b2k_malloc(g_ctx, &p, sizeof(PetscScalar) * (size_t)(n + (n & 1))));
  B2KCall(b2k_memset0(g_ctx, p, sizeof(PetscScalar) * (size_t)(n + (n & 1))));
  PetscCall(VecCreateWithArray(B2K_MEM_DEVICE, n, N, (PetscScalar *)p, v));
  (*v)->owns = PETSC_TRUE;
  return PETSC_SUCCESS;
}

PetscErrorCode VecCreateHost(PetscInt n, PetscInt N, Vec *v)
{
  PetscScalar *p = (PetscScalar *)calloc((size_t)n + 2, sizeof(PetscScalar));
  PetscCheck(p, PETSC_ERR_MEM, "out of memory");
  PetscCall(VecCreateWithArray(B2K_MEM_HOST, n, N, p, v));
  (*v)->owns = PETSC_TRUE;
  return PETSC_SUCCESS;
}

PetscErrorCode VecDestroy(Vec *v)
{
  if (!v || !*v) return PETSC_SUCCESS;
  Vec x = *v;
  if (x->owns) {
    PetscScalar *p = x->saved ? x->saved : x->array;
    if (x->mem == B2K_MEM_DEVICE) { if (g_ctx) b2k_free(g_ctx, p); }
    else free(p);
  }
  free(x);
  *v = NULL;
  return PETSC_SUCCESS;
}

PetscErrorCode VecGetSize(Vec v, PetscInt *N) { *N = v->N; return PETSC_SUCCESS; }
PetscErrorCode VecGetLocalSize(Vec v, PetscInt *n) { *n = v->n; return PETSC_SUCCESS; }
PetscErrorCode VecGetArray(Vec v, PetscScalar **a) { *a = v->array; return PETSC_SUCCESS; }
PetscErrorCode VecGetMemType(Vec v, B2KMemType *mem) { *mem = v->mem; return PETSC_SUCCESS; }
PetscErrorCode VecPlaceArray(Vec v, PetscScalar *a)
{
  PetscCheck(!v->saved, PETSC_ERR_ARG_WRONGSTATE, "VecPlaceArray() already called without VecResetArray()");
  v->saved = v->array ? v->array : (PetscScalar *)(uintptr_t)1;
  v->array = a;
  return PETSC_SUCCESS;
}
PetscErrorCode VecResetArray(Vec v)
{
  v->array = (v->saved == (PetscScalar *)(uintptr_t)1) ? NULL : v->saved;
  v->saved = NULL;
  return PETSC_SUCCESS;
}

PetscErrorCode VecSetValuesHost(Vec v, const PetscScalar *hv)
{
  if (v->mem == B2K_MEM_DEVICE) { NEED_CTX(); B2KCall(b2k_h2d(g_ctx, v->array, hv, sizeof(PetscScalar) * (size_t)v->n)); }
  else memcpy(v->array, hv, sizeof(PetscScalar) * (size_t)v->n);
  return PETSC_SUCCESS;
}
PetscErrorCode VecGetValuesHost(Vec v, PetscScalar *hv)
{
  if (v->mem == B2K_MEM_DEVICE) { NEED_CTX(); B2KCall(b2k_d2h(g_ctx, hv, v->array, sizeof(PetscScalar) * (size_t)v->n)); }
  else memcpy(hv, v->array, sizeof(PetscScalar) * (size_t)v->n);
  return PETSC_SUCCESS;
}

PetscErrorCode VecSet(Vec v, PetscScalar alpha)
{
  if (v->mem == B2K_MEM_DEVICE) { NEED_CTX(); B2KCall(b2k_fill(g_ctx, v->array, v->n, alpha)); }
  else { NEED_HOSTVEC(); PetscCall(g_hostvec.set(v, alpha)); }
  return PETSC_SUCCESS;
}

static PetscErrorCode VecSumSq_Global(Vec v, PetscReal *ss)
{
  double s = 0.0;
  if (v->mem == B2K_MEM_DEVICE) {
    NEED_CTX();
    double *d = NULL;
    B2KCall(b2k_malloc(g_ctx, (void **)&d, 16));
    B2KCall(b2k_sumsq(g_ctx, v->array, v->n > 0 ? v->n : 1, v->n, 1, d));
    PetscCall(B2KCommAllreduce(B2KCommWorld(), d, 1, 0, B2K_MEM_DEVICE));
    B2KCall(b2k_d2h(g_ctx, &s, d, 8));
    B2KCall(b2k_free(g_ctx, d));
  } else {
    NEED_HOSTVEC();
    PetscCall(g_hostvec.sumsq_local(v, &s));
    PetscCall(B2KCommAllreduce(B2KCommWorld(), &s, 1, 0, B2K_MEM_HOST));
  }
  *ss = s;
  return PETSC_SUCCESS;
}

PetscErrorCode VecNorm(Vec v, NormType type, PetscReal *val)
{
  PetscCheck(type == NORM_2 || type == NORM_FROBENIUS, PETSC_ERR_SUP, "only the 2-norm is implemented for Vec");
  PetscReal s = 0.0;
  PetscCall(VecSumSq_Global(v, &s));
  *val = sqrt(s);
  return PETSC_SUCCESS;
}

PetscErrorCode VecDot(Vec x, Vec y, PetscScalar *val)
{
  PetscCheck(x->mem == y->mem && x->n == y->n, PETSC_ERR_ARG_INCOMP, "incompatible vectors");
  double s = 0.0;
  if (x->mem == B2K_MEM_DEVICE) {
    NEED_CTX();
    double *d = NULL;
    B2KCall(b2k_malloc(g_ctx, (void **)&d, 16));
    B2KCall(b2k_dotvec(g_ctx, x->array, x->n > 0 ? x->n + (x->n & 1) : 2, x->n, 1, y->array, d));
    PetscCall(B2KCommAllreduce(B2KCommWorld(), d, 1, 0, B2K_MEM_DEVICE));
    B2KCall(b2k_d2h(g_ctx, &s, d, 8));
    B2KCall(b2k_free(g_ctx, d));
  } else {
    NEED_HOSTVEC();
    PetscCall(g_hostvec.dot_local(x, y, &s));
    PetscCall(B2KCommAllreduce(B2KCommWorld(), &s, 1, 0, B2K_MEM_HOST));
  }
  *val = s;
  return PETSC_SUCCESS;
}

/* global index of the first local entry (what PetscLayout gives a real Vec): template vectors of row-partitioned objects */
PetscErrorCode B2KVecSetOwnershipStart(Vec v, PetscInt rstart)
{
  PetscCheck(v, PETSC_ERR_ARG_NULL, "null vector");
  PetscCheck(rstart >= 0 && rstart + v->n <= v->N, PETSC_ERR_ARG_OUTOFRANGE, "rows [%d,%d) outside the global length %d", rstart, rstart + v->n, v->N);
  v->rstart = rstart;
  return PETSC_SUCCESS;
}

PetscErrorCode VecAXPY(Vec y, PetscScalar alpha, Vec x)
{
  PetscCheck(x->mem == y->mem && x->n == y->n, PETSC_ERR_ARG_INCOMP, "incompatible vectors");
  if (y->mem == B2K_MEM_DEVICE) { NEED_CTX(); B2KCall(b2k_axpby(g_ctx, y->array, y->n, x->array, x->n, y->n, 1, alpha, 1.0)); }
  else { NEED_HOSTVEC(); PetscCall(g_hostvec.axpy(y, alpha, x)); }
  return PETSC_SUCCESS;
}

PetscErrorCode VecScale(Vec v, PetscScalar alpha)
{
  if (v->mem == B2K_MEM_DEVICE) { NEED_CTX(); B2KCall(b2k_scale(g_ctx, v->array, v->n, v->n, 1, alpha)); }
  else { NEED_HOSTVEC(); PetscCall(g_hostvec.scale(v, alpha)); }
  return PETSC_SUCCESS;
}

PetscErrorCode VecAYPX(Vec y, PetscScalar beta, Vec x)
{
  PetscCheck(x->mem == y->mem && x->n == y->n, PETSC_ERR_ARG_INCOMP, "incompatible vectors");
  if (y->mem == B2K_MEM_DEVICE) { NEED_CTX(); B2KCall(b2k_axpby(g_ctx, y->array, y->n, x->array, x->n, y->n, 1, 1.0, beta)); }
  else { PetscCall(VecScale(y, beta)); PetscCall(VecAXPY(y, 1.0, x)); }
  return PETSC_SUCCESS;
}

PetscErrorCode VecPointwiseMult(Vec w, Vec x, Vec y)
{
  PetscCheck(w->mem == x->mem && x->mem == y->mem && w->n == x->n && x->n == y->n, PETSC_ERR_ARG_INCOMP, "incompatible vectors");
  if (w->mem == B2K_MEM_DEVICE) { NEED_CTX(); B2KCall(b2k_pointwise_mult(g_ctx, w->array, x->array, y->array, w->n)); }
  else { NEED_HOSTVEC(); PetscCheck(g_hostvec.pointwisemult, PETSC_ERR_SUP, "the host Vec plug-in has no pointwise product"); PetscCall(g_hostvec.pointwisemult(w, x, y)); }
  return PETSC_SUCCESS;
}

/* set-up only (once per linear system): through the host */
PetscErrorCode VecReciprocal(Vec v)
{
  PetscScalar *h = (PetscScalar *)malloc(sizeof(PetscScalar) * (size_t)(v->n > 0 ? v->n : 1));
  PetscCheck(h, PETSC_ERR_MEM, "out of memory");
  PetscErrorCode ierr = VecGetValuesHost(v, h);
  if (!ierr) { for (PetscInt i = 0; i < v->n; i++) if (h[i] != 0.0) h[i] = 1.0 / h[i]; ierr = VecSetValuesHost(v, h); }
  free(h);
  PetscCall(ierr);
  return PETSC_SUCCESS;
}

PetscErrorCode VecCopy(Vec x, Vec y)
{
  PetscCheck(x->n == y->n, PETSC_ERR_ARG_INCOMP, "incompatible vectors");
  const size_t bytes = sizeof(PetscScalar) * (size_t)x->n;
  if (x->mem == B2K_MEM_DEVICE && y->mem == B2K_MEM_DEVICE) { NEED_CTX(); B2KCall(b2k_d2d(g_ctx, y->array, x->array, bytes)); }
  else if (x->mem == B2K_MEM_HOST && y->mem == B2K_MEM_HOST) memcpy(y->array, x->array, bytes);
  else if (x->mem == B2K_MEM_HOST) { NEED_CTX(); B2KCall(b2k_h2d(g_ctx, y->array, x->array, bytes)); }
  else { NEED_CTX(); B2KCall(b2k_d2h(g_ctx, y->array, x->array, bytes)); }
  return PETSC_SUCCESS;
}

/* ---- Mat ------------------------------------------------------------------------------------------ */
PetscErrorCode MatCreate_Private(Mat *A)
{
  Mat a = (Mat)calloc(1, sizeof(*a));
  if (a) a->hdr.classid = B2K_MAT_CLASSID;
  PetscCheck(a, PETSC_ERR_MEM, "out of memory");
  *A = a;
  return PETSC_SUCCESS;
}

PetscErrorCode MatCreateSeqDense(PetscInt m, PetscInt n, PetscScalar *data, Mat *A)
{
  PetscCheck(m >= 0 && n >= 0, PETSC_ERR_ARG_OUTOFRANGE, "negative dimension");
  Mat a;
  PetscCall(MatCreate_Private(&a));
  strcpy(a->type, "seqdense");
  a->m = a->M = m; a->n = a->N = n; a->rstart = 0; a->rend = m; a->cstart = 0; a->cend = n;
  a->lda = m > 0 ? m : 1; a->mem = B2K_MEM_HOST;
  if (data) { a->dense = data; a->owns_dense = PETSC_FALSE; }
  else {
    a->dense = (PetscScalar *)calloc((size_t)a->lda * (size_t)(n > 0 ? n : 1), sizeof(PetscScalar));
    PetscCheck(a->dense, PETSC_ERR_MEM, "out of memory");
    a->owns_dense = PETSC_TRUE;
  }
  *A = a;
  return PETSC_SUCCESS;
}

PetscErrorCode MatDenseSetLDA(Mat A, PetscInt lda)
{
  PetscCheck(A->dense, PETSC_ERR_ARG_WRONG, "not a dense matrix");
  PetscCheck(lda >= A->m, PETSC_ERR_ARG_SIZ, "LDA %d must be at least the number of rows %d", lda, A->m);
  if (A->owns_dense) {
    free(A->dense);
    A->dense = (PetscScalar *)calloc((size_t)lda * (size_t)(A->n > 0 ? A->n : 1), sizeof(PetscScalar));
    PetscCheck(A->dense, PETSC_ERR_MEM, "out of memory");
  }
  A->lda = lda;
  return PETSC_SUCCESS;
}
PetscErrorCode MatDenseGetLDA(Mat A, PetscInt *lda) { PetscCheck(A->dense, PETSC_ERR_ARG_WRONG, "not a dense matrix"); *lda = A->lda; return PETSC_SUCCESS; }
PetscErrorCode MatDenseGetArray(Mat A, PetscScalar **a) { PetscCheck(A->dense, PETSC_ERR_ARG_WRONG, "not a dense matrix"); *a = A->dense; return PETSC_SUCCESS; }
PetscErrorCode MatDenseRestoreArray(Mat A, PetscScalar **a) { (void)A; if (a) *a = NULL; return PETSC_SUCCESS; }
PetscErrorCode MatGetSize(Mat A, PetscInt *M, PetscInt *N) { if (M) *M = A->M; if (N) *N = A->N; return PETSC_SUCCESS; }
PetscErrorCode MatGetLocalSize(Mat A, PetscInt *m, PetscInt *n) { if (m) *m = A->m; if (n) *n = A->n; return PETSC_SUCCESS; }
PetscErrorCode MatGetOwnershipRange(Mat A, PetscInt *rs, PetscInt *re) { if (rs) *rs = A->rstart; if (re) *re = A->rend; return PETSC_SUCCESS; }
PetscErrorCode MatGetType(Mat A, const char **type) { *type = A->type; return PETSC_SUCCESS; }

PetscErrorCode MatDestroy(Mat *A)
{
  if (!A || !*A) return PETSC_SUCCESS;
  Mat a = *A;
  if (a->ops.destroy) PetscCall(a->ops.destroy(a));
  if (a->dense && a->owns_dense) free(a->dense);
  free(a);
  *A = NULL;
  return PETSC_SUCCESS;
}

PetscErrorCode MatMult(Mat A, Vec x, Vec y)
{
  PetscCheck(A->ops.mult, PETSC_ERR_SUP, "Mat type %s has no MatMult", A->type);
  PetscCheck(x != y, PETSC_ERR_ARG_WRONG, "x and y must be different vectors");
  PetscCheck(x->n == A->n && y->n == A->m, PETSC_ERR_ARG_SIZ, "Mat (%d x %d local) incompatible with x (%d) / y (%d)", A->m, A->n, x->n, y->n);
  PetscCall(A->ops.mult(A, x, y));
  return PETSC_SUCCESS;
}

PetscErrorCode MatGetDiagonal(Mat A, Vec d)
{
  PetscCheck(A->ops.getdiagonal, PETSC_ERR_SUP, "Mat type %s has no MatGetDiagonal", A->type);
  PetscCheck(d->n == A->m, PETSC_ERR_ARG_SIZ, "vector of %d entries for a matrix with %d local rows", d->n, A->m);
  PetscCall(A->ops.getdiagonal(A, d));
  return PETSC_SUCCESS;
}

PetscErrorCode MatMultTranspose(Mat A, Vec x, Vec y)
{
  PetscCheck(A->ops.multtranspose, PETSC_ERR_SUP, "Mat type %s has no MatMultTranspose", A->type);
  PetscCheck(x != y, PETSC_ERR_ARG_WRONG, "x and y must be different vectors");
  PetscCheck(x->n == A->m && y->n == A->n, PETSC_ERR_ARG_SIZ, "Mat^T incompatible with x / y");
  PetscCall(A->ops.multtranspose(A, x, y));
  return PETSC_SUCCESS;
}

/* virtual transpose (PETSc MatCreateHermitianTranspose, used by svdsetup.c:309-315 for the implicit transpose): MatMult of
   the result is MatMultTranspose of A and vice versa; A must outlive it */
static PetscErrorCode MatMult_Transpose(Mat T, Vec x, Vec y) { return MatMultTranspose((Mat)T->data, x, y); }
static PetscErrorCode MatMultTranspose_Transpose(Mat T, Vec x, Vec y) { return MatMult((Mat)T->data, x, y); }
PetscErrorCode MatCreateHermitianTranspose(Mat A, Mat *At)
{
  Mat t;
  PetscCheck(A->ops.multtranspose, PETSC_ERR_SUP, "Mat type %s has no MatMultTranspose", A->type);
  PetscCall(MatCreate_Private(&t));
  strcpy(t->type, "transpose");
  t->m = A->n; t->n = A->m; t->M = A->N; t->N = A->M;
  t->rstart = A->cstart; t->rend = A->cend; t->cstart = A->rstart; t->cend = A->rend;
  t->mem = A->mem; t->data = A;
  t->ops.mult = MatMult_Transpose;
  t->ops.multtranspose = MatMultTranspose_Transpose;
  *At = t;
  return PETSC_SUCCESS;
}

PetscErrorCode MatCreateVecs(Mat A, Vec *right, Vec *left)
{
  if (right) {
    if (A->mem == B2K_MEM_DEVICE) PetscCall(VecCreateB200(A->n, A->N, right)); else PetscCall(VecCreateHost(A->n, A->N, right));
    (*right)->rstart = A->cstart;
  }
  if (left) {
    if (A->mem == B2K_MEM_DEVICE) PetscCall(VecCreateB200(A->m, A->M, left)); else PetscCall(VecCreateHost(A->m, A->M, left));
    (*left)->rstart = A->rstart;
  }
  return PETSC_SUCCESS;
}

PetscErrorCode MatCreateShell(PetscInt m, PetscInt n, PetscInt M, PetscInt N, B2KMemType mem, void *ctx, Mat *A)
{
  Mat a;
  PetscCall(MatCreate_Private(&a));
  strcpy(a->type, "shell");
  a->m = m; a->n = n; a->M = (M < 0) ? m : M; a->N = (N < 0) ? n : N;
  a->rstart = 0; a->rend = m; a->cstart = 0; a->cend = n;
  a->mem = mem; a->data = ctx;
  *A = a;
  return PETSC_SUCCESS;
}
PetscErrorCode MatShellGetContext(Mat A, void **ctx) { *ctx = A->data; return PETSC_SUCCESS; }
PetscErrorCode MatShellSetMult(Mat A, MatMultFn f) { A->ops.mult = f; return PETSC_SUCCESS; }
PetscErrorCode MatShellSetMultTranspose(Mat A, MatMultFn f) { A->ops.multtranspose = f; return PETSC_SUCCESS; }
PetscErrorCode MatShellSetGetDiagonal(Mat A, MatGetDiagonalFn f) { A->ops.getdiagonal = f; return PETSC_SUCCESS; }

/* ---- names and ASCII viewers: PetscObjectSetName, PetscViewerASCIIGetStdout / PushFormat / PopFormat, VecView, MatView --------
   Only what the reference's BV test programs print (bv/tests/test1.c -verbose): the default ASCII format and INFO_DETAIL. */
PetscErrorCode PetscObjectSetName(PetscObject obj, const char name[])
{
  PetscCheck(obj && name, PETSC_ERR_ARG_NULL, "null argument");
  B2KObjectHeader *h = (B2KObjectHeader *)obj;
  PetscCheck(h->classid == B2K_VEC_CLASSID || h->classid == B2K_MAT_CLASSID || h->classid == B2K_BV_CLASSID, PETSC_ERR_ARG_WRONG,
             "PetscObjectSetName() is available for Vec, Mat and BV objects");
  strncpy(h->name, name, sizeof(h->name) - 1);
  h->name[sizeof(h->name) - 1] = 0;
  return PETSC_SUCCESS;
}

static struct _p_PetscViewer g_stdout_viewer = {{PETSC_VIEWER_DEFAULT}, 0};
PetscErrorCode PetscViewerASCIIGetStdout(PetscViewer *viewer) { *viewer = &g_stdout_viewer; return PETSC_SUCCESS; }
PetscErrorCode PetscViewerPushFormat(PetscViewer viewer, PetscViewerFormat format)
{
  if (!viewer) viewer = &g_stdout_viewer;
  PetscCheck(viewer->depth < 7, PETSC_ERR_PLIB, "Too many PetscViewerPushFormat(), perhaps you forgot PetscViewerPopFormat()?");
  viewer->fmt[++viewer->depth] = format;
  return PETSC_SUCCESS;
}
PetscErrorCode PetscViewerPopFormat(PetscViewer viewer)
{
  if (!viewer) viewer = &g_stdout_viewer;
  if (viewer->depth > 0) viewer->depth--;
  return PETSC_SUCCESS;
}
PetscErrorCode PetscViewerGetFormat(PetscViewer viewer, PetscViewerFormat *format)
{
  if (!viewer) viewer = &g_stdout_viewer;
  *format = viewer->fmt[viewer->depth];
  return PETSC_SUCCESS;
}

/* "%g" the way PETSc's ASCII output shows reals: a '.' is appended when the text has neither '.' nor an exponent
   (PetscFormatConvert/PetscVSNPrintf), so that -2 prints as "-2." */
void B2KFormatReal(double x, char buf[32])
{
  snprintf(buf, 30, "%g", x);
  if (!strpbrk(buf, ".eEn")) strcat(buf, ".");    /* 'n': inf / nan stay as they are */
}

/* first line of every view: "<Class> Object: [name ]<size> MPI process[es]" + "  type: <type>" */
PetscErrorCode B2KViewHeader_Private(const char *cls, const char *name, const char *type, int indent)
{
  int size = 1;
  PetscCall(B2KCommGetRank(B2KCommWorld(), NULL, &size));
  printf("%*s%s Object: %s%s%d MPI process%s\n", indent, "", cls, name, name[0] ? " " : "", size, size > 1 ? "es" : "");
  printf("%*s  type: %s\n", indent, "", type);
  return PETSC_SUCCESS;
}

PetscErrorCode VecView(Vec v, PetscViewer viewer)
{
  (void)viewer;
  int rank = 0, size = 1;
  PetscCall(B2KCommGetRank(B2KCommWorld(), &rank, &size));
  PetscCall(B2KViewHeader_Private("Vec", v->hdr.name, (size > 1 && v->N != v->n) ? "mpi" : "seq", 0));
  PetscScalar *h = (PetscScalar *)malloc(sizeof(PetscScalar) * (size_t)(v->n > 0 ? v->n : 1));
  PetscCheck(h, PETSC_ERR_MEM, "out of memory");
  PetscErrorCode ierr = VecGetValuesHost(v, h);
  char buf[32];
  for (PetscInt i = 0; !ierr && i < v->n; i++) { B2KFormatReal(h[i], buf); printf("%s\n", buf); }   /* this rank's entries */
  free(h);
  PetscCall(ierr);
  return PETSC_SUCCESS;
}

PetscErrorCode MatView(Mat A, PetscViewer viewer)
{
  (void)viewer;
  PetscCheck(A->dense, PETSC_ERR_SUP, "MatView is implemented for sequential dense matrices");
  PetscCall(B2KViewHeader_Private("Mat", A->hdr.name, "seqdense", 0));
  for (PetscInt i = 0; i < A->m; i++) {
    for (PetscInt j = 0; j < A->n; j++) printf("%.16e ", A->dense[i + (size_t)j * A->lda]);
    printf("\n");
  }
  return PETSC_SUCCESS;
}
