/*
 * b2k_comm.cu — row-partition communicator over NCCL / NVLink.
 * Replaces the MPI layer the reference uses on this path:
 *   MPIU_Allreduce of the k-vector per Gram-Schmidt pass   bvcuda.cu:228-248 (host-staged there),
 *   MPIU_LAPY2 norm reduction                              bvlapack.c:50,
 *   PETSc VecScatter halo exchange inside MatMult_MPIAIJ   (PETSc, reached from bvops.c:879).
 * One process per GPU; the process group rendezvous (exchange of the NCCL unique id) is done by
 * the caller, normally torch.distributed in slepc_b200/dist.py.  NCCL is resolved at run time with
 * dlopen so that the library also loads on a box without NCCL (the symbol test on CPU).
 */
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#include "b2k_internal.h"

/* minimal NCCL ABI (stable since 2.x): enough to avoid a build-time dependency on nccl.h */
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat64 = 8 };
enum { ncclSum = 0, ncclMax = 2 };

struct nccl_api {
  void *h;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *);
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*ReduceScatter)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)(void);
  ncclResult_t (*GroupEnd)(void);
  const char *(*GetErrorString)(ncclResult_t);
};
static nccl_api g_nccl;

static int load_nccl(void)
{
  if (g_nccl.h) return B2K_OK;
  const char *names[] = {"libnccl.so.2", "libnccl.so", NULL};
  void *h = NULL;
  const char *env = getenv("B2K_NCCL_LIB");
  if (env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
  for (int i = 0; !h && names[i]; i++) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!h) { b2k_set_error("cannot dlopen libnccl.so.2 (set B2K_NCCL_LIB): %s", dlerror()); return B2K_ERR_COMM; }
#define SYM(field, name)                                                               \
  do {                                                                                 \
    *(void **)(&g_nccl.field) = dlsym(h, name);                                        \
    if (!g_nccl.field) { b2k_set_error("NCCL symbol %s missing", name); return B2K_ERR_COMM; } \
  } while (0)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllReduce, "ncclAllReduce");
  SYM(AllGather, "ncclAllGather");
  SYM(ReduceScatter, "ncclReduceScatter");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  g_nccl.h = h;
  return B2K_OK;
}

#define NK(call)                                                                                  \
  do {                                                                                            \
    ncclResult_t r_ = (call);                                                                     \
    if (r_ != 0) {                                                                                \
      b2k_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_));     \
      return B2K_ERR_COMM;                                                                        \
    }                                                                                             \
  } while (0)

struct b2k_comm_s {
  b2k_ctx    ctx;
  int        rank, size;
  ncclComm_t nccl;
  /* peer-memory mailboxes of the one-shot reduction (b2k_comm_p2p_*) */
  double    *box_local;                 /* cudaMalloc'ed here, exported with cudaIpcGetMemHandle  */
  void      *box_peer[B2K_XG_MAXR];     /* cudaIpcOpenMemHandle mappings of the other ranks       */
  int       *err_host;                  /* pinned + mapped                                        */
  b2k_xg_s   xg;
  int        p2p_open;
};

extern "C" int b2k_comm_unique_id(void *id_host)
{
  int rc = load_nccl();
  if (rc) return rc;
  ncclUniqueId id;
  NK(g_nccl.GetUniqueId(&id));
  memcpy(id_host, &id, sizeof(id));
  return B2K_OK;
}

extern "C" int b2k_comm_create(b2k_ctx ctx, int rank, int size, const void *id_host, b2k_comm *out)
{
  ARGCHK(size >= 1 && rank >= 0 && rank < size, "bad rank/size");
  b2k_comm c = (b2k_comm)calloc(1, sizeof(*c));
  if (!c) return B2K_ERR_MEM;
  c->ctx = ctx; c->rank = rank; c->size = size; c->nccl = NULL;
  if (size > 1) {
    int rc = load_nccl();
    if (rc) return rc;
    ncclUniqueId id;
    memcpy(&id, id_host, sizeof(id));
    CK(cudaSetDevice(ctx->device));
    NK(g_nccl.CommInitRank(&c->nccl, size, id, rank));
  }
  *out = c;
  return B2K_OK;
}

/* ---- one-shot reductions over NVLink peer memory ------------------------------------------------------------------
 * Set-up (once per communicator): every rank allocates a mailbox, exports it with CUDA IPC, the launcher all-gathers the
 * 64-byte handles (like the NCCL id) and every rank maps the mailboxes of its peers.  From then on, while the reduce
 * scope is "global" (b2k_comm_reduce_scope), the second stage of every two-stage reduction of the BV kernels
 * (k_reduce_partials_xg, b2k_bv.cu) sums over the ranks itself and the ncclAllReduce that followed it is skipped. */
extern "C" int b2k_comm_p2p_handle(b2k_comm c, void *handle_out)
{
  ARGCHK(c && handle_out, "null argument");
  ARGCHK(c->size <= B2K_XG_MAXR, "peer-memory reductions support up to 8 ranks");
  CK(cudaSetDevice(c->ctx->device));
  if (!c->box_local) {
    CK(cudaMalloc(&c->box_local, B2K_XG_BYTES));
    CK(cudaMemset(c->box_local, 0, B2K_XG_BYTES));
    CK(cudaHostAlloc(&c->err_host, sizeof(int), cudaHostAllocMapped));
    *c->err_host = 0;
    CK(cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, c->box_local));
  static_assert(sizeof(h) == B2K_COMM_P2P_HANDLE_BYTES, "CUDA IPC handle size");
  memcpy(handle_out, &h, sizeof(h));
  return B2K_OK;
}

extern "C" int b2k_comm_p2p_open(b2k_comm c, const void *all_handles)
{
  ARGCHK(c && all_handles && c->box_local, "b2k_comm_p2p_handle() must be called first");
  CK(cudaSetDevice(c->ctx->device));
  c->xg.rank = c->rank; c->xg.size = c->size;
  for (int p = 0; p < c->size; p++) {
    if (p == c->rank) { c->xg.box[p] = c->box_local; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char *)all_handles + (size_t)p * sizeof(h), sizeof(h));
    CK(cudaIpcOpenMemHandle(&c->box_peer[p], h, cudaIpcMemLazyEnablePeerAccess));
    c->xg.box[p] = (double *)c->box_peer[p];
  }
  int *derr = NULL;
  CK(cudaHostGetDevicePointer((void **)&derr, c->err_host, 0));
  c->xg.err = derr;
  c->xg.spin_limit = b2k_spin_limit();
  c->p2p_open = 1;
  c->ctx->xg = &c->xg;
  c->ctx->xg_on = 0;
  c->ctx->xg_seq = 0;
  return B2K_OK;
}

extern "C" int b2k_comm_p2p_close(b2k_comm c)
{
  if (!c) return B2K_OK;
  c->p2p_open = 0;
  if (c->ctx->xg == &c->xg) { c->ctx->xg = NULL; c->ctx->xg_on = 0; }
  return B2K_OK;
}

/* after b2k_comm_p2p_close: route the reductions through the (still mapped) mailboxes again.  Collective: every rank toggles at
   the same point of the program (tests and bench compare the two transports on the same data). */
extern "C" int b2k_comm_p2p_resume(b2k_comm c)
{
  ARGCHK(c && c->box_local && c->xg.box[c->rank] == c->box_local, "b2k_comm_p2p_open() has not mapped the mailboxes");
  CK(cudaStreamSynchronize(c->ctx->stream));
  c->p2p_open = 1;
  c->ctx->xg = &c->xg;
  c->ctx->xg_on = 0;
  return B2K_OK;
}

extern "C" int b2k_comm_p2p_enabled(b2k_comm c) { return (c && c->p2p_open) ? 1 : 0; }

/* on != 0: reductions of the BV kernels launched from now on are sums over the ranks (collective: every rank must issue
   the same reductions in the same order); 0: local again.  Returns 1 through *fused when the peer-memory path is active
   (the caller then skips its all-reduce), 0 otherwise. */
extern "C" int b2k_comm_reduce_scope(b2k_comm c, int on, int *fused)
{
  const int act = (c && c->p2p_open && c->size > 1) ? 1 : 0;
  if (act) c->ctx->xg_on = on ? 1 : 0;
  if (fused) *fused = (act && on) ? 1 : 0;
  return B2K_OK;
}

/* non-zero once a peer failed to show up within the kernel's time-out (the sums of that reduction are then invalid) */
extern "C" int b2k_comm_p2p_error(b2k_comm c, int *flag)
{
  if (flag) *flag = (c && c->err_host) ? *(volatile int *)c->err_host : 0;
  return B2K_OK;
}

extern "C" int b2k_comm_destroy(b2k_comm c)
{
  if (!c) return B2K_OK;
  cudaStreamSynchronize(c->ctx->stream);
  if (c->ctx->xg == &c->xg) { c->ctx->xg = NULL; c->ctx->xg_on = 0; }
  for (int p = 0; p < B2K_XG_MAXR; p++) if (c->box_peer[p]) cudaIpcCloseMemHandle(c->box_peer[p]);
  if (c->p2p_open && c->nccl && c->size > 1) b2k_comm_barrier(c);   /* nobody maps my mailbox any more: safe to free it */
  if (c->box_local) cudaFree(c->box_local);
  if (c->err_host) cudaFreeHost(c->err_host);
  if (c->nccl) g_nccl.CommDestroy(c->nccl);
  free(c);
  return B2K_OK;
}

int b2k_comm_ctx(b2k_comm c, b2k_ctx *ctx) { if (!c) return B2K_ERR_ARG; *ctx = c->ctx; return B2K_OK; }   /* for b2k_halo.cu */

extern "C" int b2k_comm_rank(b2k_comm c, int *rank, int *size)
{
  if (rank) *rank = c ? c->rank : 0;
  if (size) *size = c ? c->size : 1;
  return B2K_OK;
}

int b2k_xg_allreduce_inplace(b2k_ctx ctx, double *buf, int n);   /* b2k_bv.cu */

extern "C" int b2k_comm_allreduce_sum(b2k_comm c, double *buf, int count)
{
  if (!c || c->size == 1 || count == 0) return B2K_OK;
  if (c->p2p_open) {                              /* short vectors: one-shot over the NVLink mailboxes, bit-identical on every rank */
    const int rc = b2k_xg_allreduce_inplace(c->ctx, buf, count);
    if (rc != -1) return rc;
  }
  NK(g_nccl.AllReduce(buf, buf, (size_t)count, ncclFloat64, ncclSum, c->nccl, c->ctx->stream));
  return B2K_OK;
}
extern "C" int b2k_comm_allreduce_max(b2k_comm c, double *buf, int count)
{
  if (!c || c->size == 1 || count == 0) return B2K_OK;
  NK(g_nccl.AllReduce(buf, buf, (size_t)count, ncclFloat64, ncclMax, c->nccl, c->ctx->stream));
  return B2K_OK;
}
extern "C" int b2k_comm_group_start(b2k_comm c) { if (c && c->size > 1) NK(g_nccl.GroupStart()); return B2K_OK; }
extern "C" int b2k_comm_group_end(b2k_comm c) { if (c && c->size > 1) NK(g_nccl.GroupEnd()); return B2K_OK; }

extern "C" int b2k_comm_sendrecv(b2k_comm c, const double *sendbuf, int64_t nsend, int send_peer, double *recvbuf, int64_t nrecv,
                                 int recv_peer)
{
  if (!c || c->size == 1) return B2K_OK;
  if (nsend > 0) NK(g_nccl.Send(sendbuf, (size_t)nsend, ncclFloat64, send_peer, c->nccl, c->ctx->stream));
  if (nrecv > 0) NK(g_nccl.Recv(recvbuf, (size_t)nrecv, ncclFloat64, recv_peer, c->nccl, c->ctx->stream));
  return B2K_OK;
}
extern "C" int b2k_comm_allgather(b2k_comm c, const double *sendbuf, double *recvbuf, int64_t cnt)
{
  if (!c || c->size == 1) {
    if (sendbuf != recvbuf && cnt) CK(cudaMemcpyAsync(recvbuf, sendbuf, sizeof(double) * (size_t)cnt, cudaMemcpyDeviceToDevice,
                                                      c ? c->ctx->stream : 0));
    return B2K_OK;
  }
  NK(g_nccl.AllGather(sendbuf, recvbuf, (size_t)cnt, ncclFloat64, c->nccl, c->ctx->stream));
  return B2K_OK;
}
extern "C" int b2k_comm_reduce_scatter_sum(b2k_comm c, const double *sendbuf, double *recvbuf, int64_t cnt)
{
  if (!c || c->size == 1) {
    if (sendbuf != recvbuf && cnt) CK(cudaMemcpyAsync(recvbuf, sendbuf, sizeof(double) * (size_t)cnt, cudaMemcpyDeviceToDevice,
                                                      c ? c->ctx->stream : 0));
    return B2K_OK;
  }
  NK(g_nccl.ReduceScatter(sendbuf, recvbuf, (size_t)cnt, ncclFloat64, ncclSum, c->nccl, c->ctx->stream));
  return B2K_OK;
}
extern "C" int b2k_comm_barrier(b2k_comm c)
{
  if (!c || c->size == 1) return B2K_OK;
  int rc = b2k_comm_allreduce_sum(c, c->ctx->dscratch + (c->ctx->dscratch_elems - 1), 1);
  if (rc) return rc;
  CK(cudaStreamSynchronize(c->ctx->stream));
  return B2K_OK;
}
