/*
 * b2k_halo.cu — SpMV halo exchange over NVLink peer memory (default on one box; B2K_HALO_P2P=0 keeps the grouped
 * ncclSend/ncclRecv of matb200.c).  The same object with the roles exchanged carries the REVERSE halo of MatMultTranspose
 * (ghost-column contributions pushed to their owners).  Replaces the VecScatter of PETSc's MatMult_MPIAIJ /
 * MatMultTranspose_MPIAIJ (reached from bvops.c:879, gklanczos.c:80,103) for the
 * GPUs of one NVSwitch box: every rank PUSHES the entries of x its neighbours need straight into the neighbours' ghost
 * buffers (peer-mapped HBM, CUDA IPC) and raises a sequence-numbered flag; the receiver waits for the flags of its senders
 * and multiplies.  No packing buffer, no NCCL launch, 2 small kernels per MatMult.
 *
 * Protocol (sequence number s = 1, 2, … per exchange; every rank issues the same exchanges in the same order):
 *   ghost data is double-buffered by the parity of s;
 *   push(s)  [k_halo_push, stream-ordered after the SpMV of s-1]:
 *        - acknowledges to every sender that everything before s has been consumed   (ack = s-1 in the sender's block)
 *        - per destination: waits until that destination has acknowledged s-2 (its buffer of this parity is free),
 *          copies the entries, fences, and the last CTA of the destination publishes arrived = s in the destination's block
 *   wait(s)  [k_halo_wait]: spins until arrived >= s for every sender, fences; the SpMV that follows reads buffer s&1.
 * Every wait is on an event with a strictly smaller position in the peers' streams, so the protocol cannot deadlock as long
 * as all ranks call the exchange collectively; a 60 s time-out raises the error flag instead of hanging.
 *
 * Set-up is collective and needs nothing from the launcher: the 64-byte IPC handles and the slot tables travel through
 * ncclAllGather on the communicator (b2k_comm_allgather).
 */
#include <stdlib.h>
#include <string.h>
#include "b2k_internal.h"

#define HL_MAXP 8                      /* peers per direction (= ranks of one box)                  */
#define HL_CTAS 16                     /* CTAs per destination in k_halo_push: at least this many …  */
#define HL_CTAS_MAX 64                 /* … and one more per 8192 entries up to this (C5: 1.25e6 entries = 10 MB per destination) */

typedef unsigned long long u64;

struct hl_dest {                       /* one destination of my pushes */
  double       *peer_data;             /* peer's ghost buffers [2][peer_nghost]                      */
  u64          *peer_arrived;          /* peer's arrived[my recv slot there]                         */
  const u64    *my_ack;                /* my ack[q]: last sequence this destination has consumed     */
  unsigned int *done;                  /* my CTA counter for this destination                        */
  const int    *idx;                   /* device: local indices to send, or NULL = contiguous        */
  long long     xoff, count, peer_off, peer_nghost;
};
struct hl_args {
  int      nsend, nrecv;
  hl_dest  d[HL_MAXP];
  u64     *peer_ack[HL_MAXP];          /* sender p's ack[my send slot there]                          */
  const u64 *my_arrived;               /* my arrived[0..nrecv)                                        */
  int     *err;
  long long spin_limit;
};

struct b2k_halo_s {
  b2k_comm comm;
  b2k_ctx  ctx;
  int      rank, size;
  void    *block;                      /* my allocation: data[2][nghost] | arrived[HL_MAXP] | ack[HL_MAXP] | done[HL_MAXP] */
  void    *peer[HL_MAXP];              /* mapped allocations of all ranks (NULL for me)               */
  long long nghost;
  int     *d_sendidx;                  /* not owned                                                   */
  int     *err_host;
  hl_args  args;
  u64      seq;
  int      ctas;                       /* CTAs per destination                                        */
};

/* the communicator internals this file needs (b2k_comm.cu) */
int b2k_comm_ctx(b2k_comm c, b2k_ctx *ctx);

__host__ __device__ static inline long long hl_stride(long long nghost) { return (nghost + 1) & ~1LL; }   /* doubles per parity buffer: even, so both are 16-byte aligned */
static inline size_t hl_data_bytes(long long nghost) { return ((size_t)(2 * hl_stride(nghost)) * sizeof(double) + 255) & ~(size_t)255; }
static inline size_t hl_block_bytes(long long nghost) { return hl_data_bytes(nghost) + 3 * HL_MAXP * sizeof(u64); }
__host__ __device__ static inline u64 *hl_arrived(void *block, size_t data_bytes) { return (u64 *)((char *)block + data_bytes); }
__host__ __device__ static inline u64 *hl_ack(void *block, size_t data_bytes) { return hl_arrived(block, data_bytes) + HL_MAXP; }

__global__ void __launch_bounds__(256) k_halo_push(const double *__restrict__ x, const hl_args a, u64 seq)
{
  b2k_pdl_enter();
  /* acknowledgements first: this kernel runs after the SpMV of seq-1, so its ghost buffer is free again */
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < a.nrecv) {
    __threadfence_system();
    *((volatile u64 *)a.peer_ack[threadIdx.x]) = seq - 1;
  }
  if ((int)blockIdx.y >= a.nsend) return;
  const hl_dest d = a.d[blockIdx.y];
  __shared__ int go;
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    go = 1;
    while (*((volatile const u64 *)d.my_ack) + 2 < seq) {
      if (clock64() - t0 > a.spin_limit) { *((volatile int *)a.err) = 2; go = 0; break; }
    }
    __threadfence_system();
  }
  __syncthreads();
  if (go) {
    double *dst = d.peer_data + (seq & 1ull) * hl_stride(d.peer_nghost) + d.peer_off;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (d.idx) for (long long i = t0; i < d.count; i += stride) dst[i] = x[d.idx[i]];
    else if ((((unsigned long long)(x + d.xoff) | (unsigned long long)dst) & 15ull) == 0) {      /* contiguous plane, aligned: 16-byte stores over NVLink */
      const double2 *s2 = reinterpret_cast<const double2 *>(x + d.xoff);
      double2 *d2 = reinterpret_cast<double2 *>(dst);
      const long long n2 = d.count >> 1;
      for (long long i = t0; i < n2; i += stride) d2[i] = s2[i];
      if ((d.count & 1) && t0 == 0) dst[d.count - 1] = x[d.xoff + d.count - 1];
    } else for (long long i = t0; i < d.count; i += stride) dst[i] = x[d.xoff + i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(d.done, 1u) == gridDim.x - 1) {       /* last CTA of this destination: everything is on its way */
      *d.done = 0;
      __threadfence_system();
      *((volatile u64 *)d.peer_arrived) = seq;
    }
  }
}

__global__ void k_halo_wait(const hl_args a, u64 seq)
{
  b2k_pdl_enter();
  if (threadIdx.x < a.nrecv) {
    const long long t0 = clock64();
    while (*((volatile const u64 *)(a.my_arrived + threadIdx.x)) < seq) {
      if (clock64() - t0 > a.spin_limit) { *((volatile int *)a.err) = 3; break; }
    }
    __threadfence_system();
  }
}

/* bit-copy all-gather of `bytes` (multiple of 8) per rank through the communicator */
static int hl_allgather_bytes(b2k_halo h, const void *mine, void *all, size_t bytes)
{
  b2k_ctx ctx = h->ctx;
  double *dsend = NULL, *drecv = NULL;
  CK(cudaMalloc(&dsend, bytes));
  CK(cudaMalloc(&drecv, bytes * (size_t)h->size));
  CK(cudaMemcpyAsync(dsend, mine, bytes, cudaMemcpyHostToDevice, ctx->stream));
  int rc = b2k_comm_allgather(h->comm, dsend, drecv, (int64_t)(bytes / 8));
  if (!rc) {
    CK(cudaMemcpyAsync(all, drecv, bytes * (size_t)h->size, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  cudaFree(dsend); cudaFree(drecv);
  return rc;
}

extern "C" int b2k_halo_create(b2k_comm comm, int nrecv, const int *recvrank, const int *recvcount, int nsend, const int *sendrank,
                               const int *sendcount, const int *sendidx_dev, const int64_t *sendoff_host, b2k_halo *out)
{
  ARGCHK(comm && out, "null argument");
  ARGCHK(nrecv >= 0 && nrecv <= HL_MAXP && nsend >= 0 && nsend <= HL_MAXP, "at most 8 peers per direction");
  ARGCHK(sendidx_dev || sendoff_host || nsend == 0, "either the send indices (device) or the contiguous offsets (host) are needed");
  b2k_halo h = (b2k_halo)calloc(1, sizeof(*h));
  if (!h) return B2K_ERR_MEM;
  h->comm = comm;
  int rc = b2k_comm_ctx(comm, &h->ctx);
  if (rc) return rc;
  b2k_comm_rank(comm, &h->rank, &h->size);
  ARGCHK(h->size >= 2 && h->size <= HL_MAXP, "the peer-memory halo needs 2..8 ranks");
  CK(cudaSetDevice(h->ctx->device));
  for (int p = 0; p < nrecv; p++) h->nghost += recvcount[p];
  const size_t dbytes = hl_data_bytes(h->nghost), bbytes = hl_block_bytes(h->nghost);
  CK(cudaMalloc(&h->block, bbytes));
  CK(cudaMemset(h->block, 0, bbytes));
  CK(cudaHostAlloc(&h->err_host, sizeof(int), cudaHostAllocMapped));
  *h->err_host = 0;
  CK(cudaDeviceSynchronize());

  /* what every rank publishes: its IPC handle (64 B), nghost, and for every rank r: the recv slot / offset of r here and the
     send slot of r here (-1 when r is not a peer in that direction) */
  const int S = h->size;
  const size_t words = 8 + 1 + 3 * (size_t)S;
  double *mine = (double *)calloc(words, sizeof(double)), *all = (double *)calloc(words * (size_t)S, sizeof(double));
  if (!mine || !all) return B2K_ERR_MEM;
  cudaIpcMemHandle_t hd;
  CK(cudaIpcGetMemHandle(&hd, h->block));
  memcpy(mine, &hd, sizeof(hd));
  mine[8] = (double)h->nghost;
  for (int r = 0; r < S; r++) { mine[9 + 3 * r] = -1.0; mine[10 + 3 * r] = 0.0; mine[11 + 3 * r] = -1.0; }
  {
    long long off = 0;
    for (int p = 0; p < nrecv; p++) {
      ARGCHK(recvrank[p] >= 0 && recvrank[p] < S && recvrank[p] != h->rank, "bad receive rank");
      mine[9 + 3 * recvrank[p]] = (double)p; mine[10 + 3 * recvrank[p]] = (double)off;
      off += recvcount[p];
    }
    for (int q = 0; q < nsend; q++) {
      ARGCHK(sendrank[q] >= 0 && sendrank[q] < S && sendrank[q] != h->rank, "bad send rank");
      mine[11 + 3 * sendrank[q]] = (double)q;
    }
  }
  rc = hl_allgather_bytes(h, mine, all, words * sizeof(double));
  if (rc) return rc;
  for (int r = 0; r < S; r++) {
    if (r == h->rank) continue;
    cudaIpcMemHandle_t hr;
    memcpy(&hr, all + words * (size_t)r, sizeof(hr));
    CK(cudaIpcOpenMemHandle(&h->peer[r], hr, cudaIpcMemLazyEnablePeerAccess));
  }
  /* kernel arguments */
  hl_args *a = &h->args;
  a->nsend = nsend; a->nrecv = nrecv;
  int *derr = NULL;
  CK(cudaHostGetDevicePointer((void **)&derr, h->err_host, 0));
  a->err = derr;
  a->spin_limit = b2k_spin_limit();
  a->my_arrived = hl_arrived(h->block, dbytes);
  unsigned int *done = (unsigned int *)(hl_ack(h->block, dbytes) + HL_MAXP);
  long long soff = 0;
  for (int q = 0; q < nsend; q++) {
    const int r = sendrank[q];
    const double *pr = all + words * (size_t)r;
    const long long peer_nghost = (long long)pr[8];
    const int slot_there = (int)pr[9 + 3 * h->rank];
    ARGCHK(slot_there >= 0, "halo plans do not match: a destination does not expect data from this rank");
    const size_t pdbytes = hl_data_bytes(peer_nghost);
    hl_dest *d = &a->d[q];
    d->peer_data = (double *)h->peer[r];
    d->peer_arrived = hl_arrived(h->peer[r], pdbytes) + slot_there;
    d->my_ack = hl_ack(h->block, dbytes) + q;
    d->done = done + 2 * q;
    d->idx = sendidx_dev ? sendidx_dev + soff : NULL;
    d->xoff = sendoff_host ? (long long)sendoff_host[q] : 0;
    d->count = sendcount[q];
    d->peer_off = (long long)pr[10 + 3 * h->rank];
    d->peer_nghost = peer_nghost;
    soff += sendcount[q];
  }
  for (int p = 0; p < nrecv; p++) {
    const int r = recvrank[p];
    const double *pr = all + words * (size_t)r;
    const int slot_there = (int)pr[11 + 3 * h->rank];
    ARGCHK(slot_there >= 0, "halo plans do not match: a source does not send to this rank");
    a->peer_ack[p] = hl_ack(h->peer[r], hl_data_bytes((long long)pr[8])) + slot_there;
  }
  free(mine); free(all);
  {
    long long maxcount = 0;
    for (int q = 0; q < nsend; q++) if (sendcount[q] > maxcount) maxcount = sendcount[q];
    long long c = (maxcount + 8191) / 8192;
    h->ctas = (int)(c < HL_CTAS ? HL_CTAS : (c > HL_CTAS_MAX ? HL_CTAS_MAX : c));
  }
  h->seq = 0;
  *out = h;
  return B2K_OK;
}

/* push my boundary entries of x, wait for the neighbours'; *ghost_out = the buffer this SpMV must read */
extern "C" int b2k_halo_exchange(b2k_halo h, const double *x, const double **ghost_out)
{
  b2k_ctx ctx = h->ctx;
  const u64 seq = ++h->seq;
  if (h->args.nsend > 0 || h->args.nrecv > 0) {
    dim3 grid(h->ctas, h->args.nsend > 0 ? h->args.nsend : 1);
    b2k_launch_pdl(k_halo_push, grid, dim3(256), 0, ctx->stream, x, h->args, seq);
    CKLAUNCH(ctx);
  }
  if (h->args.nrecv > 0) {
    b2k_launch_pdl(k_halo_wait, dim3(1), dim3(32), 0, ctx->stream, h->args, seq);
    CKLAUNCH(ctx);
  }
  if (ghost_out) *ghost_out = (const double *)h->block + (seq & 1ull) * hl_stride(h->nghost);
  return B2K_OK;
}

extern "C" int b2k_halo_error(b2k_halo h, int *flag)
{
  if (flag) *flag = (h && h->err_host) ? *(volatile int *)h->err_host : 0;
  return B2K_OK;
}

/* collective */
extern "C" int b2k_halo_destroy(b2k_halo h)
{
  if (!h) return B2K_OK;
  cudaStreamSynchronize(h->ctx->stream);
  for (int r = 0; r < HL_MAXP; r++) if (h->peer[r]) cudaIpcCloseMemHandle(h->peer[r]);
  b2k_comm_barrier(h->comm);           /* nobody maps my block any more */
  cudaFree(h->block);
  if (h->err_host) cudaFreeHost(h->err_host);
  free(h);
  return B2K_OK;
}
