/* internal declarations shared by the .cu files of libb200krylov.so */
#ifndef B2K_INTERNAL_H
#define B2K_INTERNAL_H
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "b2k.h"

/* cross-GPU one-shot reduction over NVLink peer memory (b2k_comm.cu, k_reduce_partials_xg in b2k_bv.cu): every rank owns
   a mailbox in its HBM that all ranks of the box map through CUDA IPC; passed BY VALUE to the reduction kernel */
#define B2K_XG_MAXR 8                 /* ranks (GPUs of one NVSwitch box)                                  */
#define B2K_XG_MAXC 1032              /* doubles per (parity, source rank): >= B2K_MAX_K + 1               */
#define B2K_XG_MAXB 136               /* flags per (parity, source rank): one per reduction CTA (8 columns) */
#define B2K_XG_DATA_ELEMS (2 * B2K_XG_MAXR * B2K_XG_MAXC)
#define B2K_XG_FLAG_ELEMS (2 * B2K_XG_MAXR * B2K_XG_MAXB)
#define B2K_XG_BYTES      (sizeof(double) * B2K_XG_DATA_ELEMS + sizeof(unsigned long long) * B2K_XG_FLAG_ELEMS)
struct b2k_xg_s {
  int     rank, size;
  double *box[B2K_XG_MAXR];           /* box[p] = mailbox of rank p as mapped in THIS process (box[rank] is local) */
  int    *err;                        /* mapped pinned host flag: set when a peer did not show up in time  */
  long long spin_limit;               /* clock64 ticks a kernel waits for a peer before it raises err (env B2K_SPIN_TIMEOUT_S, default 60 s) */
};
static inline long long b2k_spin_limit(void)
{
  const char *e = getenv("B2K_SPIN_TIMEOUT_S");
  double s = e ? atof(e) : 60.0;
  if (!(s > 0.0)) s = 60.0;
  return (long long)(s * 2.0e9);
}

struct b2k_ctx_s {
  int          device;
  int          sm_count;
  cudaStream_t stream;
  double      *partials;        /* scratch for two-stage reductions: [B2K_MAX_PART_BLOCKS][pstride] */
  size_t       partials_elems;
  double      *dscratch;        /* small device scratch (host scalars/vectors staged to HBM)       */
  size_t       dscratch_elems;
  cudaEvent_t  ev0, ev1;
  uint64_t     launches;
  uint64_t     syncs;           /* host waits on the stream issued through this context (b2k_ctx_sync, blocking copies) */
  uint64_t     h2d_bytes, d2h_bytes;
  /* optional per-kernel-class timing with CUDA events on the launching stream (b2k_prof_*) */
  int          prof_on, prof_n, prof_cap;
  cudaEvent_t *prof_ev;         /* 2 events per recorded launch */
  int         *prof_id;
  double      *prof_bytes;
  double      *prof_flop;      /* floating-point operations of the recorded launch (level-3 kernels)  */
  double       prof_ms[B2K_PROF_NCLASS], prof_b[B2K_PROF_NCLASS], prof_f[B2K_PROF_NCLASS];
  uint64_t     prof_cnt[B2K_PROF_NCLASS];
  /* when xg_on, every two-stage reduction launched through b2k_launch_reduce_partials also sums over the ranks */
  b2k_xg_s    *xg;
  int          xg_on;
  unsigned long long xg_seq;
};

int b2k_launch_reduce_partials(b2k_ctx ctx, int nblk, int pstride, int ncols, double *out);   /* b2k_bv.cu */

/* Device-side DGKS decision (bvorthog.c:180, `while (l<3 && nrm && nrm < eta*onrm)`): a GATED update sweep is launched
   speculatively right behind the sweep that produces the two norms and runs only when the refinement criterion holds, so the
   host needs ONE synchronisation per column instead of one per pass.  Every CTA evaluates the predicate itself from the two
   device scalars with the host's own IEEE operations (sqrt, one multiply), so host and device always take the same decision.
   onrm2 == NULL: not gated. */
struct b2k_gate_s { const double *onrm2, *nrm2; double eta; };
#ifdef __CUDACC__
__device__ __forceinline__ bool b2k_gate_closed(const b2k_gate_s &g)
{
  if (!g.onrm2) return false;
  const double a = *reinterpret_cast<const volatile double *>(g.onrm2), b = *reinterpret_cast<const volatile double *>(g.nrm2);
  const double onrm = sqrt(a > 0.0 ? a : 0.0), nrm = sqrt(b > 0.0 ? b : 0.0);
  return !(nrm != 0.0 && fabs(nrm) < g.eta * fabs(onrm));
}
#endif

/* bracket the dominant kernel of an entry point; `bytes` = algorithmic bytes of this launch */
#define PROF_BEGIN(ctx, cls, bytes)                                                          \
  const int prof_slot_ = ((ctx)->prof_on && (ctx)->prof_n < (ctx)->prof_cap) ? (ctx)->prof_n++ : -1; \
  if (prof_slot_ >= 0) {                                                                      \
    (ctx)->prof_id[prof_slot_] = (cls);                                                       \
    (ctx)->prof_bytes[prof_slot_] = (double)(bytes);                                          \
    (ctx)->prof_flop[prof_slot_] = 0.0;                                                       \
    cudaEventRecord((ctx)->prof_ev[2 * prof_slot_], (ctx)->stream);                           \
  }
#define PROF_FLOPS(ctx, f) do { if (prof_slot_ >= 0) (ctx)->prof_flop[prof_slot_] = (double)(f); } while (0)
#define PROF_END(ctx)                                                                         \
  if (prof_slot_ >= 0) cudaEventRecord((ctx)->prof_ev[2 * prof_slot_ + 1], (ctx)->stream);

#define B2K_MAX_PART_BLOCKS 2048
#define B2K_MAX_K           B2K_MAX_COLUMNS   /* max columns in one reduction (nc+ncv+1 <= 1024), public in b2k.h */

void b2k_set_error(const char *fmt, ...);

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      b2k_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));    \
      return B2K_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

#define CKLAUNCH(ctx)                                                                         \
  do {                                                                                        \
    (ctx)->launches++;                                                                        \
    cudaError_t e_ = cudaGetLastError();                                                      \
    if (e_ != cudaSuccess) {                                                                  \
      b2k_set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_));\
      return B2K_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

#define ARGCHK(cond, msg)                                                                     \
  do {                                                                                        \
    if (!(cond)) {                                                                            \
      b2k_set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, msg);                  \
      return B2K_ERR_ARG;                                                                     \
    }                                                                                         \
  } while (0)

/* Programmatic dependent launch (sm_90+): the kernels of the Lanczos-step chain (sweeps, their reductions, scale, SpMV, halo)
   are launched with the programmatic-stream-serialization attribute and open with b2k_pdl_enter(): `launch_dependents` lets the
   NEXT kernel of the stream be scheduled while this one runs (its CTAs become resident as resources allow), `wait` blocks until
   the PREVIOUS kernel has completed and its writes are visible.  Data dependences are untouched (every chain kernel waits before
   its first global access); what disappears is the launch / drain gap between the 8-10 short kernels of a step in the
   latency-bound regime.  env B2K_PDL=0: plain stream order. */
#ifdef __CUDACC__
__device__ __forceinline__ void b2k_pdl_enter(void)
{
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
static inline int b2k_pdl_enabled(void)
{
  static int on = -1;
  if (on < 0) { const char *e = getenv("B2K_PDL"); on = (e && e[0] == '0') ? 0 : 1; }
  return on;
}
template <typename... KArgs, typename... Args>
static inline void b2k_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t shm, cudaStream_t st, Args &&...args)
{
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute at[1];
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = shm; cfg.stream = st;
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = b2k_pdl_enabled() ? 1 : 0;
  (void)cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);       /* errors surface through cudaGetLastError (CKLAUNCH) */
}
#endif

static inline int b2k_is_aligned16(const void *p) { return (((uintptr_t)p) & 15u) == 0; }

#endif
