/*
 * b2k_gs_fused.cu — single-sweep classical Gram-Schmidt update for sm_100a.
 *
 *     w  <- w - V(:,0:k) cin                       (update of pass p,   BVMultColumn  bvorthog.c:113)
 *     cout[0:k] = V(:,0:k)^T w_new, cout[k] = ||w_new||^2   (dot sweep of pass p+1, BVDotColumnInc :100,
 *                                                   and the explicit norm of :126)
 *
 * The reference reads the n x k basis block twice per pass (gemv 'C' then gemv 'N', bvcuda.cu:204,45)
 * and needs 2 passes whenever DGKS refinement triggers (bvorthog.c:180), i.e. 4 reads.  Here a CTA
 * keeps a 128-row x k tile of V in shared memory, so the update AND the next pass' inner products
 * are computed from ONE read of V from HBM: a 2-pass orthogonalisation costs 3 reads instead of 4.
 *
 * Data movement: a dedicated producer warp streams the tile columns with 1-D bulk async copies
 * (cp.async.bulk.shared::cluster.global, SASS UBLKCP) into a 3..6-stage ring, completion signalled
 * through mbarriers (complete_tx::bytes); 8 consumer warps compute out of shared memory:
 *   phase 1  (row-wise)    thread = (row pair, column quarter): partial w update, LDS.128 conflict-free
 *   phase 2  (column-wise) thread = (column, row class): accumulators live in registers for the whole
 *            kernel; column stride 130 doubles makes the 8-column quarter-warp reads conflict-free
 * Reductions are two-stage and fixed-order (bit-reproducible).  One persistent CTA per SM.
 */
#include <stdlib.h>
#include "b2k_internal.h"

#define F_RB      128                 /* rows per tile                              */
#define F_SCOL    (F_RB + 2)          /* shared-memory column stride (doubles)      */
#define F_CONS    256                 /* consumer threads                            */
#define F_THREADS (F_CONS + 32)       /* + one producer warp                         */
#define F_MAXK    64
#define F_MAXSTAGES 6

static int g_fused_enabled = -1;
int b2k_gs_fused_enabled(void)
{
  if (g_fused_enabled < 0) {
    const char *e = getenv("B2K_GS_FUSED");
    g_fused_enabled = (e && e[0] >= '0' && e[0] <= '3') ? e[0] - '0' : 3;
  }
  return g_fused_enabled;
}
/* 0: generic two-sweep kernels, 1: register-tile single sweep, 2: 1-D bulk-copy staged single sweep,
   3: 2-D tensor-map (TMA) pipelined single sweep */
extern "C" int b2k_gs_set_fused(int mode) { g_fused_enabled = (mode >= 0 && mode <= 3) ? mode : 3; return B2K_OK; }

/* ---- PTX helpers ------------------------------------------------------------------------------ */
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity)
{
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
  while (!mbar_try_wait(bar, parity)) { }
}
/* barrier among the 256 consumer threads only (the producer warp never joins) */
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

struct FusedSmem {
  unsigned long long full[F_MAXSTAGES];
  unsigned long long empty[F_MAXSTAGES];
  double psum[4][F_RB];
  double wn[F_RB];
  double red[8][8];
  double rednrm[8];
};

__global__ void __launch_bounds__(F_THREADS, 1)
k_gs_fused(const double *__restrict__ V, int64_t ld, int64_t n, int k, double *__restrict__ w, const double *__restrict__ cin,
           double *__restrict__ part, int pstride, int nstages)
{
  extern __shared__ __align__(128) unsigned char smraw[];
  FusedSmem *S = reinterpret_cast<FusedSmem *>(smraw);
  const size_t stage_doubles = (size_t)k * F_SCOL + F_RB;                  /* tile columns + w tile */
  double *stages = reinterpret_cast<double *>(smraw + ((sizeof(FusedSmem) + 127) / 128) * 128);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = n / F_RB;
  const int64_t ntail = n - ntiles * F_RB;
  /* contiguous chunk of tiles per CTA */
  const int64_t per = (ntiles + gridDim.x - 1) / gridDim.x;
  const int64_t t_begin = (int64_t)blockIdx.x * per;
  const int64_t t_end = (t_begin + per < ntiles) ? t_begin + per : ntiles;

  if (tid == 0) {
    for (int s = 0; s < nstages; s++) {
      mbar_init(smem_u32(&S->full[s]), 1);
      mbar_init(smem_u32(&S->empty[s]), 8);       /* one arrival per consumer warp */
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == 8) {
    /* ================= producer warp ================= */
    const uint32_t tile_bytes = (uint32_t)((size_t)k * F_RB * 8 + F_RB * 8);
    int s = 0;
    uint32_t ph = 0;
    for (int64_t t = t_begin; t < t_end; t++) {
      mbar_wait(smem_u32(&S->empty[s]), ph ^ 1);                       /* slot free (first round passes) */
      double *st = stages + (size_t)s * stage_doubles;
      const uint32_t bar = smem_u32(&S->full[s]);
      if (lane == 0) mbar_arrive_expect_tx(bar, tile_bytes);
      __syncwarp();
      const int64_t row0 = t * F_RB;
      for (int c = lane; c < k; c += 32)
        bulk_g2s(smem_u32(st + (size_t)c * F_SCOL), V + (int64_t)c * ld + row0, F_RB * 8, bar);
      if (lane == 0) bulk_g2s(smem_u32(st + (size_t)k * F_SCOL), w + row0, F_RB * 8, bar);
      if (++s == nstages) { s = 0; ph ^= 1; }
    }
    return;
  }

  /* ================= consumer warps ================= */
  /* phase-1 mapping: row pair rp (0..63) x column quarter cg (0..3) */
  const int rp = tid & 63, cg = tid >> 6;
  const int kq = (k + 3) >> 2;                     /* columns per quarter (<= 16) */
  double creg[16];
#pragma unroll
  for (int u = 0; u < 16; u++) {
    const int c = cg * kq + u;
    creg[u] = (u < kq && c < k) ? cin[c] : 0.0;
  }
  /* phase-2 mapping: column group g (8 columns) x row slice; thread column c2, row class qp */
  int ncg = 1;
  while (ncg * 8 < k) ncg <<= 1;                   /* 1,2,4,8 */
  const int nslice = 8 / ncg;
  const int g = warp % ncg, slice = warp / ncg;
  const int c2 = 8 * g + (lane & 7), qp = lane >> 3;
  const int it_per = (F_RB / 8) / nslice;          /* i' iterations per slice */
  const int it0 = slice * it_per;
  const bool col_ok = c2 < k;
  double acc = 0.0;       /* V(:,c2)^T w_new over my rows */
  double nrm = 0.0;       /* ||w_new||^2 over rows finalised by me */

  int s = 0;
  uint32_t ph = 0;
  for (int64_t t = t_begin; t < t_end; t++) {
    mbar_wait(smem_u32(&S->full[s]), ph);
    const double *tile = stages + (size_t)s * stage_doubles;
    const double *wt = tile + (size_t)k * F_SCOL;
    /* ---- phase 1: partial row sums over my column quarter ---- */
    {
      double ax = 0.0, ay = 0.0;
      const double *tp = tile + (size_t)(cg * kq) * F_SCOL + 2 * rp;
#pragma unroll
      for (int u = 0; u < 16; u++) {
        if (u < kq) {        /* columns beyond k have creg == 0 but must not be read: guard */
          if (cg * kq + u < k) {
            const double2 v = *reinterpret_cast<const double2 *>(tp + (size_t)u * F_SCOL);
            ax = fma(v.x, creg[u], ax);
            ay = fma(v.y, creg[u], ay);
          }
        }
      }
      *reinterpret_cast<double2 *>(&S->psum[cg][2 * rp]) = make_double2(ax, ay);
    }
    consumer_sync();
    if (tid < 64) {
      const double2 wv = *reinterpret_cast<const double2 *>(wt + 2 * tid);
      const double2 p0 = *reinterpret_cast<const double2 *>(&S->psum[0][2 * tid]);
      const double2 p1 = *reinterpret_cast<const double2 *>(&S->psum[1][2 * tid]);
      const double2 p2 = *reinterpret_cast<const double2 *>(&S->psum[2][2 * tid]);
      const double2 p3 = *reinterpret_cast<const double2 *>(&S->psum[3][2 * tid]);
      double2 o;
      o.x = wv.x - ((p0.x + p1.x) + (p2.x + p3.x));
      o.y = wv.y - ((p0.y + p1.y) + (p2.y + p3.y));
      *reinterpret_cast<double2 *>(&S->wn[2 * tid]) = o;
      *reinterpret_cast<double2 *>(w + t * F_RB + 2 * tid) = o;
      nrm = fma(o.x, o.x, nrm);
      nrm = fma(o.y, o.y, nrm);
    }
    consumer_sync();
    /* ---- phase 2: column inner products with the updated w ---- */
    if (col_ok) {
      const double *tc = tile + (size_t)c2 * F_SCOL + 2 * qp;
      const double *wq = S->wn + 2 * qp;
      double a0 = 0.0, a1 = 0.0;
#pragma unroll 4
      for (int i = it0; i < it0 + it_per; i++) {
        const double2 v = *reinterpret_cast<const double2 *>(tc + 8 * i);
        const double2 x = *reinterpret_cast<const double2 *>(wq + 8 * i);
        a0 = fma(v.x, x.x, a0);
        a1 = fma(v.y, x.y, a1);
      }
      acc += a0 + a1;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&S->empty[s]));     /* this warp is done with the stage */
    if (++s == nstages) { s = 0; ph ^= 1; }
  }

  /* ---- tail rows (n mod 128), handled by the last CTA straight from global memory ---- */
  if (ntail > 0 && blockIdx.x == gridDim.x - 1) {
    const int64_t row0 = ntiles * F_RB;
    consumer_sync();
    if (tid < ntail) {
      const int64_t r = row0 + tid;
      double sacc = 0.0;
      for (int c = 0; c < k; c++) sacc = fma(V[(int64_t)c * ld + r], cin[c], sacc);
      const double o = w[r] - sacc;
      w[r] = o;
      S->wn[tid] = o;
      nrm = fma(o, o, nrm);
    }
    consumer_sync();
    if (col_ok && qp == 0 && slice == 0) {
      double a = 0.0;
      for (int r = 0; r < (int)ntail; r++) a = fma(V[(int64_t)c2 * ld + row0 + r], S->wn[r], a);
      acc += a;
    }
  }

  /* ---- CTA reduction: over row classes (shuffle), then over row slices (shared, fixed order) ---- */
  acc += __shfl_xor_sync(0xffffffffu, acc, 8);
  acc += __shfl_xor_sync(0xffffffffu, acc, 16);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
  if (lane < 8) S->red[warp][lane] = acc;
  if (lane == 0) S->rednrm[warp] = nrm;
  consumer_sync();
  if (tid < k) {
    const int gg = tid >> 3;
    double sres = 0.0;
    for (int sl = 0; sl < nslice; sl++) sres += S->red[gg + ncg * sl][tid & 7];
    part[(int64_t)blockIdx.x * pstride + tid] = sres;
  }
  if (tid == 64) {
    double sres = 0.0;
#pragma unroll
    for (int q = 0; q < 8; q++) sres += S->rednrm[q];
    part[(int64_t)blockIdx.x * pstride + k] = sres;
  }
}

int b2k_launch_reduce_partials(b2k_ctx ctx, int nblk, int pstride, int ncols, double *out);   /* b2k_bv.cu */

int b2k_gs_update_dot_fused(b2k_ctx ctx, const double *V, int64_t ld, int64_t n, int k, double *w, const double *cin, double *cout)
{
  if (k < 1 || k > F_MAXK || n < F_RB) return -1;
  if (!b2k_is_aligned16(V) || !b2k_is_aligned16(w) || (ld & 1)) return -1;
  const size_t fixed = ((sizeof(FusedSmem) + 127) / 128) * 128;
  const size_t stage_bytes = ((size_t)k * F_SCOL + F_RB) * sizeof(double);
  const size_t budget = 220 * 1024;
  int nstages = (int)((budget - fixed) / stage_bytes);
  if (nstages > F_MAXSTAGES) nstages = F_MAXSTAGES;
  if (nstages < 2) return -1;
  const size_t shm = fixed + (size_t)nstages * stage_bytes;
  static size_t configured = 0;
  if (shm > configured) {
    CK(cudaFuncSetAttribute(k_gs_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    configured = budget;
  }
  const int64_t ntiles = n / F_RB;
  int grid = ctx->sm_count;
  if ((int64_t)grid > ntiles) grid = (int)(ntiles > 0 ? ntiles : 1);
  const int pstride = k + 1;
  PROF_BEGIN(ctx, B2K_PROF_GSFUSED, 8.0 * (double)n * (k + 2));
  k_gs_fused<<<grid, F_THREADS, shm, ctx->stream>>>(V, ld, n, k, w, cin, ctx->partials, pstride, nstages);
  PROF_END(ctx);
  CKLAUNCH(ctx);
  return b2k_launch_reduce_partials(ctx, grid, pstride, k + 1, cout);
}
