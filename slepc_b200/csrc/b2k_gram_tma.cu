/*
 * b2k_gram_tma.cu — one-sweep block inner product  M(ky x kx) = Y(:,0:ky)^T X(:,0:kx)  on the FP64 tensor cores (sm_100a).
 * Replaces BVDot_BLAS_CUDA (bvcuda.cu:140-199: ONE cublasDgemm 'C','N' on the tall panels), reached from BVDot
 * (bvglobal.c:52) by the block orthogonalisations (BVOrthogonalize CHOL / SVQB, bvorthog.c:595-680, bvlapack.c:136-345) and by
 * every block solver.  The round-1 implementation swept Y once per column of X (kx fused dot sweeps): 8 n ky kx bytes; this
 * kernel reads Y and X ONCE: 8 n (ky + kx) bytes, 8 n k when X and Y are the same block (the Gram matrix of CHOL / SVQB).
 *
 *   - a producer thread streams 64-row boxes of Y (and of X when it is a different block) through a shared-memory ring with
 *     cp.async.bulk.tensor.2d (rows past n and columns past ky / kx are zero-filled by the TMA unit: no tail code, and the
 *     zero rows add nothing to the sums);
 *   - 8 consumer warps tile the ky x kx result in 8 x 8 blocks; with rows as the contraction index a block update is one
 *     mma.sync.m8n8k4.f64 (SASS DMMA; tcgen05 has no FP64 kind) per 4 rows: A = Y-tile^T and B = X-tile fragments are the SAME
 *     access pattern (row 4s + lane%4 of column base + lane/4), conflict-free with the 68-row column stride;
 *   - every CTA keeps its ky x kx partial in registers over all of its tiles and writes it once; the partials are summed in
 *     CTA order by k_reduce_partials (bit-reproducible, no atomics).
 * Arithmetic intensity 2 ky kx / (8 (ky + kx)): 8 flop/B at 64 x 64 (16 for the Gram case) — above the FP64 ridge of B200
 * (5.7 flop/B), so the sweep is bound by the FP64 pipe, not by HBM; both fractions are reported (B2K_PROF_GEMM).
 */
#include <cuda.h>
#include <stdlib.h>
#include "b2k_internal.h"

#define GR_ROWS    64
#define GR_SROWS   68                /* rows per box = column stride in shared memory: 544 B = 32 B mod 128 B → the 8 columns x 4 rows
                                        of a fragment load fall on 32 distinct 8-byte bank pairs (two conflict-free wavefronts) */
#define GR_KB      64
#define GR_CONS    256
#define GR_THREADS (GR_CONS + 32)
#define GR_MAXSTG  6

int b2k_tm_make_map(CUtensorMap *map, const double *base, int64_t n, int64_t ncols, int64_t ld, int box_cols, int box_rows);   /* b2k_gs_tma.cu */
int b2k_launch_reduce_partials(b2k_ctx ctx, int nblk, int pstride, int ncols, double *out);                                      /* b2k_bv.cu */

__device__ __forceinline__ uint32_t gr_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t gr_try_wait(uint32_t bar, uint32_t parity)
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
__device__ __forceinline__ void gr_wait(uint32_t bar, uint32_t parity) { while (!gr_try_wait(bar, parity)) { } }
__device__ __forceinline__ void gr_dmma(double &d0, double &d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

/* one 64-row tile for a warp that owns nb <= NBP blocks: RS = 8 / NBP row splits give every warp 8 independent DMMA chains
   (accumulator u*RS + sp), so narrow results (few blocks per warp) are not bound by the latency of a single dependent chain —
   measured before the split: 32 x 32 took as long as 64 x 64 (profiles/r02/r02e_kbench2.jsonl vs r02f_kbench2.jsonl) */
template <int NBP>
__device__ __forceinline__ void gr_tile(const double *__restrict__ ap, const double *__restrict__ bp, int G, int nb, double (&c0)[8], double (&c1)[8])
{
  constexpr int RS = 8 / NBP, KS = (GR_ROWS / 4) / RS;
#pragma unroll
  for (int ks = 0; ks < KS; ks++) {
#pragma unroll
    for (int sp = 0; sp < RS; sp++) {
      const int kk = ks + sp * KS;
      const double a = ap[4 * kk];
#pragma unroll
      for (int u = 0; u < NBP; u++) {
        if (u < nb) {                                            /* warp-uniform */
          const double b = bp[(size_t)(8 * u * G) * GR_SROWS + 4 * kk];
          gr_dmma(c0[u * RS + sp], c1[u * RS + sp], a, b);
        }
      }
    }
  }
}

/* SAME: X is Y (one box per stage).  The 8 warps own (block row r = warp % nbr) x (block columns g, g + G, …), G = 8 / nbr
   groups, nbr = ceil(ky/8) block rows: 8 blocks per warp at 64 x 64, fewer for narrower results. */
template <bool SAME>
__global__ void __launch_bounds__(GR_THREADS, 1)
k_gram_tma(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX, int64_t n, int ky, int kx,
           double *__restrict__ part, int pstride, int nstages)
{
  extern __shared__ __align__(1024) unsigned char gr_raw[];
  unsigned long long *full = reinterpret_cast<unsigned long long *>(gr_raw);
  unsigned long long *empty = full + GR_MAXSTG;
  double *stages = reinterpret_cast<double *>(gr_raw + 1024);
  constexpr int BOX = GR_KB * GR_SROWS;                          /* doubles per box */
  constexpr int STAGE = SAME ? BOX : 2 * BOX;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (n + GR_ROWS - 1) / GR_ROWS;
  if (tid == 0) {
    for (int s = 0; s < nstages; s++) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gr_smem_u32(&full[s])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gr_smem_u32(&empty[s])), "r"(8));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == 8) {                                               /* ---- producer ---- */
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmY)) : "memory");
      if (!SAME) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
      const uint32_t bytes = (uint32_t)(STAGE * sizeof(double));
      int s = 0;
      uint32_t ph = 0;
      for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        gr_wait(gr_smem_u32(&empty[s]), ph ^ 1);
        double *st = stages + (size_t)s * STAGE;
        const uint32_t bar = gr_smem_u32(&full[s]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                         gr_smem_u32(st)),
                     "l"(reinterpret_cast<uint64_t>(&tmY)), "r"((int)(t * GR_ROWS)), "r"(0), "r"(bar)
                     : "memory");
        if (!SAME)
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                           gr_smem_u32(st + BOX)),
                       "l"(reinterpret_cast<uint64_t>(&tmX)), "r"((int)(t * GR_ROWS)), "r"(0), "r"(bar)
                       : "memory");
        if (++s == nstages) { s = 0; ph ^= 1; }
      }
    }
    return;
  }

  /* ---- consumers ---- */
  const int nbr = (ky + 7) >> 3, nbc = (kx + 7) >> 3;
  const int G = 8 / nbr;                                         /* column groups (nbr <= 8) */
  const bool active = warp < nbr * G;
  const int r = warp % nbr, g = warp / nbr;
  const int nb = (active && g < nbc) ? (nbc - g + G - 1) / G : 0; /* my 8 x 8 blocks: columns g, g + G, … */
  const int fr = lane >> 2, fk = lane & 3;
  double c0[8], c1[8];
#pragma unroll
  for (int u = 0; u < 8; u++) { c0[u] = 0.0; c1[u] = 0.0; }
  int s = 0;
  uint32_t ph = 0;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    gr_wait(gr_smem_u32(&full[s]), ph);
    if (nb > 0) {
      const double *Ys = stages + (size_t)s * STAGE;
      const double *Xs = SAME ? Ys : Ys + BOX;
      const double *ap = Ys + (size_t)(8 * r + fr) * GR_SROWS + fk;
      const double *bp = Xs + (size_t)(8 * g + fr) * GR_SROWS + fk;
      /* always 8 independent accumulator chains per warp: a warp with fewer than 8 blocks splits the 64 rows of the tile */
      if (nb > 4) gr_tile<8>(ap, bp, G, nb, c0, c1);
      else if (nb > 2) gr_tile<4>(ap, bp, G, nb, c0, c1);
      else if (nb > 1) gr_tile<2>(ap, bp, G, nb, c0, c1);
      else gr_tile<1>(ap, bp, G, nb, c0, c1);
    }
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(gr_smem_u32(&empty[s])) : "memory");
    if (++s == nstages) { s = 0; ph ^= 1; }
  }
  /* the CTA's partial: M(8r + fr, 8c + 2fk + {0,1}), column-major ky x kx; row splits summed in fixed order */
  if (nb > 0) {
    double *P = part + (size_t)blockIdx.x * pstride;
    const int row = 8 * r + fr;
    const int nbp = nb > 4 ? 8 : (nb > 2 ? 4 : (nb > 1 ? 2 : 1)), rs = 8 / nbp;
#pragma unroll
    for (int u = 0; u < 8; u++) {
      if (u < nb) {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int q = 0; q < 8; q++)
          if (q >= u * rs && q < (u + 1) * rs) { a0 += c0[q]; a1 += c1[q]; }
        const int col = 8 * (g + u * G) + 2 * fk;
        if (row < ky) {
          if (col < kx) P[(size_t)col * ky + row] = a0;
          if (col + 1 < kx) P[(size_t)(col + 1) * ky + row] = a1;
        }
      }
    }
  }
}

static int g_gram = -1;
/* -1: shape not supported (the caller falls back to the column-by-column sweeps).  M is ky x kx with leading dimension ldm. */
int b2k_gram_tma_launch(b2k_ctx ctx, const double *Y, int64_t ldy, const double *X, int64_t ldx, int64_t n, int ky, int kx, double *M, int ldm)
{
  if (g_gram < 0) { const char *e = getenv("B2K_GRAM_TMA"); g_gram = (e && e[0] == '0') ? 0 : 1; }
  if (!g_gram) return -1;
  if (ky < 1 || ky > GR_KB || kx < 1 || kx > GR_KB || ldm != ky) return -1;
  if (n < 32 * GR_ROWS || n >= 2147483647LL - GR_SROWS) return -1;
  if (!b2k_is_aligned16(Y) || !b2k_is_aligned16(X) || (ldy & 1) || (ldx & 1)) return -1;
  const bool same = (Y == X && ldy == ldx);
  CUtensorMap mY, mX;
  if (b2k_tm_make_map(&mY, Y, n, same ? (ky > kx ? ky : kx) : ky, ldy, GR_KB, GR_SROWS)) return -1;
  if (same) mX = mY;
  else if (b2k_tm_make_map(&mX, X, n, kx, ldx, GR_KB, GR_SROWS)) return -1;
  const int nstages = same ? 6 : 3;
  const size_t shm = 1024 + (size_t)nstages * (same ? 1 : 2) * GR_KB * GR_SROWS * sizeof(double);
  const int64_t ntiles = (n + GR_ROWS - 1) / GR_ROWS;
  int grid = ctx->sm_count;
  if ((int64_t)grid > ntiles) grid = (int)ntiles;
  const int pstride = ky * kx;
  if ((size_t)grid * (size_t)pstride > ctx->partials_elems) return -1;
  static int configured = 0;
  if (!configured) {
    const int maxshm = (int)(1024 + (size_t)6 * GR_KB * GR_SROWS * sizeof(double));
    CK(cudaFuncSetAttribute(k_gram_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxshm));
    CK(cudaFuncSetAttribute(k_gram_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxshm));
    configured = 1;
  }
  PROF_BEGIN(ctx, B2K_PROF_GEMM, 8.0 * (double)n * (same ? ky : ky + kx));
  PROF_FLOPS(ctx, 2.0 * (double)n * ky * kx);
  if (same) k_gram_tma<true><<<grid, GR_THREADS, shm, ctx->stream>>>(mY, mX, n, ky, kx, ctx->partials, pstride, nstages);
  else k_gram_tma<false><<<grid, GR_THREADS, shm, ctx->stream>>>(mY, mX, n, ky, kx, ctx->partials, pstride, nstages);
  PROF_END(ctx);
  CKLAUNCH(ctx);
  /* CTA partials → M, fixed order; local to this GPU (the caller all-reduces the ky x kx block) */
  const int xg_save = ctx->xg_on;
  ctx->xg_on = 0;
  const int rc = b2k_launch_reduce_partials(ctx, grid, pstride, pstride, M);
  ctx->xg_on = xg_save;
  return rc;
}
