/*
 * b2k_vq_tma.cu — restart GEMM  Out(n x nout) = beta*Out + alpha * In(n x kin) * Qe(kin x nout), kin, nout <= 64, Out may alias
 * columns of In (BVMultInPlace, bvops.c:220-245; replaces BVMultInPlace_BLAS_CUDA bvcuda.cu:65-112), as a TMA-fed
 * FP64 tensor-core kernel.
 *
 * ncu on k_vq (profiles/r01_ncu_summary.md): FP64 pipe 44 %, DRAM 36 % — the shape sits on the FP64 ridge (AI 5.3 flop/B)
 * and the load / DFMA / store phases of a CTA do not overlap; the DFMA inner loop is also at the shared-memory limit
 * (one LDS.128 of the tile + C/2 broadcast LDS.128 of Q per 2C DFMA).  Here:
 *   - one thread streams 128-row x 64-column boxes of In through a 3-stage shared-memory ring with
 *     cp.async.bulk.tensor.2d (rows past n and columns past kin are zero-filled by the TMA unit, no tail code);
 *   - 8 consumer warps = 2 row halves x 4 column groups run mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4; tcgen05 has no FP64
 *     kind): the Q fragments of a warp's 8 or 16 output columns stay in REGISTERS for the whole kernel, so the only
 *     shared-memory traffic is one LDS.64 of the tile per 2 DMMA (512 FMA) — the FP64 pipe is the only busy unit;
 *   - the whole tile is in shared memory before its first output is stored and a tile's rows belong to one CTA, so
 *     output columns that alias In are safe; loads of the next tiles overlap the math and the stores of this one.
 * DMMA and DFMA have the same peak on B200 (tools/dmma_probe.cu: 37.1 vs 36.5 TFLOP/s); the gain is operand delivery.
 */
#include <cuda.h>
#include <stdlib.h>
#include "b2k_internal.h"

#define VT_ROWS    128
#define VT_SROWS   132                /* rows per box = shared-memory column stride: 1056 B = 32 B mod 128 B, so that the 8 lanes
                                         of a quarter warp (2 row pairs x 4 columns of the m8n8k4 A fragment) hit 8 distinct
                                         16-byte bank groups; with the dense 1024 B stride ncu counted a 4-way conflict on every
                                         LDS.128 (816 M conflict wavefronts of 1085 M); rows 128..131 of a box are not used */
#define VT_KB      64
#define VT_STAGES  3
#define VT_THREADS 256               /* 8 warps, 2 per SM sub-partition: up to 255 registers each; lane 0 of warp 0 also feeds
                                         the ring (a ninth warp would put 3 warps on one sub-partition and cap the kernel at 168) */

int b2k_tm_make_map(CUtensorMap *map, const double *base, int64_t n, int64_t ncols, int64_t ld, int box_cols, int box_rows);   /* b2k_gs_tma.cu */

__device__ __forceinline__ uint32_t vt_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t vt_try_wait(uint32_t bar, uint32_t parity)
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
__device__ __forceinline__ void vt_wait(uint32_t bar, uint32_t parity) { while (!vt_try_wait(bar, parity)) { } }
__device__ __forceinline__ void vt_dmma(double &d0, double &d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

/* KB = columns of a TMA box (= of a shared-memory stage), STAGES = depth of the ring.  <.,.,64,3> is the general shape;
   <.,.,32,6> (default for kin, nout <= 32) halves the box so that twice as many tiles are in flight.
   MT = m-tiles (8 rows each) per warp, NS = 8-column slabs per warp.  The 8 consumer warps tile the 128 x nout output as
   (16/MT row groups) x (MT/2 column groups): <8,1> nout <= 32, <4,3> nout <= 48, <8,2> nout <= 64 — the shape that wastes
   the fewest DMMAs on padding columns. */
template <int MT, int NS, int KB, int STAGES>
__global__ void __launch_bounds__(VT_THREADS, 1)
k_vq_tma(const __grid_constant__ CUtensorMap tmIn, double *Out, int64_t ldo, int64_t n, int kin, int nout, const double *__restrict__ Q,
         int ldq, int qtrans, double alpha, double beta)
{
  extern __shared__ __align__(1024) unsigned char vt_raw[];
  unsigned long long *full = reinterpret_cast<unsigned long long *>(vt_raw);
  unsigned long long *empty = full + STAGES;
  double *stages = reinterpret_cast<double *>(vt_raw + 1024);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (n + VT_ROWS - 1) / VT_ROWS;
  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(vt_smem_u32(&full[s])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(vt_smem_u32(&empty[s])), "r"(8));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  /* producer duty (thread 0): tile `it` of this CTA goes to stage it % STAGES */
  const int64_t nlocal = (ntiles > blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto issue = [&](int64_t it) {
    const int st = (int)(it % STAGES);
    const uint32_t bar = vt_smem_u32(&full[st]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(KB * VT_SROWS * sizeof(double))) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     vt_smem_u32(stages + (size_t)st * KB * VT_SROWS)),
                 "l"(reinterpret_cast<uint64_t>(&tmIn)), "r"((int)((blockIdx.x + it * gridDim.x) * VT_ROWS)), "r"(0), "r"(bar)
                 : "memory");
  };
  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmIn)) : "memory");
    for (int64_t it = 0; it < STAGES && it < nlocal; it++) issue(it);
  }

  /* consumers: warp = (row group rh of 8*MT rows, column group cg of NS slabs); fragment coordinates of m8n8k4: */
  constexpr int CG = MT / 2, WROWS = 8 * MT;
  const int rh = warp / CG, cg = warp % CG;
  const int fr = lane >> 2, fk = lane & 3;        /* A: row fr, k fk;  B: k fk, col fr;  C: row fr, cols 2*fk, 2*fk+1 */
  const int ksteps = (kin + 3) >> 2;
  double bq[KB / 4][NS];
#pragma unroll
  for (int ks = 0; ks < KB / 4; ks++) {
#pragma unroll
    for (int sl = 0; sl < NS; sl++) {
      const int k = 4 * ks + fk, c = (cg * NS + sl) * 8 + fr;
      double v = 0.0;
      if (k < kin && c < nout) v = qtrans ? Q[(int64_t)c + (int64_t)k * ldq] : Q[(int64_t)k + (int64_t)c * ldq];
      bq[ks][sl] = alpha * v;
    }
  }
  int s = 0;
  uint32_t ph = 0;
  for (int64_t it = 0; it < nlocal; it++) {
    const int64_t t = blockIdx.x + it * gridDim.x;
    vt_wait(vt_smem_u32(&full[s]), ph);
    const double *tile = stages + (size_t)s * KB * VT_SROWS;
    const int64_t rbase = t * VT_ROWS + rh * WROWS;
    /* all MT m-tiles of the warp's rows at once: MT*NS independent accumulator chains per warp keep the FP64 tensor pipe
       busy (one chain is bound by the DMMA latency).  Row mapping: m-tile (mp, parity) holds rows 16*mp + 2*fr + parity, so
       ONE conflict-free LDS.128 (rows 2fr, 2fr+1 of column 4ks+fk) feeds two m-tiles and a lane's two results per column
       are a double2. */
    {
      const double *ap = tile + (size_t)fk * VT_SROWS + rh * WROWS + 2 * fr;
      const int64_t row = rbase + 2 * fr;                        /* + 16*mp, and +1 for the odd m-tile */
      double c0[MT][NS], c1[MT][NS];                             /* [2*mp + parity][slab] */
#pragma unroll
      for (int m = 0; m < MT; m++) {
#pragma unroll
        for (int sl = 0; sl < NS; sl++) {
          c0[m][sl] = 0.0; c1[m][sl] = 0.0;
          const int64_t r = row + 16 * (m >> 1) + (m & 1);
          if (beta != 0.0 && r < n) {
            const int col = (cg * NS + sl) * 8 + 2 * fk;
            if (col < nout) c0[m][sl] = beta * Out[(int64_t)col * ldo + r];
            if (col + 1 < nout) c1[m][sl] = beta * Out[(int64_t)(col + 1) * ldo + r];
          }
        }
      }
#pragma unroll
      for (int ks = 0; ks < KB / 4; ks++) {
        if (ks < ksteps) {
          double2 a[MT / 2];
#pragma unroll
          for (int mp = 0; mp < MT / 2; mp++) a[mp] = *reinterpret_cast<const double2 *>(ap + (size_t)(4 * ks) * VT_SROWS + 16 * mp);
#pragma unroll
          for (int sl = 0; sl < NS; sl++) {
#pragma unroll
            for (int mp = 0; mp < MT / 2; mp++) {
              vt_dmma(c0[2 * mp][sl], c1[2 * mp][sl], a[mp].x, bq[ks][sl]);
              vt_dmma(c0[2 * mp + 1][sl], c1[2 * mp + 1][sl], a[mp].y, bq[ks][sl]);
            }
          }
        }
      }
      __syncwarp();                                 /* every read of this stage has been issued and consumed */
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(vt_smem_u32(&empty[s])) : "memory");
#pragma unroll
      for (int mp = 0; mp < MT / 2; mp++) {
        const int64_t r = row + 16 * mp;
#pragma unroll
        for (int sl = 0; sl < NS; sl++) {
          const int col = (cg * NS + sl) * 8 + 2 * fk;
          if (r + 1 < n) {
            if (col < nout) *reinterpret_cast<double2 *>(Out + (int64_t)col * ldo + r) = make_double2(c0[2 * mp][sl], c0[2 * mp + 1][sl]);
            if (col + 1 < nout) *reinterpret_cast<double2 *>(Out + (int64_t)(col + 1) * ldo + r) = make_double2(c1[2 * mp][sl], c1[2 * mp + 1][sl]);
          } else if (r < n) {
            if (col < nout) Out[(int64_t)col * ldo + r] = c0[2 * mp][sl];
            if (col + 1 < nout) Out[(int64_t)(col + 1) * ldo + r] = c1[2 * mp][sl];
          }
        }
      }
    }
    /* refill this stage with the tile STAGES ahead once all 8 warps have released it (they are at most one tile
       behind; the tiles they need next were requested earlier, so this wait cannot deadlock) */
    if (tid == 0 && it + STAGES < nlocal) {
      vt_wait(vt_smem_u32(&empty[s]), ph);
      issue(it + STAGES);
    }
    if (++s == STAGES) { s = 0; ph ^= 1; }
  }
}

static int g_vq_tma = -1;
static int vq_tma_enabled(void)
{
  if (g_vq_tma < 0) {
    const char *e = getenv("B2K_VQ_TMA");
    g_vq_tma = (e && e[0] == '0') ? 0 : 1;
  }
  return g_vq_tma;
}
extern "C" int b2k_vq_set_tma(int on) { g_vq_tma = on ? 1 : 0; return B2K_OK; }

/* -1: shape not supported / disabled → k_vq */
int b2k_vq_tma_launch(b2k_ctx ctx, double *Out, int64_t ldo, const double *In, int64_t ldi, int64_t n, int kin, int nout, const double *Q,
                      int ldq, int qtrans, double alpha, double beta)
{
  if (!vq_tma_enabled()) return -1;
  if (kin < 1 || kin > 64 || nout < 1 || nout > 64 || n < 32 * VT_ROWS || n >= 2147483647LL - VT_ROWS) return -1;
  if (!b2k_is_aligned16(In) || !b2k_is_aligned16(Out) || (ldi & 1) || (ldo & 1)) return -1;
  /* in place is safe only when Out's rows are In's rows: same leading dimension, Out a column of In's block */
  static int narrow = -1;             /* 32-column boxes, 6 stages, <4,1> tiling for kin <= 32: twice the tiles in flight.  Measured on B200 at
                                         25 -> 13 columns (the C3 restart), n = 4.2 M: 5410 GB/s against 3189 GB/s with 64-column boxes
                                         (profiles/r02_kernels.md); env B2K_VQ_NARROW=0 turns it off */
  if (narrow < 0) { const char *e = getenv("B2K_VQ_NARROW"); narrow = (e && e[0] == '0') ? 0 : 1; }
  const bool use_narrow = narrow && kin <= 32 && nout <= 32;
  const int kb = use_narrow ? 32 : VT_KB, nst = use_narrow ? 6 : VT_STAGES;
  CUtensorMap mIn;
  if (b2k_tm_make_map(&mIn, In, n, kin, ldi, kb, VT_SROWS)) return -1;
  const size_t shm = 1024 + (size_t)nst * kb * VT_SROWS * sizeof(double);
  const int64_t ntiles = (n + VT_ROWS - 1) / VT_ROWS;
  int grid = ctx->sm_count;
  if ((int64_t)grid > ntiles) grid = (int)ntiles;
  static int configured = 0;
  if (!configured) {
    CK(cudaFuncSetAttribute(k_vq_tma<8, 1, VT_KB, VT_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(1024 + (size_t)VT_STAGES * VT_KB * VT_SROWS * sizeof(double))));
    CK(cudaFuncSetAttribute(k_vq_tma<4, 3, VT_KB, VT_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(1024 + (size_t)VT_STAGES * VT_KB * VT_SROWS * sizeof(double))));
    CK(cudaFuncSetAttribute(k_vq_tma<8, 2, VT_KB, VT_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(1024 + (size_t)VT_STAGES * VT_KB * VT_SROWS * sizeof(double))));
    if (narrow) {
      CK(cudaFuncSetAttribute(k_vq_tma<4, 1, 32, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(1024 + (size_t)6 * 32 * VT_SROWS * sizeof(double))));
      CK(cudaFuncSetAttribute(k_vq_tma<8, 1, 32, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(1024 + (size_t)6 * 32 * VT_SROWS * sizeof(double))));
    }
    configured = 1;
  }
  PROF_BEGIN(ctx, B2K_PROF_GEMM, 8.0 * (double)n * (kin + nout));
  PROF_FLOPS(ctx, 2.0 * (double)n * kin * nout);
  if (use_narrow && nout <= 16) k_vq_tma<4, 1, 32, 6><<<grid, VT_THREADS, shm, ctx->stream>>>(mIn, Out, ldo, n, kin, nout, Q, ldq, qtrans, alpha, beta);
  else if (use_narrow) k_vq_tma<8, 1, 32, 6><<<grid, VT_THREADS, shm, ctx->stream>>>(mIn, Out, ldo, n, kin, nout, Q, ldq, qtrans, alpha, beta);
  else if (nout <= 32) k_vq_tma<8, 1, VT_KB, VT_STAGES><<<grid, VT_THREADS, shm, ctx->stream>>>(mIn, Out, ldo, n, kin, nout, Q, ldq, qtrans, alpha, beta);
  else if (nout <= 48) k_vq_tma<4, 3, VT_KB, VT_STAGES><<<grid, VT_THREADS, shm, ctx->stream>>>(mIn, Out, ldo, n, kin, nout, Q, ldq, qtrans, alpha, beta);
  else k_vq_tma<8, 2, VT_KB, VT_STAGES><<<grid, VT_THREADS, shm, ctx->stream>>>(mIn, Out, ldo, n, kin, nout, Q, ldq, qtrans, alpha, beta);
  PROF_END(ctx);
  CKLAUNCH(ctx);
  return B2K_OK;
}
