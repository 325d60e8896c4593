/*
 * b2k_vq_tma.cu — restart GEMM  Out(n x nout) = beta*Out + alpha * In(n x kin) * Qe(kin x nout), kin, nout <= 64, Out may alias
 * columns of In (BVMultInPlace, bvops.c:220-245; replaces BVMultInPlace_BLAS_CUDA bvcuda.cu:65-112), as a TMA-fed
 * FP64 tensor-core kernel.
 *
 * ncu on k_vq (profiles/r01_ncu_summary.md): FP64 pipe 44 %, DRAM 36 % — the shape sits on the FP64 ridge (AI 5.3 flop/B)
 * and the load / DFMA / store phases of a CTA do not overlap; the DFMA inner loop is also at the shared-memory limit
 * (one LDS.128 of the tile + C/2 broadcast LDS.128 of Q per 2C DFMA).  Here:
 *   - a producer thread streams 128-row x 64-column boxes of In through a 3-stage shared-memory ring with
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
#define VT_KB      64
#define VT_STAGES  3
#define VT_THREADS 288

int b2k_tm_make_map(CUtensorMap *map, const double *base, int64_t n, int64_t ncols, int64_t ld, int box_cols);   /* b2k_gs_tma.cu */

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

template <int NS>    /* 8-column slabs per warp: 1 (nout <= 32) or 2 (nout <= 64) */
__global__ void __launch_bounds__(VT_THREADS, 1)
k_vq_tma(const __grid_constant__ CUtensorMap tmIn, double *Out, int64_t ldo, int64_t n, int kin, int nout, const double *__restrict__ Q,
         int ldq, int qtrans, double alpha, double beta)
{
  extern __shared__ __align__(1024) unsigned char vt_raw[];
  unsigned long long *full = reinterpret_cast<unsigned long long *>(vt_raw);
  unsigned long long *empty = full + VT_STAGES;
  double *stages = reinterpret_cast<double *>(vt_raw + 1024);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (n + VT_ROWS - 1) / VT_ROWS;
  if (tid == 0) {
    for (int s = 0; s < VT_STAGES; s++) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(vt_smem_u32(&full[s])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(vt_smem_u32(&empty[s])), "r"(8));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == 8) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmIn)) : "memory");
      int s = 0;
      uint32_t ph = 0;
      for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        vt_wait(vt_smem_u32(&empty[s]), ph ^ 1);
        const uint32_t bar = vt_smem_u32(&full[s]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(VT_KB * VT_ROWS * sizeof(double))) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                         vt_smem_u32(stages + (size_t)s * VT_KB * VT_ROWS)),
                     "l"(reinterpret_cast<uint64_t>(&tmIn)), "r"((int)(t * VT_ROWS)), "r"(0), "r"(bar)
                     : "memory");
        if (++s == VT_STAGES) { s = 0; ph ^= 1; }
      }
    }
    return;
  }

  /* consumers: warp = (row half rh, column group cg); fragment coordinates of m8n8k4: */
  const int rh = warp >> 2, cg = warp & 3;
  const int fr = lane >> 2, fk = lane & 3;        /* A: row fr, k fk;  B: k fk, col fr;  C: row fr, cols 2*fk, 2*fk+1 */
  const int ksteps = (kin + 3) >> 2;
  double bq[16][NS];
#pragma unroll
  for (int ks = 0; ks < 16; ks++) {
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
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    vt_wait(vt_smem_u32(&full[s]), ph);
    const double *tile = stages + (size_t)s * VT_KB * VT_ROWS;
    const int64_t rbase = t * VT_ROWS + rh * 64;
    /* 4 m-tiles (32 rows) at a time: 4*NS independent accumulator chains keep the FP64 tensor pipe busy (a single chain is
       bound by the DMMA latency) */
#pragma unroll 1
    for (int mg = 0; mg < 2; mg++) {
      const double *ap = tile + (size_t)fk * VT_ROWS + rh * 64 + mg * 32 + fr;
      const int64_t row = rbase + mg * 32 + fr;                  /* + 8*m for m-tile m */
      double c0[4][NS], c1[4][NS];
#pragma unroll
      for (int m = 0; m < 4; m++) {
#pragma unroll
        for (int sl = 0; sl < NS; sl++) {
          c0[m][sl] = 0.0; c1[m][sl] = 0.0;
          if (beta != 0.0 && row + 8 * m < n) {
            const int col = (cg * NS + sl) * 8 + 2 * fk;
            if (col < nout) c0[m][sl] = beta * Out[(int64_t)col * ldo + row + 8 * m];
            if (col + 1 < nout) c1[m][sl] = beta * Out[(int64_t)(col + 1) * ldo + row + 8 * m];
          }
        }
      }
#pragma unroll
      for (int ks = 0; ks < 16; ks++) {
        if (ks < ksteps) {
          double a[4];
#pragma unroll
          for (int m = 0; m < 4; m++) a[m] = ap[(size_t)(4 * ks) * VT_ROWS + 8 * m];
#pragma unroll
          for (int m = 0; m < 4; m++) {
#pragma unroll
            for (int sl = 0; sl < NS; sl++) vt_dmma(c0[m][sl], c1[m][sl], a[m], bq[ks][sl]);
          }
        }
      }
      if (mg == 1) {                                /* every read of this stage has been issued and consumed */
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(vt_smem_u32(&empty[s])) : "memory");
      }
#pragma unroll
      for (int m = 0; m < 4; m++) {
        if (row + 8 * m < n) {
#pragma unroll
          for (int sl = 0; sl < NS; sl++) {
            const int col = (cg * NS + sl) * 8 + 2 * fk;
            if (col < nout) Out[(int64_t)col * ldo + row + 8 * m] = c0[m][sl];
            if (col + 1 < nout) Out[(int64_t)(col + 1) * ldo + row + 8 * m] = c1[m][sl];
          }
        }
      }
    }
    if (++s == VT_STAGES) { s = 0; ph ^= 1; }
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
  if (!b2k_is_aligned16(In) || (ldi & 1)) return -1;
  /* in place is safe only when Out's rows are In's rows: same leading dimension, Out a column of In's block */
  CUtensorMap mIn;
  if (b2k_tm_make_map(&mIn, In, n, kin, ldi, VT_KB)) return -1;
  const size_t shm = 1024 + (size_t)VT_STAGES * VT_KB * VT_ROWS * sizeof(double);
  const int64_t ntiles = (n + VT_ROWS - 1) / VT_ROWS;
  int grid = ctx->sm_count;
  if ((int64_t)grid > ntiles) grid = (int)ntiles;
  static int configured = 0;
  if (!configured) {
    CK(cudaFuncSetAttribute(k_vq_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
    CK(cudaFuncSetAttribute(k_vq_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
    configured = 1;
  }
  PROF_BEGIN(ctx, B2K_PROF_GEMM, 8.0 * (double)n * (kin + nout));
  if (nout <= 32) k_vq_tma<1><<<grid, VT_THREADS, shm, ctx->stream>>>(mIn, Out, ldo, n, kin, nout, Q, ldq, qtrans, alpha, beta);
  else k_vq_tma<2><<<grid, VT_THREADS, shm, ctx->stream>>>(mIn, Out, ldo, n, kin, nout, Q, ldq, qtrans, alpha, beta);
  PROF_END(ctx);
  CKLAUNCH(ctx);
  return B2K_OK;
}
