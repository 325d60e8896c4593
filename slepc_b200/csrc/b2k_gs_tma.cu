/*
 * b2k_gs_tma.cu — Gram-Schmidt update sweep with a TMA tensor-map pipeline (sm_100a).
 *
 *     w <- beta*w + alpha * V(:,0:k) q                       (BVMultVec / BVMultColumn, bvcuda.cu:45-60)
 *     optional, from the SAME read of V:
 *       cout[0:k] = V(:,0:k)^T w_new                         (dot sweep of the next DGKS pass, bvorthog.c:100)
 *       cout[k]   = ||w_new||^2                              (explicit norm, bvorthog.c:126)
 *
 * Same arithmetic and the same register tile as k_gs_rt (b2k_gs_rt.cu), different data movement: there the
 * latency of HBM is hidden by 16 warps per SM that each hold 16 loads in flight, and the next tile cannot be
 * requested before the dot products of the current one have retired its registers (ncu: 70 % of DRAM peak for
 * the DOT variant, long-scoreboard + barrier stalls).  Here ONE producer thread per SM keeps a ring of 3..8
 * shared-memory stages full with `cp.async.bulk.tensor.2d` copies — one instruction moves a 128-row x KB-column
 * box of V (up to 64 KB; the earlier 1-D bulk-copy variant needed 64 copies of 1 KB per tile and was bound by the
 * copy issue rate) and one more the 128 rows of w — so 130-200 KB per SM are always in flight no matter what the
 * consumer warps are doing.  Out-of-bounds rows (tail tile) and columns (k < KB) are zero-filled by the TMA unit
 * without touching DRAM, so there is no tail code.  8 consumer warps = 2 row halves x 4 column groups pull their
 * 64-row x CPT-column sub-tile into registers with conflict-free LDS.128 and release the stage immediately.
 * Reductions are fixed-order (bit-reproducible): lanes by shuffle, row halves through shared memory, CTAs by
 * k_reduce_partials.
 */
#include <cuda.h>
#include <stdlib.h>
#include "b2k_internal.h"

#define TM_ROWS      128
#define TM_CONS      256                 /* consumer threads (8 warps)   */
#define TM_THREADS   (TM_CONS + 32)      /* + producer warp              */
#define TM_MAXSTAGES 8

__device__ __forceinline__ uint32_t tm_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tm_mbar_init(uint32_t bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void tm_mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tm_mbar_arrive(uint32_t bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t tm_mbar_try_wait(uint32_t bar, uint32_t parity)
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
__device__ __forceinline__ void tm_mbar_wait(uint32_t bar, uint32_t parity)
{
  while (!tm_mbar_try_wait(bar, parity)) { }
}
__device__ __forceinline__ void tm_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar)
{
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tm_half_sync(int rh) { asm volatile("bar.sync %0, 128;" ::"r"(1 + rh) : "memory"); }

struct TmSmem {
  unsigned long long full[TM_MAXSTAGES];
  unsigned long long empty[TM_MAXSTAGES];
  double  qs[64];
  double2 psum[2][2][4][32];
  double  red[2][66];
};
#define TM_FIXED (((sizeof(TmSmem) + 1023) / 1024) * 1024)

template <int CPT, bool DOT, bool NRM>
__global__ void __launch_bounds__(TM_THREADS, 1)
k_gs_tma(const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmW, int64_t n, int k, double alpha, double beta,
         double *__restrict__ w, const double *__restrict__ q, double *__restrict__ part, int pstride, int nstages, const b2k_gate_s gate)
{
  b2k_pdl_enter();
  if (b2k_gate_closed(gate)) return;                            /* DGKS does not refine: the whole grid leaves (uniform) */
  constexpr int KB = 4 * CPT;                                   /* columns per stage (box width)   */
  constexpr int STAGE_DOUBLES = (KB + 1) * TM_ROWS;            /* V box + w box                   */
  extern __shared__ __align__(1024) unsigned char tm_raw[];
  TmSmem *S = reinterpret_cast<TmSmem *>(tm_raw);
  double *stages = reinterpret_cast<double *>(tm_raw + TM_FIXED);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (n + TM_ROWS - 1) / TM_ROWS;
  const bool have_w = (beta != 0.0);

  if (tid == 0) {
    for (int s = 0; s < nstages; s++) {
      tm_mbar_init(tm_smem_u32(&S->full[s]), 1);
      tm_mbar_init(tm_smem_u32(&S->empty[s]), 8);               /* one arrival per consumer warp */
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 64) S->qs[tid] = (tid < k) ? q[tid] : 0.0;
  __syncthreads();

  if (warp == 8) {
    /* ================= producer: one thread feeds the ring ================= */
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmV)) : "memory");
      if (have_w) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
      const uint32_t bytes = (uint32_t)((have_w ? KB + 1 : KB) * TM_ROWS * sizeof(double));
      int s = 0;
      uint32_t ph = 0;
      for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        tm_mbar_wait(tm_smem_u32(&S->empty[s]), ph ^ 1);        /* slot free (first round passes) */
        double *st = stages + (size_t)s * STAGE_DOUBLES;
        const uint32_t bar = tm_smem_u32(&S->full[s]);
        tm_mbar_expect_tx(bar, bytes);
        tm_load_2d(tm_smem_u32(st), &tmV, (int)(t * TM_ROWS), 0, bar);
        if (have_w) tm_load_2d(tm_smem_u32(st + (size_t)KB * TM_ROWS), &tmW, (int)(t * TM_ROWS), 0, bar);
        if (++s == nstages) { s = 0; ph ^= 1; }
      }
    }
    return;
  }

  /* ================= consumers: warp = (row half rh, column group cg) ================= */
  const int rh = warp >> 2, cg = warp & 3;
  double acc[CPT];
#pragma unroll
  for (int u = 0; u < CPT; u++) acc[u] = 0.0;
  double nrm = 0.0;
  int s = 0, par = 0;
  uint32_t ph = 0;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, par ^= 1) {
    const int64_t row = t * TM_ROWS + rh * 64 + 2 * lane;
    tm_mbar_wait(tm_smem_u32(&S->full[s]), ph);
    const double *tile = stages + (size_t)s * STAGE_DOUBLES;
    const double *tp = tile + (size_t)(cg * CPT) * TM_ROWS + rh * 64 + 2 * lane;
    double2 v[CPT];
#pragma unroll
    for (int u = 0; u < CPT; u++) v[u] = *reinterpret_cast<const double2 *>(tp + (size_t)u * TM_ROWS);
    double2 wv = make_double2(0.0, 0.0);
    if (have_w) wv = *reinterpret_cast<const double2 *>(tile + (size_t)KB * TM_ROWS + rh * 64 + 2 * lane);
    double ax = 0.0, ay = 0.0;
#pragma unroll
    for (int u = 0; u < CPT; u++) {
      const double c = S->qs[cg * CPT + u];                     /* warp-wide broadcast */
      ax = fma(v[u].x, c, ax);
      ay = fma(v[u].y, c, ay);
    }
    S->psum[par][rh][cg][lane] = make_double2(ax, ay);          
    __syncwarp();
    if (lane == 0) tm_mbar_arrive(tm_smem_u32(&S->empty[s]));    /* the tile now lives in registers         */
    if (++s == nstages) { s = 0; ph ^= 1; }
    tm_half_sync(rh);
    const double2 p0 = S->psum[par][rh][0][lane], p1 = S->psum[par][rh][1][lane], p2 = S->psum[par][rh][2][lane],
                  p3 = S->psum[par][rh][3][lane];
    const double sx = (p0.x + p1.x) + (p2.x + p3.x), sy = (p0.y + p1.y) + (p2.y + p3.y);
    double2 wn;
    if (!have_w) { wn.x = alpha * sx; wn.y = alpha * sy; }
    else { wn.x = fma(beta, wv.x, alpha * sx); wn.y = fma(beta, wv.y, alpha * sy); }
    if (cg == 0) {
      if (row + 1 < n) *reinterpret_cast<double2 *>(w + row) = wn;
      else if (row < n) w[row] = wn.x;                          /* rows >= n were zero-filled: wn is 0 there */
      if (NRM) { nrm = fma(wn.x, wn.x, nrm); nrm = fma(wn.y, wn.y, nrm); }
    }
    if (DOT) {
#pragma unroll
      for (int u = 0; u < CPT; u++) acc[u] = fma(v[u].x, wn.x, fma(v[u].y, wn.y, acc[u]));
    }
  }

  /* ---- CTA reduction: lanes (shuffle), then the two row halves (shared memory, fixed order) ---- */
  if (DOT) {
#pragma unroll
    for (int u = 0; u < CPT; u++) {
      double sres = acc[u];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sres += __shfl_xor_sync(0xffffffffu, sres, o);
      if (lane == 0) S->red[rh][cg * CPT + u] = sres;
    }
  }
  if (NRM && cg == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
    if (lane == 0) S->red[rh][64] = nrm;
  }
  asm volatile("bar.sync 3, 256;" ::: "memory");
  if (DOT && tid < k) part[(int64_t)blockIdx.x * pstride + tid] = S->red[0][tid] + S->red[1][tid];
  if (NRM && tid == 64) part[(int64_t)blockIdx.x * pstride + (DOT ? k : 0)] = S->red[0][64] + S->red[1][64];
}

/* ---- host side: tensor maps through the driver entry point (no link-time dependency on libcuda) ---- */
typedef CUresult (*tm_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tm_encode_fn g_encode = nullptr;
static int g_encode_tried = 0;

static tm_encode_fn tm_get_encode(void)
{
  if (!g_encode_tried) {
    g_encode_tried = 1;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      g_encode = (tm_encode_fn)fn;
    else
      cudaGetLastError();
  }
  return g_encode;
}

int b2k_tm_make_map(CUtensorMap *map, const double *base, int64_t n, int64_t ncols, int64_t ld, int box_cols, int box_rows)
{
  tm_encode_fn enc = tm_get_encode();
  if (!enc) return -1;
  const cuuint64_t gdim[2] = {(cuuint64_t)n, (cuuint64_t)ncols};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t)box_rows, (cuuint32_t)box_cols};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -1;
}

int b2k_launch_reduce_partials(b2k_ctx ctx, int nblk, int pstride, int ncols, double *out);   /* b2k_bv.cu */

template <int CPT>
static int tm_launch(b2k_ctx ctx, const CUtensorMap &mV, const CUtensorMap &mW, int64_t n, int k, double alpha, double beta, double *w,
                     const double *q, int dot, int nrm, int grid, int pstride, const b2k_gate_s &gate)
{
  constexpr int KB = 4 * CPT;
  const size_t stage_bytes = (size_t)(KB + 1) * TM_ROWS * sizeof(double);
  const size_t budget = 224 * 1024;
  int nstages = (int)((budget - TM_FIXED) / stage_bytes);
  if (nstages > TM_MAXSTAGES) nstages = TM_MAXSTAGES;
  const size_t shm = TM_FIXED + (size_t)nstages * stage_bytes;
#define TM_GO(D, N)                                                                                                             \
  do {                                                                                                                          \
    static int configured = 0;                                                                                                  \
    if (!configured) {                                                                                                          \
      CK(cudaFuncSetAttribute(k_gs_tma<CPT, D, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));                  \
      configured = 1;                                                                                                           \
    }                                                                                                                           \
    b2k_launch_pdl(k_gs_tma<CPT, D, N>, dim3(grid), dim3(TM_THREADS), shm, ctx->stream, mV, mW, n, k, alpha, beta, w, q, ctx->partials, pstride, nstages, gate); \
  } while (0)
  if (dot) TM_GO(true, true);
  else if (nrm) TM_GO(false, true);
  else TM_GO(false, false);
#undef TM_GO
  return 0;
}

/* returns -1 when the shape is not supported (k > 64, small or unaligned blocks, no driver entry point): the caller
   falls back to the register-tile kernel.  Semantics of b2k_gs_rt_launch. */
int b2k_gs_tma_launch(b2k_ctx ctx, const double *V, int64_t ld, int64_t n, int k, double alpha, double beta, double *w, const double *q,
                      int dot, double *out, const b2k_gate_s *gatep)
{
  const b2k_gate_s gate = gatep ? *gatep : b2k_gate_s{nullptr, nullptr, 0.0};
  if (k < 5 || k > 64 || n < 32 * TM_ROWS || n >= 2147483647LL - TM_ROWS) return -1;   /* k <= 4: the register tile wins */
  if (!b2k_is_aligned16(V) || !b2k_is_aligned16(w) || (ld & 1)) return -1;
  CUtensorMap mV, mW;
  const int cpt = (k <= 16) ? 4 : (k <= 32 ? 8 : (k <= 48 ? 12 : 16));
  if (b2k_tm_make_map(&mV, V, n, k, ld, 4 * cpt, TM_ROWS)) return -1;
  if (b2k_tm_make_map(&mW, w, n, 1, ld, 1, TM_ROWS)) return -1;
  const int nrm = out != nullptr;
  const int64_t ntiles = (n + TM_ROWS - 1) / TM_ROWS;
  int grid = ctx->sm_count;
  if ((int64_t)grid > ntiles) grid = (int)ntiles;
  const int pstride = dot ? k + 1 : 1;
  PROF_BEGIN(ctx, dot ? B2K_PROF_GSFUSED : B2K_PROF_MULTVEC, 8.0 * (double)n * (k + (beta == 0.0 ? 1 : 2)));
  int rc;
  switch (cpt) {
    case 4: rc = tm_launch<4>(ctx, mV, mW, n, k, alpha, beta, w, q, dot, nrm, grid, pstride, gate); break;
    case 8: rc = tm_launch<8>(ctx, mV, mW, n, k, alpha, beta, w, q, dot, nrm, grid, pstride, gate); break;
    case 12: rc = tm_launch<12>(ctx, mV, mW, n, k, alpha, beta, w, q, dot, nrm, grid, pstride, gate); break;
    default: rc = tm_launch<16>(ctx, mV, mW, n, k, alpha, beta, w, q, dot, nrm, grid, pstride, gate); break;
  }
  PROF_END(ctx);
  if (rc) return rc;
  CKLAUNCH(ctx);
  if (dot) return b2k_launch_reduce_partials(ctx, grid, pstride, k + 1, out);
  if (nrm) return b2k_launch_reduce_partials(ctx, grid, 1, 1, out);
  return B2K_OK;
}
