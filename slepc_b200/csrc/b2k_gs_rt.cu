/*
 * b2k_gs_rt.cu — register-tile Gram-Schmidt update sweep for sm_100a.
 *
 *     w <- beta*w + alpha * V(:,0:k) q                       (BVMultVec / BVMultColumn, bvcuda.cu:45-60)
 *     optional, from the SAME read of V:
 *       cout[0:k] = V(:,0:k)^T w_new                         (dot sweep of the next DGKS pass, bvorthog.c:100)
 *       cout[k]   = ||w_new||^2                              (explicit norm, bvorthog.c:126)
 *
 * Why registers and not shared memory / TMA staging: the first TMA-staged variant (round 1, since removed: one 1 KB bulk
 * copy per column per 128-row tile) is bound by the per-copy issue rate of the bulk-copy engine — ncu shows
 * 40 % of the warp samples spinning on the `full` mbarrier and 50 % DRAM utilisation (profiles/r01_ncu_summary.md).
 * Here a CTA of 4 warps owns a 64-row tile; warp g holds columns [g*kq,(g+1)*kq) of the tile in REGISTERS
 * (one 16-byte load per column per lane, all issued back to back: up to 16 independent loads in flight per
 * thread, the pattern that lets k_dotvec reach 87 % of DRAM peak), computes its partial of V q, the four
 * partials meet in shared memory (one __syncthreads per tile, double-buffered), every warp forms w_new and
 * immediately re-uses its registers for the inner products.  V is read from HBM exactly once, nothing is
 * written but w.  Reductions are fixed-order (bit-reproducible): lanes by shuffle, CTAs by k_reduce_partials.
 */
#include "b2k_internal.h"

#define RT_THREADS 128
#define RT_ROWS    64          /* rows per tile: 32 lanes x one row pair */

__device__ __forceinline__ double2 rt_ld2(const double *p)
{
  double2 r;
  asm("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));   /* not volatile: free to batch */
  return r;
}
__device__ __forceinline__ double rt_ld1(const double *p)
{
  double r;
  asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
}

template <int CPT, bool DOT, bool NRM>
__global__ void __launch_bounds__(RT_THREADS, (CPT >= 16 ? 4 : (CPT >= 8 ? 5 : 6)))
k_gs_rt(const double *__restrict__ V, int64_t ld, int64_t n, int k, double alpha, double beta, double *__restrict__ w,
        const double *__restrict__ q, double *__restrict__ part, int pstride, const b2k_gate_s gate)
{
  b2k_pdl_enter();
  if (b2k_gate_closed(gate)) return;              /* DGKS does not refine: the whole grid leaves (uniform) */
  __shared__ double  qs[4 * CPT];
  __shared__ double2 psum[2][4][32];
  const int lane = threadIdx.x & 31, cg = threadIdx.x >> 5;
  const int kq = (k + 3) >> 2;                    /* columns per warp (<= CPT) */
  const int c0 = cg * kq;
  const int nc = max(0, min(kq, k - c0));
  for (int i = threadIdx.x; i < 4 * CPT; i += RT_THREADS) {
    const int g = i / CPT, u = i - g * CPT, c = g * kq + u;
    qs[i] = (u < kq && c < k) ? q[c] : 0.0;
  }
  __syncthreads();
  double acc[CPT];
#pragma unroll
  for (int u = 0; u < CPT; u++) acc[u] = 0.0;
  double nrm = 0.0;
  const double *Vc = V + (int64_t)c0 * ld;
  const int64_t ntiles = (n + RT_ROWS - 1) / RT_ROWS;
  int par = 0;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, par ^= 1) {
    const int64_t row = t * RT_ROWS + 2 * lane;
    const bool full = row + 1 < n, half = row < n;
    double2 v[CPT];
#pragma unroll
    for (int u = 0; u < CPT; u++) {
      v[u] = make_double2(0.0, 0.0);
      if (u < nc) {
        const double *p = Vc + (int64_t)u * ld + row;
        if (full) v[u] = rt_ld2(p);
        else if (half) v[u].x = rt_ld1(p);
      }
    }
    double2 wv = make_double2(0.0, 0.0);
    if (beta != 0.0) {
      if (full) wv = *reinterpret_cast<const double2 *>(w + row);
      else if (half) wv.x = w[row];
    }
    double ax = 0.0, ay = 0.0;
#pragma unroll
    for (int u = 0; u < CPT; u++) {
      const double c = qs[cg * CPT + u];          /* same address for the whole warp: broadcast */
      ax = fma(v[u].x, c, ax);
      ay = fma(v[u].y, c, ay);
    }
    psum[par][cg][lane] = make_double2(ax, ay);
    __syncthreads();
    const double2 p0 = psum[par][0][lane], p1 = psum[par][1][lane], p2 = psum[par][2][lane], p3 = psum[par][3][lane];
    const double sx = (p0.x + p1.x) + (p2.x + p3.x), sy = (p0.y + p1.y) + (p2.y + p3.y);
    double2 wn;
    if (beta == 0.0) { wn.x = alpha * sx; wn.y = alpha * sy; }
    else { wn.x = fma(beta, wv.x, alpha * sx); wn.y = fma(beta, wv.y, alpha * sy); }
    if (!half) { wn.x = 0.0; wn.y = 0.0; }
    else if (!full) wn.y = 0.0;
    if (cg == 0) {
      if (full) *reinterpret_cast<double2 *>(w + row) = wn;
      else if (half) w[row] = wn.x;
      if (NRM) { nrm = fma(wn.x, wn.x, nrm); nrm = fma(wn.y, wn.y, nrm); }
    }
    if (DOT) {
#pragma unroll
      for (int u = 0; u < CPT; u++) acc[u] = fma(v[u].x, wn.x, fma(v[u].y, wn.y, acc[u]));
    }
  }
  if (DOT) {
#pragma unroll
    for (int u = 0; u < CPT; u++) {
      double s = acc[u];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0 && u < nc) part[(int64_t)blockIdx.x * pstride + c0 + u] = s;
    }
  }
  if (NRM && cg == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
    if (lane == 0) part[(int64_t)blockIdx.x * pstride + (DOT ? k : 0)] = nrm;
  }
}

int b2k_launch_reduce_partials(b2k_ctx ctx, int nblk, int pstride, int ncols, double *out);   /* b2k_bv.cu */

/* returns -1 when the shape is not supported (k > 64, unaligned): the caller falls back to the generic kernels.
   cout: DOT ? k+1 values (V^T w_new, ||w_new||^2) : (nrm_out ? 1 value ||w_new||^2 : nothing) */
int b2k_gs_rt_launch(b2k_ctx ctx, const double *V, int64_t ld, int64_t n, int k, double alpha, double beta, double *w,
                     const double *q, int dot, double *out, const b2k_gate_s *gatep)
{
  const b2k_gate_s gate = gatep ? *gatep : b2k_gate_s{nullptr, nullptr, 0.0};
  if (k < 1 || k > 64 || n < 1) return -1;
  if (!b2k_is_aligned16(V) || !b2k_is_aligned16(w) || (ld & 1)) return -1;
  const int kq = (k + 3) >> 2;
  const int nrm = out != nullptr;
  const int64_t ntiles = (n + RT_ROWS - 1) / RT_ROWS;
  const int per_sm = (kq > 8) ? 4 : (kq > 4 ? 5 : 6);
  int grid = ctx->sm_count * per_sm;
  if ((int64_t)grid > ntiles) grid = (int)ntiles;
  if (grid > B2K_MAX_PART_BLOCKS) grid = B2K_MAX_PART_BLOCKS;
  const int pstride = dot ? k + 1 : 1;
  PROF_BEGIN(ctx, dot ? B2K_PROF_GSFUSED : B2K_PROF_MULTVEC, 8.0 * (double)n * (k + (beta == 0.0 ? 1 : 2)));
#define RT_LAUNCH(CPT)                                                                                                          \
  do {                                                                                                                          \
    if (dot) b2k_launch_pdl(k_gs_rt<CPT, true, true>, dim3(grid), dim3(RT_THREADS), 0, ctx->stream, V, ld, n, k, alpha, beta, w, q, ctx->partials, pstride, gate); \
    else if (nrm) b2k_launch_pdl(k_gs_rt<CPT, false, true>, dim3(grid), dim3(RT_THREADS), 0, ctx->stream, V, ld, n, k, alpha, beta, w, q, ctx->partials, pstride, gate); \
    else b2k_launch_pdl(k_gs_rt<CPT, false, false>, dim3(grid), dim3(RT_THREADS), 0, ctx->stream, V, ld, n, k, alpha, beta, w, q, ctx->partials, pstride, gate); \
  } while (0)
  if (kq <= 4) RT_LAUNCH(4);
  else if (kq <= 8) RT_LAUNCH(8);
  else RT_LAUNCH(16);
#undef RT_LAUNCH
  PROF_END(ctx);
  CKLAUNCH(ctx);
  if (dot) return b2k_launch_reduce_partials(ctx, grid, pstride, k + 1, out);
  if (nrm) return b2k_launch_reduce_partials(ctx, grid, 1, 1, out);
  return B2K_OK;
}
