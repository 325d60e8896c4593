/*
 * b2k_tsqr.cu — tall-skinny QR of the active columns of a basis resident in HBM: BVOrthogonalize with BV_ORTHOG_BLOCK_TSQR and
 * BV_ORTHOG_BLOCK_TSQRCHOL (bvorthog.c:611-656).  The reference runs LAPACK geqrf/orgqr on the host copy of every rank's row
 * block and combines the ranks' triangles in a binary tree (bvlapack.c:347-472) or with a packed-Givens MPI reduction
 * (bvlapack.c:474-560); its CUDA BV has no device version (the raw array comes to the host).
 *
 * Here the tree has three levels, all Householder (unconditionally stable, unlike the Gram-matrix methods CHOL / SVQB):
 *   1. inside a CTA, a FLAT tree over tiles of 128 rows: [R ; tile] -> [R' ; 0] by k reflectors whose only non-trivial part lies
 *      in the tile (the structure of LAPACK's tpqrt with a rectangular pentagon), so the reflectors are stored IN PLACE of the
 *      tile, plus two scalars (tau, 1/(alpha-beta)) per reflector;
 *   2. over the CTAs of a GPU: the k x k triangles go to the host (<= 1184 blocks), one LAPACK geqrf/orgqr of the stack;
 *   3. over the GPUs: the same on the gathered per-rank triangles (host/bv.c).
 * Q is formed by the backward kernel: every CTA starts from its k x k block W of the upper levels' orthogonal factor and applies
 * its tiles' reflectors in reverse order; each warp carries its columns of Q through all k reflectors in registers, no block
 * barrier inside a tile.  Traffic: forward reads (and, when Q is wanted, writes) n x k once, backward reads and writes it once.
 * TSQRCHOL needs only R: the forward kernel then leaves V untouched and Q = V R^-1 is the restart GEMM (k_vq_tma).
 */
#include "b2k_internal.h"

#define TS_ROWS 128
#define TS_THREADS 256
#define TS_WARPS 8

__device__ __forceinline__ double ts_warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int KP>
__host__ __device__ constexpr size_t ts_fwd_smem() { return sizeof(double) * ((size_t)KP * TS_ROWS + (size_t)KP * KP + 2 * KP); }
template <int KP>
__host__ __device__ constexpr size_t ts_bwd_smem() { return sizeof(double) * ((size_t)KP * TS_ROWS + (size_t)KP * KP + 2 * KP); }

/* forward: R of the rows [blockIdx.x*rpc, +rpc) into Rblk[blockIdx.x] (k x k column-major, upper triangular); STORE: reflectors
   over V and (tau, s) into coef[tile][2][k] */
template <int KP, bool STORE>
__global__ void __launch_bounds__(TS_THREADS) k_tsqr_fwd(double *__restrict__ V, int64_t ld, int64_t n, int k, int64_t rpc,
                                                           double *__restrict__ Rblk, double *__restrict__ coef)
{
  extern __shared__ double ts_sm[];
  double *A = ts_sm;                         /* [KP][128] tile, column-major                                   */
  double *Rs = A + KP * TS_ROWS;             /* [KP][KP] strict upper triangle, row j at Rs + j*KP              */
  double *Rd = Rs + KP * KP;                 /* [2][KP] diagonal, double-buffered by tile parity                */
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r0 = (int64_t)blockIdx.x * rpc, r1 = min(n, r0 + rpc);
  for (int i = threadIdx.x; i < KP * KP + 2 * KP; i += TS_THREADS) Rs[i] = 0.0;
  int par = 0;
  for (int64_t tr = r0; tr < r1; tr += TS_ROWS) {
    __syncthreads();
    for (int c = warp; c < k; c += TS_WARPS) {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int64_t r = tr + lane + 32 * i;
        A[c * TS_ROWS + lane + 32 * i] = (r < r1) ? V[(int64_t)c * ld + r] : 0.0;
      }
    }
    __syncthreads();
    const double *Rcur = Rd + par * KP;
    double *Rnew = Rd + (par ^ 1) * KP;
    const int64_t tile = tr / TS_ROWS;
    for (int j = 0; j < k; j++) {
      double x[4];
#pragma unroll
      for (int i = 0; i < 4; i++) x[i] = A[j * TS_ROWS + lane + 32 * i];
      const double ss = ts_warp_sum(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);   /* same value in every warp */
      const double alpha = Rcur[j];
      double beta = alpha, tau = 0.0, s = 0.0;
      if (ss > 0.0) {
        beta = -copysign(sqrt(alpha * alpha + ss), alpha);
        tau = (beta - alpha) / beta;
        s = 1.0 / (alpha - beta);
      }
      if (warp == (j & (TS_WARPS - 1)) && lane == 0) {
        Rnew[j] = beta;
        if (STORE) { coef[tile * 2 * k + j] = tau; coef[tile * 2 * k + k + j] = s; }
      }
      if (tau != 0.0) {
        for (int c = j + 1 + warp; c < k; c += TS_WARPS) {
          double a[4];
#pragma unroll
          for (int i = 0; i < 4; i++) a[i] = A[c * TS_ROWS + lane + 32 * i];
          const double dot = ts_warp_sum(x[0] * a[0] + x[1] * a[1] + x[2] * a[2] + x[3] * a[3]);
          const double tw = tau * (Rs[j * KP + c] + s * dot);
          __syncwarp();
          if (lane == 0) Rs[j * KP + c] -= tw;
          const double f = tw * s;
#pragma unroll
          for (int i = 0; i < 4; i++) A[c * TS_ROWS + lane + 32 * i] = a[i] - f * x[i];
        }
      }
      __syncthreads();
    }
    if (STORE) {
      for (int c = warp; c < k; c += TS_WARPS) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int64_t r = tr + lane + 32 * i;
          if (r < r1) V[(int64_t)c * ld + r] = A[c * TS_ROWS + lane + 32 * i];
        }
      }
    }
    par ^= 1;
  }
  __syncthreads();
  double *out = Rblk + (size_t)blockIdx.x * k * k;
  const double *Rfin = Rd + par * KP;
  for (int e = threadIdx.x; e < k * k; e += TS_THREADS) {
    const int i = e % k, c = e / k;                   /* row i, column c, column-major */
    out[e] = (i < c) ? Rs[i * KP + c] : (i == c ? Rfin[i] : 0.0);
  }
}

/* backward: rows of CTA b <- Q_b * W_b, Q_b = product of the tiles' reflectors stored by the forward kernel */
template <int KP>
__global__ void __launch_bounds__(TS_THREADS) k_tsqr_bwd(double *__restrict__ V, int64_t ld, int64_t n, int k, int64_t rpc,
                                                           const double *__restrict__ Wblk, const double *__restrict__ coef)
{
  extern __shared__ double ts_sm[];
  double *X = ts_sm;                         /* [KP][128] reflectors of the tile                               */
  double *Ct = X + KP * TS_ROWS;             /* [KP][KP] top block, column c at Ct + c*KP                        */
  double *tau = Ct + KP * KP, *sc = tau + KP;
  constexpr int NC = KP / TS_WARPS;          /* columns a warp carries                                            */
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r0 = (int64_t)blockIdx.x * rpc, r1 = min(n, r0 + rpc);
  if (r0 >= r1) return;
  const double *W = Wblk + (size_t)blockIdx.x * k * k;
  for (int e = threadIdx.x; e < k * k; e += TS_THREADS) Ct[(e / k) * KP + (e % k)] = W[e];
  const int64_t ntile = (r1 - r0 + TS_ROWS - 1) / TS_ROWS;
  for (int64_t t = ntile - 1; t >= 0; t--) {
    const int64_t tr = r0 + t * TS_ROWS, tile = tr / TS_ROWS;
    __syncthreads();
    for (int c = warp; c < k; c += TS_WARPS) {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int64_t r = tr + lane + 32 * i;
        X[c * TS_ROWS + lane + 32 * i] = (r < r1) ? V[(int64_t)c * ld + r] : 0.0;
      }
    }
    if (threadIdx.x < 2 * k) tau[(threadIdx.x < k) ? threadIdx.x : KP + (threadIdx.x - k)] = coef[tile * 2 * k + threadIdx.x];
    __syncthreads();
    double q[NC][4];
#pragma unroll
    for (int m = 0; m < NC; m++) { q[m][0] = q[m][1] = q[m][2] = q[m][3] = 0.0; }
    for (int j = k - 1; j >= 0; j--) {
      const double tj = tau[j], sj = sc[j];
      if (tj == 0.0) continue;
      double x[4];
#pragma unroll
      for (int i = 0; i < 4; i++) x[i] = X[j * TS_ROWS + lane + 32 * i];
      double dot[NC];
#pragma unroll
      for (int m = 0; m < NC; m++) dot[m] = x[0] * q[m][0] + x[1] * q[m][1] + x[2] * q[m][2] + x[3] * q[m][3];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int m = 0; m < NC; m++) dot[m] += __shfl_xor_sync(0xffffffffu, dot[m], o);
      }
      double tw[NC];
#pragma unroll
      for (int m = 0; m < NC; m++) {
        const int c = warp + TS_WARPS * m;
        tw[m] = (c < k) ? tj * (Ct[c * KP + j] + sj * dot[m]) : 0.0;
      }
      __syncwarp();
#pragma unroll
      for (int m = 0; m < NC; m++) {
        const int c = warp + TS_WARPS * m;
        if (lane == 0 && c < k) Ct[c * KP + j] -= tw[m];
        const double f = tw[m] * sj;
#pragma unroll
        for (int i = 0; i < 4; i++) q[m][i] -= f * x[i];
      }
      __syncwarp();
    }
    /* the tile's rows were copied to shared memory before the barrier above, so Q can overwrite them in place */
#pragma unroll
    for (int m = 0; m < NC; m++) {
      const int c = warp + TS_WARPS * m;
      if (c < k) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int64_t r = tr + lane + 32 * i;
          if (r < r1) V[(int64_t)c * ld + r] = q[m][i];
        }
      }
    }
  }
}

static int ts_kp(int k) { return k <= 16 ? 16 : (k <= 32 ? 32 : 64); }

template <int KP>
static int ts_ctas_per_sm(void)
{
  static int cached = -1;
  if (cached < 0) {
    int a = 0, b = 0, c = 0;
    cudaFuncSetAttribute(k_tsqr_fwd<KP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ts_fwd_smem<KP>());
    cudaFuncSetAttribute(k_tsqr_fwd<KP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ts_fwd_smem<KP>());
    cudaFuncSetAttribute(k_tsqr_bwd<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ts_bwd_smem<KP>());
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_tsqr_fwd<KP, true>, TS_THREADS, ts_fwd_smem<KP>());
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_tsqr_fwd<KP, false>, TS_THREADS, ts_fwd_smem<KP>());
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, k_tsqr_bwd<KP>, TS_THREADS, ts_bwd_smem<KP>());
    cached = a < b ? a : b;
    if (c < cached) cached = c;
    if (cached < 1) cached = 1;
    if (cached > 8) cached = 8;
  }
  return cached;
}

/* the row blocking both kernels use: nblk CTAs of rows_per_blk rows (a multiple of 128), coef_elems doubles of reflector scalars */
extern "C" int b2k_tsqr_plan(b2k_ctx ctx, int64_t n, int k, int *nblk, int64_t *rows_per_blk, int64_t *coef_elems)
{
  ARGCHK(ctx && n >= 0 && k >= 1 && k <= B2K_TSQR_MAX_K, "TSQR takes 1..64 columns");
  CK(cudaSetDevice(ctx->device));
  const int kp = ts_kp(k);
  const int per_sm = kp == 16 ? ts_ctas_per_sm<16>() : (kp == 32 ? ts_ctas_per_sm<32>() : ts_ctas_per_sm<64>());
  const int64_t ntile = (n + TS_ROWS - 1) / TS_ROWS;
  int64_t g = (int64_t)ctx->sm_count * per_sm;
  if (g > ntile) g = ntile;
  if (g < 1) g = 1;
  const int64_t rpc = ((ntile + g - 1) / g) * TS_ROWS;
  g = rpc ? (n + rpc - 1) / rpc : 1;
  if (g < 1) g = 1;
  if (nblk) *nblk = (int)g;
  if (rows_per_blk) *rows_per_blk = rpc;
  if (coef_elems) *coef_elems = 2 * (ntile > 0 ? ntile : 1) * (int64_t)k;
  return B2K_OK;
}

extern "C" int b2k_tsqr_forward(b2k_ctx ctx, double *V, int64_t ld, int64_t n, int k, int store, double *Rblk, double *coef)
{
  int nblk = 0;
  int64_t rpc = 0;
  { const int rc = b2k_tsqr_plan(ctx, n, k, &nblk, &rpc, NULL); if (rc) return rc; }
  ARGCHK(Rblk && (coef || !store), "null output");
  const int kp = ts_kp(k);
  const int64_t ntile = (n + TS_ROWS - 1) / TS_ROWS;
  PROF_BEGIN(ctx, B2K_PROF_GEMM, 8.0 * (double)n * k * (store ? 2 : 1));
  PROF_FLOPS(ctx, 2.0 * (double)ntile * TS_ROWS * (double)k * k);
#define TS_FWD(KP)                                                                                                                  \
  do {                                                                                                                              \
    if (store) k_tsqr_fwd<KP, true><<<nblk, TS_THREADS, ts_fwd_smem<KP>(), ctx->stream>>>(V, ld, n, k, rpc, Rblk, coef);            \
    else k_tsqr_fwd<KP, false><<<nblk, TS_THREADS, ts_fwd_smem<KP>(), ctx->stream>>>(V, ld, n, k, rpc, Rblk, coef);                 \
  } while (0)
  if (kp == 16) TS_FWD(16); else if (kp == 32) TS_FWD(32); else TS_FWD(64);
#undef TS_FWD
  PROF_END(ctx);
  CKLAUNCH(ctx);
  return B2K_OK;
}

extern "C" int b2k_tsqr_backward(b2k_ctx ctx, double *V, int64_t ld, int64_t n, int k, const double *Wblk, const double *coef)
{
  int nblk = 0;
  int64_t rpc = 0;
  { const int rc = b2k_tsqr_plan(ctx, n, k, &nblk, &rpc, NULL); if (rc) return rc; }
  ARGCHK(Wblk && coef, "null input");
  if (n == 0) return B2K_OK;
  const int kp = ts_kp(k);
  const int64_t ntile = (n + TS_ROWS - 1) / TS_ROWS;
  PROF_BEGIN(ctx, B2K_PROF_GEMM, 16.0 * (double)n * k);
  PROF_FLOPS(ctx, 4.0 * (double)ntile * TS_ROWS * (double)k * k);
  if (kp == 16) k_tsqr_bwd<16><<<nblk, TS_THREADS, ts_bwd_smem<16>(), ctx->stream>>>(V, ld, n, k, rpc, Wblk, coef);
  else if (kp == 32) k_tsqr_bwd<32><<<nblk, TS_THREADS, ts_bwd_smem<32>(), ctx->stream>>>(V, ld, n, k, rpc, Wblk, coef);
  else k_tsqr_bwd<64><<<nblk, TS_THREADS, ts_bwd_smem<64>(), ctx->stream>>>(V, ld, n, k, rpc, Wblk, coef);
  PROF_END(ctx);
  CKLAUNCH(ctx);
  return B2K_OK;
}
