/*
 * b2k_bv.cu — hand-written sm_100a kernels for the BV (basis-vector) operations on SLEPc's
 * Krylov hot path.  They replace the cuBLAS wrappers of
 *   /root/reference/src/sys/classes/bv/impls/cuda/bvcuda.cu
 * (gemv 'C' :204-264, gemv 'N' :45-60, nrm2 :290-303, scal :269-285, gemm :22-40 / :65-112 /
 * :140-199, geam :117-135) and the coefficient micro-kernels (:345-548).
 *
 * Everything here is FP64 and HBM-bandwidth bound (AI <= 0.25 flop/B for the level-2 sweeps):
 * the design rules are coalesced 16-byte loads, enough independent loads in flight per SM to
 * cover HBM latency, grids sized in multiples of the SM count, and fixed-order two-stage
 * reductions (no atomics) so that results are bit-reproducible.
 */
#include "b2k_internal.h"

#define WARP 32

/* streaming 16-byte load: read-only path, do not allocate in L1 (V is touched once per sweep) */
__device__ __forceinline__ double2 ld_stream2(const double2 *p)
{
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ double ld_stream1(const double *p)
{
  double r;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

/* ------------------------------------------------------------------------------------------------
 * dotvec: part[blk][c0+c] = sum over this block's rows of V(r,c0+c)*w(r), one column tile of
 * <= CT columns per blockIdx.y; optional extra column k: w^T w (tile 0 only).
 * Each thread keeps CT accumulators and walks row PAIRS (16-byte loads) with a grid stride, so a
 * thread has up to CT independent 16 B loads in flight and V is read exactly once.
 * ---------------------------------------------------------------------------------------------- */
template <int CT, bool VEC2>
__global__ void __launch_bounds__(256, 2) k_dotvec(const double *__restrict__ V, int64_t ld, int64_t n, int k, int ctile,
                                                 const double *__restrict__ w, double *__restrict__ part, int pstride,
                                                 int with_ww)
{
  b2k_pdl_enter();
  const int tile = blockIdx.y;
  const int c0 = tile * ctile;
  const int nc = min(ctile, k - c0);
  double acc[CT];
#pragma unroll
  for (int c = 0; c < CT; c++) acc[c] = 0.0;
  double aw = 0.0;
  const bool do_ww = with_ww && tile == 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const double *Vc = V + (int64_t)c0 * ld;

  if (VEC2) {
    const int64_t npair = n >> 1;
    const double2 *w2 = reinterpret_cast<const double2 *>(w);
    for (int64_t p = t0; p < npair; p += stride) {
      const double2 wv = w2[p];
      double2 v[CT];
#pragma unroll
      for (int c = 0; c < CT; c++)
        if (c < nc) v[c] = ld_stream2(reinterpret_cast<const double2 *>(Vc + (int64_t)c * ld) + p);
#pragma unroll
      for (int c = 0; c < CT; c++)
        if (c < nc) { acc[c] = fma(v[c].x, wv.x, acc[c]); acc[c] = fma(v[c].y, wv.y, acc[c]); }
      if (do_ww) { aw = fma(wv.x, wv.x, aw); aw = fma(wv.y, wv.y, aw); }
    }
    if ((n & 1) && t0 == 0) {   /* odd tail row */
      const int64_t r = n - 1;
      const double wv = w[r];
#pragma unroll
      for (int c = 0; c < CT; c++)
        if (c < nc) acc[c] = fma(Vc[(int64_t)c * ld + r], wv, acc[c]);
      if (do_ww) aw = fma(wv, wv, aw);
    }
  } else {
    for (int64_t r = t0; r < n; r += stride) {
      const double wv = w[r];
#pragma unroll
      for (int c = 0; c < CT; c++)
        if (c < nc) acc[c] = fma(ld_stream1(Vc + (int64_t)c * ld + r), wv, acc[c]);
      if (do_ww) aw = fma(wv, wv, aw);
    }
  }

  /* block reduction: shuffle inside warps, then 8 warps through shared memory, fixed order */
  __shared__ double red[8][CT + 1];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < CT; c++) {
    double s = warp_sum(acc[c]);
    if (lane == 0) red[wid][c] = s;
  }
  {
    double s = warp_sum(aw);
    if (lane == 0) red[wid][CT] = s;
  }
  __syncthreads();
  if (threadIdx.x < nc) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 8; q++) s += red[q][threadIdx.x];
    part[(int64_t)blockIdx.x * pstride + c0 + threadIdx.x] = s;
  }
  if (do_ww && threadIdx.x == 32) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 8; q++) s += red[q][CT];
    part[(int64_t)blockIdx.x * pstride + k] = s;
  }
}

/* second stage: out[c] = sum_b part[b][c] in block order, one warp per column */
__global__ void __launch_bounds__(256) k_reduce_partials(const double *__restrict__ part, int nblk, int pstride, int ncols,
                                                         double *__restrict__ out)
{
  b2k_pdl_enter();
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= ncols) return;
  double s = 0.0;
  for (int b = lane; b < nblk; b += 32) s += part[(int64_t)b * pstride + c];
  s = warp_sum(s);
  if (lane == 0) out[c] = s;
}

/* second stage FUSED with the sum over the GPUs of the box (replaces k_reduce_partials + ncclAllReduce, i.e. the
 * MPIU_Allreduce of bvcuda.cu:228-248): the warp that has summed column c over this GPU's CTAs stores the value straight
 * into the mailbox of every rank over NVLink (peer-mapped HBM), the CTA publishes a sequence-numbered flag to every
 * rank, waits for the same flag from every rank in its OWN mailbox, and adds the `size` contributions in rank order —
 * all ranks get bit-identical sums.  One-shot, latency-bound (<= 1025 doubles): no ring, no NCCL launch.
 * Mailbox = [parity][source rank][column]; two parities suffice because a rank can only be one reduction ahead of the
 * slowest peer (it cannot finish reduction s+1 before every peer has started it, i.e. finished reading s). */
__global__ void __launch_bounds__(256) k_reduce_partials_xg(const double *part, int nblk, int pstride, int ncols,
                                                            double *out, const b2k_xg_s xg, unsigned long long seq)   /* out may be part (in-place all-reduce) */
{
  b2k_pdl_enter();
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  const size_t par = (size_t)(seq & 1ull);
  double s = 0.0;
  if (c < ncols) {
    for (int b = lane; b < nblk; b += 32) s += part[(int64_t)b * pstride + c];
    s = warp_sum(s);
    if (lane < xg.size) {                         /* lane p delivers to rank p (itself included) */
      volatile double *dst = xg.box[lane] + (par * B2K_XG_MAXR + xg.rank) * B2K_XG_MAXC + c;
      *dst = s;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < xg.size) {
    const int p = threadIdx.x;
    __threadfence_system();                       /* cumulative: orders the CTA's data stores before the flag */
    volatile unsigned long long *theirs =
        reinterpret_cast<unsigned long long *>(xg.box[p] + B2K_XG_DATA_ELEMS) + (par * B2K_XG_MAXR + xg.rank) * B2K_XG_MAXB + blockIdx.x;
    *theirs = seq;
    volatile unsigned long long *mine =
        reinterpret_cast<unsigned long long *>(xg.box[xg.rank] + B2K_XG_DATA_ELEMS) + (par * B2K_XG_MAXR + p) * B2K_XG_MAXB + blockIdx.x;
    const long long t0 = clock64();
    while (*mine < seq) {
      if (clock64() - t0 > xg.spin_limit) { *((volatile int *)xg.err) = 1; break; }   /* ~60 s: a peer is missing; fail, do not hang */
    }
    __threadfence_system();
  }
  __syncthreads();
  if (c < ncols && lane == 0) {
    const volatile double *src = xg.box[xg.rank] + par * B2K_XG_MAXR * B2K_XG_MAXC + c;
    double t = 0.0;
    for (int p = 0; p < xg.size; p++) t += src[(size_t)p * B2K_XG_MAXC];
    out[c] = t;
  }
}

/* ------------------------------------------------------------------------------------------------
 * multvec: y = beta*y + alpha * V(:,0:k) q, optionally emitting partial sums of ||y_new||^2.
 * One thread per row pair, coefficients in shared memory (broadcast reads), 8 independent 16 B
 * loads per batch.  No cross-thread reduction is needed for y itself.
 * ---------------------------------------------------------------------------------------------- */
template <bool VEC2, bool NRM>
__global__ void __launch_bounds__(256) k_multvec(const double *__restrict__ V, int64_t ld, int64_t n, int k, double alpha,
                                                  double beta, double *__restrict__ y, const double *__restrict__ q,
                                                  double *__restrict__ part, int pstride, int pcol, const b2k_gate_s gate)
{
  b2k_pdl_enter();
  if (b2k_gate_closed(gate)) return;
  extern __shared__ double qs[];
  for (int i = threadIdx.x; i < k; i += blockDim.x) qs[i] = q[i];
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double nrm = 0.0;
  if (VEC2) {
    const int64_t npair = n >> 1;
    double2 *y2 = reinterpret_cast<double2 *>(y);
    for (int64_t p = t0; p < npair; p += stride) {
      double ax = 0.0, ay = 0.0;
      int c = 0;
      for (; c + 8 <= k; c += 8) {
        double2 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = ld_stream2(reinterpret_cast<const double2 *>(V + (int64_t)(c + u) * ld) + p);
#pragma unroll
        for (int u = 0; u < 8; u++) { ax = fma(v[u].x, qs[c + u], ax); ay = fma(v[u].y, qs[c + u], ay); }
      }
      for (; c < k; c++) {
        const double2 v = ld_stream2(reinterpret_cast<const double2 *>(V + (int64_t)c * ld) + p);
        ax = fma(v.x, qs[c], ax); ay = fma(v.y, qs[c], ay);
      }
      double2 o;
      if (beta == 0.0) { o.x = alpha * ax; o.y = alpha * ay; }
      else { const double2 yo = y2[p]; o.x = fma(beta, yo.x, alpha * ax); o.y = fma(beta, yo.y, alpha * ay); }
      y2[p] = o;
      if (NRM) { nrm = fma(o.x, o.x, nrm); nrm = fma(o.y, o.y, nrm); }
    }
    if ((n & 1) && t0 == 0) {
      const int64_t r = n - 1;
      double a = 0.0;
      for (int c = 0; c < k; c++) a = fma(V[(int64_t)c * ld + r], qs[c], a);
      const double o = (beta == 0.0) ? alpha * a : fma(beta, y[r], alpha * a);
      y[r] = o;
      if (NRM) nrm = fma(o, o, nrm);
    }
  } else {
    for (int64_t r = t0; r < n; r += stride) {
      double a = 0.0;
      for (int c = 0; c < k; c++) a = fma(ld_stream1(V + (int64_t)c * ld + r), qs[c], a);
      const double o = (beta == 0.0) ? alpha * a : fma(beta, y[r], alpha * a);
      y[r] = o;
      if (NRM) nrm = fma(o, o, nrm);
    }
  }
  if (NRM) {
    __shared__ double red[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double s = warp_sum(nrm);
    if (lane == 0) red[wid] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
#pragma unroll
      for (int i = 0; i < 8; i++) t += red[i];
      part[(int64_t)blockIdx.x * pstride + pcol] = t;
    }
  }
}

/* ---- elementwise n x k block kernels (blockIdx.y = column) ------------------------------------ */
__global__ void __launch_bounds__(256) k_scale(double *__restrict__ X, int64_t ld, int64_t n, double alpha)
{
  b2k_pdl_enter();
  double *x = X + (int64_t)blockIdx.y * ld;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) x[r] *= alpha;
}
__global__ void __launch_bounds__(256) k_scale_rsqrt(double *__restrict__ x, int64_t n, const double *__restrict__ sumsq)
{
  const double s = sumsq[0];
  if (s == 0.0 || s == 1.0) return;
  const double a = 1.0 / sqrt(s);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) x[r] *= a;
}
/* the normalisation of a column whose DGKS refinement was decided on the device: sumsq = ||w||^2 after the second pass when the
   gate let it run, after the first otherwise.  Same IEEE operations as the host's 1/nrm, nrm = sqrt(sumsq). */
__global__ void __launch_bounds__(256) k_scale_rsqrt_gated(double *__restrict__ x, int64_t n, const b2k_gate_s gate, const double *__restrict__ nrm2_second)
{
  b2k_pdl_enter();
  const double s = b2k_gate_closed(gate) ? *reinterpret_cast<const volatile double *>(gate.nrm2) : nrm2_second[0];
  if (!(s > 0.0) || s == 1.0) return;
  const double a = 1.0 / sqrt(s);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) x[r] *= a;
}
__global__ void __launch_bounds__(256) k_copy(double *__restrict__ Y, int64_t ldy, const double *__restrict__ X, int64_t ldx,
                                               int64_t n)
{
  double *y = Y + (int64_t)blockIdx.y * ldy;
  const double *x = X + (int64_t)blockIdx.y * ldx;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) y[r] = x[r];
}
__global__ void __launch_bounds__(256) k_axpby(double *__restrict__ Y, int64_t ldy, const double *__restrict__ X, int64_t ldx,
                                                int64_t n, double alpha, double beta)
{
  double *y = Y + (int64_t)blockIdx.y * ldy;
  const double *x = X + (int64_t)blockIdx.y * ldx;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (beta == 0.0)
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) y[r] = alpha * x[r];
  else
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) y[r] = fma(alpha, x[r], beta * y[r]);
}
/* w = x .* y (Jacobi-preconditioned CG of the shift-and-invert ST: PETSc VecPointwiseMult) */
__global__ void __launch_bounds__(256) k_pointwise_mult(double *__restrict__ w, const double *__restrict__ x, const double *__restrict__ y, int64_t n)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) w[r] = x[r] * y[r];
}
__global__ void __launch_bounds__(256) k_fill(double *__restrict__ x, int64_t n, double v)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) x[r] = v;
}
/* splitmix64-finaliser hash → uniform [-1,1): identical to oracle/slepc_oracle.py:hash_uniform */
__host__ __device__ __forceinline__ double b2k_hash_uniform(uint64_t idx, uint64_t seed)
{
  uint64_t x = (idx + 1ull) * 0x9E3779B97F4A7C15ull + seed * 0xBF58476D1CE4E5B9ull;
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return 2.0 * ((double)(x >> 11) * (1.0 / 9007199254740992.0)) - 1.0;
}
__global__ void __launch_bounds__(256) k_set_random(double *__restrict__ x, int64_t n, int64_t row0, uint64_t seed)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride)
    x[r] = b2k_hash_uniform((uint64_t)(row0 + r), seed);
}
/* per-block partial of sum of squares of column blockIdx.y → part[blk][col] */
__global__ void __launch_bounds__(256) k_sumsq(const double *__restrict__ X, int64_t ld, int64_t n, double *__restrict__ part,
                                                int pstride)
{
  const double *x = X + (int64_t)blockIdx.y * ld;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double s = 0.0;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) { const double v = x[r]; s = fma(v, v, s); }
  __shared__ double red[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  s = warp_sum(s);
  if (lane == 0) red[wid] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) t += red[i];
    part[(int64_t)blockIdx.x * pstride + blockIdx.y] = t;
  }
}
__global__ void __launch_bounds__(256) k_colabssum(const double *__restrict__ X, int64_t ld, int64_t n, double *__restrict__ part,
                                                    int pstride)
{
  const double *x = X + (int64_t)blockIdx.y * ld;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double s = 0.0;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) s += fabs(x[r]);
  __shared__ double red[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  s = warp_sum(s);
  if (lane == 0) red[wid] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) t += red[i];
    part[(int64_t)blockIdx.x * pstride + blockIdx.y] = t;
  }
}
/* infinity norm of an n x k block = max over rows of sum_c |X(r,c)| (LAPACKlange 'I' of bvlapack.c:37-83): per-CTA maxima */
__global__ void __launch_bounds__(256) k_rowabssum_max(const double *__restrict__ X, int64_t ld, int64_t n, int k, double *__restrict__ part)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double m = 0.0;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
    double s = 0.0;
    for (int c = 0; c < k; c++) s += fabs(X[r + (int64_t)c * ld]);
    m = fmax(m, s);
  }
  __shared__ double red[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[wid] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) t = fmax(t, red[i]);
    part[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(256) k_max_partials(const double *__restrict__ part, int nblk, double *__restrict__ out)
{
  double m = 0.0;
  for (int b = threadIdx.x; b < nblk; b += 256) m = fmax(m, part[b]);
  __shared__ double red[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[wid] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) t = fmax(t, red[i]);
    out[0] = t;
  }
}
/* out[0] = sum_c in[c] (single thread; tiny) */
__global__ void k_sum_small(const double *__restrict__ in, int ncols, double *__restrict__ out)
{
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int c = 0; c < ncols; c++) s += in[c];
    out[0] = s;
  }
}

/* ------------------------------------------------------------------------------------------------
 * tall-skinny GEMM  Out(n x nout) = beta*Out + alpha * In(n x kin) * Qe(kin x nout)
 * with Qe(i,c) = Q[i + c*ldq] (qtrans=0) or Q[c + i*ldq] (qtrans=1).
 * One CTA owns RB rows: it stages the whole RB x kin input tile in shared memory FIRST and only
 * then writes its output columns, so Out may alias columns of In (BVMultInPlace, bvcuda.cu:65-112:
 * no lda x (e-s) workspace, no cudaMemcpy2D copy-back → half the reference's traffic).
 * Threads form a (RB/4) x 16 grid; each computes a 4 x 4 register tile of a 64-column output panel.
 * ---------------------------------------------------------------------------------------------- */
#define GEMM_KC 32
template <int RB>
__global__ void __launch_bounds__((RB / 4) * 16) k_gemm_ts(double *Out, int64_t ldo, const double *In, int64_t ldi, int64_t n,
                                                           int kin, int nout, const double *__restrict__ Q, int ldq,
                                                           int qtrans, double alpha, double beta)
{
  extern __shared__ double sm[];
  double *Vs = sm;                    /* [kin][RB]      */
  double *Qs = sm + (size_t)kin * RB; /* [GEMM_KC][64]  */
  const int tr = threadIdx.x;         /* 0..RB/4-1 : row group   */
  const int tc = threadIdx.y;         /* 0..15     : col group   */
  const int tid = tc * (RB / 4) + tr;
  const int nthr = (RB / 4) * 16;
  const int64_t row0 = (int64_t)blockIdx.x * RB;

  for (int idx = tid; idx < kin * RB; idx += nthr) {
    const int i = idx / RB, r = idx - i * RB;
    const int64_t gr = row0 + r;
    Vs[idx] = (gr < n) ? In[(int64_t)i * ldi + gr] : 0.0;
  }
  for (int p0 = 0; p0 < nout; p0 += 64) {
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
    for (int k0 = 0; k0 < kin; k0 += GEMM_KC) {
      __syncthreads();
      const int kc = min(GEMM_KC, kin - k0);
      for (int idx = tid; idx < GEMM_KC * 64; idx += nthr) {
        const int i = idx >> 6, c = idx & 63;
        double v = 0.0;
        if (i < kc && p0 + c < nout)
          v = qtrans ? Q[(int64_t)(p0 + c) + (int64_t)(k0 + i) * ldq] : Q[(int64_t)(k0 + i) + (int64_t)(p0 + c) * ldq];
        Qs[idx] = v;
      }
      __syncthreads();
#pragma unroll 4
      for (int i = 0; i < kc; i++) {
        const double2 a01 = *reinterpret_cast<const double2 *>(&Vs[(size_t)(k0 + i) * RB + 4 * tr]);
        const double2 a23 = *reinterpret_cast<const double2 *>(&Vs[(size_t)(k0 + i) * RB + 4 * tr + 2]);
        const double2 b01 = *reinterpret_cast<const double2 *>(&Qs[i * 64 + 4 * tc]);
        const double2 b23 = *reinterpret_cast<const double2 *>(&Qs[i * 64 + 4 * tc + 2]);
        const double a[4] = {a01.x, a01.y, a23.x, a23.y};
        const double b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
        for (int x = 0; x < 4; x++)
#pragma unroll
          for (int y = 0; y < 4; y++) acc[x][y] = fma(a[x], b[y], acc[x][y]);
      }
    }
#pragma unroll
    for (int y = 0; y < 4; y++) {
      const int c = p0 + 4 * tc + y;
      if (c < nout) {
#pragma unroll
        for (int x = 0; x < 4; x++) {
          const int64_t gr = row0 + 4 * tr + x;
          if (gr < n) {
            double *o = Out + (int64_t)c * ldo + gr;
            *o = (beta == 0.0) ? alpha * acc[x][y] : fma(beta, *o, alpha * acc[x][y]);
          }
        }
      }
    }
  }
}

/* =================================== host-side launchers ======================================== */
int b2k_launch_reduce_partials(b2k_ctx ctx, int nblk, int pstride, int ncols, double *out)
{
  if (ctx->xg && ctx->xg_on && ncols <= B2K_XG_MAXC && (ncols + 7) / 8 <= B2K_XG_MAXB)
    b2k_launch_pdl(k_reduce_partials_xg, dim3((ncols + 7) / 8), dim3(256), 0, ctx->stream, ctx->partials, nblk, pstride, ncols, out, *ctx->xg, ++ctx->xg_seq);
  else {
    ARGCHK(!(ctx->xg && ctx->xg_on), "reduction too wide for the peer-memory mailbox");
    b2k_launch_pdl(k_reduce_partials, dim3((ncols + 7) / 8), dim3(256), 0, ctx->stream, ctx->partials, nblk, pstride, ncols, out);
  }
  CKLAUNCH(ctx);
  return B2K_OK;
}

/* in-place sum over the ranks of n <= B2K_XG_MAXC doubles through the NVLink mailboxes (one launch; the reduction kernel with a
   single "CTA partial" = the buffer itself); returns -1 when the mailboxes are not open or n is too large → caller uses NCCL */
int b2k_xg_allreduce_inplace(b2k_ctx ctx, double *buf, int n)
{
  if (!ctx->xg || n < 1 || n > B2K_XG_MAXC || (n + 7) / 8 > B2K_XG_MAXB) return -1;
  b2k_launch_pdl(k_reduce_partials_xg, dim3((n + 7) / 8), dim3(256), 0, ctx->stream, buf, 1, n, n, buf, *ctx->xg, ++ctx->xg_seq);
  CKLAUNCH(ctx);
  return B2K_OK;
}

static inline int grid_rows(b2k_ctx ctx, int64_t work_items, int per_sm)
{
  int64_t need = (work_items + 255) / 256;
  int64_t cap = (int64_t)ctx->sm_count * per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

static int launch_dotvec(b2k_ctx ctx, const double *V, int64_t ld, int64_t n, int k, const double *w, double *out, int with_ww)
{
  ARGCHK(k >= 0 && k <= B2K_MAX_K, "k out of range");
  const int ncols = k + (with_ww ? 1 : 0);
  if (ncols == 0) return B2K_OK;
  if (n == 0) {                                   /* a rank without rows still takes part in a cross-GPU reduction */
    if (ctx->xg && ctx->xg_on) return b2k_launch_reduce_partials(ctx, 0, ncols, ncols, out);
    CK(cudaMemsetAsync(out, 0, sizeof(double) * ncols, ctx->stream));
    return B2K_OK;
  }
  const int CT = 16;
  int ntiles = (k + CT - 1) / CT;
  if (ntiles < 1) ntiles = 1;
  const int ctile = (k + ntiles - 1) / ntiles > 0 ? (k + ntiles - 1) / ntiles : 1;
  const bool vec2 = b2k_is_aligned16(V) && b2k_is_aligned16(w) && (ld % 2 == 0);
  /* exactly one wave: 2 resident CTAs per SM (launch bounds), shared among the column tiles */
  int64_t items = vec2 ? (n >> 1) : n;
  int gx = (ctx->sm_count * 2) / ntiles;
  if (gx < 1) gx = 1;
  int64_t need = (items + 255) / 256;
  if (need < 1) need = 1;
  if (gx > need) gx = (int)need;
  if (gx > B2K_MAX_PART_BLOCKS) gx = B2K_MAX_PART_BLOCKS;
  const int pstride = ncols;
  dim3 grid(gx, ntiles);
  PROF_BEGIN(ctx, B2K_PROF_DOTVEC, 8.0 * (double)n * (k + 1));
  if (vec2) b2k_launch_pdl(k_dotvec<16, true>, grid, dim3(256), 0, ctx->stream, V, ld, n, k, ctile, w, ctx->partials, pstride, with_ww);
  else      b2k_launch_pdl(k_dotvec<16, false>, grid, dim3(256), 0, ctx->stream, V, ld, n, k, ctile, w, ctx->partials, pstride, with_ww);
  PROF_END(ctx);
  CKLAUNCH(ctx);
  return b2k_launch_reduce_partials(ctx, gx, pstride, ncols, out);
}

extern "C" int b2k_dotvec(b2k_ctx ctx, const double *V, int64_t ld, int64_t n, int k, const double *y, double *q)
{
  return launch_dotvec(ctx, V, ld, n, k, y, q, 0);
}
extern "C" int b2k_gs_dot(b2k_ctx ctx, const double *V, int64_t ld, int64_t n, int k, const double *w, double *c)
{
  return launch_dotvec(ctx, V, ld, n, k, w, c, 1);
}

static int launch_multvec(b2k_ctx ctx, const double *V, int64_t ld, int64_t n, int k, double alpha, double beta, double *y,
                          const double *q, double *nrm_out, const b2k_gate_s *gatep = nullptr)
{
  const b2k_gate_s gate = gatep ? *gatep : b2k_gate_s{nullptr, nullptr, 0.0};
  ARGCHK(k >= 0 && k <= B2K_MAX_K, "k out of range");
  if (n == 0) {
    if (nrm_out && ctx->xg && ctx->xg_on) return b2k_launch_reduce_partials(ctx, 0, 1, 1, nrm_out);
    if (nrm_out) CK(cudaMemsetAsync(nrm_out, 0, sizeof(double), ctx->stream));
    return B2K_OK;
  }
  const bool vec2 = b2k_is_aligned16(V) && b2k_is_aligned16(y) && (ld % 2 == 0);
  const int64_t items = vec2 ? (n >> 1) : n;
  int gx = grid_rows(ctx, items > 0 ? items : 1, 6);
  if (gx > B2K_MAX_PART_BLOCKS) gx = B2K_MAX_PART_BLOCKS;
  const size_t shm = sizeof(double) * (size_t)(k > 0 ? k : 1);
  PROF_BEGIN(ctx, B2K_PROF_MULTVEC, 8.0 * (double)n * (k + (beta == 0.0 ? 1 : 2)));
  if (nrm_out) {
    if (vec2) b2k_launch_pdl(k_multvec<true, true>, dim3(gx), dim3(256), shm, ctx->stream, V, ld, n, k, alpha, beta, y, q, ctx->partials, 1, 0, gate);
    else      b2k_launch_pdl(k_multvec<false, true>, dim3(gx), dim3(256), shm, ctx->stream, V, ld, n, k, alpha, beta, y, q, ctx->partials, 1, 0, gate);
    PROF_END(ctx);
    CKLAUNCH(ctx);
    { const int rc_ = b2k_launch_reduce_partials(ctx, gx, 1, 1, nrm_out); if (rc_) return rc_; }
  } else {
    if (vec2) b2k_launch_pdl(k_multvec<true, false>, dim3(gx), dim3(256), shm, ctx->stream, V, ld, n, k, alpha, beta, y, q, (double *)nullptr, 0, 0, gate);
    else      b2k_launch_pdl(k_multvec<false, false>, dim3(gx), dim3(256), shm, ctx->stream, V, ld, n, k, alpha, beta, y, q, (double *)nullptr, 0, 0, gate);
    PROF_END(ctx);
    CKLAUNCH(ctx);
  }
  return B2K_OK;
}

int b2k_gs_rt_launch(b2k_ctx ctx, const double *V, int64_t ld, int64_t n, int k, double alpha, double beta, double *w,
                     const double *q, int dot, double *out, const b2k_gate_s *gate);   /* b2k_gs_rt.cu (register tile) */
int b2k_gs_tma_launch(b2k_ctx ctx, const double *V, int64_t ld, int64_t n, int k, double alpha, double beta, double *w,
                      const double *q, int dot, double *out, const b2k_gate_s *gate);  /* b2k_gs_tma.cu (2-D tensor-map pipeline) */

/* implementation of the update sweeps; env B2K_GS_FUSED: 0 generic kernels (two sweeps for update+dot), 1 register-tile single
   sweep, 3 (default) 2-D tensor-map (TMA) pipelined single sweep with the register tile below 4096 rows or k <= 4 */
static int g_fused_enabled = -1;
static int b2k_gs_fused_enabled(void)
{
  if (g_fused_enabled < 0) {
    const char *e = getenv("B2K_GS_FUSED");
    g_fused_enabled = (e && (e[0] == '0' || e[0] == '1' || e[0] == '3')) ? e[0] - '0' : 3;
  }
  return g_fused_enabled;
}
extern "C" int b2k_gs_set_fused(int mode) { g_fused_enabled = (mode == 0 || mode == 1 || mode == 3) ? mode : 3; return B2K_OK; }

/* single-sweep update kernels in order of preference for the selected mode; -1 = shape not supported */
static int gs_single_sweep(b2k_ctx ctx, int mode, const double *V, int64_t ld, int64_t n, int k, double alpha, double beta, double *w,
                           const double *q, int dot, double *out, const b2k_gate_s *gate = nullptr)
{
  if (mode == 3) {
    int rc = b2k_gs_tma_launch(ctx, V, ld, n, k, alpha, beta, w, q, dot, out, gate);
    if (rc != -1) return rc;
  }
  return b2k_gs_rt_launch(ctx, V, ld, n, k, alpha, beta, w, q, dot, out, gate);
}

extern "C" int b2k_multvec(b2k_ctx ctx, const double *V, int64_t ld, int64_t n, int k, double alpha, double beta, double *y,
                           const double *q)
{
  if (b2k_gs_fused_enabled() && k > 0 && n > 0) {
    int rc = gs_single_sweep(ctx, b2k_gs_fused_enabled(), V, ld, n, k, alpha, beta, y, q, 0, nullptr);
    if (rc != -1) return rc;
  }
  return launch_multvec(ctx, V, ld, n, k, alpha, beta, y, q, nullptr);
}

extern "C" int b2k_gs_update_dot(b2k_ctx ctx, const double *V, int64_t ld, int64_t n, int k, double *w, const double *cin,
                                 double *cout)
{
  const int mode = b2k_gs_fused_enabled();
  if (mode && k > 0 && n > 0) {
    int rc = gs_single_sweep(ctx, mode, V, ld, n, k, -1.0, 1.0, w, cin, 1, cout);
    if (rc != -1) return rc;       /* -1: shape not supported by the single-sweep kernels → two-sweep path */
  }
  /* two-sweep path: update sweep, then dot sweep (V read twice) */
  int rc = launch_multvec(ctx, V, ld, n, k, -1.0, 1.0, w, cin, nullptr);
  if (rc) return rc;
  return launch_dotvec(ctx, V, ld, n, k, w, cout, 1);
}

/* w -= V cin ; out[0] = ||w_new||^2 : the update sweep with the explicit norm folded in (one read of V) */
extern "C" int b2k_gs_update_norm(b2k_ctx ctx, const double *V, int64_t ld, int64_t n, int k, double *w, const double *cin,
                                  double *nrm2_out)
{
  if (b2k_gs_fused_enabled() && k > 0 && n > 0) {
    int rc = gs_single_sweep(ctx, b2k_gs_fused_enabled(), V, ld, n, k, -1.0, 1.0, w, cin, 0, nrm2_out);
    if (rc != -1) return rc;
  }
  return launch_multvec(ctx, V, ld, n, k, -1.0, 1.0, w, cin, nrm2_out);
}

/* the same sweep launched SPECULATIVELY behind the sweep that produced *onrm2 and *nrm2 (device scalars): it runs only if the
   DGKS criterion of bvorthog.c:180 asks for a refinement, nrm != 0 && nrm < eta*onrm with onrm = sqrt(max(*onrm2,0)),
   nrm = sqrt(max(*nrm2,0)); otherwise w and (apart from the reduction of stale partials) nrm2_out are left alone */
extern "C" int b2k_gs_update_norm_gated(b2k_ctx ctx, const double *V, int64_t ld, int64_t n, int k, double *w, const double *cin,
                                        double *nrm2_out, const double *onrm2, const double *nrm2, double eta)
{
  ARGCHK(onrm2 && nrm2 && nrm2_out, "null gate / output pointer");
  const b2k_gate_s gate = {onrm2, nrm2, eta};
  if (b2k_gs_fused_enabled() && k > 0 && n > 0) {
    int rc = gs_single_sweep(ctx, b2k_gs_fused_enabled(), V, ld, n, k, -1.0, 1.0, w, cin, 0, nrm2_out, &gate);
    if (rc != -1) return rc;
  }
  return launch_multvec(ctx, V, ld, n, k, -1.0, 1.0, w, cin, nrm2_out, &gate);
}

extern "C" int b2k_sumsq(b2k_ctx ctx, const double *X, int64_t ld, int64_t n, int k, double *out)
{
  ARGCHK(k >= 1 && k <= B2K_MAX_K, "k out of range");
  if (n == 0) {
    /* a rank without rows issues the SAME cross-GPU reduction as the ranks with rows (k columns, then their sum) */
    if (ctx->xg && ctx->xg_on) {
      const int rc_ = b2k_launch_reduce_partials(ctx, 0, k, k, ctx->dscratch);
      if (rc_) return rc_;
      k_sum_small<<<1, 32, 0, ctx->stream>>>(ctx->dscratch, k, out);
      CKLAUNCH(ctx);
      return B2K_OK;
    }
    CK(cudaMemsetAsync(out, 0, sizeof(double), ctx->stream));
    return B2K_OK;
  }
  int gx = grid_rows(ctx, n, 4);
  if ((int64_t)gx * k > (int64_t)ctx->sm_count * 8) gx = (ctx->sm_count * 8 + k - 1) / k;
  if (gx < 1) gx = 1;
  if (gx > B2K_MAX_PART_BLOCKS) gx = B2K_MAX_PART_BLOCKS;
  dim3 grid(gx, k);
  k_sumsq<<<grid, 256, 0, ctx->stream>>>(X, ld, n, ctx->partials, k);
  CKLAUNCH(ctx);
  double *tmp = ctx->dscratch;           /* per-column sums */
  { const int rc_ = b2k_launch_reduce_partials(ctx, gx, k, k, tmp); if (rc_) return rc_; }
  k_sum_small<<<1, 32, 0, ctx->stream>>>(tmp, k, out);
  CKLAUNCH(ctx);
  return B2K_OK;
}

extern "C" int b2k_colabssum(b2k_ctx ctx, const double *X, int64_t ld, int64_t n, int k, double *out_k)
{
  ARGCHK(k >= 1 && k <= B2K_MAX_K, "k out of range");
  if (n == 0) {
    if (ctx->xg && ctx->xg_on) return b2k_launch_reduce_partials(ctx, 0, k, k, out_k);
    CK(cudaMemsetAsync(out_k, 0, sizeof(double) * k, ctx->stream));
    return B2K_OK;
  }
  int gx = grid_rows(ctx, n, 4);
  if ((int64_t)gx * k > (int64_t)ctx->sm_count * 8) gx = (ctx->sm_count * 8 + k - 1) / k;
  if (gx < 1) gx = 1;
  dim3 grid(gx, k);
  k_colabssum<<<grid, 256, 0, ctx->stream>>>(X, ld, n, ctx->partials, k);
  CKLAUNCH(ctx);
  { const int rc_ = b2k_launch_reduce_partials(ctx, gx, k, k, out_k); if (rc_) return rc_; }
  return B2K_OK;
}

/* out[0] = max_r sum_c |X(r,c)| over this GPU's rows (the caller combines ranks with a MAX all-reduce) */
extern "C" int b2k_rowabssum_max(b2k_ctx ctx, const double *X, int64_t ld, int64_t n, int k, double *out)
{
  ARGCHK(k >= 1 && k <= B2K_MAX_K, "k out of range");
  if (n == 0) { CK(cudaMemsetAsync(out, 0, sizeof(double), ctx->stream)); return B2K_OK; }
  int gx = grid_rows(ctx, n, 4);
  if (gx > B2K_MAX_PART_BLOCKS) gx = B2K_MAX_PART_BLOCKS;
  k_rowabssum_max<<<gx, 256, 0, ctx->stream>>>(X, ld, n, k, ctx->partials);
  CKLAUNCH(ctx);
  k_max_partials<<<1, 256, 0, ctx->stream>>>(ctx->partials, gx, out);
  CKLAUNCH(ctx);
  return B2K_OK;
}

static inline dim3 grid2d(b2k_ctx ctx, int64_t n, int k)
{
  int64_t need = (n + 255) / 256;
  int64_t cap = ((int64_t)ctx->sm_count * 8 + k - 1) / k;
  if (cap < 1) cap = 1;
  if (need < 1) need = 1;
  return dim3((unsigned)(need < cap ? need : cap), (unsigned)k);
}

extern "C" int b2k_scale(b2k_ctx ctx, double *X, int64_t ld, int64_t n, int k, double alpha)
{
  if (n == 0 || k == 0) return B2K_OK;
  if (alpha == 0.0) {   /* bvcuda.cu:277: memset */
    CK(cudaMemset2DAsync(X, ld * sizeof(double), 0, n * sizeof(double), k, ctx->stream));
    return B2K_OK;
  }
  PROF_BEGIN(ctx, B2K_PROF_ELEMWISE, 16.0 * (double)n * k);
  b2k_launch_pdl(k_scale, grid2d(ctx, n, k), dim3(256), 0, ctx->stream, X, ld, n, alpha);
  PROF_END(ctx);
  CKLAUNCH(ctx);
  return B2K_OK;
}
extern "C" int b2k_scale_rsqrt(b2k_ctx ctx, double *x, int64_t n, const double *sumsq)
{
  if (n == 0) return B2K_OK;
  PROF_BEGIN(ctx, B2K_PROF_ELEMWISE, 16.0 * (double)n);
  k_scale_rsqrt<<<grid2d(ctx, n, 1), 256, 0, ctx->stream>>>(x, n, sumsq);
  PROF_END(ctx);
  CKLAUNCH(ctx);
  return B2K_OK;
}
extern "C" int b2k_scale_rsqrt_gated(b2k_ctx ctx, double *x, int64_t n, const double *onrm2, const double *nrm2_first, const double *nrm2_second,
                                     double eta)
{
  ARGCHK(onrm2 && nrm2_first && nrm2_second, "null argument");
  if (n == 0) return B2K_OK;
  b2k_gate_s g;
  g.onrm2 = onrm2; g.nrm2 = nrm2_first; g.eta = eta;
  PROF_BEGIN(ctx, B2K_PROF_ELEMWISE, 16.0 * (double)n);
  b2k_launch_pdl(k_scale_rsqrt_gated, grid2d(ctx, n, 1), dim3(256), 0, ctx->stream, x, n, g, nrm2_second);
  PROF_END(ctx);
  CKLAUNCH(ctx);
  return B2K_OK;
}
extern "C" int b2k_copy(b2k_ctx ctx, double *Y, int64_t ldy, const double *X, int64_t ldx, int64_t n, int k)
{
  if (n == 0 || k == 0) return B2K_OK;
  PROF_BEGIN(ctx, B2K_PROF_ELEMWISE, 16.0 * (double)n * k);
  k_copy<<<grid2d(ctx, n, k), 256, 0, ctx->stream>>>(Y, ldy, X, ldx, n);
  PROF_END(ctx);
  CKLAUNCH(ctx);
  return B2K_OK;
}
extern "C" int b2k_axpby(b2k_ctx ctx, double *Y, int64_t ldy, const double *X, int64_t ldx, int64_t n, int k, double alpha,
                         double beta)
{
  if (n == 0 || k == 0) return B2K_OK;
  PROF_BEGIN(ctx, B2K_PROF_ELEMWISE, (beta == 0.0 ? 16.0 : 24.0) * (double)n * k);
  k_axpby<<<grid2d(ctx, n, k), 256, 0, ctx->stream>>>(Y, ldy, X, ldx, n, alpha, beta);
  PROF_END(ctx);
  CKLAUNCH(ctx);
  return B2K_OK;
}
extern "C" int b2k_pointwise_mult(b2k_ctx ctx, double *w, const double *x, const double *y, int64_t n)
{
  if (n == 0) return B2K_OK;
  PROF_BEGIN(ctx, B2K_PROF_ELEMWISE, 24.0 * (double)n);
  k_pointwise_mult<<<grid2d(ctx, n, 1), 256, 0, ctx->stream>>>(w, x, y, n);
  PROF_END(ctx);
  CKLAUNCH(ctx);
  return B2K_OK;
}
extern "C" int b2k_fill(b2k_ctx ctx, double *x, int64_t n, double v)
{
  if (n == 0) return B2K_OK;
  PROF_BEGIN(ctx, B2K_PROF_ELEMWISE, 8.0 * (double)n);
  k_fill<<<grid2d(ctx, n, 1), 256, 0, ctx->stream>>>(x, n, v);
  PROF_END(ctx);
  CKLAUNCH(ctx);
  return B2K_OK;
}
extern "C" int b2k_set_random(b2k_ctx ctx, double *x, int64_t n, int64_t row0, uint64_t seed)
{
  if (n == 0) return B2K_OK;
  k_set_random<<<grid2d(ctx, n, 1), 256, 0, ctx->stream>>>(x, n, row0, seed);
  CKLAUNCH(ctx);
  return B2K_OK;
}

int b2k_vq_launch(b2k_ctx ctx, double *Out, int64_t ldo, const double *In, int64_t ldi, int64_t n, int kin, int nout, const double *Q,
                  int ldq, int qtrans, double alpha, double beta);   /* b2k_vq.cu */
int b2k_vq_tma_launch(b2k_ctx ctx, double *Out, int64_t ldo, const double *In, int64_t ldi, int64_t n, int kin, int nout, const double *Q,
                      int ldq, int qtrans, double alpha, double beta);   /* b2k_vq_tma.cu (TMA ring + FP64 tensor cores) */

static int launch_gemm_ts(b2k_ctx ctx, double *Out, int64_t ldo, const double *In, int64_t ldi, int64_t n, int kin, int nout,
                          const double *Q, int ldq, int qtrans, double alpha, double beta)
{
  if (n == 0 || nout == 0) return B2K_OK;
  ARGCHK(kin >= 0 && kin <= 2048, "kin out of range");
  if (kin == 0) return b2k_scale(ctx, Out, ldo, n, nout, beta);
  {
    int rc = b2k_vq_tma_launch(ctx, Out, ldo, In, ldi, n, kin, nout, Q, ldq, qtrans, alpha, beta);   /* kin, nout <= 64, large n */
    if (rc != -1) return rc;
    rc = b2k_vq_launch(ctx, Out, ldo, In, ldi, n, kin, nout, Q, ldq, qtrans, alpha, beta);           /* kin, nout <= 64: the restart shape */
    if (rc != -1) return rc;
  }
  const size_t qbytes = sizeof(double) * GEMM_KC * 64;
  const size_t lim = 200 * 1024;
#define LAUNCH_GEMM(RB)                                                                                         \
  do {                                                                                                          \
    const size_t shm = sizeof(double) * (size_t)kin * RB + qbytes;                                              \
    CK(cudaFuncSetAttribute(k_gemm_ts<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));             \
    dim3 blk(RB / 4, 16);                                                                                       \
    const int64_t gx = (n + RB - 1) / RB;                                                                       \
    PROF_BEGIN(ctx, B2K_PROF_GEMM, 8.0 * (double)n * (kin + nout));                                             \
    PROF_FLOPS(ctx, 2.0 * (double)n * kin * nout);                                                              \
    k_gemm_ts<RB><<<(unsigned)gx, blk, shm, ctx->stream>>>(Out, ldo, In, ldi, n, kin, nout, Q, ldq, qtrans, alpha, beta); \
    PROF_END(ctx);                                                                                              \
    CKLAUNCH(ctx);                                                                                              \
  } while (0)
  if (sizeof(double) * (size_t)kin * 64 + qbytes <= 96 * 1024) LAUNCH_GEMM(64);
  else if (sizeof(double) * (size_t)kin * 32 + qbytes <= lim) LAUNCH_GEMM(32);
  else if (sizeof(double) * (size_t)kin * 8 + qbytes <= lim) LAUNCH_GEMM(8);
  else { b2k_set_error("b2k gemm: kin=%d too large for the shared-memory row tile", kin); return B2K_ERR_ARG; }
#undef LAUNCH_GEMM
  return B2K_OK;
}

extern "C" int b2k_mult(b2k_ctx ctx, double *Y, int64_t ldy, const double *X, int64_t ldx, int64_t n, int ky, int kx, double alpha,
                        double beta, const double *Q, int ldq)
{
  return launch_gemm_ts(ctx, Y, ldy, X, ldx, n, kx, ky, Q, ldq, 0, alpha, beta);
}

extern "C" int b2k_mult_inplace(b2k_ctx ctx, double *V, int64_t ld, int64_t n, int k, int s, int e, const double *Q, int ldq,
                                int trans)
{
  ARGCHK(s >= 0 && e >= s, "bad column range");
  if (e == s) return B2K_OK;
  const double *Qb = trans ? Q + s : Q + (int64_t)s * ldq;
  return launch_gemm_ts(ctx, V + (int64_t)s * ld, ld, V, ld, n, k, e - s, Qb, ldq, trans, 1.0, 0.0);
}

int b2k_gram_tma_launch(b2k_ctx ctx, const double *Y, int64_t ldy, const double *X, int64_t ldx, int64_t n, int ky, int kx, double *M,
                        int ldm);   /* b2k_gram_tma.cu (one sweep, FP64 tensor cores) */

extern "C" int b2k_dot(b2k_ctx ctx, const double *Y, int64_t ldy, const double *X, int64_t ldx, int64_t n, int ky, int kx, double *M,
                       int ldm)
{
  /* one sweep of Y and X per 64 x 64 block of M (k_gram_tma); blocks that do not start at M(0,0) with ldm == ky go through
     the context's scratch and a strided copy */
  if (ky > 0 && kx > 0 && n > 0) {
    bool ok = true;
    for (int y0 = 0; ok && y0 < ky; y0 += 64) {
      for (int x0 = 0; ok && x0 < kx; x0 += 64) {
        const int by = ky - y0 < 64 ? ky - y0 : 64, bx = kx - x0 < 64 ? kx - x0 : 64;
        const bool direct = (ky <= 64 && kx <= 64 && ldm == ky);
        double *dst = direct ? M : ctx->dscratch;
        const int rc = b2k_gram_tma_launch(ctx, Y + (int64_t)y0 * ldy, ldy, X + (int64_t)x0 * ldx, ldx, n, by, bx, dst, by);
        if (rc == -1) { ok = false; break; }
        if (rc) return rc;
        if (!direct)
          CK(cudaMemcpy2DAsync(M + (int64_t)x0 * ldm + y0, sizeof(double) * (size_t)ldm, ctx->dscratch, sizeof(double) * (size_t)by,
                               sizeof(double) * (size_t)by, (size_t)bx, cudaMemcpyDeviceToDevice, ctx->stream));
      }
    }
    if (ok) return B2K_OK;
  }
  /* shapes the tensor-map kernel does not take (short or unaligned blocks): M(:,j) = Y^T X(:,j), kx fused-reduction sweeps */
  for (int j = 0; j < kx; j++) {
    int rc = launch_dotvec(ctx, Y, ldy, n, ky, X + (int64_t)j * ldx, M + (int64_t)j * ldm, 0);
    if (rc) return rc;
  }
  return B2K_OK;
}
