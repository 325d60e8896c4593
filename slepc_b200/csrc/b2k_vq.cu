/*
 * b2k_vq.cu — tall-skinny FP64 GEMM for the Krylov restart:  Out(n x nout) = beta*Out + alpha * In(n x kin) * Qe(kin x nout)
 * with kin, nout <= 64 (ncv ~ 64).  Replaces BVMultInPlace_BLAS_CUDA / BVMult_BLAS_CUDA (bvcuda.cu:65-112, 22-40:
 * cublasDgemm into an lda x (e-s) workspace + cudaMemcpy2D copy-back).  Out may alias columns of In (in place).
 *
 * Arithmetic intensity 2*kin*nout / (8*(kin+nout)) ~ 5.3 flop/B at 64 x 32: right at the FP64 ridge of B200
 * (37 TFLOP/s DFMA / 6.5 TB/s = 5.7), so the kernel has to keep BOTH the DFMA pipe and HBM busy:
 *   - a persistent CTA (2 per SM) stages a 128-row x kin tile of In in shared memory with 16-byte loads issued
 *     back to back (up to 16 in flight per thread), Q once per CTA;
 *   - 8 warps = 2 row halves x 4 column groups; a thread owns one row pair x C columns (C = nout/4): per k step
 *     ONE conflict-free LDS.128 of the tile (lanes along rows) + C/2 broadcast LDS.128 of Q feed 2*C DFMAs, so the
 *     shared-memory pipe (4 + C/2 wavefronts) stays at or below the DFMA pipe (C cycles per warp step);
 *   - the whole tile is in shared memory before the first store, so columns of Out that alias In are safe
 *     (rows are private to the CTA); the second resident CTA overlaps its loads/stores with this one's math.
 */
#include "b2k_internal.h"

#define VQ_ROWS    128
#define VQ_THREADS 256

__device__ __forceinline__ double2 vq_ld2(const double *p)
{
  double2 r;
  /* no .nc: Out may alias columns of In (BVMultInPlace) and the non-coherent path requires data that is read-only for
     the whole kernel */
  asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p) : "memory");
  return r;
}

template <int C>
__global__ void __launch_bounds__(VQ_THREADS, 2)
k_vq(double *Out, int64_t ldo, const double *In, int64_t ldi, int64_t n, int kin, int nout, const double *__restrict__ Q, int ldq,
     int qtrans, double alpha, double beta)
{
  extern __shared__ __align__(16) double vq_sm[];
  double *Vs = vq_sm;                              /* [kin][128]  */
  double *Qs = vq_sm + (size_t)kin * VQ_ROWS;      /* [kin][4*C]  */
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rh = warp >> 2, wc = warp & 3;
  const int QLD = 4 * C;
  for (int idx = tid; idx < kin * QLD; idx += VQ_THREADS) {
    const int i = idx / QLD, c = idx - i * QLD;
    double v = 0.0;
    if (c < nout) v = qtrans ? Q[(int64_t)c + (int64_t)i * ldq] : Q[(int64_t)i + (int64_t)c * ldq];
    Qs[idx] = v;
  }
  const int64_t ntiles = (n + VQ_ROWS - 1) / VQ_ROWS;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t row0 = t * VQ_ROWS;
    __syncthreads();                               /* previous tile fully consumed (and Qs visible) */
    /* ---- stage the tile: warp w takes columns w, w+8, …; a lane moves two row pairs per column ---- */
    {
      const int64_t ra = row0 + 2 * lane, rb = ra + 64;
      for (int cb = warp; cb < kin; cb += 64) {
        double2 va[8], vb[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int c = cb + 8 * u;
          va[u] = make_double2(0.0, 0.0);
          vb[u] = make_double2(0.0, 0.0);
          if (c < kin) {
            const double *p = In + (int64_t)c * ldi;
            if (ra + 1 < n) va[u] = vq_ld2(p + ra); else if (ra < n) va[u].x = p[ra];
            if (rb + 1 < n) vb[u] = vq_ld2(p + rb); else if (rb < n) vb[u].x = p[rb];
          }
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int c = cb + 8 * u;
          if (c < kin) {
            *reinterpret_cast<double2 *>(&Vs[(size_t)c * VQ_ROWS + 2 * lane]) = va[u];
            *reinterpret_cast<double2 *>(&Vs[(size_t)c * VQ_ROWS + 64 + 2 * lane]) = vb[u];
          }
        }
      }
    }
    __syncthreads();
    /* ---- math: one row pair x C columns per thread ---- */
    double acc0[C], acc1[C];
#pragma unroll
    for (int c = 0; c < C; c++) { acc0[c] = 0.0; acc1[c] = 0.0; }
    const double *ap = Vs + rh * 64 + 2 * lane;
    const double *bp = Qs + wc * C;
#pragma unroll 4
    for (int i = 0; i < kin; i++) {
      const double2 a = *reinterpret_cast<const double2 *>(ap + (size_t)i * VQ_ROWS);
#pragma unroll
      for (int c = 0; c < C; c += 2) {
        const double2 b = *reinterpret_cast<const double2 *>(bp + (size_t)i * QLD + c);
        acc0[c] = fma(a.x, b.x, acc0[c]);
        acc1[c] = fma(a.y, b.x, acc1[c]);
        acc0[c + 1] = fma(a.x, b.y, acc0[c + 1]);
        acc1[c + 1] = fma(a.y, b.y, acc1[c + 1]);
      }
    }
    /* ---- store (all reads of this tile came from shared memory, so aliasing In is safe) ---- */
    const int64_t r = row0 + rh * 64 + 2 * lane;
#pragma unroll
    for (int c = 0; c < C; c++) {
      const int col = wc * C + c;
      if (col < nout && r < n) {
        double *o = Out + (int64_t)col * ldo + r;
        if (r + 1 < n) {
          double2 v;
          if (beta == 0.0) { v.x = alpha * acc0[c]; v.y = alpha * acc1[c]; }
          else { const double2 old = *reinterpret_cast<double2 *>(o); v.x = fma(beta, old.x, alpha * acc0[c]); v.y = fma(beta, old.y, alpha * acc1[c]); }
          *reinterpret_cast<double2 *>(o) = v;
        } else {
          *o = (beta == 0.0) ? alpha * acc0[c] : fma(beta, *o, alpha * acc0[c]);
        }
      }
    }
  }
}

/* -1: shape not supported here (the generic k_gemm_ts handles it) */
int b2k_vq_launch(b2k_ctx ctx, double *Out, int64_t ldo, const double *In, int64_t ldi, int64_t n, int kin, int nout, const double *Q,
                  int ldq, int qtrans, double alpha, double beta)
{
  if (kin < 1 || kin > 64 || nout < 1 || nout > 64 || n < 1) return -1;
  if (!b2k_is_aligned16(Out) || !b2k_is_aligned16(In) || (ldo & 1) || (ldi & 1)) return -1;
  const int C = (nout <= 16) ? 4 : (nout <= 32 ? 8 : (nout <= 48 ? 12 : 16));
  const size_t shm = sizeof(double) * ((size_t)kin * VQ_ROWS + (size_t)kin * 4 * C);
  const int64_t ntiles = (n + VQ_ROWS - 1) / VQ_ROWS;
  int grid = ctx->sm_count * 2;
  if ((int64_t)grid > ntiles) grid = (int)ntiles;
  PROF_BEGIN(ctx, B2K_PROF_GEMM, 8.0 * (double)n * (kin + nout));
  PROF_FLOPS(ctx, 2.0 * (double)n * kin * nout);
#define VQ_LAUNCH(CC)                                                                                              \
  do {                                                                                                             \
    CK(cudaFuncSetAttribute(k_vq<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));                     \
    k_vq<CC><<<grid, VQ_THREADS, shm, ctx->stream>>>(Out, ldo, In, ldi, n, kin, nout, Q, ldq, qtrans, alpha, beta); \
  } while (0)
  if (C == 4) VQ_LAUNCH(4);
  else if (C == 8) VQ_LAUNCH(8);
  else if (C == 12) VQ_LAUNCH(12);
  else VQ_LAUNCH(16);
#undef VQ_LAUNCH
  PROF_END(ctx);
  CKLAUNCH(ctx);
  return B2K_OK;
}
