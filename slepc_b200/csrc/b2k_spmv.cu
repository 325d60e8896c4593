/*
 * b2k_spmv.cu — FP64 CSR sparse matrix-vector product for sm_100a and device-side stencil
 * generators.  Replaces the PETSc MatMult (AIJCUSPARSE → cusparseSpMV) that SLEPc reaches through
 *   BVMatMultColumn  /root/reference/src/sys/classes/bv/interface/bvops.c:862-885
 *   MatMult_STOperator / STApply_Generic  src/sys/classes/st/interface/stsolve.c:244-258, 16-25.
 *
 * Algorithm ("CSR-stream"): the rows are cut on the host into blocks whose non-zeros fit a
 * shared-memory buffer.  A CTA streams its block's (col,val) pairs with fully coalesced loads,
 * multiplies by the gathered x entries, parks the products in shared memory, and then each thread
 * sums the products of its own row(s).  Matrix bytes are read exactly once and contiguously; the
 * x gather hits L1/L2 (stencil locality).  HBM-bound: 12 B per non-zero + 4 B rowptr + 8 B y + 8 B x
 * per row.
 */
#include <stdlib.h>
#include <algorithm>
#include <string.h>
#include "b2k_internal.h"

#define SPMV_THREADS 256
#define SPMV_CAP     2560      /* products per CTA held in shared memory (20 KB)            */
#define SPMV_MAXROWS 512       /* rows per CTA                                               */

struct b2k_csr_s {
  int64_t nrows, ncols_local, nghost, nnz;
  int    *rowptr, *colidx;
  double *val;
  int    *blkrow;     /* [nblk+1] first row of each row block */
  int     nblk;
  /* SELL-32 copy (sliced ELLPACK, slice height 32 = one warp, no row sorting): built when padding is small */
  int64_t  nslices, sell_elems;
  int64_t *sl_off;    /* [nslices+1 (+3 padding)] element offset of each slice; width = (off[s+1]-off[s])/32 */
  int     *sl_col;
  double  *sl_val;
  /* chunks of whole slices (even first slice, <= SP_CAP entries, <= SP_MAXS slices) for the bulk-copy pipeline kernel */
  int     *sp_chunk;  /* [nchunks+1] first slice of each chunk */
  int      nchunks;
  int      sp_cap;    /* entries of the largest chunk, rounded up to 32 */
  b2k_ctx  ctx;       /* owner (b2k_csr_arrays rebuilds the CSR copy on this context's stream)                     */
  int      csr_dropped;   /* colidx/val were freed once the SELL copy existed: the matrix is stored ONCE            */
  int      last_kernel;   /* B2K_SPMV_KERNEL_* of the most recent product                                           */
};

/* ------------------------------------------------------------------------------------------------
 * SELL-32 SpMV: one warp per slice of 32 consecutive rows, lane = row.  Entry w of the slice's rows is
 * stored at off + 32*w + lane, so every (col,val) load of a warp is one fully coalesced 128 B / 256 B
 * request, there is no row pointer, no shared memory and no barrier, and a thread has `width`
 * independent matrix loads + gathers in flight.  Padding entries carry val = 0 and the row's own first
 * column.  Same algorithmic bytes as CSR minus the row pointers.
 * ---------------------------------------------------------------------------------------------- */
__device__ __forceinline__ int ld_stream_i32(const int *p)
{
  int r;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ double ld_stream_f64(const double *p)
{
  double r;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
}

#define SELL_CHUNK 8
/* persistent warps: each warp walks the slices with a grid stride, TWO slices per iteration so that a lane keeps up to
   16 matrix loads (192 B) in flight before it touches x */
__global__ void __launch_bounds__(256) k_spmv_sell(const int64_t *__restrict__ sl_off, const int *__restrict__ col,
                                                    const double *__restrict__ val, const double *__restrict__ x,
                                                    const double *__restrict__ xg, int ncl, double *__restrict__ y, int64_t nrows,
                                                    int64_t nslices, double sigma)
{
  b2k_pdl_enter();
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t s0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; s0 < nslices; s0 += 2 * nwarps) {
    const int64_t s1 = s0 + nwarps;
    const bool two = s1 < nslices;
    const int64_t off0 = sl_off[s0], end0 = sl_off[s0 + 1];
    const int64_t off1 = two ? sl_off[s1] : 0, end1 = two ? sl_off[s1 + 1] : 0;
    const int width0 = (int)((end0 - off0) >> 5), width1 = (int)((end1 - off1) >> 5);
    const int *cp0 = col + off0 + lane, *cp1 = col + off1 + lane;
    const double *vp0 = val + off0 + lane, *vp1 = val + off1 + lane;
    double acc0 = 0.0, acc1 = 0.0;
    const int wmax = max(width0, width1);
    for (int w = 0; w < wmax; w += SELL_CHUNK) {
      /* issue both slices' matrix loads back to back, then the gathers */
      int c0[SELL_CHUNK], c1[SELL_CHUNK];
      double v0[SELL_CHUNK], v1[SELL_CHUNK];
#pragma unroll
      for (int u = 0; u < SELL_CHUNK; u++) {
        const bool on0 = w + u < width0, on1 = w + u < width1;
        c0[u] = on0 ? ld_stream_i32(cp0 + 32 * (w + u)) : 0;
        v0[u] = on0 ? ld_stream_f64(vp0 + 32 * (w + u)) : 0.0;
        c1[u] = on1 ? ld_stream_i32(cp1 + 32 * (w + u)) : 0;
        v1[u] = on1 ? ld_stream_f64(vp1 + 32 * (w + u)) : 0.0;
      }
      /* branch-free gathers (entries past the slice width read a valid dummy address with v = 0) so that all of them are
         in flight together */
      double x0[SELL_CHUNK], x1[SELL_CHUNK];
#pragma unroll
      for (int u = 0; u < SELL_CHUNK; u++) {
        const double *p0 = (c0[u] < ncl) ? x + c0[u] : xg + (c0[u] - ncl);
        const double *p1 = (c1[u] < ncl) ? x + c1[u] : xg + (c1[u] - ncl);
        x0[u] = __ldg(p0);
        x1[u] = __ldg(p1);
      }
#pragma unroll
      for (int u = 0; u < SELL_CHUNK; u++) { acc0 = fma(v0[u], x0[u], acc0); acc1 = fma(v1[u], x1[u], acc1); }
    }
    const int64_t r0 = s0 * 32 + lane, r1 = s1 * 32 + lane;
    if (r0 < nrows) { if (sigma != 0.0) acc0 -= sigma * x[r0]; y[r0] = acc0; }
    if (two && r1 < nrows) { if (sigma != 0.0) acc1 -= sigma * x[r1]; y[r1] = acc1; }
  }
}

/* ------------------------------------------------------------------------------------------------
 * SELL-32 SpMV, bulk-copy pipeline: the matrix stream (the 12 B per entry that dominate the traffic) is moved by the
 * TMA unit, not by the warps.  Chunks of whole slices are contiguous in sl_col / sl_val, so ONE producer thread per SM
 * keeps a 4-stage shared-memory ring full with three `cp.async.bulk` copies per chunk (slice offsets, columns, values;
 * completion through an mbarrier with complete_tx::bytes) while 16 consumer warps read (col,val) from shared memory, gather
 * x through L1/L2, and write y.  HBM latency of the stream is covered by ~150 KB in flight per SM instead of by
 * per-thread loads that sit two dependent round trips (offset -> entry -> x) away from the arithmetic.
 * ---------------------------------------------------------------------------------------------- */
#define SP_CAP     8192                 /* upper bound of entries per chunk (32 slices of width 8); the stage size of a matrix is
                                           its largest chunk (cap), the ring depth what fits 200 KB: 2..4 stages              */
#define SP_MAXS    32                   /* slices per chunk: 2 per consumer warp, processed together */
#define SP_STAGES  4                    /* maximum ring depth */
#define SP_CWARPS  16
#define SP_THREADS (32 * SP_CWARPS + 32)
#define SP_OFFS    (SP_MAXS + 2)
#define SP_STAGE_BYTES(cap) (SP_OFFS * 8 + (cap) * 12)
#define SP_SMEM_BUDGET (200 * 1024)

__device__ __forceinline__ uint32_t sp_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t sp_try_wait(uint32_t bar, uint32_t parity)
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
__device__ __forceinline__ void sp_wait(uint32_t bar, uint32_t parity) { while (!sp_try_wait(bar, parity)) { } }
__device__ __forceinline__ void sp_bulk(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}

/* two slices of the same compile-time width W from shared memory: no predicates, no width loop; GHOST = the matrix has
   halo columns (otherwise every column index addresses x directly) */
template <int W, bool GHOST>
__device__ __forceinline__ void sp_pair(const int *__restrict__ cp0, const double *__restrict__ vp0, const int *__restrict__ cp1,
                                        const double *__restrict__ vp1, const double *__restrict__ x, const double *__restrict__ xg,
                                        int ncl, double &acc0, double &acc1)
{
  int c0[W], c1[W];
  double v0[W], v1[W], x0[W], x1[W];
#pragma unroll
  for (int u = 0; u < W; u++) { c0[u] = cp0[32 * u]; c1[u] = cp1[32 * u]; v0[u] = vp0[32 * u]; v1[u] = vp1[32 * u]; }
#pragma unroll
  for (int u = 0; u < W; u++) {
    if (GHOST) {
      x0[u] = __ldg((c0[u] < ncl) ? x + c0[u] : xg + (c0[u] - ncl));
      x1[u] = __ldg((c1[u] < ncl) ? x + c1[u] : xg + (c1[u] - ncl));
    } else {
      x0[u] = __ldg(x + c0[u]);
      x1[u] = __ldg(x + c1[u]);
    }
  }
#pragma unroll
  for (int u = 0; u < W; u++) { acc0 = fma(v0[u], x0[u], acc0); acc1 = fma(v1[u], x1[u], acc1); }
}

/* MULTI: a block of kcols vectors (column-major, leading dimensions ldx / ldg / ldy): the matrix chunk staged in shared memory is
   multiplied by every column before the stage is released, so the matrix is read from HBM ONCE for the whole block
   (BVMatMult in BV_MATMULT_MAT mode, svec.c:203-231).  MULTI = false is the SpMV: kcols = 1 at compile time. */
template <bool GHOST, bool MULTI>
__global__ void __launch_bounds__(SP_THREADS, 1)
k_spmv_sell_pipe(const int *__restrict__ chunk, int nchunks, const int64_t *__restrict__ sl_off, const int *__restrict__ col,
                 const double *__restrict__ val, const double *__restrict__ xb, const double *__restrict__ xgb, int ncl,
                 double *__restrict__ yb, int64_t nrows, double sigma, int cap, int nstages, int kcols_, int64_t ldx, int64_t ldg, int64_t ldy)
{
  b2k_pdl_enter();
  const int kcols = MULTI ? kcols_ : 1;
  extern __shared__ __align__(128) unsigned char sp_raw[];
  const size_t stage_bytes = SP_STAGE_BYTES((size_t)cap);
  unsigned long long *full = reinterpret_cast<unsigned long long *>(sp_raw);
  unsigned long long *empty = full + SP_STAGES;
  volatile int *meta = reinterpret_cast<volatile int *>(empty + SP_STAGES);      /* [stage][2]: first slice, slice count */
  unsigned char *stages = sp_raw + 128;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < nstages; s++) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sp_smem_u32(&full[s])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sp_smem_u32(&empty[s])), "r"(SP_CWARPS));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == SP_CWARPS) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      int c = blockIdx.x;
      int s0 = 0, s1 = 0;
      int64_t e0 = 0, e1 = 0;
      if (c < nchunks) { s0 = chunk[c]; s1 = chunk[c + 1]; e0 = sl_off[s0]; e1 = sl_off[s1]; }
      while (c < nchunks) {
        /* look one chunk ahead so that the dependent loads (chunk -> offsets) are off the critical path */
        const int cn = c + gridDim.x;
        int n0 = 0, n1 = 0;
        int64_t f0 = 0, f1 = 0;
        if (cn < nchunks) { n0 = chunk[cn]; n1 = chunk[cn + 1]; f0 = sl_off[n0]; f1 = sl_off[n1]; }
        sp_wait(sp_smem_u32(&empty[s]), ph ^ 1);
        unsigned char *st = stages + (size_t)s * stage_bytes;
        const uint32_t bar = sp_smem_u32(&full[s]);
        const uint32_t nent = (uint32_t)(e1 - e0);
        const uint32_t noff = (uint32_t)(((s1 - s0 + 1) + 1) & ~1);               /* even count: 16-byte multiples */
        meta[2 * s] = s0; meta[2 * s + 1] = s1 - s0;                              /* released by the arrive below */
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(noff * 8u + nent * 12u) : "memory");
        sp_bulk(sp_smem_u32(st), sl_off + s0, noff * 8u, bar);
        if (nent) {                                                               /* a chunk of empty slices moves its offsets only */
          sp_bulk(sp_smem_u32(st + SP_OFFS * 8), col + e0, nent * 4u, bar);
          sp_bulk(sp_smem_u32(st + SP_OFFS * 8 + (size_t)cap * 4), val + e0, nent * 8u, bar);
        }
        if (++s == nstages) { s = 0; ph ^= 1; }
        c = cn; s0 = n0; s1 = n1; e0 = f0; e1 = f1;
      }
    }
    return;
  }

  int s = 0;
  uint32_t ph = 0;
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    sp_wait(sp_smem_u32(&full[s]), ph);
    const int s0 = meta[2 * s], ns = meta[2 * s + 1];
    const unsigned char *st = stages + (size_t)s * stage_bytes;
    const int64_t *offs = reinterpret_cast<const int64_t *>(st);
    const int *scol = reinterpret_cast<const int *>(st + SP_OFFS * 8);
    const double *sval = reinterpret_cast<const double *>(st + SP_OFFS * 8 + (size_t)cap * 4);
    const int64_t base = offs[0];
    for (int j = warp; j < ns; j += 2 * SP_CWARPS) {
      /* two slices per warp at a time (j and j + 16): twice the gathers in flight, and a chunk of 16 or 32 slices keeps
         all 16 warps equally busy */
      const int j1 = j + SP_CWARPS;
      const bool two = j1 < ns;
      const int o0 = (int)(offs[j] - base), width0 = (int)((offs[j + 1] - offs[j]) >> 5);
      const int o1 = two ? (int)(offs[j1] - base) : 0, width1 = two ? (int)((offs[j1 + 1] - offs[j1]) >> 5) : 0;
      const int *cp0 = scol + o0 + lane, *cp1 = scol + o1 + lane;
      const double *vp0 = sval + o0 + lane, *vp1 = sval + o1 + lane;
      for (int jc = 0; jc < kcols; jc++) {
        const double *x = xb + (MULTI ? (int64_t)jc * ldx : 0), *xg = xgb + (MULTI ? (int64_t)jc * ldg : 0);
        double *y = yb + (MULTI ? (int64_t)jc * ldy : 0);
        double acc0 = 0.0, acc1 = 0.0;
        int wmax = max(width0, width1);
        if (two && width0 == width1 && width0 <= 8) {       /* the common case: specialised on the width */
          switch (width0) {
            case 1: sp_pair<1, GHOST>(cp0, vp0, cp1, vp1, x, xg, ncl, acc0, acc1); break;
            case 2: sp_pair<2, GHOST>(cp0, vp0, cp1, vp1, x, xg, ncl, acc0, acc1); break;
            case 3: sp_pair<3, GHOST>(cp0, vp0, cp1, vp1, x, xg, ncl, acc0, acc1); break;
            case 4: sp_pair<4, GHOST>(cp0, vp0, cp1, vp1, x, xg, ncl, acc0, acc1); break;
            case 5: sp_pair<5, GHOST>(cp0, vp0, cp1, vp1, x, xg, ncl, acc0, acc1); break;
            case 6: sp_pair<6, GHOST>(cp0, vp0, cp1, vp1, x, xg, ncl, acc0, acc1); break;
            case 7: sp_pair<7, GHOST>(cp0, vp0, cp1, vp1, x, xg, ncl, acc0, acc1); break;
            case 8: sp_pair<8, GHOST>(cp0, vp0, cp1, vp1, x, xg, ncl, acc0, acc1); break;
            default: break;
          }
          wmax = 0;
        }
        for (int w = 0; w < wmax; w += SELL_CHUNK) {
          int c0[SELL_CHUNK], c1[SELL_CHUNK];
          double v0[SELL_CHUNK], v1[SELL_CHUNK], x0[SELL_CHUNK], x1[SELL_CHUNK];
#pragma unroll
          for (int u = 0; u < SELL_CHUNK; u++) {
            const bool on0 = w + u < width0, on1 = w + u < width1;
            c0[u] = on0 ? cp0[32 * (w + u)] : 0;
            v0[u] = on0 ? vp0[32 * (w + u)] : 0.0;
            c1[u] = on1 ? cp1[32 * (w + u)] : 0;
            v1[u] = on1 ? vp1[32 * (w + u)] : 0.0;
          }
#pragma unroll
          for (int u = 0; u < SELL_CHUNK; u++) {
            const double *p0 = (c0[u] < ncl) ? x + c0[u] : xg + (c0[u] - ncl);
            const double *p1 = (c1[u] < ncl) ? x + c1[u] : xg + (c1[u] - ncl);
            x0[u] = __ldg(p0);
            x1[u] = __ldg(p1);
          }
#pragma unroll
          for (int u = 0; u < SELL_CHUNK; u++) { acc0 = fma(v0[u], x0[u], acc0); acc1 = fma(v1[u], x1[u], acc1); }
        }
        const int64_t r0 = (int64_t)(s0 + j) * 32 + lane, r1 = (int64_t)(s0 + j1) * 32 + lane;
        if (r0 < nrows) { if (sigma != 0.0) acc0 -= sigma * x[r0]; y[r0] = acc0; }
        if (two && r1 < nrows) { if (sigma != 0.0) acc1 -= sigma * x[r1]; y[r1] = acc1; }
      }
    }
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sp_smem_u32(&empty[s])) : "memory");
    if (++s == nstages) { s = 0; ph ^= 1; }
  }
}

__global__ void __launch_bounds__(256) k_sell_width(const int *__restrict__ rowptr, int64_t nrows, int64_t nslices, int *__restrict__ width)
{
  const int lane = threadIdx.x & 31;
  const int64_t slice = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (slice >= nslices) return;
  const int64_t row = slice * 32 + lane;
  int len = (row < nrows) ? rowptr[row + 1] - rowptr[row] : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
  if (lane == 0) width[slice] = len;
}

__global__ void __launch_bounds__(256) k_sell_fill(const int *__restrict__ rowptr, const int *__restrict__ colidx, const double *__restrict__ val,
                                                    int64_t nrows, int64_t nslices, const int64_t *__restrict__ sl_off, int *__restrict__ scol,
                                                    double *__restrict__ sval)
{
  const int lane = threadIdx.x & 31;
  const int64_t slice = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (slice >= nslices) return;
  const int64_t off = sl_off[slice];
  const int width = (int)((sl_off[slice + 1] - off) >> 5);
  const int64_t row = slice * 32 + lane;
  int a = 0, len = 0;
  if (row < nrows) { a = rowptr[row]; len = rowptr[row + 1] - a; }
  const int padcol = len > 0 ? colidx[a] : 0;
  for (int w = 0; w < width; w++) {
    const bool in = w < len;
    scol[off + 32 * w + lane] = in ? colidx[a + w] : padcol;
    sval[off + 32 * w + lane] = in ? val[a + w] : 0.0;
  }
}

/* inverse of k_sell_fill: the CSR (col,val) arrays from the SELL-32 copy and the row pointer (the CSR copy is dropped once
   the SELL copy exists; the transpose builders of the host layer ask for it back through b2k_csr_arrays) */
__global__ void __launch_bounds__(256) k_sell_to_csr(const int *__restrict__ rowptr, int64_t nrows, int64_t nslices, const int64_t *__restrict__ sl_off,
                                                      const int *__restrict__ scol, const double *__restrict__ sval, int *__restrict__ colidx,
                                                      double *__restrict__ val)
{
  const int lane = threadIdx.x & 31;
  const int64_t slice = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (slice >= nslices) return;
  const int64_t off = sl_off[slice];
  const int64_t row = slice * 32 + lane;
  if (row >= nrows) return;
  const int a = rowptr[row], len = rowptr[row + 1] - a;
  for (int w = 0; w < len; w++) {
    colidx[a + w] = scol[off + 32 * w + lane];
    val[a + w] = sval[off + 32 * w + lane];
  }
}

static int g_sell_mode = -1;      /* env B2K_SPMV_SELL: 0 never, 1 (default) when padding <= 25 %, 2 always */
static int sell_mode(void)
{
  if (g_sell_mode < 0) {
    const char *e = getenv("B2K_SPMV_SELL");
    g_sell_mode = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
  }
  return g_sell_mode;
}
extern "C" int b2k_spmv_set_sell(int mode) { g_sell_mode = (mode >= 0 && mode <= 2) ? mode : 1; return B2K_OK; }

static int build_sell(b2k_ctx ctx, b2k_csr A)
{
  A->nslices = 0; A->sl_off = NULL; A->sl_col = NULL; A->sl_val = NULL; A->sell_elems = 0; A->sp_chunk = NULL; A->nchunks = 0;
  if (!sell_mode() || A->nrows == 0 || A->nnz == 0) return B2K_OK;
  const int64_t ns = (A->nrows + 31) / 32;
  int *dwidth = NULL;
  CK(cudaMalloc(&dwidth, sizeof(int) * (size_t)ns));
  const unsigned grid = (unsigned)((ns + 7) / 8);
  k_sell_width<<<grid, 256, 0, ctx->stream>>>(A->rowptr, A->nrows, ns, dwidth);
  CKLAUNCH(ctx);
  int *hw = (int *)malloc(sizeof(int) * (size_t)ns);
  int64_t *hoff = (int64_t *)malloc(sizeof(int64_t) * (size_t)(ns + 4));
  if (!hw || !hoff) { free(hw); free(hoff); cudaFree(dwidth); return B2K_ERR_MEM; }
  CK(cudaMemcpyAsync(hw, dwidth, sizeof(int) * (size_t)ns, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  cudaFree(dwidth);
  int64_t tot = 0;
  for (int64_t s = 0; s < ns; s++) { hoff[s] = tot; tot += 32 * (int64_t)hw[s]; }
  hoff[ns] = hoff[ns + 1] = hoff[ns + 2] = hoff[ns + 3] = tot;   /* padding: the pipeline kernel copies an even count */
  /* chunks of whole slices for k_spmv_sell_pipe: even first slice, <= SP_CAP entries, <= SP_MAXS slices */
  int *hchunk = (int *)malloc(sizeof(int) * (size_t)(ns / 2 + 3));
  int nchunks = 0;
  int64_t maxent = 0;
  bool pipe_ok = hchunk != NULL && ns < 2147483000LL;
  for (int64_t s = 0; pipe_ok && s < ns;) {
    hchunk[nchunks++] = (int)s;
    int64_t e = s;
    while (e < ns && e - s < SP_MAXS && hoff[e + 1] - hoff[s] <= SP_CAP) e++;
    if (e < ns) e = s + ((e - s) >= 16 ? ((e - s) & ~(int64_t)15) : ((e - s) & ~(int64_t)1));   /* 16 or 32 slices = equal work for
                                                                       the 16 consumer warps; always an even boundary */
    if (e == s) pipe_ok = false;                                    /* a slice (pair) wider than a stage: no pipeline */
    if (hoff[e] - hoff[s] > maxent) maxent = hoff[e] - hoff[s];
    s = e;
  }
  if (pipe_ok) hchunk[nchunks] = (int)ns;
  free(hw);
  if (sell_mode() == 1 && (double)tot > 1.25 * (double)A->nnz) { free(hoff); free(hchunk); return B2K_OK; }   /* too much padding: stay on CSR-stream */
  size_t fr = 0, to = 0;
  cudaMemGetInfo(&fr, &to);
  if ((double)tot * 12.0 + 8.0 * (double)(ns + 1) > 0.5 * (double)fr) { free(hoff); free(hchunk); return B2K_OK; }   /* keep room for the basis */
  CK(cudaMalloc(&A->sl_off, sizeof(int64_t) * (size_t)(ns + 4)));
  CK(cudaMalloc(&A->sl_col, sizeof(int) * (size_t)tot));
  CK(cudaMalloc(&A->sl_val, sizeof(double) * (size_t)tot));
  CK(cudaMemcpyAsync(A->sl_off, hoff, sizeof(int64_t) * (size_t)(ns + 4), cudaMemcpyHostToDevice, ctx->stream));
  if (pipe_ok && nchunks > 0) {
    CK(cudaMalloc(&A->sp_chunk, sizeof(int) * (size_t)(nchunks + 1)));
    CK(cudaMemcpyAsync(A->sp_chunk, hchunk, sizeof(int) * (size_t)(nchunks + 1), cudaMemcpyHostToDevice, ctx->stream));
    A->nchunks = nchunks;
    A->sp_cap = (int)((maxent + 31) & ~(int64_t)31);
  }
  k_sell_fill<<<grid, 256, 0, ctx->stream>>>(A->rowptr, A->colidx, A->val, A->nrows, ns, A->sl_off, A->sl_col, A->sl_val);
  CKLAUNCH(ctx);
  CK(cudaStreamSynchronize(ctx->stream));
  free(hoff);
  free(hchunk);
  A->nslices = ns; A->sell_elems = tot;
  /* the matrix is stored once: the SELL copy is what every product reads, so the CSR (col,val) arrays go (the row pointer
     stays: 4 B/row); b2k_csr_arrays brings them back on demand.  env B2K_CSR_KEEP=1 keeps both (debugging). */
  const char *keep = getenv("B2K_CSR_KEEP");
  if (!(keep && keep[0] == '1')) {
    cudaFree(A->colidx); cudaFree(A->val);
    A->colidx = NULL; A->val = NULL;
    A->csr_dropped = 1;
  }
  return B2K_OK;
}

/* bring the CSR (col,val) arrays back when they were dropped */
static int csr_restore(b2k_csr A)
{
  if (!A->csr_dropped) return B2K_OK;
  b2k_ctx ctx = A->ctx;
  CK(cudaMalloc(&A->colidx, sizeof(int) * (size_t)(A->nnz ? A->nnz : 1)));
  CK(cudaMalloc(&A->val, sizeof(double) * (size_t)(A->nnz ? A->nnz : 1)));
  const unsigned grid = (unsigned)((A->nslices + 7) / 8);
  k_sell_to_csr<<<grid, 256, 0, ctx->stream>>>(A->rowptr, A->nrows, A->nslices, A->sl_off, A->sl_col, A->sl_val, A->colidx, A->val);
  CKLAUNCH(ctx);
  A->csr_dropped = 0;
  return B2K_OK;
}

__global__ void __launch_bounds__(SPMV_THREADS) k_spmv_csr_stream(const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                                   const double *__restrict__ val, const int *__restrict__ blkrow,
                                                                   const double *__restrict__ x, const double *__restrict__ xg,
                                                                   int ncl, double *__restrict__ y, double sigma)
{
  __shared__ double prod[SPMV_CAP];
  __shared__ double red[SPMV_THREADS / 32];
  const int b = blockIdx.x;
  const int r0 = blkrow[b], r1 = blkrow[b + 1];
  const int k0 = rowptr[r0], k1 = rowptr[r1];
  if (k1 - k0 > SPMV_CAP) {
    /* a single long row (r1 == r0+1 by construction): block-wide reduction */
    double s = 0.0;
    for (int k = k0 + threadIdx.x; k < k1; k += SPMV_THREADS) {
      const int c = colidx[k];
      const double xv = (c < ncl) ? x[c] : xg[c - ncl];
      s = fma(val[k], xv, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int i = 0; i < SPMV_THREADS / 32; i++) t += red[i];
      if (sigma != 0.0) t -= sigma * x[r0];
      y[r0] = t;
    }
    return;
  }
  for (int k = k0 + threadIdx.x; k < k1; k += SPMV_THREADS) {
    const int c = colidx[k];
    const double xv = (c < ncl) ? __ldg(x + c) : __ldg(xg + (c - ncl));
    prod[k - k0] = val[k] * xv;
  }
  __syncthreads();
  for (int r = r0 + threadIdx.x; r < r1; r += SPMV_THREADS) {
    const int a = rowptr[r] - k0, e = rowptr[r + 1] - k0;
    double s = 0.0;
    for (int k = a; k < e; k++) s += prod[k];
    if (sigma != 0.0) s -= sigma * x[r];
    y[r] = s;
  }
}

static int build_blocks(int64_t nrows, const int *rowptr, int **blk_out, int *nblk_out)
{
  int cap = 1024, nb = 0;
  int *blk = (int *)malloc(sizeof(int) * (cap + 1));
  if (!blk) return B2K_ERR_MEM;
  int64_t r = 0;
  while (r < nrows) {
    if (nb + 1 >= cap) {
      cap *= 2;
      int *t = (int *)realloc(blk, sizeof(int) * (cap + 1));
      if (!t) { free(blk); return B2K_ERR_MEM; }
      blk = t;
    }
    blk[nb++] = (int)r;
    int64_t e = r + 1;  /* always take at least one row */
    const int base = rowptr[r];
    while (e < nrows && e - r < SPMV_MAXROWS && rowptr[e + 1] - base <= SPMV_CAP) e++;
    r = e;
  }
  blk[nb] = (int)nrows;
  *blk_out = blk;
  *nblk_out = nb;
  return B2K_OK;
}

/* the row blocks of k_spmv_csr_stream: a host walk over the row pointer (0.1 s for 1.7e7 rows), so they are built only when that
   kernel is going to run — a matrix with a SELL copy never needs them (b2k_csr_create_global / b2k_csr_adopt do not even bring the
   row pointer to the host) */
static int ensure_blocks(b2k_csr A, const int *rowptr_host)
{
  if (A->blkrow) return B2K_OK;
  b2k_ctx ctx = A->ctx;
  int *rp = NULL;
  if (!rowptr_host) {
    rp = (int *)malloc(sizeof(int) * (size_t)(A->nrows + 1));
    if (!rp) return B2K_ERR_MEM;
    CK(cudaMemcpyAsync(rp, A->rowptr, sizeof(int) * (size_t)(A->nrows + 1), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    rowptr_host = rp;
  }
  int *blk = NULL, nb = 0;
  int rc = build_blocks(A->nrows, rowptr_host, &blk, &nb);
  free(rp);
  if (rc) return rc;
  A->nblk = nb;
  CK(cudaMalloc(&A->blkrow, sizeof(int) * (size_t)(nb + 1)));
  CK(cudaMemcpyAsync(A->blkrow, blk, sizeof(int) * (size_t)(nb + 1), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  free(blk);
  return B2K_OK;
}

static int finish_create(b2k_ctx ctx, b2k_csr A, const int *rowptr_host)
{
  const int rc = build_sell(ctx, A);
  if (rc) return rc;
  if (A->nslices > 0) return B2K_OK;               /* the products read the SELL copy */
  return ensure_blocks(A, rowptr_host);
}

extern "C" int b2k_csr_create(b2k_ctx ctx, int64_t nrows, int64_t ncols_local, int64_t nghost, const int *rowptr_host,
                              const int *colidx_host, const double *val_host, b2k_csr *out)
{
  ARGCHK(nrows >= 0 && nrows < 2147483647LL, "nrows must fit int32");
  ARGCHK(ncols_local + nghost < 2147483647LL, "column count must fit int32");
  CK(cudaSetDevice(ctx->device));
  b2k_csr A = (b2k_csr)calloc(1, sizeof(*A));
  if (!A) return B2K_ERR_MEM;
  A->nrows = nrows; A->ncols_local = ncols_local; A->nghost = nghost; A->ctx = ctx;
  A->nnz = nrows ? rowptr_host[nrows] : 0;
  CK(cudaMalloc(&A->rowptr, sizeof(int) * (size_t)(nrows + 1)));
  CK(cudaMalloc(&A->colidx, sizeof(int) * (size_t)(A->nnz ? A->nnz : 1)));
  CK(cudaMalloc(&A->val, sizeof(double) * (size_t)(A->nnz ? A->nnz : 1)));
  CK(cudaMemcpyAsync(A->rowptr, rowptr_host, sizeof(int) * (size_t)(nrows + 1), cudaMemcpyHostToDevice, ctx->stream));
  if (A->nnz) {
    CK(cudaMemcpyAsync(A->colidx, colidx_host, sizeof(int) * (size_t)A->nnz, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(A->val, val_host, sizeof(double) * (size_t)A->nnz, cudaMemcpyHostToDevice, ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->h2d_bytes += sizeof(int) * (size_t)(nrows + 1) + (sizeof(int) + sizeof(double)) * (size_t)A->nnz;
  int rc = finish_create(ctx, A, rowptr_host);
  if (rc) return rc;
  *out = A;
  return B2K_OK;
}

extern "C" int b2k_csr_adopt(b2k_ctx ctx, int64_t nrows, int64_t ncols_local, int64_t nghost, int64_t nnz, int *rowptr,
                             int *colidx, double *val, b2k_csr *out)
{
  ARGCHK(nrows >= 0 && nrows < 2147483647LL && nnz < 2147483647LL, "sizes must fit int32");
  b2k_csr A = (b2k_csr)calloc(1, sizeof(*A));
  if (!A) return B2K_ERR_MEM;
  A->nrows = nrows; A->ncols_local = ncols_local; A->nghost = nghost; A->nnz = nnz; A->ctx = ctx;
  A->rowptr = rowptr; A->colidx = colidx; A->val = val;
  int rc = finish_create(ctx, A, NULL);
  if (rc) return rc;
  *out = A;
  return B2K_OK;
}

extern "C" int b2k_csr_destroy(b2k_ctx ctx, b2k_csr A)
{
  if (!A) return B2K_OK;
  CK(cudaStreamSynchronize(ctx->stream));
  cudaFree(A->rowptr); cudaFree(A->colidx); cudaFree(A->val); cudaFree(A->blkrow);
  cudaFree(A->sl_off); cudaFree(A->sl_col); cudaFree(A->sl_val); cudaFree(A->sp_chunk);
  free(A);
  return B2K_OK;
}

extern "C" int b2k_csr_info(b2k_csr A, int64_t *nrows, int64_t *ncl, int64_t *nghost, int64_t *nnz)
{
  if (nrows) *nrows = A->nrows;
  if (ncl) *ncl = A->ncols_local;
  if (nghost) *nghost = A->nghost;
  if (nnz) *nnz = A->nnz;
  return B2K_OK;
}

extern "C" int b2k_csr_arrays(b2k_csr A, int **rowptr, int **colidx, double **val)
{
  if (colidx || val) { const int rc = csr_restore(A); if (rc) return rc; }
  if (rowptr) *rowptr = A->rowptr;
  if (colidx) *colidx = A->colidx;
  if (val) *val = A->val;
  return B2K_OK;
}

/* diag[r] = A(r, r + diag_col_offset) for the local rows (0 where the entry is not stored): MatGetDiagonal for the Jacobi
   preconditioner of the shift-and-invert ST.  Works on whichever copy exists (SELL first). */
__global__ void __launch_bounds__(256) k_sell_diag(const int64_t *__restrict__ sl_off, const int *__restrict__ col, const double *__restrict__ val,
                                                    int64_t nrows, int64_t nslices, int coloff, double *__restrict__ diag)
{
  const int lane = threadIdx.x & 31;
  const int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (s >= nslices) return;
  const int64_t off = sl_off[s], r = s * 32 + lane;
  const int width = (int)((sl_off[s + 1] - off) >> 5);
  double d = 0.0;
  for (int w = 0; w < width; w++) {
    const int c = col[off + 32 * w + lane];
    const double v = val[off + 32 * w + lane];
    if ((int64_t)c == r + coloff && v != 0.0) d = v;        /* padding repeats a real column with value 0 */
  }
  if (r < nrows) diag[r] = d;
}
__global__ void __launch_bounds__(256) k_csr_diag(const int *__restrict__ rowptr, const int *__restrict__ colidx, const double *__restrict__ val,
                                                   int64_t nrows, int coloff, double *__restrict__ diag)
{
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  double d = 0.0;
  for (int k = rowptr[r]; k < rowptr[r + 1]; k++) if ((int64_t)colidx[k] == r + coloff) d = val[k];
  diag[r] = d;
}
extern "C" int b2k_csr_get_diagonal(b2k_ctx ctx, b2k_csr A, int64_t diag_col_offset, double *diag)
{
  if (A->nrows == 0) return B2K_OK;
  if (A->nslices > 0) k_sell_diag<<<(unsigned)((A->nslices + 7) / 8), 256, 0, ctx->stream>>>(A->sl_off, A->sl_col, A->sl_val, A->nrows, A->nslices, (int)diag_col_offset, diag);
  else k_csr_diag<<<(unsigned)((A->nrows + 255) / 256), 256, 0, ctx->stream>>>(A->rowptr, A->colidx, A->val, A->nrows, (int)diag_col_offset, diag);
  CKLAUNCH(ctx);
  return B2K_OK;
}

/* the caller is done with the arrays b2k_csr_arrays lent: drop the CSR copy again when a SELL copy serves the products */
extern "C" int b2k_csr_release_arrays(b2k_csr A)
{
  if (!A || A->csr_dropped || A->nslices == 0) return B2K_OK;
  const char *keep = getenv("B2K_CSR_KEEP");
  if (keep && keep[0] == '1') return B2K_OK;
  CK(cudaStreamSynchronize(A->ctx->stream));
  cudaFree(A->colidx); cudaFree(A->val);
  A->colidx = NULL; A->val = NULL;
  A->csr_dropped = 1;
  return B2K_OK;
}

/* HBM bytes held by the matrix (all copies) */
extern "C" int b2k_csr_bytes(b2k_csr A, int64_t *bytes)
{
  int64_t b = 4 * (A->nrows + 1) + 4 * (int64_t)(A->nblk + 1);
  if (!A->csr_dropped) b += 12 * (A->nnz ? A->nnz : 1);
  if (A->nslices) b += 12 * A->sell_elems + 8 * (A->nslices + 4) + 4 * (int64_t)(A->nchunks + 1);
  *bytes = b;
  return B2K_OK;
}

extern "C" int b2k_csr_last_kernel(b2k_csr A, int *which)
{
  *which = A ? A->last_kernel : 0;
  return B2K_OK;
}

static int g_pipe_min_chunks = -2;   /* -2: read env B2K_SPMV_PIPE_MIN_CHUNKS once; -1: default (4 chunks per SM); >= 0: explicit */
extern "C" int b2k_spmv_set_pipe_min_chunks(int min_chunks)
{
  g_pipe_min_chunks = min_chunks < 0 ? -1 : min_chunks;
  return B2K_OK;
}

static int g_pipe_mode = -1;          /* env B2K_SPMV_PIPE: 1 (default) bulk-copy pipeline over the SELL copy, 0 plain SELL kernel */
/* ring depth of the bulk-copy pipeline for this matrix, 0 = the pipeline does not take it (too few chunks, no SELL copy, switched off) */
static int sp_pipe_stages(b2k_ctx ctx, b2k_csr A, size_t *shm)
{
  if (g_pipe_mode < 0) {
    const char *e = getenv("B2K_SPMV_PIPE");
    g_pipe_mode = (e && e[0] == '0') ? 0 : 1;
    if (g_pipe_mode) {
      cudaError_t ce = cudaFuncSetAttribute(k_spmv_sell_pipe<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 + SP_SMEM_BUDGET);
      if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_spmv_sell_pipe<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 + SP_SMEM_BUDGET);
      if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_spmv_sell_pipe<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 + SP_SMEM_BUDGET);
      if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_spmv_sell_pipe<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 + SP_SMEM_BUDGET);
      if (ce != cudaSuccess) { cudaGetLastError(); g_pipe_mode = 0; }
    }
  }
  if (g_pipe_min_chunks == -2) {
    const char *e = getenv("B2K_SPMV_PIPE_MIN_CHUNKS");
    g_pipe_min_chunks = e ? atoi(e) : -1;
  }
  const int min_chunks = g_pipe_min_chunks >= 0 ? (g_pipe_min_chunks > 0 ? g_pipe_min_chunks : 1) : 4 * ctx->sm_count;
  int st = 0;
  if (A->nslices > 0 && sell_mode() && g_pipe_mode && A->nchunks >= min_chunks) {
    st = (int)(SP_SMEM_BUDGET / SP_STAGE_BYTES((size_t)A->sp_cap));
    if (st > SP_STAGES) st = SP_STAGES;
    *shm = 128 + (size_t)st * SP_STAGE_BYTES((size_t)A->sp_cap);
  }
  return st >= 2 ? st : 0;
}

extern "C" int b2k_csr_spmv_shift(b2k_ctx ctx, b2k_csr A, const double *x, const double *xghost, double *y, double sigma)
{
  if (A->nrows == 0) return B2K_OK;
  ARGCHK(x != y, "SpMV cannot run in place");
  if (!(A->nslices > 0 && sell_mode())) {                                           /* CSR-stream (no SELL copy, or forced) */
    if (A->csr_dropped) { const int rc = csr_restore(A); if (rc) return rc; }
    { const int rc = ensure_blocks(A, NULL); if (rc) return rc; }
  }
  /* algorithmic bytes of the CSR product whichever storage runs: the SELL copy moves 12 B per stored
     entry (padding included) and no row pointers */
  PROF_BEGIN(ctx, B2K_PROF_SPMV, 12.0 * (double)A->nnz + 4.0 * (double)(A->nrows + 1) + 8.0 * (double)(A->ncols_local + A->nghost) + 8.0 * (double)A->nrows);
  size_t sp_shm = 0;
  const int sp_stages = sp_pipe_stages(ctx, A, &sp_shm);
  A->last_kernel = (sp_stages >= 2) ? (A->nghost > 0 ? B2K_SPMV_KERNEL_SELL_PIPE_GHOST : B2K_SPMV_KERNEL_SELL_PIPE)
                                     : ((A->nslices > 0 && sell_mode()) ? B2K_SPMV_KERNEL_SELL : B2K_SPMV_KERNEL_CSR_STREAM);
  if (sp_stages >= 2 && A->nghost > 0)
    b2k_launch_pdl(k_spmv_sell_pipe<true, false>, dim3(ctx->sm_count), dim3(SP_THREADS), sp_shm, ctx->stream,
                   A->sp_chunk, A->nchunks, A->sl_off, A->sl_col, A->sl_val, x, xghost ? xghost : x, (int)A->ncols_local, y, A->nrows, sigma,
                   A->sp_cap, sp_stages, 1, (int64_t)0, (int64_t)0, (int64_t)0);
  else if (sp_stages >= 2)
    b2k_launch_pdl(k_spmv_sell_pipe<false, false>, dim3(ctx->sm_count), dim3(SP_THREADS), sp_shm, ctx->stream,
                   A->sp_chunk, A->nchunks, A->sl_off, A->sl_col, A->sl_val, x, x, (int)A->ncols_local, y, A->nrows, sigma, A->sp_cap, sp_stages,
                   1, (int64_t)0, (int64_t)0, (int64_t)0);
  else if (A->nslices > 0 && sell_mode())
    b2k_launch_pdl(k_spmv_sell, dim3((unsigned)std::min<int64_t>((A->nslices + 15) / 16, (int64_t)ctx->sm_count * 8)), dim3(256), 0, ctx->stream,
                   A->sl_off, A->sl_col, A->sl_val, x, xghost ? xghost : x, (int)A->ncols_local, y, A->nrows, A->nslices, sigma);
  else
    k_spmv_csr_stream<<<A->nblk, SPMV_THREADS, 0, ctx->stream>>>(A->rowptr, A->colidx, A->val, A->blkrow, x,
                                                                   xghost ? xghost : x, (int)A->ncols_local, y, sigma);
  PROF_END(ctx);
  CKLAUNCH(ctx);
  return B2K_OK;
}
/* ------------------------------------------------------------------------------------------------
 * SELL-32 SpMM: Y(:,0:k) = A [X(:,0:k) ; XG(:,0:k)] for a block of k vectors, the matrix read once per tile of KT columns
 * instead of once per column (BVMatMult with BV_MATMULT_MAT, svec.c:203-231: PETSc MatMatMult on the dense block).
 * One warp per slice, lane = row: the (col,val) entries of a row pass through registers in chunks of 8 and feed KT
 * accumulators; the gathers of the KT columns of X hit the same L1/L2 lines of neighbouring rows.  Algorithmic bytes
 * 12 nnz ceil(k/KT) + 16 n k against 12 nnz k + 16 n k for the column loop.
 * ---------------------------------------------------------------------------------------------- */
template <int KT>
__global__ void __launch_bounds__(256) k_spmm_sell(const int64_t *__restrict__ sl_off, const int *__restrict__ col, const double *__restrict__ val,
                                                    const double *__restrict__ X, int64_t ldx, const double *__restrict__ XG, int64_t ldg, int ncl,
                                                    double *__restrict__ Y, int64_t ldy, int64_t nrows, int64_t nslices, int k0, int k)
{
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int kt = min(KT, k - k0);
  const double *Xb = X + (int64_t)k0 * ldx, *Gb = XG ? XG + (int64_t)k0 * ldg : X;
  for (int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < nslices; s += nwarps) {
    const int64_t off = sl_off[s];
    const int width = (int)((sl_off[s + 1] - off) >> 5);
    const int *cp = col + off + lane;
    const double *vp = val + off + lane;
    double acc[KT];
#pragma unroll
    for (int j = 0; j < KT; j++) acc[j] = 0.0;
    for (int w = 0; w < width; w += SELL_CHUNK) {
      int c[SELL_CHUNK];
      double v[SELL_CHUNK];
#pragma unroll
      for (int u = 0; u < SELL_CHUNK; u++) {
        const bool on = w + u < width;
        c[u] = on ? ld_stream_i32(cp + 32 * (w + u)) : 0;
        v[u] = on ? ld_stream_f64(vp + 32 * (w + u)) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < SELL_CHUNK; u++) {
        const bool own = c[u] < ncl;
        const double *p = own ? Xb + c[u] : Gb + (c[u] - ncl);
        const int64_t st = own ? ldx : ldg;
#pragma unroll
        for (int j = 0; j < KT; j++)
          if (j < kt) acc[j] = fma(v[u], __ldg(p + (int64_t)j * st), acc[j]);
      }
    }
    const int64_t r = s * 32 + lane;
    if (r < nrows) {
#pragma unroll
      for (int j = 0; j < KT; j++)
        if (j < kt) Y[(int64_t)(k0 + j) * ldy + r] = acc[j];
    }
  }
}

/* Y(:,0:k) = A [X ; XG](:,0:k): block product over the SELL copy (any matrix that has one), column loop otherwise */
extern "C" int b2k_csr_spmm(b2k_ctx ctx, b2k_csr A, const double *X, int64_t ldx, const double *XG, int64_t ldg, double *Y, int64_t ldy, int k)
{
  ARGCHK(k >= 0, "negative column count");
  if (A->nrows == 0 || k == 0) return B2K_OK;
  ARGCHK(X != Y, "SpMM cannot run in place");
  ARGCHK(A->nghost == 0 || XG, "the matrix has ghost columns: a ghost block is needed");
  if (!(A->nslices > 0 && sell_mode())) {
    for (int j = 0; j < k; j++) {
      const int rc = b2k_csr_spmv_shift(ctx, A, X + (int64_t)j * ldx, XG ? XG + (int64_t)j * ldg : NULL, Y + (int64_t)j * ldy, 0.0);
      if (rc) return rc;
    }
    return B2K_OK;
  }
  size_t sp_shm = 0;
  const int sp_stages = sp_pipe_stages(ctx, A, &sp_shm);
  if (sp_stages) {
    /* large matrices: the bulk-copy pipeline with the column loop inside the consumers — the matrix streams through shared
       memory once for the whole block */
    PROF_BEGIN(ctx, B2K_PROF_SPMV, 12.0 * (double)A->nnz + 8.0 * (double)(A->ncols_local + A->nghost) * k + 8.0 * (double)A->nrows * k);
    if (A->nghost > 0)
      b2k_launch_pdl(k_spmv_sell_pipe<true, true>, dim3(ctx->sm_count), dim3(SP_THREADS), sp_shm, ctx->stream, A->sp_chunk, A->nchunks, A->sl_off,
                     A->sl_col, A->sl_val, X, XG, (int)A->ncols_local, Y, A->nrows, 0.0, A->sp_cap, sp_stages, k, ldx, ldg, ldy);
    else
      b2k_launch_pdl(k_spmv_sell_pipe<false, true>, dim3(ctx->sm_count), dim3(SP_THREADS), sp_shm, ctx->stream, A->sp_chunk, A->nchunks, A->sl_off,
                     A->sl_col, A->sl_val, X, X, (int)A->ncols_local, Y, A->nrows, 0.0, A->sp_cap, sp_stages, k, ldx, (int64_t)0, ldy);
    PROF_END(ctx);
    CKLAUNCH(ctx);
    A->last_kernel = B2K_SPMV_KERNEL_SPMM_PIPE;
    return B2K_OK;
  }
  /* small matrices: one launch per 16 columns instead of one per column */
  const unsigned grid = (unsigned)std::min<int64_t>((A->nslices + 7) / 8, (int64_t)ctx->sm_count * 8);
  const int KT = (k > 8) ? 16 : 8;
  const int passes = (k + KT - 1) / KT;
  PROF_BEGIN(ctx, B2K_PROF_SPMV, 12.0 * (double)A->nnz * passes + 8.0 * (double)(A->ncols_local + A->nghost) * k + 8.0 * (double)A->nrows * k);
  for (int k0 = 0; k0 < k; k0 += KT) {
    if (KT == 16) k_spmm_sell<16><<<grid, 256, 0, ctx->stream>>>(A->sl_off, A->sl_col, A->sl_val, X, ldx, XG, ldg, (int)A->ncols_local, Y, ldy, A->nrows, A->nslices, k0, k);
    else k_spmm_sell<8><<<grid, 256, 0, ctx->stream>>>(A->sl_off, A->sl_col, A->sl_val, X, ldx, XG, ldg, (int)A->ncols_local, Y, ldy, A->nrows, A->nslices, k0, k);
    ctx->launches++;
  }
  PROF_END(ctx);
  cudaError_t e_ = cudaGetLastError();
  if (e_ != cudaSuccess) { b2k_set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_)); return B2K_ERR_CUDA; }
  A->last_kernel = B2K_SPMV_KERNEL_SPMM;
  return B2K_OK;
}

extern "C" int b2k_csr_spmv(b2k_ctx ctx, b2k_csr A, const double *x, const double *xghost, double *y)
{
  return b2k_csr_spmv_shift(ctx, A, x, xghost, y, 0.0);
}

__global__ void __launch_bounds__(256) k_gather(double *__restrict__ out, const double *__restrict__ x, const int *__restrict__ idx,
                                                 int64_t n)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = x[idx[i]];
}
extern "C" int b2k_gather(b2k_ctx ctx, double *out, const double *x, const int *idx, int64_t count)
{
  if (count == 0) return B2K_OK;
  int64_t g = (count + 255) / 256;
  if (g > ctx->sm_count * 8) g = ctx->sm_count * 8;
  k_gather<<<(unsigned)g, 256, 0, ctx->stream>>>(out, x, idx, count);
  CKLAUNCH(ctx);
  return B2K_OK;
}

/* out[idx[i]] += in[i]: accumulation of the reverse halo (A^T products); idx entries are unique within one call */
__global__ void __launch_bounds__(256) k_scatter_add(double *__restrict__ out, const int *__restrict__ idx, const double *__restrict__ in,
                                                      int64_t n)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[idx[i]] += in[i];
}
extern "C" int b2k_scatter_add(b2k_ctx ctx, double *out, const int *idx, const double *in, int64_t count)
{
  if (count == 0) return B2K_OK;
  int64_t g = (count + 255) / 256;
  if (g > ctx->sm_count * 8) g = ctx->sm_count * 8;
  k_scatter_add<<<(unsigned)g, 256, 0, ctx->stream>>>(out, idx, in, count);
  CKLAUNCH(ctx);
  return B2K_OK;
}

/* ------------------------------------------------------------------------------------------------
 * Device generator for the d-dimensional Laplacian stencil (ex1.c:37-48, ex2.c:39-54 and the 7-point
 * analogue): grid nx (slowest) x ny x nz (fastest; nz=1 for 2-D, ny=nz=1 for 1-D), natural ordering,
 * diagonal 2*dim, off-diagonals -1.  This rank owns global rows [row0,row0+nrows) which must be whole
 * planes of the slowest index when more than one GPU is used.  Column indices are LOCAL: owned
 * columns first, then ghosts [lower plane | upper plane].
 * ---------------------------------------------------------------------------------------------- */
__device__ __forceinline__ int lap_row_count(int dim, int64_t nx, int64_t ny, int64_t nz, int64_t g)
{
  const int64_t plane = ny * nz;
  const int64_t i = g / plane, rem = g - i * plane, j = rem / nz, kk = rem - j * nz;
  int c = 1 + (i > 0) + (i < nx - 1);
  if (dim >= 2) c += (j > 0) + (j < ny - 1);
  if (dim >= 3) c += (kk > 0) + (kk < nz - 1);
  return c;
}
__global__ void k_lap_count(int dim, int64_t nx, int64_t ny, int64_t nz, int64_t row0, int64_t nrows, int *__restrict__ cnt)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += stride)
    cnt[r] = lap_row_count(dim, nx, ny, nz, row0 + r);
}
/* closed-form row pointer: sum of counts of rows before r.  Cheap formula: each row has 2*dim+1
   entries minus one per boundary face it touches; we instead run an exclusive scan on the device. */
__global__ void k_scan_block(const int *__restrict__ in, int *__restrict__ out, int64_t n, int *__restrict__ blocksum)
{
  /* 1024 elements per block, Hillis-Steele in shared memory */
  __shared__ int s[1024];
  const int64_t base = (int64_t)blockIdx.x * 1024;
  const int t = threadIdx.x;
  s[t] = (base + t < n) ? in[base + t] : 0;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    int v = (t >= o) ? s[t - o] : 0;
    __syncthreads();
    s[t] += v;
    __syncthreads();
  }
  if (base + t < n) out[base + t + 1] = s[t];   /* inclusive → shifted by one = exclusive rowptr */
  if (t == 1023) blocksum[blockIdx.x] = s[t];
}
__global__ void k_scan_add(int *__restrict__ out, int64_t n, const int *__restrict__ blockoff)
{
  const int64_t base = (int64_t)blockIdx.x * 1024;
  const int t = threadIdx.x;
  if (base + t < n) out[base + t + 1] += blockoff[blockIdx.x];
}
__global__ void k_lap_fill(int dim, int64_t nx, int64_t ny, int64_t nz, int64_t row0, int64_t nrows, const int *__restrict__ rowptr,
                           int *__restrict__ col, double *__restrict__ val)
{
  const int64_t plane = ny * nz;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t nloc = nrows;
  const bool has_lo = row0 > 0;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += stride) {
    const int64_t g = row0 + r;
    const int64_t i = g / plane, rem = g - i * plane, j = rem / nz, kk = rem - j * nz;
    int p = rowptr[r];
    /* ascending global column order: -plane, -nz, -1, 0, +1, +nz, +plane */
    if (i > 0) {
      const int64_t lc = r - plane;     /* local index if owned */
      col[p] = (lc >= 0) ? (int)lc : (int)(nloc + (lc + plane));   /* lower ghost plane */
      val[p++] = -1.0;
    }
    if (dim >= 2 && j > 0) { col[p] = (int)(r - nz); val[p++] = -1.0; }
    if (dim >= 3 && kk > 0) { col[p] = (int)(r - 1); val[p++] = -1.0; }
    col[p] = (int)r; val[p++] = 2.0 * dim;
    if (dim >= 3 && kk < nz - 1) { col[p] = (int)(r + 1); val[p++] = -1.0; }
    if (dim >= 2 && j < ny - 1) { col[p] = (int)(r + nz); val[p++] = -1.0; }
    if (i < nx - 1) {
      const int64_t lc = r + plane;
      col[p] = (lc < nloc) ? (int)lc : (int)(nloc + (has_lo ? plane : 0) + (lc - nloc));   /* upper ghost plane */
      val[p++] = -1.0;
    }
  }
}

extern "C" int b2k_csr_laplacian(b2k_ctx ctx, int dim, int64_t nx, int64_t ny, int64_t nz, int64_t row0, int64_t nrows, b2k_csr *out,
                                 int64_t *nghost_lo, int64_t *nghost_hi)
{
  ARGCHK(dim >= 1 && dim <= 3, "dim must be 1,2,3");
  if (dim < 3) nz = 1;
  if (dim < 2) ny = 1;
  const int64_t plane = ny * nz, N = nx * plane;
  ARGCHK(row0 >= 0 && row0 + nrows <= N, "row range outside the grid");
  ARGCHK(nrows > 0 && nrows < 2147483647LL / 8, "local row count out of range");
  ARGCHK(row0 % plane == 0 && nrows % plane == 0, "slab partition must own whole planes of the slowest index");
  const int64_t glo = (row0 > 0) ? plane : 0, ghi = (row0 + nrows < N) ? plane : 0;
  CK(cudaSetDevice(ctx->device));
  int *cnt = NULL, *rowptr = NULL, *bs = NULL, *col = NULL;
  double *val = NULL;
  const int64_t nb = (nrows + 1023) / 1024;
  CK(cudaMalloc(&cnt, sizeof(int) * (size_t)nrows));
  CK(cudaMalloc(&rowptr, sizeof(int) * (size_t)(nrows + 1)));
  CK(cudaMalloc(&bs, sizeof(int) * (size_t)(nb + 1)));
  int g = ctx->sm_count * 8;
  k_lap_count<<<g, 256, 0, ctx->stream>>>(dim, nx, ny, nz, row0, nrows, cnt);
  CKLAUNCH(ctx);
  CK(cudaMemsetAsync(rowptr, 0, sizeof(int), ctx->stream));
  k_scan_block<<<(unsigned)nb, 1024, 0, ctx->stream>>>(cnt, rowptr, nrows, bs);
  CKLAUNCH(ctx);
  /* block offsets on the host (nb <= 131072 ints) */
  int *hbs = (int *)malloc(sizeof(int) * (size_t)(nb + 1));
  if (!hbs) return B2K_ERR_MEM;
  CK(cudaMemcpyAsync(hbs, bs, sizeof(int) * (size_t)nb, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  int64_t run = 0;
  for (int64_t b = 0; b < nb; b++) { const int v = hbs[b]; hbs[b] = (int)run; run += v; }
  if (run >= 2147483647LL) { free(hbs); b2k_set_error("nnz exceeds int32"); return B2K_ERR_ARG; }
  CK(cudaMemcpyAsync(bs, hbs, sizeof(int) * (size_t)nb, cudaMemcpyHostToDevice, ctx->stream));
  k_scan_add<<<(unsigned)nb, 1024, 0, ctx->stream>>>(rowptr, nrows, bs);
  CKLAUNCH(ctx);
  const int64_t nnz = run;
  CK(cudaMalloc(&col, sizeof(int) * (size_t)nnz));
  CK(cudaMalloc(&val, sizeof(double) * (size_t)nnz));
  k_lap_fill<<<g, 256, 0, ctx->stream>>>(dim, nx, ny, nz, row0, nrows, rowptr, col, val);
  CKLAUNCH(ctx);
  CK(cudaStreamSynchronize(ctx->stream));
  free(hbs);
  cudaFree(cnt);
  cudaFree(bs);
  if (nghost_lo) *nghost_lo = glo;
  if (nghost_hi) *nghost_hi = ghi;
  return b2k_csr_adopt(ctx, nrows, nrows, glo + ghi, nnz, rowptr, col, val, out);
}
