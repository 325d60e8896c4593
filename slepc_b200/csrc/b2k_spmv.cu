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
  int64_t *sl_off;    /* [nslices+1] element offset of each slice; width = (off[s+1]-off[s])/32 */
  int     *sl_col;
  double  *sl_val;
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

__global__ void __launch_bounds__(256) k_spmv_sell(const int64_t *__restrict__ sl_off, const int *__restrict__ col,
                                                    const double *__restrict__ val, const double *__restrict__ x,
                                                    const double *__restrict__ xg, int ncl, double *__restrict__ y, int64_t nrows,
                                                    int64_t nslices, double sigma)
{
  const int lane = threadIdx.x & 31;
  const int64_t slice = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (slice >= nslices) return;
  const int64_t off = sl_off[slice];
  const int width = (int)((sl_off[slice + 1] - off) >> 5);
  const int *cp = col + off + lane;
  const double *vp = val + off + lane;
  double acc = 0.0;
  int w = 0;
  for (; w + 4 <= width; w += 4) {
    const int c0 = ld_stream_i32(cp + 32 * w), c1 = ld_stream_i32(cp + 32 * (w + 1)), c2 = ld_stream_i32(cp + 32 * (w + 2)),
              c3 = ld_stream_i32(cp + 32 * (w + 3));
    const double v0 = ld_stream_f64(vp + 32 * w), v1 = ld_stream_f64(vp + 32 * (w + 1)), v2 = ld_stream_f64(vp + 32 * (w + 2)),
                 v3 = ld_stream_f64(vp + 32 * (w + 3));
    const double x0 = (c0 < ncl) ? __ldg(x + c0) : __ldg(xg + (c0 - ncl));
    const double x1 = (c1 < ncl) ? __ldg(x + c1) : __ldg(xg + (c1 - ncl));
    const double x2 = (c2 < ncl) ? __ldg(x + c2) : __ldg(xg + (c2 - ncl));
    const double x3 = (c3 < ncl) ? __ldg(x + c3) : __ldg(xg + (c3 - ncl));
    acc = fma(v0, x0, acc);
    acc = fma(v1, x1, acc);
    acc = fma(v2, x2, acc);
    acc = fma(v3, x3, acc);
  }
  for (; w < width; w++) {
    const int c = ld_stream_i32(cp + 32 * w);
    const double v = ld_stream_f64(vp + 32 * w);
    acc = fma(v, (c < ncl) ? __ldg(x + c) : __ldg(xg + (c - ncl)), acc);
  }
  const int64_t row = slice * 32 + lane;
  if (row < nrows) {
    if (sigma != 0.0) acc -= sigma * x[row];
    y[row] = acc;
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
  A->nslices = 0; A->sl_off = NULL; A->sl_col = NULL; A->sl_val = NULL; A->sell_elems = 0;
  if (!sell_mode() || A->nrows == 0 || A->nnz == 0) return B2K_OK;
  const int64_t ns = (A->nrows + 31) / 32;
  int *dwidth = NULL;
  CK(cudaMalloc(&dwidth, sizeof(int) * (size_t)ns));
  const unsigned grid = (unsigned)((ns + 7) / 8);
  k_sell_width<<<grid, 256, 0, ctx->stream>>>(A->rowptr, A->nrows, ns, dwidth);
  CKLAUNCH(ctx);
  int *hw = (int *)malloc(sizeof(int) * (size_t)ns);
  int64_t *hoff = (int64_t *)malloc(sizeof(int64_t) * (size_t)(ns + 1));
  if (!hw || !hoff) { free(hw); free(hoff); cudaFree(dwidth); return B2K_ERR_MEM; }
  CK(cudaMemcpyAsync(hw, dwidth, sizeof(int) * (size_t)ns, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  cudaFree(dwidth);
  int64_t tot = 0;
  for (int64_t s = 0; s < ns; s++) { hoff[s] = tot; tot += 32 * (int64_t)hw[s]; }
  hoff[ns] = tot;
  free(hw);
  if (sell_mode() == 1 && (double)tot > 1.25 * (double)A->nnz) { free(hoff); return B2K_OK; }   /* too much padding: stay on CSR-stream */
  size_t fr = 0, to = 0;
  cudaMemGetInfo(&fr, &to);
  if ((double)tot * 12.0 + 8.0 * (double)(ns + 1) > 0.5 * (double)fr) { free(hoff); return B2K_OK; }   /* keep room for the basis */
  CK(cudaMalloc(&A->sl_off, sizeof(int64_t) * (size_t)(ns + 1)));
  CK(cudaMalloc(&A->sl_col, sizeof(int) * (size_t)tot));
  CK(cudaMalloc(&A->sl_val, sizeof(double) * (size_t)tot));
  CK(cudaMemcpyAsync(A->sl_off, hoff, sizeof(int64_t) * (size_t)(ns + 1), cudaMemcpyHostToDevice, ctx->stream));
  k_sell_fill<<<grid, 256, 0, ctx->stream>>>(A->rowptr, A->colidx, A->val, A->nrows, ns, A->sl_off, A->sl_col, A->sl_val);
  CKLAUNCH(ctx);
  CK(cudaStreamSynchronize(ctx->stream));
  free(hoff);
  A->nslices = ns; A->sell_elems = tot;
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

static int finish_create(b2k_ctx ctx, b2k_csr A, const int *rowptr_host)
{
  int *blk = NULL, nb = 0;
  int rc = build_blocks(A->nrows, rowptr_host, &blk, &nb);
  if (rc) return rc;
  A->nblk = nb;
  CK(cudaMalloc(&A->blkrow, sizeof(int) * (size_t)(nb + 1)));
  CK(cudaMemcpyAsync(A->blkrow, blk, sizeof(int) * (size_t)(nb + 1), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  free(blk);
  return build_sell(ctx, A);
}

extern "C" int b2k_csr_create(b2k_ctx ctx, int64_t nrows, int64_t ncols_local, int64_t nghost, const int *rowptr_host,
                              const int *colidx_host, const double *val_host, b2k_csr *out)
{
  ARGCHK(nrows >= 0 && nrows < 2147483647LL, "nrows must fit int32");
  ARGCHK(ncols_local + nghost < 2147483647LL, "column count must fit int32");
  CK(cudaSetDevice(ctx->device));
  b2k_csr A = (b2k_csr)calloc(1, sizeof(*A));
  if (!A) return B2K_ERR_MEM;
  A->nrows = nrows; A->ncols_local = ncols_local; A->nghost = nghost;
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
  A->nrows = nrows; A->ncols_local = ncols_local; A->nghost = nghost; A->nnz = nnz;
  A->rowptr = rowptr; A->colidx = colidx; A->val = val;
  int *rp = (int *)malloc(sizeof(int) * (size_t)(nrows + 1));
  if (!rp) return B2K_ERR_MEM;
  CK(cudaMemcpyAsync(rp, rowptr, sizeof(int) * (size_t)(nrows + 1), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  int rc = finish_create(ctx, A, rp);
  free(rp);
  if (rc) return rc;
  *out = A;
  return B2K_OK;
}

extern "C" int b2k_csr_destroy(b2k_ctx ctx, b2k_csr A)
{
  if (!A) return B2K_OK;
  CK(cudaStreamSynchronize(ctx->stream));
  cudaFree(A->rowptr); cudaFree(A->colidx); cudaFree(A->val); cudaFree(A->blkrow);
  cudaFree(A->sl_off); cudaFree(A->sl_col); cudaFree(A->sl_val);
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
  if (rowptr) *rowptr = A->rowptr;
  if (colidx) *colidx = A->colidx;
  if (val) *val = A->val;
  return B2K_OK;
}

extern "C" int b2k_csr_spmv_shift(b2k_ctx ctx, b2k_csr A, const double *x, const double *xghost, double *y, double sigma)
{
  if (A->nrows == 0) return B2K_OK;
  ARGCHK(x != y, "SpMV cannot run in place");
  /* algorithmic bytes of the CSR product whichever storage runs: the SELL copy moves 12 B per stored
     entry (padding included) and no row pointers */
  PROF_BEGIN(ctx, B2K_PROF_SPMV, 12.0 * (double)A->nnz + 4.0 * (double)(A->nrows + 1) + 8.0 * (double)(A->ncols_local + A->nghost) + 8.0 * (double)A->nrows);
  if (A->nslices > 0 && sell_mode())
    k_spmv_sell<<<(unsigned)((A->nslices + 7) / 8), 256, 0, ctx->stream>>>(A->sl_off, A->sl_col, A->sl_val, x, xghost ? xghost : x,
                                                                            (int)A->ncols_local, y, A->nrows, A->nslices, sigma);
  else
    k_spmv_csr_stream<<<A->nblk, SPMV_THREADS, 0, ctx->stream>>>(A->rowptr, A->colidx, A->val, A->blkrow, x,
                                                                   xghost ? xghost : x, (int)A->ncols_local, y, sigma);
  PROF_END(ctx);
  CKLAUNCH(ctx);
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
