/*
 * b2k_transpose.cu — A_loc^T built in HBM (set-up, once per matrix).
 *
 * The thick-restart Lanczos bidiagonalisation multiplies by A and by A^T in turn (gklanczos.c:67,80,90,103).  With an implicit
 * transpose (svdsetup.c:273-279,309-315) PETSc's MatMultTranspose_MPIAIJ walks the CSR rows and scatters into y; on a GPU that
 * is an atomic scatter, slow and not reproducible.  Here the local transpose is stored instead, split like MPIAIJ's diagonal /
 * off-diagonal blocks: rows of AT_own = the owned columns of A, rows of AT_ghost = its ghost columns, columns of both = the
 * local rows of A.  Products with them are ordinary SELL SpMVs (k_spmv_sell_pipe); only the ghost part travels.
 *
 * Round 1 built it on the host (device -> host copy of the matrix, counting sort, upload: 1 s for 2e7 entries, inside the first
 * MatMultTranspose of the solve).  Now: one stable LSB radix sort of (local column -> entry index) pairs in HBM, then a gather.
 * Stable => inside a transposed row the entries keep the ascending row order of A => the summation order of y = A^T x, and
 * with it every bit of the result, is fixed.  The radix sort is cub::DeviceRadixSort (header-only, ships with nvcc): set-up
 * code outside the per-step path, like the host LAPACK of the projected problem.
 */
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <time.h>
#include "b2k_internal.h"

/* idx[k] = k */
__global__ void __launch_bounds__(256) k_iota(int *__restrict__ idx, int64_t n)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) idx[k] = (int)k;
}

/* row pointer of the transpose from the sorted column keys: trp[c] = first position whose key >= c, c = 0..nt */
__global__ void __launch_bounds__(256) k_tr_rowptr(const int *__restrict__ keys, int64_t nnz, int64_t nt, int *__restrict__ trp)
{
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c > nt) return;
  int64_t lo = 0, hi = nnz;
  while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (keys[mid] < (int)c) lo = mid + 1; else hi = mid; }
  trp[c] = (int)lo;
}

/* out[c] = in[c] - base, c = 0..n */
__global__ void __launch_bounds__(256) k_tr_rebase(const int *__restrict__ in, int base, int64_t n, int *__restrict__ out)
{
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c <= n) out[c] = in[c] - base;
}

/* entries p0..p0+cnt-1 of the sorted order: column index of the transpose = the row of A that held entry perm[p] (binary search
   in A's row pointer), value = val[perm[p]] */
__global__ void __launch_bounds__(256) k_tr_gather(const int *__restrict__ perm, int64_t p0, int64_t cnt, const int *__restrict__ rowptr,
                                                    int64_t nrows, const double *__restrict__ val, int *__restrict__ tci, double *__restrict__ tv)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < cnt; q += stride) {
    const int k = perm[p0 + q];
    int64_t lo = 0, hi = nrows;                      /* largest r with rowptr[r] <= k (empty rows share a pointer: take the last) */
    while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if (rowptr[mid] <= k) lo = mid; else hi = mid; }
    tci[q] = (int)lo;
    tv[q] = val[k];
  }
}

/* one block of the transpose: transposed rows [c0, c0+nr) of the sorted order become a b2k_csr of nr rows x nrows columns */
static int tr_block(b2k_ctx ctx, const int *trp_all, const int *perm, int c0, int64_t nr, int base, int cnt, const int *rowptr, int64_t nrows,
                    const double *val, b2k_csr *out)
{
  int *trp = NULL, *tci = NULL;
  double *tv = NULL;
  CK(cudaMalloc(&trp, sizeof(int) * (size_t)(nr + 1)));
  CK(cudaMalloc(&tci, sizeof(int) * (size_t)(cnt ? cnt : 1)));
  CK(cudaMalloc(&tv, sizeof(double) * (size_t)(cnt ? cnt : 1)));
  k_tr_rebase<<<(unsigned)((nr + 1 + 255) / 256), 256, 0, ctx->stream>>>(trp_all + c0, base, nr, trp);
  CKLAUNCH(ctx);
  if (cnt) {
    k_tr_gather<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(perm, base, cnt, rowptr, nrows, val, tci, tv);
    CKLAUNCH(ctx);
  }
  return b2k_csr_adopt(ctx, nr, nrows, 0, cnt, trp, tci, tv, out);      /* takes the three arrays; builds the SELL copy */
}

/* AT_own (ncols_local x nrows) and AT_ghost (nghost x nrows; NULL when A has no ghost columns) from A, all in HBM.
   Replaces the host counting sort of round 1 (MatBuildLocalTranspose_B200CSR / MatB200CSRTranspose, matb200.c). */
extern "C" int b2k_csr_transpose_split(b2k_ctx ctx, b2k_csr A, b2k_csr *ATown, b2k_csr *ATghost)
{
  ARGCHK(ctx && A && ATown, "null argument");
  int64_t nrows = 0, ncl = 0, ng = 0, nnz = 0;
  b2k_csr_info(A, &nrows, &ncl, &ng, &nnz);
  ARGCHK(ng == 0 || ATghost, "the matrix has ghost columns: ATghost must be given");
  *ATown = NULL;
  if (ATghost) *ATghost = NULL;
  CK(cudaSetDevice(ctx->device));
  const int64_t nt = ncl + ng;
  int *rowptr = NULL, *colidx = NULL;
  double *val = NULL;
  { const int rc = b2k_csr_arrays(A, &rowptr, &colidx, &val); if (rc) return rc; }   /* the CSR copy is rebuilt from the SELL copy if it was dropped */
  int *keys = NULL, *perm_in = NULL, *perm = NULL, *trp_all = NULL;
  void *tmp = NULL;
  size_t tmp_bytes = 0;
  const size_t ne = (size_t)(nnz ? nnz : 1);
  CK(cudaMalloc(&keys, sizeof(int) * ne));
  CK(cudaMalloc(&perm_in, sizeof(int) * ne));
  CK(cudaMalloc(&perm, sizeof(int) * ne));
  CK(cudaMalloc(&trp_all, sizeof(int) * (size_t)(nt + 1)));
  int end_bit = 1;
  while (end_bit < 31 && ((int64_t)1 << end_bit) < nt) end_bit++;       /* only the bits a column index can have */
  if (nnz) {
    k_iota<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(perm_in, nnz);
    CKLAUNCH(ctx);
    CK(cub::DeviceRadixSort::SortPairs(NULL, tmp_bytes, colidx, keys, perm_in, perm, (int)nnz, 0, end_bit, ctx->stream));
    CK(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1));
    CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, colidx, keys, perm_in, perm, (int)nnz, 0, end_bit, ctx->stream));
    ctx->launches++;
  }
  k_tr_rowptr<<<(unsigned)((nt + 1 + 255) / 256), 256, 0, ctx->stream>>>(keys, nnz, nt, trp_all);
  CKLAUNCH(ctx);
  int base = 0;
  CK(cudaMemcpyAsync(&base, trp_all + ncl, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->syncs++;
  int rc = tr_block(ctx, trp_all, perm, 0, ncl, 0, base, rowptr, nrows, val, ATown);
  if (!rc && ng) rc = tr_block(ctx, trp_all, perm, (int)ncl, ng, base, (int)(nnz - base), rowptr, nrows, val, ATghost);
  CK(cudaStreamSynchronize(ctx->stream));
  cudaFree(keys); cudaFree(perm_in); cudaFree(perm); cudaFree(trp_all); cudaFree(tmp);
  b2k_csr_release_arrays(A);
  return rc;
}


/* ------------------------------------------------------------------------------------------------------------------------------
 * Local column numbering built in HBM.  A row block arrives with GLOBAL column indices (what MatMPIAIJ's users hand over); the
 * product kernels want [owned | ghosts] numbering with the ghosts = the sorted distinct off-range columns (MPIAIJ's garray,
 * MatSetUpMultiply_MPIAIJ).  Round 1 did this on the host: two passes over the entries plus a remapped copy, 0.3 s per GPU for
 * the 8.4e7 entries of C2 — the largest item of the end-to-end leg after the solve itself (bench.py phases).  Here the
 * arrays are uploaded as they are and the off-range columns are selected, sorted, made unique and looked up on the device.
 * ---------------------------------------------------------------------------------------------------------------------------- */
struct tr_offrange {
  int c0, c1;
  __host__ __device__ bool operator()(const int &c) const { return c < c0 || c >= c1; }
};

__global__ void __launch_bounds__(256) k_col_minmax(const int *__restrict__ col, int64_t nnz, int *__restrict__ mm)
{
  int lo = 2147483647, hi = -2147483647 - 1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) { const int c = col[k]; lo = min(lo, c); hi = max(hi, c); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
  if ((threadIdx.x & 31) == 0) { atomicMin(mm, lo); atomicMax(mm + 1, hi); }
}

__global__ void __launch_bounds__(256) k_col_localize(int *__restrict__ col, int64_t nnz, int c0, int c1, const int *__restrict__ ghosts, int ng)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int ncl = c1 - c0;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) {
    const int c = col[k];
    if (c >= c0 && c < c1) { col[k] = c - c0; continue; }
    int lo = 0, hi = ng - 1;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (ghosts[mid] < c) lo = mid + 1; else hi = mid; }
    col[k] = ncl + lo;
  }
}

/* rows of a row block with GLOBAL column indices in [0, ncols_global); owned columns [cstart, cend).  On return *A multiplies
   [x_owned ; x_ghost], *nghost / *ghosts_host (malloc'ed by this call, the caller frees it; NULL when there are none) list the
   global indices of the ghost columns in local order (ascending). */
extern "C" int b2k_csr_create_global(b2k_ctx ctx, int64_t nrows, int64_t ncols_global, int64_t cstart, int64_t cend, const int *rowptr_host,
                                     const int *colidx_host, const double *val_host, b2k_csr *A, int64_t *nghost, int **ghosts_host)
{
  ARGCHK(ctx && A && nghost && ghosts_host, "null argument");
  ARGCHK(nrows >= 0 && nrows < 2147483647LL && ncols_global >= 0 && ncols_global < 2147483647LL, "sizes must fit int32");
  ARGCHK(cstart >= 0 && cend >= cstart && cend <= ncols_global, "owned column range outside the matrix");
  CK(cudaSetDevice(ctx->device));
  *A = NULL; *nghost = 0; *ghosts_host = NULL;
  const char *tdbg = getenv("B2K_TIMING");                 /* B2K_TIMING=1: wall-clock of the set-up phases on stderr */
  struct timespec ts0, ts1;
  if (tdbg) clock_gettime(CLOCK_MONOTONIC, &ts0);
#define TR_TICK(what)                                                                                                              \
  do {                                                                                                                             \
    if (tdbg) {                                                                                                                    \
      cudaStreamSynchronize(ctx->stream);                                                                                          \
      clock_gettime(CLOCK_MONOTONIC, &ts1);                                                                                        \
      fprintf(stderr, "[b2k_csr_create_global] %-28s %8.2f ms\n", what, 1e3 * (ts1.tv_sec - ts0.tv_sec) + 1e-6 * (ts1.tv_nsec - ts0.tv_nsec)); \
      ts0 = ts1;                                                                                                                   \
    }                                                                                                                              \
  } while (0)
  const int64_t nnz = nrows ? rowptr_host[nrows] : 0;
  const size_t ne = (size_t)(nnz ? nnz : 1);
  int *rp = NULL, *ci = NULL, *off = NULL, *srt = NULL, *cnt = NULL;
  double *va = NULL;
  void *tmp = NULL;
  CK(cudaMalloc(&rp, sizeof(int) * (size_t)(nrows + 1)));
  CK(cudaMalloc(&ci, sizeof(int) * ne));
  CK(cudaMalloc(&va, sizeof(double) * ne));
  CK(cudaMalloc(&cnt, sizeof(int) * 4));
  CK(cudaMemcpyAsync(rp, rowptr_host, sizeof(int) * (size_t)(nrows + 1), cudaMemcpyHostToDevice, ctx->stream));
  if (nnz) {
    CK(cudaMemcpyAsync(ci, colidx_host, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(va, val_host, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, ctx->stream));
  }
  ctx->h2d_bytes += sizeof(int) * (size_t)(nrows + 1) + (sizeof(int) + sizeof(double)) * (size_t)nnz;
  TR_TICK("cudaMalloc + upload");
  int hc[4] = {2147483647, -2147483647 - 1, 0, 0};         /* min, max, off-range entries, distinct ghosts */
  CK(cudaMemcpyAsync(cnt, hc, sizeof(hc), cudaMemcpyHostToDevice, ctx->stream));
  int ng = 0;
  if (nnz) {
    k_col_minmax<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ci, nnz, cnt);
    CKLAUNCH(ctx);
    const bool all_owned = (cstart == 0 && cend == ncols_global);
    if (!all_owned) {
      size_t tb = 0, tb2 = 0, tb3 = 0;
      const tr_offrange pred = {(int)cstart, (int)cend};
      CK(cudaMalloc(&off, sizeof(int) * ne));
      CK(cub::DeviceSelect::If(NULL, tb, ci, off, cnt + 2, (int)nnz, pred, ctx->stream));
      int end_bit = 1;
      while (end_bit < 31 && ((int64_t)1 << end_bit) < ncols_global) end_bit++;
      CK(cub::DeviceRadixSort::SortKeys(NULL, tb2, off, off, (int)nnz, 0, end_bit, ctx->stream));
      CK(cub::DeviceSelect::Unique(NULL, tb3, off, off, cnt + 3, (int)nnz, ctx->stream));
      if (tb2 > tb) tb = tb2;
      if (tb3 > tb) tb = tb3;
      CK(cudaMalloc(&tmp, tb ? tb : 1));
      CK(cub::DeviceSelect::If(tmp, tb, ci, off, cnt + 2, (int)nnz, pred, ctx->stream));
      CK(cudaMemcpyAsync(hc, cnt, sizeof(hc), cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      ctx->syncs++;
      const int noff = hc[2];
      if (noff > 0) {
        CK(cudaMalloc(&srt, sizeof(int) * (size_t)noff));
        CK(cub::DeviceRadixSort::SortKeys(tmp, tb, off, srt, noff, 0, end_bit, ctx->stream));
        CK(cub::DeviceSelect::Unique(tmp, tb, srt, off, cnt + 3, noff, ctx->stream));   /* off now holds the distinct ghosts, ascending */
        CK(cudaMemcpyAsync(hc, cnt, sizeof(hc), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->syncs++;
        ng = hc[3];
      }
    } else {
      CK(cudaMemcpyAsync(hc, cnt, sizeof(hc), cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      ctx->syncs++;
    }
    if (hc[0] < 0 || (int64_t)hc[1] >= ncols_global) {
      cudaFree(rp); cudaFree(ci); cudaFree(va); cudaFree(cnt); cudaFree(off); cudaFree(srt); cudaFree(tmp);
      b2k_set_error("column index %d outside [0,%lld)", hc[0] < 0 ? hc[0] : hc[1], (long long)ncols_global);
      return B2K_ERR_ARG;
    }
    if (!all_owned) {
      k_col_localize<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ci, nnz, (int)cstart, (int)cend, off, ng);
      CKLAUNCH(ctx);
      if (ng > 0) {
        int *g = (int *)malloc(sizeof(int) * (size_t)ng);
        if (!g) return B2K_ERR_MEM;
        CK(cudaMemcpyAsync(g, off, sizeof(int) * (size_t)ng, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->d2h_bytes += sizeof(int) * (size_t)ng;
        *ghosts_host = g;
      }
    }
  }
  CK(cudaStreamSynchronize(ctx->stream));
  cudaFree(cnt); cudaFree(off); cudaFree(srt); cudaFree(tmp);
  *nghost = ng;
  TR_TICK("range check + local numbering");
  const int rc = b2k_csr_adopt(ctx, nrows, cend - cstart, ng, nnz, rp, ci, va, A);   /* takes the three arrays; builds the SELL copy */
  TR_TICK("row blocks + SELL copy");
#undef TR_TICK
  if (rc && *ghosts_host) { free(*ghosts_host); *ghosts_host = NULL; }
  return rc;
}
