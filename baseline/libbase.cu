/*
 * libbase.cu — LIBRARY baseline, not part of the product: the cuBLAS / cuSPARSE calls that SLEPc's own CUDA BV back-end
 * and PETSc's AIJCUSPARSE MatMult would issue for one Lanczos step of the Krylov-Schur expansion, in the reference's
 * schedule, on the same B200 and the same shapes as bench.py's workload ("whatever cuBLAS/cuSPARSE dispatch
 * on sm_100" is the bar the hand-written kernels have to beat).
 *
 * Schedule per Lanczos step at column j (classical Gram-Schmidt with one DGKS refinement, bvorthog.c:91-132,174-203):
 *     w = A v_j                               cusparseSpMV, CSR fp64/int32        (PETSc MatMult behind bvops.c:879)
 *     2 x { c = V(:,0:j+1)^T w                cublasDgemv 'T'                     (BVDotVec_BLAS_CUDA   bvcuda.cu:204-264)
 *           w = w - V(:,0:j+1) c              cublasDgemv 'N'                     (BVMultVec_BLAS_CUDA  bvcuda.cu:45-60)  }
 *     ||w||                                   cublasDnrm2                         (BVNorm_BLAS_CUDA     bvcuda.cu:290-303)
 *     w = w / ||w||                           cublasDscal                         (BVScale_BLAS_CUDA    bvcuda.cu:269-285)
 * and once per restart  V(:,0:nout) = V(:,0:k) Q   cublasDgemm into a workspace + cudaMemcpy2D back (bvcuda.cu:65-112).
 * Device pointer mode, no host synchronisation inside the timed region: the reference's 5 coefficient micro-kernels and 3
 * blocking 8-byte copies per pass (bvcuda.cu:345-548) are NOT counted, so this is a lower bound of the reference's GPU time.
 * Built by `make` into baseline/libbase.so (links the CUDA toolkit's libcublas / libcusparse); called by bench.py only.
 */
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cusparse.h>
#include <stdint.h>
#include <stdio.h>

#define CKC(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { snprintf(g_err, sizeof g_err, "%s:%d %s -> %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return 1; } } while (0)
#define CKB(x) do { cublasStatus_t s_ = (x); if (s_ != CUBLAS_STATUS_SUCCESS) { snprintf(g_err, sizeof g_err, "%s:%d %s -> cublas status %d", __FILE__, __LINE__, #x, (int)s_); return 2; } } while (0)
#define CKS(x) do { cusparseStatus_t s_ = (x); if (s_ != CUSPARSE_STATUS_SUCCESS) { snprintf(g_err, sizeof g_err, "%s:%d %s -> cusparse status %d", __FILE__, __LINE__, #x, (int)s_); return 3; } } while (0)
static char g_err[512];
extern "C" const char *libbase_last_error(void) { return g_err; }

__global__ void lb_lap_count(int64_t gx, int64_t gy, int64_t gz, int dim, int *cnt)
{
  const int64_t n = gx * gy * gz;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = r / (gy * gz), rem = r % (gy * gz), j = rem / gz, k = rem % gz;
    int c = 1 + (i > 0) + (i < gx - 1);
    if (dim >= 2) c += (j > 0) + (j < gy - 1);
    if (dim >= 3) c += (k > 0) + (k < gz - 1);
    cnt[r] = c;
  }
}
__global__ void lb_lap_fill(int64_t gx, int64_t gy, int64_t gz, int dim, const int *rowptr, int *col, double *val)
{
  const int64_t n = gx * gy * gz, plane = gy * gz;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = r / plane, rem = r % plane, j = rem / gz, k = rem % gz;
    int p = rowptr[r];
    if (i > 0) { col[p] = (int)(r - plane); val[p++] = -1.0; }
    if (dim >= 2 && j > 0) { col[p] = (int)(r - gz); val[p++] = -1.0; }
    if (dim >= 3 && k > 0) { col[p] = (int)(r - 1); val[p++] = -1.0; }
    col[p] = (int)r; val[p++] = 2.0 * dim;
    if (dim >= 3 && k < gz - 1) { col[p] = (int)(r + 1); val[p++] = -1.0; }
    if (dim >= 2 && j < gy - 1) { col[p] = (int)(r + gz); val[p++] = -1.0; }
    if (i < gx - 1) { col[p] = (int)(r + plane); val[p++] = -1.0; }
  }
}
__global__ void lb_fill(double *x, int64_t n, uint64_t seed)
{
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    uint64_t z = (uint64_t)(r + 1) * 0x9E3779B97F4A7C15ull + seed;
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 27; z *= 0x94D049BB133111EBull; z ^= z >> 31;
    x[r] = 2.0 * ((double)(z >> 11) * (1.0 / 9007199254740992.0)) - 1.0;
  }
}
struct ev_timer {
  cudaEvent_t a, b;
  int init() { return cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess; }
  void start(cudaStream_t s) { cudaEventRecord(a, s); }
  double stop(cudaStream_t s) { float f = 0; cudaEventRecord(b, s); cudaEventSynchronize(b); cudaEventElapsedTime(&f, a, b); return (double)f; }
};

/*
 * dim-dimensional Laplacian on gx*gy*gz, basis of kmax+1 columns; runs `cycles` restart cycles of the library schedule with
 * j = kmin .. kmax-1 and a kmax -> nout restart GEMM.  out[]:
 *   0 ms per restart cycle   1 ms per Lanczos step   2 Lanczos steps per cycle
 *   3 ms gemv'T' (k = kprobe)   4 ms gemv'N'   5 ms SpMV   6 ms nrm2   7 ms scal   8 ms gemm + copy-back   9 nnz   10 rows
 */
extern "C" int libbase_run(int dim, int64_t gx, int64_t gy, int64_t gz, int kmin, int kmax, int nout, int cycles, int warmup, int kprobe, double *out)
{
  if (dim < 3) gz = 1;
  if (dim < 2) gy = 1;
  const int64_t n = gx * gy * gz;
  cudaStream_t st;
  CKC(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  /* matrix: counts -> row pointer (host scan; set-up) -> fill */
  int *cnt = NULL, *rowptr = NULL, *col = NULL;
  double *val = NULL;
  CKC(cudaMalloc(&cnt, sizeof(int) * (size_t)n));
  CKC(cudaMalloc(&rowptr, sizeof(int) * (size_t)(n + 1)));
  lb_lap_count<<<1184, 256, 0, st>>>(gx, gy, gz, dim, cnt);
  int *hcnt = (int *)malloc(sizeof(int) * (size_t)(n + 1));
  if (!hcnt) { snprintf(g_err, sizeof g_err, "out of host memory"); return 4; }
  CKC(cudaMemcpyAsync(hcnt + 1, cnt, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));
  CKC(cudaStreamSynchronize(st));
  hcnt[0] = 0;
  long long run = 0;
  for (int64_t r = 1; r <= n; r++) { run += hcnt[r]; hcnt[r] = (int)run; }
  const int64_t nnz = run;
  if (nnz >= 2147483647LL) { snprintf(g_err, sizeof g_err, "nnz exceeds int32"); return 4; }
  CKC(cudaMemcpyAsync(rowptr, hcnt, sizeof(int) * (size_t)(n + 1), cudaMemcpyHostToDevice, st));
  CKC(cudaMalloc(&col, sizeof(int) * (size_t)nnz));
  CKC(cudaMalloc(&val, sizeof(double) * (size_t)nnz));
  lb_lap_fill<<<1184, 256, 0, st>>>(gx, gy, gz, dim, rowptr, col, val);
  CKC(cudaStreamSynchronize(st));
  free(hcnt);
  cudaFree(cnt);

  const int64_t ld = n;
  double *V = NULL, *W = NULL, *coef = NULL, *Q = NULL, *scal = NULL;
  CKC(cudaMalloc(&V, sizeof(double) * (size_t)ld * (size_t)(kmax + 1)));
  CKC(cudaMalloc(&W, sizeof(double) * (size_t)ld * (size_t)nout));      /* the reference's lda x (e-s) workspace */
  CKC(cudaMalloc(&coef, sizeof(double) * (size_t)(kmax + 8)));
  CKC(cudaMalloc(&Q, sizeof(double) * (size_t)kmax * (size_t)kmax));
  CKC(cudaMalloc(&scal, sizeof(double) * 8));
  lb_fill<<<1184, 256, 0, st>>>(V, ld * (int64_t)(kmax + 1), 1);
  lb_fill<<<8, 256, 0, st>>>(Q, (int64_t)kmax * kmax, 2);
  const double hs[4] = {1.0, 0.0, -1.0, 1e-3};
  CKC(cudaMemcpyAsync(scal, hs, sizeof(hs), cudaMemcpyHostToDevice, st));
  const double *one = scal, *zero = scal + 1, *mone = scal + 2, *small = scal + 3;

  cublasHandle_t bl;
  cusparseHandle_t sp;
  CKB(cublasCreate(&bl));
  CKB(cublasSetStream(bl, st));
  CKB(cublasSetPointerMode(bl, CUBLAS_POINTER_MODE_DEVICE));
  CKS(cusparseCreate(&sp));
  CKS(cusparseSetStream(sp, st));
  CKS(cusparseSetPointerMode(sp, CUSPARSE_POINTER_MODE_DEVICE));
  cusparseSpMatDescr_t mA;
  CKS(cusparseCreateCsr(&mA, n, n, nnz, rowptr, col, val, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, CUDA_R_64F));
  cusparseDnVecDescr_t vx[2];        /* descriptors are re-pointed per column */
  CKS(cusparseCreateDnVec(&vx[0], n, V, CUDA_R_64F));
  CKS(cusparseCreateDnVec(&vx[1], n, V + ld, CUDA_R_64F));
  size_t bufsz = 0;
  CKS(cusparseSpMV_bufferSize(sp, CUSPARSE_OPERATION_NON_TRANSPOSE, one, mA, vx[0], zero, vx[1], CUDA_R_64F, CUSPARSE_SPMV_ALG_DEFAULT, &bufsz));
  void *spbuf = NULL;
  CKC(cudaMalloc(&spbuf, bufsz ? bufsz : 16));
  CKS(cusparseSpMV_preprocess(sp, CUSPARSE_OPERATION_NON_TRANSPOSE, one, mA, vx[0], zero, vx[1], CUDA_R_64F, CUSPARSE_SPMV_ALG_DEFAULT, spbuf));

  ev_timer T;
  if (T.init()) { snprintf(g_err, sizeof g_err, "event creation failed"); return 1; }

  auto step = [&](int j) -> int {                 /* columns 0..j are the basis, column j+1 receives w */
    double *vj = V + (size_t)j * ld, *w = V + (size_t)(j + 1) * ld;
    CKS(cusparseDnVecSetValues(vx[0], vj));
    CKS(cusparseDnVecSetValues(vx[1], w));
    CKS(cusparseSpMV(sp, CUSPARSE_OPERATION_NON_TRANSPOSE, one, mA, vx[0], zero, vx[1], CUDA_R_64F, CUSPARSE_SPMV_ALG_DEFAULT, spbuf));
    for (int pass = 0; pass < 2; pass++) {
      CKB(cublasDgemv(bl, CUBLAS_OP_T, (int)n, j + 1, small, V, (int)ld, w, 1, zero, coef, 1));      /* `small`: keeps w bounded over many cycles */
      CKB(cublasDgemv(bl, CUBLAS_OP_N, (int)n, j + 1, mone, V, (int)ld, coef, 1, one, w, 1));
    }
    CKB(cublasDnrm2(bl, (int)n, w, 1, coef + kmax + 1));
    CKB(cublasDscal(bl, (int)n, one, w, 1));       /* alpha = 1 on the device: same traffic as 1/nrm, keeps the data bounded */
    return 0;
  };
  auto restart = [&]() -> int {
    CKB(cublasDgemm(bl, CUBLAS_OP_N, CUBLAS_OP_N, (int)n, nout, kmax, small, V, (int)ld, Q, kmax, zero, W, (int)ld));
    CKC(cudaMemcpy2DAsync(V, sizeof(double) * (size_t)ld, W, sizeof(double) * (size_t)ld, sizeof(double) * (size_t)n, (size_t)nout, cudaMemcpyDeviceToDevice, st));
    return 0;
  };
  for (int c = 0; c < warmup; c++) { for (int j = kmin; j < kmax; j++) if (step(j)) return 5; if (restart()) return 5; }
  CKC(cudaStreamSynchronize(st));
  T.start(st);
  for (int c = 0; c < cycles; c++) { for (int j = kmin; j < kmax; j++) if (step(j)) return 5; if (restart()) return 5; }
  const double ms_all = T.stop(st);
  out[0] = ms_all / cycles;
  out[2] = (double)(kmax - kmin);
  out[1] = out[0] / out[2];

  /* per-operation times at k = kprobe columns */
  const int reps = 10, k = kprobe;
  double *w = V + (size_t)kmax * ld;
  T.start(st); for (int r = 0; r < reps; r++) CKB(cublasDgemv(bl, CUBLAS_OP_T, (int)n, k, small, V, (int)ld, w, 1, zero, coef, 1)); out[3] = T.stop(st) / reps;
  T.start(st); for (int r = 0; r < reps; r++) CKB(cublasDgemv(bl, CUBLAS_OP_N, (int)n, k, mone, V, (int)ld, coef, 1, one, w, 1)); out[4] = T.stop(st) / reps;
  CKS(cusparseDnVecSetValues(vx[0], V)); CKS(cusparseDnVecSetValues(vx[1], w));
  T.start(st); for (int r = 0; r < reps; r++) CKS(cusparseSpMV(sp, CUSPARSE_OPERATION_NON_TRANSPOSE, one, mA, vx[0], zero, vx[1], CUDA_R_64F, CUSPARSE_SPMV_ALG_DEFAULT, spbuf)); out[5] = T.stop(st) / reps;
  T.start(st); for (int r = 0; r < reps; r++) CKB(cublasDnrm2(bl, (int)n, w, 1, coef)); out[6] = T.stop(st) / reps;
  T.start(st); for (int r = 0; r < reps; r++) CKB(cublasDscal(bl, (int)n, one, w, 1)); out[7] = T.stop(st) / reps;
  T.start(st); for (int r = 0; r < reps; r++) if (restart()) return 5; out[8] = T.stop(st) / reps;
  out[9] = (double)nnz; out[10] = (double)n;

  cusparseDestroyDnVec(vx[0]); cusparseDestroyDnVec(vx[1]); cusparseDestroySpMat(mA);
  cusparseDestroy(sp); cublasDestroy(bl);
  cudaFree(spbuf); cudaFree(V); cudaFree(W); cudaFree(coef); cudaFree(Q); cudaFree(scal); cudaFree(rowptr); cudaFree(col); cudaFree(val);
  cudaEventDestroy(T.a); cudaEventDestroy(T.b);
  cudaStreamDestroy(st);
  return 0;
}
