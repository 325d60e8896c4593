/*
 * b2k.h — C ABI of libb200krylov.so: sm_100a CUDA kernels for SLEPc's Krylov hot path.
 *
 * This is the drop-in boundary.  Every entry point replaces one function of the
 * reference's CUDA BV backend (src/sys/classes/bv/impls/cuda/bvcuda.cu, cuBLAS wrappers) or the
 * PETSc MatMult it calls (src/sys/classes/bv/interface/bvops.c:879).  File:line citations are
 * relative to the SLEPc 3.22 tree.
 *
 * Conventions
 *   - plain C: pointers and sizes only, no torch / PETSc types;
 *   - every function returns 0 (B2K_OK) or a B2K_ERR_* code; b2k_last_error() gives the text;
 *   - pointers are DEVICE pointers unless the parameter name ends in _host;
 *   - basis blocks are column-major, leading dimension `ld` in elements, 64-bit offsets
 *     (m*ld exceeds 2^31 at 1.3e8 rows, cf. the PetscIntMultError guard at svec.c:425);
 *   - all work is queued on the context's stream; nothing synchronises unless stated;
 *   - reductions are LOCAL to this GPU (the caller all-reduces the k+1 doubles, b2k_comm_*);
 *   - results are bit-reproducible run to run (fixed-order two-stage reductions, no atomics).
 */
#ifndef B2K_H
#define B2K_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2K_OK        0
#define B2K_ERR_CUDA  1   /* a CUDA runtime call failed */
#define B2K_ERR_ARG   2   /* invalid argument */
#define B2K_ERR_NOGPU 3   /* no CUDA device: the product path has no CPU fallback */
#define B2K_ERR_COMM  4   /* NCCL / peer-memory failure */
#define B2K_ERR_MEM   5
/* widest block one reduction / projection takes (columns of V in b2k_dotvec, b2k_gs_*, b2k_sumsq …): callers validate
   nc + ncv + 1 against it up front instead of failing inside the first orthogonalisation */
#define B2K_MAX_COLUMNS 1024

typedef struct b2k_ctx_s  *b2k_ctx;    /* one per process/GPU: device, stream, scratch */
typedef struct b2k_csr_s  *b2k_csr;    /* FP64 CSR matrix resident in HBM */
typedef struct b2k_comm_s *b2k_comm;   /* row-partition communicator (NCCL over NVLink) */

const char *b2k_last_error(void);
int  b2k_version(void);
int  b2k_device_count(int *count);

/* ---- context / memory ----------------------------------------------------------------- */
int  b2k_ctx_create(int device, b2k_ctx *ctx);
int  b2k_ctx_destroy(b2k_ctx ctx);
int  b2k_ctx_sync(b2k_ctx ctx);
void *b2k_ctx_stream(b2k_ctx ctx);                       /* cudaStream_t */
int  b2k_ctx_sm_count(b2k_ctx ctx);
int  b2k_ctx_launches(b2k_ctx ctx, uint64_t *count);     /* kernels launched through this ctx */
int  b2k_ctx_syncs(b2k_ctx ctx, uint64_t *count);        /* host waits on the stream (b2k_ctx_sync + blocking copies) */
int  b2k_malloc(b2k_ctx ctx, void **dptr, size_t bytes);
int  b2k_free(b2k_ctx ctx, void *dptr);
int  b2k_memset0(b2k_ctx ctx, void *dptr, size_t bytes);
int  b2k_h2d(b2k_ctx ctx, void *dst, const void *src_host, size_t bytes);        /* blocking */
int  b2k_d2h(b2k_ctx ctx, void *dst_host, const void *src, size_t bytes);        /* blocking */
int  b2k_h2d_async(b2k_ctx ctx, void *dst, const void *src_host, size_t bytes);
int  b2k_d2h_async(b2k_ctx ctx, void *dst_host, const void *src, size_t bytes);
int  b2k_d2d(b2k_ctx ctx, void *dst, const void *src, size_t bytes);             /* async */
int  b2k_host_alloc(void **hptr, size_t bytes);                                  /* pinned */
int  b2k_host_free(void *hptr);
int  b2k_mem_info(b2k_ctx ctx, size_t *free_bytes, size_t *total_bytes);
/* CUDA-event timing on the context's stream (bench / roofline) */
int  b2k_timer_start(b2k_ctx ctx);
int  b2k_timer_stop_ms(b2k_ctx ctx, double *ms);                                 /* blocking */

/* bytes moved by b2k_h2d* / b2k_d2h* since the context was created (bench: e2e accounting) */
int  b2k_ctx_copy_bytes(b2k_ctx ctx, uint64_t *h2d_bytes, uint64_t *d2h_bytes);
/* per-kernel-class timing: CUDA events around every launch of the class on the context's stream
   (PetscLogGpuTimeBegin/End stand-in, bvcuda.cu:35-38).  b2k_prof_get synchronises. */
#define B2K_PROF_DOTVEC   0   /* k_dotvec: V^T w (+ w^T w)                                        */
#define B2K_PROF_MULTVEC  1   /* k_gs_tma<.,0,.> / k_gs_rt / k_multvec: y = beta y + alpha V q (+ norm) */
#define B2K_PROF_GSFUSED  2   /* k_gs_tma<.,1,1>: update + next pass' V^T w + norm, V read once     */
#define B2K_PROF_SPMV     3   /* k_spmv_sell_pipe / k_spmv_sell / k_spmv_csr_stream                */
#define B2K_PROF_GEMM     4   /* k_vq_tma / k_vq: V Q (restart); k_gram_tma: Y^T X; k_tsqr_*        */
#define B2K_PROF_ELEMWISE 5   /* scale / copy / axpby / fill                                      */
#define B2K_PROF_NCLASS   6
int  b2k_prof_enable(b2k_ctx ctx, int on);               /* on: start a fresh recording            */
int  b2k_prof_get(b2k_ctx ctx, int cls, uint64_t *launches, double *ms, double *algorithmic_bytes);
int  b2k_prof_get_flops(b2k_ctx ctx, int cls, double *flops);   /* level-3 classes: 2 n kin nout per launch (FP64 roofline next to the HBM one) */

/* ---- BV level-1/2/3 (replace bvcuda.cu) ------------------------------------------------------ */
/* q[0:k] = V(:,0:k)^T y                      — BVDotVec_BLAS_CUDA  bvcuda.cu:204-264 (gemv 'C')  */
int  b2k_dotvec(b2k_ctx ctx, const double *V, int64_t ld, int64_t n, int k, const double *y, double *q);
/* y = beta*y + alpha*V(:,0:k) q              — BVMultVec_BLAS_CUDA bvcuda.cu:45-60   (gemv 'N')  */
int  b2k_multvec(b2k_ctx ctx, const double *V, int64_t ld, int64_t n, int k, double alpha, double beta,
                 double *y, const double *q);
/* out[0] = sum of squares of the n x k block  — BVNorm_BLAS_CUDA    bvcuda.cu:290-303 (nrm2)      */
int  b2k_sumsq(b2k_ctx ctx, const double *X, int64_t ld, int64_t n, int k, double *out);
/* out[0] = max_j sum_i |X(i,j)| (local part of NORM_1), out[1] = max |X(i,j)| … used by BVNorm   */
int  b2k_colabssum(b2k_ctx ctx, const double *X, int64_t ld, int64_t n, int k, double *out_k);
/* out[0] = max over rows of sum_c |X(r,c)|: NORM_INFINITY of the block (BVNorm_LAPACK_Private bvlapack.c:37-83, lange 'I') */
int  b2k_rowabssum_max(b2k_ctx ctx, const double *X, int64_t ld, int64_t n, int k, double *out_dev);
/* X = alpha*X on an n x k block               — BVScale_BLAS_CUDA   bvcuda.cu:269-285 (scal)      */
int  b2k_scale(b2k_ctx ctx, double *X, int64_t ld, int64_t n, int k, double alpha);
/* Y = X on an n x k block                     — BVCopy_Svec_CUDA    sveccuda.cu:305-330           */
int  b2k_copy(b2k_ctx ctx, double *Y, int64_t ldy, const double *X, int64_t ldx, int64_t n, int k);
/* Y = alpha*X + beta*Y on an n x k block      — BVAXPY_BLAS_CUDA    bvcuda.cu:117-135 (geam)      */
int  b2k_axpby(b2k_ctx ctx, double *Y, int64_t ldy, const double *X, int64_t ldx, int64_t n, int k,
               double alpha, double beta);
/* Y(n x ky) = beta*Y + alpha*X(n x kx) Q(kx x ky) — BVMult_BLAS_CUDA bvcuda.cu:22-40 (gemm)       */
int  b2k_mult(b2k_ctx ctx, double *Y, int64_t ldy, const double *X, int64_t ldx, int64_t n, int ky, int kx,
              double alpha, double beta, const double *Q, int ldq);
/* V(:,s:e) = V(:,0:k) Q(0:k,s:e)  (Q^T if trans) in place, no workspace, no copy-back
                                               — BVMultInPlace_BLAS_CUDA bvcuda.cu:65-112          */
int  b2k_mult_inplace(b2k_ctx ctx, double *V, int64_t ld, int64_t n, int k, int s, int e,
                      const double *Q, int ldq, int trans);
/* 1 (default; env B2K_VQ_TMA): b2k_mult / b2k_mult_inplace with <= 64 columns on >= 4096 rows run the TMA-fed FP64
   tensor-core kernel (k_vq_tma); 0: the DFMA kernel k_vq                                                        */
int  b2k_vq_set_tma(int on);
/* M(ky x kx) = Y^T X (local)                  — BVDot_BLAS_CUDA     bvcuda.cu:140-199 (gemm 'C')  */
int  b2k_dot(b2k_ctx ctx, const double *Y, int64_t ldy, const double *X, int64_t ldx, int64_t n, int ky, int kx,
             double *M, int ldm);
/* x[i] = hash_uniform(row0+i, seed) in [-1,1): deterministic stand-in for BVSetRandomColumn
   (bvops.c:482, PetscRandom) shared bit-for-bit with the oracle and the host code              */
int  b2k_set_random(b2k_ctx ctx, double *x, int64_t n, int64_t row0, uint64_t seed);
int  b2k_fill(b2k_ctx ctx, double *x, int64_t n, double value);
/* w = x .* y (VecPointwiseMult: Jacobi preconditioner of the CG behind the shift-and-invert ST)            */
int  b2k_pointwise_mult(b2k_ctx ctx, double *w, const double *x, const double *y, int64_t n);

/* ---- fused classical Gram-Schmidt sweeps (replace bvorthog.c:91-132 + bvcuda.cu:345-548) ---- */
/* c[0:k] = V(:,0:k)^T w, c[k] = w^T w        — BVDotColumnInc bvorthog.c:32-47: one sweep, one
   reduction (h and ||w||^2 together)                                                            */
int  b2k_gs_dot(b2k_ctx ctx, const double *V, int64_t ld, int64_t n, int k, const double *w, double *c);
/* w -= V(:,0:k) cin ; cout[0:k] = V^T w_new ; cout[k] = ||w_new||^2 — update of pass p fused
   with the dot sweep of pass p+1 (DGKS refinement) and with the explicit norm: V is read ONCE  */
int  b2k_gs_update_dot(b2k_ctx ctx, const double *V, int64_t ld, int64_t n, int k, double *w,
                       const double *cin, double *cout);
/* w -= V(:,0:k) cin ; nrm2_out[0] = ||w_new||^2 — update sweep with the explicit norm of bvorthog.c:126
   folded in (used when no refinement pass is expected)                                              */
int  b2k_gs_update_norm(b2k_ctx ctx, const double *V, int64_t ld, int64_t n, int k, double *w,
                        const double *cin, double *nrm2_out);
/* b2k_gs_update_norm launched speculatively: every CTA evaluates the DGKS refinement criterion of bvorthog.c:180 from two device
   scalars (onrm2 = w^T w before the previous pass, nrm2 = ||w||^2 after it) and the sweep runs only if
   nrm != 0 && nrm < eta*onrm — the host learns the outcome of BOTH passes from one read (one synchronisation per column) */
int  b2k_gs_update_norm_gated(b2k_ctx ctx, const double *V, int64_t ld, int64_t n, int k, double *w, const double *cin,
                              double *nrm2_out, const double *onrm2, const double *nrm2, double eta);
/* x *= 1/sqrt(sumsq[0]) guarded (no-op if sumsq is 0 or 1): normalisation with the norm still on
   the device                                  — BVOrthonormalizeColumn bvorthog.c:417-422        */
int  b2k_scale_rsqrt(b2k_ctx ctx, double *x, int64_t n, const double *sumsq);
/* x *= 1/sqrt(s), s = nrm2_second[0] if the DGKS gate (onrm2, nrm2_first, eta) of b2k_gs_update_norm_gated let the refinement run,
   nrm2_first[0] otherwise; no-op when s is 0 or 1.  With it a whole Lanczos/Arnoldi step is enqueued without the host knowing
   the outcome of the refinement test (BVOrthonormalizeColumn bvorthog.c:417-422 after BVOrthogonalizeGS :174-203)        */
int  b2k_scale_rsqrt_gated(b2k_ctx ctx, double *x, int64_t n, const double *onrm2, const double *nrm2_first, const double *nrm2_second,
                           double eta);
/* implementation of the update sweeps (b2k_multvec, b2k_gs_update_norm, b2k_gs_update_dot); env B2K_GS_FUSED:
   0 generic kernels, two sweeps for update+dot; 1 register-tile single sweep;
   3 (default) 2-D tensor-map (TMA) pipelined single sweep, register tile for k <= 4 or fewer than 4096 rows       */
int  b2k_gs_set_fused(int mode);

/* ---- tall-skinny QR (BVOrthogonalize with BV_ORTHOG_BLOCK_TSQR / TSQRCHOL, bvorthog.c:611-656, bvlapack.c:347-560) ---- */
#define B2K_TSQR_MAX_K 64
/* the row blocking of the two kernels: nblk blocks of rows_per_blk rows (a multiple of 128; the last one shorter);
   coef_elems = doubles of reflector scalars the forward kernel writes when it stores the reflectors                     */
int  b2k_tsqr_plan(b2k_ctx ctx, int64_t n, int k, int *nblk, int64_t *rows_per_blk, int64_t *coef_elems);
/* Householder QR of every row block of V(:,0:k): Rblk_dev[b] = its k x k upper-triangular factor (column-major, ld k).
   store != 0: the reflectors overwrite V and coef_dev receives their scalars (then b2k_tsqr_backward forms Q);
   store == 0: V is left untouched (only R is wanted: TSQRCHOL, BVOrthogonalize_LAPACK_TSQR_OnlyR bvlapack.c:520-560)    */
int  b2k_tsqr_forward(b2k_ctx ctx, double *V, int64_t ld, int64_t n, int k, int store, double *Rblk_dev, double *coef_dev);
/* rows of block b <- Q_b * Wblk_dev[b] (k x k, column-major, ld k): with W = the blocks of the orthogonal factor of the stacked
   triangles this is the explicit Q of the whole basis (orgqr + the tree accumulation of bvlapack.c:386-462)            */
int  b2k_tsqr_backward(b2k_ctx ctx, double *V, int64_t ld, int64_t n, int k, const double *Wblk_dev, const double *coef_dev);

/* ---- sparse matrix-vector product (replaces PETSc MatMult behind bvops.c:879 / stsolve.c:22) -- */
/* CSR with int32 indices.  Column indices < ncols_local address x, the rest address
   xghost[col-ncols_local] (halo entries received from the neighbouring GPUs).                     */
int  b2k_csr_create(b2k_ctx ctx, int64_t nrows, int64_t ncols_local, int64_t nghost,
                    const int *rowptr_host, const int *colidx_host, const double *val_host, b2k_csr *A);
/* rows with GLOBAL column indices in [0, ncols_global), owned columns [cstart, cend): the local numbering [owned | ghosts] is built
   in HBM (off-range columns selected, sorted, made unique and looked up on the device — MPIAIJ's garray, MatSetUpMultiply_MPIAIJ);
   *ghosts_host (malloc'ed here, freed by the caller; NULL when nghost = 0) = global index of every ghost column, ascending     */
int  b2k_csr_create_global(b2k_ctx ctx, int64_t nrows, int64_t ncols_global, int64_t cstart, int64_t cend, const int *rowptr_host,
                           const int *colidx_host, const double *val_host, b2k_csr *A, int64_t *nghost, int **ghosts_host);
/* same, adopting arrays that already live in HBM (device generators)                              */
int  b2k_csr_adopt(b2k_ctx ctx, int64_t nrows, int64_t ncols_local, int64_t nghost, int64_t nnz,
                   int *rowptr, int *colidx, double *val, b2k_csr *A);
int  b2k_csr_destroy(b2k_ctx ctx, b2k_csr A);
int  b2k_csr_info(b2k_csr A, int64_t *nrows, int64_t *ncols_local, int64_t *nghost, int64_t *nnz);
/* device pointers of the CSR arrays (owned by A).  The matrix is stored ONCE: when the SELL-32 copy that the products read
   exists, the CSR (col,val) arrays are freed after the conversion (the row pointer stays); asking for colidx/val here
   rebuilds them from the SELL copy on A's stream, b2k_csr_release_arrays drops them again.  env B2K_CSR_KEEP=1 keeps both. */
int  b2k_csr_arrays(b2k_csr A, int **rowptr, int **colidx, double **val);
int  b2k_csr_release_arrays(b2k_csr A);
/* local transpose built in HBM (set-up): ATown = (ncols_local x nrows) rows of A^T that this GPU owns, ATghost = (nghost x nrows)
   rows that belong to the ghost columns (NULL when A has none; pass ATghost = NULL then).  Entries of a transposed row keep the
   ascending row order of A, so y = A^T x sums in a fixed order.  Replaces MatMultTranspose_MPIAIJ's scatter (reached from
   gklanczos.c:80,103 with the implicit transpose of svdsetup.c:273-279) and MatTranspose for the explicit one (svdsetup.c:300-306) */
int  b2k_csr_transpose_split(b2k_ctx ctx, b2k_csr A, b2k_csr *ATown, b2k_csr *ATghost);
/* diag[r] = A(r, r + diag_col_offset) of the local rows, 0 where not stored (MatGetDiagonal; offset = position of the diagonal
   block in the local column numbering, 0 for a square matrix whose owned columns are its owned rows)          */
int  b2k_csr_get_diagonal(b2k_ctx ctx, b2k_csr A, int64_t diag_col_offset, double *diag);
/* HBM bytes held by the matrix, all copies (bench: footprint)                                      */
int  b2k_csr_bytes(b2k_csr A, int64_t *bytes);
/* which kernel the most recent product of A ran (tests assert the dispatch they mean to cover)     */
#define B2K_SPMV_KERNEL_NONE            0
#define B2K_SPMV_KERNEL_CSR_STREAM      1   /* k_spmv_csr_stream                      */
#define B2K_SPMV_KERNEL_SELL            2   /* k_spmv_sell                            */
#define B2K_SPMV_KERNEL_SELL_PIPE       3   /* k_spmv_sell_pipe<false> (no ghosts)    */
#define B2K_SPMV_KERNEL_SELL_PIPE_GHOST 4   /* k_spmv_sell_pipe<true>  (halo columns) */
#define B2K_SPMV_KERNEL_SPMM            5   /* k_spmm_sell (block of vectors, small matrices)               */
#define B2K_SPMV_KERNEL_SPMM_PIPE       6   /* k_spmv_sell_pipe<.,true> (block of vectors, matrix read once) */
int  b2k_csr_last_kernel(b2k_csr A, int *which);
/* the bulk-copy pipeline kernel runs when the SELL copy has at least this many chunks (default 4 per SM, i.e. about 6e5 rows;
   env B2K_SPMV_PIPE_MIN_CHUNKS); < 0 restores the default.  Tests lower it to push small matrices through the pipeline. */
int  b2k_spmv_set_pipe_min_chunks(int min_chunks);
/* y = A [x ; xghost]                                                                               */
int  b2k_csr_spmv(b2k_ctx ctx, b2k_csr A, const double *x, const double *xghost, double *y);
/* Y(:,0:k) = A [X ; XG](:,0:k): sparse matrix times a block of k vectors (column-major, leading dimensions ldx / ldg / ldy),
   the matrix read once per 16 columns — BVMatMult in BV_MATMULT_MAT mode (svec.c:203-231, PETSc MatMatMult)            */
int  b2k_csr_spmm(b2k_ctx ctx, b2k_csr A, const double *X, int64_t ldx, const double *XG, int64_t ldg, double *Y, int64_t ldy, int k);
/* y = A x - sigma*xdiag   (shifted operator of STSHIFT, shift.c:79; xdiag = x rows owned here)    */
int  b2k_csr_spmv_shift(b2k_ctx ctx, b2k_csr A, const double *x, const double *xghost, double *y, double sigma);
/* storage used by b2k_csr_spmv: 0 CSR-stream only, 1 (default) SELL-32 copy when its padding is <= 25 %, 2 SELL-32 always;
   read when a matrix is created (env B2K_SPMV_SELL) and at every product (0 forces CSR-stream)                      */
int  b2k_spmv_set_sell(int mode);
/* device generator: rows [row0,row0+nrows) of the d-dimensional Laplacian stencil (d=1,2,3) on an
   nx*ny*nz grid, natural ordering, slab partition along the slowest index: ex1.c:37-48, ex2.c:39-54.
   Ghost layout: [lower neighbour plane | upper neighbour plane].                                   */
int  b2k_csr_laplacian(b2k_ctx ctx, int dim, int64_t nx, int64_t ny, int64_t nz, int64_t row0, int64_t nrows,
                       b2k_csr *A, int64_t *nghost_lo, int64_t *nghost_hi);
/* gather: out[i] = x[idx[i]] (packs halo send buffers)                                             */
int  b2k_gather(b2k_ctx ctx, double *out, const double *x, const int *idx, int64_t count);

/* scatter-add: out[idx[i]] += in[i] (idx unique within a call): accumulates the reverse halo of y = A^T x            */
int  b2k_scatter_add(b2k_ctx ctx, double *out, const int *idx, const double *in, int64_t count);

/* ---- row-partition communicator (replaces MPIU_Allreduce bvcuda.cu:228-248, VecScatter) -------- */
#define B2K_COMM_ID_BYTES 128
int  b2k_comm_unique_id(void *id_host /* B2K_COMM_ID_BYTES */);
int  b2k_comm_create(b2k_ctx ctx, int rank, int size, const void *id_host, b2k_comm *comm);
int  b2k_comm_destroy(b2k_comm comm);
int  b2k_comm_rank(b2k_comm comm, int *rank, int *size);
/* in-place sum over ranks of count doubles in HBM, queued on the ctx stream                        */
int  b2k_comm_allreduce_sum(b2k_comm comm, double *buf, int count);
int  b2k_comm_allreduce_max(b2k_comm comm, double *buf, int count);
/* neighbour halo exchange: send `nsend` doubles to `peer`, receive `nrecv` from it (either may be 0) */
int  b2k_comm_sendrecv(b2k_comm comm, const double *sendbuf, int64_t nsend, int send_peer,
                       double *recvbuf, int64_t nrecv, int recv_peer);
int  b2k_comm_group_start(b2k_comm comm);
int  b2k_comm_group_end(b2k_comm comm);
int  b2k_comm_allgather(b2k_comm comm, const double *sendbuf, double *recvbuf, int64_t count_per_rank);
int  b2k_comm_reduce_scatter_sum(b2k_comm comm, const double *sendbuf, double *recvbuf, int64_t count_per_rank);
int  b2k_comm_barrier(b2k_comm comm);
/* One-shot reductions over NVLink peer memory (replace the device ncclAllReduce of <= 1025 doubles that follows every
   Gram-Schmidt sweep, i.e. the MPIU_Allreduce of bvcuda.cu:228-248): each rank exports a mailbox (CUDA IPC handle,
   B2K_COMM_P2P_HANDLE_BYTES), the launcher all-gathers the handles, each rank maps its peers (<= 8 ranks, one box).
   While the reduce scope is global the BV reductions (b2k_dotvec, b2k_gs_*, b2k_sumsq, b2k_colabssum) return sums over
   all ranks directly, bit-identical on every rank; collective: all ranks issue the same reductions in the same order. */
#define B2K_COMM_P2P_HANDLE_BYTES 64
int  b2k_comm_p2p_handle(b2k_comm comm, void *handle_out_host);
int  b2k_comm_p2p_open(b2k_comm comm, const void *all_handles_host /* size x 64 bytes, rank order */);
int  b2k_comm_p2p_close(b2k_comm comm);           /* back to NCCL for the reductions (mappings are released by destroy) */
int  b2k_comm_p2p_resume(b2k_comm comm);          /* mailboxes on again after a close (collective)                      */
int  b2k_comm_p2p_enabled(b2k_comm comm);
int  b2k_comm_reduce_scope(b2k_comm comm, int global_on, int *fused_out);
int  b2k_comm_p2p_error(b2k_comm comm, int *flag_out);

/* SpMV halo over NVLink peer memory (opt-in, B2K_HALO_P2P=1 in the host layer; default: grouped ncclSend/ncclRecv):
   replaces the VecScatter of MatMult_MPIAIJ.  Collective set-up from the halo plan (who sends what to whom; handles and slot
   tables travel through ncclAllGather), then per MatMult one push kernel + one wait kernel; the ghost entries arrive in a
   double-buffered array owned by the object, *ghost_out is the buffer the SpMV of this exchange must read. */
typedef struct b2k_halo_s *b2k_halo;
int  b2k_halo_create(b2k_comm comm, int nrecv, const int *recvrank_host, const int *recvcount_host, int nsend, const int *sendrank_host,
                     const int *sendcount_host, const int *sendidx_dev /* or NULL */, const int64_t *sendoff_host /* contiguous sends */,
                     b2k_halo *halo);
int  b2k_halo_exchange(b2k_halo halo, const double *x_dev, const double **ghost_out);
int  b2k_halo_error(b2k_halo halo, int *flag_out);
int  b2k_halo_destroy(b2k_halo halo);

#ifdef __cplusplus
}
#endif
#endif
