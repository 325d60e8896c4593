/*
 * b2kslepc.h — host-side mirror of the SLEPc interface for the Krylov hot path
 * (libb2kslepc.so, plain C).  Names, argument meaning and error behaviour follow SLEPc 3.22
 * (include/slepcbv.h, slepcds.h, slepcst.h, slepceps.h, slepcsvd.h) so that callers and tests read
 * like the reference's own; the arithmetic is delegated through the BV ops table
 * (include/slepc/private/bvimpl.h:25-61) to a registered BV type:
 *     "b200"  — device-resident basis, sm_100a kernels through include/b2k.h   (the product)
 * and through the Mat ops (mult / multtranspose) to the Mat type "b200csr".
 * PETSc is not available in this image, so the few PETSc objects the path needs (Vec, dense
 * sequential Mat, error codes) are provided here in minimal form.  With a real PETSc/SLEPc the
 * same ops plug into BVRegister()/MatShell — see INTEGRATION.md.
 *
 * Real scalars, 32-bit PetscInt (SLEPc's default build), column-major dense storage.
 */
#ifndef B2KSLEPC_H
#define B2KSLEPC_H
#include <stddef.h>
#include <stdint.h>
#include "b2k.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef int    PetscErrorCode;
typedef int    PetscInt;
typedef int    PetscBool;
typedef double PetscScalar;
typedef double PetscReal;
#define PETSC_SUCCESS 0
#define PETSC_TRUE    1
#define PETSC_FALSE   0
#define PETSC_DECIDE    (-1)
#define PETSC_DETERMINE (-1)
#define PETSC_DEFAULT   (-2)
#define PETSC_CURRENT   (-3)
#define PETSC_MACHINE_EPSILON 2.2204460492503131e-16
#define PETSC_MAX_REAL        1.7976931348623157e308
#define SLEPC_DEFAULT_TOL     1e-8                      /* include/slepcmath.h:25 */

/* PETSc error codes used on this path (petscerror.h numbering) */
#define PETSC_ERR_MEM             55
#define PETSC_ERR_SUP             56
#define PETSC_ERR_ORDER           58
#define PETSC_ERR_ARG_SIZ         60
#define PETSC_ERR_ARG_WRONG       62
#define PETSC_ERR_ARG_OUTOFRANGE  63
#define PETSC_ERR_FP              72
#define PETSC_ERR_ARG_WRONGSTATE  73
#define PETSC_ERR_ARG_INCOMP      75
#define PETSC_ERR_LIB             76
#define PETSC_ERR_PLIB            77
#define PETSC_ERR_CONV_FAILED     82
#define PETSC_ERR_ARG_NULL        85
#define PETSC_ERR_ARG_UNKNOWN_TYPE 86
#define PETSC_ERR_USER_INPUT      110
#define PETSC_ERR_GPU             97

/* text of the last error raised by this library on the calling thread (with a call trace) */
const char *B2KGetLastErrorMessage(void);
void        B2KClearError(void);

typedef enum { B2K_MEM_HOST = 0, B2K_MEM_DEVICE = 1 } B2KMemType;
typedef enum { NORM_1 = 0, NORM_2 = 1, NORM_FROBENIUS = 2, NORM_INFINITY = 3 } NormType;

typedef void *PetscObject;                          /* a Vec, Mat or BV handle cast as in the reference: (PetscObject)X */
typedef struct _p_PetscViewer *PetscViewer;         /* ASCII on stdout; NULL means the same */
typedef enum { PETSC_VIEWER_DEFAULT = 0, PETSC_VIEWER_ASCII_INFO = 1, PETSC_VIEWER_ASCII_INFO_DETAIL = 2 } PetscViewerFormat;
typedef struct _p_B2KComm *B2KComm;
typedef struct _p_Vec *Vec;
typedef struct _p_Mat *Mat;
typedef struct _p_BV  *BV;
typedef struct _p_DS  *DS;
typedef struct _p_ST  *ST;
typedef struct _p_EPS *EPS;
typedef struct _p_SVD *SVD;

/* names and viewers: only what the reference's BV programs print (bv/tests/test1.c -verbose) */
PetscErrorCode PetscObjectSetName(PetscObject obj, const char name[]);
PetscErrorCode PetscViewerASCIIGetStdout(PetscViewer *viewer);
PetscErrorCode PetscViewerPushFormat(PetscViewer viewer, PetscViewerFormat format);
PetscErrorCode PetscViewerPopFormat(PetscViewer viewer);
PetscErrorCode PetscViewerGetFormat(PetscViewer viewer, PetscViewerFormat *format);
void           B2KFormatReal(double x, char buf[32]);     /* "%g" as PETSc's ASCII output shows reals: -2 prints as "-2." */

/* ---- device context and communicator (stand-ins for PETSc's CUDA handle and MPI_Comm) -------- */
PetscErrorCode B2KInitialize(int device);           /* creates the process-wide b2k_ctx; errors if no GPU */
PetscErrorCode B2KFinalize(void);
b2k_ctx        B2KGetContext(void);                 /* NULL before B2KInitialize */
PetscErrorCode B2KDeviceSynchronize(void);
/* world communicator: size 1 by default; one process per GPU otherwise */
B2KComm        B2KCommWorld(void);
PetscErrorCode B2KCommInitNCCL(int rank, int size, const void *nccl_unique_id /* B2K_COMM_ID_BYTES */);
typedef int (*B2KAllreduceFn)(double *buf, int count, int op /*0 sum,1 max*/, void *user);
typedef int (*B2KSendrecvFn)(const double *sbuf, int nsend, int dest, double *rbuf, int nrecv, int src, void *user);
/* host-buffer communicator driven by callbacks (used with CPU BV types, e.g. gloo in the tests) */
PetscErrorCode B2KCommInitCallbacks(int rank, int size, B2KAllreduceFn ar, B2KSendrecvFn sr, void *user);
PetscErrorCode B2KCommReset(void);
/* optional, after B2KCommInitNCCL on the GPUs of one box: peer-memory (NVLink) mailboxes for the k-vector reductions of
   every Gram-Schmidt sweep.  Handle: 64 bytes per rank (include/b2k.h b2k_comm_p2p_*), all-gathered by the launcher. */
PetscErrorCode B2KCommP2PHandle(void *handle_out /* 64 bytes */);
PetscErrorCode B2KCommP2POpen(const void *all_handles /* size x 64 bytes in rank order */);
PetscErrorCode B2KCommDisableP2P(void);           /* reductions through ncclAllReduce again (collective)                   */
PetscErrorCode B2KCommEnableP2P(void);            /* back to the mapped mailboxes after a B2KCommDisableP2P (collective)   */
/* reductions of the BV kernels issued between Begin(global) and End are sums over the ranks when *fused comes back true */
PetscErrorCode B2KCommReduceScope(B2KComm comm, PetscBool global, PetscBool *fused);
PetscErrorCode B2KCommGetRank(B2KComm comm, int *rank, int *size);
PetscErrorCode B2KCommAllreduce(B2KComm comm, double *buf, int count, int op, B2KMemType where);
PetscErrorCode B2KCommSendrecv(B2KComm comm, const double *sbuf, PetscInt nsend, int dest, double *rbuf, PetscInt nrecv, int src,
                               B2KMemType where);

/* ---- Vec (minimal) --------------------------------------------------------------------------- */
PetscErrorCode VecCreateWithArray(B2KMemType mem, PetscInt n, PetscInt N, PetscScalar *array, Vec *v); /* array may be NULL */
PetscErrorCode VecCreateB200(PetscInt n, PetscInt N, Vec *v);        /* device vector, owns storage */
PetscErrorCode VecCreateHost(PetscInt n, PetscInt N, Vec *v);        /* host vector, owns storage   */
PetscErrorCode VecDestroy(Vec *v);
PetscErrorCode VecGetSize(Vec v, PetscInt *N);
PetscErrorCode VecGetLocalSize(Vec v, PetscInt *n);
PetscErrorCode VecGetArray(Vec v, PetscScalar **a);                  /* raw pointer in the vector's memory space */
PetscErrorCode VecGetMemType(Vec v, B2KMemType *mem);
PetscErrorCode VecPlaceArray(Vec v, PetscScalar *a);
PetscErrorCode VecResetArray(Vec v);
PetscErrorCode VecSetValuesHost(Vec v, const PetscScalar *host_values);   /* upload n values   */
PetscErrorCode VecGetValuesHost(Vec v, PetscScalar *host_values);         /* download n values */
PetscErrorCode VecSet(Vec v, PetscScalar alpha);
PetscErrorCode VecView(Vec v, PetscViewer viewer);
PetscErrorCode B2KVecSetOwnershipStart(Vec v, PetscInt rstart);          /* first global row of a row-partitioned template vector */
PetscErrorCode VecNorm(Vec v, NormType type, PetscReal *val);             /* collective */
PetscErrorCode VecDot(Vec x, Vec y, PetscScalar *val);                    /* collective */
PetscErrorCode VecAXPY(Vec y, PetscScalar alpha, Vec x);
PetscErrorCode VecScale(Vec v, PetscScalar alpha);
PetscErrorCode VecCopy(Vec x, Vec y);
PetscErrorCode VecAYPX(Vec y, PetscScalar beta, Vec x);                   /* y = x + beta y */
PetscErrorCode VecPointwiseMult(Vec w, Vec x, Vec y);                     /* w = x .* y     */
PetscErrorCode VecReciprocal(Vec v);                                      /* v = 1 ./ v (zeros stay zero), set-up only: goes through the host */

/* ---- Mat (dense sequential host matrices for Q/H/T/M arguments; operator matrices) ----------- */
PetscErrorCode MatCreateSeqDense(PetscInt m, PetscInt n, PetscScalar *data /* or NULL */, Mat *A);
PetscErrorCode MatDenseSetLDA(Mat A, PetscInt lda);
PetscErrorCode MatDenseGetLDA(Mat A, PetscInt *lda);
PetscErrorCode MatDenseGetArray(Mat A, PetscScalar **a);
PetscErrorCode MatDenseRestoreArray(Mat A, PetscScalar **a);
PetscErrorCode MatView(Mat A, PetscViewer viewer);                       /* sequential dense matrices */
PetscErrorCode MatGetSize(Mat A, PetscInt *M, PetscInt *N);
PetscErrorCode MatGetLocalSize(Mat A, PetscInt *m, PetscInt *n);
PetscErrorCode MatGetOwnershipRange(Mat A, PetscInt *rstart, PetscInt *rend);
PetscErrorCode MatDestroy(Mat *A);
PetscErrorCode MatMult(Mat A, Vec x, Vec y);
PetscErrorCode MatMultTranspose(Mat A, Vec x, Vec y);
PetscErrorCode MatCreateVecs(Mat A, Vec *right, Vec *left);
PetscErrorCode MatCreateHermitianTranspose(Mat A, Mat *At);          /* virtual: MatMult(At) = MatMultTranspose(A) */
PetscErrorCode MatGetType(Mat A, const char **type);
PetscErrorCode MatGetDiagonal(Mat A, Vec d);                         /* Jacobi preconditioner of the ST's linear solves */
/* operator plug-in, the MatShell route (cf. src/eps/tutorials/ex3.c:46-49,140-168) */
typedef PetscErrorCode (*MatMultFn)(Mat A, Vec x, Vec y);
PetscErrorCode MatCreateShell(PetscInt m, PetscInt n, PetscInt M, PetscInt N, B2KMemType mem, void *ctx, Mat *A);
PetscErrorCode MatShellGetContext(Mat A, void **ctx);
PetscErrorCode MatShellSetMult(Mat A, MatMultFn mult);
PetscErrorCode MatShellSetMultTranspose(Mat A, MatMultFn multtranspose);
typedef PetscErrorCode (*MatGetDiagonalFn)(Mat A, Vec d);
PetscErrorCode MatShellSetGetDiagonal(Mat A, MatGetDiagonalFn getdiagonal);
/* Mat type "b200csr": CSR rows [rstart,rend) of a global M x N matrix resident in HBM.
   colidx holds GLOBAL column indices.  The column space is partitioned by `colstarts`
   (size+1 entries; NULL = same as the row partition for square matrices / single rank).
   On >1 rank the halo plan must be given with MatB200CSRSetHalo before the first MatMult. */
PetscErrorCode MatCreateB200CSR(PetscInt M, PetscInt N, PetscInt rstart, PetscInt rend, const PetscInt *rowptr,
                                const PetscInt *colidx_global, const PetscScalar *val, PetscInt cstart, PetscInt cend, Mat *A);
/* ghosts: sorted global column indices outside [cstart,cend) used by the local rows (filled by the constructor) */
PetscErrorCode MatB200CSRGetGhosts(Mat A, PetscInt *nghost, const PetscInt **ghosts_global);
/* per neighbour p: receive recvcount[p] consecutive ghosts (in ghost order) from recvrank[p];
   send sendcount[q] entries x[sendidx[...]] (LOCAL indices) to sendrank[q]                        */
PetscErrorCode MatB200CSRSetHalo(Mat A, PetscInt nrecv, const PetscInt *recvrank, const PetscInt *recvcount,
                                 PetscInt nsend, const PetscInt *sendrank, const PetscInt *sendcount, const PetscInt *sendidx);
/* d-dimensional Laplacian generated on the device, slab-partitioned over the world communicator
   (ex1.c:37-48, ex2.c:39-54; 7-point analogue).  nx is the slowest index and is what gets split. */
PetscErrorCode MatCreateB200Laplacian(PetscInt dim, PetscInt nx, PetscInt ny, PetscInt nz, Mat *A);
/* explicit transpose of a b200csr matrix (single rank): what SVDSetUp builds by default, svdsetup.c:300-306 */
PetscErrorCode MatB200CSRTranspose(Mat A, Mat *At);
PetscErrorCode MatB200CSRGetInfo(Mat A, int64_t *nnz, int64_t *nghost);

/* ---- BV (include/slepcbv.h) ------------------------------------------------------------------- */
typedef enum { BV_ORTHOG_CGS = 0, BV_ORTHOG_MGS = 1 } BVOrthogType;
typedef enum { BV_ORTHOG_REFINE_IFNEEDED = 0, BV_ORTHOG_REFINE_NEVER = 1, BV_ORTHOG_REFINE_ALWAYS = 2 } BVOrthogRefineType;
typedef enum { BV_ORTHOG_BLOCK_GS = 0, BV_ORTHOG_BLOCK_CHOL = 1, BV_ORTHOG_BLOCK_TSQR = 2, BV_ORTHOG_BLOCK_TSQRCHOL = 3,
               BV_ORTHOG_BLOCK_SVQB = 4 } BVOrthogBlockType;   /* include/slepcbv.h:84-88; TSQR/TSQRCHOL: PETSC_ERR_SUP here */
#define BVB200 "b200"

PetscErrorCode BVRegister(const char *name, PetscErrorCode (*ctor)(BV));      /* slepcbv.h:259, bvfunc.c:646 */
PetscErrorCode BVCreate(BV *bv);
PetscErrorCode BVDestroy(BV *bv);
PetscErrorCode BVSetType(BV bv, const char *type);                            /* bvbasic.c:35 */
PetscErrorCode BVGetType(BV bv, const char **type);
PetscErrorCode BVSetSizes(BV bv, PetscInt n, PetscInt N, PetscInt m);         /* bvbasic.c:105 */
PetscErrorCode BVSetSizesFromVec(BV bv, Vec t, PetscInt m);                   /* bvbasic.c:166 */
PetscErrorCode BVGetSizes(BV bv, PetscInt *n, PetscInt *N, PetscInt *m);
PetscErrorCode BVGetLeadingDimension(BV bv, PetscInt *ld);
PetscErrorCode BVSetActiveColumns(BV bv, PetscInt l, PetscInt k);             /* bvbasic.c:350 */
PetscErrorCode BVGetActiveColumns(BV bv, PetscInt *l, PetscInt *k);
PetscErrorCode BVSetOrthogonalization(BV bv, BVOrthogType type, BVOrthogRefineType refine, PetscReal eta, BVOrthogBlockType block);
PetscErrorCode BVGetOrthogonalization(BV bv, BVOrthogType *type, BVOrthogRefineType *refine, PetscReal *eta, BVOrthogBlockType *block);
PetscErrorCode BVResize(BV bv, PetscInt m, PetscBool copy);                   /* bvbasic.c:253 */
PetscErrorCode BVDuplicate(BV V, BV *W);
PetscErrorCode BVDuplicateResize(BV V, PetscInt m, BV *W);
PetscErrorCode BVGetColumn(BV bv, PetscInt j, Vec *v);                        /* bvbasic.c:1116: at most two at a time */
PetscErrorCode BVRestoreColumn(BV bv, PetscInt j, Vec *v);
PetscErrorCode BVGetArray(BV bv, PetscScalar **a);                            /* pointer in the BV's memory space */
PetscErrorCode BVRestoreArray(BV bv, PetscScalar **a);
PetscErrorCode BVGetMemType(BV bv, B2KMemType *mem);
PetscErrorCode BVCreateVec(BV bv, Vec *v);
PetscErrorCode BVSetColumnHost(BV bv, PetscInt j, const PetscScalar *host_values);   /* convenience: upload column j   */
PetscErrorCode BVGetColumnHost(BV bv, PetscInt j, PetscScalar *host_values);         /* convenience: download column j */
PetscErrorCode BVInsertVec(BV bv, PetscInt j, Vec w);                         /* bvbasic.c:1566 */
PetscErrorCode BVCopyVec(BV bv, PetscInt j, Vec w);
PetscErrorCode BVCopy(BV V, BV W);                                            /* bvbasic.c:1797 */
PetscErrorCode BVCopyColumn(BV V, PetscInt j, PetscInt i);                    /* bvbasic.c:1871 */
PetscErrorCode BVSetRandomColumn(BV bv, PetscInt j);                          /* bvops.c:482 (deterministic hash here) */
PetscErrorCode BVSetRandomSeed(BV bv, uint64_t seed);
PetscErrorCode BVMult(BV Y, PetscScalar alpha, PetscScalar beta, BV X, Mat Q);          /* bvops.c:22  */
PetscErrorCode BVMultVec(BV X, PetscScalar alpha, PetscScalar beta, Vec y, PetscScalar q[]);   /* bvops.c:110 */
PetscErrorCode BVMultColumn(BV X, PetscScalar alpha, PetscScalar beta, PetscInt j, PetscScalar *q); /* bvops.c:165 */
PetscErrorCode BVMultInPlace(BV V, Mat Q, PetscInt s, PetscInt e);                      /* bvops.c:220 */
PetscErrorCode BVMultInPlaceHermitianTranspose(BV V, Mat Q, PetscInt s, PetscInt e);    /* bvops.c:268 */
PetscErrorCode BVScale(BV bv, PetscScalar alpha);                                       /* bvops.c:379 */
PetscErrorCode BVScaleColumn(BV bv, PetscInt j, PetscScalar alpha);                     /* bvops.c:419 */
PetscErrorCode BVDot(BV X, BV Y, Mat M);                                                /* bvglobal.c:52  */
PetscErrorCode BVDotVec(BV X, Vec y, PetscScalar m[]);                                  /* bvglobal.c:151 */
PetscErrorCode BVDotColumn(BV X, PetscInt j, PetscScalar *q);                           /* bvglobal.c:299 */
PetscErrorCode BVNorm(BV bv, NormType type, PetscReal *val);                            /* bvglobal.c:455 */
PetscErrorCode BVNormVec(BV bv, Vec v, NormType type, PetscReal *val);                  /* bvglobal.c:590 */
PetscErrorCode BVNormColumn(BV bv, PetscInt j, NormType type, PetscReal *val);          /* bvglobal.c:523 */
PetscErrorCode BVNormalize(BV bv, PetscScalar *eigi);                                   /* bvglobal.c:836 */
/* non-standard inner product <x,y> = y^T B x (B symmetric positive definite): BVDot*, BVNorm*, the orthogonalisations and
   BVNormalize use it (bvbasic.c:497, BV_IPMatMult bvimpl.h:147-157) */
PetscErrorCode BVSetMatrix(BV bv, Mat B, PetscBool indef);                              /* bvbasic.c:497 */
PetscErrorCode BVGetMatrix(BV bv, Mat *B, PetscBool *indef);                            /* bvbasic.c:565 */
/* split-phase reductions: every Begin queues its local part, the first End performs ONE global reduction for all of them
   (PetscSplitReduction; a BV type may take the pair over with the dotvec_begin/end, norm_begin/end slots, bvimpl.h:33-39) */
PetscErrorCode BVDotVecBegin(BV X, Vec y, PetscScalar *m);                              /* bvglobal.c:188 */
PetscErrorCode BVDotVecEnd(BV X, Vec y, PetscScalar *m);                                /* bvglobal.c:238 */
PetscErrorCode BVDotColumnBegin(BV X, PetscInt j, PetscScalar *m);                      /* bvglobal.c:350 */
PetscErrorCode BVDotColumnEnd(BV X, PetscInt j, PetscScalar *m);                        /* bvglobal.c:410 */
PetscErrorCode BVNormColumnBegin(BV bv, PetscInt j, NormType type, PetscReal *val);     /* bvglobal.c:703 */
PetscErrorCode BVNormVecBegin(BV bv, Vec v, NormType type, PetscReal *val);             /* bvglobal.c:575 (= VecNormBegin)      */
PetscErrorCode BVNormVecEnd(BV bv, Vec v, NormType type, PetscReal *val);               /* bvglobal.c:616                       */
/* PETSc's split-phase Vec reductions as bv/tests/test10.c uses them: stand-ins that evaluate at Begin (see host/bv.c) */
PetscErrorCode VecDotBegin(Vec x, Vec y, PetscScalar *val);
PetscErrorCode VecDotEnd(Vec x, Vec y, PetscScalar *val);
PetscErrorCode VecNormBegin(Vec x, NormType type, PetscReal *val);
PetscErrorCode VecNormEnd(Vec x, NormType type, PetscReal *val);
PetscErrorCode BVNormColumnEnd(BV bv, PetscInt j, NormType type, PetscReal *val);       /* bvglobal.c:760 */
PetscErrorCode BVMatMult(BV V, Mat A, BV Y);                                            /* bvops.c:767 */
PetscErrorCode BVMatMultColumn(BV V, Mat A, PetscInt j);                                /* bvops.c:862 */
PetscErrorCode BVOrthogonalizeVec(BV bv, Vec v, PetscScalar *H, PetscReal *norm, PetscBool *lindep);       /* bvorthog.c:249 */
PetscErrorCode BVOrthogonalizeColumn(BV bv, PetscInt j, PetscScalar *H, PetscReal *norm, PetscBool *lindep); /* bvorthog.c:315 */
PetscErrorCode BVOrthonormalizeColumn(BV bv, PetscInt j, PetscBool replace, PetscReal *norm, PetscBool *lindep); /* bvorthog.c:380 */
PetscErrorCode BVView(BV bv, PetscViewer viewer);                                       /* bvfunc.c:548 + svec.c:353 */
PetscErrorCode BVOrthogonalize(BV V, Mat R);                                            /* bvorthog.c:682 (block GS)   */
PetscErrorCode BVMatArnoldi(BV V, Mat A, Mat H, PetscInt k, PetscInt *m, PetscReal *beta, PetscBool *breakdown);  /* bvkrylov.c:56  */
PetscErrorCode BVMatLanczos(BV V, Mat A, Mat T, PetscInt k, PetscInt *m, PetscReal *beta, PetscBool *breakdown);  /* bvkrylov.c:165 */
/* instrumentation (PetscLogEvent stand-in): Gram-Schmidt passes and MatMults issued through this BV */
PetscErrorCode BVGetCounters(BV bv, int64_t *gs_passes, int64_t *matmults);

/* ---- DS (include/slepcds.h): HEP (compact), NHEP and SVD (compact), host LAPACK ---------------- */
#define DSHEP  "hep"
#define DSNHEP "nhep"
#define DSSVD  "svd"
typedef enum { DS_STATE_RAW = 0, DS_STATE_INTERMEDIATE, DS_STATE_CONDENSED, DS_STATE_TRUNCATED } DSStateType;
typedef enum { DS_MAT_A = 0, DS_MAT_B, DS_MAT_C, DS_MAT_T, DS_MAT_D, DS_MAT_Q, DS_MAT_Z, DS_MAT_X, DS_MAT_Y, DS_MAT_U, DS_MAT_V,
               DS_MAT_W, DS_NUM_MAT } DSMatType;
typedef PetscErrorCode (*SlepcEigenvalueComparisonFn)(PetscScalar ar, PetscScalar ai, PetscScalar br, PetscScalar bi, PetscInt *res, void *ctx);
PetscErrorCode SlepcCompareLargestMagnitude(PetscScalar, PetscScalar, PetscScalar, PetscScalar, PetscInt *, void *);   /* slepcsc.c:152 */
PetscErrorCode SlepcCompareSmallestMagnitude(PetscScalar, PetscScalar, PetscScalar, PetscScalar, PetscInt *, void *);
PetscErrorCode SlepcCompareLargestReal(PetscScalar, PetscScalar, PetscScalar, PetscScalar, PetscInt *, void *);
PetscErrorCode SlepcCompareSmallestReal(PetscScalar, PetscScalar, PetscScalar, PetscScalar, PetscInt *, void *);
PetscErrorCode SlepcCompareLargestImaginary(PetscScalar, PetscScalar, PetscScalar, PetscScalar, PetscInt *, void *);
PetscErrorCode SlepcCompareSmallestImaginary(PetscScalar, PetscScalar, PetscScalar, PetscScalar, PetscInt *, void *);
PetscErrorCode SlepcCompareTargetMagnitude(PetscScalar, PetscScalar, PetscScalar, PetscScalar, PetscInt *, void *);
PetscErrorCode SlepcCompareTargetReal(PetscScalar, PetscScalar, PetscScalar, PetscScalar, PetscInt *, void *);

PetscErrorCode DSCreate(DS *ds);
PetscErrorCode DSDestroy(DS *ds);
PetscErrorCode DSSetType(DS ds, const char *type);
PetscErrorCode DSSetCompact(DS ds, PetscBool comp);
PetscErrorCode DSSetExtraRow(DS ds, PetscBool ext);
PetscErrorCode DSAllocate(DS ds, PetscInt ld);
PetscErrorCode DSGetLeadingDimension(DS ds, PetscInt *ld);
PetscErrorCode DSSetDimensions(DS ds, PetscInt n, PetscInt l, PetscInt k);         /* dsops.c:129 */
PetscErrorCode DSGetDimensions(DS ds, PetscInt *n, PetscInt *l, PetscInt *k, PetscInt *t);
PetscErrorCode DSSVDSetDimensions(DS ds, PetscInt m);                              /* dssvd.c:636 */
PetscErrorCode DSSetState(DS ds, DSStateType state);
PetscErrorCode DSGetState(DS ds, DSStateType *state);
PetscErrorCode DSSetEigenvalueComparison(DS ds, SlepcEigenvalueComparisonFn fn, void *ctx);   /* DSSetSlepcSC equivalent */
PetscErrorCode DSGetArray(DS ds, DSMatType m, PetscScalar **a);
PetscErrorCode DSRestoreArray(DS ds, DSMatType m, PetscScalar **a);
PetscErrorCode DSGetArrayReal(DS ds, DSMatType m, PetscReal **a);
PetscErrorCode DSRestoreArrayReal(DS ds, DSMatType m, PetscReal **a);
PetscErrorCode DSGetMat(DS ds, DSMatType m, Mat *A);                               /* dsops.c:417 */
PetscErrorCode DSRestoreMat(DS ds, DSMatType m, Mat *A);
PetscErrorCode DSSolve(DS ds, PetscScalar eigr[], PetscScalar eigi[]);             /* dsops.c:723 */
PetscErrorCode DSSort(DS ds, PetscScalar eigr[], PetscScalar eigi[], PetscScalar rr[], PetscScalar ri[], PetscInt *k);  /* dsops.c:777 */
PetscErrorCode DSUpdateExtraRow(DS ds);
PetscErrorCode DSSynchronize(DS ds, PetscScalar eigr[], PetscScalar eigi[]);       /* redundant mode: no-op */
PetscErrorCode DSVectors(DS ds, DSMatType mat, PetscInt *j, PetscReal *rnorm);
PetscErrorCode DSTruncate(DS ds, PetscInt n, PetscBool trim);                      /* dsops.c:232 */
PetscErrorCode DSGetTruncateSize(DS ds, PetscInt l, PetscInt n, PetscInt *k);      /* dsops.c:368 */

/* ---- KSP: the linear solves behind ST for generalized problems and shift-and-invert.  The reference defaults to a direct
        solve (preonly + LU, stsles.c); a sparse factorisation is outside this path, so the solver here is Jacobi-preconditioned
        conjugate gradients on the device vectors (symmetric positive definite coefficient matrices: B of a GHEP, A - sigma B
        below the spectrum), tolerance SLEPC_DEFAULT_TOL * 1e-2 like STSetDefaultKSP (stsles.c:104-131) -------------------- */
typedef struct _p_KSP *KSP;
#define KSPCG "cg"
PetscErrorCode KSPCreate(KSP *ksp);
PetscErrorCode KSPDestroy(KSP *ksp);
PetscErrorCode KSPSetOperators(KSP ksp, Mat A, Mat P /* ignored: the Jacobi preconditioner is taken from A */);
PetscErrorCode KSPSetTolerances(KSP ksp, PetscReal rtol, PetscReal abstol, PetscReal dtol, PetscInt maxits);
PetscErrorCode KSPSetUp(KSP ksp);
PetscErrorCode KSPSolve(KSP ksp, Vec b, Vec x);
PetscErrorCode KSPGetIterationNumber(KSP ksp, PetscInt *its);
PetscErrorCode KSPGetTotalIterations(KSP ksp, PetscInt *its);

/* ---- ST (include/slepcst.h): shift and shift-and-invert ---------------------------------------------- */
#define STSHIFT   "shift"
#define STSINVERT "sinvert"
PetscErrorCode STCreate(ST *st);
PetscErrorCode STDestroy(ST *st);
PetscErrorCode STSetType(ST st, const char *type);                                 /* stfunc.c / stregis.c */
PetscErrorCode STGetType(ST st, const char **type);
PetscErrorCode STGetKSP(ST st, KSP *ksp);                                          /* stsles.c:133 */
PetscErrorCode STGetBilinearForm(ST st, Mat *B);                                   /* stfunc.c:578: B (NULL for a standard problem) */
PetscErrorCode STSetMatrices(ST st, PetscInt n, Mat A[]);
PetscErrorCode STSetShift(ST st, PetscScalar shift);
PetscErrorCode STGetShift(ST st, PetscScalar *shift);
PetscErrorCode STSetUp(ST st);
PetscErrorCode STApply(ST st, Vec x, Vec y);                                       /* stsolve.c:45 */
PetscErrorCode STGetOperator(ST st, Mat *Op);                                      /* stsolve.c:378 */
PetscErrorCode STRestoreOperator(ST st, Mat *Op);
PetscErrorCode STBackTransform(ST st, PetscInt n, PetscScalar *eigr, PetscScalar *eigi);   /* shift.c:49 */

/* ---- EPS (include/slepceps.h): Krylov-Schur, standard problems ------------------------------------- */
#define EPSKRYLOVSCHUR "krylovschur"
typedef enum { EPS_HEP = 1, EPS_GHEP = 2, EPS_NHEP = 3 } EPSProblemType;
typedef enum { EPS_LARGEST_MAGNITUDE = 1, EPS_SMALLEST_MAGNITUDE, EPS_LARGEST_REAL, EPS_SMALLEST_REAL, EPS_LARGEST_IMAGINARY,
               EPS_SMALLEST_IMAGINARY, EPS_TARGET_MAGNITUDE, EPS_TARGET_REAL } EPSWhich;
typedef enum { EPS_CONVERGED_TOL = 1, EPS_CONVERGED_USER = 2, EPS_DIVERGED_ITS = -1, EPS_DIVERGED_BREAKDOWN = -2,
               EPS_DIVERGED_SYMMETRY_LOST = -3, EPS_CONVERGED_ITERATING = 0 } EPSConvergedReason;
typedef enum { EPS_ERROR_ABSOLUTE = 0, EPS_ERROR_RELATIVE = 1, EPS_ERROR_BACKWARD = 2 } EPSErrorType;
typedef enum { EPS_CONV_ABS = 0, EPS_CONV_REL = 1 } EPSConv;
typedef PetscErrorCode (*EPSMonitorFn)(EPS eps, PetscInt its, PetscInt nconv, PetscScalar *eigr, PetscScalar *eigi, PetscReal *errest,
                                       PetscInt nest, void *ctx);

PetscErrorCode EPSCreate(EPS *eps);
PetscErrorCode EPSDestroy(EPS *eps);
PetscErrorCode EPSSetOperators(EPS eps, Mat A, Mat B /* NULL: standard problem; else A x = k B x with B symmetric positive definite (EPS_GHEP) */);
PetscErrorCode EPSSetPurify(EPS eps, PetscBool purify);                            /* epsopts.c:1469 */
PetscErrorCode EPSSetProblemType(EPS eps, EPSProblemType type);
PetscErrorCode EPSSetType(EPS eps, const char *type);
PetscErrorCode EPSSetDimensions(EPS eps, PetscInt nev, PetscInt ncv, PetscInt mpd);
PetscErrorCode EPSGetDimensions(EPS eps, PetscInt *nev, PetscInt *ncv, PetscInt *mpd);
PetscErrorCode EPSSetTolerances(EPS eps, PetscReal tol, PetscInt maxits);
PetscErrorCode EPSGetTolerances(EPS eps, PetscReal *tol, PetscInt *maxits);
PetscErrorCode EPSSetWhichEigenpairs(EPS eps, EPSWhich which);
PetscErrorCode EPSSetTarget(EPS eps, PetscScalar target);
PetscErrorCode EPSSetConvergenceTest(EPS eps, EPSConv conv);
PetscErrorCode EPSSetInitialSpace(EPS eps, PetscInt n, Vec is[]);                  /* epssetup.c:592 */
PetscErrorCode EPSKrylovSchurSetRestart(EPS eps, PetscReal keep);
PetscErrorCode EPSKrylovSchurSetLocking(EPS eps, PetscBool lock);
PetscErrorCode EPSMonitorSet(EPS eps, EPSMonitorFn monitor, void *ctx);
PetscErrorCode EPSGetBV(EPS eps, BV *V);
PetscErrorCode EPSGetDS(EPS eps, DS *ds);
PetscErrorCode EPSGetST(EPS eps, ST *st);
PetscErrorCode EPSSetUp(EPS eps);
PetscErrorCode EPSSolve(EPS eps);                                                  /* epssolve.c:119 */
/* run at most `nrestarts` more restart cycles of the current solve (bench: a "step" = one cycle) */
PetscErrorCode EPSSolveRestartCycles(EPS eps, PetscInt nrestarts, PetscInt *done);
PetscErrorCode EPSGetIterationNumber(EPS eps, PetscInt *its);
PetscErrorCode EPSGetConverged(EPS eps, PetscInt *nconv);
PetscErrorCode EPSGetConvergedReason(EPS eps, EPSConvergedReason *reason);
PetscErrorCode EPSGetEigenvalue(EPS eps, PetscInt i, PetscScalar *eigr, PetscScalar *eigi);
PetscErrorCode EPSGetEigenpair(EPS eps, PetscInt i, PetscScalar *eigr, PetscScalar *eigi, Vec Vr, Vec Vi);   /* epssolve.c:406 */
PetscErrorCode EPSGetErrorEstimate(EPS eps, PetscInt i, PetscReal *errest);
PetscErrorCode EPSComputeError(EPS eps, PetscInt i, EPSErrorType type, PetscReal *error);                     /* epssolve.c:742 */

/* ---- SVD (include/slepcsvd.h): thick-restart Lanczos ----------------------------------------------- */
#define SVDTRLANCZOS "trlanczos"
typedef enum { SVD_LARGEST = 0, SVD_SMALLEST = 1 } SVDWhich;
typedef enum { SVD_CONVERGED_TOL = 1, SVD_DIVERGED_ITS = -1, SVD_DIVERGED_BREAKDOWN = -2, SVD_CONVERGED_ITERATING = 0 } SVDConvergedReason;
typedef enum { SVD_ERROR_ABSOLUTE = 0, SVD_ERROR_RELATIVE = 1 } SVDErrorType;
PetscErrorCode SVDCreate(SVD *svd);
PetscErrorCode SVDDestroy(SVD *svd);
PetscErrorCode SVDSetOperators(SVD svd, Mat A, Mat B /* must be NULL */);
PetscErrorCode SVDSetTransposeMatrix(SVD svd, Mat AT);     /* optional explicit A^T supplied by the caller (multi-GPU) */
PetscErrorCode SVDSetImplicitTranspose(SVD svd, PetscBool impl);
PetscErrorCode SVDSetType(SVD svd, const char *type);
PetscErrorCode SVDSetDimensions(SVD svd, PetscInt nsv, PetscInt ncv, PetscInt mpd);
PetscErrorCode SVDGetDimensions(SVD svd, PetscInt *nsv, PetscInt *ncv, PetscInt *mpd);
PetscErrorCode SVDSetTolerances(SVD svd, PetscReal tol, PetscInt maxits);
PetscErrorCode SVDSetWhichSingularTriplets(SVD svd, SVDWhich which);
PetscErrorCode SVDSetInitialSpaces(SVD svd, PetscInt nr, Vec isr[], PetscInt nl, Vec isl[]);
PetscErrorCode SVDTRLanczosSetRestart(SVD svd, PetscReal keep);
PetscErrorCode SVDTRLanczosSetLocking(SVD svd, PetscBool lock);
PetscErrorCode SVDTRLanczosSetOneSide(SVD svd, PetscBool oneside);
PetscErrorCode SVDGetBV(SVD svd, BV *V, BV *U);
PetscErrorCode SVDSetUp(SVD svd);
PetscErrorCode SVDSolve(SVD svd);                                                  /* svdsolve.c:116 */
PetscErrorCode SVDGetIterationNumber(SVD svd, PetscInt *its);
PetscErrorCode SVDGetConverged(SVD svd, PetscInt *nconv);
PetscErrorCode SVDGetConvergedReason(SVD svd, SVDConvergedReason *reason);
PetscErrorCode SVDGetSingularTriplet(SVD svd, PetscInt i, PetscReal *sigma, Vec u, Vec v);
PetscErrorCode SVDComputeError(SVD svd, PetscInt i, SVDErrorType type, PetscReal *error);   /* svdsolve.c:486 */

#ifdef __cplusplus
}
#endif
#endif
