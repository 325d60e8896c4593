/* private structures of libb2kslepc.so — counterpart of include/slepc/private/{bv,ds,st,eps,svd}impl.h */
#ifndef B2KIMPL_H
#define B2KIMPL_H
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "b2kslepc.h"

/* ---- error handling (PetscCall / SETERRQ look-alikes) ------------------------------------------- */
PetscErrorCode B2KErrorSet(PetscErrorCode code, const char *func, const char *file, int line, const char *fmt, ...);
PetscErrorCode B2KErrorTrace(PetscErrorCode code, const char *func, const char *file, int line);
#define SETERRQ(code, ...) return B2KErrorSet((code), __func__, __FILE__, __LINE__, __VA_ARGS__)
#define PetscCheck(cond, code, ...)                                                          \
  do {                                                                                       \
    if (!(cond)) return B2KErrorSet((code), __func__, __FILE__, __LINE__, __VA_ARGS__);      \
  } while (0)
#define PetscCall(call)                                                                      \
  do {                                                                                       \
    PetscErrorCode ierr_ = (call);                                                           \
    if (ierr_) return B2KErrorTrace(ierr_, __func__, __FILE__, __LINE__);                    \
  } while (0)
/* calls into libb200krylov (include/b2k.h) */
#define B2KCall(call)                                                                        \
  do {                                                                                       \
    int rc_ = (call);                                                                        \
    if (rc_) return B2KErrorSet(PETSC_ERR_GPU, __func__, __FILE__, __LINE__, "%s failed (%d): %s", #call, rc_, b2k_last_error()); \
  } while (0)
#define PetscMax(a, b) (((a) < (b)) ? (b) : (a))
#define PetscMin(a, b) (((a) < (b)) ? (a) : (b))
#define PetscAbsReal(a) fabs(a)
#define PetscSqrtReal(a) sqrt(a)
static inline PetscReal SlepcAbsEigenvalue(PetscScalar re, PetscScalar im) { return hypot(re, im); }

/* ---- host BLAS/LAPACK (scipy-bundled OpenBLAS, LP64, `scipy_` prefix) ------------------------------ */
#define LAPACK(name) scipy_##name##_
void LAPACK(dgeqrf)(const int *, const int *, double *, const int *, double *, double *, const int *, int *);
void LAPACK(dorgqr)(const int *, const int *, const int *, double *, const int *, const double *, double *, const int *, int *);
void LAPACK(dormqr)(const char *, const char *, const int *, const int *, const int *, const double *, const int *, const double *, double *,
                    const int *, double *, const int *, int *, size_t, size_t);
void LAPACK(dsteqr)(const char *, const int *, double *, double *, double *, const int *, double *, int *, size_t);
void LAPACK(dlartg)(const double *, const double *, double *, double *, double *);
void LAPACK(drot)(const int *, double *, const int *, double *, const int *, const double *, const double *);
void LAPACK(dgehrd)(const int *, const int *, const int *, double *, const int *, double *, double *, const int *, int *);
void LAPACK(dorghr)(const int *, const int *, const int *, double *, const int *, const double *, double *, const int *, int *);
void LAPACK(dhseqr)(const char *, const char *, const int *, const int *, const int *, double *, const int *, double *, double *,
                    double *, const int *, double *, const int *, int *, size_t, size_t);
void LAPACK(dtrexc)(const char *, const int *, double *, const int *, double *, const int *, int *, int *, double *, int *, size_t);
void LAPACK(dtrevc)(const char *, const char *, int *, const int *, double *, const int *, double *, const int *, double *,
                    const int *, const int *, int *, double *, int *, size_t, size_t);
void LAPACK(dbdsqr)(const char *, const int *, const int *, const int *, const int *, double *, double *, double *, const int *,
                    double *, const int *, double *, const int *, double *, int *, size_t);
void LAPACK(dgemv)(const char *, const int *, const int *, const double *, const double *, const int *, const double *,
                   const int *, const double *, double *, const int *, size_t);
double LAPACK(dnrm2)(const int *, const double *, const int *);
void LAPACK(dpotrf)(const char *, const int *, double *, const int *, int *, size_t);
void LAPACK(dtrtri)(const char *, const char *, const int *, double *, const int *, int *, size_t, size_t);
void LAPACK(dsyev)(const char *, const char *, const int *, double *, const int *, double *, double *, const int *, int *, size_t, size_t);

/* ---- communicator ------------------------------------------------------------------------------- */
struct _p_B2KComm {
  int            rank, size;
  int            kind;          /* 0 self, 1 nccl (device buffers), 2 callbacks (host buffers) */
  b2k_comm       nccl;
  B2KAllreduceFn ar;
  B2KSendrecvFn  sr;
  void          *user;
};

/* ---- common object header (stands in for PetscObject: class id + the name given with PetscObjectSetName) ---- */
#define B2K_VEC_CLASSID 0x56454300
#define B2K_MAT_CLASSID 0x4d415400
#define B2K_BV_CLASSID  0x42560000
typedef struct { int classid; char name[64]; } B2KObjectHeader;     /* FIRST member of _p_Vec, _p_Mat, _p_BV */
struct _p_PetscViewer { PetscViewerFormat fmt[8]; int depth; };

/* ---- Vec / Mat ------------------------------------------------------------------------------------ */
struct _p_Vec {
  B2KObjectHeader hdr;
  PetscInt     n, N;
  PetscInt     rstart;     /* global index of the first local entry (PetscLayout rstart) */
  B2KMemType   mem;
  PetscScalar *array;
  PetscScalar *saved;      /* VecPlaceArray / VecResetArray */
  PetscBool    owns;
};

typedef struct _MatOps {
  PetscErrorCode (*mult)(Mat, Vec, Vec);
  PetscErrorCode (*multtranspose)(Mat, Vec, Vec);
  PetscErrorCode (*destroy)(Mat);
  /* optional: Y(:,0:k) = A X(:,0:k) on column-major blocks in the Mat's own memory space (MatMatMult on a dense block, what
     BVMatMult uses in BV_MATMULT_MAT mode, svec.c:203-231); NULL = the caller loops over the columns with mult */
  PetscErrorCode (*multblock)(Mat, const PetscScalar *, PetscInt, PetscScalar *, PetscInt, PetscInt);
  PetscErrorCode (*getdiagonal)(Mat, Vec);
} MatOps;

struct _p_Mat {
  B2KObjectHeader hdr;
  MatOps       ops;
  char         type[24];
  PetscInt     m, n, M, N;           /* local / global sizes            */
  PetscInt     rstart, rend;         /* owned row range                 */
  PetscInt     cstart, cend;         /* owned range of the column space */
  B2KMemType   mem;                  /* where its vectors live          */
  /* seqdense */
  PetscScalar *dense;
  PetscInt     lda;
  PetscBool    owns_dense;
  void        *data;                 /* type-specific / shell context   */
};

/* ---- BV (bvimpl.h:25-113) -------------------------------------------------------------------------- */
typedef struct _BVOps {
  PetscErrorCode (*mult)(BV, PetscScalar, PetscScalar, BV, Mat);
  PetscErrorCode (*multvec)(BV, PetscScalar, PetscScalar, Vec, PetscScalar *);
  PetscErrorCode (*multinplace)(BV, Mat, PetscInt, PetscInt);
  PetscErrorCode (*multinplacetrans)(BV, Mat, PetscInt, PetscInt);
  PetscErrorCode (*dot)(BV, BV, Mat);
  PetscErrorCode (*dotvec)(BV, Vec, PetscScalar *);
  PetscErrorCode (*dotvec_local)(BV, Vec, PetscScalar *);
  PetscErrorCode (*dotvec_begin)(BV, Vec, PetscScalar *);                    /* bvimpl.h:33 */
  PetscErrorCode (*dotvec_end)(BV, Vec, PetscScalar *);                      /* bvimpl.h:34 */
  PetscErrorCode (*scale)(BV, PetscInt, PetscScalar);
  PetscErrorCode (*norm)(BV, PetscInt, NormType, PetscReal *);
  PetscErrorCode (*norm_local)(BV, PetscInt, NormType, PetscReal *);
  PetscErrorCode (*norm_begin)(BV, PetscInt, NormType, PetscReal *);         /* bvimpl.h:38 */
  PetscErrorCode (*norm_end)(BV, PetscInt, NormType, PetscReal *);           /* bvimpl.h:39 */
  PetscErrorCode (*normalize)(BV, PetscScalar *);
  PetscErrorCode (*matmult)(BV, Mat, BV);
  PetscErrorCode (*copy)(BV, BV);
  PetscErrorCode (*copycolumn)(BV, PetscInt, PetscInt);
  PetscErrorCode (*resize)(BV, PetscInt, PetscBool);
  PetscErrorCode (*getcolumn)(BV, PetscInt, Vec *);
  PetscErrorCode (*restorecolumn)(BV, PetscInt, Vec *);
  PetscErrorCode (*getarray)(BV, PetscScalar **);
  PetscErrorCode (*restorearray)(BV, PetscScalar **);
  PetscErrorCode (*gramschmidt)(BV, PetscInt, Vec, PetscBool *, PetscScalar *, PetscScalar *, PetscReal *, PetscReal *);
  PetscErrorCode (*destroy)(BV);
  /* extensions of this build (NULL is always allowed) */
  PetscErrorCode (*setrandomcolumn)(BV, PetscInt);        /* deterministic hash fill          */
  PetscErrorCode (*duplicate)(BV, BV);                    /* bvimpl.h:56                      */
  /* Householder QR of the LOCAL rows of the active columns l..k-1 (what BVOrthogonalize_LAPACK_TSQR / _TSQR_OnlyR do on the raw
     array, bvlapack.c:347-560): tsqr_factor returns the (k-l) x (k-l) upper-triangular factor of this rank's rows in R (column-major,
     leading dimension k-l); with wantq the type keeps what it needs so that tsqr_formq(W) overwrites the columns with Q_local * W */
  /* Krylov steps j = k..m-1 (v_{j+1} = A v_j, orthonormalised against v_0..v_j with the DGKS rule, coefficients into the buffer
     exactly as BVMatMultColumn + BVOrthonormalizeColumn leave them) enqueued WITHOUT host round trips; *jnext = first step the
     type did not complete (k when it declines, m when all went through; a breakdown or a third DGKS pass stops it early and the
     front-end's step-by-step loop takes over from there), *beta = the norm of the last completed step */
  PetscErrorCode (*krylov_steps)(BV, Mat A, PetscInt k, PetscInt m, PetscInt *jnext, PetscReal *beta);
  PetscErrorCode (*tsqr_factor)(BV, PetscBool wantq, PetscScalar *R);
  PetscErrorCode (*tsqr_formq)(BV, const PetscScalar *W);
} BVOps;

struct _p_BV {
  B2KObjectHeader    hdr;
  BVOps              ops;
  char               type_name[24];
  B2KComm            comm;
  PetscInt           n, N;          /* local / global rows                 */
  PetscInt           m;             /* columns                             */
  PetscInt           l, k;          /* active window                        */
  PetscInt           nc;            /* constraints (always 0 on this path)  */
  PetscInt           ld;            /* leading dimension (elements)         */
  PetscInt           row0;          /* global index of the first local row  */
  BVOrthogType       orthog_type;
  BVOrthogRefineType orthog_ref;
  PetscReal          orthog_eta;
  BVOrthogBlockType  orthog_block;
  PetscScalar       *h, *c;         /* coefficient work arrays for BVOrthogonalizeVec (host) */
  PetscScalar       *buffer;        /* (nc+m) x m host coefficients: col 0 scratch, col j = h_j   */
  Vec                cv[2];         /* lent column views                    */
  PetscInt           ci[2];
  B2KMemType         mem;
  uint64_t           rng_seed;
  uint64_t           state;         /* PetscObjectState stand-in           */
  int64_t            n_gs_passes, n_matmult;
  PetscBool          sizes_set, type_set;
  PetscErrorCode   (*ctor)(BV);   /* constructor deferred until the sizes are known (bvbasic.c:56-62) */
  PetscScalar       *work;        /* BVAllocateWork_Private (bvfunc.c:654) */
  size_t             lwork;
  Mat                matrix;      /* inner-product matrix B (not owned), bvimpl.h:79 */
  PetscBool          indef;
  Vec                Bx;          /* B times the vector of the current inner product (BV_IPMatMult, bvimpl.h:147) */
  BV                 cached;      /* B times the active columns (BV_IPMatMultBV, bvimpl.h:164) */
  void              *data;
};
#define BV_BUF(bv, i, j) ((bv)->buffer[(size_t)(i) + (size_t)(j) * ((bv)->nc + (bv)->m)])
PetscErrorCode BVCreate_B200(BV bv);
PetscErrorCode BVForgetSizes_Private(BV bv);     /* bv.c: the basis of a solver whose operator changed size */
/* h += c on rows 0..nc+j-1 (NULL = the buffer: column j += column 0) — BV_AddCoefficients bvimpl.h:308-322 */
static inline void BV_AddCoefficients(BV bv, PetscInt j, PetscScalar *h, PetscScalar *c)
{
  if (!h) { h = bv->buffer + (size_t)j * (size_t)(bv->nc + bv->m); c = bv->buffer; }
  for (PetscInt i = 0; i < bv->nc + j; i++) h[i] += c[i];
}

/* ---- DS -------------------------------------------------------------------------------------------- */
struct _p_DS {
  char         type[8];
  PetscInt     ld, n, l, k, t;
  PetscInt     m, tm;            /* DSSVD: columns, and columns before truncation (dssvd.c:14-17) */
  DSStateType  state;
  PetscBool    compact, extrarow, allocated;
  PetscScalar *mat[DS_NUM_MAT];  /* ld x ld each (T: ld x 3)                                        */
  PetscInt    *perm;
  PetscScalar *work;
  PetscInt    *iwork;
  SlepcEigenvalueComparisonFn cmp;
  void        *cmpctx;
  Mat          lent[DS_NUM_MAT];
};

/* ---- ST -------------------------------------------------------------------------------------------- */
struct _p_KSP {
  Mat       A;
  PetscReal rtol, abstol;
  PetscInt  max_it, its, total_its;
  Vec       dinv, r, z, p, q;    /* Jacobi preconditioner and the CG work vectors */
  PetscBool setup;
};

struct _p_ST {
  Mat         A, B;              /* the problem matrices (B = NULL: standard problem)            */
  PetscInt    nmat;
  PetscScalar sigma;
  PetscBool   sinvert;
  Mat         Op;                /* shell of the transformed operator (NULL: A itself)            */
  Mat         T;                 /* shell A - sigma B (coefficient matrix of the solve, or the multiplied one) */
  KSP         ksp;
  Vec         w;                 /* work vector of STApply                                        */
  PetscBool   setup;
};

/* ---- EPS ------------------------------------------------------------------------------------------- */
struct _p_EPS {
  PetscInt     nev, ncv, mpd, max_it, n, nloc;
  PetscReal    tol;
  EPSWhich     which;
  EPSProblemType problem_type;
  EPSConv      conv;
  PetscScalar  target;
  PetscBool    ishermitian, isgeneralized, purify;
  Mat          B;
  PetscReal    keep;
  PetscBool    lock;
  ST           st;
  BV           V;
  DS           ds;
  PetscScalar *eigr, *eigi;
  PetscReal   *errest;
  PetscInt    *perm;
  PetscInt     nconv, its;
  EPSConvergedReason reason;
  PetscInt     nini;
  Vec          inivec;           /* first initial vector (copied at setup)     */
  PetscBool    setup_done, vectors_done;
  /* state of a solve in progress (EPSSolveRestartCycles) */
  PetscInt     l;
  PetscBool    started;
  EPSMonitorFn monitor;
  void        *monitorctx;
  Vec          work[5];
  PetscBool    solved;
  PetscInt     allocated;        /* size of eigr/eigi/errest/perm */
  struct { SlepcEigenvalueComparisonFn fn; void *ctx; PetscScalar sigma; } sc;   /* SlepcSC with map = STBackTransform */
  int64_t      n_restarts_bv, n_dssolve;
};

/* ---- SVD ------------------------------------------------------------------------------------------- */
struct _p_SVD {
  PetscInt     nsv, ncv, mpd, max_it;
  PetscReal    tol;
  SVDWhich     which;
  PetscBool    impltrans, swapped, oneside, lock, owns_AT;
  PetscScalar *swork;         /* ncv+1 coefficients of the one-sided recurrence (trlanczos.c:463) */
  PetscReal    keep;
  Mat          OP, A, AT;        /* user matrix; working pair after the M<N swap (svdsetup.c:301-306) */
  Mat          userAT;
  BV           U, V;
  DS           ds;
  PetscReal   *sigma, *errest;
  PetscInt    *perm;
  PetscInt     nconv, its;
  SVDConvergedReason reason;
  Vec          iniV, iniU;
  PetscBool    setup_done, solved, started;
  PetscInt     l;                /* state of a solve in progress */
  PetscInt     allocated;
  Vec          work[4];
};

/* helpers shared between the files */
uint64_t     B2KHashMix(uint64_t idx, uint64_t seed);
double       B2KHashUniform(uint64_t idx, uint64_t seed);
PetscErrorCode MatCreate_Private(Mat *A);
#endif
