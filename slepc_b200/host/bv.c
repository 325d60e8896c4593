/*
 * bv.c — front-end of the BV class for the Krylov hot path.  Same role as SLEPc's
 * src/sys/classes/bv/interface/{bvbasic,bvops,bvglobal,bvorthog,bvkrylov,bvfunc,bvregis}.c:
 * argument validation, active-window bookkeeping, the DGKS Gram-Schmidt driver and the
 * Arnoldi/Lanczos recurrences.  All arithmetic on the n-long dimension is delegated through the
 * ops table (b2kimpl.h, counterpart of bvimpl.h:25-61) to the registered BV type; the only type
 * this library ships is "b200" (bvb200.c, sm_100a kernels).  There is no CPU arithmetic here.
 */
#include "b2kimpl.h"

/* ---- registry: BVRegister bvfunc.c:646, BVRegisterAll bvregis.c:29-39 ------------------------- */
typedef struct { char name[24]; PetscErrorCode (*ctor)(BV); } BVRegEntry;
static BVRegEntry g_bvlist[16];
static int g_nbv = 0;
static PetscBool g_bv_registered_all = PETSC_FALSE;

PetscErrorCode BVRegister(const char *name, PetscErrorCode (*ctor)(BV))
{
  PetscCheck(name && ctor, PETSC_ERR_ARG_NULL, "null argument");
  PetscCheck(strlen(name) < sizeof(g_bvlist[0].name), PETSC_ERR_ARG_SIZ, "BV type name too long");
  for (int i = 0; i < g_nbv; i++)
    if (!strcmp(g_bvlist[i].name, name)) { g_bvlist[i].ctor = ctor; return PETSC_SUCCESS; }
  PetscCheck(g_nbv < 16, PETSC_ERR_MEM, "too many BV types registered");
  strcpy(g_bvlist[g_nbv].name, name);
  g_bvlist[g_nbv++].ctor = ctor;
  return PETSC_SUCCESS;
}

static PetscErrorCode BVRegisterAll(void)
{
  if (g_bv_registered_all) return PETSC_SUCCESS;
  g_bv_registered_all = PETSC_TRUE;
  PetscCall(BVRegister(BVB200, BVCreate_B200));
  return PETSC_SUCCESS;
}

#define BVCheckSizes(bv) PetscCheck((bv)->sizes_set && (bv)->type_set, PETSC_ERR_ARG_WRONGSTATE, "BV sizes and type must be set first")
#define BVCheckOp(bv, op) PetscCheck((bv)->ops.op, PETSC_ERR_SUP, "BV type %s does not implement operation " #op, (bv)->type_name)

/* ---- lifecycle (bvbasic.c, bvfunc.c) ------------------------------------------------------------ */
PetscErrorCode BVCreate(BV *out)
{
  PetscCall(BVRegisterAll());
  BV bv = (BV)calloc(1, sizeof(*bv));
  if (bv) bv->hdr.classid = B2K_BV_CLASSID;
  PetscCheck(bv, PETSC_ERR_MEM, "out of memory");
  bv->comm = B2KCommWorld();
  bv->orthog_type = BV_ORTHOG_CGS;                /* bvfunc.c:176-179 */
  bv->orthog_ref = BV_ORTHOG_REFINE_IFNEEDED;
  bv->orthog_eta = 0.7071;
  bv->orthog_block = BV_ORTHOG_BLOCK_GS;
  bv->ci[0] = bv->ci[1] = -1;
  bv->rng_seed = 0x5EED;
  *out = bv;
  return PETSC_SUCCESS;
}

static PetscErrorCode BVReset_Private(BV bv)
{
  if (bv->type_set && bv->ops.destroy) PetscCall(bv->ops.destroy(bv));
  memset(&bv->ops, 0, sizeof(bv->ops));
  bv->data = NULL;
  bv->type_set = PETSC_FALSE;
  PetscCall(VecDestroy(&bv->cv[0]));
  PetscCall(VecDestroy(&bv->cv[1]));
  free(bv->buffer); bv->buffer = NULL;
  free(bv->h); bv->h = NULL;
  free(bv->c); bv->c = NULL;
  free(bv->work); bv->work = NULL; bv->lwork = 0;
  PetscCall(VecDestroy(&bv->Bx));
  PetscCall(BVDestroy(&bv->cached));
  return PETSC_SUCCESS;
}

PetscErrorCode BVDestroy(BV *bv)
{
  if (!bv || !*bv) return PETSC_SUCCESS;
  PetscCall(BVReset_Private(*bv));
  free(*bv);
  *bv = NULL;
  return PETSC_SUCCESS;
}

/* what EPSReset / SVDReset do to the basis when a solver is given an operator of another size (epssetup.c:441-458,
   svdsetup.c:107-116: the BV is destroyed and created again at the next set-up): the storage and the sizes go, the type, the
   orthogonalisation options and the random seed stay */
PetscErrorCode BVForgetSizes_Private(BV bv)
{
  PetscErrorCode (*ctor)(BV) = bv->ctor;
  PetscCall(BVReset_Private(bv));
  bv->sizes_set = PETSC_FALSE;
  bv->n = bv->N = bv->m = 0; bv->l = bv->k = 0; bv->nc = 0; bv->ld = 0; bv->row0 = 0;
  bv->ci[0] = bv->ci[1] = -1;
  bv->matrix = NULL;
  bv->ctor = ctor;                                /* BVConstruct_Private runs it again once the new sizes are known */
  return PETSC_SUCCESS;
}

static PetscErrorCode BVAllocateBuffers_Private(BV bv)
{
  const size_t ldb = (size_t)(bv->nc + bv->m);
  free(bv->buffer); free(bv->h); free(bv->c);
  bv->buffer = (PetscScalar *)calloc(ldb * (size_t)(bv->m > 0 ? bv->m : 1) + 1, sizeof(PetscScalar));   /* bvbasic.c:757-789 */
  bv->h = (PetscScalar *)calloc(ldb + 1, sizeof(PetscScalar));
  bv->c = (PetscScalar *)calloc(ldb + 1, sizeof(PetscScalar));
  PetscCheck(bv->buffer && bv->h && bv->c, PETSC_ERR_MEM, "out of memory");
  return PETSC_SUCCESS;
}

static PetscErrorCode BVConstruct_Private(BV bv)
{
  if (!bv->ctor || !bv->sizes_set || bv->type_set) return PETSC_SUCCESS;
  bv->ld = bv->n + (bv->n & 1);                   /* BV_SetDefaultLD bvimpl.h:471-484: 16-byte multiple */
  if (bv->ld < 2) bv->ld = 2;
  bv->l = 0; bv->k = bv->m;
  PetscCall(BVAllocateBuffers_Private(bv));
  PetscCall(bv->ctor(bv));
  bv->type_set = PETSC_TRUE;
  PetscCall(VecCreateWithArray(bv->mem, bv->n, bv->N, NULL, &bv->cv[0]));   /* BVCreateVecEmpty svec.c:486-487 */
  PetscCall(VecCreateWithArray(bv->mem, bv->n, bv->N, NULL, &bv->cv[1]));
  bv->cv[0]->rstart = bv->cv[1]->rstart = bv->row0;
  return PETSC_SUCCESS;
}

PetscErrorCode BVSetType(BV bv, const char *type)
{
  PetscCheck(type, PETSC_ERR_ARG_NULL, "null type");
  if (bv->type_set && !strcmp(bv->type_name, type)) return PETSC_SUCCESS;
  PetscCall(BVRegisterAll());
  PetscErrorCode (*ctor)(BV) = NULL;
  for (int i = 0; i < g_nbv; i++)
    if (!strcmp(g_bvlist[i].name, type)) ctor = g_bvlist[i].ctor;
  PetscCheck(ctor, PETSC_ERR_ARG_UNKNOWN_TYPE, "Unable to find requested BV type %s", type);
  PetscCall(BVReset_Private(bv));
  strcpy(bv->type_name, type);
  bv->ctor = ctor;
  PetscCall(BVConstruct_Private(bv));              /* deferred until sizes are known, bvbasic.c:56-62 */
  return PETSC_SUCCESS;
}

PetscErrorCode BVGetType(BV bv, const char **type) { *type = bv->type_name; return PETSC_SUCCESS; }

PetscErrorCode BVSetSizes(BV bv, PetscInt n, PetscInt N, PetscInt m)
{
  PetscCheck(!(N >= 0 && n > N), PETSC_ERR_ARG_INCOMP, "Local size %d cannot be larger than global size %d", n, N);
  PetscCheck(m > 0, PETSC_ERR_ARG_INCOMP, "Number of columns %d must be positive", m);
  PetscCheck(n >= 0, PETSC_ERR_ARG_INCOMP, "Local size must be given");
  PetscCheck(!bv->sizes_set, PETSC_ERR_ARG_WRONGSTATE, "Cannot change/reset vector sizes to %d local %d global after previously setting them to %d local %d global", n, N, bv->n, bv->N);
  int size = 1;
  PetscCall(B2KCommGetRank(bv->comm, NULL, &size));
  PetscCheck(N >= 0 || size == 1, PETSC_ERR_ARG_WRONG, "the global size must be given on more than one rank");
  bv->n = n; bv->N = (N < 0) ? n : N; bv->m = m; bv->k = m; bv->l = 0;
  bv->sizes_set = PETSC_TRUE;
  PetscCall(BVConstruct_Private(bv));
  return PETSC_SUCCESS;
}

PetscErrorCode BVSetSizesFromVec(BV bv, Vec t, PetscInt m)
{
  PetscCheck(t, PETSC_ERR_ARG_NULL, "null template vector");
  PetscCheck(!bv->sizes_set, PETSC_ERR_ARG_WRONGSTATE, "Vector layout was already defined by a previous call to BVSetSizes/FromVec");
  bv->row0 = t->rstart;
  bv->mem = t->mem;
  PetscCall(BVSetSizes(bv, t->n, t->N, m));
  return PETSC_SUCCESS;
}

PetscErrorCode BVGetSizes(BV bv, PetscInt *n, PetscInt *N, PetscInt *m)
{
  if (n) *n = bv->n;
  if (N) *N = bv->N;
  if (m) *m = bv->m;
  return PETSC_SUCCESS;
}
PetscErrorCode BVGetLeadingDimension(BV bv, PetscInt *ld) { BVCheckSizes(bv); *ld = bv->ld; return PETSC_SUCCESS; }
PetscErrorCode BVGetMemType(BV bv, B2KMemType *mem) { *mem = bv->mem; return PETSC_SUCCESS; }

PetscErrorCode BVSetActiveColumns(BV bv, PetscInt l, PetscInt k)
{
  BVCheckSizes(bv);
  if (k == PETSC_DECIDE || k == PETSC_DEFAULT) k = bv->m;
  else PetscCheck(k >= 0 && k <= bv->m, PETSC_ERR_ARG_OUTOFRANGE, "Illegal value of k (%d). Must be between 0 and m (%d)", k, bv->m);
  if (l == PETSC_DECIDE || l == PETSC_DEFAULT) l = 0;
  else PetscCheck(l >= 0 && l <= k, PETSC_ERR_ARG_OUTOFRANGE, "Illegal value of l (%d). Must be between 0 and k (%d)", l, k);
  bv->l = l; bv->k = k;
  return PETSC_SUCCESS;
}
PetscErrorCode BVGetActiveColumns(BV bv, PetscInt *l, PetscInt *k) { if (l) *l = bv->l; if (k) *k = bv->k; return PETSC_SUCCESS; }

PetscErrorCode BVSetOrthogonalization(BV bv, BVOrthogType type, BVOrthogRefineType refine, PetscReal eta, BVOrthogBlockType block)
{
  PetscCheck(type == BV_ORTHOG_CGS || type == BV_ORTHOG_MGS, PETSC_ERR_ARG_WRONG, "Unknown orthogonalization type");
  PetscCheck(refine == BV_ORTHOG_REFINE_NEVER || refine == BV_ORTHOG_REFINE_IFNEEDED || refine == BV_ORTHOG_REFINE_ALWAYS, PETSC_ERR_ARG_WRONG, "Unknown refinement type");
  if (eta == (PetscReal)PETSC_DEFAULT || eta == (PetscReal)PETSC_DETERMINE) eta = 0.7071;
  else if (eta != (PetscReal)PETSC_CURRENT) { PetscCheck(eta > 0.0 && eta <= 1.0, PETSC_ERR_ARG_OUTOFRANGE, "Invalid eta value"); }
  else eta = bv->orthog_eta;
  PetscCheck(block == BV_ORTHOG_BLOCK_GS || block == BV_ORTHOG_BLOCK_CHOL || block == BV_ORTHOG_BLOCK_SVQB || block == BV_ORTHOG_BLOCK_TSQR || block == BV_ORTHOG_BLOCK_TSQRCHOL,
             PETSC_ERR_ARG_WRONG, "Unknown block orthogonalization type");
  bv->orthog_type = type; bv->orthog_ref = refine; bv->orthog_eta = eta; bv->orthog_block = block;
  return PETSC_SUCCESS;
}
PetscErrorCode BVGetOrthogonalization(BV bv, BVOrthogType *type, BVOrthogRefineType *refine, PetscReal *eta, BVOrthogBlockType *block)
{
  if (type) *type = bv->orthog_type;
  if (refine) *refine = bv->orthog_ref;
  if (eta) *eta = bv->orthog_eta;
  if (block) *block = bv->orthog_block;
  return PETSC_SUCCESS;
}

PetscErrorCode BVResize(BV bv, PetscInt m, PetscBool copy)
{
  BVCheckSizes(bv);
  PetscCheck(m > 0, PETSC_ERR_ARG_OUTOFRANGE, "Number of columns %d must be positive", m);
  PetscCheck(bv->nc == 0 || !copy, PETSC_ERR_ARG_WRONGSTATE, "Cannot resize a BV with constraints");
  if (bv->m == m) return PETSC_SUCCESS;
  BVCheckOp(bv, resize);
  PetscCall(bv->ops.resize(bv, m, copy));
  bv->m = m;
  bv->k = PetscMin(bv->k, m);
  bv->l = PetscMin(bv->l, m);
  PetscCall(BVAllocateBuffers_Private(bv));
  bv->state++;
  return PETSC_SUCCESS;
}

static PetscErrorCode BVDuplicate_Private(BV V, PetscInt m, BV *W)
{
  BVCheckSizes(V);
  BV w;
  PetscCall(BVCreate(&w));
  w->comm = V->comm; w->row0 = V->row0; w->mem = V->mem;
  w->orthog_type = V->orthog_type; w->orthog_ref = V->orthog_ref; w->orthog_eta = V->orthog_eta; w->orthog_block = V->orthog_block;
  w->rng_seed = V->rng_seed;
  PetscCall(BVSetSizes(w, V->n, V->N, m));
  PetscCall(BVSetType(w, V->type_name));
  if (V->ops.duplicate) PetscCall(V->ops.duplicate(V, w));
  w->matrix = V->matrix; w->indef = V->indef;            /* bvbasic.c:1413 */
  *W = w;
  return PETSC_SUCCESS;
}
PetscErrorCode BVDuplicate(BV V, BV *W) { return BVDuplicate_Private(V, V->m, W); }
PetscErrorCode BVDuplicateResize(BV V, PetscInt m, BV *W) { return BVDuplicate_Private(V, m, W); }

/* ---- column access: bvbasic.c:1116-1200 ----------------------------------------------------------- */
PetscErrorCode BVGetColumn(BV bv, PetscInt j, Vec *v)
{
  BVCheckSizes(bv);
  BVCheckOp(bv, getcolumn);
  PetscCheck(j >= 0 || -j <= bv->nc, PETSC_ERR_ARG_OUTOFRANGE, "You requested constraint %d but only %d are available", -j, bv->nc);
  PetscCheck(j < bv->m, PETSC_ERR_ARG_OUTOFRANGE, "You requested column %d but only %d are available", j, bv->m);
  PetscCheck(j != bv->ci[0] && j != bv->ci[1], PETSC_ERR_SUP, "Column %d already fetched in a previous call to BVGetColumn", j);
  const PetscInt none = -bv->nc - 1;
  const int l = (bv->ci[0] == none) ? 0 : ((bv->ci[1] == none) ? 1 : -1);
  PetscCheck(l != -1, PETSC_ERR_SUP, "Too many requested columns; you must call BVRestoreColumn for one of the previously fetched columns");
  bv->ci[l] = j;                                  /* the impl places the array into cv[l] */
  PetscErrorCode ierr = bv->ops.getcolumn(bv, j, v);
  if (ierr) { bv->ci[l] = none; return B2KErrorTrace(ierr, __func__, __FILE__, __LINE__); }
  *v = bv->cv[l];
  return PETSC_SUCCESS;
}

PetscErrorCode BVRestoreColumn(BV bv, PetscInt j, Vec *v)
{
  BVCheckSizes(bv);
  PetscCheck(v && *v, PETSC_ERR_ARG_NULL, "null vector");
  PetscCheck(j >= 0 || -j <= bv->nc, PETSC_ERR_ARG_OUTOFRANGE, "You requested constraint %d but only %d are available", -j, bv->nc);
  PetscCheck(j < bv->m, PETSC_ERR_ARG_OUTOFRANGE, "You requested column %d but only %d are available", j, bv->m);
  PetscCheck(j == bv->ci[0] || j == bv->ci[1], PETSC_ERR_ARG_WRONG, "Column %d has not been fetched with a call to BVGetColumn", j);
  const int l = (j == bv->ci[0]) ? 0 : 1;
  PetscCheck(*v == bv->cv[l], PETSC_ERR_ARG_WRONG, "Argument 3 is not the same Vec that was obtained with BVGetColumn");
  PetscCall(bv->ops.restorecolumn(bv, j, v));
  bv->ci[l] = -bv->nc - 1;
  bv->state++;                                    /* the column may have been written through the Vec */
  *v = NULL;
  return PETSC_SUCCESS;
}

PetscErrorCode BVGetArray(BV bv, PetscScalar **a)
{
  BVCheckSizes(bv);
  BVCheckOp(bv, getarray);
  PetscCall(bv->ops.getarray(bv, a));
  return PETSC_SUCCESS;
}
PetscErrorCode BVRestoreArray(BV bv, PetscScalar **a)
{
  BVCheckSizes(bv);
  if (bv->ops.restorearray) PetscCall(bv->ops.restorearray(bv, a));
  if (a) *a = NULL;
  bv->state++;
  return PETSC_SUCCESS;
}

PetscErrorCode BVCreateVec(BV bv, Vec *v)
{
  BVCheckSizes(bv);
  if (bv->mem == B2K_MEM_DEVICE) PetscCall(VecCreateB200(bv->n, bv->N, v));
  else PetscCall(VecCreateHost(bv->n, bv->N, v));
  (*v)->rstart = bv->row0;
  return PETSC_SUCCESS;
}

PetscErrorCode BVSetColumnHost(BV bv, PetscInt j, const PetscScalar *hv)
{
  Vec v;
  PetscCall(BVGetColumn(bv, j, &v));
  PetscErrorCode ierr = VecSetValuesHost(v, hv);
  PetscCall(BVRestoreColumn(bv, j, &v));
  PetscCall(ierr);
  return PETSC_SUCCESS;
}
PetscErrorCode BVGetColumnHost(BV bv, PetscInt j, PetscScalar *hv)
{
  Vec v;
  PetscCall(BVGetColumn(bv, j, &v));
  PetscErrorCode ierr = VecGetValuesHost(v, hv);
  PetscCall(BVRestoreColumn(bv, j, &v));
  PetscCall(ierr);
  return PETSC_SUCCESS;
}

PetscErrorCode BVInsertVec(BV bv, PetscInt j, Vec w)
{
  BVCheckSizes(bv);
  PetscCheck(w->n == bv->n, PETSC_ERR_ARG_INCOMP, "Vec argument has local dimension %d, should be %d", w->n, bv->n);
  PetscCheck(j >= -bv->nc && j < bv->m, PETSC_ERR_ARG_OUTOFRANGE, "Argument j has wrong value %d, should be between %d and %d", j, -bv->nc, bv->m - 1);
  Vec v;
  PetscCall(BVGetColumn(bv, j, &v));
  PetscErrorCode ierr = VecCopy(w, v);
  PetscCall(BVRestoreColumn(bv, j, &v));
  PetscCall(ierr);
  return PETSC_SUCCESS;
}
PetscErrorCode BVCopyVec(BV bv, PetscInt j, Vec w)
{
  BVCheckSizes(bv);
  PetscCheck(w->n == bv->n, PETSC_ERR_ARG_INCOMP, "Vec argument has local dimension %d, should be %d", w->n, bv->n);
  Vec v;
  PetscCall(BVGetColumn(bv, j, &v));
  PetscErrorCode ierr = VecCopy(v, w);
  PetscCall(BVRestoreColumn(bv, j, &v));
  PetscCall(ierr);
  return PETSC_SUCCESS;
}

PetscErrorCode BVCopy(BV V, BV W)
{
  BVCheckSizes(V); BVCheckSizes(W);
  BVCheckOp(V, copy);
  PetscCheck(V->n == W->n, PETSC_ERR_ARG_INCOMP, "Mismatching local dimension V %d, W %d", V->n, W->n);
  PetscCheck(V->k - V->l <= W->m - W->l, PETSC_ERR_ARG_SIZ, "W has %d non-leading columns, not enough to store %d columns", W->m - W->l, V->k - V->l);
  PetscCheck(!strcmp(V->type_name, W->type_name), PETSC_ERR_ARG_INCOMP, "V and W must have the same type");
  if (V == W || !V->n) return PETSC_SUCCESS;
  PetscCall(V->ops.copy(V, W));
  W->state++;
  return PETSC_SUCCESS;
}

PetscErrorCode BVCopyColumn(BV V, PetscInt j, PetscInt i)
{
  BVCheckSizes(V);
  PetscCheck(j >= -V->nc && j < V->m && i >= -V->nc && i < V->m, PETSC_ERR_ARG_OUTOFRANGE, "column index out of range");
  if (j == i) return PETSC_SUCCESS;
  BVCheckOp(V, copycolumn);
  PetscCall(V->ops.copycolumn(V, j, i));
  V->state++;
  return PETSC_SUCCESS;
}

PetscErrorCode BVSetRandomSeed(BV bv, uint64_t seed) { bv->rng_seed = seed; return PETSC_SUCCESS; }

PetscErrorCode BVSetRandomColumn(BV bv, PetscInt j)
{
  BVCheckSizes(bv);
  PetscCheck(j >= 0 && j < bv->m, PETSC_ERR_ARG_OUTOFRANGE, "Argument j has wrong value %d, should be between 0 and %d", j, bv->m - 1);
  BVCheckOp(bv, setrandomcolumn);
  PetscCall(bv->ops.setrandomcolumn(bv, j));
  bv->state++;
  return PETSC_SUCCESS;
}

/* ---- non-standard inner product: bvbasic.c:497-600, bvimpl.h:147-175, bvglobal.c:20-46 ------------------------------- */
PetscErrorCode BVSetMatrix(BV bv, Mat B, PetscBool indef)
{
  if (B) {
    PetscCheck(B->M == B->N, PETSC_ERR_ARG_SIZ, "Matrix argument is not square, it has %d rows and %d columns", B->M, B->N);
    PetscCheck(!indef, PETSC_ERR_SUP, "indefinite inner products (GHIEP) are outside this path");
    if (bv->sizes_set) PetscCheck(B->n == bv->n, PETSC_ERR_ARG_INCOMP, "Mismatching local dimension BV %d, Mat %d", bv->n, B->n);
  }
  if (B != bv->matrix || (B && indef != bv->indef)) {
    bv->matrix = B;
    bv->indef = B ? indef : PETSC_FALSE;
    PetscCall(VecDestroy(&bv->Bx));
    PetscCall(BVDestroy(&bv->cached));
  }
  return PETSC_SUCCESS;
}
PetscErrorCode BVGetMatrix(BV bv, Mat *B, PetscBool *indef) { if (B) *B = bv->matrix; if (indef) *indef = bv->indef; return PETSC_SUCCESS; }

/* bv->Bx = B x (BV_IPMatMult; recomputed on every call: Vec objects carry no state counter here) */
static PetscErrorCode BV_IPMatMult(BV bv, Vec x)
{
  if (!bv->Bx) PetscCall(MatCreateVecs(bv->matrix, &bv->Bx, NULL));
  PetscCall(MatMult(bv->matrix, x, bv->Bx));
  return PETSC_SUCCESS;
}
/* bv->cached(:, l:k) = B bv(:, l:k) (BV_IPMatMultBV) */
static PetscErrorCode BV_IPMatMultBV(BV bv)
{
  if (!bv->cached) {
    Mat B = bv->matrix;
    bv->matrix = NULL;                            /* the cache itself uses the standard inner product */
    PetscErrorCode ierr = BVDuplicate(bv, &bv->cached);
    bv->matrix = B;
    PetscCall(ierr);
  }
  PetscCall(BVSetActiveColumns(bv->cached, bv->l, bv->k));
  PetscCall(BVMatMult(bv, bv->matrix, bv->cached));
  return PETSC_SUCCESS;
}
/* sqrt(z^T B z): BVNorm_Private bvglobal.c:417-438 */
static PetscErrorCode BVNorm_IP_Private(BV bv, Vec z, PetscReal *val)
{
  PetscScalar p;
  PetscCall(BV_IPMatMult(bv, z));
  PetscCall(VecDot(bv->Bx, z, &p));
  const PetscReal deftol = 10 * PETSC_MACHINE_EPSILON;
  PetscCheck(p > -deftol * fabs(p) - deftol, PETSC_ERR_FP, "The inner product is not well defined: indefinite matrix %g", p);
  *val = (p < 0.0) ? 0.0 : sqrt(p);
  return PETSC_SUCCESS;
}

/* ---- level-2/3 operations: bvops.c ------------------------------------------------------------------ */
PetscErrorCode BVMult(BV Y, PetscScalar alpha, PetscScalar beta, BV X, Mat Q)
{
  BVCheckSizes(Y); BVCheckSizes(X);
  BVCheckOp(Y, mult);
  PetscCheck(!strcmp(X->type_name, Y->type_name), PETSC_ERR_ARG_INCOMP, "X and Y must have the same type");
  PetscCheck(X != Y, PETSC_ERR_ARG_WRONG, "X and Y arguments must be different");
  if (Q) {
    PetscCheck(Q->dense, PETSC_ERR_SUP, "Mat argument must be sequential dense");
    PetscCheck(Q->m >= X->k, PETSC_ERR_ARG_SIZ, "Mat argument has %d rows, should have at least %d", Q->m, X->k);
    PetscCheck(Q->n >= Y->k, PETSC_ERR_ARG_SIZ, "Mat argument has %d columns, should have at least %d", Q->n, Y->k);
  }
  PetscCheck(X->n == Y->n, PETSC_ERR_ARG_INCOMP, "Mismatching local dimension X %d, Y %d", X->n, Y->n);
  PetscCall(Y->ops.mult(Y, alpha, beta, X, Q));
  Y->state++;
  return PETSC_SUCCESS;
}

PetscErrorCode BVMultVec(BV X, PetscScalar alpha, PetscScalar beta, Vec y, PetscScalar q[])
{
  BVCheckSizes(X);
  BVCheckOp(X, multvec);
  PetscCheck(y, PETSC_ERR_ARG_NULL, "null vector");
  PetscCheck(X->n == y->n, PETSC_ERR_ARG_INCOMP, "Mismatching local dimension X %d, y %d", X->n, y->n);
  PetscCall(X->ops.multvec(X, alpha, beta, y, q));
  return PETSC_SUCCESS;
}

PetscErrorCode BVMultColumn(BV X, PetscScalar alpha, PetscScalar beta, PetscInt j, PetscScalar *q)
{
  BVCheckSizes(X);
  PetscCheck(j >= 0, PETSC_ERR_ARG_OUTOFRANGE, "Index j must be non-negative");
  PetscCheck(j < X->m, PETSC_ERR_ARG_OUTOFRANGE, "Index j=%d but BV only has %d columns", j, X->m);
  const PetscInt ksave = X->k;
  X->k = j;
  Vec y;
  PetscCall(BVGetColumn(X, j, &y));
  PetscErrorCode ierr = X->ops.multvec(X, alpha, beta, y, q);
  PetscCall(BVRestoreColumn(X, j, &y));
  X->k = ksave;
  PetscCall(ierr);
  return PETSC_SUCCESS;
}

PetscErrorCode BVMultInPlace(BV V, Mat Q, PetscInt s, PetscInt e)
{
  BVCheckSizes(V);
  PetscCheck(Q && Q->dense, PETSC_ERR_SUP, "Mat argument must be sequential dense");
  PetscCheck(s >= V->l && s <= V->m, PETSC_ERR_ARG_OUTOFRANGE, "Argument s has wrong value %d, should be between %d and %d", s, V->l, V->m);
  PetscCheck(e >= V->l && e <= V->m, PETSC_ERR_ARG_OUTOFRANGE, "Argument e has wrong value %d, should be between %d and %d", e, V->l, V->m);
  PetscCheck(Q->m >= V->k, PETSC_ERR_ARG_SIZ, "Mat argument has %d rows, should have at least %d", Q->m, V->k);
  PetscCheck(e <= Q->n, PETSC_ERR_ARG_SIZ, "Mat argument only has %d columns, the requested value of e is larger: %d", Q->n, e);
  BVCheckOp(V, multinplace);
  PetscCall(V->ops.multinplace(V, Q, s, e));
  V->state++;
  return PETSC_SUCCESS;
}

PetscErrorCode BVMultInPlaceHermitianTranspose(BV V, Mat Q, PetscInt s, PetscInt e)
{
  BVCheckSizes(V);
  PetscCheck(Q && Q->dense, PETSC_ERR_SUP, "Mat argument must be sequential dense");
  PetscCheck(s >= V->l && s <= V->m, PETSC_ERR_ARG_OUTOFRANGE, "Argument s has wrong value %d, should be between %d and %d", s, V->l, V->m);
  PetscCheck(e >= V->l && e <= V->m, PETSC_ERR_ARG_OUTOFRANGE, "Argument e has wrong value %d, should be between %d and %d", e, V->l, V->m);
  PetscCheck(Q->n >= V->k, PETSC_ERR_ARG_SIZ, "Mat argument has %d columns, should have at least %d", Q->n, V->k);
  PetscCheck(e <= Q->m, PETSC_ERR_ARG_SIZ, "Mat argument only has %d rows, the requested value of e is larger: %d", Q->m, e);
  BVCheckOp(V, multinplacetrans);
  PetscCall(V->ops.multinplacetrans(V, Q, s, e));
  V->state++;
  return PETSC_SUCCESS;
}

PetscErrorCode BVScale(BV bv, PetscScalar alpha)
{
  BVCheckSizes(bv);
  if (alpha == 1.0) return PETSC_SUCCESS;
  BVCheckOp(bv, scale);
  PetscCall(bv->ops.scale(bv, -1, alpha));
  bv->state++;
  return PETSC_SUCCESS;
}

PetscErrorCode BVScaleColumn(BV bv, PetscInt j, PetscScalar alpha)
{
  BVCheckSizes(bv);
  PetscCheck(j >= 0 && j < bv->m, PETSC_ERR_ARG_OUTOFRANGE, "Argument j has wrong value %d, the number of columns is %d", j, bv->m);
  if (alpha == 1.0) return PETSC_SUCCESS;
  BVCheckOp(bv, scale);
  PetscCall(bv->ops.scale(bv, j, alpha));
  bv->state++;
  return PETSC_SUCCESS;
}

/* ---- reductions: bvglobal.c --------------------------------------------------------------------------- */
PetscErrorCode BVDot(BV X, BV Y, Mat M)
{
  BVCheckSizes(X); BVCheckSizes(Y);
  BVCheckOp(X, dot);
  PetscCheck(M && M->dense, PETSC_ERR_SUP, "Mat argument must be sequential dense");
  PetscCheck(!strcmp(X->type_name, Y->type_name), PETSC_ERR_ARG_INCOMP, "X and Y must have the same type");
  PetscCheck(M->m >= Y->k, PETSC_ERR_ARG_SIZ, "Mat argument has %d rows, should have at least %d", M->m, Y->k);
  PetscCheck(M->n >= X->k, PETSC_ERR_ARG_SIZ, "Mat argument has %d columns, should have at least %d", M->n, X->k);
  PetscCheck(X->n == Y->n, PETSC_ERR_ARG_INCOMP, "Mismatching local dimension X %d, Y %d", X->n, Y->n);
  if (X->l == X->k || Y->l == Y->k) return PETSC_SUCCESS;
  if (X->matrix) {                                /* M = Y^T (B X): bvglobal.c:84-91 */
    PetscCall(BV_IPMatMultBV(X));
    PetscCall(X->ops.dot(X->cached, Y, M));
  } else PetscCall(X->ops.dot(X, Y, M));
  return PETSC_SUCCESS;
}

PetscErrorCode BVDotVec(BV X, Vec y, PetscScalar m[])
{
  BVCheckSizes(X);
  BVCheckOp(X, dotvec);
  PetscCheck(y, PETSC_ERR_ARG_NULL, "null vector");
  PetscCheck(X->n == y->n, PETSC_ERR_ARG_INCOMP, "Mismatching local dimension X %d, y %d", X->n, y->n);
  if (X->matrix) { PetscCall(BV_IPMatMult(X, y)); PetscCall(X->ops.dotvec(X, X->Bx, m)); }   /* svec.c:117-120 */
  else PetscCall(X->ops.dotvec(X, y, m));
  return PETSC_SUCCESS;
}

PetscErrorCode BVDotColumn(BV X, PetscInt j, PetscScalar *q)
{
  BVCheckSizes(X);
  BVCheckOp(X, dotvec);
  PetscCheck(j >= 0, PETSC_ERR_ARG_OUTOFRANGE, "Index j must be non-negative");
  PetscCheck(j < X->m, PETSC_ERR_ARG_OUTOFRANGE, "Index j=%d but BV only has %d columns", j, X->m);
  const PetscInt ksave = X->k;
  X->k = j;
  Vec y;
  PetscCall(BVGetColumn(X, j, &y));
  PetscErrorCode ierr = PETSC_SUCCESS;
  if (X->matrix) { ierr = BV_IPMatMult(X, y); if (!ierr) ierr = X->ops.dotvec(X, X->Bx, q); }
  else ierr = X->ops.dotvec(X, y, q);
  PetscCall(BVRestoreColumn(X, j, &y));
  X->k = ksave;
  PetscCall(ierr);
  return PETSC_SUCCESS;
}

/* ---- split-phase reductions: bvglobal.c:188-260, 350-440, 703-800 ------------------------------------------------
   PetscSplitReduction stand-in for BV types WITHOUT the *_begin/_end slots: local parts (dotvec_local / norm_local) are
   queued in host memory, the first End reduces all of them in one collective per reduction type. */
#define B2K_SR_MAX 4096
static struct {
  double lv[B2K_SR_MAX], gv[B2K_SR_MAX];
  char   ismax[B2K_SR_MAX];
  void  *owner[B2K_SR_MAX];
  int    nbegin, nend, reduced;
} g_sr;

static PetscErrorCode BVSplitReductionEnd_Private(BV bv)
{
  if (g_sr.reduced) return PETSC_SUCCESS;
  int nmax = 0;
  for (int i = 0; i < g_sr.nbegin; i++) nmax += g_sr.ismax[i];
  if (!nmax) {
    memcpy(g_sr.gv, g_sr.lv, sizeof(double) * (size_t)g_sr.nbegin);
    PetscCall(B2KCommAllreduce(bv->comm, g_sr.gv, g_sr.nbegin, 0, B2K_MEM_HOST));
  } else {                                        /* mixed SUM / MAX: one collective per type */
    static double tmp[B2K_SR_MAX];
    for (int op = 0; op < 2; op++) {
      int n = 0;
      for (int i = 0; i < g_sr.nbegin; i++) if (g_sr.ismax[i] == op) tmp[n++] = g_sr.lv[i];
      if (n) PetscCall(B2KCommAllreduce(bv->comm, tmp, n, op, B2K_MEM_HOST));
      n = 0;
      for (int i = 0; i < g_sr.nbegin; i++) if (g_sr.ismax[i] == op) g_sr.gv[i] = tmp[n++];
    }
  }
  g_sr.reduced = 1;
  return PETSC_SUCCESS;
}
static void BVSplitReductionPop_Private(int n)
{
  g_sr.nend += n;
  if (g_sr.nend == g_sr.nbegin) g_sr.nbegin = g_sr.nend = g_sr.reduced = 0;   /* all results handed out: ready for the next batch */
}

PetscErrorCode BVDotVecBegin(BV X, Vec y, PetscScalar *m)
{
  BVCheckSizes(X);
  PetscCheck(y, PETSC_ERR_ARG_NULL, "null vector");
  PetscCheck(X->n == y->n, PETSC_ERR_ARG_INCOMP, "Mismatching local dimension X %d, y %d", X->n, y->n);
  PetscCheck(!X->matrix, PETSC_ERR_SUP, "split-phase reductions with a non-standard inner product are not available");
  if (X->ops.dotvec_begin) { PetscCall(X->ops.dotvec_begin(X, y, m)); return PETSC_SUCCESS; }
  BVCheckOp(X, dotvec_local);
  const PetscInt nv = X->k - X->l;
  PetscCheck(!g_sr.reduced, PETSC_ERR_ORDER, "Called before all BVxxxEnd() called");
  PetscCheck(g_sr.nbegin + nv <= B2K_SR_MAX, PETSC_ERR_ARG_SIZ, "too many outstanding split reductions");
  for (PetscInt i = 0; i < nv; i++) { g_sr.ismax[g_sr.nbegin + i] = 0; g_sr.owner[g_sr.nbegin + i] = (void *)X; }
  PetscCall(X->ops.dotvec_local(X, y, g_sr.lv + g_sr.nbegin));
  g_sr.nbegin += nv;
  return PETSC_SUCCESS;
}

PetscErrorCode BVDotVecEnd(BV X, Vec y, PetscScalar *m)
{
  BVCheckSizes(X);
  if (X->ops.dotvec_end) { PetscCall(X->ops.dotvec_end(X, y, m)); return PETSC_SUCCESS; }
  const PetscInt nv = X->k - X->l;
  PetscCall(BVSplitReductionEnd_Private(X));
  PetscCheck(g_sr.nend + nv <= g_sr.nbegin, PETSC_ERR_ARG_WRONGSTATE, "Called BVxxxEnd() more times than BVxxxBegin()");
  PetscCheck(nv == 0 || (void *)X == g_sr.owner[g_sr.nend], PETSC_ERR_ARG_WRONGSTATE, "Called BVxxxEnd() in a different order or with a different BV than BVxxxBegin()");
  PetscScalar *mm = m ? m : X->buffer;
  for (PetscInt i = 0; i < nv; i++) mm[i] = g_sr.gv[g_sr.nend + i];
  BVSplitReductionPop_Private(nv);
  return PETSC_SUCCESS;
}

PetscErrorCode BVDotColumnBegin(BV X, PetscInt j, PetscScalar *m)
{
  BVCheckSizes(X);
  PetscCheck(j >= 0, PETSC_ERR_ARG_OUTOFRANGE, "Index j must be non-negative");
  PetscCheck(j < X->m, PETSC_ERR_ARG_OUTOFRANGE, "Index j=%d but BV only has %d columns", j, X->m);
  const PetscInt ksave = X->k;
  X->k = j;
  Vec y;
  PetscCall(BVGetColumn(X, j, &y));
  PetscErrorCode ierr = BVDotVecBegin(X, y, m);
  PetscCall(BVRestoreColumn(X, j, &y));
  X->k = ksave;
  PetscCall(ierr);
  return PETSC_SUCCESS;
}

PetscErrorCode BVDotColumnEnd(BV X, PetscInt j, PetscScalar *m)
{
  BVCheckSizes(X);
  PetscCheck(j >= 0 && j < X->m, PETSC_ERR_ARG_OUTOFRANGE, "Index j=%d out of range", j);
  const PetscInt ksave = X->k;
  X->k = j;
  PetscErrorCode ierr = BVDotVecEnd(X, NULL, m);
  X->k = ksave;
  PetscCall(ierr);
  return PETSC_SUCCESS;
}

PetscErrorCode BVNormColumnBegin(BV bv, PetscInt j, NormType type, PetscReal *val)
{
  BVCheckSizes(bv);
  PetscCheck(j >= 0 && j < bv->m, PETSC_ERR_ARG_OUTOFRANGE, "Argument j has wrong value %d, the number of columns is %d", j, bv->m);
  PetscCheck(!bv->matrix, PETSC_ERR_SUP, "split-phase reductions with a non-standard inner product are not available");
  if (bv->ops.norm_begin) { PetscCall(bv->ops.norm_begin(bv, j, type, val)); return PETSC_SUCCESS; }
  BVCheckOp(bv, norm_local);
  PetscCheck(!g_sr.reduced, PETSC_ERR_ORDER, "Called before all BVxxxEnd() called");
  PetscCheck(g_sr.nbegin < B2K_SR_MAX, PETSC_ERR_ARG_SIZ, "too many outstanding split reductions");
  PetscReal lres;
  PetscCall(bv->ops.norm_local(bv, j, type, &lres));
  if (type == NORM_2 || type == NORM_FROBENIUS) lres = lres * lres;
  g_sr.ismax[g_sr.nbegin] = (type == NORM_INFINITY) ? 1 : 0;
  g_sr.owner[g_sr.nbegin] = (void *)bv;
  g_sr.lv[g_sr.nbegin++] = lres;
  return PETSC_SUCCESS;
}

PetscErrorCode BVNormColumnEnd(BV bv, PetscInt j, NormType type, PetscReal *val)
{
  BVCheckSizes(bv);
  if (bv->ops.norm_end) { PetscCall(bv->ops.norm_end(bv, j, type, val)); return PETSC_SUCCESS; }
  PetscCall(BVSplitReductionEnd_Private(bv));
  PetscCheck(g_sr.nend < g_sr.nbegin, PETSC_ERR_ARG_WRONGSTATE, "Called BVxxxEnd() more times than BVxxxBegin()");
  PetscCheck((void *)bv == g_sr.owner[g_sr.nend], PETSC_ERR_ARG_WRONGSTATE, "Called BVxxxEnd() in a different order or with a different BV than BVxxxBegin()");
  const PetscReal g = g_sr.gv[g_sr.nend];
  *val = (type == NORM_2 || type == NORM_FROBENIUS) ? sqrt(g) : g;
  BVSplitReductionPop_Private(1);
  return PETSC_SUCCESS;
}

/* BVNormVecBegin / BVNormVecEnd (bvglobal.c:575-636) are VecNormBegin / VecNormEnd for the standard inner product, i.e. PETSc's
   split-phase Vec reductions, and bv/tests/test10.c pairs them with VecDotBegin / VecDotEnd.  PETSc is not here, so the four Vec
   entry points are stand-ins that evaluate at Begin (the complete, reduced value) and hand the number out at End in FIFO order —
   what PETSc itself does when split reductions are switched off; the BV-level pairs above are the ones that really merge. */
#define B2K_EAGER_MAX 64
static struct { double v[B2K_EAGER_MAX]; void *owner[B2K_EAGER_MAX]; int head, tail; } g_eager;
static PetscErrorCode EagerPush_Private(void *owner, double v)
{
  PetscCheck(g_eager.tail - g_eager.head < B2K_EAGER_MAX, PETSC_ERR_ARG_SIZ, "too many outstanding split reductions");
  g_eager.owner[g_eager.tail % B2K_EAGER_MAX] = owner; g_eager.v[g_eager.tail++ % B2K_EAGER_MAX] = v;
  return PETSC_SUCCESS;
}
static PetscErrorCode EagerPop_Private(void *owner, double *v)
{
  PetscCheck(g_eager.head < g_eager.tail, PETSC_ERR_ARG_WRONGSTATE, "Called xxxEnd() more times than xxxBegin()");
  PetscCheck(g_eager.owner[g_eager.head % B2K_EAGER_MAX] == owner, PETSC_ERR_ARG_WRONGSTATE, "Called xxxEnd() in a different order or on a different object than xxxBegin()");
  *v = g_eager.v[g_eager.head++ % B2K_EAGER_MAX];
  return PETSC_SUCCESS;
}
PetscErrorCode VecDotBegin(Vec x, Vec y, PetscScalar *val) { PetscScalar d; (void)val; PetscCall(VecDot(x, y, &d)); return EagerPush_Private((void *)x, d); }
PetscErrorCode VecDotEnd(Vec x, Vec y, PetscScalar *val) { (void)y; return EagerPop_Private((void *)x, val); }
PetscErrorCode VecNormBegin(Vec x, NormType type, PetscReal *val) { PetscReal d; (void)val; PetscCall(VecNorm(x, type, &d)); return EagerPush_Private((void *)x, d); }
PetscErrorCode VecNormEnd(Vec x, NormType type, PetscReal *val) { (void)type; return EagerPop_Private((void *)x, val); }

PetscErrorCode BVNormVecBegin(BV bv, Vec v, NormType type, PetscReal *val)
{
  BVCheckSizes(bv);
  PetscCheck(v, PETSC_ERR_ARG_NULL, "null vector");
  if (bv->matrix) {                                /* BVNorm_Begin_Private, bvglobal.c:597-601: sqrt(v' B v), evaluated here */
    PetscReal d;
    (void)val;
    PetscCall(BVNormVec(bv, v, type, &d));
    PetscCall(EagerPush_Private((void *)v, d));
  } else PetscCall(VecNormBegin(v, type, val));    /* bvglobal.c:602 */
  return PETSC_SUCCESS;
}
PetscErrorCode BVNormVecEnd(BV bv, Vec v, NormType type, PetscReal *val)
{
  BVCheckSizes(bv);
  if (bv->matrix) PetscCall(EagerPop_Private((void *)v, val));
  else PetscCall(VecNormEnd(v, type, val));        /* bvglobal.c:634 */
  return PETSC_SUCCESS;
}

PetscErrorCode BVNorm(BV bv, NormType type, PetscReal *val)
{
  BVCheckSizes(bv);
  BVCheckOp(bv, norm);
  PetscCheck(type != NORM_2 || bv->k - bv->l <= 1, PETSC_ERR_SUP, "Requested norm not available");   /* bvglobal.c:496 */
  PetscCheck(!bv->matrix, PETSC_ERR_SUP, "Matrix norm not available for non-standard inner product");   /* bvglobal.c:497 */
  PetscCall(bv->ops.norm(bv, -1, type, val));
  return PETSC_SUCCESS;
}

PetscErrorCode BVNormColumn(BV bv, PetscInt j, NormType type, PetscReal *val)
{
  BVCheckSizes(bv);
  BVCheckOp(bv, norm);
  PetscCheck(j >= 0 && j < bv->m, PETSC_ERR_ARG_OUTOFRANGE, "Argument j has wrong value %d, the number of columns is %d", j, bv->m);
  if (bv->matrix) {                               /* sqrt(V[j]' B V[j]), the type is ignored: bvglobal.c:547-551 */
    Vec z;
    PetscCall(BVGetColumn(bv, j, &z));
    PetscErrorCode ierr = BVNorm_IP_Private(bv, z, val);
    PetscCall(BVRestoreColumn(bv, j, &z));
    PetscCall(ierr);
  } else PetscCall(bv->ops.norm(bv, j, type, val));
  return PETSC_SUCCESS;
}

PetscErrorCode BVNormVec(BV bv, Vec v, NormType type, PetscReal *val)
{
  BVCheckSizes(bv);
  PetscCheck(v->n == bv->n, PETSC_ERR_ARG_INCOMP, "Vec argument has local dimension %d, should be %d", v->n, bv->n);
  if (bv->matrix) PetscCall(BVNorm_IP_Private(bv, v, val));   /* bvglobal.c:607-609 */
  else PetscCall(VecNorm(v, type, val));
  return PETSC_SUCCESS;
}

PetscErrorCode BVNormalize(BV bv, PetscScalar *eigi)
{
  BVCheckSizes(bv);
  if (bv->matrix) {                               /* BVNormalize_Private bvglobal.c:800-834: B-norms, one column at a time */
    PetscCheck(!eigi, PETSC_ERR_SUP, "conjugate pairs with a non-standard inner product are outside this path");
    BVCheckOp(bv, scale);
    for (PetscInt i = bv->l; i < bv->k; i++) {
      PetscReal nrm;
      PetscCall(BVNormColumn(bv, i, NORM_2, &nrm));
      if (nrm != 0.0 && nrm != 1.0) PetscCall(bv->ops.scale(bv, i, 1.0 / nrm));
    }
    bv->state++;
    return PETSC_SUCCESS;
  }
  BVCheckOp(bv, normalize);
  PetscCall(bv->ops.normalize(bv, eigi));
  bv->state++;
  return PETSC_SUCCESS;
}

/* ---- operator application: bvops.c:767-885 ------------------------------------------------------------ */
PetscErrorCode BVMatMult(BV V, Mat A, BV Y)
{
  BVCheckSizes(V); BVCheckSizes(Y);
  BVCheckOp(V, matmult);
  PetscCheck(V->n == A->n, PETSC_ERR_ARG_INCOMP, "Mat argument has %d local columns, should be %d", A->n, V->n);
  PetscCheck(V->k - V->l <= Y->m - Y->l, PETSC_ERR_ARG_SIZ, "Y has %d non-leading columns, not enough to store %d columns", Y->m - Y->l, V->k - V->l);
  PetscCheck(A->m == Y->n, PETSC_ERR_ARG_INCOMP, "Mismatching local dimension A %d, Y %d", A->m, Y->n);
  PetscCall(V->ops.matmult(V, A, Y));
  V->n_matmult += V->k - V->l;
  Y->state++;
  return PETSC_SUCCESS;
}

PetscErrorCode BVMatMultColumn(BV V, Mat A, PetscInt j)
{
  BVCheckSizes(V);
  PetscCheck(j >= 0, PETSC_ERR_ARG_OUTOFRANGE, "Index j must be non-negative");
  PetscCheck(j + 1 < V->m, PETSC_ERR_ARG_OUTOFRANGE, "Result should go in index j+1=%d but BV only has %d columns", j + 1, V->m);
  Vec vj, vj1;
  PetscCall(BVGetColumn(V, j, &vj));
  PetscCall(BVGetColumn(V, j + 1, &vj1));
  PetscErrorCode ierr = MatMult(A, vj, vj1);
  PetscCall(BVRestoreColumn(V, j, &vj));
  PetscCall(BVRestoreColumn(V, j + 1, &vj1));
  PetscCall(ierr);
  V->n_matmult++;
  return PETSC_SUCCESS;
}

/* ---- Gram-Schmidt: bvorthog.c --------------------------------------------------------------------------- */
/* BV_SafeSqrt bvimpl.h:121-141 (definite inner product) */
static PetscErrorCode BV_SafeSqrt(PetscScalar alpha, PetscReal *res)
{
  const PetscReal deftol = 10 * PETSC_MACHINE_EPSILON;
  PetscCheck(alpha > -deftol, PETSC_ERR_FP, "The inner product is not well defined: indefinite matrix %g", alpha);   /* also catches NaN */
  *res = (alpha < 0.0) ? 0.0 : sqrt(alpha);
  return PETSC_SUCCESS;
}

static PetscErrorCode BV_NormVecOrColumn(BV bv, PetscInt j, Vec v, PetscReal *nrm)
{
  if (v) PetscCall(BVNormVec(bv, v, NORM_2, nrm));
  else PetscCall(BVNormColumn(bv, j, NORM_2, nrm));
  return PETSC_SUCCESS;
}

/* one classical GS step with a single reduction — bvorthog.c:91-132 */
static PetscErrorCode BVOrthogonalizeCGS1(BV bv, PetscInt j, Vec v, PetscBool *which, PetscScalar *h, PetscScalar *c, PetscReal *onorm, PetscReal *norm)
{
  (void)which;
  PetscReal beta = 0.0, sum = 0.0;
  PetscScalar *cc = c ? c : bv->buffer;
  bv->k = j;
  if (onorm || norm) {
    if (!v) {                                     /* BVDotColumnInc :32-47: k = j+1, column j times itself included */
      Vec y;
      bv->k = j + 1;
      PetscCall(BVGetColumn(bv, j, &y));
      PetscErrorCode ierr = PETSC_SUCCESS;
      if (bv->matrix) { ierr = BV_IPMatMult(bv, y); if (!ierr) ierr = bv->ops.dotvec(bv, bv->Bx, c); }   /* svec.c:117-120 */
      else ierr = bv->ops.dotvec(bv, y, c);
      PetscCall(BVRestoreColumn(bv, j, &y));
      bv->k = j;
      PetscCall(ierr);
      PetscCall(BV_SafeSqrt(cc[bv->nc + j], &beta));           /* BV_SquareRoot */
    } else {
      PetscCall(BVDotVec(bv, v, c));
      PetscCall(BVNormVec(bv, v, NORM_2, &beta));
    }
  } else {
    if (!v) PetscCall(BVDotColumn(bv, j, c));
    else PetscCall(BVDotVec(bv, v, c));
  }
  if (!v) PetscCall(BVMultColumn(bv, -1.0, 1.0, j, c));
  else PetscCall(BVMultVec(bv, -1.0, 1.0, v, c));
  if (onorm) *onorm = beta;
  if (norm) {
    for (PetscInt i = 0; i < bv->nc + j; i++) sum += cc[i] * cc[i];   /* BV_SquareSum */
    *norm = beta * beta - sum;
    if (*norm <= 0.0) PetscCall(BV_NormVecOrColumn(bv, j, v, norm));
    else *norm = sqrt(*norm);
  }
  BV_AddCoefficients(bv, j, h, c);
  return PETSC_SUCCESS;
}

/* one modified GS step — bvorthog.c:52-85 */
static PetscErrorCode BVOrthogonalizeMGS1(BV bv, PetscInt j, Vec v, PetscBool *which, PetscScalar *h, PetscScalar *c, PetscReal *onrm, PetscReal *nrm)
{
  Vec w = v, vi;
  PetscScalar *cc = c ? c : bv->buffer;
  PetscScalar dot;
  if (!v) PetscCall(BVGetColumn(bv, j, &w));
  if (onrm) PetscCall(BVNormVec(bv, w, NORM_2, onrm));
  for (PetscInt i = -bv->nc; i < j; i++) {
    if (which && i >= 0 && !which[i]) continue;
    PetscCall(BVGetColumn(bv, i, &vi));
    if (bv->matrix) { PetscCall(BV_IPMatMult(bv, w)); PetscCall(VecDot(bv->Bx, vi, &dot)); }   /* bvorthog.c:69-73 */
    else PetscCall(VecDot(w, vi, &dot));
    cc[bv->nc + i] = dot;
    PetscCall(VecAXPY(w, -dot, vi));
    PetscCall(BVRestoreColumn(bv, i, &vi));
  }
  if (nrm) PetscCall(BVNormVec(bv, w, NORM_2, nrm));
  if (!v) PetscCall(BVRestoreColumn(bv, j, &w));
  BV_AddCoefficients(bv, j, h, c);
  return PETSC_SUCCESS;
}

/* DGKS driver — bvorthog.c:145-217 */
static PetscErrorCode BVOrthogonalizeGS(BV bv, PetscInt j, Vec v, PetscBool *which, PetscReal *norm, PetscBool *lindep)
{
  PetscScalar *h, *c;
  PetscReal onrm = 0.0, nrm = 0.0;
  PetscInt k, l;
  const PetscBool mgs = (bv->orthog_type == BV_ORTHOG_MGS) ? PETSC_TRUE : PETSC_FALSE;
  PetscErrorCode (*gs1)(BV, PetscInt, Vec, PetscBool *, PetscScalar *, PetscScalar *, PetscReal *, PetscReal *) =
      (bv->ops.gramschmidt && !mgs && !bv->matrix) ? bv->ops.gramschmidt : (mgs ? BVOrthogonalizeMGS1 : BVOrthogonalizeCGS1);   /* bvorthog.c:134; the
      fused slot of type b200 implements the STANDARD inner product: with BVSetMatrix the generic pass (B x through MatMult) runs */
  if (v) { k = bv->k; h = bv->h; c = bv->c; }
  else { k = j; h = NULL; c = NULL; }
  const PetscBool dolindep = lindep ? PETSC_TRUE : PETSC_FALSE;

  /* BV_CleanCoefficients bvimpl.h:289 */
  {
    PetscScalar *hh = h ? h : bv->buffer + (size_t)k * (bv->nc + bv->m);
    for (PetscInt i = 0; i < bv->nc + k; i++) hh[i] = 0.0;
  }
  switch (bv->orthog_ref) {
  case BV_ORTHOG_REFINE_IFNEEDED:
    PetscCall(gs1(bv, k, v, which, h, c, &onrm, &nrm));
    bv->n_gs_passes++;
    l = 1;
    while (l < 3 && nrm && fabs(nrm) < bv->orthog_eta * fabs(onrm)) {   /* repeat if ||q|| < eta ||h||, :180 */
      l++;
      if (mgs) onrm = nrm;
      PetscCall(gs1(bv, k, v, which, h, c, mgs ? NULL : &onrm, &nrm));
      bv->n_gs_passes++;
    }
    if (dolindep) *lindep = !(nrm && fabs(nrm) >= bv->orthog_eta * fabs(onrm));
    break;
  case BV_ORTHOG_REFINE_NEVER:
    PetscCall(gs1(bv, k, v, which, h, c, NULL, NULL));
    bv->n_gs_passes++;
    if (norm || dolindep) PetscCall(BV_NormVecOrColumn(bv, k, v, &nrm));
    if (dolindep) *lindep = !nrm;
    break;
  case BV_ORTHOG_REFINE_ALWAYS:
    PetscCall(gs1(bv, k, v, which, h, c, NULL, NULL));
    PetscCall(gs1(bv, k, v, which, h, c, dolindep ? &onrm : NULL, (norm || dolindep) ? &nrm : NULL));
    bv->n_gs_passes += 2;
    if (dolindep) *lindep = !(nrm && fabs(nrm) >= bv->orthog_eta * fabs(onrm));
    break;
  }
  if (norm) {
    *norm = nrm;
    if (!v) BV_BUF(bv, bv->nc + k, k) = (dolindep && *lindep) ? 0.0 : nrm;   /* BV_SetValue(bv,k,k,h,nrm) :212-213 */
  }
  return PETSC_SUCCESS;
}

PetscErrorCode BVOrthogonalizeVec(BV bv, Vec v, PetscScalar *H, PetscReal *norm, PetscBool *lindep)
{
  BVCheckSizes(bv);
  PetscCheck(v, PETSC_ERR_ARG_NULL, "null vector");
  PetscCheck(v->n == bv->n, PETSC_ERR_ARG_INCOMP, "Vec argument has local dimension %d, should be %d", v->n, bv->n);
  const PetscInt ksave = bv->k, lsave = bv->l;
  bv->l = -bv->nc;
  PetscErrorCode ierr = BVOrthogonalizeGS(bv, 0, v, NULL, norm, lindep);
  bv->k = ksave; bv->l = lsave;
  PetscCall(ierr);
  if (H) for (PetscInt i = bv->l; i < bv->k; i++) H[i - bv->l] = bv->h[bv->nc + i];   /* BV_StoreCoefficients bvimpl.h:400-415 */
  return PETSC_SUCCESS;
}

PetscErrorCode BVOrthogonalizeColumn(BV bv, PetscInt j, PetscScalar *H, PetscReal *norm, PetscBool *lindep)
{
  BVCheckSizes(bv);
  PetscCheck(j >= 0, PETSC_ERR_ARG_OUTOFRANGE, "Index j must be non-negative");
  PetscCheck(j < bv->m, PETSC_ERR_ARG_OUTOFRANGE, "Index j=%d but BV only has %d columns", j, bv->m);
  const PetscInt ksave = bv->k, lsave = bv->l;
  bv->l = -bv->nc;
  PetscErrorCode ierr = BVOrthogonalizeGS(bv, j, NULL, NULL, norm, lindep);
  bv->k = ksave; bv->l = lsave;
  PetscCall(ierr);
  if (H) for (PetscInt i = bv->l; i < j + 1; i++) H[i - bv->l] = BV_BUF(bv, bv->nc + i, j);   /* j-l+1 coefficients, the last one is the norm */
  bv->state++;
  return PETSC_SUCCESS;
}

PetscErrorCode BVOrthonormalizeColumn(BV bv, PetscInt j, PetscBool replace, PetscReal *norm, PetscBool *lindep)
{
  BVCheckSizes(bv);
  PetscCheck(j >= 0, PETSC_ERR_ARG_OUTOFRANGE, "Index j must be non-negative");
  PetscCheck(j < bv->m, PETSC_ERR_ARG_OUTOFRANGE, "Index j=%d but BV only has %d columns", j, bv->m);
  PetscReal nrm = 0.0;
  PetscBool lndep = PETSC_FALSE;
  const PetscInt ksave = bv->k, lsave = bv->l;
  bv->l = -bv->nc;
  PetscErrorCode ierr = BVOrthogonalizeGS(bv, j, NULL, NULL, &nrm, &lndep);
  if (!ierr && replace && (nrm == 0.0 || lndep)) {
    for (int attempt = 0; attempt < 2 && !ierr && (nrm == 0.0 || lndep); attempt++) {   /* bvorthog.c:403-411 */
      ierr = bv->ops.setrandomcolumn ? bv->ops.setrandomcolumn(bv, j) : PETSC_ERR_SUP;
      bv->state++;
      if (!ierr) ierr = BVOrthogonalizeGS(bv, j, NULL, NULL, &nrm, &lndep);
    }
  }
  bv->k = ksave; bv->l = lsave;
  PetscCall(ierr);
  if (nrm != 1.0 && nrm != 0.0) {
    BVCheckOp(bv, scale);
    PetscCall(bv->ops.scale(bv, j, 1.0 / nrm));
  }
  if (norm) *norm = nrm;
  if (lindep) *lindep = lndep;
  bv->state++;
  return PETSC_SUCCESS;
}

/* BVOrthogonalize with BV_ORTHOG_BLOCK_GS — bvorthog.c:560-594, 682-740 */
/* ---- BVView: bvfunc.c:548-603 + the column loop of BVView_Svec svec.c:353-382 ------------------------------------------ */
PetscErrorCode B2KViewHeader_Private(const char *cls, const char *name, const char *type, int indent);
PetscErrorCode BVView(BV bv, PetscViewer viewer)
{
  static const char *orthname[2] = {"classical", "modified"}, *refname[3] = {"if needed", "never", "always"};
  static const char *blockname[5] = {"GS", "CHOL", "TSQR", "TSQRCHOL", "SVQB"};
  PetscViewerFormat format;
  BVCheckSizes(bv);
  PetscCall(PetscViewerGetFormat(viewer, &format));
  PetscCall(B2KViewHeader_Private("BV", bv->hdr.name, bv->type_name, 0));
  if (format == PETSC_VIEWER_ASCII_INFO || format == PETSC_VIEWER_ASCII_INFO_DETAIL) {
    printf("  %d columns of global length %d\n", bv->m, bv->N);
    if (bv->nc > 0) printf("  number of constraints: %d\n", bv->nc);
    printf("  vector orthogonalization method: %s Gram-Schmidt\n", orthname[bv->orthog_type]);
    if (bv->orthog_ref == BV_ORTHOG_REFINE_IFNEEDED) printf("  orthogonalization refinement: %s (eta: %g)\n", refname[bv->orthog_ref], bv->orthog_eta);
    else printf("  orthogonalization refinement: %s\n", refname[bv->orthog_ref]);
    printf("  block orthogonalization method: %s\n", blockname[bv->orthog_block]);
    printf("  doing matmult as matrix-vector products\n");          /* ops.matmult is a column loop of MatMult (BV_MATMULT_VECS) */
    return PETSC_SUCCESS;                                             /* the type's view prints nothing in the INFO formats, svec.c:365 */
  }
  for (PetscInt j = 0; j < bv->m; j++) {
    Vec v;
    PetscCall(BVGetColumn(bv, j, &v));
    PetscErrorCode ierr = VecView(v, viewer);
    PetscCall(BVRestoreColumn(bv, j, &v));
    PetscCall(ierr);
  }
  return PETSC_SUCCESS;
}

/* ---- block orthogonalisation: bvorthog.c:488-765, bvlapack.c:136-345 ------------------------------------------
   CHOL and SVQB are two BLAS-3 sweeps of the basis (BVDot = V^T V, BVMultInPlace = V S) around a k x k host
   factorisation, i.e. the level-3 kernels of the restart; TSQR/TSQRCHOL (host Householder panels over a raw
   BVGetArray pointer in the reference, bvlapack.c:347-560) go through the type's tsqr_factor / tsqr_formq. */

/* the buffer seen as an (nc+m) x m dense matrix, BV_GetBufferMat bvorthog.c:545-559 */
static PetscErrorCode BV_GetBufferMat(BV bv, Mat *R)
{
  PetscCall(MatCreateSeqDense(bv->nc + bv->m, bv->m, bv->buffer, R));
  return PETSC_SUCCESS;
}

/* BV_StoreCoeffsBlock_Default bvorthog.c:566-581 */
static PetscErrorCode BV_StoreCoeffsBlock(BV bv, Mat R, PetscBool tri)
{
  const PetscInt ldb = bv->m + bv->nc, ldr = R->lda;
  for (PetscInt j = bv->l; j < bv->k; j++)
    memcpy(R->dense + (size_t)j * ldr, bv->buffer + (size_t)j * ldb, sizeof(PetscScalar) * (size_t)((tri ? (j + 1) : bv->k) + bv->nc));
  return PETSC_SUCCESS;
}

/* Block Gram-Schmidt against the leading columns: V2 = V2 - V1*R12, R12 = V1'*V2 (bvorthog.c:488-499).  The split view V1
   of BVGetSplit is an alias of V restricted to columns 0..l-1 (same storage, svec.c:427-461). */
static PetscErrorCode BVOrthogonalize_BlockGS(BV V, Mat R)
{
  struct _p_BV V1 = *V;
  V1.l = 0; V1.k = V->l;
  V1.ci[0] = V1.ci[1] = -V->nc - 1;
  PetscCall(BVDot(V, &V1, R));
  PetscCall(BVMult(V, -1.0, 1.0, &V1, R));
  return PETSC_SUCCESS;
}

/* BVMatCholInv_LAPACK_Private bvlapack.c:136-202: upper Cholesky factor in R(l:k,l:k), its inverse in S */
static PetscErrorCode BVMatCholInv_Private(BV bv, Mat R, Mat S)
{
  const PetscInt l = bv->l, k = bv->k, n = k - l, ld = R->lda;
  int n_ = n, ld_ = ld, lds_, info = 0;
  PetscScalar *pR = R->dense, *pS;
  PetscInt lds;
  PetscScalar *tmp = NULL;
  if (S == R) {
    tmp = (PetscScalar *)calloc((size_t)ld * (size_t)k + 1, sizeof(PetscScalar));
    PetscCheck(tmp, PETSC_ERR_MEM, "out of memory");
    pS = tmp; lds = ld;
  } else { pS = S->dense; lds = S->lda; }
  lds_ = lds;
  for (PetscInt i = l; i < k; i++) memcpy(pS + (size_t)i * lds + l, pR + (size_t)i * ld + l, sizeof(PetscScalar) * (size_t)n);   /* save a copy */
  LAPACK(dpotrf)("U", &n_, pR + (size_t)l * ld + l, &ld_, &info, 1);
  if (info) {                                     /* retry on a diagonally perturbed matrix, bvlapack.c:172-180 */
    for (PetscInt i = l; i < k; i++) {
      memcpy(pR + (size_t)i * ld + l, pS + (size_t)i * lds + l, sizeof(PetscScalar) * (size_t)n);
      pR[i + (size_t)i * ld] += 50.0 * PETSC_MACHINE_EPSILON;
    }
    LAPACK(dpotrf)("U", &n_, pR + (size_t)l * ld + l, &ld_, &info, 1);
    if (info) { free(tmp); SETERRQ(PETSC_ERR_LIB, "Error in LAPACK subroutine potrf: info=%d", info); }
  }
  if (S == R) LAPACK(dtrtri)("U", "N", &n_, pR + (size_t)l * ld + l, &ld_, &info, 1, 1);
  else {
    memset(pS + (size_t)l * lds, 0, sizeof(PetscScalar) * (size_t)(k - l) * (size_t)lds);
    for (PetscInt i = l; i < k; i++) memcpy(pS + (size_t)i * lds + l, pR + (size_t)i * ld + l, sizeof(PetscScalar) * (size_t)n);
    LAPACK(dtrtri)("U", "N", &n_, pS + (size_t)l * lds + l, &lds_, &info, 1, 1);
  }
  free(tmp);
  PetscCheck(!info, PETSC_ERR_LIB, "Error in LAPACK subroutine trtri: info=%d", info);
  for (PetscInt i = l; i < k - 1; i++) {          /* zero out entries below the diagonal */
    memset(pR + (size_t)i * ld + i + 1, 0, sizeof(PetscScalar) * (size_t)(k - i - 1));
    if (S != R) memset(pS + (size_t)i * lds + i + 1, 0, sizeof(PetscScalar) * (size_t)(k - i - 1));
  }
  return PETSC_SUCCESS;
}

/* BVMatSVQB_LAPACK_Private bvlapack.c:261-345: on input R = V'V; on output S = D U Lambda^{-1/2}, (U,Lambda) the
   eigendecomposition of D R D, D = diag(R)^{-1/2}; and R = inv(S) = Lambda^{1/2} U' / D when S != R */
static PetscErrorCode BVMatSVQB_Private(BV bv, Mat R, Mat S)
{
  const PetscInt l = bv->l, k = bv->k, n = k - l, ld = R->lda;
  PetscScalar *pR = R->dense, *pS = (S == R) ? pR : S->dense;
  const PetscInt lds = (S == R) ? ld : S->lda;
  int n_ = n, lds_ = lds, info = 0, lwork = -1;
  double wq = 0.0, dummy = 0.0;
  LAPACK(dsyev)("V", "L", &n_, pS, &lds_, &dummy, &wq, &lwork, &info, 1, 1);
  lwork = (int)wq;
  if (lwork < 3 * n + 8) lwork = 3 * n + 8;
  double *eig = (double *)malloc(sizeof(double) * (size_t)(2 * n + lwork + 2));
  PetscCheck(eig, PETSC_ERR_MEM, "out of memory");
  double *D = eig + n, *work = D + n;
  for (PetscInt i = l; i < k; i++) D[i - l] = 1.0 / sqrt(pR[i + (size_t)i * ld]);
  for (PetscInt i = l; i < k; i++) for (PetscInt j = l; j < k; j++) pS[i + (size_t)j * lds] = pR[i + (size_t)j * ld] * D[i - l];
  for (PetscInt j = l; j < k; j++) for (PetscInt i = l; i < k; i++) pS[i + (size_t)j * lds] *= D[j - l];
  LAPACK(dsyev)("V", "L", &n_, pS + (size_t)l * lds + l, &lds_, eig, work, &lwork, &info, 1, 1);
  if (info) { free(eig); SETERRQ(PETSC_ERR_LIB, "Error in LAPACK subroutine syev: info=%d", info); }
  if (S != R) for (PetscInt i = l; i < k; i++) for (PetscInt j = l; j < k; j++) pR[i + (size_t)j * ld] = pS[j + (size_t)i * lds];   /* R = U' */
  for (PetscInt i = l; i < k; i++) for (PetscInt j = l; j < k; j++) pS[i + (size_t)j * lds] *= D[i - l];
  for (PetscInt j = l; j < k; j++) for (PetscInt i = l; i < k; i++) pS[i + (size_t)j * lds] /= sqrt(eig[j - l]);
  if (S != R) {
    for (PetscInt i = l; i < k; i++) for (PetscInt j = l; j < k; j++) pR[i + (size_t)j * ld] *= sqrt(eig[i - l]);
    for (PetscInt j = l; j < k; j++) for (PetscInt i = l; i < k; i++) pR[i + (size_t)j * ld] /= D[j - l];
  }
  free(eig);
  return PETSC_SUCCESS;
}

/* BVOrthogonalize_Chol bvorthog.c:586-601 (svqb = PETSC_FALSE) and BVOrthogonalize_SVQB :660-675 */
static PetscErrorCode BVOrthogonalize_Gram(BV V, Mat Rin, PetscBool svqb)
{
  Mat R, S;
  PetscCall(BV_GetBufferMat(V, &R));
  S = Rin ? Rin : R;                              /* use Rin as a workspace for S */
  PetscErrorCode ierr = PETSC_SUCCESS;
  if (V->l) ierr = BVOrthogonalize_BlockGS(V, R);
  if (!ierr) ierr = BVDot(V, V, R);
  if (!ierr) ierr = svqb ? BVMatSVQB_Private(V, R, S) : BVMatCholInv_Private(V, R, S);
  if (!ierr) ierr = BVMultInPlace(V, S, V->l, V->k);
  if (!ierr && Rin) ierr = BV_StoreCoeffsBlock(V, Rin, svqb ? PETSC_FALSE : PETSC_TRUE);
  PetscCall(MatDestroy(&R));
  PetscCall(ierr);
  return PETSC_SUCCESS;
}

/* every rank's block of `count` doubles, in rank order, on every rank (the communicators of this build reduce: a zero-padded sum) */
static PetscErrorCode BVAllgatherHost_Private(BV bv, const double *mine, PetscInt count, double *all)
{
  int rank = 0, size = 1;
  PetscCall(B2KCommGetRank(bv->comm, &rank, &size));
  memset(all, 0, sizeof(double) * (size_t)count * (size_t)size);
  memcpy(all + (size_t)rank * (size_t)count, mine, sizeof(double) * (size_t)count);
  if (size == 1) return PETSC_SUCCESS;
  if (bv->comm->kind == 2) return B2KCommAllreduce(bv->comm, all, (int)(count * size), 0, B2K_MEM_HOST);
  b2k_ctx ctx = B2KGetContext();                 /* NCCL communicator: device buffers */
  double *d = NULL;
  const size_t bytes = sizeof(double) * (size_t)count * (size_t)size;
  B2KCall(b2k_malloc(ctx, (void **)&d, bytes));
  B2KCall(b2k_h2d(ctx, d, all, bytes));
  PetscErrorCode ierr = B2KCommAllreduce(bv->comm, d, (int)(count * size), 0, B2K_MEM_DEVICE);
  if (!ierr) B2KCall(b2k_d2h(ctx, all, d, bytes));
  B2KCall(b2k_free(ctx, d));
  return ierr;
}

/* BVMatTriInv_LAPACK_Private bvlapack.c:208-250: S(l:k,l:k) = inv(R(l:k,l:k)), R upper triangular; S == R inverts in place */
static PetscErrorCode BVMatTriInv_Private(BV bv, Mat R, Mat S)
{
  const PetscInt l = bv->l, k = bv->k, n = k - l, ld = R->lda;
  int n_ = n, ld_ = ld, info = 0;
  PetscScalar *pR = R->dense;
  if (S == R) LAPACK(dtrtri)("U", "N", &n_, pR + (size_t)l * ld + l, &ld_, &info, 1, 1);
  else {
    PetscScalar *pS = S->dense;
    const PetscInt lds = S->lda;
    int lds_ = lds;
    memset(pS + (size_t)l * lds, 0, sizeof(PetscScalar) * (size_t)(k - l) * (size_t)lds);
    for (PetscInt i = l; i < k; i++) memcpy(pS + (size_t)i * lds + l, pR + (size_t)i * ld + l, sizeof(PetscScalar) * (size_t)n);
    LAPACK(dtrtri)("U", "N", &n_, pS + (size_t)l * lds + l, &lds_, &info, 1, 1);
  }
  PetscCheck(!info, PETSC_ERR_LIB, "Error in LAPACK subroutine trtri: info=%d", info);
  return PETSC_SUCCESS;
}

/* BVOrthogonalize_TSQR bvorthog.c:611-631 (onlyr = PETSC_FALSE: Q formed from the reflectors) and BVOrthogonalize_TSQRCHOL :636-656
   (onlyr = PETSC_TRUE: only R from the Householder tree, then Q = V inv(R) with BVMultInPlace).  The local factorisation is the
   type's (ops.tsqr_factor / ops.tsqr_formq: device kernels for BV type b200, LAPACK for a host type); the level over the ranks —
   the binary tree of bvlapack.c:386-462 and the packed-Givens reduction of :474-560 in the reference — is ONE QR of the gathered
   k x k triangles here (at most 8 of them), done redundantly by every rank, so all ranks hold the same R. */
static PetscErrorCode BVOrthogonalize_TSQR(BV V, Mat Rin, PetscBool onlyr)
{
  Mat R, S;
  const PetscInt l = V->l, k = V->k, n = k - l;
  int rank = 0, size = 1;
  PetscCheck(V->ops.tsqr_factor && (onlyr || V->ops.tsqr_formq), PETSC_ERR_SUP, "BV type %s has no TSQR block orthogonalization; use chol or svqb", V->type_name);
  PetscCall(B2KCommGetRank(V->comm, &rank, &size));
  PetscCall(BV_GetBufferMat(V, &R));
  S = Rin ? Rin : R;
  double *Rl = (double *)calloc((size_t)n * (size_t)n + 1, sizeof(double)), *W = (double *)calloc((size_t)n * (size_t)n + 1, sizeof(double));
  double *all = NULL, *stack = NULL;
  PetscErrorCode ierr = (Rl && W) ? PETSC_SUCCESS : PETSC_ERR_MEM;
  if (!ierr && V->l) ierr = BVOrthogonalize_BlockGS(V, R);
  if (!ierr) ierr = V->ops.tsqr_factor(V, onlyr ? PETSC_FALSE : PETSC_TRUE, Rl);
  for (PetscInt i = 0; i < n; i++) W[i + (size_t)i * n] = 1.0;
  if (!ierr && size > 1 && n > 0) {
    const PetscInt rows = (PetscInt)size * n;
    int rows_ = rows, n_ = n, lwork = 64 * (int)n + 64, info = 0;
    all = (double *)malloc(sizeof(double) * (size_t)rows * (size_t)n);
    stack = (double *)malloc(sizeof(double) * ((size_t)rows * (size_t)n + (size_t)n + (size_t)lwork));
    if (!all || !stack) ierr = PETSC_ERR_MEM;
    if (!ierr) ierr = BVAllgatherHost_Private(V, Rl, n * n, all);
    if (!ierr) {
      double *tau = stack + (size_t)rows * n, *work = tau + n;
      for (int r = 0; r < size; r++)
        for (PetscInt c = 0; c < n; c++) memcpy(stack + (size_t)r * n + (size_t)c * rows, all + (size_t)r * n * n + (size_t)c * n, sizeof(double) * (size_t)n);
      LAPACK(dgeqrf)(&rows_, &n_, stack, &rows_, tau, work, &lwork, &info);
      if (!info) {
        for (PetscInt c = 0; c < n; c++) for (PetscInt i = 0; i < n; i++) Rl[i + (size_t)c * n] = (i <= c) ? stack[i + (size_t)c * rows] : 0.0;
        if (!onlyr) {
          LAPACK(dorgqr)(&rows_, &n_, &n_, stack, &rows_, tau, work, &lwork, &info);
          for (PetscInt c = 0; c < n; c++) memcpy(W + (size_t)c * n, stack + (size_t)rank * n + (size_t)c * rows, sizeof(double) * (size_t)n);
        }
      }
      if (info) ierr = PETSC_ERR_LIB;
    }
  }
  if (!ierr) {
    for (PetscInt i = 0; i < n; i++) {            /* R with a non-negative diagonal (the unique factor; LAPACK's sign is arbitrary) */
      if (Rl[i + (size_t)i * n] < 0.0) {
        for (PetscInt c = i; c < n; c++) Rl[i + (size_t)c * n] = -Rl[i + (size_t)c * n];
        for (PetscInt r = 0; r < n; r++) W[r + (size_t)i * n] = -W[r + (size_t)i * n];
      }
    }
    const PetscInt ldr = R->lda;
    for (PetscInt c = 0; c < n; c++) for (PetscInt i = 0; i < n; i++) R->dense[(l + i) + (size_t)(l + c) * ldr] = Rl[i + (size_t)c * n];
    if (onlyr) {
      ierr = BVMatTriInv_Private(V, R, S);
      if (!ierr) ierr = BVMultInPlace(V, S, V->l, V->k);
    } else ierr = V->ops.tsqr_formq(V, W);
  }
  if (!ierr && Rin) ierr = BV_StoreCoeffsBlock(V, Rin, PETSC_TRUE);
  free(Rl); free(W); free(all); free(stack);
  PetscCall(MatDestroy(&R));
  PetscCall(ierr);
  return PETSC_SUCCESS;
}

/* BVOrthogonalize_GS bvorthog.c:504-540: column by column */
static PetscErrorCode BVOrthogonalize_GS(BV V, Mat R)
{
  PetscInt ldr = 0;
  PetscScalar *r = NULL;
  if (R) { ldr = R->lda; r = R->dense; }
  const PetscInt lsave = V->l, ksave = V->k;
  for (PetscInt j = lsave; j < ksave; j++) {
    PetscReal norm;
    V->l = lsave; V->k = ksave;
    PetscCall(BVOrthogonalizeColumn(V, j, NULL, &norm, NULL));       /* the reference does not ask for lindep here (bvorthog.c:531-538): a
                                                                         dependent column is reduced to noise and normalised, bv/tests/test12.c */
    PetscCheck(norm != 0.0, PETSC_ERR_CONV_FAILED, "Breakdown in BVOrthogonalize due to a linearly dependent column");
    PetscCall(V->ops.scale(V, j, 1.0 / norm));
    if (r) {
      for (PetscInt i = 0; i < j; i++) r[i + (size_t)j * ldr] = BV_BUF(V, V->nc + i, j);
      r[j + (size_t)j * ldr] = norm;
      for (PetscInt i = j + 1; i < ksave; i++) r[i + (size_t)j * ldr] = 0.0;
    }
  }
  V->l = lsave; V->k = ksave;
  return PETSC_SUCCESS;
}

PetscErrorCode BVOrthogonalize(BV V, Mat R)
{
  BVCheckSizes(V);
  if (R) {
    PetscCheck(R->dense, PETSC_ERR_SUP, "Mat argument must be sequential dense");
    PetscCheck(R->m == R->n, PETSC_ERR_ARG_SIZ, "Mat argument is not square, it has %d rows and %d columns", R->m, R->n);
    PetscCheck(R->n >= V->k, PETSC_ERR_ARG_SIZ, "Mat size %d is smaller than the number of BV active columns %d", R->n, V->k);
  }
  PetscCheck(V->nc == 0, PETSC_ERR_SUP, "Not implemented for BV with constraints, use BVOrthogonalizeColumn() instead");
  switch (V->orthog_block) {
  case BV_ORTHOG_BLOCK_GS: PetscCall(BVOrthogonalize_GS(V, R)); break;
  case BV_ORTHOG_BLOCK_CHOL: PetscCall(BVOrthogonalize_Gram(V, R, PETSC_FALSE)); break;
  case BV_ORTHOG_BLOCK_SVQB: PetscCall(BVOrthogonalize_Gram(V, R, PETSC_TRUE)); break;
  case BV_ORTHOG_BLOCK_TSQR: PetscCall(BVOrthogonalize_TSQR(V, R, PETSC_FALSE)); break;
  case BV_ORTHOG_BLOCK_TSQRCHOL: PetscCall(BVOrthogonalize_TSQR(V, R, PETSC_TRUE)); break;
  }
  V->state++;
  return PETSC_SUCCESS;
}

/* ---- Krylov recurrences: bvkrylov.c ------------------------------------------------------------------------ */
/* BV_OrthogonalizeColumn_Safe bvimpl.h:452-465: when the full basis is requested the last vector is
   numerically zero, so it is orthogonalised without refinement and flagged dependent */
static PetscErrorCode BV_OrthogonalizeColumn_Safe(BV bv, PetscInt j, PetscReal *norm, PetscBool *lindep)
{
  const BVOrthogRefineType ref = bv->orthog_ref;
  bv->orthog_ref = BV_ORTHOG_REFINE_NEVER;
  PetscErrorCode ierr = BVOrthogonalizeColumn(bv, j, NULL, NULL, NULL);
  bv->orthog_ref = ref;
  PetscCall(ierr);
  if (norm) *norm = 0.0;
  if (lindep) *lindep = PETSC_TRUE;
  return PETSC_SUCCESS;
}

static PetscErrorCode BVKrylovLoop_Private(BV V, Mat A, PetscInt k, PetscInt *m, PetscReal *beta, PetscBool *lindep_out)
{
  PetscBool lindep = PETSC_FALSE;
  PetscReal b = 0.0;
  /* ops.krylov_steps: the type runs the recurrence on its own (BV type b200: the whole cycle is enqueued on the stream, the DGKS
     decision and the normalisation are taken on the device, the host reads every coefficient once at the end); whatever it
     leaves — the full-basis step j = N-1, a breakdown, a refinement pass it had not prepared — is done here, one synchronised
     step, after which the type is asked again for the rest */
  const PetscInt mm = (*m > V->N - 1) ? V->N - 1 : *m;
  const PetscBool ask = V->ops.krylov_steps ? PETSC_TRUE : PETSC_FALSE;
  for (PetscInt j = k; j < *m; j++) {
    if (ask && j < mm) {
      PetscInt jn = j;
      PetscCall(V->ops.krylov_steps(V, A, j, mm, &jn, &b));
      j = jn;
      if (j >= *m) break;
    }
    PetscCall(BVMatMultColumn(V, A, j));
    if (j == V->N - 1) PetscCall(BV_OrthogonalizeColumn_Safe(V, j + 1, &b, &lindep));
    else PetscCall(BVOrthonormalizeColumn(V, j + 1, PETSC_FALSE, &b, &lindep));
    if (lindep) { *m = j + 1; break; }
  }
  if (beta) *beta = b;
  *lindep_out = lindep;
  return PETSC_SUCCESS;
}

PetscErrorCode BVMatArnoldi(BV V, Mat A, Mat H, PetscInt k, PetscInt *m, PetscReal *beta, PetscBool *breakdown)
{
  BVCheckSizes(V);
  PetscCheck(A && m, PETSC_ERR_ARG_NULL, "null argument");
  PetscCheck(k >= 0 && k <= V->m, PETSC_ERR_ARG_OUTOFRANGE, "Argument k has wrong value %d, should be between 0 and %d", k, V->m);
  PetscCheck(*m > 0 && *m <= V->m, PETSC_ERR_ARG_OUTOFRANGE, "Argument m has wrong value %d, should be between 1 and %d", *m, V->m);
  PetscCheck(*m > k, PETSC_ERR_ARG_OUTOFRANGE, "Argument m should be at least equal to k+1");
  if (H) {
    PetscCheck(H->dense, PETSC_ERR_SUP, "Matrix H must be sequential dense");
    PetscCheck(H->m >= *m, PETSC_ERR_ARG_SIZ, "Matrix H has %d rows, should have at least %d", H->m, *m);
    PetscCheck(H->n >= *m, PETSC_ERR_ARG_SIZ, "Matrix H has %d columns, should have at least %d", H->n, *m);
  }
  PetscBool lindep;
  PetscCall(BVKrylovLoop_Private(V, A, k, m, beta, &lindep));
  if (breakdown) *breakdown = lindep;
  if (H) {
    PetscScalar *h = H->dense;
    const size_t ldh = (size_t)H->lda, ldb = (size_t)(V->nc + V->m);
    const PetscScalar *a = V->buffer;
    for (PetscInt j = k; j < *m - 1; j++) memcpy(h + j * ldh, a + V->nc + (size_t)(j + 1) * ldb, sizeof(PetscScalar) * (size_t)(j + 2));
    memcpy(h + (size_t)(*m - 1) * ldh, a + V->nc + (size_t)(*m) * ldb, sizeof(PetscScalar) * (size_t)(*m));
    if (H->lda > *m) h[(*m) + (size_t)(*m - 1) * ldh] = a[V->nc + (*m) + (size_t)(*m) * ldb];
  }
  V->state++;
  return PETSC_SUCCESS;
}

PetscErrorCode BVMatLanczos(BV V, Mat A, Mat T, PetscInt k, PetscInt *m, PetscReal *beta, PetscBool *breakdown)
{
  BVCheckSizes(V);
  PetscCheck(A && m, PETSC_ERR_ARG_NULL, "null argument");
  PetscCheck(k >= 0 && k <= V->m, PETSC_ERR_ARG_OUTOFRANGE, "Argument k has wrong value %d, should be between 0 and %d", k, V->m);
  PetscCheck(*m > 0 && *m <= V->m, PETSC_ERR_ARG_OUTOFRANGE, "Argument m has wrong value %d, should be between 1 and %d", *m, V->m);
  PetscCheck(*m > k, PETSC_ERR_ARG_OUTOFRANGE, "Argument m should be at least equal to k+1");
  if (T) {
    PetscCheck(T->dense, PETSC_ERR_SUP, "Matrix T must be sequential dense");
    PetscCheck(T->m >= *m, PETSC_ERR_ARG_SIZ, "Matrix T has %d rows, should have at least %d", T->m, *m);
    PetscCheck(T->n >= 2, PETSC_ERR_ARG_SIZ, "Matrix T has %d columns, should have at least %d", T->n, 2);
  }
  PetscBool lindep;
  PetscCall(BVKrylovLoop_Private(V, A, k, m, beta, &lindep));
  if (breakdown) *breakdown = lindep;
  if (T) {
    PetscReal *alpha = T->dense, *betat = T->dense + T->lda;
    const size_t ldb = (size_t)(V->nc + V->m);
    for (PetscInt j = k; j < *m; j++) {
      alpha[j] = V->buffer[V->nc + j + (size_t)(j + 1) * ldb];
      betat[j] = V->buffer[V->nc + j + 1 + (size_t)(j + 1) * ldb];
    }
  }
  V->state++;
  return PETSC_SUCCESS;
}

PetscErrorCode BVGetCounters(BV bv, int64_t *gs_passes, int64_t *matmults)
{
  if (gs_passes) *gs_passes = bv->n_gs_passes;
  if (matmults) *matmults = bv->n_matmult;
  return PETSC_SUCCESS;
}
