/* MINIMAL PETSc STAND-IN (see ../../petscsys.h): the object header that PETSCHEADER(ops) expands to */
#pragma once
#include <petscsys.h>
typedef struct {
  PetscErrorCode (*view)(PetscObject, PetscViewer);
  PetscErrorCode (*destroy)(PetscObject *);
} PetscOps;
struct _p_PetscObject { PetscClassId classid; PetscOps bops[1]; MPI_Comm comm; PetscObjectId id; int refct; char *type_name, *name, *prefix; PetscObjectState state; void *options; };
#define PETSCHEADER(ObjectOps) \
  struct _p_PetscObject hdr; \
  ObjectOps ops[1]
#define PetscHeaderCreate(h, classid, class_name, descr, mansec, comm, destroy, view) PetscHeaderCreate_Stub((void **)&(h), sizeof(*(h)), classid, class_name, comm)
PetscErrorCode PetscHeaderCreate_Stub(void **, size_t, PetscClassId, const char *, MPI_Comm);
PetscErrorCode PetscHeaderDestroy_Stub(void **);
#define PetscHeaderDestroy(h) PetscHeaderDestroy_Stub((void **)(h))
PetscErrorCode PetscFunctionListAdd_Private(PetscFunctionList *, const char *, void (*)(void));
#define PetscFunctionListAdd(list, name, fptr) PetscFunctionListAdd_Private((list), (name), (void (*)(void))(fptr))
PetscErrorCode PetscFunctionListFind_Private(PetscFunctionList, const char *, void (**)(void));
#define PetscFunctionListFind(list, name, fptr) PetscFunctionListFind_Private((list), (name), (void (**)(void))(fptr))
PetscErrorCode PetscObjectChangeTypeName(PetscObject, const char *);
PetscErrorCode PetscSplitReductionGet(MPI_Comm, PetscSplitReduction **);
