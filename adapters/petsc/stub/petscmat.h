/* MINIMAL PETSc STAND-IN (see petscsys.h in this directory): Vec / Mat / KSP handles and the prototypes the adapter uses */
#pragma once
#include <petscsys.h>
typedef struct _p_Vec *Vec;
typedef struct _p_Mat *Mat;
typedef struct _p_KSP *KSP;
typedef struct _p_PC *PC;
typedef struct _p_IS *IS;
typedef struct _p_VecScatter *VecScatter;
typedef struct _p_MatNullSpace *MatNullSpace;
typedef const char *VecType;
typedef const char *MatType;
typedef const char *KSPType;
typedef const char *PCType;
typedef const char *MatSolverType;
typedef enum { MAT_INITIAL_MATRIX, MAT_REUSE_MATRIX, MAT_IGNORE_MATRIX, MAT_INPLACE_MATRIX } MatReuse;
typedef enum { DIFFERENT_NONZERO_PATTERN, SUBSET_NONZERO_PATTERN, SAME_NONZERO_PATTERN, UNKNOWN_NONZERO_PATTERN } MatStructure;
typedef enum { MAT_DO_NOT_COPY_VALUES, MAT_COPY_VALUES, MAT_SHARE_NONZERO_PATTERN } MatDuplicateOption;
typedef enum { MATOP_MULT = 3, MATOP_MULT_TRANSPOSE = 5, MATOP_GET_DIAGONAL = 17, MATOP_DESTROY = 60, MATOP_CREATE_VECS = 88 } MatOperation;
typedef enum { MAT_FACTOR_NONE, MAT_FACTOR_LU, MAT_FACTOR_CHOLESKY } MatFactorType;
typedef int MatProductType;
typedef int MatOption;
typedef int MatAssemblyType;
#define VECSTANDARD "standard"
#define VECSEQ "seq"
#define VECMPI "mpi"
#define VECCUDA "cuda"
#define VECSEQCUDA "seqcuda"
#define VECMPICUDA "mpicuda"
#define MATSEQDENSE "seqdense"
#define MATSEQDENSECUDA "seqdensecuda"
#define MATDENSE "dense"
#define MATSHELL "shell"
#define MATAIJ "aij"
#define MATAIJCUSPARSE "aijcusparse"
PetscErrorCode VecGetLocalSize(Vec, PetscInt *);
PetscErrorCode VecGetSize(Vec, PetscInt *);
PetscErrorCode VecGetOwnershipRange(Vec, PetscInt *, PetscInt *);
PetscErrorCode VecGetType(Vec, VecType *);
PetscErrorCode VecCUDAGetArray(Vec, PetscScalar **);
PetscErrorCode VecCUDARestoreArray(Vec, PetscScalar **);
PetscErrorCode VecCUDAGetArrayRead(Vec, const PetscScalar **);
PetscErrorCode VecCUDARestoreArrayRead(Vec, const PetscScalar **);
PetscErrorCode VecCUDAGetArrayWrite(Vec, PetscScalar **);
PetscErrorCode VecCUDARestoreArrayWrite(Vec, PetscScalar **);
PetscErrorCode VecCUDAPlaceArray(Vec, const PetscScalar *);
PetscErrorCode VecCUDAResetArray(Vec);
PetscErrorCode VecGetArray(Vec, PetscScalar **);
PetscErrorCode VecRestoreArray(Vec, PetscScalar **);
PetscErrorCode VecGetArrayRead(Vec, const PetscScalar **);
PetscErrorCode VecRestoreArrayRead(Vec, const PetscScalar **);
PetscErrorCode VecCreateMPICUDAWithArray(MPI_Comm, PetscInt, PetscInt, PetscInt, const PetscScalar *, Vec *);
PetscErrorCode VecCreateSeqCUDAWithArray(MPI_Comm, PetscInt, PetscInt, const PetscScalar *, Vec *);
PetscErrorCode VecCreate(MPI_Comm, Vec *);
PetscErrorCode VecSetSizes(Vec, PetscInt, PetscInt);
PetscErrorCode VecSetType(Vec, VecType);
PetscErrorCode VecDestroy(Vec *);
PetscErrorCode VecDuplicate(Vec, Vec *);
PetscErrorCode VecCopy(Vec, Vec);
PetscErrorCode VecNorm(Vec, NormType, PetscReal *);
PetscErrorCode VecDot(Vec, Vec, PetscScalar *);
PetscErrorCode VecAXPY(Vec, PetscScalar, Vec);
PetscErrorCode VecScale(Vec, PetscScalar);
PetscErrorCode VecSet(Vec, PetscScalar);
PetscErrorCode VecView(Vec, PetscViewer);
PetscErrorCode VecSetRandom(Vec, PetscRandom);
PetscErrorCode VecBindToCPU(Vec, PetscBool);
PetscErrorCode VecGetLayout(Vec, PetscLayout *);
PetscErrorCode MatCreateShell(MPI_Comm, PetscInt, PetscInt, PetscInt, PetscInt, void *, Mat *);
PetscErrorCode MatShellGetContext(Mat, void *);
PetscErrorCode MatShellSetOperation(Mat, MatOperation, void (*)(void));
PetscErrorCode MatShellSetVecType(Mat, VecType);
PetscErrorCode MatGetSize(Mat, PetscInt *, PetscInt *);
PetscErrorCode MatGetLocalSize(Mat, PetscInt *, PetscInt *);
PetscErrorCode MatGetOwnershipRange(Mat, PetscInt *, PetscInt *);
PetscErrorCode MatGetOwnershipRangeColumn(Mat, PetscInt *, PetscInt *);
PetscErrorCode MatGetRow(Mat, PetscInt, PetscInt *, const PetscInt **, const PetscScalar **);
PetscErrorCode MatRestoreRow(Mat, PetscInt, PetscInt *, const PetscInt **, const PetscScalar **);
PetscErrorCode MatGetType(Mat, MatType *);
PetscErrorCode MatMult(Mat, Vec, Vec);
PetscErrorCode MatMultTranspose(Mat, Vec, Vec);
PetscErrorCode MatCreateVecs(Mat, Vec *, Vec *);
PetscErrorCode MatDestroy(Mat *);
PetscErrorCode MatDenseGetArrayRead(Mat, const PetscScalar **);
PetscErrorCode MatDenseRestoreArrayRead(Mat, const PetscScalar **);
PetscErrorCode MatDenseGetArray(Mat, PetscScalar **);
PetscErrorCode MatDenseRestoreArray(Mat, PetscScalar **);
PetscErrorCode MatDenseGetArrayWrite(Mat, PetscScalar **);
PetscErrorCode MatDenseRestoreArrayWrite(Mat, PetscScalar **);
PetscErrorCode MatDenseGetLDA(Mat, PetscInt *);
PetscErrorCode MatCreateDense(MPI_Comm, PetscInt, PetscInt, PetscInt, PetscInt, PetscScalar *, Mat *);
PetscErrorCode MatCreateDenseCUDA(MPI_Comm, PetscInt, PetscInt, PetscInt, PetscInt, PetscScalar *, Mat *);
PetscErrorCode MatDenseCUDAGetArray(Mat, PetscScalar **);
PetscErrorCode MatDenseCUDARestoreArray(Mat, PetscScalar **);
PetscErrorCode MatRegister(const char *, PetscErrorCode (*)(Mat));
/* PETSc calls made by inline helpers of the SLEPc headers themselves (never by the adapter): old-style declarations, enough
   for -fsyntax-only */
PetscErrorCode MatCreateRedundantMatrix();
PetscErrorCode MatConvert();
PetscErrorCode PetscObjectQuery();
PetscErrorCode PetscContainerGetPointer();
PetscErrorCode PetscSNPrintfCount();
PetscErrorCode PetscSNPrintf();
PetscErrorCode PetscObjectGetId();
PetscErrorCode VecGetState();
PetscErrorCode PetscMalloc2();
PetscErrorCode PetscFree2();
PetscErrorCode VecCreateSeqCUDA();
PetscErrorCode VecCreateSeq();
PetscErrorCode MatCreate();
PetscErrorCode MatSetSizes();
PetscErrorCode MatSetType();
PetscErrorCode MatDiagonalSet();
PetscErrorCode PetscIntCast();
