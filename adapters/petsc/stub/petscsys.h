/* MINIMAL PETSc STAND-IN — for `gcc -fsyntax-only` of the adapter sources against the reference's own SLEPc headers on a
 * machine without PETSc (tests/test_adapter_syntax.py).  It declares just enough of petscsys.h / petscmat.h /
 * petsc/private/petscimpl.h for include/slepc*.h and include/slepc/private/{slepcimpl,bvimpl}.h to parse and for the adapter
 * to type-check: types, the object header, the error macros, and prototypes of the PETSc calls the adapter makes.
 * It is NOT an implementation and is never linked; with a real PETSc 3.22 installation this directory is simply left out
 * of the include path. */
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define PETSC_EXTERN extern
#define PETSC_EXTERN_TYPEDEF
#define PETSC_DLLIMPORT
#define MPIAPI
#define PETSC_MEMALIGN 16
#define PETSC_VISIBILITY_INTERNAL
#define PETSC_DLLEXPORT
#define PETSC_USE_REAL_DOUBLE 1
#define PETSC_HAVE_CUDA 1
#define PETSC_HAVE_CUPM 1
#define PETSC_CLANG_STATIC_ANALYZER_IGNORE 0
#define PETSC_INTERN extern
#define PETSC_VISIBILITY_PUBLIC
#define PETSC_VERSION_MAJOR 3
#define PETSC_VERSION_MINOR 22
#define PETSC_VERSION_SUBMINOR 0
#define PETSC_VERSION_RELEASE 1
#define PETSC_VERSION_LT(a, b, c) (0)
#define PETSC_VERSION_LE(a, b, c) (0)
#define PETSC_VERSION_GE(a, b, c) (1)
#define PETSC_VERSION_GT(a, b, c) (1)
#define PETSC_VERSION_(a, b, c) (1)
#define PETSC_ATTRIBUTE_FORMAT(a, b)
#define PETSC_DEPRECATED_FUNCTION(a, b, c, ...)
#define PETSC_DEPRECATED_MACRO(a, b, c, ...)
#define PETSC_DEPRECATED_ENUM(a, b, c, ...)
#define PETSC_DEPRECATED_TYPEDEF(a, b, c, ...)
#define PETSC_NODISCARD
#define PETSC_SINGLE_LIBRARY_INTERN extern
#define PETSC_UNUSED __attribute__((unused))
#define PETSC_STATIC_INLINE static inline
#define PETSC_NULLPTR NULL
#define PetscDefined(x) 0
#define PetscHasAttribute(x) 0
#define PetscUnlikely(x) (x)
#define PetscLikely(x) (x)

typedef int       PetscErrorCode;
typedef int       PetscInt;
typedef int       PetscMPIInt;
typedef int       PetscBLASInt;
typedef int64_t   PetscInt64;
typedef int64_t   PetscCount;
typedef double    PetscReal;
typedef double    PetscScalar;
typedef double    PetscLogDouble;
typedef int       PetscClassId;
typedef int       PetscLogEvent;
typedef int       PetscObjectState;
typedef int       PetscObjectId;
typedef enum { PETSC_FALSE, PETSC_TRUE } PetscBool;
typedef enum { PETSC_COPY_VALUES, PETSC_OWN_POINTER, PETSC_USE_POINTER } PetscCopyMode;
typedef enum { PETSC_MEMTYPE_HOST = 0, PETSC_MEMTYPE_DEVICE = 1, PETSC_MEMTYPE_CUDA = 1 } PetscMemType;
typedef int MPI_Comm;
typedef int MPI_Op;
typedef int MPI_Datatype;
#define MPI_COMM_NULL 0
#define PETSC_COMM_SELF 1
#define PETSC_COMM_WORLD 2
#define MPIU_SCALAR 0
#define MPIU_REAL 0
#define MPIU_SUM 0
#define MPIU_MAX 0
#define MPIU_MIN 0
#define MPI_IN_PLACE ((void *)1)
#define PETSC_SUCCESS 0
#define PETSC_ERR_MEM 55
#define PETSC_ERR_SUP 56
#define PETSC_ERR_ORDER 58
#define PETSC_ERR_LIB 76
#define PETSC_ERR_PLIB 77
#define PETSC_ERR_ARG_SIZ 60
#define PETSC_ERR_ARG_WRONG 62
#define PETSC_ERR_ARG_INCOMP 75
#define PETSC_ERR_ARG_OUTOFRANGE 63
#define PETSC_ERR_ARG_WRONGSTATE 73
#define PETSC_ERR_GPU 97
#define PETSC_ERR_USER_INPUT 83
#define PETSC_ERR_FP 72
#define PETSC_ERR_USER 83
#define PETSC_ERR_CONV_FAILED 82
#define PETSC_ERR_ARG_NULL 85
#define PETSC_DECIDE (-1)
#define PETSC_DETERMINE PETSC_DECIDE
#define PETSC_DEFAULT (-2)
#define PETSC_CURRENT (-2)
#define PETSC_MAX_PATH_LEN 4096
#define PETSC_MACHINE_EPSILON 2.220446049250313e-16
#define PETSC_SQRT_MACHINE_EPSILON 1.490116119384766e-08
#define PETSC_SMALL 1.e-20
#define PETSC_MAX_REAL 1.7976931348623157e+308
#define PETSC_MIN_REAL (-PETSC_MAX_REAL)
#define PETSC_PI 3.1415926535897932384626433832795029
#define PETSC_i 0
#define PetscInt_FMT "d"
#define PetscMax(a, b) (((a) < (b)) ? (b) : (a))
#define PetscMin(a, b) (((a) < (b)) ? (a) : (b))
#define PetscAbsScalar(a) fabs(a)
#define PetscAbsReal(a) fabs(a)
#define PetscRealPart(a) (a)
#define PetscImaginaryPart(a) ((PetscReal)0)
#define PetscConj(a) (a)
#define PetscSqrtReal(a) sqrt(a)
#define PetscSqrtScalar(a) sqrt(a)
#define PetscSqr(a) ((a) * (a))
#define PetscLogReal(a) log(a)
#define PetscExpReal(a) exp(a)
#define PetscPowReal(a, b) pow(a, b)
#define PetscPowScalar(a, b) pow(a, b)
#define PetscCosReal(a) cos(a)
#define PetscSinReal(a) sin(a)
#define PetscAtan2Real(a, b) atan2(a, b)
#define PetscFloorReal(a) floor(a)
#define PetscCeilReal(a) ceil(a)
#define PetscIsInfOrNanReal(a) (!isfinite(a))
#define PetscIsInfOrNanScalar(a) (!isfinite(a))
#define PetscIsNanReal(a) isnan(a)
#define PetscIsNanScalar(a) isnan(a)
#define PetscCMPLX(a, b) (a)
#define PetscSign(a) (((a) >= 0) ? 1 : -1)
#define PetscAbs(a) (((a) >= 0) ? (a) : (-(a)))
#define PetscSafePointerPlusOffset(p, o) ((p) ? (p) + (o) : NULL)

typedef struct _p_PetscObject *PetscObject;
typedef struct _n_PetscSubcomm *PetscSubcomm;
typedef int PetscSubcommType;
typedef struct _p_PetscViewer *PetscViewer;
typedef struct _p_PetscRandom *PetscRandom;
typedef struct _p_PetscOptions *PetscOptions;
typedef struct _n_PetscFunctionList *PetscFunctionList;
typedef struct _p_PetscOptionItems PetscOptionItems;
typedef struct _p_PetscLayout *PetscLayout;
typedef struct _p_PetscContainer *PetscContainer;
typedef struct _p_PetscDraw *PetscDraw;
typedef struct _p_PetscDrawLG *PetscDrawLG;
typedef struct _p_PetscDrawSP *PetscDrawSP;
typedef struct _n_PetscViewerAndFormat PetscViewerAndFormat;
typedef struct _p_PetscDeviceContext *PetscDeviceContext;
typedef const char *PetscViewerType;
typedef int PetscViewerFormat;
typedef int PetscDataType;
typedef void (*PetscVoidFn)(void);
typedef PetscVoidFn *PetscVoidFunction;
typedef PetscErrorCode (*PetscErrorCodeFn)(void);
typedef struct { int dummy; } PetscSplitReduction;
typedef enum { NORM_1 = 0, NORM_2 = 1, NORM_FROBENIUS = 2, NORM_INFINITY = 3, NORM_1_AND_2 = 4 } NormType;
#define NORM_MAX NORM_INFINITY
typedef enum { INSERT_VALUES = 1, ADD_VALUES = 2 } InsertMode;

#define PetscFunctionBegin do { } while (0)
#define PetscFunctionBeginUser do { } while (0)
#define PetscFunctionReturn(x) return (x)
#define PetscFunctionReturnVoid() return
PetscErrorCode PetscError(MPI_Comm, int, const char *, const char *, PetscErrorCode, int, const char *, ...);
#define SETERRQ(comm, ierr, ...) return PetscError(comm, __LINE__, __func__, __FILE__, ierr, 0, __VA_ARGS__)
#define PetscCheck(cond, comm, ierr, ...) do { if (!(cond)) SETERRQ(comm, ierr, __VA_ARGS__); } while (0)
#define PetscAssert(cond, comm, ierr, ...) do { if (!(cond)) SETERRQ(comm, ierr, __VA_ARGS__); } while (0)
#define PetscCall(...) do { PetscErrorCode ierr_petsc_call_ = __VA_ARGS__; if (ierr_petsc_call_) return ierr_petsc_call_; } while (0)
#define PetscCallMPI(...) PetscCall(__VA_ARGS__)
#define PetscCallExternal(func, ...) do { if (func(__VA_ARGS__)) return PETSC_ERR_LIB; } while (0)
#define PetscCallAbort(comm, ...) do { if (__VA_ARGS__) abort(); } while (0)
#define PetscCallVoid(...) do { if (__VA_ARGS__) return; } while (0)
#define CHKERRQ(ierr) do { if (ierr) return ierr; } while (0)
#define PetscValidHeaderSpecific(h, ck, arg) do { (void)(h); } while (0)
#define PetscValidHeader(h, arg) do { (void)(h); } while (0)
#define PetscValidType(h, arg) do { (void)(h); } while (0)
#define PetscAssertPointer(p, arg) do { (void)(p); } while (0)
#define PetscValidLogicalCollectiveInt(a, b, c) do { } while (0)
#define PetscValidLogicalCollectiveEnum(a, b, c) do { } while (0)
#define PetscValidLogicalCollectiveBool(a, b, c) do { } while (0)
#define PetscValidLogicalCollectiveScalar(a, b, c) do { } while (0)
#define PetscValidLogicalCollectiveReal(a, b, c) do { } while (0)
#define PetscCheckSameComm(a, b, c, d) do { } while (0)
#define PetscCheckSameTypeAndComm(a, b, c, d) do { } while (0)
#define PetscCheckSameType(a, b, c, d) do { } while (0)
#define PetscUseTypeMethod(obj, op, ...) PetscCall((*(obj)->ops->op)(obj, ##__VA_ARGS__))
#define PetscTryTypeMethod(obj, op, ...) do { if ((obj)->ops->op) PetscCall((*(obj)->ops->op)(obj, ##__VA_ARGS__)); } while (0)
#define PetscUseMethod(obj, name, args, vals) do { } while (0)
#define PetscTryMethod(obj, name, args, vals) do { } while (0)
#define PetscDesignatedInitializer(name, ...) .name = __VA_ARGS__
#define PetscObjectComm(o) ((o) ? 2 : 0)
#define PetscObjectComposeFunction(a, b, c) PetscObjectComposeFunction_Private(a, b, (void (*)(void))(c))

PetscErrorCode PetscObjectComposeFunction_Private(PetscObject, const char *, void (*)(void));
PetscErrorCode PetscObjectStateIncrease(PetscObject);
PetscErrorCode PetscObjectStateGet(PetscObject, PetscObjectState *);
PetscErrorCode PetscObjectGetComm(PetscObject, MPI_Comm *);
PetscErrorCode PetscObjectReference(PetscObject);
PetscErrorCode PetscObjectTypeCompare(PetscObject, const char *, PetscBool *);
PetscErrorCode PetscObjectTypeCompareAny(PetscObject, PetscBool *, const char *, ...);
PetscErrorCode PetscObjectGetName(PetscObject, const char **);
PetscErrorCode PetscObjectSetName(PetscObject, const char *);
PetscErrorCode PetscMalloc(size_t, void *);
PetscErrorCode PetscFree_Private(void *);
#define PetscFree(p) (PetscFree_Private((void *)(p)) || ((p) = NULL, 0))
#define PetscNew(p) PetscMalloc(sizeof(**(p)), (p))
#define PetscMalloc1(n, p) PetscMalloc((size_t)(n) * sizeof(**(p)), (p))
#define PetscCalloc1(n, p) PetscMalloc((size_t)(n) * sizeof(**(p)), (p))
PetscErrorCode PetscArraycpy_Private(void *, const void *, size_t);
#define PetscArraycpy(a, b, n) PetscArraycpy_Private((a), (b), (size_t)(n) * sizeof(*(a)))
#define PetscArrayzero(a, n) (memset((a), 0, (size_t)(n) * sizeof(*(a))), PETSC_SUCCESS)
PetscErrorCode PetscStrcmp(const char *, const char *, PetscBool *);
PetscErrorCode PetscStrallocpy(const char *, char **);
PetscErrorCode PetscInfo_Private(void *, const char *, ...);
#define PetscInfo(o, ...) PetscInfo_Private((void *)(o), __VA_ARGS__)
PetscErrorCode PetscLogEventBegin(PetscLogEvent, void *, void *, void *, void *);
PetscErrorCode PetscLogEventEnd(PetscLogEvent, void *, void *, void *, void *);
PetscErrorCode PetscLogGpuTimeBegin(void);
PetscErrorCode PetscLogGpuTimeEnd(void);
PetscErrorCode PetscLogGpuFlops(PetscLogDouble);
PetscErrorCode PetscLogFlops(PetscLogDouble);
PetscErrorCode PetscLogGpuToCpu(PetscLogDouble);
PetscErrorCode PetscLogCpuToGpu(PetscLogDouble);
PetscErrorCode PetscOptionsGetString(PetscOptions, const char *, const char *, char *, size_t, PetscBool *);
PetscErrorCode PetscOptionsGetInt(PetscOptions, const char *, const char *, PetscInt *, PetscBool *);
PetscErrorCode PetscOptionsGetBool(PetscOptions, const char *, const char *, PetscBool *, PetscBool *);
PetscErrorCode PetscIntMultError(PetscInt, PetscInt, PetscInt *);
PetscErrorCode PetscBLASIntCast(PetscInt, PetscBLASInt *);
PetscErrorCode PetscMPIIntCast(PetscInt, PetscMPIInt *);
PetscErrorCode PetscViewerASCIIPrintf(PetscViewer, const char *, ...);
PetscErrorCode PetscViewerGetFormat(PetscViewer, PetscViewerFormat *);
PetscErrorCode PetscViewerASCIIPushTab(PetscViewer);
PetscErrorCode PetscViewerASCIIPopTab(PetscViewer);
PetscErrorCode PetscViewerASCIIUseTabs(PetscViewer, PetscBool);
PetscErrorCode PetscDeviceInitialize(int);
PetscErrorCode PetscDLLibraryRegister_b200(void);
int MPI_Comm_size(MPI_Comm, int *);
int MPI_Comm_rank(MPI_Comm, int *);
int MPI_Allreduce(const void *, void *, int, MPI_Datatype, MPI_Op, MPI_Comm);
#define MPIU_Allreduce(a, b, c, d, e, f) MPI_Allreduce(a, b, c, d, e, f)
#define PETSC_VIEWER_ASCII_MATLAB 3
#define PETSC_VIEWER_ASCII_INFO 1
#define PETSC_VIEWER_ASCII_INFO_DETAIL 2
#define PETSC_CUDA 1
/* the MPI calls the adapter makes (real signatures) */
#define MPIU_INT 0
#define MPI_BYTE 0
#define MPI_STATUS_IGNORE NULL
typedef struct { int dummy; } MPI_Status;
int MPI_Allgather(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, MPI_Comm);
int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm);
int MPI_Sendrecv(const void *, int, MPI_Datatype, int, int, void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *);
