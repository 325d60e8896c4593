/* MINIMAL PETSc STAND-IN (see petscsys.h): the CUDA bits that include/slepc/private/bvimpl.h touches in a CUDA-enabled build */
#pragma once
#include <petscsys.h>
#include <petscmat.h>
typedef int PetscCuBLASInt;
typedef int cudaError_t;
typedef struct { unsigned int x, y, z; } dim3;
typedef struct CUstream_st *cudaStream_t;
enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
struct cudaDeviceProp { int maxGridSize[3]; int maxThreadsDim[3]; int maxThreadsPerBlock; };
cudaError_t cudaGetDevice(int *);
cudaError_t cudaGetDeviceProperties(struct cudaDeviceProp *, int);
cudaError_t cudaMalloc(void **, size_t);
cudaError_t cudaFree(void *);
cudaError_t cudaMemcpy(void *, const void *, size_t, int);
#define PetscCallCUDA(...) do { if (__VA_ARGS__) return PETSC_ERR_GPU; } while (0)
PetscErrorCode PetscCuBLASIntCast(PetscInt, PetscCuBLASInt *);
PetscErrorCode MatDenseCUDAGetArrayRead(Mat, const PetscScalar **);
PetscErrorCode MatDenseCUDARestoreArrayRead(Mat, const PetscScalar **);
