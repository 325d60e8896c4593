/* what SLEPc's configure would generate for a real, double-precision, CUDA-enabled build (stand-in, see petscsys.h) */
#pragma once
#define SLEPC_PETSC_DIR "stub"
#define SLEPC_PETSC_ARCH "stub"
#define SLEPC_DIR "stub"
#define SLEPC_LIB_DIR "stub"
