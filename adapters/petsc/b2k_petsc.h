/* glue shared by the PETSc-typed plug-ins (bvb200_petsc.c, matb200_petsc.c): one b2k context per process (= per GPU) and one
   row-partition communicator per MPI communicator, built the way slepc_b200/dist.py builds them for the stand-alone host layer —
   the NCCL unique id and the 64-byte CUDA-IPC handles of the reduction mailboxes travel through MPI instead of torch.distributed */
#pragma once
#include <petscmat.h>
#include "b2k.h"

PETSC_EXTERN PetscErrorCode B2KPetscInitialize(MPI_Comm comm);          /* idempotent: picks the device with rank % device count */
PETSC_EXTERN b2k_ctx        B2KPetscContext(void);
PETSC_EXTERN PetscErrorCode B2KPetscCommGet(MPI_Comm comm, b2k_comm *c); /* cached per communicator; NULL-equivalent for one rank */
PETSC_EXTERN PetscErrorCode MatCreate_B200CSR(Mat A);                    /* MatRegister("b200csr", …) constructor                   */
/* convenience: an operator for EPSSetOperators / SVDSetOperators from any assembled PETSc matrix (rows are read with MatGetRow
   once, at set-up): a MatShell with MATOP_MULT / MATOP_MULT_TRANSPOSE = the sm_100a SpMV and CUDA vectors */
PETSC_EXTERN PetscErrorCode MatCreateB200FromMat(Mat Ain, Mat *Aout);
