/* b2k_petsc.c — process-wide glue of the PETSc-typed plug-ins: the b2k context of this rank's GPU and one row-partition
   communicator per MPI communicator.  The NCCL unique id (128 bytes) is broadcast and the CUDA-IPC handles of the reduction
   mailboxes (64 bytes per rank) are all-gathered with MPI — what slepc_b200/dist.py does with torch.distributed for the
   stand-alone host layer.  Type-checked only (no PETSc / MPI in this image). */
#include <petscsys.h>
#include "b2k.h"
#include "b2k_petsc.h"

#define B2KCall(call) do { int rc_ = (call); PetscCheck(!rc_, PETSC_COMM_SELF, PETSC_ERR_GPU, "%s failed (%d): %s", #call, rc_, b2k_last_error()); } while (0)

static b2k_ctx g_ctx = NULL;
static struct { MPI_Comm comm; b2k_comm c; } g_comms[16];
static int g_ncomms = 0;

b2k_ctx B2KPetscContext(void) { return g_ctx; }

PetscErrorCode B2KPetscInitialize(MPI_Comm comm)
{
  int ndev = 0, rank = 0;

  PetscFunctionBegin;
  if (g_ctx) PetscFunctionReturn(PETSC_SUCCESS);
  B2KCall(b2k_device_count(&ndev));
  PetscCheck(ndev > 0, comm, PETSC_ERR_GPU, "no CUDA device: the b200 plug-ins have no CPU fallback");
  PetscCallMPI(MPI_Comm_rank(comm, &rank));
  B2KCall(b2k_ctx_create(rank % ndev, &g_ctx));                   /* one process per GPU of the box */
  PetscFunctionReturn(PETSC_SUCCESS);
}

PetscErrorCode B2KPetscCommGet(MPI_Comm comm, b2k_comm *c)
{
  int  size, rank, ok = 1, allok = 0;
  char id[B2K_COMM_ID_BYTES], handle[B2K_COMM_P2P_HANDLE_BYTES], *all;

  PetscFunctionBegin;
  for (int i = 0; i < g_ncomms; i++) if (g_comms[i].comm == comm) { *c = g_comms[i].c; PetscFunctionReturn(PETSC_SUCCESS); }
  PetscCheck(g_ncomms < 16, comm, PETSC_ERR_SUP, "too many communicators");
  PetscCallMPI(MPI_Comm_size(comm, &size));
  PetscCallMPI(MPI_Comm_rank(comm, &rank));
  memset(id, 0, sizeof(id));
  if (rank == 0 && size > 1) B2KCall(b2k_comm_unique_id(id));
  if (size > 1) PetscCallMPI(MPI_Bcast(id, B2K_COMM_ID_BYTES, MPI_BYTE, 0, comm));
  B2KCall(b2k_comm_create(g_ctx, rank, size, id, c));
  if (size > 1 && size <= 8) {                                   /* NVLink mailboxes for the k-vector reductions; every rank must agree */
    PetscCall(PetscMalloc1((size_t)size * B2K_COMM_P2P_HANDLE_BYTES, &all));
    if (b2k_comm_p2p_handle(*c, handle)) ok = 0;
    PetscCallMPI(MPI_Allgather(handle, B2K_COMM_P2P_HANDLE_BYTES, MPI_BYTE, all, B2K_COMM_P2P_HANDLE_BYTES, MPI_BYTE, comm));
    PetscCallMPI(MPI_Allreduce(&ok, &allok, 1, MPIU_INT, MPIU_MIN, comm));
    if (allok) {
      ok = b2k_comm_p2p_open(*c, all) ? 0 : 1;
      PetscCallMPI(MPI_Allreduce(&ok, &allok, 1, MPIU_INT, MPIU_MIN, comm));
      if (!allok) B2KCall(b2k_comm_p2p_close(*c));
    }
    PetscCall(PetscFree(all));
  }
  g_comms[g_ncomms].comm = comm; g_comms[g_ncomms++].c = *c;
  PetscFunctionReturn(PETSC_SUCCESS);
}
