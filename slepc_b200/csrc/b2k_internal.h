/* internal declarations shared by the .cu files of libb200krylov.so */
#ifndef B2K_INTERNAL_H
#define B2K_INTERNAL_H
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "b2k.h"

struct b2k_ctx_s {
  int          device;
  int          sm_count;
  cudaStream_t stream;
  double      *partials;        /* scratch for two-stage reductions: [B2K_MAX_PART_BLOCKS][pstride] */
  size_t       partials_elems;
  double      *dscratch;        /* small device scratch (host scalars/vectors staged to HBM)       */
  size_t       dscratch_elems;
  cudaEvent_t  ev0, ev1;
  uint64_t     launches;
};

#define B2K_MAX_PART_BLOCKS 2048
#define B2K_MAX_K           1024     /* max columns in one reduction (ncv+1 <= 1024)              */

void b2k_set_error(const char *fmt, ...);

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      b2k_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));    \
      return B2K_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

#define CKLAUNCH(ctx)                                                                         \
  do {                                                                                        \
    (ctx)->launches++;                                                                        \
    cudaError_t e_ = cudaGetLastError();                                                      \
    if (e_ != cudaSuccess) {                                                                  \
      b2k_set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_));\
      return B2K_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

#define ARGCHK(cond, msg)                                                                     \
  do {                                                                                        \
    if (!(cond)) {                                                                            \
      b2k_set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, msg);                  \
      return B2K_ERR_ARG;                                                                     \
    }                                                                                         \
  } while (0)

static inline int b2k_is_aligned16(const void *p) { return (((uintptr_t)p) & 15u) == 0; }

#endif
