/* context, memory and timing entry points of libb200krylov.so (see include/b2k.h) */
#include <stdarg.h>
#include <string.h>
#include "b2k_internal.h"

static thread_local char g_err[512] = "";

void b2k_set_error(const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char *b2k_last_error(void) { return g_err; }
extern "C" int b2k_version(void) { return 100; }

extern "C" int b2k_device_count(int *count)
{
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { n = 0; (void)cudaGetLastError(); }
  *count = n;
  return B2K_OK;
}

extern "C" int b2k_ctx_create(int device, b2k_ctx *out)
{
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    (void)cudaGetLastError();
    b2k_set_error("no CUDA device visible: libb200krylov has no CPU fallback");
    return B2K_ERR_NOGPU;
  }
  ARGCHK(device >= 0 && device < n, "device index out of range");
  CK(cudaSetDevice(device));
  b2k_ctx c = (b2k_ctx)calloc(1, sizeof(*c));
  if (!c) return B2K_ERR_MEM;
  c->device = device;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  c->partials_elems = (size_t)B2K_MAX_PART_BLOCKS * (B2K_MAX_K + 8);
  CK(cudaMalloc(&c->partials, c->partials_elems * sizeof(double)));
  c->dscratch_elems = 1 << 16;
  CK(cudaMalloc(&c->dscratch, c->dscratch_elems * sizeof(double)));
  CK(cudaEventCreate(&c->ev0));
  CK(cudaEventCreate(&c->ev1));
  *out = c;
  return B2K_OK;
}

extern "C" int b2k_ctx_destroy(b2k_ctx c)
{
  if (!c) return B2K_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  cudaFree(c->partials);
  cudaFree(c->dscratch);
  if (c->prof_ev) { for (int i = 0; i < 2 * c->prof_cap; i++) cudaEventDestroy(c->prof_ev[i]); free(c->prof_ev); free(c->prof_id); free(c->prof_bytes); free(c->prof_flop); }
  cudaEventDestroy(c->ev0);
  cudaEventDestroy(c->ev1);
  cudaStreamDestroy(c->stream);
  free(c);
  return B2K_OK;
}

extern "C" int b2k_ctx_sync(b2k_ctx c) { c->syncs++; CK(cudaStreamSynchronize(c->stream)); return B2K_OK; }
extern "C" int b2k_ctx_syncs(b2k_ctx c, uint64_t *n) { *n = c->syncs; return B2K_OK; }
extern "C" void *b2k_ctx_stream(b2k_ctx c) { return (void *)c->stream; }
extern "C" int b2k_ctx_sm_count(b2k_ctx c) { return c->sm_count; }
extern "C" int b2k_ctx_launches(b2k_ctx c, uint64_t *n) { *n = c->launches; return B2K_OK; }
extern "C" int b2k_ctx_copy_bytes(b2k_ctx c, uint64_t *h2d, uint64_t *d2h)
{
  if (h2d) *h2d = c->h2d_bytes;
  if (d2h) *d2h = c->d2h_bytes;
  return B2K_OK;
}

#define B2K_PROF_CAP 16384
static int prof_flush(b2k_ctx c)
{
  if (c->prof_n == 0) return B2K_OK;
  CK(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < c->prof_n; i++) {
    float f = 0.f;
    CK(cudaEventElapsedTime(&f, c->prof_ev[2 * i], c->prof_ev[2 * i + 1]));
    const int id = c->prof_id[i];
    c->prof_ms[id] += (double)f;
    c->prof_b[id] += c->prof_bytes[i];
    c->prof_f[id] += c->prof_flop[i];
    c->prof_cnt[id]++;
  }
  c->prof_n = 0;
  return B2K_OK;
}
extern "C" int b2k_prof_enable(b2k_ctx c, int on)
{
  if (on) {
    if (!c->prof_ev) {
      c->prof_cap = B2K_PROF_CAP;
      c->prof_ev = (cudaEvent_t *)calloc(2 * (size_t)c->prof_cap, sizeof(cudaEvent_t));
      c->prof_id = (int *)calloc((size_t)c->prof_cap, sizeof(int));
      c->prof_bytes = (double *)calloc((size_t)c->prof_cap, sizeof(double));
      c->prof_flop = (double *)calloc((size_t)c->prof_cap, sizeof(double));
      if (!c->prof_ev || !c->prof_id || !c->prof_bytes || !c->prof_flop) return B2K_ERR_MEM;
      for (int i = 0; i < 2 * c->prof_cap; i++) CK(cudaEventCreate(&c->prof_ev[i]));
    }
    c->prof_n = 0;
    for (int i = 0; i < B2K_PROF_NCLASS; i++) { c->prof_ms[i] = 0.0; c->prof_b[i] = 0.0; c->prof_f[i] = 0.0; c->prof_cnt[i] = 0; }
    c->prof_on = 1;
  } else {
    int rc = prof_flush(c);
    c->prof_on = 0;
    return rc;
  }
  return B2K_OK;
}
extern "C" int b2k_prof_get(b2k_ctx c, int cls, uint64_t *launches, double *ms, double *bytes)
{
  ARGCHK(cls >= 0 && cls < B2K_PROF_NCLASS, "kernel class out of range");
  int rc = prof_flush(c);
  if (rc) return rc;
  if (launches) *launches = c->prof_cnt[cls];
  if (ms) *ms = c->prof_ms[cls];
  if (bytes) *bytes = c->prof_b[cls];
  return B2K_OK;
}

extern "C" int b2k_prof_get_flops(b2k_ctx c, int cls, double *flops)
{
  ARGCHK(cls >= 0 && cls < B2K_PROF_NCLASS, "kernel class out of range");
  int rc = prof_flush(c);
  if (rc) return rc;
  *flops = c->prof_f[cls];
  return B2K_OK;
}

extern "C" int b2k_malloc(b2k_ctx c, void **p, size_t bytes)
{
  CK(cudaSetDevice(c->device));
  if (bytes == 0) bytes = 16;
  CK(cudaMalloc(p, bytes));
  return B2K_OK;
}
extern "C" int b2k_free(b2k_ctx c, void *p)
{
  if (!p) return B2K_OK;
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaFree(p));
  return B2K_OK;
}
extern "C" int b2k_memset0(b2k_ctx c, void *p, size_t bytes) { CK(cudaMemsetAsync(p, 0, bytes, c->stream)); return B2K_OK; }
extern "C" int b2k_h2d(b2k_ctx c, void *dst, const void *src, size_t bytes)
{
  c->h2d_bytes += bytes;
  c->syncs++;
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return B2K_OK;
}
extern "C" int b2k_d2h(b2k_ctx c, void *dst, const void *src, size_t bytes)
{
  c->d2h_bytes += bytes;
  c->syncs++;
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return B2K_OK;
}
extern "C" int b2k_h2d_async(b2k_ctx c, void *dst, const void *src, size_t bytes)
{
  c->h2d_bytes += bytes;
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  return B2K_OK;
}
extern "C" int b2k_d2h_async(b2k_ctx c, void *dst, const void *src, size_t bytes)
{
  c->d2h_bytes += bytes;
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  return B2K_OK;
}
extern "C" int b2k_d2d(b2k_ctx c, void *dst, const void *src, size_t bytes)
{
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, c->stream));
  return B2K_OK;
}
extern "C" int b2k_host_alloc(void **p, size_t bytes) { CK(cudaHostAlloc(p, bytes ? bytes : 16, cudaHostAllocDefault)); return B2K_OK; }
extern "C" int b2k_host_free(void *p) { if (p) CK(cudaFreeHost(p)); return B2K_OK; }
extern "C" int b2k_mem_info(b2k_ctx c, size_t *f, size_t *t)
{
  CK(cudaSetDevice(c->device));
  CK(cudaMemGetInfo(f, t));
  return B2K_OK;
}
extern "C" int b2k_timer_start(b2k_ctx c) { CK(cudaEventRecord(c->ev0, c->stream)); return B2K_OK; }
extern "C" int b2k_timer_stop_ms(b2k_ctx c, double *ms)
{
  float f = 0.f;
  CK(cudaEventRecord(c->ev1, c->stream));
  CK(cudaEventSynchronize(c->ev1));
  CK(cudaEventElapsedTime(&f, c->ev0, c->ev1));
  *ms = (double)f;
  return B2K_OK;
}
