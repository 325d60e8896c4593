// FP64 throughput probe on B200: DFMA (vector pipe) vs mma.sync.m8n8k4.f64 (legacy tensor path; tcgen05 has no FP64 kind).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_probe tools/dmma_probe.cu && tools/dmma_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double *out, int iters)
{
  double a[16], x = 1.0 + threadIdx.x * 1e-9, y = 0.999999;
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = fma(a[i], x, y);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma(double *out, int iters)
{
  double c[8][2], a = 1.0 + threadIdx.x * 1e-9, b = 0.5;
#pragma unroll
  for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
  double *out;
  cudaMalloc(&out, 148 * 8 * 256 * sizeof(double));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, grid = 148 * 4, block = 256;
  float ms;
  for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(e0); k_dfma<<<grid, block>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep) printf("DFMA : %.2f TFLOP/s (%.3f ms)\n", 2.0 * 16 * iters * (double)grid * block / ms / 1e9, ms);
    cudaEventRecord(e0); k_dmma<<<grid, block>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep) printf("DMMA : %.2f TFLOP/s (%.3f ms)\n", 2.0 * 8 * 256 * iters * (double)grid * (block / 32) / ms / 1e9, ms);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
