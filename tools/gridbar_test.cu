// Microbenchmark: cost of a software grid barrier among 148 co-resident CTAs (sense-reversing, one thread per CTA).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void grid_barrier(unsigned* count, unsigned* gen, unsigned nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned g = ld_acquire(gen);
    __threadfence();
    if (atomicAdd(count, 1) == nblocks - 1) {
      *count = 0;
      __threadfence();
      atomicAdd(gen, 1);
    } else {
      while (ld_acquire(gen) == g) {}
    }
  }
  __syncthreads();
}
__global__ void bar_kernel(unsigned* count, unsigned* gen, int iters, float* sink) {
  float acc = 0.f;
  for (int i = 0; i < iters; ++i) {
    grid_barrier(count, gen, gridDim.x);
    acc += 1.f;
  }
  if (threadIdx.x == 0) sink[blockIdx.x] = acc;
}
int main() {
  unsigned* c; float* sink;
  cudaMalloc(&c, 8); cudaMemset(c, 0, 8); cudaMalloc(&sink, 1024);
  for (int threads : {32, 192}) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      const int iters = 2000;
      cudaEventRecord(e0);
      bar_kernel<<<148, threads, 200 * 1024 * 0>>>(c, c + 1, iters, sink);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("threads=%d: %s  %.3f us per grid barrier\n", threads, cudaGetErrorString(e), ms * 1e3 / iters);
    }
  }
  return 0;
}
