// Microbenchmark: tcgen05.ld / tcgen05.st throughput per SM as a function of the number of warps per sub-partition.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I clearconverse_b200/csrc tmem_bw.cu -o tmem_bw
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"
using namespace resep::ptx;

template <int MODE>   // 0: ld x32 + wait each; 1: 4 x ld x32 in flight then wait; 2: st x32 + wait
__global__ void __launch_bounds__(512, 1) k_bw(int iters, long long* out, float* sink) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t v[32], acc = 0;
  for (int j = 0; j < 32; ++j) v[j] = threadIdx.x + j;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) { tmem_ld32(tmem + 32 * ((i + warp) & 7), v); tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 8) acc += v[j]; }
    if (MODE == 1) {
      uint32_t a[32], b[32], c[32], d[32];
      tmem_ld32(tmem, a); tmem_ld32(tmem + 32, b); tmem_ld32(tmem + 64, c); tmem_ld32(tmem + 96, d);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 8) acc += a[j] + b[j] + c[j] + d[j];
    }
    if (MODE == 2) { tmem_st32(tmem + 32 * ((i + warp) & 7), v); tmem_st_wait(); }
  }
  const long long t1 = clock64();
  __syncthreads();
  sink[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem_slot); }
}

int main() {
  long long* d_out; float* sink;
  cudaMalloc(&d_out, 8); cudaMalloc(&sink, 148 * 512 * 4);
  const int iters = 2000;
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {4, 8, 16}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k_bw<0><<<148, warps * 32>>>(iters, d_out, sink);
        if (mode == 1) k_bw<1><<<148, warps * 32>>>(iters, d_out, sink);
        if (mode == 2) k_bw<2><<<148, warps * 32>>>(iters, d_out, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      }
      long long c; cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
      const double bytes = (double)iters * warps * 32 * 32 * 4 * (mode == 1 ? 4 : 1);
      printf("%s, %2d warps per SM (%d per sub-partition): %.1f bytes per clock per SM, %.0f cycles per warp-instruction\n",
             mode == 0 ? "tcgen05.ld x32 + wait      " : mode == 1 ? "4 x tcgen05.ld x32, one wait" : "tcgen05.st x32 + wait      ", warps, warps / 4,
             bytes / c, (double)c / iters / (mode == 1 ? 4 : 1));
    }
  return 0;
}
