// Microbenchmark: which issue pipes do the softmax inner-loop instructions share on sm_100a?  Per-SM throughput of
// MUFU.EX2, F2FP.BF16.F32.PACK_AB (cvt.rn.bf16x2.f32), a PRMT-based truncating pack and SHFL.BFLY, alone and mixed.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t f2fp(float a, float b) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ uint32_t prmt_pack(float a, float b) { uint32_t r; asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(__float_as_uint(b)), "r"(__float_as_uint(a))); return r; }
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  float v[8]; uint32_t u[4] = {0, 0, 0, 0};
  for (int i = 0; i < 8; ++i) v[i] = -0.001f * (threadIdx.x + i);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2 || MODE == 4) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = ex2(v[i]);
    }
    if (MODE == 1 || MODE == 2) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { u[i] ^= f2fp(v[2 * i], v[2 * i + 1]); }
#pragma unroll
      for (int i = 0; i < 4; ++i) { u[i] ^= f2fp(v[2 * i + 1], v[2 * i]); }
    }
    if (MODE == 3 || MODE == 4) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { u[i] ^= prmt_pack(v[2 * i], v[2 * i + 1]); }
#pragma unroll
      for (int i = 0; i < 4; ++i) { u[i] ^= prmt_pack(v[2 * i + 1], v[2 * i]); }
    }
    if (MODE == 5) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __shfl_xor_sync(0xffffffffu, v[i], 1 + (i & 3));
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += v[i];
  for (int i = 0; i < 4; ++i) s += __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  const char* names[6] = {"8 MUFU.EX2", "8 F2FP", "8 MUFU.EX2 + 8 F2FP", "8 PRMT pack (+8 LOP)", "8 MUFU.EX2 + 8 PRMT pack", "8 SHFL.BFLY"};
  for (int mode = 0; mode < 6; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) k<0><<<148 * 2, 1024>>>(out, iters, cyc);
      if (mode == 1) k<1><<<148 * 2, 1024>>>(out, iters, cyc);
      if (mode == 2) k<2><<<148 * 2, 1024>>>(out, iters, cyc);
      if (mode == 3) k<3><<<148 * 2, 1024>>>(out, iters, cyc);
      if (mode == 4) k<4><<<148 * 2, 1024>>>(out, iters, cyc);
      if (mode == 5) k<5><<<148 * 2, 1024>>>(out, iters, cyc);
      cudaDeviceSynchronize();
    }
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    // per SM: 64 warps, each issuing the listed instructions per iteration
    printf("%-28s %.1f cycles per warp-iteration per SMSP (%lld cycles total)\n", names[mode], (double)c / iters / 16.0, c);
  }
  return 0;
}
