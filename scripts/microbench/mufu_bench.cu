// Microbenchmark: MUFU.EX2 throughput per SM (ex2.approx.ftz.f32), alone and mixed with FFMA, vs a degree-3
// polynomial exp2 on the FMA pipe (Cody-Waite split: 2^x = 2^floor(x) * p(x - floor(x))).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_poly(float x) {   // x <= 0
  x = fmaxf(x, -126.f);
  const float fl = floorf(x);
  const float f = x - fl;                               // [0, 1)
  float p = fmaf(f, 0.0555041f, 0.2402265f);            // minimax-ish cubic for 2^f
  p = fmaf(p, f, 0.6931472f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + ((int)fl << 23));
}
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  float v[8];
  for (int i = 0; i < 8; ++i) v[i] = -0.001f * (threadIdx.x + i);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) v[i] = ex2(v[i]) - 1.5f;
      else if (MODE == 1) v[i] = ex2_poly(v[i]) - 1.5f;
      else v[i] = (i & 1) ? ex2(v[i]) - 1.5f : ex2_poly(v[i]) - 1.5f;
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  const char* names[3] = {"MUFU.EX2 only", "polynomial only", "half MUFU / half polynomial"};
  for (int mode = 0; mode < 3; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) k<0><<<148 * 2, 1024>>>(out, iters, cyc);
      if (mode == 1) k<1><<<148 * 2, 1024>>>(out, iters, cyc);
      if (mode == 2) k<2><<<148 * 2, 1024>>>(out, iters, cyc);
      cudaDeviceSynchronize();
    }
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    // per SM: 2 CTAs x 1024 threads x 8 exps per iteration
    printf("%-30s %.2f exp2 per clock per SM (%lld cycles)\n", names[mode], 2.0 * 1024 * 8 * iters / (double)c, c);
  }
  return 0;
}
