// Microbenchmark: issue rate of the legacy warp-level HMMA.16816.F32.BF16 (mma.sync m16n8k16) on sm_100a, per SM
// sub-partition, with 1..8 resident warps per sub-partition and 4 independent accumulator chains per warp; alone and
// with 8 MUFU.EX2 per 5 HMMAs (the attention inner loop's mix).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void hmma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  float d[5][4]; uint32_t a[4] = {0x3c003c00u + threadIdx.x, 0x3c003c01u, 0x3c003c02u, 0x3c003c03u};
  float v[8], f[8];
  for (int i = 0; i < 8; ++i) f[i] = 0.001f * (threadIdx.x + i);
  for (int i = 0; i < 8; ++i) v[i] = -0.001f * (threadIdx.x + i);
  for (int j = 0; j < 5; ++j) for (int i = 0; i < 4; ++i) d[j][i] = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 5; ++j) hmma(d[j], a, a[j & 3], a[(j + 1) & 3]);
    constexpr int NMUFU = MODE & 15, NFMA = MODE >> 4;
#pragma unroll
    for (int i = 0; i < NMUFU; ++i) v[i] = ex2(v[i]);
#pragma unroll
    for (int i = 0; i < NFMA; ++i) f[i & 7] = fmaf(f[i & 7], 1.0001f, 0.5f);
  }
  long long t1 = clock64();
  float s = 0; for (int j = 0; j < 5; ++j) for (int i = 0; i < 4; ++i) s += d[j][i];
  for (int i = 0; i < 8; ++i) s += v[i] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE> void run(float* out, long long* cyc, const char* name) {
  const int iters = 2000;
  for (int wps = 2; wps <= 8; wps *= 2) {
    for (int rep = 0; rep < 2; ++rep) { k<MODE><<<148, 128 * wps>>>(out, iters, cyc); cudaDeviceSynchronize(); }
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-34s %d warps/SMSP: %.1f cycles per block per SMSP\n", name, wps, (double)c / iters / wps);
  }
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4 * 4); cudaMalloc(&cyc, 8);
  run<0>(out, cyc, "5 HMMA");
  run<8>(out, cyc, "5 HMMA + 8 MUFU");
  run<4>(out, cyc, "5 HMMA + 4 MUFU");
  run<16 * 16>(out, cyc, "5 HMMA + 16 FFMA");
  run<32 * 16>(out, cyc, "5 HMMA + 32 FFMA");
  run<64 * 16>(out, cyc, "5 HMMA + 64 FFMA");
  run<8 + 16 * 16>(out, cyc, "5 HMMA + 8 MUFU + 16 FFMA");
  run<8 + 32 * 16>(out, cyc, "5 HMMA + 8 MUFU + 32 FFMA");
  run<4 + 32 * 16>(out, cyc, "5 HMMA + 4 MUFU + 32 FFMA");
  run<0 + 64 * 16>(out, cyc, "5 HMMA + 0 MUFU + 64 FFMA");
  return 0;
}
