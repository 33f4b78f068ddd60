// Probe for the tcgen05 attention kernel (round 2).  One CTA per launch; answers, on a B200:
//  (1) what a tcgen05.mma (kind::f16, cta_group::1, M = 128, K = 16) costs as a function of N when the issue
//      loop is not the bottleneck (unrolled runs under one elect.sync), A from shared memory (SS) and from
//      tensor memory (TS);
//  (2) whether the operand forms the attention kernel wants are read correctly:
//        S = Q . K^T  with Q / K as K-major 32-byte slices INSIDE a 128B-swizzled [rows x 128 B] tile
//                     (start address advanced by 0 / 32 bytes, SBO = 1024), N = 80 and 160;
//        O = P . V    with P packed bf16 in tensor memory (TS form) and V as an MN-major 32-byte slice of the same
//                     tile (start + 64 bytes), N = 16, several (LBO, SBO) encodings tried;
//  (3) MUFU.EX2 vs a packed-f32x2 polynomial exp2 on the FMA pipe, elements per clock per SM.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I clearconverse_b200/csrc attn_probe.cu -o attn_probe
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"
using namespace resep::ptx;

// ---------------------------------------------------------------------------------------------- (1) MMA pacing
template <bool TS, int N>
__global__ void __launch_bounds__(128, 1) k_pace(int n_groups, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  constexpr uint32_t IDESC = umma_idesc(UMMA_BF16, UMMA_BF16, 128, N);
  if (warp == 1) {
    const uint64_t adesc = umma_desc_k_sw128(smem_u32(smem));
    const uint64_t bdesc = umma_desc_k_sw128(smem_u32(smem + 32768));
    long long t0 = clock64();
    for (int g = 0; g < n_groups; ++g) {
      if (elect_one()) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint32_t d = tmem + (i & 1) * 256;            // two accumulators, alternating
          if (TS) umma_bf16_ts(d, tmem + 480 + 8 * (i & 3), bdesc + 2 * (i & 3), IDESC, i >= 2);
          else umma_bf16(d, adesc + 2 * (i & 3), bdesc + 2 * (i & 3), IDESC, i >= 2);
        }
      }
      __syncwarp();
    }
    long long t1 = clock64();
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (threadIdx.x == 32) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

template <bool TS, int N>
static void pace(long long* d_out) {
  auto kern = k_pace<TS, N>;
  const int smem = 96 * 1024;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int groups = 128;
  long long h[2] = {0, 0};
  for (int rep = 0; rep < 2; ++rep) {
    kern<<<1, 128, smem>>>(groups, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("pace<%d,%d>: %s\n", (int)TS, N, cudaGetErrorString(e)); exit(1); }
  }
  cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
  printf("pace %s M=128 N=%3d: issue %6.1f cyc/mma, complete %6.1f cyc/mma (N-proportional floor %.0f)\n", TS ? "TS" : "SS", N,
         (double)h[0] / (groups * 16), (double)h[1] / (groups * 16), 128.0 * N / 256.0);
}

// ---------------------------------------------------------------------------------------------- (2) operand forms
// tile: [256 rows][64 bf16] with the 128B swizzle (16-byte chunk c of row r stored at chunk c ^ (r & 7))
struct FormArgs {
  const __nv_bfloat16* tile;   // [256][64] logical (unswizzled) values
  const __nv_bfloat16* p;      // [128][160] P values for the TS test
  float* s_out;                // [128][160]
  float* o_out;                // [8 variants][128][16]
  int q_row0;                  // first row of the A operand of the S test
  int variants;                // bit mask of the PV descriptor variants to run
};

__device__ __forceinline__ uint64_t desc_generic(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;      // SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(160, 1) k_forms(FormArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 4) tmem_alloc<512>(&tmem_slot);
  // fill the swizzled tile
  for (int i = threadIdx.x; i < 256 * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(smem + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(a.tile + r * 64 + c * 8);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t tile = smem_u32(smem);
  uint32_t phase = 0;
  // ---- S test: columns [0,160) = one N=160 MMA; columns [160,240) and [240,320) = two N=80 MMAs (keys 0-79, 80-159)
  if (warp == 4) {
    if (elect_one()) {
      const uint64_t qd = umma_desc_k_sw128(tile + a.q_row0 * 128);
      umma_bf16(tmem + 0, qd, umma_desc_k_sw128(tile + 32), umma_idesc(UMMA_BF16, UMMA_BF16, 128, 160), false);
      umma_bf16(tmem + 160, qd, umma_desc_k_sw128(tile + 32), umma_idesc(UMMA_BF16, UMMA_BF16, 128, 80), false);
      umma_bf16(tmem + 240, qd, umma_desc_k_sw128(tile + 32 + 80 * 128), umma_idesc(UMMA_BF16, UMMA_BF16, 128, 80), false);
      umma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, phase); phase ^= 1;
  tc_fence_after();
  if (warp < 4) {
    const int r = warp * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < 320; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(lane_base + c0, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) {
        const int c = c0 + j;
        if (c < 160) a.s_out[r * 160 + c] = __uint_as_float(v[j]);
        else if (a.s_out[r * 160 + c - 160] != __uint_as_float(v[j])) a.s_out[r * 160 + c - 160] = nanf("");   // N=80 halves must equal the N=160 MMA
      }
    }
    // ---- P into tensor memory: packed bf16, 80 columns at 384
    for (int c0 = 0; c0 < 80; c0 += 16) {
      uint32_t p[16];
      for (int j = 0; j < 16; ++j) p[j] = *reinterpret_cast<const uint32_t*>(a.p + r * 160 + 2 * (c0 + j));
      tmem_st16(lane_base + 384 + c0, p);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // ---- PV test: 8 descriptor variants for the MN-major V slice, accumulators at columns 16 * variant
  if (warp == 4) {
    if (elect_one()) {
      constexpr uint32_t IDESC_MN = umma_idesc(UMMA_BF16, UMMA_BF16, 128, 16) | (1u << 16);   // B is MN-major
      const uint32_t lbos[8] = {0, 1024, 128, 2048, 0, 1024, 16, 64};
      const uint32_t sbos[8] = {1024, 1024, 1024, 1024, 2048, 128, 1024, 1024};
      for (int v = 0; v < 8; ++v)
        for (int ks = 0; ks < 10 && ((a.variants >> v) & 1); ++ks)
          umma_bf16_ts(tmem + 16 * v, tmem + 384 + 8 * ks, desc_generic(tile + 64 + ks * 16 * 128, lbos[v], sbos[v]), IDESC_MN, ks > 0);
      umma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, phase); phase ^= 1;
  tc_fence_after();
  if (warp < 4) {
    const int r = warp * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    for (int v = 0; v < 8; ++v) {
      uint32_t o[16];
      tmem_ld16(lane_base + 16 * v, o);
      tmem_ld_wait();
      for (int j = 0; j < 16; ++j) a.o_out[(v * 128 + r) * 16 + j] = __uint_as_float(o[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

static void forms(int variants) {
  std::vector<__nv_bfloat16> tile(256 * 64), p(128 * 160);
  std::vector<float> tf(256 * 64), pf(128 * 160);
  srand(1);
  for (size_t i = 0; i < tile.size(); ++i) { tile[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); tf[i] = __bfloat162float(tile[i]); }
  for (size_t i = 0; i < p.size(); ++i) { p[i] = __float2bfloat16((rand() % 1000) / 1000.f); pf[i] = __bfloat162float(p[i]); }
  __nv_bfloat16 *d_tile, *d_p; float *d_s, *d_o;
  cudaMalloc(&d_tile, tile.size() * 2); cudaMalloc(&d_p, p.size() * 2); cudaMalloc(&d_s, 128 * 160 * 4); cudaMalloc(&d_o, 8 * 128 * 16 * 4);
  cudaMemcpy(d_tile, tile.data(), tile.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(d_p, p.data(), p.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k_forms, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int q_row0 : {0, 96}) {
    FormArgs a{d_tile, d_p, d_s, d_o, q_row0, variants};
    cudaMemset(d_s, 0, 128 * 160 * 4); cudaMemset(d_o, 0, 8 * 128 * 16 * 4);
    k_forms<<<1, 160, 64 * 1024>>>(a);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("forms: %s\n", cudaGetErrorString(e)); exit(1); }
    std::vector<float> s(128 * 160), o(8 * 128 * 16);
    cudaMemcpy(s.data(), d_s, s.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(o.data(), d_o, o.size() * 4, cudaMemcpyDeviceToHost);
    double es = 0; int nan_s = 0;
    for (int r = 0; r < 128; ++r)
      for (int k = 0; k < 160; ++k) {
        double ref = 0;
        for (int e2 = 0; e2 < 16; ++e2) ref += (double)tf[(q_row0 + r) * 64 + e2] * tf[k * 64 + 16 + e2];
        if (std::isnan(s[r * 160 + k])) ++nan_s; else es = fmax(es, fabs(ref - s[r * 160 + k]));
      }
    printf("forms q_row0=%3d: S = Q.K^T (K-major slices at +0 / +32 B inside the SW128 tile): max err %.3e, N=80 halves != N=160 at %d entries\n", q_row0, es, nan_s);
    for (int v = 0; v < 8; ++v) {
      if (!((variants >> v) & 1)) continue;
      double eo = 0;
      for (int r = 0; r < 128; ++r)
        for (int j = 0; j < 16; ++j) {
          double ref = 0;
          for (int k = 0; k < 160; ++k) ref += (double)pf[r * 160 + k] * tf[k * 64 + 32 + j];
          eo = fmax(eo, fabs(ref - o[(v * 128 + r) * 16 + j]));
        }
      const unsigned lbos[8] = {0, 1024, 128, 2048, 0, 1024, 16, 64}, sbos[8] = {1024, 1024, 1024, 1024, 2048, 128, 1024, 1024};
      printf("   O = P.V (P in TMEM, V MN-major slice at +64 B, LBO %4u SBO %4u): max err %.3e %s\n", lbos[v], sbos[v], eo, eo < 1e-3 ? "OK" : "");
    }
  }
}

// ---------------------------------------------------------------------------------------------- (3) exp2 throughput
__device__ __forceinline__ float2 ex2_poly2(float2 x) {   // x <= 0; degree-3 minimax of 2^f on [-0.5, 0.5], packed f32x2
  const float2 magic = make_float2(12582912.f, 12582912.f);
  x.x = fmaxf(x.x, -125.f); x.y = fmaxf(x.y, -125.f);
  const float2 r = fadd2(x, magic);                      // low mantissa bits = round(x)
  const float2 xi = fadd2(r, make_float2(-12582912.f, -12582912.f));
  const float2 f = fadd2(x, make_float2(-xi.x, -xi.y));
  float2 p = ffma2(f, make_float2(0.05517167f, 0.05517167f), make_float2(0.24261113f, 0.24261113f));
  p = ffma2(p, f, make_float2(0.69326097f, 0.69326097f));
  p = ffma2(p, f, make_float2(0.99992806f, 0.99992806f));
  p.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(r.x) << 23));
  p.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(r.y) << 23));
  return p;
}
template <int MODE>
__global__ void k_exp(float* out, int iters, long long* cyc) {
  float2 v[8];
  for (int i = 0; i < 8; ++i) v[i] = make_float2(-0.001f * (threadIdx.x + i), -0.002f * (threadIdx.x + i));
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0 || (MODE == 2 && (i & 1)) || (MODE == 3 && (i & 3) == 3)) v[i] = make_float2(ex2_approx(v[i].x) - 1.5f, ex2_approx(v[i].y) - 1.5f);
      else { const float2 p = ex2_poly2(v[i]); v[i] = make_float2(p.x - 1.5f, p.y - 1.5f); }
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += v[i].x + v[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
static void exps() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  const char* names[4] = {"MUFU.EX2 only", "packed polynomial only", "half MUFU / half polynomial", "1/4 MUFU, 3/4 polynomial"};
  for (int mode = 0; mode < 4; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) k_exp<0><<<148, 512>>>(out, iters, cyc);
      if (mode == 1) k_exp<1><<<148, 512>>>(out, iters, cyc);
      if (mode == 2) k_exp<2><<<148, 512>>>(out, iters, cyc);
      if (mode == 3) k_exp<3><<<148, 512>>>(out, iters, cyc);
      cudaDeviceSynchronize();
    }
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("exp2 %-30s %.2f per clock per SM (16 warps/SM, %lld cycles)\n", names[mode], 512.0 * 16 * iters / (double)c, c);
  }
  // accuracy of the polynomial on [-30, 0]
  double worst = 0;
  for (int i = 0; i <= 300000; ++i) {
    const float x = -i * 1e-4f;
    const float xr = nearbyintf(x), f = x - xr;
    float p = fmaf(f, 0.05517167f, 0.24261113f); p = fmaf(p, f, 0.69326097f); p = fmaf(p, f, 0.99992806f);
    const double got = ldexp((double)p, (int)xr), ref = exp2((double)x);
    worst = fmax(worst, fabs(got - ref) / ref);
  }
  printf("exp2 polynomial (degree 3 on [-0.5, 0.5]): worst relative error %.3e (bf16 half-ulp is 1.95e-03)\n", worst);
}

int main(int argc, char** argv) {
  // a wrong descriptor may fault and poison the context: every section / variant can be run as its own process
  const char* what = argc > 1 ? argv[1] : "all";
  long long* d_out;
  cudaMalloc(&d_out, 64);
  if (!strcmp(what, "pace") || !strcmp(what, "all")) {
    pace<false, 16>(d_out); pace<false, 32>(d_out); pace<false, 64>(d_out); pace<false, 80>(d_out); pace<false, 128>(d_out);
    pace<false, 160>(d_out); pace<false, 256>(d_out);
    pace<true, 16>(d_out); pace<true, 32>(d_out); pace<true, 64>(d_out); pace<true, 128>(d_out);
  }
  if (!strcmp(what, "forms") || !strcmp(what, "all")) forms(argc > 2 ? 1 << atoi(argv[2]) : 0xFF);
  if (!strcmp(what, "exps") || !strcmp(what, "all")) exps();
  return 0;
}
