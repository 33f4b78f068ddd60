// Microbenchmark 2: cost of the ISSUE pattern around tcgen05.mma.  Per "unit": wait on an (already complete)
// mbarrier, tcgen05.fence::after_thread_sync, U MMAs (cta_group::2, A from TMEM, N columns), one commit.
// STYLE 0: the role code runs under `if (lane == 0)`;  STYLE 1: the whole warp runs the loop, only the MMAs and
// the commit are issued by the elect.sync thread (CUTLASS style).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"
using namespace resep::ptx;

template <int STYLE, int N, int U>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k_bench(int n_units, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[16];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc_pair<512>(&tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  constexpr uint32_t IDESC = umma_idesc(UMMA_BF16, UMMA_BF16, 256, N);
  if (threadIdx.x == 0) for (int i = 0; i < 8; ++i) mbar_arrive(&bar[i]);   // "full" barriers: phase 0 complete forever
  __syncthreads();
  if (warp == 1 && rank == 0) {
    long long t0 = 0, t1 = 0, t2 = 0;
    if (STYLE == 0) {
      if (lane == 0) {
        t0 = clock64();
        int st = 0;
        for (int u = 0; u < n_units; ++u) {
          mbar_wait(&bar[st], 0);
          tc_fence_after();
          const uint64_t bdesc = umma_desc_k_sw128(smem_u32(smem + st * 8192));
          const uint32_t d = tmem + (u & 1) * 256;
#pragma unroll
          for (int k = 0; k < U; ++k) umma_bf16_ts_pair(d, tmem + 480 + 8 * (k & 3), bdesc + 2 * (k & 3), IDESC, k != 0);
          umma_commit_pair(&bar[8 + st]);
          if (++st == 8) st = 0;
        }
        t1 = clock64();
        umma_commit_pair(&bar[8]);
      }
      __syncwarp();
    } else {
      t0 = clock64();
      int st = 0;
      for (int u = 0; u < n_units; ++u) {
        mbar_wait(&bar[st], 0);
        tc_fence_after();
        const uint64_t bdesc = umma_desc_k_sw128(smem_u32(smem + st * 8192));
        const uint32_t d = tmem + (u & 1) * 256;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < U; ++k) umma_bf16_ts_pair(d, tmem + 480 + 8 * (k & 3), bdesc + 2 * (k & 3), IDESC, k != 0);
          umma_commit_pair(&bar[8 + st]);
        }
        __syncwarp();
        if (++st == 8) st = 0;
      }
      t1 = clock64();
      if (elect_one()) umma_commit_pair(&bar[8]);
      __syncwarp();
    }
    // wait for everything: bar[8] got n_units/8 (+1) arrivals; just spin on the clock long enough instead
    long long tw = clock64();
    while (clock64() - tw < 2000000) {}
    t2 = clock64();
    if (lane == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) { tc_fence_after(); tmem_dealloc_pair<512>(tmem); }
}

template <int STYLE, int N, int U>
void run(long long* d_out) {
  auto kern = k_bench<STYLE, N, U>;
  const int smem = 96 * 1024;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int n_units = 4096 / U;
  long long h[2];
  for (int rep = 0; rep < 2; ++rep) {
    kern<<<2, 128, smem>>>(n_units, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("sync: %s\n", cudaGetErrorString(e)); exit(1); }
  }
  cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
  printf("style %d (%s) N=%3d U=%2d: issue loop %6.1f cyc/mma (floor %d)\n", STYLE, STYLE ? "elect    " : "lane == 0", N, U,
         (double)h[0] / (n_units * U), N / 2 < 64 ? 64 : N / 2);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 64);
  run<0, 128, 4>(d_out); run<1, 128, 4>(d_out);
  run<0, 128, 8>(d_out); run<1, 128, 8>(d_out);
  run<0, 256, 4>(d_out); run<1, 256, 4>(d_out);
  run<0, 256, 8>(d_out); run<1, 256, 8>(d_out);
  run<0, 64, 8>(d_out);  run<1, 64, 8>(d_out);
  return 0;
}
