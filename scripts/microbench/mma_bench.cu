// Microbenchmark: pacing of tcgen05.mma (kind::f16, bf16) as a function of cta_group, operand source of A
// (shared memory descriptor vs tensor memory), N, and commit frequency.  One CTA (or CTA pair) per SM, one
// issuing thread, operands are whatever is in shared / tensor memory (values do not matter for timing).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I clearconverse_b200/csrc mma_bench.cu -o mma_bench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"
using namespace resep::ptx;

template <bool PAIR, bool TS, int N>
__global__ void __launch_bounds__(128, 1) k_bench(int n_mma, int commit_every, int same_acc, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
  if (warp == 1) { if (PAIR) tmem_alloc_pair<512>(&tmem_slot); else tmem_alloc<512>(&tmem_slot); }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  constexpr uint32_t IDESC = umma_idesc(UMMA_BF16, UMMA_BF16, PAIR ? 256 : 128, N);
  if (warp == 1 && lane == 0 && rank == 0) {
    const uint64_t adesc = umma_desc_k_sw128(smem_u32(smem));
    const uint64_t bdesc = umma_desc_k_sw128(smem_u32(smem + 32768));
    uint32_t ph = 0;
    long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t d = tmem + (same_acc ? 0 : ((i / 8) & 1) * 256);
      const int k = i & 3;
      if (TS) {
        if (PAIR) umma_bf16_ts_pair(d, tmem + 480 + 8 * k, bdesc + 2 * k, IDESC, (i & 7) != 0);
        else umma_bf16_ts(d, tmem + 480 + 8 * k, bdesc + 2 * k, IDESC, (i & 7) != 0);
      } else {
        if (PAIR) umma_bf16_pair(d, adesc + 2 * k, bdesc + 2 * k, IDESC, (i & 7) != 0);
        else umma_bf16(d, adesc + 2 * k, bdesc + 2 * k, IDESC, (i & 7) != 0);
      }
      if (commit_every > 0 && (i + 1) % commit_every == 0 && i + 1 < n_mma) {
        if (PAIR) umma_commit_pair(&bar[1]); else umma_commit(&bar[1]);   // nobody waits on it
      }
    }
    long long t1 = clock64();
    if (PAIR) umma_commit_pair(&bar[0]); else umma_commit(&bar[0]);
    mbar_wait(&bar[0], ph);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  __syncwarp();
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 1) { tc_fence_after(); if (PAIR) tmem_dealloc_pair<512>(tmem); else tmem_dealloc<512>(tmem); }
}

template <bool PAIR, bool TS, int N>
void run(const char* name, int grid, long long* d_out) {
  auto kern = k_bench<PAIR, TS, N>;
  const int smem = 96 * 1024;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int commit_every : {0, 4, 8}) {
    for (int same_acc : {1, 0}) {
      const int n = 2048;
      long long h[2] = {0, 0};
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute attr; attr.id = cudaLaunchAttributeClusterDimension;
      attr.val.clusterDim.x = PAIR ? 2 : 1; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
      cfg.attrs = &attr; cfg.numAttrs = 1;
      for (int rep = 0; rep < 2; ++rep) {
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, n, commit_every, same_acc, d_out);
        if (e != cudaSuccess) { printf("%s launch: %s\n", name, cudaGetErrorString(e)); return; }
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s sync: %s\n", name, cudaGetErrorString(e)); exit(1); }
      }
      cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
      const double ideal = (PAIR ? 256.0 : 128.0) * N / (256.0 * (PAIR ? 2 : 1));
      printf("%-22s grid %3d commit/%d %s: issue %6.1f cyc/mma, complete %6.1f cyc/mma (floor %.0f)\n", name, grid, commit_every,
             same_acc ? "same-acc" : "alt-acc ", (double)h[0] / n, (double)h[1] / n, ideal);
    }
  }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 64);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int grid : {2, sms & ~1}) {
    run<false, false, 128>("1cta SS N=128", grid, d_out);
    run<false, true, 128>("1cta TS N=128", grid, d_out);
    run<false, true, 256>("1cta TS N=256", grid, d_out);
    run<false, true, 64>("1cta TS N=64", grid, d_out);
    run<true, false, 128>("2cta SS N=128", grid, d_out);
    run<true, false, 256>("2cta SS N=256", grid, d_out);
    run<true, true, 64>("2cta TS N=64", grid, d_out);
    run<true, true, 128>("2cta TS N=128", grid, d_out);
    run<true, true, 256>("2cta TS N=256", grid, d_out);
  }
  return 0;
}
