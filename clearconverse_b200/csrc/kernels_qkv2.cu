// k_qkv2_tc: y = LayerNorm(o; g1, be1, eps 1e-6); qkv = y . Win^T + bin  -- norm1 + the packed in-projection of
// nn.MultiheadAttention (speechbrain Transformer.py TransformerEncoderLayer.forward, normalize_before=True) on a
// PAIR of CTAs (tcgen05 cta_group::2), bf16 out.
//
// The kernel is HBM-bound (512 B in + 768 B out per token row against 98,304 FLOP), so the design goal is to keep
// the loads and stores streaming and everything else out of their way:
//   * the in-projection weights (hi + lo, 192 KB bf16) are RESIDENT: each CTA of the pair holds its half (N/2 rows
//     of every [128 x 128] B tile, 96 KB) for the whole kernel -- no weight ring, no per-tile L2 traffic;
//   * LayerNorm warps (thread = row x column half, one pass over the shared tile, packed f32x2 math, one exchange
//     per row) write the bf16 A operand straight to TMEM, double-buffered, and run one tile ahead of the MMAs;
//   * 48 MMAs (M=256, N=128, K=16) per pair-tile, issued elect-style by the whole MMA warp (see kernels_post2.cu);
//   * drain warps empty accumulator n of tile t while the tensor pipe works on the next one; output tiles leave
//     through double-buffered swizzled staging + TMA stores; the bias comes from the constant bank.
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "tc_common.cuh"

namespace resep {

using namespace ptx;

namespace qkv2 {
constexpr int DRAIN_WARPS = 8, LN_WARPS = 8;
constexpr int THREADS = 96 + 32 * (DRAIN_WARPS + LN_WARPS);   // 608
constexpr int DRAIN_THREADS = 32 * DRAIN_WARPS, LN_THREADS = 32 * LN_WARPS;
constexpr int ATOM = 128 * 128;                  // [128 rows x 128 B]
constexpr int WUNIT = 16384;                     // this CTA's [64 B-rows x 128 K] of one (n, part): two [64 x 64] atoms
constexpr int OFF_W = 0;                         // 6 units: (n, part) = n * 2 + part
constexpr int OFF_OT = 6 * WUNIT;                // [128 x 128] fp32 residual tile (4 atoms)
constexpr int OFF_OUT = OFF_OT + 4 * ATOM;       // 2 x [128 x 128] bf16 output staging (2 atoms each)
constexpr int OFF_RED = OFF_OUT + 4 * ATOM;      // [2][2][128] LayerNorm partial sums / squares
constexpr int OFF_BAR = OFF_RED + 4 * 128 * 4;
constexpr int NBAR = 13;
constexpr int SMEM = OFF_BAR + NBAR * 8 + 16;
constexpr int TM_ACC = 0, TM_Y2 = 384;           // 3 x 128 accumulator columns, 2 x 64 packed LN columns
static_assert(SMEM <= 227 * 1024, "shared memory budget");
}  // namespace qkv2

struct Qkv2Args {
  // by value = constant bank; every index below is a compile-time constant, so these become immediate operands
  float bin[3 * D];      // in-proj bias (drain warps)
  float g1[D], be1[D];   // norm1 weight / bias (LayerNorm warps)
  int64_t M;
};

template <int C, int N, class F>
__device__ __forceinline__ void static_for_q(F&& f) {
  if constexpr (C < N) {
    f(std::integral_constant<int, C>{});
    static_for_q<C + 1, N>(f);
  }
}

__device__ __forceinline__ uint32_t sw128_f32_q(int row, int col) {
  return (uint32_t)((col >> 5) * qkv2::ATOM + row * 128 + ((((col & 31) >> 2) ^ (row & 7)) << 4));
}
__device__ __forceinline__ uint32_t sw128_bf16_q(int row, int col) {   // 16-byte chunk of 8 bf16 columns, two 64-column atoms
  return (uint32_t)((col >> 6) * qkv2::ATOM + row * 128 + ((((col & 63) >> 3) ^ (row & 7)) << 4));
}

template <bool SPLIT, bool F16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(qkv2::THREADS, 1)
k_qkv2_tc(const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmW,
          const __grid_constant__ CUtensorMap tmWL, const __grid_constant__ CUtensorMap tmQ,
          const __grid_constant__ Qkv2Args args) {
  using namespace qkv2;
  constexpr int PARTS = SPLIT ? 2 : 1;
  constexpr uint32_t IDESC = umma_idesc(F16 ? UMMA_F16 : UMMA_BF16, F16 ? UMMA_F16 : UMMA_BF16, 256, 128);
  extern __shared__ __align__(1024) uint8_t smem[];
  float* s_sum = reinterpret_cast<float*>(smem + OFF_RED);
  float* s_sq = s_sum + 256;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* w_full = bars;                 // leader: both CTAs' resident weights landed (once)
  uint64_t* ot_full = bars + 1;            // local: residual tile landed (TMA)
  uint64_t* ot_empty = bars + 2;           // local: the LN warps have read it (8 warp arrivals)
  uint64_t* y_full = bars + 3;             // [2] leader: packed LN output of both CTAs stored in TMEM (16 warp arrivals)
  uint64_t* y_empty = bars + 5;            // [2] both: the MMAs that read it retired (pair commit)
  uint64_t* acc_full = bars + 7;           // [3] both: output tile accumulated (pair commit)
  uint64_t* acc_empty = bars + 10;         // [3] leader: drain warps of both CTAs have it in registers (16 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();                                   // the next kernel's prologue may overlap this kernel's tail
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int m_ptiles = (int)((args.M + 255) / 256);
  const int n_iters = pair < m_ptiles ? (m_ptiles - pair + npairs - 1) / npairs : 0;
  auto row0_of = [&](int it) { return ((pair + it * npairs) * 2 + (int)rank) * 128; };

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmO); prefetch_tmap(&tmW); prefetch_tmap(&tmQ);
    if (SPLIT) prefetch_tmap(&tmWL);
    mbar_init(w_full, 1); mbar_init(ot_full, 1); mbar_init(ot_empty, LN_WARPS);
    for (int i = 0; i < 2; ++i) { mbar_init(&y_full[i], 2 * LN_WARPS); mbar_init(&y_empty[i], 1); }
    for (int i = 0; i < 3; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 2 * DRAIN_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair<512>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // Programmatic dependent launch: everything above and the weight loads below are independent of the previous
  // kernel (which produced `o` and may still be reading the qkv buffer this kernel overwrites).
  if (warp != 0) pdl_wait();
  if (warp == 0) {
    if (lane == 0 && n_iters > 0) {                // resident weights: this CTA's 64 rows of every (n, part) tile
      const uint32_t wfull = mapa_u32(smem_u32(w_full), 0);
      if (leader) mbar_arrive_expect_tx(w_full, 2 * 3 * PARTS * WUNIT);
      for (int n = 0; n < 3; ++n)
        for (int part = 0; part < PARTS; ++part) {
          uint8_t* dst = smem + OFF_W + (n * 2 + part) * WUNIT;
          const CUtensorMap* m = part ? &tmWL : &tmW;
          tma_load_2d_pair(dst, m, wfull, 0, n * 128 + (int)rank * 64);
          tma_load_2d_pair(dst + WUNIT / 2, m, wfull, 64, n * 128 + (int)rank * 64);
        }
    }
    __syncwarp();
  } else if (warp == 2) {
    if (lane == 0) {                               // residual tile producer
      for (int it = 0; it < n_iters; ++it) {
        mbar_wait(ot_empty, (it & 1) ^ 1);
        mbar_arrive_expect_tx(ot_full, 4 * ATOM);
#pragma unroll
        for (int j = 0; j < 4; ++j) tma_load_2d(smem + OFF_OT + j * ATOM, &tmO, ot_full, 32 * j, row0_of(it));
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (leader && n_iters > 0) {                   // MMA issuer: whole warp runs the schedule, elect.sync lane issues
      mbar_wait(w_full, 0);
#pragma unroll 1
      for (int it = 0; it < n_iters; ++it) {
        const int b = it & 1;
        mbar_wait(&y_full[b], (it >> 1) & 1);
        tc_fence_after();
        const uint32_t a = tmem + TM_Y2 + 64 * b;
#pragma unroll
        for (int n = 0; n < 3; ++n) {
          mbar_wait(&acc_empty[n], (it & 1) ^ 1);  // slot n is used once per tile
          tc_fence_after();
          const uint32_t d = tmem + TM_ACC + 128 * n;
          if (elect_one()) {
#pragma unroll
            for (int part = 0; part < PARTS; ++part) {
              const uint32_t w = smem_u32(smem + OFF_W + (n * 2 + part) * WUNIT);
              const uint64_t b0 = umma_desc_k_sw128(w), b1 = umma_desc_k_sw128(w + WUNIT / 2);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_bf16_ts_pair(d, a + 8 * k, b0 + 2 * k, IDESC, !(part == 0 && k == 0));
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_bf16_ts_pair(d, a + 32 + 8 * k, b1 + 2 * k, IDESC, true);
            }
            umma_commit_pair(&acc_full[n]);
            if (n == 2) umma_commit_pair(&y_empty[b]);
          }
          __syncwarp();
        }
      }
    }
    __syncwarp();
  } else if (warp >= 3 + DRAIN_WARPS) {
    // ------------------------------------------------------------------ LayerNorm warps: thread = (row, column half)
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t yfull0 = mapa_u32(smem_u32(&y_full[0]), 0);
    auto lbar = [&]() { asm volatile("bar.sync 3, %0;" ::"n"(LN_THREADS) : "memory"); };
    auto run = [&](auto hf_c) {
    constexpr int hf = decltype(hf_c)::value;
#pragma unroll 1
    for (int it = 0; it < n_iters; ++it) {
      const int b = it & 1;
      mbar_wait(ot_full, it & 1);
      float2 v[32];
      float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float4 o4 = *reinterpret_cast<const float4*>(smem + OFF_OT + sw128_f32_q(r, 64 * hf + 4 * j));
        v[2 * j] = make_float2(o4.x, o4.y);
        v[2 * j + 1] = make_float2(o4.z, o4.w);
        s1 = fadd2(s1, fadd2(v[2 * j], v[2 * j + 1]));
        s2 = ffma2(v[2 * j], v[2 * j], s2);
        s2 = ffma2(v[2 * j + 1], v[2 * j + 1], s2);
      }
      s_sum[hf * 128 + r] = s1.x + s1.y;
      s_sq[hf * 128 + r] = s2.x + s2.y;
      __syncwarp();
      if (lane == 0) mbar_arrive(ot_empty);          // tile buffer may be refilled
      lbar();
      // Chan's combination of the two halves' (mean, M2)
      const float m0 = s_sum[r] * (1.f / 64), m1 = s_sum[128 + r] * (1.f / 64);
      const float M0 = s_sq[r] - s_sum[r] * m0, M1 = s_sq[128 + r] - s_sum[128 + r] * m1;
      const float mean = 0.5f * (m0 + m1), dm = m0 - m1;
      const float var = fmaxf((M0 + M1 + 32.f * dm * dm) * (1.f / D), 0.f);
      const float rstd = rsqrtf(var + LN_EPS);
      const float2 rs2 = make_float2(rstd, rstd), nm2 = make_float2(-mean * rstd, -mean * rstd);
      mbar_wait(&y_empty[b], ((it >> 1) & 1) ^ 1);   // the MMAs of tile it - 2 no longer read this buffer
      tc_fence_after();
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        uint32_t p[16];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int j = 8 * h2 + jj;
          const int c = 64 * hf + 4 * j;
          const float2 y0 = ffma2(ffma2(v[2 * j], rs2, nm2), make_float2(args.g1[c], args.g1[c + 1]), make_float2(args.be1[c], args.be1[c + 1]));
          const float2 y1 = ffma2(ffma2(v[2 * j + 1], rs2, nm2), make_float2(args.g1[c + 2], args.g1[c + 3]), make_float2(args.be1[c + 2], args.be1[c + 3]));
          p[2 * jj] = pack16<F16>(y0.x, y0.y);
          p[2 * jj + 1] = pack16<F16>(y1.x, y1.y);
        }
        tmem_st16(lane_base + TM_Y2 + 64 * b + 32 * hf + 16 * h2, p);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(yfull0 + 8 * b);
      lbar();                                        // s_sum / s_sq are rewritten by the next tile
    }
    };
    if (((warp - 3 - DRAIN_WARPS) >> 2) == 0) run(std::integral_constant<int, 0>{});
    else run(std::integral_constant<int, 1>{});
  } else {
    // ------------------------------------------------------------------ drain warps: (row r, column half hf)
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t accempty0 = mapa_u32(smem_u32(&acc_empty[0]), 0);
    const bool elected = warp == 3 && lane == 0;
    auto dbar = [&]() { asm volatile("bar.sync 2, %0;" ::"n"(DRAIN_THREADS) : "memory"); };
    auto run = [&](auto hf_c) {
      constexpr int hf = decltype(hf_c)::value;
      constexpr int cb = hf * 64;
      uint32_t nstore = 0;
#pragma unroll 1
      for (int it = 0; it < n_iters; ++it) {
        const int row0 = row0_of(it);
        static_for_q<0, 3>([&](auto n_c) {
          constexpr int n = decltype(n_c)::value;
          uint8_t* out = smem + OFF_OUT + (nstore & 1) * 2 * ATOM;
          mbar_wait(&acc_full[n], it & 1);
          tc_fence_after();
          float2 v[32];
          tmem_ld32(lane_base + TM_ACC + n * 128 + cb, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
          tmem_ld32(lane_base + TM_ACC + n * 128 + cb + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[16]));
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(accempty0 + 8 * n);
          if (elected) tma_store_wait_read<1>();     // the store issued two output tiles ago has read this staging buffer
          dbar();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float2 x[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
              x[i] = fadd2(v[4 * j + i], make_float2(args.bin[n * 128 + cb + 8 * j + 2 * i], args.bin[n * 128 + cb + 8 * j + 2 * i + 1]));
            *reinterpret_cast<uint4*>(out + sw128_bf16_q(r, cb + 8 * j)) =
                make_uint4(pack16<F16>(x[0].x, x[0].y), pack16<F16>(x[1].x, x[1].y), pack16<F16>(x[2].x, x[2].y), pack16<F16>(x[3].x, x[3].y));
          }
          fence_proxy_async();
          dbar();
          if (elected) {
            tma_store_2d(&tmQ, out, n * 128, row0);
            tma_store_2d(&tmQ, out + ATOM, n * 128 + 64, row0);
            tma_store_commit();
          }
          ++nstore;
        });
      }
    };
    if (((warp - 3) >> 2) == 0) run(std::integral_constant<int, 0>{});
    else run(std::integral_constant<int, 1>{});
    if (elected) tma_store_wait<0>();
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair<512>(tmem);
  }
}

int launch_qkv2_tc(ResepHandle* h, const LayerDev& lw, const float* o, bf16* qkv, int64_t rows, cudaStream_t st) {
  if (rows <= 0) return RESEP_OK;
  const bool split = h->w16_mode >= 1;
  CUtensorMap tmO, tmW, tmWL, tmQ;
  int rc;
  if ((rc = make_tmap<float>(h, &tmO, o, rows, D, 128))) return rc;
  if ((rc = make_tmap<bf16>(h, &tmW, h->fmt16 ? lw.in_w_h[0] : lw.in_w_bf, 3 * D, D, 64))) return rc;
  if ((rc = make_tmap<bf16>(h, &tmWL, h->fmt16 ? lw.in_w_h[1] : lw.in_w_bl, 3 * D, D, 64))) return rc;
  if ((rc = make_tmap<bf16>(h, &tmQ, qkv, rows, 3 * D, 128))) return rc;
  Qkv2Args a;
  std::memcpy(a.bin, lw.h_in_b, 3 * D * 4);
  std::memcpy(a.g1, lw.h_in_b + 3 * D, D * 4); std::memcpy(a.be1, lw.h_in_b + 4 * D, D * 4);
  a.M = rows;
  auto kern = h->fmt16 ? (split ? k_qkv2_tc<true, true> : k_qkv2_tc<false, true>) : (split ? k_qkv2_tc<true, false> : k_qkv2_tc<false, false>);
  RESEP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, qkv2::SMEM));
  const int max_pairs = h->sm_count / 2;
  const int ptiles = (int)((rows + 255) / 256);
  const int npairs = ptiles < max_pairs ? ptiles : max_pairs;
  // profiling events bracket the launch itself, not the host-side descriptor encoding above
  ProfScope prof_scope(h, rows >= 8192 ? "k_qkv2_tc" : "k_qkv2_tc(small)", st);
  RESEP_CUDA(h, launch_pdl(kern, dim3(2 * npairs), dim3(qkv2::THREADS), qkv2::SMEM, st, tmO, tmW, tmWL, tmQ, a));
  RESEP_LAUNCH_CHECK(h, "k_qkv2_tc");
  return RESEP_OK;
}

}  // namespace resep
