// k_post2_tc: the post-attention half of one pre-norm TransformerEncoderLayer (speechbrain
// Transformer.py TransformerEncoderLayer.forward, normalize_before=True) on a PAIR of CTAs
// (thread-block cluster of 2, tcgen05 cta_group::2):
//
//     o' = o + ctx . Wo^T + bo                       (out-proj + residual)
//     y  = LayerNorm(o'; g2, be2, eps 1e-6)          (norm2; g2, be2 are folded into W1g, b1g on the host)
//     o  = o' + relu(y . W1^T + b1) . W2^T + b2      (pos_ffn + residual)
//
// Why a pair.  With one CTA per 128-row tile every M=128,N=128,K=16 MMA reads a 4 KB B slice from shared
// memory in 64 cycles while the weight TMA writes the next slice at the same rate (the 128 B/clk shared-memory
// port is saturated) and every SM pulls the whole layer's weights through L2 per tile.  With cta_group::2 one
// MMA covers 256 rows (128 per CTA, accumulators in each CTA's own TMEM) and each CTA holds only HALF of every
// weight tile (N/2 rows of B), so the shared-memory and L2 weight traffic per FLOP halve.
//
// TMEM (512 columns per CTA): three [128 x 128] fp32 regions whose roles rotate every tile, plus region X (128
// columns).  Role 0 ("Y") holds o' (written by E1) and FFN2 accumulates straight onto it, so the residual lives in
// TMEM for the whole tile.  Roles 1 and 2 are the FFN1 accumulators of the even / odd 128-wide hidden chunks,
// re-packed in place to bf16 (A operand of FFN2).  X is the out-proj accumulator of a tile and then, in its first
// 64 columns, the packed LayerNorm output Y2 (A operand of that tile's FFN1s): it is dead from the last FFN1 of
// tile t on, so the out-proj of tile t+1 is issued right there, in front of F2(t,6) and F2(t,7), whose 1,024 MMA
// cycles cover the LayerNorm epilogue E1(t+1).  Rotation: Y(t+1) = the even slot of tile t (free after F2(t,6), long
// before E1(t+1) stores o' there), even slot(t+1) = odd slot(t) (last read by F2(t,7), which precedes F1(t+1,0) on the
// in-order tensor pipe), odd slot(t+1) = Y(t), which group B must have read first: it loads its 64 columns per thread
// in one go and releases the region ~400 cycles after the last FFN2 retires, then stages and stores at leisure
// (LSU traffic starves next to the MMA operand reads and weight fills: ~2.9k cycles per tile, off the critical path).
// The activation operands never touch shared memory (TS-form MMAs); only ctx (A of the out-proj) does.
//
// Warp roles (28 warps): 0 = weight TMA producer (16 KB units = this CTA's 64 B-rows x 128 K = 8 MMAs),
// 1 = MMA issuer (leader CTA only) + TMEM owner, 2 = tile producer (ctx + residual tiles), 3 = idle,
// 4..19 = group A (thread = token row x 32-column quarter): E2 = FFN chunk epilogue (bias + relu + bf16 pack,
// TMEM -> TMEM) and E1 = tile prologue (out-proj accumulator + bias + residual = o', LayerNorm -> Y2, o' -> Y),
// 20..27 = group B (thread = token row x 64-column half): E3 = tile epilogue (Y + b2 -> staged slabs -> TMA store).
//
// MMA issue order per tile (the weight producer streams units in exactly this order):
//     F2(0) F1(2) | F2(1) F1(3) | ... | F2(5) F1(7) | OUT(t+1) F2(6) | F2(7) F1(t+1,0) F1(t+1,1)
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "tc_common.cuh"

namespace resep {

using namespace ptx;

namespace post2 {
constexpr int A_WARPS = 16, B_WARPS = 8;
constexpr int THREADS = 128 + 32 * (A_WARPS + B_WARPS);   // 896
constexpr int A_THREADS = 32 * A_WARPS, B_THREADS = 32 * B_WARPS;
constexpr int ATOM = 128 * 128;          // [128 rows x 128 B] swizzle-128B tile (16 KB)
constexpr int UNIT = 16384;              // one weight unit of this CTA: two [64 rows x 64 K] bf16 atoms = 8 MMAs
constexpr int NW = 5;                    // weight units in flight per CTA (512 MMA-cycles each)
constexpr int NCH = FFN / 128;           // 8 hidden chunks of 128
constexpr int OFF_CTX = 0;               // [128 x 128] bf16 ctx tile (2 atoms): A operand of the out-proj
constexpr int OFF_RES = 2 * ATOM;        // [128 x 128] fp32 residual tile (4 atoms of 32 columns)
constexpr int OFF_STG = 6 * ATOM;        // result staging for the TMA store: slabs 2, 3 of the [128 x 128] fp32 tile (slabs 0, 1 reuse the ctx buffer)
constexpr int OFF_W = 8 * ATOM;
constexpr int OFF_PAR = OFF_W + NW * UNIT;
constexpr int PAR_FLOATS = 2 * D + FFN;  // bo, b2, b1g (b1 with norm2's shift folded in)
constexpr int OFF_RED = OFF_PAR + PAR_FLOATS * 4;   // [2][4][128] floats: LayerNorm partial sums / squares of the column quarters
constexpr int OFF_BAR = OFF_RED + 8 * 128 * 4;
constexpr int NBAR = 2 * NW + 16;
constexpr int SMEM = OFF_BAR + NBAR * 8 + 16;
// TMEM: three [128 x 128] fp32 regions whose roles rotate every tile (role 0 = Y: o' -> FFN2 accumulator; roles 1, 2 =
// FFN1 chunk accumulators of the even / odd chunks, re-packed in place to bf16), region of (tile t, role) =
// 128 * ((role + t) % 3); and X: out-proj accumulator [128 x 128] fp32, then the packed LayerNorm output Y2 in its
// first 64 columns.
constexpr int TM_X = 384;
static_assert(SMEM <= 227 * 1024, "shared memory budget");
}  // namespace post2

struct Post2Args {
  const float* par;   // DEVICE: bo[128] | b2[128] | b1g[1024] (FFN1 bias with norm2's shift folded in), one bulk copy at kernel entry
                      // (staging them from kernel parameters cost 3.8k cycles of register-indexed constant loads per launch)
  int64_t M;
  int dbg;            // development aid (RESEP_DBG): bit 0 = no weight TMA / no w_full waits
  long long* trace;   // development aid (RESEP_TRACE): [3 roles][256] (tag, clock) pairs of CTA 0; null in production
};
// role 0 = MMA issuer, 1 = group A warp 3, 2 = group B warp 11 (lane 0 each)
#ifndef RESEP_TRACE_BUILD
#define TR(role, tag) do { } while (0)
#else
#define TR(role, tag)                                                                     \
  do {                                                                                    \
    if (args.trace != nullptr && blockIdx.x == 0 && tr_n < 256) {                         \
      args.trace[(role) * 512 + 2 * tr_n] = (tag);                                        \
      args.trace[(role) * 512 + 2 * tr_n + 1] = clock64();                                \
      ++tr_n;                                                                             \
    }                                                                                     \
  } while (0)
#endif

template <int C, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (C < N) {
    f(std::integral_constant<int, C>{});
    static_for<C + 1, N>(f);
  }
}

__device__ __forceinline__ uint32_t sw128_f32_off(int row, int col) {   // fp32 [128 x 128] tile as four 32-column swizzled atoms
  return (uint32_t)((col >> 5) * post2::ATOM + row * 128 + ((((col & 31) >> 2) ^ (row & 7)) << 4));
}

template <bool SPLIT, int FFN_SPLIT, bool F16>   // FFN_SPLIT: 0 = FFN weights single, 1 = both hi + lo, 2 = FFN2 only, 3 = FFN1 only
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(post2::THREADS, 1)
k_post2_tc(const __grid_constant__ CUtensorMap tmCtx, const __grid_constant__ CUtensorMap tmO,
           const __grid_constant__ CUtensorMap tmWo, const __grid_constant__ CUtensorMap tmWoL,
           const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW1L,
           const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmW2L,
           const __grid_constant__ Post2Args args) {
  using namespace post2;
  constexpr int PARTS = SPLIT ? 2 : 1;          // out-proj weight operand: bf16 hi (+ lo)
  constexpr int F1PARTS = (FFN_SPLIT == 1 || FFN_SPLIT == 3) ? 2 : 1, F2PARTS = (FFN_SPLIT == 1 || FFN_SPLIT == 2) ? 2 : 1;   // FFN weight operands
  constexpr uint32_t IDESC = umma_idesc(F16 ? UMMA_F16 : UMMA_BF16, F16 ? UMMA_F16 : UMMA_BF16, 256, 128);

  extern __shared__ __align__(1024) uint8_t smem[];
  float* par = reinterpret_cast<float*>(smem + OFF_PAR);
  float *s_bo = par, *s_b2 = par + D, *s_b1 = par + 2 * D;
  float* s_sum = reinterpret_cast<float*>(smem + OFF_RED);   // [4][128]
  float* s_sq = s_sum + 512;                                  // [4][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* w_full = bars;                 // [NW] leader: both CTAs' TMA bytes of the unit landed
  uint64_t* w_empty = bars + NW;           // [NW] both: the MMAs that read the unit retired (pair commit)
  uint64_t* ctx_full = bars + 2 * NW;      // leader: both ctx tiles landed
  uint64_t* ctx_empty = ctx_full + 1;      // both: out-proj MMAs retired
  uint64_t* res_full = ctx_full + 2;       // local: residual tile landed
  uint64_t* res_empty = ctx_full + 3;      // local: group A has read it (8 warp arrivals)
  uint64_t* out_full = ctx_full + 4;       // both: out-proj accumulator of the tile ready (pair commit)
  uint64_t* y_full = ctx_full + 5;         // leader: LN2(o') of the tile in Y2, both CTAs (32 warp arrivals)
  uint64_t* acch_full = ctx_full + 6;      // [2] both: FFN1 chunk accumulator ready (pair commit)
  uint64_t* hs_full = ctx_full + 8;        // [2] leader: relu'd bf16 chunk stored by both CTAs (32 warp arrivals)
  uint64_t* accy_done = ctx_full + 10;     // both: every FFN2 of the tile retired
  uint64_t* yp_full = ctx_full + 11;       // leader: o'(t) stored in the tile's Y region by both CTAs (32 warp arrivals)
  uint64_t* yreg_free = ctx_full + 12;     // leader: both CTAs have the tile's result in registers (16 warp arrivals)
  uint64_t* stg_free = ctx_full + 13;      // local: the tile's result has left the staging area (which includes the ctx buffer)
  uint64_t* par_full = ctx_full + 15;      // local: the parameter block has landed
  uint64_t* f26_done = ctx_full + 14;      // both: F2(t, 6) retired: its FFN1 slot is the Y region of tile t + 1 and may take o'(t + 1)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();                           // the next kernel's prologue may overlap this kernel's tail
  int tr_n = 0;
#ifdef RESEP_TRACE_BUILD
  if (args.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) args.trace[1535] = clock64();
#endif
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int m_ptiles = (int)((args.M + 255) / 256);
  const int n_iters = pair < m_ptiles ? (m_ptiles - pair + npairs - 1) / npairs : 0;
  auto row0_of = [&](int it) { return ((pair + it * npairs) * 2 + (int)rank) * 128; };

#ifdef RESEP_TRACE_BUILD
  if (args.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) args.trace[1531] = clock64();
#endif
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmCtx); prefetch_tmap(&tmO); prefetch_tmap(&tmWo); prefetch_tmap(&tmW1); prefetch_tmap(&tmW2);
    if (SPLIT) prefetch_tmap(&tmWoL);
    if (F1PARTS == 2) prefetch_tmap(&tmW1L);
    if (F2PARTS == 2) prefetch_tmap(&tmW2L);
    for (int i = 0; i < NW; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    mbar_init(ctx_full, 1); mbar_init(ctx_empty, 1); mbar_init(res_full, 1); mbar_init(res_empty, A_WARPS);
    mbar_init(out_full, 1); mbar_init(y_full, 2 * A_WARPS);
    for (int i = 0; i < 2; ++i) { mbar_init(&acch_full[i], 1); mbar_init(&hs_full[i], 2 * A_WARPS); }
    mbar_init(accy_done, 1); mbar_init(yp_full, 2 * A_WARPS); mbar_init(yreg_free, 2 * B_WARPS); mbar_init(stg_free, 2); mbar_init(f26_done, 1);
    mbar_init(par_full, 1);
    fence_barrier_init();
    mbar_arrive_expect_tx(par_full, PAR_FLOATS * 4);       // weights, not the previous kernel's output: no PDL wait needed
    bulk_load_1d(par, args.par, PAR_FLOATS * 4, par_full);
#ifdef RESEP_TRACE_BUILD
    if (args.trace != nullptr && blockIdx.x == 0) args.trace[1530] = clock64();
#endif
  }
  if (warp == 1) {
    tmem_alloc_pair<512>(tmem_slot);
#ifdef RESEP_TRACE_BUILD
    if (args.trace != nullptr && blockIdx.x == 0 && lane == 0) args.trace[1529] = clock64();
#endif
  }
  tc_fence_before();
  cluster_sync_all();                      // barriers of both CTAs are initialised before any remote arrive / TMA
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
#ifdef RESEP_TRACE_BUILD
  if (args.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) args.trace[1534] = clock64();
#endif

  // Programmatic dependent launch: the weight ring is filled without waiting for the previous kernel (attention,
  // which produces ctx); every other role touches ctx / o and waits for it to complete first.
  if (warp != 0) pdl_wait();
  if (warp == 0) {
    // ------------------------------------------------------------------ weight producer (both CTAs: own halves)
    if (lane == 0 && n_iters > 0 && !(args.dbg & 1)) {
      int st = 0;
      uint32_t ph = 0;
      const uint32_t wfull0 = mapa_u32(smem_u32(&w_full[0]), 0);
      // one unit = this CTA's 64 B-rows x 128 K: two [64 x 64] boxes at (c0, r0) and (c0 + 64, r0)
      auto put = [&](const CUtensorMap* hi, const CUtensorMap* lo, int c0, int r0, int parts) {
        for (int part = 0; part < parts; ++part) {
          mbar_wait(&w_empty[st], ph ^ 1);
          if (leader) mbar_arrive_expect_tx(&w_full[st], 2 * UNIT);
          uint8_t* dst = smem + OFF_W + st * UNIT;
          const CUtensorMap* m = part ? lo : hi;
          tma_load_2d_pair(dst, m, wfull0 + 8 * st, c0, r0);
          tma_load_2d_pair(dst + UNIT / 2, m, wfull0 + 8 * st, c0 + 64, r0);
          if (++st == NW) { st = 0; ph ^= 1; }
        }
      };
      auto put_out = [&]() { put(&tmWo, &tmWoL, 0, (int)rank * 64, PARTS); };
      auto put_f1 = [&](int c) { put(&tmW1, &tmW1L, 0, c * 128 + (int)rank * 64, F1PARTS); };
      auto put_f2 = [&](int c) { put(&tmW2, &tmW2L, c * 128, (int)rank * 64, F2PARTS); };
      put_out();
      put_f1(0);
      put_f1(1);
      for (int t = 0; t < n_iters; ++t)
        for (int c = 0; c < NCH; ++c) {
          if (c == NCH - 2 && t + 1 < n_iters) put_out();
          put_f2(c);
          if (c + 2 < NCH) put_f1(c + 2);
          else if (c == NCH - 1 && t + 1 < n_iters) { put_f1(0); put_f1(1); }
        }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ------------------------------------------------------------------ tile producer: ctx + residual
    if (lane == 0) {
      const uint32_t ctxfull = mapa_u32(smem_u32(ctx_full), 0);
      for (int it = 0; it < n_iters; ++it) {
        const int r0 = row0_of(it);
        mbar_wait(ctx_empty, (it & 1) ^ 1);                         // out-proj of tile it - 1 has read the buffer
        if (it >= 2) mbar_wait(stg_free, it & 1);                    // ... and the result of tile it - 2 has been stored from it
        if (leader) mbar_arrive_expect_tx(ctx_full, 4 * ATOM);     // both CTAs' 32 KB
        tma_load_2d_pair(smem + OFF_CTX, &tmCtx, ctxfull, 0, r0);
        tma_load_2d_pair(smem + OFF_CTX + ATOM, &tmCtx, ctxfull, 64, r0);
        mbar_wait(res_empty, (it & 1) ^ 1);
        mbar_arrive_expect_tx(res_full, 4 * ATOM);
#pragma unroll
        for (int j = 0; j < 4; ++j) tma_load_2d(smem + OFF_RES + j * ATOM, &tmO, res_full, 32 * j, r0);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    // The WHOLE warp runs the schedule (waits, fences, ring bookkeeping); only the 8 MMAs and the commits of a
    // unit are issued by the elect.sync lane.  Under `if (lane == 0)` ptxas wraps every UTCHMMA in an election
    // loop and the issue cost (113 cycles per N=128 MMA with 4 MMAs per unit, scripts/microbench/mma_issue_bench.cu)
    // exceeds the 64 cycles the MMA takes; in this form with 8 MMAs per unit it is at the 64-cycle floor.
    // Every tcgen05.mma costs at least 64 cycles whatever its N (scripts/microbench/mma_bench.cu), so all are N=128.
    if (leader && n_iters > 0) {
      int st = 0;
      uint32_t ph = 0;
      auto wait_unit = [&]() -> uint32_t {
        if (!(args.dbg & 1)) mbar_wait(&w_full[st], ph);
        tc_fence_after();
        return smem_u32(smem + OFF_W + st * UNIT);
      };
      auto advance = [&]() { if (++st == NW) { st = 0; ph ^= 1; } };
      auto region = [&](int t, int role) -> uint32_t { return tmem + 128u * (uint32_t)((role + t) % 3); };
      auto do_out = [&](int t) {             // X = ctx(t) . Wo^T   (both operands in shared memory)
        if (lane == 0) TR(0, 1000 + t);
        mbar_wait(ctx_full, t & 1);
        if (lane == 0) TR(0, 1100 + t);
        const uint32_t d = tmem + TM_X;
#pragma unroll
        for (int part = 0; part < PARTS; ++part) {
          const uint32_t b = wait_unit();
          const uint64_t b0 = umma_desc_k_sw128(b), b1 = umma_desc_k_sw128(b + UNIT / 2);
          const uint64_t a0 = umma_desc_k_sw128(smem_u32(smem + OFF_CTX)), a1 = umma_desc_k_sw128(smem_u32(smem + OFF_CTX + ATOM));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_pair(d, a0 + 2 * k, b0 + 2 * k, IDESC, !(part == 0 && k == 0));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_pair(d, a1 + 2 * k, b1 + 2 * k, IDESC, true);
            umma_commit_pair(&w_empty[st]);
            if (part == PARTS - 1) {
              umma_commit_pair(out_full);
              umma_commit_pair(ctx_empty);
            }
          }
          __syncwarp();
          advance();
        }
        if (lane == 0) TR(0, 1200 + t);
      };
      auto do_f1 = [&](int t, int c) {       // slot(t, c) = LN2(o'(t)) [Y2, TMEM] . W1_chunk^T
        const uint32_t d = region(t, 1 + (c & 1));
        const uint32_t a = tmem + TM_X;
#pragma unroll
        for (int part = 0; part < F1PARTS; ++part) {
          const uint32_t b = wait_unit();
          const uint64_t b0 = umma_desc_k_sw128(b), b1 = umma_desc_k_sw128(b + UNIT / 2);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_ts_pair(d, a + 8 * k, b0 + 2 * k, IDESC, !(part == 0 && k == 0));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_ts_pair(d, a + 32 + 8 * k, b1 + 2 * k, IDESC, true);
            umma_commit_pair(&w_empty[st]);
            if (part == F1PARTS - 1) umma_commit_pair(&acch_full[c & 1]);
          }
          __syncwarp();
          advance();
        }
      };
      auto do_f2 = [&](int t, int c) {       // Y(t) += relu(H chunk) [packed in its accumulator] . W2_chunk^T
        const uint32_t a = region(t, 1 + (c & 1)), d = region(t, 0);
        if (lane == 0) TR(0, 3000 + t * NCH + c);
        // slot (c & 1) is used 4x per tile: use index (t * 4 + c / 2)
        mbar_wait(&hs_full[c & 1], (t * (NCH / 2) + (c >> 1)) & 1);
        if (c == 0) mbar_wait(yp_full, t & 1);           // o'(t) is in the Y region (stored after Y2, off E1's critical path)
        if (lane == 0) TR(0, 4000 + t * NCH + c);
#pragma unroll
        for (int part = 0; part < F2PARTS; ++part) {
          const uint32_t b = wait_unit();
          const uint64_t b0 = umma_desc_k_sw128(b), b1 = umma_desc_k_sw128(b + UNIT / 2);
          if (elect_one()) {
            // hidden 32 q .. 32 q + 31 of the chunk are packed in columns [32 q, 32 q + 16) of the slot
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_ts_pair(d, a + 32 * (k >> 1) + 8 * (k & 1), b0 + 2 * k, IDESC, true);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_ts_pair(d, a + 64 + 32 * (k >> 1) + 8 * (k & 1), b1 + 2 * k, IDESC, true);
            umma_commit_pair(&w_empty[st]);
            if (part == F2PARTS - 1 && c == NCH - 2) umma_commit_pair(f26_done);
            if (part == F2PARTS - 1 && c == NCH - 1) umma_commit_pair(accy_done);
          }
          __syncwarp();
          advance();
        }
      };
      auto wait_y = [&](int t) {
        if (lane == 0) TR(0, 2000 + t);
        mbar_wait(y_full, t & 1);              // LN2(o'(t)) in Y2, both CTAs
        if (lane == 0) TR(0, 2100 + t);
      };
      do_out(0);
      wait_y(0);
      do_f1(0, 0);
      do_f1(0, 1);
#pragma unroll 1
      for (int t = 0; t < n_iters; ++t) {
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
          if (c == NCH - 2 && t + 1 < n_iters) do_out(t + 1);   // X is dead: F1(t, 7) was the last reader of Y2(t)
          do_f2(t, c);
          if (c + 2 < NCH) do_f1(t, c + 2);
          else if (c == NCH - 1 && t + 1 < n_iters) {
            wait_y(t + 1);
            do_f1(t + 1, 0);                            // into the slot F2(t, 7) has read (in-order pipe: no wait)
            mbar_wait(yreg_free, t & 1);                // group B has Y(t) in registers: it becomes the odd-chunk slot of tile t + 1
            do_f1(t + 1, 1);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 4 + A_WARPS) {
    // ------------------------------------------------------------------ group A: FFN chunk epilogue (E2) + tile prologue (E1)
    // thread = (token row r, column quarter qq): TMEM lane r, 32 of the 128 columns of a region
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    // A broadcast LDS.128 costs 4 cycles of the SM's one load-return path whatever it broadcasts, so parameter loads
    // are kept off the tile-boundary critical path: norm2's scale and shift are folded into W1 / b1 on the host.
    // (Constant-bank operands were tried: with immediate offsets the code grows 8 x 4-fold and misses the instruction
    // cache, 60 us; register-indexed LDC, 52 us, is slower than the LDS it replaces, 47 us.)
    const int qq = (warp - 4) >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t hsfull0 = mapa_u32(smem_u32(&hs_full[0]), 0);
    const uint32_t yfull = mapa_u32(smem_u32(y_full), 0);
    const uint32_t ypfull = mapa_u32(smem_u32(yp_full), 0);
    const bool tracer = warp == 4 && lane == 0;
    mbar_wait(par_full, 0);
    auto abar = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(A_THREADS) : "memory"); };
    auto E1 = [&](int t, float2 (&v)[16]) {
      if (tracer) TR(1, 1000 + t);
      // While the out-proj of this tile is still on the tensor pipe: v = o + bo for this thread's 32 columns, so the
      // critical path after `out_full` has no shared-memory loads before the statistics.
      mbar_wait(res_full, t & 1);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b = *reinterpret_cast<const float4*>(s_bo + 32 * qq + 4 * j);
        const float4 o4 = *reinterpret_cast<const float4*>(smem + OFF_RES + sw128_f32_off(r, 32 * qq + 4 * j));
        v[2 * j] = fadd2(make_float2(b.x, b.y), make_float2(o4.x, o4.y));
        v[2 * j + 1] = fadd2(make_float2(b.z, b.w), make_float2(o4.z, o4.w));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(res_empty);     // the residual tile may be refilled
      mbar_wait(out_full, t & 1);
      tc_fence_after();
      if (tracer) TR(1, 1100 + t);
      // o' = acc + v; row statistics from the plain sums of this thread's 32 columns (one exchange per row)
      float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
      {
        float2 acc[16];
        tmem_ld32(lane_base + TM_X + 32 * qq, *reinterpret_cast<uint32_t(*)[32]>(&acc[0]));
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float2 x = fadd2(acc[j], v[j]);
          s1 = fadd2(s1, x);
          s2 = ffma2(x, x, s2);
          v[j] = x;
        }
      }
      s_sum[qq * 128 + r] = s1.x + s1.y;
      s_sq[qq * 128 + r] = s2.x + s2.y;
      if (tracer) TR(1, 1120 + t);
      abar();                                  // also: every thread has read its accumulator columns of X before Y2 overwrites them
      if (tracer) TR(1, 1130 + t);
      // Chan's combination of the four quarters' (mean, M2): no cancellation between the quarters
      float mq[4], M2 = 0.f, mean = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float su = s_sum[i * 128 + r];
        mq[i] = su * (1.f / 32);
        M2 += s_sq[i * 128 + r] - su * mq[i];
        mean += mq[i];
      }
      mean *= 0.25f;
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float dm = mq[i] - mean; M2 = fmaf(32.f * dm, dm, M2); }
      const float var = fmaxf(M2 * (1.f / D), 0.f);
      const float rstd = rsqrtf(var + LN_EPS);
      const float2 rs2 = make_float2(rstd, rstd), nm2 = make_float2(-mean * rstd, -mean * rstd);
      {
        uint32_t p[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          // (x - mean) * rstd; norm2's scale and shift live in W1 / b1
          const float2 y0 = ffma2(v[2 * j], rs2, nm2);
          const float2 y1 = ffma2(v[2 * j + 1], rs2, nm2);
          p[2 * j] = pack16<F16>(y0.x, y0.y);
          p[2 * j + 1] = pack16<F16>(y1.x, y1.y);
        }
        tmem_st16(lane_base + TM_X + 16 * qq, p);
      }
      if (tracer) TR(1, 1140 + t);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(yfull);
      if (tracer) TR(1, 1200 + t);
    };
    // o'(t) -> the tile's Y region (FFN2 accumulates onto it).  For t > 0 that region was the even FFN1 slot of tile
    // t - 1, last read by F2(t - 1, 6), which was issued right behind this tile's out-proj.
    auto put_oprime = [&](int t, const float2 (&v)[16]) {
      if (t > 0) { mbar_wait(f26_done, (t - 1) & 1); tc_fence_after(); }
      tmem_st32(lane_base + 128u * (uint32_t)(t % 3) + 32 * qq, *reinterpret_cast<const uint32_t(*)[32]>(&v[0]));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(ypfull);
      if (tracer) TR(1, 1300 + t);
    };
    auto E2 = [&](int t, int c) {
      const int sb = c & 1;
      const uint32_t hcol = lane_base + 128u * (uint32_t)((1 + sb + t) % 3) + 32 * qq;
      if (tracer) TR(1, 3000 + t * NCH + c);
      mbar_wait(&acch_full[sb], (t * (NCH / 2) + (c >> 1)) & 1);
      tc_fence_after();
      if (tracer) TR(1, 4000 + t * NCH + c);
      const float* bias = s_b1 + c * 128 + qq * 32;
      float2 v[16];
      tmem_ld32(hcol, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
      tmem_ld_wait();
      uint32_t p[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b4 = *reinterpret_cast<const float4*>(bias + 4 * j);
        const float2 x0 = fadd2(v[2 * j], make_float2(b4.x, b4.y)), x1 = fadd2(v[2 * j + 1], make_float2(b4.z, b4.w));
        p[2 * j] = pack16_relu<F16>(x0.x, x0.y);
        p[2 * j + 1] = pack16_relu<F16>(x1.x, x1.y);
      }
      tmem_st16(hcol, p);                    // this thread's own 32 fp32 columns -> their first 16 columns, packed
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(hsfull0 + 8 * sb);
      if (tracer) TR(1, 5000 + t * NCH + c);
    };
    if (n_iters > 0) {
      float2 v[16];
      E1(0, v);
      put_oprime(0, v);
#pragma unroll 1
      for (int t = 0; t < n_iters; ++t) {
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) E2(t, c);
        if (t + 1 < n_iters) { E1(t + 1, v); put_oprime(t + 1, v); }
      }
    }
  } else if (warp >= 4 + A_WARPS) {
    // ------------------------------------------------------------------ group B: tile epilogue (E3)
    // thread = (token row r, column half hf): TMEM lane r, 64 columns, read in one go so that Y(t) is released early
    const int q = warp & 3;
    const int hf = (warp - 4 - A_WARPS) >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t yregfree = mapa_u32(smem_u32(yreg_free), 0);
    const bool elected = ((warp - 4 - A_WARPS) & 3) == 0 && lane == 0;   // one per column half: stores that half's two slabs
    const bool tracer_b = warp == 4 + A_WARPS && lane == 0;
    auto hbar = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(2 + hf), "n"(B_THREADS / 2) : "memory"); };
    mbar_wait(par_full, 0);
    auto slab = [&](int g) -> uint8_t* { return smem + (g < 2 ? OFF_CTX + g * ATOM : OFF_STG + (g - 2) * ATOM); };
#pragma unroll 1
    for (int t = 0; t < n_iters; ++t) {
      const uint32_t ycol = lane_base + 128u * (uint32_t)(t % 3) + 64 * hf;
      const int row0 = row0_of(t);
      if (tracer_b) TR(2, 2000 + t);
      mbar_wait(accy_done, t & 1);             // every MMA up to F2(t, 7) has retired: Y(t) is final, ctx(t + 1) has been read
      tc_fence_after();
      if (tracer_b) TR(2, 2100 + t);
      float2 v[32];
      tmem_ld32(ycol, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
      tmem_ld32(ycol + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[16]));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(yregfree);   // the region may become the odd FFN1 slot of tile t + 1
      if (tracer_b) TR(2, 2150 + t);
      // (the previous tile's TMA stores have read this half's slabs: its elected thread waited, everyone of the half passed
      // the barrier at the loop's end.)  Slab by slab, so that the first store runs under the second slab's staging; the
      // two column halves stage and store independently.
#pragma unroll
      for (int g2 = 0; g2 < 2; ++g2) {
        uint8_t* srow = slab(2 * hf + g2) + r * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = *reinterpret_cast<const float4*>(s_b2 + 64 * hf + 32 * g2 + 4 * j);
          const float2 x0 = fadd2(v[16 * g2 + 2 * j], make_float2(b.x, b.y));
          const float2 x1 = fadd2(v[16 * g2 + 2 * j + 1], make_float2(b.z, b.w));
          *reinterpret_cast<float4*>(srow + ((j ^ (r & 7)) << 4)) = make_float4(x0.x, x0.y, x1.x, x1.y);
        }
        fence_proxy_async();
        hbar();
        if (elected) tma_store_2d(&tmO, slab(2 * hf + g2), 32 * (2 * hf + g2), row0);   // rows >= M are clipped
        if (tracer_b) TR(2, 2160 + 10 * g2 + t);
      }
      if (elected) {
        tma_store_commit();
        tma_store_wait_read<0>();              // this half's slabs (for hf = 0: the ctx buffer) are free again
        mbar_arrive(stg_free);                 // (two arrivals per tile)
        if (tracer_b) TR(2, 2200 + t);
      }
      hbar();                                  // nobody writes the next tile's result before the store has read this one
    }
    if (elected) tma_store_wait<0>();
    if (tracer_b) TR(2, 2300);
  }
  tc_fence_before();
#ifdef RESEP_TRACE_BUILD
  if (args.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) args.trace[1533] = clock64();
#endif
  cluster_sync_all();                      // the leader's MMAs read the peer's shared memory: nobody leaves early
#ifdef RESEP_TRACE_BUILD
  if (args.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) args.trace[1532] = clock64();
#endif
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair<512>(tmem);
  }
}

int launch_post2_tc(ResepHandle* h, const LayerDev& lw, const bf16* ctx, float* o, int64_t rows, cudaStream_t st) {
  if (rows <= 0) return RESEP_OK;
  const bool split = h->w16_mode >= 1;
  const int ffn_split = h->w16_mode == 1 ? 1 : h->w16_mode == 3 ? 2 : h->w16_mode == 4 ? 3 : 0;
  CUtensorMap tmCtx, tmO, tmWo, tmWoL, tmW1, tmW1L, tmW2, tmW2L;
  int rc;
  if ((rc = make_tmap<bf16>(h, &tmCtx, ctx, rows, D, 128))) return rc;
  if ((rc = make_tmap<float>(h, &tmO, o, rows, D, 128))) return rc;
  const bool f16 = h->fmt16 != 0;
  if ((rc = make_tmap<bf16>(h, &tmWo, f16 ? lw.out_w_h[0] : lw.out_w_bf, D, D, 64))) return rc;
  if ((rc = make_tmap<bf16>(h, &tmWoL, f16 ? lw.out_w_h[1] : lw.out_w_bl, D, D, 64))) return rc;
  if ((rc = make_tmap<bf16>(h, &tmW1, f16 ? lw.f1g_w_h[0] : lw.f1g_w_bf, FFN, D, 64))) return rc;
  if ((rc = make_tmap<bf16>(h, &tmW1L, f16 ? lw.f1g_w_h[1] : lw.f1g_w_bl, FFN, D, 64))) return rc;
  if ((rc = make_tmap<bf16>(h, &tmW2, f16 ? lw.f2_w_h[0] : lw.f2_w_bf, D, FFN, 64))) return rc;
  if ((rc = make_tmap<bf16>(h, &tmW2L, f16 ? lw.f2_w_h[1] : lw.f2_w_bl, D, FFN, 64))) return rc;
  static long long* trace_buf = nullptr;
  if (getenv("RESEP_TRACE") && !trace_buf) { cudaMalloc(&trace_buf, 1536 * 8); cudaMemset(trace_buf, 0, 1536 * 8); g_post_trace = trace_buf; }
  static const int dbg = getenv("RESEP_DBG") ? atoi(getenv("RESEP_DBG")) : 0;
  Post2Args a;
  a.par = lw.post_par;
  a.M = rows; a.dbg = dbg; a.trace = trace_buf;
  auto pick = [&](auto f16_c) {
    constexpr bool F = decltype(f16_c)::value;
    return !split ? k_post2_tc<false, 0, F> : ffn_split == 1 ? k_post2_tc<true, 1, F> : ffn_split == 2 ? k_post2_tc<true, 2, F>
         : ffn_split == 3 ? k_post2_tc<true, 3, F> : k_post2_tc<true, 0, F>;
  };
  auto kern = f16 ? pick(std::true_type{}) : pick(std::false_type{});
  RESEP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, post2::SMEM));
  static int max_pairs = 0;                // CTA pairs the device can hold at once (one CTA per SM)
  if (max_pairs == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(h->sm_count & ~1));
    cfg.blockDim = dim3(post2::THREADS);
    cfg.dynamicSmemBytes = post2::SMEM;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) { n = h->sm_count / 2; cudaGetLastError(); }
    max_pairs = n < h->sm_count / 2 ? n : h->sm_count / 2;
  }
  const int ptiles = (int)((rows + 255) / 256);
  const int npairs = ptiles < max_pairs ? ptiles : max_pairs;
  // profiling events bracket the launch itself, not the host-side descriptor encoding above
  ProfScope prof_scope(h, rows >= 8192 ? "k_post2_tc" : "k_post2_tc(small)", st);
  RESEP_CUDA(h, launch_pdl(kern, dim3(2 * npairs), dim3(post2::THREADS), post2::SMEM, st, tmCtx, tmO, tmWo, tmWoL, tmW1, tmW1L, tmW2, tmW2L, a));
  RESEP_LAUNCH_CHECK(h, "k_post2_tc");
  return RESEP_OK;
}

}  // namespace resep
