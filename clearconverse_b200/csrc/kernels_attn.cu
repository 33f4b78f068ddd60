// k_attn_tc: multi-head self-attention of the intra-chunk transformer on tcgen05 / TMEM / TMA.
//
//     ctx[c, i, h] = softmax_j(q[c,i,h] . k[c,j,h] / sqrt(16)) v[c,j,h]      chunk c of 150 rows, 8 heads of 16
//
// (the nn.MultiheadAttention call of speechbrain's TransformerEncoderLayer, restated at
// oracle/resepformer_oracle.py MultiheadAttention.forward; call site /root/reference/back/api.py:1077).  Input is the
// bf16 in-projection buffer of k_qkv2_tc: rows head-interleaved [head][q16 | k16 | v16], keys pre-scaled by
// log2(e) / 4 so that q . k is the softmax exponent in base 2.
//
// Why this shape.  With head_dim 16 the contraction work is tiny; what costs is the softmax: 150 x 160 x 8
// exponentials per chunk.  The mma.sync predecessor (k_attention_bf16_tma) held the warp's dispatch port for every
// HMMA and was bound by the SUM of its HMMA, MUFU and packing instructions (33 us per 64,800-row launch, XU pipe
// 74 % busy).  Here the two contractions are asynchronous tcgen05.mma instructions issued by one thread:
//   S = Q K^T   SS form: A = the head's q slice, B = its k slice, both K-major 32-byte slices INSIDE one 128B-swizzled
//               [160 rows x 128 B] TMA box (start address advanced by the slice's byte offset); M = 128, N = 80
//               (one key half), K = 16: one MMA per block, 52 cycles;
//   O = P V     TS form: A = P, packed bf16 in tensor memory, written there by the softmax threads over the scores
//               they just read; B = the head's v slice of the same box as an MN-major operand; N = 16, K = 16:
//               five MMAs of 13 cycles per block (measured: scripts/microbench/attn_probe.cu, profiles/r2_attn_probe.txt);
// and the softmax is thread-per-query-row (a TMEM lane is a row: no shuffles, no shared memory), with the
// Cauchy-Schwarz stabiliser m_i = ceil(1.02 |q_i| max_j |k_j|) known before any score exists, so the two key halves
// of a head are independent blocks and the pipeline never waits for a row maximum.  Half of the exponentials go
// through MUFU.EX2, the other half through a cubic on the FMA pipe (packed f32x2; Cody-Waite split with the rounding
// done by the add of 1.5 * 2^23 - m_i, which is exact because m_i is an integer): 16 -> 21.8 exp2 per clock per SM.
//
// A chunk has 150 query rows = one full M = 128 tile + 22 rows.  The 22 extra rows of FOUR heads share one 128-lane
// tile ("tail tile"): the softmax thread of lane quarter j copies q row 128 + lane of head 4g + j into a small A
// operand in tensor memory (QT_j: zero outside quarter j), and S_tail = sum_j QT_j K_j^T (four TS-form MMAs
// accumulating into the same columns) leaves head 4g + j's scores in quarter j.  One softmax pass then serves four
// heads, and P V_j into four 16-column accumulators gives each quarter its head's output.  (A first version gave the
// tail rows their own warps, one head at a time: that serial chain paced the whole kernel at 53 us per launch.)
// Every softmax warp therefore sees the same block stream: per group of four heads 8 head blocks + 2 tail blocks.
//
// Concurrency.  A softmax warp on its own is latency-bound: per 80-key block it waits for the scores (mbarrier), for
// the tensor-memory load, for the tensor-memory store and for the arrive -- about 800 cycles around 450 cycles of
// arithmetic (scripts/gpu_attn_trace.py; one warp per sub-partition ran at 0.45 IPC).  So the CTA runs THREE softmax
// warpgroups (12 warps, three per SM sub-partition) over a round-robin of jobs (the four heads and the tail tile of a
// group: five jobs of two blocks), each warpgroup with its own issuer warp, score block and P block.  P does NOT
// overwrite the scores: a warpgroup pulls a whole score block into registers and releases it at once, so its issuer
// puts the next block's S on the tensor pipe while the exponentials of this one are still being computed, and the
// P.V of block n follows S(n + 1) -- the warpgroup never waits for its issuer.  (Versions that reused the score
// columns for P idled through every issuer turnaround: 36 us per launch with four warpgroups, 53 us with one issuer
// polling all four.)  tcgen05.mma of one thread execute in order, and a warpgroup's stream needs no ordering against
// the others', so three issuer warps need no scheduler.
//
// Warp roles (16 warps, one CTA per SM, persistent over (chunk, four heads) items):
//   0..11  softmax + epilogue: warpgroup warp / 4, lane quarter = warp % 4, thread = query row
//   12..14 MMA issuer of warpgroup 0..2 (whole warp runs the schedule, elect.sync lane issues); it also computes
//          max_j |k_j|^2 of its head jobs' units (the stabiliser's second factor) before their scores are needed
//   15     TMA producer: one [160 x 128 B] box per (chunk, head) into a ring of NU = 8 units (two groups); TMEM owner
// The issuers have the HIGHEST warp ids on purpose: the sub-partition arbiter prefers the highest eligible warp id,
// and an issuer placed below the ever-eligible softmax warps took ~2,000 cycles for its ~150 instructions per block
// (scripts/gpu_attn_trace.py) -- longer than the block it was supposed to stay ahead of.
// Tensor memory (504 of 512 columns): per warpgroup a score block [128 x 80] fp32, a P block (40 columns of packed
// 16-bit pairs) and a [128 x 16] output accumulator; four accumulators for the tail tile; the four 8-column QT operands.
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "tc_common.cuh"

namespace resep {

using namespace ptx;

namespace attn {
constexpr int NWG = 3;                         // softmax warpgroups (each with its own issuer warp)
constexpr int NU = 8;                          // (chunk, head) units in flight: two groups of four heads
constexpr int UNIT = 160 * 128;                // one TMA box: 160 rows x 128 B (q | k | v of the head + 32 B of the next head)
constexpr int OFF_KMAX = NU * UNIT;            // float[NU]
constexpr int OFF_BAR = OFF_KMAX + 64;
constexpr int THREADS = 32 * (NWG + 1) + 128 * NWG;   // 512: three issuers, the TMA producer, 12 softmax warps
constexpr int KB = 80;                         // keys per score block
constexpr int TMEM_COLS = 512;
constexpr int TM_S = 0, TM_P = NWG * KB, TM_O0 = TM_P + NWG * (KB / 2), TM_O1 = TM_O0 + 16 * NWG, TM_QT = TM_O1 + 64;   // 0, 240, 360, 408, 472 (.. 504)
// Barriers.  Everything that belongs to a warpgroup's stream is per warpgroup, so every waiter sees every phase of
// the barriers it waits on (a parity is only enough for that).
// "Unit landed" and "unit's key norm ready" have waiters that do not look at every unit (an issuer sees its own jobs'
// units, a tail warp one unit per tail).  A parity wait is only right if the barrier's PREVIOUS phase is over when
// the wait starts, so these two get TWO barriers per ring slot, used alternately (index u % 16): the previous user of
// a barrier is then unit u - 16, and that one has landed before unit u - 8 could even be requested -- which every
// waiter knows has happened, because it has already worked on a unit of group G - 1 or later.
constexpr int B_UNIT_FULL = 0 /* [2 NU] */, B_NORM_FULL = 2 * NU /* [2 NU] */, B_UNIT_EMPTY = 4 * NU /* [NU], count 2: head user + tail user */,
              B_S_FULL = 5 * NU, B_S_FREE = B_S_FULL + NWG /* 4 warps */, B_P_FULL = B_S_FREE + NWG /* 4 warps */,
              B_P_FREE = B_P_FULL + NWG, B_O0_FULL = B_P_FREE + NWG, B_O0_FREE = B_O0_FULL + NWG /* 4 warps */,
              B_O1_FULL = B_O0_FREE + NWG /* [owner] */, B_O1_FREE = B_O1_FULL + NWG /* [owner], 4 warps */,
              B_QT_FULL = B_O1_FREE + NWG /* [owner], 4 warps */, NBAR = B_QT_FULL + NWG;
constexpr int SMEM = OFF_BAR + NBAR * 8 + 16;
static_assert(SMEM <= 227 * 1024, "shared memory budget");
static_assert(TM_QT + 32 <= TMEM_COLS, "tensor memory budget");
constexpr float MAGIC = 12582912.f;            // 1.5 * 2^23: adding it rounds to an integer in the low mantissa bits
constexpr float M_LIMIT = 57.f;                // stabiliser above this: online row maxima instead (2 m < 115 keeps 2^(s - m) normal)
}  // namespace attn

struct AttnArgs {
  bf16* ctx;            // [rows, 128] bf16
  int n_chunks;
  long long* trace;     // development aid (RESEP_TRACE + a -DRESEP_TRACE_BUILD build): [2 roles][512] (tag, clock) pairs of CTA 0
};
long long* g_attn_trace = nullptr;
#ifndef RESEP_TRACE_BUILD
#define ATR(role, tag) do { } while (0)
#else
#define ATR(role, tag)                                                                    \
  do {                                                                                    \
    if (args.trace != nullptr && blockIdx.x == 0 && lane == 0 && atr_n < 512) {           \
      args.trace[(role) * 1024 + 2 * atr_n] = (tag);                                      \
      args.trace[(role) * 1024 + 2 * atr_n + 1] = clock64();                              \
      ++atr_n;                                                                            \
    }                                                                                     \
  } while (0)
#endif

// squared norm of 16 bf16 values in packed bf16 arithmetic (an upper bound is all the stabiliser needs; the 1.02
// margin covers the roundings)
template <bool F16>
__device__ __forceinline__ float attn_sqnorm16(const uint4& a, const uint4& b) {
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint32_t acc = 0u;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc = hfma2_sq<F16>(w[i], acc);
  const float2 f = unpack16<F16>(acc);
  return f.x + f.y;
}
__device__ __forceinline__ float attn_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float attn_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// p = 2^(s - m) for a pair of scores, MUFU path
template <bool F16>
__device__ __forceinline__ uint32_t attn_exp_mufu(uint32_t sa, uint32_t sb, float2 negm, float2& lsum) {
  const float2 x = fadd2(make_float2(__uint_as_float(sa), __uint_as_float(sb)), negm);
  const float2 e = make_float2(ex2_approx(x.x), ex2_approx(x.y));
  lsum = fadd2(lsum, e);
  return pack16<F16>(e.x, e.y);
}
// the same on the FMA pipe: s + (MAGIC - m) rounds s to the nearest integer n_s (m is an integer, so the constant is
// exact); 2^(s - m) = 2^(n_s - m) * 2^f with f = s - n_s in [-0.5, 0.5]; 2^f by a cubic (relative error 7.5e-5, the
// bf16 rounding that follows is 2e-3); the integer part goes straight into the exponent field
template <bool F16>
__device__ __forceinline__ uint32_t attn_exp_poly(uint32_t sa, uint32_t sb, float2 c, float2 negc, float2& lsum) {
  const float2 s = make_float2(__uint_as_float(sa), __uint_as_float(sb));
  const float2 r = fadd2(s, c);
  const float2 ns = fadd2(r, negc);
  const float2 f = ffma2(ns, make_float2(-1.f, -1.f), s);
  float2 p = ffma2(f, make_float2(0.05517167f, 0.05517167f), make_float2(0.24261113f, 0.24261113f));
  p = ffma2(p, f, make_float2(0.69326097f, 0.69326097f));
  p = ffma2(p, f, make_float2(0.99992806f, 0.99992806f));
  const float2 e = make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(r.x) << 23)),
                               __int_as_float(__float_as_int(p.y) + (__float_as_int(r.y) << 23)));
  lsum = fadd2(lsum, e);
  return pack16<F16>(e.x, e.y);
}

// One score block [this thread's row x 80 keys], already in registers, -> P (packed 16-bit pairs), also in registers:
// the exponentials do not wait for the P block of tensor memory to come free (the previous block's P.V may still be
// reading it); only the three stores that follow do.  TAIL: the block holds keys 80..159 of a 150-key chunk, its last 10 columns are padding (the
// next chunk's rows inside the TMA box): P = 0.  POLY: odd pairs take the polynomial (only with the integer
// Cauchy-Schwarz stabiliser, which bounds the exponent).
template <bool TAIL, bool POLY, bool F16>
__device__ __forceinline__ void attn_softmax_block(const uint32_t (&v)[80], uint32_t (&p)[40], float m, float2& lsum) {
  const float2 negm = make_float2(-m, -m);
  const float2 c = make_float2(attn::MAGIC - m, attn::MAGIC - m), negc = make_float2(m - attn::MAGIC, m - attn::MAGIC);
#pragma unroll
  for (int i = 0; i < 40; ++i) {
    if (TAIL && i >= 35) p[i] = 0u;       // keys 150..159
    else p[i] = (POLY && (i & 1)) ? attn_exp_poly<F16>(v[2 * i], v[2 * i + 1], c, negc, lsum) : attn_exp_mufu<F16>(v[2 * i], v[2 * i + 1], negm, lsum);
  }
}
__device__ __forceinline__ void attn_store_p(uint32_t pbuf, const uint32_t (&p)[40]) {
  tmem_st16(pbuf, *reinterpret_cast<const uint32_t(*)[16]>(&p[0]));
  tmem_st16(pbuf + 16, *reinterpret_cast<const uint32_t(*)[16]>(&p[16]));
  tmem_st8(pbuf + 32, *reinterpret_cast<const uint32_t(*)[8]>(&p[32]));
}

// row maximum of one score block held in registers (online path)
template <bool TAIL>
__device__ __forceinline__ float attn_block_max(const uint32_t (&v)[80], float mx) {
#pragma unroll
  for (int j = 0; j < (TAIL ? 70 : 80); ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
  return mx;
}

template <bool POLY, bool F16>
__global__ void __launch_bounds__(attn::THREADS, 1) k_attn_tc(const __grid_constant__ CUtensorMap tmQKV, const AttnArgs args) {
  using namespace attn;
  constexpr uint32_t FMT = F16 ? UMMA_F16 : UMMA_BF16;
  constexpr uint32_t IDESC_S = umma_idesc(FMT, FMT, 128, KB);
  constexpr uint32_t IDESC_PV = umma_idesc(FMT, FMT, 128, 16) | UMMA_IDESC_B_MN_MAJOR;
  extern __shared__ __align__(1024) uint8_t smem[];
  float* s_kmax = reinterpret_cast<float*>(smem + OFF_KMAX);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBAR);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int atr_n = 0;
  (void)atr_n;
  pdl_trigger();

  // Work items ("groups"): (chunk, four heads).  Group G of this CTA is item blockIdx.x + G * gridDim.x; its units (one
  // per head) are u = 4 G .. 4 G + 3, in ring slot u % 8.  Jobs: J = 5 G + r; r < 4: the head blocks of unit 4 G + r
  // (key halves 0, 1); r == 4: the group's tail tile (two more blocks).  Warpgroup x takes the jobs J = x (mod 3).
  const int n_items = 2 * args.n_chunks;
  const int total_groups = (int)blockIdx.x < n_items ? (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int total_units = 4 * total_groups, total_jobs = 5 * total_groups;
  auto item_of = [&](int u) { return (int)blockIdx.x + (u >> 2) * (int)gridDim.x; };
  auto chunk_of = [&](int u) { return item_of(u) >> 1; };
  auto head_of = [&](int u) { return ((item_of(u) & 1) << 2) + (u & 3); };
  auto unit_bar = [&](int u) -> uint64_t* { return &bars[B_UNIT_FULL + (u & (2 * NU - 1))]; };
  auto norm_bar = [&](int u) -> uint64_t* { return &bars[B_NORM_FULL + (u & (2 * NU - 1))]; };
  auto unit_par = [&](int u) -> uint32_t { return (uint32_t)(u / (2 * NU)) & 1; };

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQKV);
    for (int i = 0; i < 2 * NU; ++i) { mbar_init(&bars[B_UNIT_FULL + i], 1); mbar_init(&bars[B_NORM_FULL + i], 1); }
    for (int i = 0; i < NU; ++i) mbar_init(&bars[B_UNIT_EMPTY + i], 2);
    for (int i = 0; i < NWG; ++i) {
      mbar_init(&bars[B_S_FULL + i], 1); mbar_init(&bars[B_S_FREE + i], 4); mbar_init(&bars[B_P_FULL + i], 4); mbar_init(&bars[B_P_FREE + i], 1);
      mbar_init(&bars[B_O0_FULL + i], 1); mbar_init(&bars[B_O0_FREE + i], 4);
      mbar_init(&bars[B_O1_FULL + i], 1); mbar_init(&bars[B_O1_FREE + i], 4); mbar_init(&bars[B_QT_FULL + i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 4 * NWG + NWG) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t smem_base = smem_u32(smem);

  pdl_wait();   // qkv comes from the previous kernel in the stream; ctx is still being read by the kernel before it
  if (warp == 4 * NWG + NWG) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int u = 0; u < total_units; ++u) {
        const int slot = u & (NU - 1);
        mbar_wait(&bars[B_UNIT_EMPTY + slot], ((uint32_t)(u / NU) & 1) ^ 1);
        mbar_arrive_expect_tx(unit_bar(u), UNIT);
        tma_load_2d(smem + slot * UNIT, &tmQKV, unit_bar(u), head_of(u) * QKV_HEAD_STRIDE, chunk_of(u) * CHUNK);
      }
    }
    __syncwarp();
  } else if (warp >= 4 * NWG) {
    // ------------------------------------------------------------------ MMA issuer of warpgroup x.  Per block: wait until
    // the warpgroup has the previous scores in registers, put this block's S on the pipe, THEN the previous block's
    // P.V (its P arrives while the warpgroup already has these scores to work on).
    const int x = warp - 4 * NWG;
    const uint32_t sd = tmem + TM_S + KB * x, pa = tmem + TM_P + (KB / 2) * x;
    uint32_t nb = 0, nh = 0;                  // blocks issued; head jobs started (phases of the per-warpgroup barriers)
    int pend = -1, pend_G = 0, pend_r = 0;    // the block whose P.V is still to be issued: key half, group, job kind
    uint32_t pend_nb = 0, pend_nh = 0;
    auto do_pv = [&]() {
      const int kh = pend, u0 = 4 * pend_G;
      if (x == 0) ATR(0, 1000 + pend_nb);
      mbar_wait(&bars[B_P_FULL + x], pend_nb & 1);
      if (x == 0) ATR(0, 2000 + pend_nb);
      if (pend_r < 4) {
        const int u = u0 + pend_r, slot = u & (NU - 1);
        if (kh == 0) mbar_wait(&bars[B_O0_FREE + x], (pend_nh & 1) ^ 1);     // the warpgroup's previous head output is in registers
        tc_fence_after();
        const uint32_t vb = smem_base + slot * UNIT + kh * KB * 128 + 64;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < KB / 16; ++ks)
            umma_bf16_ts(tmem + TM_O0 + 16 * x, pa + 8 * ks, umma_desc_mn_sw128(vb + ks * 16 * 128), IDESC_PV, !(kh == 0 && ks == 0));
          umma_commit(&bars[B_P_FREE + x]);
          if (kh == 1) {
            umma_commit(&bars[B_O0_FULL + x]);
            umma_commit(&bars[B_UNIT_EMPTY + slot]);                         // the unit's head user is done (its tail user arrives too)
          }
        }
      } else {
        // (the previous tail's accumulators were drained before this tail's QT operands were written)
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {      // every lane gets P . V_j; quarter j's lanes are head j's rows
            const uint32_t vb = smem_base + ((u0 + j) & (NU - 1)) * UNIT + kh * KB * 128 + 64;
#pragma unroll
            for (int ks = 0; ks < KB / 16; ++ks)
              umma_bf16_ts(tmem + TM_O1 + 16 * j, pa + 8 * ks, umma_desc_mn_sw128(vb + ks * 16 * 128), IDESC_PV, !(kh == 0 && ks == 0));
          }
          umma_commit(&bars[B_P_FREE + x]);
          if (kh == 1) {
            umma_commit(&bars[B_O1_FULL + x]);
#pragma unroll
            for (int j = 0; j < 4; ++j) umma_commit(&bars[B_UNIT_EMPTY + ((u0 + j) & (NU - 1))]);
          }
        }
      }
      __syncwarp();
      pend = -1;
    };
#pragma unroll 1
    for (int J = x; J < total_jobs; J += NWG) {
      const int G = J / 5, r = J - 5 * G, u0 = 4 * G;
#pragma unroll 1
      for (int kh = 0; kh < 2; ++kh) {
        if (r < 4) {
          const int u = u0 + r, slot = u & (NU - 1);
          if (kh == 0) {
            mbar_wait(unit_bar(u), unit_par(u));
            // max_j |k_j|^2 over the unit's 150 keys (the warpgroup is still busy with its previous block)
            const uint8_t* up = smem + slot * UNIT;
            float kn2 = 0.f;
#pragma unroll
            for (int i = 0; i < 5; ++i) {
              const int rr = lane + 32 * i;
              if (rr < CHUNK) {
                const uint4 ka = *reinterpret_cast<const uint4*>(up + rr * 128 + ((2 ^ (rr & 7)) << 4));
                const uint4 kb = *reinterpret_cast<const uint4*>(up + rr * 128 + ((3 ^ (rr & 7)) << 4));
                kn2 = fmaxf(kn2, attn_sqnorm16<F16>(ka, kb));
              }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) kn2 = fmaxf(kn2, __shfl_xor_sync(0xffffffffu, kn2, o));
            if (lane == 0) {
              s_kmax[slot] = kn2;           // the previous use of the slot was read before its unit was released
              mbar_arrive(norm_bar(u));
            }
            __syncwarp();
          }
          if (nb > 0) mbar_wait(&bars[B_S_FREE + x], (nb - 1) & 1);          // the previous scores are in registers
          tc_fence_after();
          const uint32_t ub = smem_base + slot * UNIT;
          const uint64_t adesc = umma_desc_k_sw128(ub), bdesc = umma_desc_k_sw128(ub + kh * KB * 128 + 32);
          if (elect_one()) {
            umma_bf16(sd, adesc, bdesc, IDESC_S, false);
            umma_commit(&bars[B_S_FULL + x]);
          }
          __syncwarp();
        } else {
          if (kh == 0) {
            for (int j = 0; j < 4; ++j) mbar_wait(unit_bar(u0 + j), unit_par(u0 + j));   // (another warpgroup's head may not have been issued yet)
            mbar_wait(&bars[B_QT_FULL + x], (uint32_t)(G / 3) & 1);          // this warpgroup's (G / 3)-th tail
          }
          if (nb > 0) mbar_wait(&bars[B_S_FREE + x], (nb - 1) & 1);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {    // sum over the group's heads of QT_j . K_j^T (TS form)
              const uint32_t ub = smem_base + ((u0 + j) & (NU - 1)) * UNIT;
              umma_bf16_ts(sd, tmem + TM_QT + 8 * j, umma_desc_k_sw128(ub + kh * KB * 128 + 32), IDESC_S, j > 0);
            }
            umma_commit(&bars[B_S_FULL + x]);
          }
          __syncwarp();
        }
        if (pend >= 0) do_pv();
        pend = kh; pend_G = G; pend_r = r; pend_nb = nb; pend_nh = nh;
        ++nb;
      }
      if (r < 4) ++nh;
    }
    if (pend >= 0) do_pv();
  } else {
    // ------------------------------------------------------------------ softmax + epilogue: thread = query row = TMEM lane
    const int x = warp >> 2, q = warp & 3;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t sbuf = lane_base + TM_S + KB * x, pbuf = lane_base + TM_P + (KB / 2) * x;
    const int row0 = q * 32 + lane;          // this thread's row in the head blocks
    const int row1 = 128 + lane;             // ... and in the tail tile (head 4 G + q), valid for lane < 22
    const bool valid1 = lane < CHUNK - 128;
    const bool tracer = warp == 0;
    uint32_t nb = 0, nh = 0, nt = 0;         // blocks / head jobs / tail jobs of this warpgroup so far
    // stabiliser of one query row against one unit: m = ceil(1.02 |q| max_j |k_j|) (Cauchy-Schwarz; 1.02 covers the
    // packed 16-bit norm roundings; an integer keeps the polynomial's rounding constant exact)
    auto stabiliser = [&](const uint4& qa, const uint4& qb, int slot) -> float {
      return ceilf(attn_sqrt(attn_sqnorm16<F16>(qa, qb) * s_kmax[slot]) * 1.02f);
    };
    auto store_row = [&](const uint32_t (&o)[16], float l, int u, int row) {
      const float inv = attn_rcp(l);
      uint32_t wd[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) wd[j] = pack16<F16>(__uint_as_float(o[2 * j]) * inv, __uint_as_float(o[2 * j + 1]) * inv);
      uint4* dst = reinterpret_cast<uint4*>(args.ctx + ((int64_t)chunk_of(u) * CHUNK + row) * D + head_of(u) * DH);
      dst[0] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
      dst[1] = make_uint4(wd[4], wd[5], wd[6], wd[7]);
    };
    // A job's output is read after the FIRST hand-over of the warpgroup's next job: by then its last P.V has long completed.
    int pend_kind = 0, pend_u = 0;            // 1: head unit pend_u (O0), 2: tail row of unit pend_u (O1)
    uint32_t pend_idx = 0;
    float pend_l = 1.f;
    auto flush = [&]() {
      if (pend_kind == 0) return;
      uint32_t o[16];
      if (pend_kind == 1) {
        mbar_wait(&bars[B_O0_FULL + x], pend_idx & 1);
        tc_fence_after();
        tmem_ld16(lane_base + TM_O0 + 16 * x, o);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[B_O0_FREE + x]);
        store_row(o, pend_l, pend_u, row0);
      } else {
        mbar_wait(&bars[B_O1_FULL + x], pend_idx & 1);
        tc_fence_after();
        tmem_ld16(lane_base + TM_O1 + 16 * q, o);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[B_O1_FREE + x]);
        if (valid1) store_row(o, pend_l, pend_u, row1);
      }
      pend_kind = 0;
      if (tracer) ATR(1, 5000);
    };
    // Both key halves of one row set.  Ordinary path: the stabiliser m is known up front and the halves are independent.
    // Scores outside its range (`online`): running row maximum; before the second half's P is handed over the thread
    // rescales its own accumulator row (the first half's P.V has completed: its "P block free" commit has been seen).
    auto two_blocks = [&](float m, bool online, uint32_t o_cols, float& l_out) {
      float2 lsum = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int kh = 0; kh < 2; ++kh) {
        uint32_t v[80];
        if (tracer) ATR(1, 1000 + nb);
        mbar_wait(&bars[B_S_FULL + x], nb & 1);
        tc_fence_after();
        if (tracer) ATR(1, 2000 + nb);
        tmem_ld32(sbuf, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tmem_ld32(sbuf + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
        tmem_ld16(sbuf + 64, *reinterpret_cast<uint32_t(*)[16]>(&v[64]));
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[B_S_FREE + x]);          // the issuer may put the next scores here
        float sc = 1.f;
        if (online) {
          const float m_new = kh == 0 ? attn_block_max<false>(v, -INFINITY) : fmaxf(m, attn_block_max<true>(v, -INFINITY));
          if (kh == 1) sc = ex2_approx(m - m_new);
          m = m_new;
        }
        uint32_t pk[40];
        if (online) {
          if (kh == 1) lsum = make_float2(lsum.x * sc, lsum.y * sc);
          if (kh == 0) attn_softmax_block<false, false, F16>(v, pk, m, lsum); else attn_softmax_block<true, false, F16>(v, pk, m, lsum);
        } else {
          if (kh == 0) attn_softmax_block<false, POLY, F16>(v, pk, m, lsum); else attn_softmax_block<true, POLY, F16>(v, pk, m, lsum);
        }
        if (tracer) ATR(1, 2500 + nb);
        if (nb > 0) {                                             // the previous P.V has read the P block
          mbar_wait(&bars[B_P_FREE + x], (nb - 1) & 1);
          tc_fence_after();
        }
        if (online && kh == 1) {                                  // (the first half's P.V is complete: rescale its output)
          uint32_t o[16];
          tmem_ld16(o_cols, o);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * sc);
          tmem_st16(o_cols, o);
        }
        attn_store_p(pbuf, pk);
        if (tracer) ATR(1, 3000 + nb);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[B_P_FULL + x]);
        if (tracer) ATR(1, 4000 + nb);
        ++nb;
        if (kh == 0) flush();
      }
      l_out = lsum.x + lsum.y;
    };
#pragma unroll 1
    for (int J = x; J < total_jobs; J += NWG) {
      const int G = J / 5, r = J - 5 * G, u0 = 4 * G;
      if (r < 4) {
        // ---- head blocks of unit u0 + r
        const int u = u0 + r, slot = u & (NU - 1);
        mbar_wait(&bars[B_S_FULL + x], nb & 1);                   // (the scores exist: the unit has landed)
        mbar_wait(norm_bar(u), unit_par(u));
        const uint8_t* ub = smem + slot * UNIT;
        const uint4 qa = *reinterpret_cast<const uint4*>(ub + row0 * 128 + ((0 ^ (row0 & 7)) << 4));
        const uint4 qb = *reinterpret_cast<const uint4*>(ub + row0 * 128 + ((1 ^ (row0 & 7)) << 4));
        const float m = stabiliser(qa, qb, slot);
        const bool online = __any_sync(0xffffffffu, !(m < M_LIMIT));         // also taken for NaN / inf inputs
        float l;
        two_blocks(m, online, lane_base + TM_O0 + 16 * x, l);
        pend_kind = 1; pend_u = u; pend_idx = nh; pend_l = l;
        ++nh;
      } else {
        // ---- tail tile of the group: quarter q holds rows 128..149 of head u0 + q
        const int u = u0 + q, slot = u & (NU - 1);
        // (two barriers per slot: see B_UNIT_FULL)
        mbar_wait(unit_bar(u), unit_par(u));
        mbar_wait(norm_bar(u), unit_par(u));
        // the previous tail tile (group G - 1, warpgroup (5 (G - 1) + 4) % 3, its ((G - 1) / 3)-th tail) has drained the
        // tail accumulators, and its score MMAs are done with the QT operands
        if (G > 0) mbar_wait(&bars[B_O1_FREE + (5 * (G - 1) + 4) % NWG], (uint32_t)((G - 1) / 3) & 1);
        uint4 qa = make_uint4(0, 0, 0, 0), qb = qa;
        float m1 = 0.f;
        if (valid1) {
          const uint8_t* ub = smem + slot * UNIT;
          qa = *reinterpret_cast<const uint4*>(ub + row1 * 128 + ((0 ^ (row1 & 7)) << 4));
          qb = *reinterpret_cast<const uint4*>(ub + row1 * 128 + ((1 ^ (row1 & 7)) << 4));
          m1 = stabiliser(qa, qb, slot);
        }
        {
          const uint32_t wd[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
          uint32_t z[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) z[j] = (j >> 3) == q ? wd[j & 7] : 0u;
          tmem_st32(lane_base + TM_QT, z);                        // QT_q <- this quarter's tail rows; zeros in the other three
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[B_QT_FULL + x]);
        }
        const bool online = __any_sync(0xffffffffu, valid1 && !(m1 < M_LIMIT));
        float l;
        two_blocks(m1, online, lane_base + TM_O1 + 16 * q, l);
        pend_kind = 2; pend_u = u; pend_idx = nt; pend_l = l;
        ++nt;
      }
    }
    flush();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4 * NWG + NWG) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem);
  }
}

// Measured on a B200 (config 2, 64,800 rows per launch): k_attn_tc 42 us, k_attention_bf16_tma 33.5 us -- the mma.sync
// kernel stays the default until this one beats it (profiles/r2_attention.md).
bool attn_tc_default() { return false; }

int launch_attn_tc(ResepHandle* h, const bf16* qkv, bf16* ctx, int n_chunks, cudaStream_t st) {
  if (n_chunks <= 0) return RESEP_OK;
  ProfScope prof_scope(h, n_chunks >= 54 ? "k_attn_tc" : "k_attn_tc(small)", st);
  CUtensorMap tmQKV;
  int rc = make_tmap_head(h, &tmQKV, qkv, (int64_t)n_chunks * CHUNK, 3 * D, 160);
  if (rc) return rc;
  static const bool poly = !(getenv("RESEP_ATTN_POLY") && getenv("RESEP_ATTN_POLY")[0] == '0');
  auto kern = h->fmt16 ? (poly ? k_attn_tc<true, true> : k_attn_tc<false, true>) : (poly ? k_attn_tc<true, false> : k_attn_tc<false, false>);
  RESEP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::SMEM));
  static long long* trace_buf = nullptr;
  if (getenv("RESEP_TRACE") && !trace_buf) { cudaMalloc(&trace_buf, 2048 * 8); cudaMemset(trace_buf, 0, 2048 * 8); g_attn_trace = trace_buf; }
  AttnArgs a{ctx, n_chunks, trace_buf};
  // items of four heads (one tail tile each): 432 chunks = 864 items = 5.84 per CTA, six rounds at 97 %
  const int items = 2 * n_chunks;
  const int grid = items < h->sm_count ? items : h->sm_count;
  RESEP_CUDA(h, launch_pdl(kern, dim3((unsigned)grid), dim3(attn::THREADS), attn::SMEM, st, tmQKV, a));
  RESEP_LAUNCH_CHECK(h, "k_attn_tc");
  return RESEP_OK;
}

}  // namespace resep
