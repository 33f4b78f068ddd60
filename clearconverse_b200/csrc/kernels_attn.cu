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
// Warp roles (8 warps, one CTA per SM, persistent over (chunk, head-group) items):
//   0      TMA producer: one [160 x 128 B] box per (chunk, head) into a ring of NU units
//   1      MMA issuer (whole warp runs the schedule, elect.sync lane issues) + TMEM owner
//   2      key norms: max_j |k_j|^2 of every unit (the stabiliser's second factor)
//   3      spare (keeps the softmax warps' TMEM lane quarters = warp % 4)
//   4..7   softmax + epilogue, lane quarter = warp % 4, thread = query row
// Tensor memory (512 columns): a ring of 4 score blocks [128 x 80] fp32 (P overwrites the first 40 columns of its
// block), two [128 x 16] output accumulators for the head blocks (alternating per head), four for the tail tile, and
// the four 8-column QT operands.  All tcgen05.mma of the CTA come from one thread and execute in order, so a score
// block is recycled by program order alone (S(n+4) is issued after P.V(n)); only "P ready" and "O drained" need barriers.
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "tc_common.cuh"

namespace resep {

using namespace ptx;

namespace attn {
// TWO CTAs share an SM (256 TMEM columns, < 113 KB of shared memory and <= 128 registers each): a softmax warp is
// latency-bound on its own (one warp per sub-partition ran at 0.45 IPC and 1,700 cycles per block in the first
// single-CTA versions, scripts/gpu_attn_trace.py), two per sub-partition fill each other's waits.
constexpr int NU = 5;                          // (chunk, head) units in flight: one group of four heads + one prefetched
constexpr int UNIT = 160 * 128;                // one TMA box: 160 rows x 128 B (q | k | v of the head + 32 B of the next head)
constexpr int OFF_KMAX = NU * UNIT;            // float[NU]
constexpr int OFF_BAR = OFF_KMAX + 64;
constexpr int THREADS = 256;
constexpr int KB = 80;                         // keys per score block
constexpr int RING = 2, LOOK = 1;              // score blocks in tensor memory / how far S runs ahead of P.V
constexpr int TMEM_COLS = 256;
// score ring | head-block accumulators (two, alternating per head) | tail accumulators (four heads); the QT operands
// alias the upper half of the tail accumulators: their last reader (the tail's second score MMA) precedes the first
// tail P.V MMA in the in-order tensor pipe, and the softmax threads rewrite them only after the tail epilogue
constexpr int TM_S = 0, TM_O0 = RING * KB, TM_O1 = TM_O0 + 32, TM_QT = TM_O1 + 32;   // 0, 160, 192, 224 (.. 256)
constexpr int B_UNIT_FULL = 0, B_UNIT_EMPTY = NU, B_NORM_FULL = 2 * NU, B_S_FULL = 3 * NU, B_P_FULL = B_S_FULL + RING,
              B_O0_FULL = B_P_FULL + RING /* [2] */, B_O0_FREE = B_O0_FULL + 2 /* [2] */, B_O1_FULL = B_O0_FREE + 2,
              B_O1_FREE = B_O1_FULL + 1, B_QT_FULL = B_O1_FREE + 1, NBAR = B_QT_FULL + 1;
constexpr int SMEM = OFF_BAR + NBAR * 8 + 16;
static_assert(2 * (SMEM + 1024) <= 227 * 1024, "two CTAs per SM");
constexpr float MAGIC = 12582912.f;            // 1.5 * 2^23: adding it rounds to an integer in the low mantissa bits
constexpr float M_LIMIT = 57.f;                // stabiliser above this: exact row maxima instead (2 m < 115 keeps 2^(s - m) normal)
}  // namespace attn

struct AttnArgs {
  bf16* ctx;            // [rows, 128] bf16
  int n_chunks;
  int log2_hpi;         // heads per work item = 1 << log2_hpi (8, 4, 2 or 1): small batches split a chunk's heads over several CTAs
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
__device__ __forceinline__ float attn_sqnorm16(const uint4& a, const uint4& b) {
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  __nv_bfloat162 acc = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
    acc = __hfma2(v, v, acc);
  }
  const float2 f = __bfloat1622float2(acc);
  return f.x + f.y;
}
__device__ __forceinline__ float attn_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float attn_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// p = 2^(s - m) for a pair of scores, MUFU path
__device__ __forceinline__ uint32_t attn_exp_mufu(uint32_t sa, uint32_t sb, float2 negm, float2& lsum) {
  const float2 x = fadd2(make_float2(__uint_as_float(sa), __uint_as_float(sb)), negm);
  const float2 e = make_float2(ex2_approx(x.x), ex2_approx(x.y));
  lsum = fadd2(lsum, e);
  return pack_bf16(e.x, e.y);
}
// the same on the FMA pipe: s + (MAGIC - m) rounds s to the nearest integer n_s (m is an integer, so the constant is
// exact); 2^(s - m) = 2^(n_s - m) * 2^f with f = s - n_s in [-0.5, 0.5]; 2^f by a cubic (relative error 7.5e-5, the
// bf16 rounding that follows is 2e-3); the integer part goes straight into the exponent field
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
  return pack_bf16(e.x, e.y);
}

// One score block [this thread's row x 80 keys] -> P (packed bf16) written over it.  TAIL: the block holds keys
// 80..159 of a 150-key chunk, its last 10 columns are padding (the next chunk's rows inside the TMA box): P = 0.
// POLY: odd pairs take the polynomial (only with the integer Cauchy-Schwarz stabiliser, which bounds the exponent).
template <bool TAIL, bool POLY>
__device__ __forceinline__ void attn_softmax_block(uint32_t sbuf, float m, float2& lsum) {
  const float2 negm = make_float2(-m, -m);
  const float2 c = make_float2(attn::MAGIC - m, attn::MAGIC - m), negc = make_float2(m - attn::MAGIC, m - attn::MAGIC);
  auto pairs16 = [&](const uint32_t (&v)[32], uint32_t (&p)[16]) {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      p[j] = (POLY && (j & 1)) ? attn_exp_poly(v[2 * j], v[2 * j + 1], c, negc, lsum) : attn_exp_mufu(v[2 * j], v[2 * j + 1], negm, lsum);
  };
  // Three column groups; the next group's load is in flight while the current one is exponentiated.  P group g
  // overwrites score columns that groups <= g have already put in registers.
  uint32_t va[32], vb[32];
  tmem_ld32(sbuf, va);
  tmem_ld_wait();
  tmem_ld32(sbuf + 32, vb);
  {
    uint32_t p[16];
    pairs16(va, p);
    tmem_ld_wait();
    tmem_st16(sbuf, p);                   // P columns 0..15 <- keys 0..31
  }
  uint32_t vc[16];
  tmem_ld16(sbuf + 64, vc);
  {
    uint32_t p[16];
    pairs16(vb, p);
    tmem_ld_wait();
    tmem_st16(sbuf + 16, p);              // keys 32..63
  }
  {
    uint32_t p[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (TAIL && j >= 3) p[j] = 0u;      // keys 150..159
      else p[j] = (POLY && (j & 1)) ? attn_exp_poly(vc[2 * j], vc[2 * j + 1], c, negc, lsum) : attn_exp_mufu(vc[2 * j], vc[2 * j + 1], negm, lsum);
    }
    tmem_st8(sbuf + 32, p);               // keys 64..79
  }
}

// row maximum of one score block (fallback path)
template <bool TAIL>
__device__ __forceinline__ float attn_block_max(uint32_t sbuf, float mx) {
  {
    uint32_t v[32];
    tmem_ld32(sbuf, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
    tmem_ld32(sbuf + 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
  }
  uint32_t v2[16];
  tmem_ld16(sbuf + 64, v2);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < (TAIL ? 6 : 16); ++j) mx = fmaxf(mx, __uint_as_float(v2[j]));
  return mx;
}

template <bool POLY>
__global__ void __launch_bounds__(attn::THREADS, 2) k_attn_tc(const __grid_constant__ CUtensorMap tmQKV, const AttnArgs args) {
  using namespace attn;
  constexpr uint32_t IDESC_S = umma_idesc(UMMA_BF16, UMMA_BF16, 128, KB);
  constexpr uint32_t IDESC_PV = umma_idesc(UMMA_BF16, UMMA_BF16, 128, 16) | UMMA_IDESC_B_MN_MAJOR;
  extern __shared__ __align__(1024) uint8_t smem[];
  float* s_kmax = reinterpret_cast<float*>(smem + OFF_KMAX);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBAR);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int atr_n = 0;
  (void)atr_n;
  pdl_trigger();

  // work items: (chunk, group of hpi heads), hpi a power of two; an item's heads are processed in groups of
  // gs = min(4, hpi) that share a tail tile.  Block stream of one group: head blocks (head 0 half 0, head 0 half 1,
  // head 1 half 0, ...), then the tail tile's two halves.
  const int lh = args.log2_hpi, hpi = 1 << lh, li = 3 - lh;  // heads per item; items per chunk = 1 << li
  const int gs = hpi < 4 ? hpi : 4, bpg = 2 * gs + 2;        // heads per group, blocks per group
  const int n_items = args.n_chunks << li;
  const int my_items = (int)blockIdx.x < n_items ? (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int total_units = my_items << lh, total_groups = total_units / gs, total_blocks = total_groups * bpg;
  auto item_of = [&](int u) { return (int)blockIdx.x + (u >> lh) * (int)gridDim.x; };
  auto chunk_of = [&](int u) { return item_of(u) >> li; };
  auto head_of = [&](int u) { return ((item_of(u) & ((1 << li) - 1)) << lh) + (u & (hpi - 1)); };

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQKV);
    for (int i = 0; i < NU; ++i) { mbar_init(&bars[B_UNIT_FULL + i], 1); mbar_init(&bars[B_UNIT_EMPTY + i], 1); mbar_init(&bars[B_NORM_FULL + i], 1); }
    for (int i = 0; i < RING; ++i) { mbar_init(&bars[B_S_FULL + i], 1); mbar_init(&bars[B_P_FULL + i], 4); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars[B_O0_FULL + i], 1); mbar_init(&bars[B_O0_FREE + i], 4); }
    mbar_init(&bars[B_O1_FULL], 1); mbar_init(&bars[B_O1_FREE], 4); mbar_init(&bars[B_QT_FULL], 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t smem_base = smem_u32(smem);

  pdl_wait();   // qkv comes from the previous kernel in the stream; ctx is still being read by the kernel before it
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int u = 0; u < total_units; ++u) {
        const int slot = u % NU;
        mbar_wait(&bars[B_UNIT_EMPTY + slot], ((u / NU) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars[B_UNIT_FULL + slot], UNIT);
        tma_load_2d(smem + slot * UNIT, &tmQKV, &bars[B_UNIT_FULL + slot], head_of(u) * QKV_HEAD_STRIDE, chunk_of(u) * CHUNK);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    struct Cur { int n, G, i; };              // block n = block i of group G
    auto next = [&](Cur& c) { ++c.n; if (++c.i == bpg) { c.i = 0; ++c.G; } };
    auto issue_s = [&](const Cur& c) {       // scores of block c.n into ring buffer n & 1
      const int n = c.n, i = c.i, u0 = c.G * gs;
      const uint32_t d = tmem + TM_S + KB * (n & 1);
      if (i < 2 * gs) {                       // head block: rows 0..127 of head u0 + i / 2, key half i % 2 (SS form)
        const int u = u0 + (i >> 1), kh = i & 1, slot = u % NU;
        if (kh == 0) {
          mbar_wait(&bars[B_UNIT_FULL + slot], (u / NU) & 1);
          tc_fence_after();
        }
        const uint32_t ub = smem_base + slot * UNIT;
        const uint64_t adesc = umma_desc_k_sw128(ub);
        const uint64_t bdesc = umma_desc_k_sw128(ub + kh * KB * 128 + 32);
        if (elect_one()) {
          umma_bf16(d, adesc, bdesc, IDESC_S, false);
          umma_commit(&bars[B_S_FULL + (n & 1)]);
        }
      } else {                                // tail block: sum over the group's heads of QT_j . K_j^T (TS form, A in tensor memory)
        const int kh = i - 2 * gs;
        if (kh == 0) {
          mbar_wait(&bars[B_QT_FULL], c.G & 1);
          tc_fence_after();
        }
        if (elect_one()) {
          for (int j = 0; j < gs; ++j) {
            const uint32_t ub = smem_base + ((u0 + j) % NU) * UNIT;
            umma_bf16_ts(d, tmem + TM_QT + 8 * j, umma_desc_k_sw128(ub + kh * KB * 128 + 32), IDESC_S, j > 0);
          }
          umma_commit(&bars[B_S_FULL + (n & 1)]);
        }
      }
      __syncwarp();
    };
    auto issue_pv = [&](const Cur& c) {      // O (+)= P(block n) . V(keys of block n)
      const int n = c.n, i = c.i, u0 = c.G * gs;
      ATR(0, 2000 + n);
      mbar_wait(&bars[B_P_FULL + (n & 1)], (n >> 1) & 1);
      ATR(0, 2500 + n);
      const uint32_t pa = tmem + TM_S + KB * (n & 1);
      if (i < 2 * gs) {
        const int u = u0 + (i >> 1), kh = i & 1, slot = u % NU, ob = u & 1;
        if (kh == 0) mbar_wait(&bars[B_O0_FREE + ob], ((u >> 1) & 1) ^ 1);   // the epilogue of unit u - 2 has the accumulator in registers
        tc_fence_after();
        const uint32_t vb = smem_base + slot * UNIT + kh * KB * 128 + 64;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < KB / 16; ++ks)
            umma_bf16_ts(tmem + TM_O0 + 16 * ob, pa + 8 * ks, umma_desc_mn_sw128(vb + ks * 16 * 128), IDESC_PV, !(kh == 0 && ks == 0));
          if (kh == 1) umma_commit(&bars[B_O0_FULL + ob]);
        }
      } else {
        // (the tail accumulators were drained before the group's QT operands were written, which the tail's score
        // MMAs have already waited for: no "free" barrier needed)
        const int kh = i - 2 * gs;
        tc_fence_after();
        if (elect_one()) {
          for (int j = 0; j < gs; ++j) {     // every lane gets P . V_j; quarter j's lanes are head j's rows
            const uint32_t vb = smem_base + ((u0 + j) % NU) * UNIT + kh * KB * 128 + 64;
#pragma unroll
            for (int ks = 0; ks < KB / 16; ++ks)
              umma_bf16_ts(tmem + TM_O1 + 16 * j, pa + 8 * ks, umma_desc_mn_sw128(vb + ks * 16 * 128), IDESC_PV, !(kh == 0 && ks == 0));
          }
          if (kh == 1) {
            umma_commit(&bars[B_O1_FULL]);
            for (int j = 0; j < gs; ++j) umma_commit(&bars[B_UNIT_EMPTY + (u0 + j) % NU]);   // every MMA that reads the group's units has been issued
          }
        }
      }
      __syncwarp();
      ATR(0, 3000 + n);
    };
    if (total_blocks > 0) {
      Cur cs{0, 0, 0}, cp{0, 0, 0};
      issue_s(cs);
      next(cs);
#pragma unroll 1
      while (cp.n < total_blocks) {
        if (cs.n < total_blocks) {            // into the block P.V(n - 1) has read (in-order tensor pipe)
          issue_s(cs);
          next(cs);
        }
        issue_pv(cp);
        next(cp);
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ key norms: max_j |k_j|^2 over the chunk's 150 keys
    for (int u = 0; u < total_units; ++u) {
      const int slot = u % NU;
      mbar_wait(&bars[B_UNIT_FULL + slot], (u / NU) & 1);
      const uint8_t* ub = smem + slot * UNIT;
      float kn2 = 0.f;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const int r = lane + 32 * i;
        if (r < CHUNK) {
          const uint4 ka = *reinterpret_cast<const uint4*>(ub + r * 128 + ((2 ^ (r & 7)) << 4));
          const uint4 kb = *reinterpret_cast<const uint4*>(ub + r * 128 + ((3 ^ (r & 7)) << 4));
          kn2 = fmaxf(kn2, attn_sqnorm16(ka, kb));
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) kn2 = fmaxf(kn2, __shfl_xor_sync(0xffffffffu, kn2, o));
      if (lane == 0) {
        s_kmax[slot] = kn2;               // the previous use of the slot was read before its unit was released
        mbar_arrive(&bars[B_NORM_FULL + slot]);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax + epilogue: thread = query row = TMEM lane
    const int q = warp & 3;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    const int row0 = q * 32 + lane;          // this thread's row in the head blocks
    const int row1 = 128 + lane;             // ... and in the tail tile (head u0 + q), valid for lane < 22
    // stabiliser of one query row against one unit: m = ceil(1.02 |q| max_j |k_j|) (Cauchy-Schwarz; 1.02 covers the
    // packed-bf16 norm roundings; an integer keeps the polynomial's rounding constant exact)
    auto stabiliser = [&](const uint4& qa, const uint4& qb, int slot) -> float {
      return ceilf(attn_sqrt(attn_sqnorm16(qa, qb) * s_kmax[slot]) * 1.02f);
    };
    // one block: wait for its scores, P over them, signal
    auto block = [&](int n, float m, bool fallback, bool tail_half, float2& lsum) {
      if (warp == 4) ATR(1, 1000 + n);
      mbar_wait(&bars[B_S_FULL + (n & 1)], (n >> 1) & 1);
      tc_fence_after();
      if (warp == 4) ATR(1, 2000 + n);
      const uint32_t sbuf = lane_base + TM_S + KB * (n & 1);
      if (fallback) {
        if (!tail_half) attn_softmax_block<false, false>(sbuf, m, lsum); else attn_softmax_block<true, false>(sbuf, m, lsum);
      } else {
        if (!tail_half) attn_softmax_block<false, POLY>(sbuf, m, lsum); else attn_softmax_block<true, POLY>(sbuf, m, lsum);
      }
      if (warp == 4) ATR(1, 3000 + n);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_P_FULL + (n & 1)]);
      if (warp == 4) ATR(1, 4000 + n);
    };
    // exact row maximum over both halves (blocks n, n + 1): the path for scores outside the stabiliser's range
    auto exact_max = [&](int n) -> float {
      mbar_wait(&bars[B_S_FULL + (n & 1)], (n >> 1) & 1);
      mbar_wait(&bars[B_S_FULL + ((n + 1) & 1)], ((n + 1) >> 1) & 1);
      tc_fence_after();
      float mx = attn_block_max<false>(lane_base + TM_S + KB * (n & 1), -INFINITY);
      return attn_block_max<true>(lane_base + TM_S + KB * ((n + 1) & 1), mx);
    };
    auto store_row = [&](const uint32_t (&o)[16], float l, int u, int row) {
      const float inv = attn_rcp(l);
      uint32_t w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] = pack_bf16(__uint_as_float(o[2 * j]) * inv, __uint_as_float(o[2 * j + 1]) * inv);
      uint4* dst = reinterpret_cast<uint4*>(args.ctx + ((int64_t)chunk_of(u) * CHUNK + row) * D + head_of(u) * DH);
      dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
      dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
    };
    // A head's output is read one block after its last P was handed over: the P.V MMAs run under the next block's
    // exponentials.  The tail tile's output is read at the start of the next group (its columns hold the next QT).
    int pend_kind = 0, pend_id = 0;          // 1: head unit pend_id, 2: tail tile of group pend_id
    float pend_l = 1.f;
    auto flush = [&]() {
      if (pend_kind == 1) {
        const int u = pend_id, ob = u & 1;
        mbar_wait(&bars[B_O0_FULL + ob], (u >> 1) & 1);
        tc_fence_after();
        uint32_t o[16];
        tmem_ld16(lane_base + TM_O0 + 16 * ob, o);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[B_O0_FREE + ob]);
        store_row(o, pend_l, u, row0);
      } else if (pend_kind == 2) {
        const int G = pend_id;
        mbar_wait(&bars[B_O1_FULL], G & 1);
        tc_fence_after();
        uint32_t o[16];
        tmem_ld16(lane_base + TM_O1 + 16 * q, o);
        tmem_ld_wait();
        if (q < gs && lane < CHUNK - 128) store_row(o, pend_l, G * gs + q, row1);
      }
      pend_kind = 0;
      if (warp == 4) ATR(1, 5000);
    };
#pragma unroll 1
    for (int G = 0; G < total_groups; ++G) {
      const int u0 = G * gs, nb = G * bpg;
      float m1 = 0.f;
      const bool valid1 = q < gs && lane < CHUNK - 128;
      // this quarter's tail rows (rows 128..149 of head u0 + q) become the QT_q operand; zeros in the other three
      auto write_qt = [&](int slot) {
        uint4 qa = make_uint4(0, 0, 0, 0), qb = qa;
        if (valid1) {
          const uint8_t* ub = smem + slot * UNIT;
          qa = *reinterpret_cast<const uint4*>(ub + row1 * 128 + ((0 ^ (row1 & 7)) << 4));
          qb = *reinterpret_cast<const uint4*>(ub + row1 * 128 + ((1 ^ (row1 & 7)) << 4));
          m1 = stabiliser(qa, qb, slot);
        }
        const uint32_t w[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
        uint32_t z[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) z[j] = (j >> 3) == q ? w[j & 7] : 0u;
        tmem_st32(lane_base + TM_QT, z);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[B_QT_FULL]);
      };
      if (pend_kind == 2) flush();           // the previous tail's accumulators share columns with the QT operands
      if (q >= gs) write_qt(0);              // (no head for this quarter: all zeros)
      // ---- head blocks
#pragma unroll 1
      for (int hg = 0; hg < gs; ++hg) {
        const int u = u0 + hg, slot = u % NU, n0 = nb + 2 * hg;
        mbar_wait(&bars[B_UNIT_FULL + slot], (u / NU) & 1);
        mbar_wait(&bars[B_NORM_FULL + slot], (u / NU) & 1);
        if (hg == q) {
          if (pend_kind == 2) flush();
          write_qt(slot);
        }
        const uint8_t* ub = smem + slot * UNIT;
        const uint4 qa = *reinterpret_cast<const uint4*>(ub + row0 * 128 + ((0 ^ (row0 & 7)) << 4));
        const uint4 qb = *reinterpret_cast<const uint4*>(ub + row0 * 128 + ((1 ^ (row0 & 7)) << 4));
        float m = stabiliser(qa, qb, slot);
        const bool fallback = __any_sync(0xffffffffu, !(m < M_LIMIT));     // also taken for NaN / inf inputs
        if (fallback) m = exact_max(n0);
        float2 lsum = make_float2(0.f, 0.f);
        block(n0, m, fallback, false, lsum);
        if (pend_kind == 1) flush();
        block(n0 + 1, m, fallback, true, lsum);
        pend_kind = 1; pend_id = u; pend_l = lsum.x + lsum.y;
      }
      // ---- tail tile: quarter q holds rows 128..149 of head u0 + q
      {
        const int n0 = nb + 2 * gs;
        const bool fallback = __any_sync(0xffffffffu, valid1 && !(m1 < M_LIMIT));
        if (fallback) m1 = exact_max(n0);
        float2 lsum = make_float2(0.f, 0.f);
        block(n0, m1, fallback, false, lsum);
        if (pend_kind == 1) flush();
        block(n0 + 1, m1, fallback, true, lsum);
        pend_kind = 2; pend_id = G; pend_l = lsum.x + lsum.y;
      }
    }
    flush();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
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
  auto kern = poly ? k_attn_tc<true> : k_attn_tc<false>;
  RESEP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::SMEM));
  // a chunk's heads are independent: small batches split them over more CTAs (the result does not depend on the split)
  const int slots = 2 * h->sm_count;       // two CTAs per SM
  // Items of four heads (one tail tile each) balance better over the 296 CTA slots than whole chunks (432 chunks = 864
  // items = 2.92 per CTA: three rounds at 97 %); fewer heads per item only when there are not enough items to go round.
  int lh = 2;
  while (lh > 0 && (n_chunks << (3 - lh)) * 2 <= slots) --lh;
  static long long* trace_buf = nullptr;
  if (getenv("RESEP_TRACE") && !trace_buf) { cudaMalloc(&trace_buf, 2048 * 8); cudaMemset(trace_buf, 0, 2048 * 8); g_attn_trace = trace_buf; }
  AttnArgs a{ctx, n_chunks, lh, trace_buf};
  const int items = n_chunks << (3 - lh);
  const int grid = items < slots ? items : slots;
  RESEP_CUDA(h, launch_pdl(kern, dim3((unsigned)grid), dim3(attn::THREADS), attn::SMEM, st, tmQKV, a));
  RESEP_LAUNCH_CHECK(h, "k_attn_tc");
  return RESEP_OK;
}

}  // namespace resep
