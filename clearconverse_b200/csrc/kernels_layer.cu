// k_post_tc: the post-attention half of one pre-norm TransformerEncoderLayer as ONE tcgen05 kernel
// (speechbrain Transformer.py TransformerEncoderLayer.forward, normalize_before=True):
//
//     o' = o + ctx . Wo^T + bo                       (out-proj + residual)
//     y  = LayerNorm(o'; g2, be2, eps 1e-6)          (norm2)
//     o  = o' + relu(y . W1^T + b1) . W2^T + b2      (pos_ffn + residual)
//
// One CTA owns a tile of 128 token rows at a time (persistent over tiles).  Nothing of the layer's
// interior touches HBM, and the activation operands do not even touch shared memory:
//   * the out-proj accumulator is rewritten in place with o' (tcgen05.st) and FFN2 accumulates
//     straight onto it, so the residual lives in TMEM for the whole tile;
//   * LN2(o') and every relu(FFN1) chunk are packed to bf16 by the epilogue warps and stored back to
//     TMEM, and FFN1 / FFN2 read them as the A operand FROM TENSOR MEMORY (tcgen05.mma [d],[a],b-desc).
//     With both operands in shared memory an M=128,N=128,K=16 MMA reads 8 KB of smem in 64 cycles
//     (= the whole 128 B/cycle port) while the weight TMA writes another 64 B/cycle: measured 99
//     cycles per MMA instead of 64.  With A in TMEM the port carries the B tile and the TMA only;
//   * the fp32 residual tile comes in and goes out through one 128B-swizzled shared tile moved by
//     TMA (a thread-per-row global access costs a full L1 wavefront per 16 bytes).
//
// Warp roles (11 warps): 0 = weight TMA producer (16 KB units: one [128 x 64] bf16 tile; the hi
// and lo parts of a split weight are separate units), 1 = MMA issuer + TMEM owner, 2 = tile
// producer (ctx tile + residual tile), 3..10 = epilogue (two threads per token row: TMEM lane ==
// row, each thread owns 64 of the 128 columns; LayerNorm partial sums go through shared memory).
//
// MMA issue order per tile (the weight producer streams tiles in exactly this order):
//     OUT | F1(0) F1(1) | F2(0) F1(2) | F2(1) F1(3) | ... | F2(5) F1(7) | F2(6) | F2(7)
// so the tensor pipe always has the next chunk's FFN1 queued while the epilogue converts the
// current chunk.
#include <cstdlib>

#include "tc_common.cuh"

namespace resep {

using namespace ptx;

namespace post {
constexpr int EPI_WARPS = 8;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int THREADS = 96 + EPI_THREADS;
constexpr int TILE = 128 * 128;          // bytes of one [128 rows x 128 B] swizzled atom (16 KB)
constexpr int NWST = 7;                  // weight stage units in flight
constexpr int NCHUNK = FFN / 128;        // 8 hidden chunks of 128
constexpr int OFF_CTX = 0;               // [128 x 128] bf16 ctx tile (2 atoms), A operand of the out-proj
constexpr int OFF_OT = OFF_CTX + 2 * TILE;   // [128 x 128] fp32 residual tile in / result tile out (4 atoms)
constexpr int OFF_W = OFF_OT + 4 * TILE;
constexpr int OFF_PAR = OFF_W + NWST * TILE;
constexpr int PAR_FLOATS = 4 * D + FFN;  // bo, g2, be2, b2, b1
constexpr int OFF_RED = OFF_PAR + PAR_FLOATS * 4;      // [2][128] floats: LN partial sums of the two column halves
constexpr int OFF_BAR = OFF_RED + 2 * 128 * 4;
constexpr int NBAR = 2 * NWST + 12;
constexpr int SMEM = OFF_BAR + NBAR * 8 + 16;
// TMEM columns: residual / result accumulator, two FFN1 chunk accumulators (each later holds its own
// relu'd bf16 copy in columns [0,32) and [64,96)), and the packed LN2 output
constexpr int TM_Y = 0, TM_H0 = 128, TM_H1 = 256, TM_Y2 = 384;
constexpr int TMEM_COLS = 512;
static_assert(SMEM <= 227 * 1024, "shared memory budget");
}  // namespace post

struct PostArgs {
  const float *bo, *g2, *be2, *b1, *b2;
  int64_t M;
  long long* trace;   // optional [128] clock stamps of CTA 0 (development aid; null in production)
};

// byte offset of the 16-byte chunk holding 4 fp32 columns [col, col+4) of `row` in the [128 x 128] fp32
// tile stored as four 128B-swizzled atoms of 32 columns (the layout TMA SWIZZLE_128B produces)
__device__ __forceinline__ uint32_t sw128_f32(int row, int col) {
  return (uint32_t)((col >> 5) * post::TILE + row * 128 + ((((col & 31) >> 2) ^ (row & 7)) << 4));
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(post::EPI_THREADS) : "memory"); }

template <bool SPLIT, bool SPLIT_FFN>
__global__ void __launch_bounds__(post::THREADS, 1)
k_post_tc(const __grid_constant__ CUtensorMap tmCtx, const __grid_constant__ CUtensorMap tmO,
          const __grid_constant__ CUtensorMap tmWo, const __grid_constant__ CUtensorMap tmWoL,
          const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW1L,
          const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmW2L, const PostArgs args) {
  using namespace post;
  constexpr int PARTS = SPLIT ? 2 : 1;          // out-proj weight operand: bf16 hi (+ lo)
  constexpr int FPARTS = SPLIT_FFN ? 2 : 1;     // FFN weight operands
  constexpr uint32_t IDESC = umma_idesc(UMMA_BF16, UMMA_BF16, 128, 128);

  // 128B-swizzled tiles need 1024-byte alignment; declaring it on the dynamic window keeps every access
  // below a true shared-memory (LDS/STS) access instead of a generic one.
  extern __shared__ __align__(1024) uint8_t smem[];
  float* par = reinterpret_cast<float*>(smem + OFF_PAR);
  float *s_bo = par, *s_g2 = par + D, *s_be2 = par + 2 * D, *s_b2 = par + 3 * D, *s_b1 = par + 4 * D;
  float* s_red = reinterpret_cast<float*>(smem + OFF_RED);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* w_full = bars;                 // [NWST] TMA -> MMA
  uint64_t* w_empty = bars + NWST;         // [NWST] MMA commit -> producer
  uint64_t* a_full = bars + 2 * NWST;      // ctx tile landed (TMA)
  uint64_t* a_empty = a_full + 1;          // out-proj MMAs retired -> ctx buffer free (MMA commit)
  uint64_t* o_full = a_full + 2;           // residual tile landed (TMA)
  uint64_t* ot_empty = a_full + 3;         // result tile has been read by its TMA store (1 arrival)
  uint64_t* y_full = a_full + 4;           // o' in TM_Y and LN2(o') in TM_Y2 (epilogue arrivals)
  uint64_t* hs_full = a_full + 5;          // [2] relu'd bf16 chunk stored into its accumulator columns (epilogue)
  uint64_t* acch_full = a_full + 7;        // [2] FFN1 chunk accumulator ready (MMA commit)
  uint64_t* accy_full = a_full + 9;        // out-proj accumulator ready
  uint64_t* accy_done = a_full + 10;       // all FFN2 of the tile retired
  uint64_t* accy_empty = a_full + 11;      // epilogue has the tile's result in registers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (int)((args.M + 127) / 128);

  for (int i = threadIdx.x; i < D; i += THREADS) {
    s_bo[i] = args.bo[i]; s_g2[i] = args.g2[i]; s_be2[i] = args.be2[i]; s_b2[i] = args.b2[i];
  }
  for (int i = threadIdx.x; i < FFN; i += THREADS) s_b1[i] = args.b1[i];
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmCtx); prefetch_tmap(&tmO); prefetch_tmap(&tmWo); prefetch_tmap(&tmW1); prefetch_tmap(&tmW2);
    if (SPLIT) prefetch_tmap(&tmWoL);
    if (SPLIT_FFN) { prefetch_tmap(&tmW1L); prefetch_tmap(&tmW2L); }
    for (int i = 0; i < NWST; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    mbar_init(a_full, 1); mbar_init(a_empty, 1); mbar_init(o_full, 1); mbar_init(ot_empty, 1);
    mbar_init(y_full, EPI_THREADS);
    for (int i = 0; i < 2; ++i) { mbar_init(&hs_full[i], EPI_THREADS); mbar_init(&acch_full[i], 1); }
    mbar_init(accy_full, 1); mbar_init(accy_done, 1); mbar_init(accy_empty, EPI_THREADS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ weight producer
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      auto put = [&](const CUtensorMap* hi, const CUtensorMap* lo, int c0, int c1, int parts) {
        for (int part = 0; part < parts; ++part) {
          mbar_wait(&w_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&w_full[st], TILE);
          tma_load_2d(smem + OFF_W + st * TILE, part ? lo : hi, &w_full[st], c0, c1);
          if (++st == NWST) { st = 0; ph ^= 1; }
        }
      };
      auto put_f1 = [&](int c) { put(&tmW1, &tmW1L, 0, c * 128, FPARTS); put(&tmW1, &tmW1L, 64, c * 128, FPARTS); };
      auto put_f2 = [&](int c) { put(&tmW2, &tmW2L, c * 128, 0, FPARTS); put(&tmW2, &tmW2L, c * 128 + 64, 0, FPARTS); };
      for (int t = blockIdx.x; t < m_tiles; t += gridDim.x) {
        put(&tmWo, &tmWoL, 0, 0, PARTS);
        put(&tmWo, &tmWoL, 64, 0, PARTS);
        put_f1(0);
        put_f1(1);
        for (int c = 0; c < NCHUNK; ++c) {
          put_f2(c);
          if (c + 2 < NCHUNK) put_f1(c + 2);
        }
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ------------------------------------------------------------------ tile producer: ctx + residual
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < m_tiles; t += gridDim.x, ++it) {
        mbar_wait(a_empty, (it & 1) ^ 1);
        mbar_arrive_expect_tx(a_full, 2 * TILE);
        tma_load_2d(smem + OFF_CTX, &tmCtx, a_full, 0, t * 128);
        tma_load_2d(smem + OFF_CTX + TILE, &tmCtx, a_full, 64, t * 128);
        mbar_wait(ot_empty, (it & 1) ^ 1);
        mbar_arrive_expect_tx(o_full, 4 * TILE);
#pragma unroll
        for (int j = 0; j < 4; ++j) tma_load_2d(smem + OFF_OT + j * TILE, &tmO, o_full, 32 * j, t * 128);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0, it = 0;
      // one weight unit ([128 x 64] bf16 B tile in the ring) against 4 K-slices of A; `a` gives the A operand
      // of K-slice k of k-block kb: a shared-memory descriptor (SS) or a TMEM address (TS)
      auto unit_ss = [&](uint32_t tmem_d, uint64_t adesc, bool fresh) {
        mbar_wait(&w_full[st], ph);
        tc_fence_after();
        const uint64_t bdesc = umma_desc_k_sw128(smem_u32(smem + OFF_W + st * TILE));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, IDESC, !(fresh && k == 0));
        umma_commit(&w_empty[st]);
        if (++st == NWST) { st = 0; ph ^= 1; }
      };
      int tu = 0;
      auto unit_ts = [&](uint32_t tmem_d, uint32_t a_col, bool fresh) {   // a_col: TMEM column of K-slice 0 (8 columns per slice)
        const bool trf = args.trace != nullptr && blockIdx.x == 0 && it == 1 && tu < 64;
        if (trf) args.trace[128 + 2 * tu] = clock64();
        mbar_wait(&w_full[st], ph);
        if (trf) args.trace[128 + 2 * tu + 1] = clock64();
        ++tu;
        tc_fence_after();
        const uint64_t bdesc = umma_desc_k_sw128(smem_u32(smem + OFF_W + st * TILE));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ts(tmem_d, a_col + 8 * k, bdesc + 2 * k, IDESC, !(fresh && k == 0));
        umma_commit(&w_empty[st]);
        if (++st == NWST) { st = 0; ph ^= 1; }
      };
      // FFN1 chunk: D = LN2(o') [TM_Y2, 64 packed columns] . W1_chunk^T
      auto ffn1 = [&](uint32_t d) {
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int part = 0; part < FPARTS; ++part) unit_ts(d, tmem + TM_Y2 + 32 * kb, kb == 0 && part == 0);
      };
      // FFN2 chunk: Y += relu(H) . W2_chunk^T; the packed chunk sits in columns [0,32) and [64,96) of its accumulator
      auto ffn2 = [&](uint32_t h) {
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int part = 0; part < FPARTS; ++part) unit_ts(tmem + TM_Y, h + 64 * kb, false);
      };
      const bool tr = args.trace != nullptr && blockIdx.x == 0;
      int ti = 0;
#define TRACE_M() do { if (tr && it == 1 && ti < 64) args.trace[ti++] = clock64(); } while (0)
      for (int t = blockIdx.x; t < m_tiles; t += gridDim.x, ++it) {
        TRACE_M();                                // 0: tile start
        tu = 0;
        mbar_wait(accy_empty, (it & 1) ^ 1);      // previous tile's result has been read out of TMEM
        TRACE_M();                                // 1
        mbar_wait(a_full, it & 1);                // ctx tile landed
        TRACE_M();                                // 2
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)            // out-proj (both operands in shared memory)
#pragma unroll
          for (int part = 0; part < PARTS; ++part)
            unit_ss(tmem + TM_Y, umma_desc_k_sw128(smem_u32(smem + OFF_CTX + kb * TILE)), kb == 0 && part == 0);
        umma_commit(accy_full);
        umma_commit(a_empty);
        TRACE_M();                                // 3: out-proj issued
        mbar_wait(y_full, it & 1);                // o' is in TM_Y, LN2(o') in TM_Y2
        TRACE_M();                                // 4: y_full seen
        tc_fence_after();
        ffn1(tmem + TM_H0);
        umma_commit(&acch_full[0]);
        ffn1(tmem + TM_H1);
        umma_commit(&acch_full[1]);
        for (int c = 0; c < NCHUNK; ++c) {
          const int b = c & 1;
          const uint32_t h = tmem + (b ? TM_H1 : TM_H0);
          // slot b is used 4x per tile (chunks b, b+2, b+4, b+6): the parity of its use count is (c >> 1) & 1
          mbar_wait(&hs_full[b], (c >> 1) & 1);   // relu(F1(c)) is packed in its accumulator's columns
          TRACE_M();                              // 5+2c: hs_full(c) seen
          tc_fence_after();
          ffn2(h);
          if (c + 2 < NCHUNK) {
            ffn1(h);                              // F1(c+2) overwrites the chunk after F2(c) (MMAs execute in order)
            umma_commit(&acch_full[b]);
          }
          TRACE_M();                              // 6+2c: round c issued
        }
        umma_commit(accy_done);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue
    // thread = (token row r, column half hf): TMEM lane r, columns [64 hf, 64 hf + 64) of every 128-wide tile
    const int ew = warp - 3;
    const int q = warp & 3;                        // TMEM lane quarter this warp may access
    const int hf = ew >> 2;
    const int r = q * 32 + lane;
    const int cb = hf * 64;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t it = 0;
    const bool tr = args.trace != nullptr && blockIdx.x == 0 && ew == 1 && lane == 0;
    int ti = 64;
#define TRACE_E() do { if (tr && it == 1 && ti < 128) args.trace[ti++] = clock64(); } while (0)
    uint32_t v[64];
    auto ld64 = [&](uint32_t col) {
      tmem_ld32(lane_base + col, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
      tmem_ld32(lane_base + col + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
      tmem_ld_wait();
    };
    for (int t = blockIdx.x; t < m_tiles; t += gridDim.x, ++it) {
      TRACE_E();                                   // 64: tile start
      // ---- E1: o' = out-proj + bo + o -> back into TM_Y; LN2(o') -> TM_Y2 (packed bf16)
      mbar_wait(accy_full, it & 1);
      mbar_wait(o_full, it & 1);
      TRACE_E();                                   // 65: accumulator + residual tile ready
      tc_fence_after();
      ld64(TM_Y + cb);
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float4 b = *reinterpret_cast<const float4*>(s_bo + cb + 4 * j);
        const float4 o4 = *reinterpret_cast<const float4*>(smem + OFF_OT + sw128_f32(r, cb + 4 * j));
        const float x0 = __uint_as_float(v[4 * j + 0]) + b.x + o4.x;
        const float x1 = __uint_as_float(v[4 * j + 1]) + b.y + o4.y;
        const float x2 = __uint_as_float(v[4 * j + 2]) + b.z + o4.z;
        const float x3 = __uint_as_float(v[4 * j + 3]) + b.w + o4.w;
        sum += (x0 + x1) + (x2 + x3);
        v[4 * j + 0] = __float_as_uint(x0); v[4 * j + 1] = __float_as_uint(x1);
        v[4 * j + 2] = __float_as_uint(x2); v[4 * j + 3] = __float_as_uint(x3);
      }
      tmem_st32(lane_base + TM_Y + cb, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
      tmem_st32(lane_base + TM_Y + cb + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
      s_red[hf * 128 + r] = sum;
      epi_bar_sync();
      const float mean = (s_red[r] + s_red[128 + r]) * (1.f / D);
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        const float d = __uint_as_float(v[j]) - mean;
        sq = fmaf(d, d, sq);
      }
      epi_bar_sync();                              // everyone has read the sums before they are overwritten
      s_red[hf * 128 + r] = sq;
      epi_bar_sync();
      const float rstd = rsqrtf((s_red[r] + s_red[128 + r]) * (1.f / D) + LN_EPS);
      {
        uint32_t p[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int c = cb + 2 * j;
          const float y0 = (__uint_as_float(v[2 * j]) - mean) * rstd * s_g2[c] + s_be2[c];
          const float y1 = (__uint_as_float(v[2 * j + 1]) - mean) * rstd * s_g2[c + 1] + s_be2[c + 1];
          p[j] = pack_bf16(y0, y1);
        }
        tmem_st32(lane_base + TM_Y2 + hf * 32, p);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(y_full);
      TRACE_E();                                   // 66: E1 done
      // ---- E2: hidden chunks: relu(F1(c) + b1) packed to bf16, back into the chunk's own TMEM columns
#pragma unroll 1
      for (int c = 0; c < NCHUNK; ++c) {
        const int b = c & 1;
        const uint32_t hcol = (b ? TM_H1 : TM_H0) + cb;
        mbar_wait(&acch_full[b], (c >> 1) & 1);
        TRACE_E();                                 // 67+2c: acch_full(c) seen
        tc_fence_after();
        const float* bias = s_b1 + c * 128 + cb;
        ld64(hcol);
        uint32_t p[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias + 4 * j);
          p[2 * j] = pack_bf16(fmaxf(__uint_as_float(v[4 * j + 0]) + b4.x, 0.f), fmaxf(__uint_as_float(v[4 * j + 1]) + b4.y, 0.f));
          p[2 * j + 1] = pack_bf16(fmaxf(__uint_as_float(v[4 * j + 2]) + b4.z, 0.f), fmaxf(__uint_as_float(v[4 * j + 3]) + b4.w, 0.f));
        }
        tmem_st32(lane_base + hcol, p);            // this thread's own 64 fp32 columns -> 32 packed columns
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&hs_full[b]);
        TRACE_E();                                 // 68+2c: chunk c stored
      }
      // ---- E3: result = TM_Y + b2 -> swizzled fp32 tile in shared memory -> one TMA store
      mbar_wait(accy_done, it & 1);
      TRACE_E();                                   // 83: accy_done seen
      tc_fence_after();
      ld64(TM_Y + cb);
      tc_fence_before();
      mbar_arrive(accy_empty);                     // TM_Y is in registers: the next tile's out-proj may start
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float4 b = *reinterpret_cast<const float4*>(s_b2 + cb + 4 * j);
        *reinterpret_cast<float4*>(smem + OFF_OT + sw128_f32(r, cb + 4 * j)) =
            make_float4(__uint_as_float(v[4 * j + 0]) + b.x, __uint_as_float(v[4 * j + 1]) + b.y,
                        __uint_as_float(v[4 * j + 2]) + b.z, __uint_as_float(v[4 * j + 3]) + b.w);
      }
      fence_proxy_async();
      epi_bar_sync();
      if (ew == 0 && lane == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) tma_store_2d(&tmO, smem + OFF_OT + j * TILE, 32 * j, t * 128);   // rows >= M are clipped
        tma_store_commit();
        tma_store_wait_read<0>();                  // the tile may be overwritten by the next residual load
        mbar_arrive(ot_empty);
      }
      TRACE_E();                                   // 84: tile end
    }
    if (ew == 0 && lane == 0) tma_store_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<post::TMEM_COLS>(tmem);
  }
}

// ------------------------------------------------------------------------------------------------
// k_qkv_tc: y = LayerNorm(o; g1, be1); qkv = y . Win^T + bin  (norm1 + the packed in-projection of
// nn.MultiheadAttention) in one kernel, bf16 out.  Same building blocks as k_post_tc: the fp32 residual
// tile arrives by TMA, dedicated LayerNorm warps (one token row per thread) store the packed bf16 result
// to TMEM (double-buffered, so LN runs one tile ahead of the MMAs), the three 128-wide output tiles are
// TS-form MMAs with one accumulator slot each, and the drain warps empty slot n of tile t while the tensor
// pipe already works on tile t+1.  Output tiles leave through double-buffered swizzled staging + TMA stores.
namespace qkvk {
constexpr int DRAIN_THREADS = 256;               // warps 3..10: accumulator -> bias -> bf16 -> staging -> TMA store
constexpr int LN_THREADS = 128;                  // warps 11..14: LayerNorm, one token row per thread
constexpr int THREADS = 96 + DRAIN_THREADS + LN_THREADS;
constexpr int TILE = 128 * 128;
constexpr int NWST = 5;                          // weight ring depth (16 KB units)
constexpr int OFF_OT = 0;                        // [128 x 128] fp32 residual tile (4 atoms); LN runs one tile ahead of the
                                                 // MMAs, so a single buffer already gives a full tile of prefetch distance
constexpr int OFF_OUT = OFF_OT + 4 * TILE;       // 2 x [128 x 128] bf16 output staging (2 atoms each)
constexpr int OFF_W = OFF_OUT + 4 * TILE;
constexpr int OFF_PAR = OFF_W + NWST * TILE;     // g1[128], be1[128], bin[384]
constexpr int OFF_BAR = OFF_PAR + (2 * D + 3 * D) * 4;
constexpr int NBAR = 2 * NWST + 2 + 4 + 6;
constexpr int SMEM = OFF_BAR + NBAR * 8 + 16;
constexpr int TM_ACC = 0, TM_Y2 = 384;           // 3 x 128 accumulator columns, 2 x 64 packed LN columns
static_assert(SMEM <= 227 * 1024, "shared memory budget");
}  // namespace qkvk

struct QkvArgs {
  const float *g1, *be1, *bin;
  int64_t M;
};

__device__ __forceinline__ uint32_t sw128_bf16(int row, int col) {   // 16-byte chunk of 8 bf16 columns, two 64-column atoms
  return (uint32_t)((col >> 6) * qkvk::TILE + row * 128 + ((((col & 63) >> 3) ^ (row & 7)) << 4));
}
__device__ __forceinline__ void drain_bar_sync() { asm volatile("bar.sync 2, %0;" ::"n"(qkvk::DRAIN_THREADS) : "memory"); }

template <bool SPLIT>
__global__ void __launch_bounds__(qkvk::THREADS, 1)
k_qkv_tc(const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmW,
         const __grid_constant__ CUtensorMap tmWL, const __grid_constant__ CUtensorMap tmQ, const QkvArgs args) {
  using namespace qkvk;
  constexpr int PARTS = SPLIT ? 2 : 1;
  constexpr uint32_t IDESC = umma_idesc(UMMA_BF16, UMMA_BF16, 128, 128);
  extern __shared__ __align__(1024) uint8_t smem[];
  float* s_g1 = reinterpret_cast<float*>(smem + OFF_PAR);
  float *s_be1 = s_g1 + D, *s_bin = s_g1 + 2 * D;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* w_full = bars;
  uint64_t* w_empty = bars + NWST;
  uint64_t* ot_full = bars + 2 * NWST;     // residual tile landed (TMA)
  uint64_t* ot_empty = ot_full + 1;        // the LN warps have read it (128 arrivals)
  uint64_t* y_full = ot_full + 2;          // [2] packed LN output stored in TMEM (128 arrivals)
  uint64_t* y_empty = ot_full + 4;         // [2] the MMAs that read it retired (commit)
  uint64_t* acc_full = ot_full + 6;        // [3] output tile accumulated (commit)
  uint64_t* acc_empty = ot_full + 9;       // [3] drain warps have it in registers (256 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (int)((args.M + 127) / 128);
  for (int i = threadIdx.x; i < D; i += THREADS) { s_g1[i] = args.g1[i]; s_be1[i] = args.be1[i]; }
  for (int i = threadIdx.x; i < 3 * D; i += THREADS) s_bin[i] = args.bin[i];
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmO); prefetch_tmap(&tmW); prefetch_tmap(&tmQ);
    if (SPLIT) prefetch_tmap(&tmWL);
    for (int i = 0; i < NWST; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    mbar_init(ot_full, 1); mbar_init(ot_empty, LN_THREADS);
    for (int i = 0; i < 2; ++i) { mbar_init(&y_full[i], LN_THREADS); mbar_init(&y_empty[i], 1); }
    for (int i = 0; i < 3; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], DRAIN_THREADS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {                               // weight producer: Win tile (n, kb, part) in MMA order
      int st = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < m_tiles; t += gridDim.x)
        for (int n = 0; n < 3; ++n)
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int part = 0; part < PARTS; ++part) {
              mbar_wait(&w_empty[st], ph ^ 1);
              mbar_arrive_expect_tx(&w_full[st], TILE);
              tma_load_2d(smem + OFF_W + st * TILE, part ? &tmWL : &tmW, &w_full[st], kb * 64, n * 128);
              if (++st == NWST) { st = 0; ph ^= 1; }
            }
    }
    __syncwarp();
  } else if (warp == 2) {
    if (lane == 0) {                               // residual tile producer
      uint32_t it = 0;
      for (int t = blockIdx.x; t < m_tiles; t += gridDim.x, ++it) {
        mbar_wait(ot_empty, (it & 1) ^ 1);
        mbar_arrive_expect_tx(ot_full, 4 * TILE);
#pragma unroll
        for (int j = 0; j < 4; ++j) tma_load_2d(smem + OFF_OT + j * TILE, &tmO, ot_full, 32 * j, t * 128);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {                               // MMA issuer
      int st = 0;
      uint32_t ph = 0, it = 0;
      for (int t = blockIdx.x; t < m_tiles; t += gridDim.x, ++it) {
        const int b = it & 1;
        mbar_wait(&y_full[b], (it >> 1) & 1);
        tc_fence_after();
        for (int n = 0; n < 3; ++n) {
          mbar_wait(&acc_empty[n], (it & 1) ^ 1);  // slot n is used once per tile
          tc_fence_after();
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int part = 0; part < PARTS; ++part) {
              mbar_wait(&w_full[st], ph);
              tc_fence_after();
              const uint64_t bdesc = umma_desc_k_sw128(smem_u32(smem + OFF_W + st * TILE));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_ts(tmem + TM_ACC + n * 128, tmem + TM_Y2 + b * 64 + kb * 32 + 8 * k, bdesc + 2 * k, IDESC,
                             !(kb == 0 && part == 0 && k == 0));
              umma_commit(&w_empty[st]);
              if (++st == NWST) { st = 0; ph ^= 1; }
            }
          umma_commit(&acc_full[n]);
        }
        umma_commit(&y_empty[b]);
      }
    }
    __syncwarp();
  } else if (warp >= 11) {
    // ------------------------------------------------------------------ LayerNorm warps: thread == token row
    // The row is re-read from the swizzled shared tile for each of the three passes (sum, centred sum of
    // squares, normalise) instead of being held in 128 registers; no cross-thread exchange is needed.
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t it = 0;
    for (int t = blockIdx.x; t < m_tiles; t += gridDim.x, ++it) {
      const int b = it & 1;
      mbar_wait(ot_full, it & 1);
      float sum = 0.f;
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        const float4 o4 = *reinterpret_cast<const float4*>(smem + OFF_OT + sw128_f32(r, 4 * j));
        sum += (o4.x + o4.y) + (o4.z + o4.w);
      }
      const float mean = sum * (1.f / D);
      float sq = 0.f;
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        const float4 o4 = *reinterpret_cast<const float4*>(smem + OFF_OT + sw128_f32(r, 4 * j));
        const float d0 = o4.x - mean, d1 = o4.y - mean, d2 = o4.z - mean, d3 = o4.w - mean;
        sq = fmaf(d0, d0, sq); sq = fmaf(d1, d1, sq); sq = fmaf(d2, d2, sq); sq = fmaf(d3, d3, sq);
      }
      const float rstd = rsqrtf(sq * (1.f / D) + LN_EPS);
      mbar_wait(&y_empty[b], ((it >> 1) & 1) ^ 1);   // the MMAs of tile it - 2 no longer read this buffer
      tc_fence_after();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t p[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int c = half * 64 + 4 * j;
          const float4 o4 = *reinterpret_cast<const float4*>(smem + OFF_OT + sw128_f32(r, c));
          const float4 g4 = *reinterpret_cast<const float4*>(s_g1 + c);
          const float4 b4 = *reinterpret_cast<const float4*>(s_be1 + c);
          p[2 * j] = pack_bf16((o4.x - mean) * rstd * g4.x + b4.x, (o4.y - mean) * rstd * g4.y + b4.y);
          p[2 * j + 1] = pack_bf16((o4.z - mean) * rstd * g4.z + b4.z, (o4.w - mean) * rstd * g4.w + b4.w);
        }
        tmem_st32(lane_base + TM_Y2 + b * 64 + half * 32, p);
      }
      mbar_arrive(ot_empty);                        // tile buffer may be refilled
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&y_full[b]);
    }
  } else {
    // ------------------------------------------------------------------ drain warps: (row r, column half hf)
    const int ew = warp - 3;
    const int q = warp & 3;
    const int hf = ew >> 2;
    const int r = q * 32 + lane;
    const int cb = hf * 64;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    const bool elected = ew == 0 && lane == 0;
    uint32_t it = 0, nstore = 0;
    for (int t = blockIdx.x; t < m_tiles; t += gridDim.x, ++it) {
#pragma unroll 1
      for (int n = 0; n < 3; ++n, ++nstore) {
        uint8_t* out = smem + OFF_OUT + (nstore & 1) * 2 * TILE;
        mbar_wait(&acc_full[n], it & 1);
        tc_fence_after();
        uint32_t v[64];
        tmem_ld32(lane_base + TM_ACC + n * 128 + cb, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tmem_ld32(lane_base + TM_ACC + n * 128 + cb + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&acc_empty[n]);
        if (elected) tma_store_wait_read<1>();     // the store issued two output tiles ago has read this staging buffer
        drain_bar_sync();
        const float* bias = s_bin + n * 128 + cb;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b0 = *reinterpret_cast<const float4*>(bias + 8 * j);
          const float4 b1 = *reinterpret_cast<const float4*>(bias + 8 * j + 4);
          *reinterpret_cast<uint4*>(out + sw128_bf16(r, cb + 8 * j)) =
              make_uint4(pack_bf16(__uint_as_float(v[8 * j + 0]) + b0.x, __uint_as_float(v[8 * j + 1]) + b0.y),
                         pack_bf16(__uint_as_float(v[8 * j + 2]) + b0.z, __uint_as_float(v[8 * j + 3]) + b0.w),
                         pack_bf16(__uint_as_float(v[8 * j + 4]) + b1.x, __uint_as_float(v[8 * j + 5]) + b1.y),
                         pack_bf16(__uint_as_float(v[8 * j + 6]) + b1.z, __uint_as_float(v[8 * j + 7]) + b1.w));
        }
        fence_proxy_async();
        drain_bar_sync();
        if (elected) {
          tma_store_2d(&tmQ, out, n * 128, t * 128);
          tma_store_2d(&tmQ, out + TILE, n * 128 + 64, t * 128);
          tma_store_commit();
        }
      }
    }
    if (elected) tma_store_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

int launch_qkv_tc(ResepHandle* h, const LayerDev& lw, const float* o, bf16* qkv, int64_t rows, cudaStream_t st) {
  if (rows <= 0) return RESEP_OK;
  ProfScope prof_scope(h, "k_qkv_tc", st);
  const bool split = h->w16_mode >= 1;
  CUtensorMap tmO, tmW, tmWL, tmQ;
  int rc;
  if ((rc = make_tmap<float>(h, &tmO, o, rows, D, 128))) return rc;
  if ((rc = make_tmap<bf16>(h, &tmW, lw.in_w_bf, 3 * D, D, 128))) return rc;
  if ((rc = make_tmap<bf16>(h, &tmWL, lw.in_w_bl, 3 * D, D, 128))) return rc;
  if ((rc = make_tmap<bf16>(h, &tmQ, qkv, rows, 3 * D, 128))) return rc;
  QkvArgs a{lw.norm1_w, lw.norm1_b, lw.in_b_hi, rows};
  const int tiles = (int)((rows + 127) / 128);
  const int grid = tiles < h->sm_count ? tiles : h->sm_count;
  if (split) {
    RESEP_CUDA(h, cudaFuncSetAttribute(k_qkv_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, qkvk::SMEM));
    k_qkv_tc<true><<<grid, qkvk::THREADS, qkvk::SMEM, st>>>(tmO, tmW, tmWL, tmQ, a);
  } else {
    RESEP_CUDA(h, cudaFuncSetAttribute(k_qkv_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, qkvk::SMEM));
    k_qkv_tc<false><<<grid, qkvk::THREADS, qkvk::SMEM, st>>>(tmO, tmW, tmWL, tmQ, a);
  }
  RESEP_LAUNCH_CHECK(h, "k_qkv_tc");
  return RESEP_OK;
}

long long* g_post_trace = nullptr;

int launch_post_tc(ResepHandle* h, const LayerDev& lw, const bf16* ctx, float* o, int64_t rows, cudaStream_t st) {
  if (rows <= 0) return RESEP_OK;
  ProfScope prof_scope(h, "k_post_tc", st);
  const bool split = h->w16_mode >= 1, split_ffn = h->w16_mode == 1;
  CUtensorMap tmCtx, tmO, tmWo, tmWoL, tmW1, tmW1L, tmW2, tmW2L;
  int rc;
  if ((rc = make_tmap<bf16>(h, &tmCtx, ctx, rows, D, 128))) return rc;
  if ((rc = make_tmap<float>(h, &tmO, o, rows, D, 128))) return rc;
  if ((rc = make_tmap<bf16>(h, &tmWo, lw.out_w_bf, D, D, 128))) return rc;
  if ((rc = make_tmap<bf16>(h, &tmWoL, lw.out_w_bl, D, D, 128))) return rc;
  if ((rc = make_tmap<bf16>(h, &tmW1, lw.f1_w_bf, FFN, D, 128))) return rc;
  if ((rc = make_tmap<bf16>(h, &tmW1L, lw.f1_w_bl, FFN, D, 128))) return rc;
  if ((rc = make_tmap<bf16>(h, &tmW2, lw.f2_w_bf, D, FFN, 128))) return rc;
  if ((rc = make_tmap<bf16>(h, &tmW2L, lw.f2_w_bl, D, FFN, 128))) return rc;
  static long long* trace_buf = nullptr;
  if (getenv("RESEP_TRACE") && !trace_buf) { cudaMalloc(&trace_buf, 256 * 8); cudaMemset(trace_buf, 0, 256 * 8); }
  PostArgs a{lw.out_b, lw.norm2_w, lw.norm2_b, lw.f1_b, lw.f2_b, rows, trace_buf};
  if (trace_buf) g_post_trace = trace_buf;
  const int tiles = (int)((rows + 127) / 128);
  const int grid = tiles < h->sm_count ? tiles : h->sm_count;
  auto kern = split_ffn ? k_post_tc<true, true> : split ? k_post_tc<true, false> : k_post_tc<false, false>;
  RESEP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, post::SMEM));
  kern<<<grid, post::THREADS, post::SMEM, st>>>(tmCtx, tmO, tmWo, tmWoL, tmW1, tmW1L, tmW2, tmW2L, a);
  RESEP_LAUNCH_CHECK(h, "k_post_tc");
  return RESEP_OK;
}

}  // namespace resep
