// k_maskdec_tc: the separator's tail in one kernel (bf16 mode) --
//   mask  = ReLU(PReLU(out) . Wfc^T + bfc)                 speechbrain dual_path.py Dual_Path_Model.forward: output_fc
//                                                           (PReLU -> Conv1d 128 -> 2 x 128, k = 1) and the ReLU mask
//   h     = mask[spk] * mix_w                               speechbrain inference/separation.py separate_batch:
//                                                           sep_h = mix_w * est_mask
//   est   = ConvTranspose1d(h; 128 -> 1, k = 16, stride 8)  speechbrain lobes/models/dual_path.py Decoder
// The unfused path wrote the fp32 mask (1,024 B per token row) and read it back in the decoder: 132 MB of the
// 182 MB those two kernels moved at config 2.  Here the mask never leaves the SM: the output_fc accumulators
// ([128 rows x 256] fp32, double-buffered in TMEM) are read back by four epilogue warps (thread = token row, both
// speakers), multiplied with the encoder features (TMA boxes) and contracted against the 128 x 16 decoder filters
// with packed f32x2 FMAs; the overlap-add (each output slot of 8 samples = lower taps of its frame + upper taps of
// the previous frame) is a warp shuffle, and rows leave through a transposed staging tile as coalesced float2
// (sample, speaker) stores.  Tiles advance by 127 rows: row 0 of a tile only supplies its upper taps.
//   warp 0: TMA producer for the A operand (PReLU output, bf16) and, once, the resident output_fc weights (hi + lo)
//   warp 1: MMA issuer (M=128, N=256, K=16; 8 or 16 per tile), TMEM owner
//   warp 2: TMA producer for the encoder-feature boxes
//   warps 3-6: epilogue
#include <cstdlib>
#include <cstring>

#include "tc_common.cuh"

namespace resep {

using namespace ptx;

namespace md {
constexpr int THREADS = 224;
constexpr int ROWS_OUT = 127;                    // output slots per tile (tile = 128 rows, first row is the halo)
constexpr int WHALF = 256 * 128;                 // [256 rows x 128 B]: one K half of Wfc (hi or lo)
constexpr int OFF_W = 0;                         // hi k0, hi k1, lo k0, lo k1
constexpr int NA = 2, ABOX = 128 * 128;          // A ring: [128 rows x 64 bf16] K halves (one tile)
constexpr int OFF_A = OFF_W + 4 * WHALF;
constexpr int NX = 3, XBOX = 128 * 128;          // feature ring: [128 rows x 32 fp32]
constexpr int OFF_X = OFF_A + NA * ABOX;
constexpr int OFF_DW = OFF_X + NX * XBOX;        // decoder filters fp32 [128][16]
constexpr int OFF_B = OFF_DW + D * KSZ * 4;      // output_fc bias fp32 [256] (channel = 2 * filter + speaker)
constexpr int STG_P = 136;                       // staging pitch (floats): [spk][tap][row], 2-way conflicts at most
constexpr int OFF_STG = OFF_B + NSPK * D * 4;
constexpr int OFF_EDGE = OFF_STG + NSPK * 8 * STG_P * 4;   // upper taps of lane 31 of each lane quarter: [4][16]
constexpr int OFF_ROW = OFF_EDGE + 4 * 16 * 4;   // per row: (sample offset of the slot in est, valid samples 0..8)
constexpr int OFF_BAR = OFF_ROW + 128 * 8;
constexpr int NBAR = 1 + 2 * NA + 2 * NX + 4;
constexpr int SMEM = OFF_BAR + NBAR * 8 + 8;
static_assert(SMEM <= 227 * 1024, "shared memory budget");
static_assert(STRIDE == 8 && KSZ == 16 && NSPK == 2 && D == 128, "decoder geometry");
}  // namespace md

struct MaskDecArgs {
  const float* fc_b;
  const float* dec_w;
  const int64_t* item_off;
  const int64_t* item_len;
  const int* item_row0;
  float* est;
  int64_t M;
  int B;
  int n_tiles;
};

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <bool SPLIT, bool F16>
__global__ void __launch_bounds__(md::THREADS, 1)
k_maskdec_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
             const __grid_constant__ CUtensorMap tmWL, const __grid_constant__ CUtensorMap tmX, const MaskDecArgs a) {
  using namespace md;
  constexpr uint32_t IDESC = umma_idesc(F16 ? UMMA_F16 : UMMA_BF16, F16 ? UMMA_F16 : UMMA_BF16, 128, 256);
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* w_full = bars;
  uint64_t* a_full = bars + 1;
  uint64_t* a_empty = a_full + NA;
  uint64_t* x_full = a_empty + NA;
  uint64_t* x_empty = x_full + NX;
  uint64_t* acc_full = x_empty + NX;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBAR);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  pdl_trigger();
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmW); prefetch_tmap(&tmX);
    if (SPLIT) prefetch_tmap(&tmWL);
    mbar_init(w_full, 1);
    for (int i = 0; i < NA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < NX; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 4); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 128); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  if (warp >= 3) {   // static parameters (not produced by the previous kernel): all loads in flight before the stores
    const int t = threadIdx.x - 96;
    float4* dw4 = reinterpret_cast<float4*>(smem + OFF_DW);
    float4* bs4 = reinterpret_cast<float4*>(smem + OFF_B);
    const float4* gdw = reinterpret_cast<const float4*>(a.dec_w);
    const float4 w0 = __ldg(gdw + t), w1 = __ldg(gdw + t + 128), w2 = __ldg(gdw + t + 256), w3 = __ldg(gdw + t + 384);
    const float4 b4 = t < NSPK * D / 4 ? __ldg(reinterpret_cast<const float4*>(a.fc_b) + t) : make_float4(0.f, 0.f, 0.f, 0.f);
    dw4[t] = w0; dw4[t + 128] = w1; dw4[t + 256] = w2; dw4[t + 384] = w3;
    if (t < NSPK * D / 4) bs4[t] = b4;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------- weights (once) + A operand
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, (SPLIT ? 4 : 2) * WHALF);
      tma_load_2d(smem + OFF_W, &tmW, w_full, 0, 0);
      tma_load_2d(smem + OFF_W + WHALF, &tmW, w_full, 64, 0);
      if (SPLIT) {
        tma_load_2d(smem + OFF_W + 2 * WHALF, &tmWL, w_full, 0, 0);
        tma_load_2d(smem + OFF_W + 3 * WHALF, &tmWL, w_full, 64, 0);
      }
      pdl_wait();                                    // the PReLU output comes from the block epilogue before us
      int slot = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
        const int row0 = t * ROWS_OUT;
#pragma unroll
        for (int kh = 0; kh < 2; ++kh) {
          mbar_wait(&a_empty[slot], phase ^ 1);
          mbar_arrive_expect_tx(&a_full[slot], ABOX);
          tma_load_2d(smem + OFF_A + slot * ABOX, &tmA, &a_full[slot], kh * 64, row0);
          if (++slot == NA) { slot = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      mbar_wait(w_full, 0);
      int slot = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
        mbar_wait(&acc_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 256;
#pragma unroll
        for (int kh = 0; kh < 2; ++kh) {
          mbar_wait(&a_full[slot], phase);
          tc_fence_after();
          const uint64_t adesc = umma_desc_k_sw128(smem_u32(smem + OFF_A + slot * ABOX));
          const uint64_t hdesc = umma_desc_k_sw128(smem_u32(smem + OFF_W + kh * WHALF));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, adesc + 2 * k, hdesc + 2 * k, IDESC, (kh | k) != 0);
          if (SPLIT) {
            const uint64_t ldesc = umma_desc_k_sw128(smem_u32(smem + OFF_W + (2 + kh) * WHALF));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, adesc + 2 * k, ldesc + 2 * k, IDESC, true);
          }
          umma_commit(&a_empty[slot]);
          if (++slot == NA) { slot = 0; phase ^= 1; }
        }
        umma_commit(&acc_full[as]);
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ------------------------------------------------------------- encoder-feature boxes
    if (lane == 0) {
      pdl_wait();
      int xs = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
        const int row0 = t * ROWS_OUT;
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          mbar_wait(&x_empty[xs], phase ^ 1);
          mbar_arrive_expect_tx(&x_full[xs], XBOX);
          tma_load_2d(smem + OFF_X + xs * XBOX, &tmX, &x_full[xs], cc * 32, row0);
          if (++xs == NX) { xs = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------- epilogue: thread = token row, both speakers
    const int q = warp & 3;                            // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;
    const int tid = threadIdx.x - 96;                  // 0..127 (store role; not the row)
    const float* dw = reinterpret_cast<const float*>(smem + OFF_DW);
    const float* bs = reinterpret_cast<const float*>(smem + OFF_B);
    float* stg = reinterpret_cast<float*>(smem + OFF_STG);
    float* edge = reinterpret_cast<float*>(smem + OFF_EDGE);
    int2* rowinfo = reinterpret_cast<int2*>(smem + OFF_ROW);
    int as = 0, xs = 0;
    uint32_t aphase = 0, xphase = 0;
    for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
      float2 acc[NSPK][8];
#pragma unroll
      for (int s = 0; s < NSPK; ++s)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[s][k] = make_float2(0.f, 0.f);
      mbar_wait(&acc_full[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * 256 + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        // output_fc channel of (filter n, speaker s) is 2 n + s (upstream reshapes [.., 256] to [.., N, spk]): the 32
        // filters of this chunk are accumulator columns [64 cc, 64 cc + 64)
        uint32_t v0[32], v1[32];
        tmem_ld32(taddr + cc * 64, v0);
        tmem_ld32(taddr + cc * 64 + 32, v1);
        mbar_wait(&x_full[xs], xphase);
        tmem_ld_wait();
        const uint8_t* xrow = smem + OFF_X + xs * XBOX + r * 128;
        const float* dwc = dw + cc * 32 * KSZ;
        const float* bc = bs + cc * 64;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 x4 = *reinterpret_cast<const float4*>(xrow + ((j ^ (r & 7)) << 4));
          const float4 ba = *reinterpret_cast<const float4*>(bc + 8 * j);
          const float4 bb = *reinterpret_cast<const float4*>(bc + 8 * j + 4);
          const float xs4[4] = {x4.x, x4.y, x4.z, x4.w};
          const float b8[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int col = (8 * j + 2 * i) & 31;
            const uint32_t d0 = j < 4 ? v0[col] : v1[col], d1 = j < 4 ? v0[col + 1] : v1[col + 1];
            const float m0 = fmaxf(__uint_as_float(d0) + b8[2 * i], 0.f) * xs4[i];
            const float m1 = fmaxf(__uint_as_float(d1) + b8[2 * i + 1], 0.f) * xs4[i];
            const float2 mm0 = make_float2(m0, m0), mm1 = make_float2(m1, m1);
            const float4* wp = reinterpret_cast<const float4*>(dwc + (4 * j + i) * KSZ);
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              const float4 w4 = wp[w];
              const float2 wa = make_float2(w4.x, w4.y), wb = make_float2(w4.z, w4.w);
              acc[0][2 * w] = ffma2(mm0, wa, acc[0][2 * w]);
              acc[0][2 * w + 1] = ffma2(mm0, wb, acc[0][2 * w + 1]);
              acc[1][2 * w] = ffma2(mm1, wa, acc[1][2 * w]);
              acc[1][2 * w + 1] = ffma2(mm1, wb, acc[1][2 * w + 1]);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&x_empty[xs]);
        if (++xs == NX) { xs = 0; xphase ^= 1; }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }

      // ---- overlap-add: slot of this row = its taps 0..7 + taps 8..15 of the previous row
      float up[NSPK][8];
#pragma unroll
      for (int s = 0; s < NSPK; ++s)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          up[s][2 * k] = __shfl_up_sync(0xffffffffu, acc[s][4 + k].x, 1);
          up[s][2 * k + 1] = __shfl_up_sync(0xffffffffu, acc[s][4 + k].y, 1);
        }
      if (lane == 31) {
#pragma unroll
        for (int s = 0; s < NSPK; ++s)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            edge[q * 16 + s * 8 + 2 * k] = acc[s][4 + k].x;
            edge[q * 16 + s * 8 + 2 * k + 1] = acc[s][4 + k].y;
          }
      }
      // which item / frame this row is
      const int64_t g = (int64_t)t * ROWS_OUT + r;
      int item = 0;
      {
        int lo = 0, hi = a.B - 1;
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if ((int64_t)__ldg(a.item_row0 + mid) <= g) lo = mid; else hi = mid - 1;
        }
        item = lo;
      }
      const int l = (int)(g - __ldg(a.item_row0 + item));
      const int64_t T = __ldg(a.item_len + item);
      int nv = (int)(T - 8 * (int64_t)l < 8 ? T - 8 * (int64_t)l : 8);
      if (nv < 0 || g >= a.M || (r == 0 && t != 0)) nv = 0;     // row 0 is the halo (tile 0: the very first frame)
      epi_bar_sync();                                  // edge[] written; the previous tile's stores have read stg / rowinfo
      if (lane == 0 && q > 0) {
#pragma unroll
        for (int s = 0; s < NSPK; ++s)
#pragma unroll
          for (int k = 0; k < 8; ++k) up[s][k] = edge[(q - 1) * 16 + s * 8 + k];
      }
      const bool first = l == 0 || r == 0;             // no previous frame in this item (or none in this tile: halo row)
#pragma unroll
      for (int s = 0; s < NSPK; ++s)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          stg[(s * 8 + 2 * k) * STG_P + r] = acc[s][k].x + (first ? 0.f : up[s][2 * k]);
          stg[(s * 8 + 2 * k + 1) * STG_P + r] = acc[s][k].y + (first ? 0.f : up[s][2 * k + 1]);
        }
      rowinfo[r] = make_int2((int)(__ldg(a.item_off + item) + 8 * (int64_t)l), nv);
      epi_bar_sync();
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int f = tid + 128 * jj;
        const int rr = f >> 3, k = f & 7;
        const int2 ri = rowinfo[rr];
        if (k < ri.y)
          *reinterpret_cast<float2*>(a.est + 2 * ((int64_t)ri.x + k)) = make_float2(stg[k * STG_P + rr], stg[(8 + k) * STG_P + rr]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

int launch_maskdec(ResepHandle* h, const bf16* prelu, const float* x0, const Plan& p, float* est, cudaStream_t st) {
  if (p.M <= 0) return RESEP_OK;
  CUtensorMap tmA, tmW, tmWL, tmX;
  int rc;
  const bool split = h->w16_mode != 0;
  if ((rc = make_tmap<bf16>(h, &tmA, prelu, p.M, D, 128))) return rc;
  const bf16* w_hi = h->fmt16 ? h->w.fc_w_h[0] : h->w.fc_w_bf;
  const bf16* w_lo = h->fmt16 ? h->w.fc_w_h[1] : h->w.fc_w_bl;
  if ((rc = make_tmap<bf16>(h, &tmW, w_hi, NSPK * D, D, 256))) return rc;
  if ((rc = make_tmap<bf16>(h, &tmWL, split ? w_lo : w_hi, NSPK * D, D, 256))) return rc;
  if ((rc = make_tmap<float>(h, &tmX, x0, p.M, D, 128))) return rc;
  MaskDecArgs a;
  a.fc_b = h->w.fc_b; a.dec_w = h->w.dec_w;
  a.item_off = p.d_item_off; a.item_len = p.d_item_len; a.item_row0 = p.d_item_row0;
  a.est = est; a.M = p.M; a.B = p.B;
  a.n_tiles = (int)((p.M - 1 + md::ROWS_OUT - 1) / md::ROWS_OUT);
  if (a.n_tiles < 1) a.n_tiles = 1;
  const int grid = a.n_tiles < h->sm_count ? a.n_tiles : h->sm_count;
  ProfScope prof_scope(h, "k_maskdec_tc", st);
  auto kern = h->fmt16 ? (split ? k_maskdec_tc<true, true> : k_maskdec_tc<false, true>) : (split ? k_maskdec_tc<true, false> : k_maskdec_tc<false, false>);
  RESEP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, md::SMEM));
  RESEP_CUDA(h, launch_pdl(kern, dim3(grid), dim3(md::THREADS), md::SMEM, st, tmA, tmW, tmWL, tmX, a));
  RESEP_LAUNCH_CHECK(h, "k_maskdec_tc");
  return RESEP_OK;
}

}  // namespace resep
