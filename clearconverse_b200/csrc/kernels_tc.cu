// Tensor-core kernels of the RE-SepFormer path (sm_100a only).
//
//  * k_gemm_tc     persistent, warp-specialised GEMM  out = epi(A[M,K] . W[N,K]^T + bias):
//                  TMA (cp.async.bulk.tensor, 128B swizzle) -> 4-stage smem ring -> tcgen05.mma
//                  (cta_group::1, M=128 x N=128 per instruction, kind::f16 for bf16 operands or
//                  kind::tf32) -> double-buffered fp32 accumulators in TMEM -> tcgen05.ld epilogue
//                  (bias, ReLU, residual add, bf16 / tf32 rounding) overlapped with the next tile's MMAs.
//  * k_attention_bf16  per (sequence tile, head) softmax(q k^T / 4) v with bf16 mma.sync and fp32
//                  online softmax; head_dim is 16, so the kernel is exp/softmax-bound, not MMA-bound.
//
// Upstream arithmetic restated (speechbrain Transformer.py TransformerEncoderLayer, pre-norm):
//   y = LN1(x); x += OutProj(MHA(y)); y = LN2(x); x += W2 relu(W1 y + b1) + b2
#include <cuda.h>

#include <cstdlib>
#include <type_traits>

#include "tc_common.cuh"

namespace resep {

using namespace ptx;

PFN_encodeTiled g_encode = nullptr;

int tc_init(ResepHandle* h) {
  if (h->tc_ready) return RESEP_OK;
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
      return set_err(h, RESEP_ECUDA, "cuTensorMapEncodeTiled entry point not available");
    g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
  }
  h->tc_ready = true;
  return RESEP_OK;
}

void tc_destroy(ResepHandle* h) { h->tc_ready = false; }

// ------------------------------------------------------------------------------------------------
// GEMM
constexpr int BM = 128, BN = 128;
constexpr int STAGES = 4;
constexpr int STAGE_A_BYTES = BM * 128;               // 128 rows x 128 B of K
constexpr int STAGE_B_BYTES = BN * 128;
constexpr int stage_bytes(bool split_w) { return STAGE_A_BYTES + (split_w ? 2 : 1) * STAGE_B_BYTES; }
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = ACC_STAGES * BN;            // 256 fp32 columns
constexpr int GEMM_THREADS = 192;                     // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue
constexpr int gemm_smem(bool split_w) { return STAGES * stage_bytes(split_w) + 1024 /*alignment slack*/ + 256 /*barriers*/; }

enum { EPI_STORE_F32 = 0, EPI_STORE_BF16 = 1, EPI_RESID_F32 = 2, EPI_STORE_TF32 = 3 };

// SPLITW (tf32 only): the weight is given as two tf32 matrices W = W_hi + W_lo and every K-slice issues two
// MMAs into the same accumulator; this removes the weight-rounding error, which dominates plain TF32
// (measured: max-abs 8e-4 -> 2e-4 on the forward pass), at the cost of one more B tile per stage.
// (Mixing a bf16 A with an fp16 B in one kind::f16 MMA was tried for the same purpose: it raises an
// illegal-instruction fault on sm_100a, so both operands always share one format.)
template <typename TIn, int EPI, bool SPLITW>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
          const __grid_constant__ CUtensorMap tmWlo, const float* __restrict__ bias, void* out_, int64_t M, int N, int K,
          int relu, int f16) {
  constexpr int STAGE_BYTES = stage_bytes(SPLITW);
  constexpr int BK = 128 / sizeof(TIn);               // elements per 128-byte swizzle row: 64 bf16 / 32 tf32
  constexpr int UK = 32 / sizeof(TIn);                // K per tcgen05.mma: 16 bf16 / 8 tf32
  // 16-bit operands: bf16 or (f16 != 0) IEEE fp16 -- same instruction, another format field
  const uint32_t FMT = sizeof(TIn) == 2 ? (f16 ? UMMA_F16 : UMMA_BF16) : UMMA_TF32;
  const uint32_t IDESC = umma_idesc(FMT, FMT, BM, BN);

  extern __shared__ __align__(1024) uint8_t smem[];   // 128B-swizzled tiles need 1024-byte alignment
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;                          // [STAGES]
  uint64_t* empty_bar = bars + STAGES;                // [STAGES]
  uint64_t* acc_full = bars + 2 * STAGES;             // [ACC_STAGES]
  uint64_t* acc_empty = bars + 2 * STAGES + ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * ACC_STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = N / BN;
  const int m_tiles = (int)((M + BM - 1) / BM);
  const int total_tiles = m_tiles * n_tiles;
  const int kblocks = K / BK;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmW);
    if constexpr (SPLITW) prefetch_tmap(&tmWlo);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < ACC_STAGES; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 128); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer (one thread)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int m0 = (t / n_tiles) * BM, n0 = (t % n_tiles) * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m0);
          tma_load_2d(sa + STAGE_A_BYTES, &tmW, &full_bar[stage], kb * BK, n0);
          if constexpr (SPLITW) tma_load_2d(sa + STAGE_A_BYTES + STAGE_B_BYTES, &tmWlo, &full_bar[stage], kb * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer (one thread)
    if (lane == 0) {
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        mbar_wait(&acc_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t adesc = umma_desc_k_sw128(sa);
          const uint64_t bdesc = umma_desc_k_sw128(sa + STAGE_A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            // advancing K inside the 128B swizzle atom = +32 B on the start address (>>4 -> +2)
            if constexpr (sizeof(TIn) == 2)
              umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, (kb | k) != 0);
            else
              umma_tf32(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, (kb | k) != 0);
          }
          if constexpr (SPLITW) {
            const uint64_t ldesc = umma_desc_k_sw128(sa + STAGE_A_BYTES + STAGE_B_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UK; ++k) {
              if constexpr (sizeof(TIn) == 2) umma_bf16(d_tmem, adesc + 2 * k, ldesc + 2 * k, IDESC, true);
              else umma_tf32(d_tmem, adesc + 2 * k, ldesc + 2 * k, IDESC, true);
            }
          }
          umma_commit(&empty_bar[stage]);             // frees the smem slot when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&acc_full[as]);                   // accumulator ready for the epilogue
        if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------- epilogue (4 warps, one row per thread)
    const int q = warp & 3;                            // TMEM lane quarter this warp may access
    int as = 0;
    uint32_t aphase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int m0 = (t / n_tiles) * BM, n0 = (t % n_tiles) * BN;
      const int64_t row = (int64_t)m0 + q * 32 + lane;
      mbar_wait(&acc_full[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        tmem_ld_wait();
        if (row < M) {
          const float* bp = bias + n0 + c0;
          if constexpr (EPI == EPI_STORE_BF16) {
            uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(out_) + row * N + n0 + c0);
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float f[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                f[i] = __uint_as_float(v[j + i]) + __ldg(bp + j + i);
                if (relu) f[i] = fmaxf(f[i], 0.f);
              }
              op[j / 8] = f16 ? make_uint4(pack16<true>(f[0], f[1]), pack16<true>(f[2], f[3]), pack16<true>(f[4], f[5]), pack16<true>(f[6], f[7]))
                              : make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
            }
          } else {
            float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(out_) + row * N + n0 + c0);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float f[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                f[i] = __uint_as_float(v[j + i]) + __ldg(bp + j + i);
                if (relu) f[i] = fmaxf(f[i], 0.f);
                if constexpr (EPI == EPI_STORE_TF32) f[i] = round_tf32(f[i]);
              }
              if constexpr (EPI == EPI_RESID_F32) {
                const float4 r = op[j / 4];
                f[0] += r.x; f[1] += r.y; f[2] += r.z; f[3] += r.w;
              }
              op[j / 4] = make_float4(f[0], f[1], f[2], f[3]);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[as]);
      if (++as == ACC_STAGES) { as = 0; aphase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

template <typename TIn, int EPI, bool SPLITW = false>
static int launch_gemm_tc(ResepHandle* h, const TIn* A, const TIn* W, const float* bias, void* out, int64_t M, int N,
                          int K, bool relu, cudaStream_t st, const TIn* Wlo = nullptr) {
  if (M <= 0) return RESEP_OK;
  constexpr int BK = 128 / sizeof(TIn);
  if (N % BN != 0 || K % BK != 0) return set_err(h, RESEP_EINVAL, "gemm_tc: N % 128 or K % (128 B) != 0");
  CUtensorMap tmA, tmW, tmWlo;
  int rc;
  if ((rc = make_tmap<TIn>(h, &tmA, A, M, K, BM))) return rc;
  if ((rc = make_tmap<TIn>(h, &tmW, W, N, K, BN))) return rc;
  if ((rc = make_tmap<TIn>(h, &tmWlo, SPLITW ? Wlo : W, N, K, BN))) return rc;
  constexpr int GEMM_SMEM = gemm_smem(SPLITW);
  auto kern = k_gemm_tc<TIn, EPI, SPLITW>;
  RESEP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
  const int tiles = (int)((M + BM - 1) / BM) * (N / BN);
  const int grid = tiles < h->sm_count ? tiles : h->sm_count;
  static const char* const kEpiName[4] = {"f32", "bf16", "resid", "tf32"};
  static const std::string kname = std::string("k_gemm_tc<") + (sizeof(TIn) == 2 ? "bf16" : "tf32") + "," + kEpiName[EPI] +
                                   (SPLITW ? ",hi+lo>" : ">");
  ProfScope prof_scope(h, kname.c_str(), st);
  kern<<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(tmA, tmW, tmWlo, bias, out, M, N, K, relu ? 1 : 0, sizeof(TIn) == 2 ? h->fmt16 : 0);
  RESEP_LAUNCH_CHECK(h, "k_gemm_tc");
  return RESEP_OK;
}

// bf16-activation GEMM; mode != 0: W as bf16 hi + lo (two MMAs per K-slice), 0: bf16(W) only
template <int EPI>
static int gemm_bf16(ResepHandle* h, int mode, const bf16* A, const bf16* W, const bf16* Wlo, const float* bias, void* out,
                     int64_t M, int N, int K, bool relu, cudaStream_t st) {
  if (mode != 0) return launch_gemm_tc<bf16, EPI, true>(h, A, W, bias, out, M, N, K, relu, st, Wlo);
  return launch_gemm_tc<bf16, EPI, false>(h, A, W, bias, out, M, N, K, relu, st);
}

// ------------------------------------------------------------------------------------------------
// Attention, bf16 mma.sync m16n8k16.  One CTA = 160 query rows (5 warps x 2 m16 tiles) of one
// sequence for one head; keys/values are streamed through shared memory in tiles of 160 with an
// online (running max / sum) softmax, so a 150-row chunk is a single pass and the memory
// transformer's longer sequences loop.  K is staged row-major [key][16] (row stride 24 halves,
// conflict-free B-fragment reads), V transposed [16][key] so that P.V B-fragments are 32-bit loads.
constexpr int AKT = 160;        // keys per tile == queries per CTA
constexpr int KS_STRIDE = 24;   // halves
constexpr int VT_STRIDE = AKT + 8;
template <bool F16>
__global__ void __launch_bounds__(160) k_attention_bf16(const bf16* __restrict__ qkv, bf16* __restrict__ ctx, int seq_len,
                                                        const int* __restrict__ seq_off, const int* __restrict__ tile_seq,
                                                        const int* __restrict__ tile_q0) {
  __shared__ __align__(16) bf16 Ks[AKT * KS_STRIDE];
  __shared__ __align__(16) bf16 Vt[DH * VT_STRIDE];
  int q0, off, len;
  if (tile_seq != nullptr) {
    const int seq = tile_seq[blockIdx.x];
    q0 = tile_q0[blockIdx.x];
    off = seq_off[seq];
    len = seq_off[seq + 1] - off;
  } else {
    const int tps = (seq_len + AKT - 1) / AKT;
    q0 = (blockIdx.x % tps) * AKT;
    off = (blockIdx.x / tps) * seq_len;
    len = seq_len;
  }
  const int head = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const bf16* qbase = qkv + (int64_t)off * (3 * D) + head * QKV_HEAD_STRIDE;

  // Q fragments of this warp's two m16 tiles (rows beyond the sequence read as zero)
  uint32_t qa[2][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    const int r0 = q0 + warp * 32 + mt * 16 + g;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int r = r0 + hh * 8;
      uint32_t lo = 0, hi = 0;
      if (r < len) {
        const uint32_t* p = reinterpret_cast<const uint32_t*>(qbase + (int64_t)r * (3 * D));
        lo = p[t4];
        hi = p[t4 + 4];
      }
      qa[mt][hh] = lo;
      qa[mt][hh + 2] = hi;
    }
  }
  float m_run[2][2], l_run[2][2], o[2][2][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      m_run[mt][hh] = -INFINITY;
      l_run[mt][hh] = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) o[mt][hh][i] = 0.f;   // o[mt][ntile][c]
    }

  for (int kt = 0; kt < len; kt += AKT) {
    __syncthreads();
    {  // stage one key/value row per thread
      const int key = kt + threadIdx.x;
      uint4 k0 = make_uint4(0, 0, 0, 0), k1 = k0, v0 = k0, v1 = k0;
      if (key < len) {
        const uint4* p = reinterpret_cast<const uint4*>(qbase + (int64_t)key * (3 * D) + QKV_K);
        k0 = p[0]; k1 = p[1];
        const uint4* pv = reinterpret_cast<const uint4*>(qbase + (int64_t)key * (3 * D) + QKV_V);
        v0 = pv[0]; v1 = pv[1];
      }
      uint32_t* kd = reinterpret_cast<uint32_t*>(Ks + threadIdx.x * KS_STRIDE);
      kd[0] = k0.x; kd[1] = k0.y; kd[2] = k0.z; kd[3] = k0.w; kd[4] = k1.x; kd[5] = k1.y; kd[6] = k1.z; kd[7] = k1.w;
      const uint32_t vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const __nv_bfloat162 pr = *reinterpret_cast<const __nv_bfloat162*>(&vv[i]);
        Vt[(2 * i) * VT_STRIDE + threadIdx.x] = pr.x;
        Vt[(2 * i + 1) * VT_STRIDE + threadIdx.x] = pr.y;
      }
    }
    __syncthreads();
    const int nvalid = min(AKT, len - kt);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      if (q0 + warp * 32 + mt * 16 >= len) continue;   // warp-uniform: whole m-tile is padding
      float s[AKT / 8][4];
#pragma unroll
      for (int j = 0; j < AKT / 8; ++j) {
        s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
        const uint32_t* kp = reinterpret_cast<const uint32_t*>(Ks + (8 * j + g) * KS_STRIDE);
        mma16_16816<F16>(s[j], qa[mt], kp[t4], kp[t4 + 4]);
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int j = 0; j < AKT / 8; ++j) {
        const int c = 8 * j + 2 * t4;
        s[j][0] = (c < nvalid) ? s[j][0] : -INFINITY;
        s[j][1] = (c + 1 < nvalid) ? s[j][1] : -INFINITY;
        s[j][2] = (c < nvalid) ? s[j][2] : -INFINITY;
        s[j][3] = (c + 1 < nvalid) ? s[j][3] : -INFINITY;
        mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
        mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m_run[mt][0], mx0), mn1 = fmaxf(m_run[mt][1], mx1);
      const float c0 = exp2f(m_run[mt][0] - mn0), c1 = exp2f(m_run[mt][1] - mn1);
      m_run[mt][0] = mn0; m_run[mt][1] = mn1;
      l_run[mt][0] *= c0; l_run[mt][1] *= c1;
#pragma unroll
      for (int nn = 0; nn < 2; ++nn) { o[mt][nn][0] *= c0; o[mt][nn][1] *= c0; o[mt][nn][2] *= c1; o[mt][nn][3] *= c1; }
      float ls0 = 0.f, ls1 = 0.f;
#pragma unroll
      for (int j = 0; j < AKT / 8; ++j) {
        s[j][0] = exp2f(s[j][0] - mn0); s[j][1] = exp2f(s[j][1] - mn0);
        s[j][2] = exp2f(s[j][2] - mn1); s[j][3] = exp2f(s[j][3] - mn1);
        ls0 += s[j][0] + s[j][1];
        ls1 += s[j][2] + s[j][3];
      }
      l_run[mt][0] += ls0; l_run[mt][1] += ls1;
#pragma unroll
      for (int kk = 0; kk < AKT / 16; ++kk) {
        uint32_t pa[4];
        pa[0] = pack16<F16>(s[2 * kk][0], s[2 * kk][1]);
        pa[1] = pack16<F16>(s[2 * kk][2], s[2 * kk][3]);
        pa[2] = pack16<F16>(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pa[3] = pack16<F16>(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int nn = 0; nn < 2; ++nn) {
          const uint32_t* vp = reinterpret_cast<const uint32_t*>(Vt + (8 * nn + g) * VT_STRIDE + 16 * kk);
          mma16_16816<F16>(o[mt][nn], pa, vp[t4], vp[t4 + 4]);
        }
      }
    }
  }
  // normalise and store: C fragment rows g / g+8, columns 2*t4, 2*t4+1 of each 8-wide dh tile
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      float l = l_run[mt][hh];
      l += __shfl_xor_sync(0xffffffffu, l, 1);
      l += __shfl_xor_sync(0xffffffffu, l, 2);
      const int r = q0 + warp * 32 + mt * 16 + hh * 8 + g;
      if (r < len) {
        const float inv = 1.f / l;
        bf16* op = ctx + (int64_t)(off + r) * D + head * DH;
#pragma unroll
        for (int nn = 0; nn < 2; ++nn)
          *reinterpret_cast<uint32_t*>(op + 8 * nn + 2 * t4) =
              pack16<F16>(o[mt][nn][2 * hh] * inv, o[mt][nn][2 * hh + 1] * inv);
      }
    }
  }
}

// Short-sequence attention (len <= 160: every intra chunk and most memory sequences), one CTA per
// (sequence, head), 5 warps x 2 m16 query tiles.  Two passes over the keys with Q.K^T recomputed
// (19 extra HMMA per tile, the tensor pipe is nearly idle here) instead of holding the 16 x 160 score
// tile in registers: ~50 registers per thread, so 6+ CTAs are resident per SM and the MUFU.EX2 /
// LDS / HMMA latencies overlap across warps.  Fragments come from ldmatrix (K as is, V transposed
// by the .trans form), rows padded to 48 B so the 8 x 16 B row reads are bank-conflict free.
// Per score: FMNMX (pass 1); FFMA + MUFU.EX2 + FADD + half a pack (pass 2) -- the MUFU unit (16/clk/SM)
// is the floor.
constexpr int AS_ROW = 24;   // halves per staged K / V row (16 used)
// Template: KMAX keys staged per (sequence, head) CTA, WARPS warps of MT m16 query tiles each (WARPS * MT * 16 query rows
// per CTA; longer sequences take `q_tiles` CTAs).  <160, 5, 2> is the per-chunk kernel; <448, 4, 1> covers the coupled
// memory transformer's single sequence of all chunk summaries (432 rows at config 2) with 7 x 8 CTAs instead of the
// 3 x 8 of the streaming kernel below, which is latency-bound at that size.
__device__ __forceinline__ float sqrt_approx(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// Squared norms for the softmax stabiliser, accumulated in the packed 16-bit operand format (HFMA2: two elements per
// instruction, no conversions).  Only an upper bound is needed downstream; the roundings are covered by the 1.02 margin.
// (An fp16 sum that overflows becomes inf, which selects the exact-row-maximum path.)
template <bool F16>
__device__ __forceinline__ float sqnorm16_h(const uint32_t (&w)[8]) {
  uint32_t acc = 0u;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc = hfma2_sq<F16>(w[i], acc);
  const float2 f = unpack16<F16>(acc);
  return f.x + f.y;
}
template <bool F16>
__device__ __forceinline__ float sqnorm4_h(uint32_t a, uint32_t b) {
  const float2 f = unpack16<F16>(hfma2_sq<F16>(b, hfma2_sq<F16>(a, 0u)));
  return f.x + f.y;
}

template <int KMAX, int WARPS, int MT, bool F16>
__global__ void __launch_bounds__(32 * WARPS, KMAX <= 160 ? 6 : 2)
k_attention_bf16_short(const bf16* __restrict__ qkv, bf16* __restrict__ ctx, int seq_len, const int* __restrict__ seq_off,
                       int q_tiles) {
  constexpr int THREADS = 32 * WARPS, QROWS = WARPS * MT * 16;
  __shared__ __align__(16) bf16 Ks[KMAX * AS_ROW];
  __shared__ __align__(16) bf16 Vs[KMAX * AS_ROW];
  pdl_trigger();
  pdl_wait();                                        // qkv comes from the previous kernel in the stream
  const int seq = blockIdx.x / q_tiles, q0 = (blockIdx.x % q_tiles) * QROWS;
  int off, len;
  if (seq_off != nullptr) {
    off = seq_off[seq];
    len = seq_off[seq + 1] - off;
  } else {
    off = seq * seq_len;
    len = seq_len;
  }
  if (q0 >= len) return;
  const int head = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const bf16* qbase = qkv + (int64_t)off * (3 * D) + head * QKV_HEAD_STRIDE;
  __shared__ float s_kmax[WARPS];
  // Q fragments of this warp's two m16 tiles, requested before the K / V rows so that the CTA pays one global-memory
  // latency, not three (rows past the sequence read as zero)
  uint32_t qf[MT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int r = q0 + warp * (16 * MT) + mt * 16 + g + hh * 8;
      uint32_t lo = 0, hi = 0;
      if (r < len) {
        const uint32_t* p = reinterpret_cast<const uint32_t*>(qbase + (int64_t)r * (3 * D));
        lo = __ldg(p + t4);
        hi = __ldg(p + t4 + 4);
      }
      qf[mt][hh] = lo;
      qf[mt][hh + 2] = hi;
    }
  {  // stage the key / value rows (rows past the sequence up to the next multiple of 16 are zero); max squared key norm
    float kn2max = 0.f;
    const int kp = (len + 15) & ~15;
    for (int key = threadIdx.x; key < kp; key += THREADS) {
      uint4 k0 = make_uint4(0, 0, 0, 0), k1 = k0, v0 = k0, v1 = k0;
      if (key < len) {
        const uint4* p = reinterpret_cast<const uint4*>(qbase + (int64_t)key * (3 * D) + QKV_K);
        k0 = p[0]; k1 = p[1];
        const uint4* pv = reinterpret_cast<const uint4*>(qbase + (int64_t)key * (3 * D) + QKV_V);
        v0 = pv[0]; v1 = pv[1];
      }
      uint4* kd = reinterpret_cast<uint4*>(Ks + key * AS_ROW);
      uint4* vd = reinterpret_cast<uint4*>(Vs + key * AS_ROW);
      kd[0] = k0; kd[1] = k1;
      vd[0] = v0; vd[1] = v1;
      const uint32_t kw[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
      kn2max = fmaxf(kn2max, sqnorm16_h<F16>(kw));     // (same arithmetic as k_attention_bf16_tma: the two kernels must agree bit for bit)
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kn2max = fmaxf(kn2max, __shfl_xor_sync(0xffffffffu, kn2max, o));
    if (lane == 0) s_kmax[warp] = kn2max;
  }
  __syncthreads();
  float kmax2 = s_kmax[0];
#pragma unroll
  for (int i = 1; i < WARPS; ++i) kmax2 = fmaxf(kmax2, s_kmax[i]);
  const int nkk = (len + 15) >> 4;                    // 16-key blocks holding at least one valid key
  // ldmatrix lane addressing: matrix m = lane / 8, row lane % 8
  const int lm = lane >> 3, lr = lane & 7;
  const bf16* k_lane = Ks + ((lm >> 1) * 8 + lr) * AS_ROW + (lm & 1) * 8;   // K: m0/m1 = dh halves of tile A, m2/m3 of tile B
  const bf16* v_lane = Vs + ((lm & 1) * 8 + lr) * AS_ROW + (lm >> 1) * 8;   // V^T: m0/m1 = key halves for dh 0-7, m2/m3 for dh 8-15
  constexpr uint32_t ONES = F16 ? 0x3C003C00u : 0x3F803F80u;              // bf16 (1.0, 1.0): B fragment of an all-ones [16 keys x 8] matrix
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int row0 = q0 + warp * (16 * MT) + mt * 16;
    if (row0 >= len) break;                            // warp-uniform
    const uint32_t (&qa)[4] = qf[mt];
    // Softmax stabiliser.  Any m_i >= max_j s_ij gives the same softmax; by Cauchy-Schwarz s_ij <= |q_i| |k_j| so
    // m_i = |q_i| * max_j |k_j| (computed from norms: no pass over the scores) never overflows.  Scores lie in
    // [-m_i, m_i], so exp((s_ij - m_i) / 4) cannot flush to zero as long as 2 m_i / 4 < 80 nats; a warp whose rows
    // violate that (never seen with LayerNorm'ed inputs) takes the exact row-maximum pass instead.
    float qn0, qn1;
    {
      qn0 = sqnorm4_h<F16>(qa[0], qa[2]);               // row g:     dims 2 t4, 2 t4 + 1, 8 + 2 t4, 9 + 2 t4
      qn1 = sqnorm4_h<F16>(qa[1], qa[3]);               // row g + 8
      qn0 += __shfl_xor_sync(0xffffffffu, qn0, 1); qn0 += __shfl_xor_sync(0xffffffffu, qn0, 2);
      qn1 += __shfl_xor_sync(0xffffffffu, qn1, 1); qn1 += __shfl_xor_sync(0xffffffffu, qn1, 2);
    }
    // upper bounds of the raw scores; margin: the packed-bf16 squared norms may each come out up to 1.6 % low
    float mx0 = sqrt_approx(qn0 * kmax2) * 1.02f, mx1 = sqrt_approx(qn1 * kmax2) * 1.02f;
    float mb = fmaxf(mx0, mx1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, o));
    // Only the last 16-key block can hold keys past the sequence: it is peeled off (MASK = true), so the other
    // blocks carry no compare / select instructions.
    auto max_block = [&](int kk, auto mask_c) {
      constexpr bool MASK = decltype(mask_c)::value;
      uint32_t kf[4];
      ldmatrix_x4(kf, k_lane + kk * 16 * AS_ROW);
      float s0[4], s1[4];
      mma16_16816_z<F16>(s0, qa, kf[0], kf[1]);
      mma16_16816_z<F16>(s1, qa, kf[2], kf[3]);
      if (MASK) {
        const int c = kk * 16 + 2 * t4;
        if (c >= len) { s0[0] = -INFINITY; s0[2] = -INFINITY; }
        if (c + 1 >= len) { s0[1] = -INFINITY; s0[3] = -INFINITY; }
        if (c + 8 >= len) { s1[0] = -INFINITY; s1[2] = -INFINITY; }
        if (c + 9 >= len) { s1[1] = -INFINITY; s1[3] = -INFINITY; }
      }
      mx0 = fmaxf(mx0, fmaxf(fmaxf(s0[0], s0[1]), fmaxf(s1[0], s1[1])));
      mx1 = fmaxf(mx1, fmaxf(fmaxf(s0[2], s0[3]), fmaxf(s1[2], s1[3])));
    };
    if (!(mb * (0.5f / QK_PRESCALE) < 80.f)) {                         // warp-uniform; also taken for NaN / inf inputs
      // ---- exact pass: row maxima of the raw scores
      mx0 = -INFINITY; mx1 = -INFINITY;
#pragma unroll 1
      for (int kk = 0; kk < nkk - 1; ++kk) max_block(kk, std::false_type{});
      max_block(nkk - 1, std::true_type{});
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    }
    const float nb0 = -mx0, nb1 = -mx1;               // k is pre-scaled (QK_PRESCALE): scores are base-2 exponents
    // ---- pass 2: p = exp2(s - max) (k pre-scaled); O += P . V; row sums of the bf16-rounded P from a third MMA against an
    // all-ones B fragment (every column of that accumulator is the row sum, so no add chain and no shuffle)
    float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f}, ol[4] = {0.f, 0.f, 0.f, 0.f};
    auto pv_block = [&](int kk, auto mask_c) {
      constexpr bool MASK = decltype(mask_c)::value;
      uint32_t kf[4], vf[4];
      ldmatrix_x4(kf, k_lane + kk * 16 * AS_ROW);
      ldmatrix_x4_trans(vf, v_lane + kk * 16 * AS_ROW);
      float s0[4], s1[4];
      mma16_16816_c<F16>(s0, qa, kf[0], kf[1], nb0, nb1);   // s - stabiliser, straight out of the tensor core
      mma16_16816_c<F16>(s1, qa, kf[2], kf[3], nb0, nb1);
#pragma unroll
      for (int i = 0; i < 4; ++i) { s0[i] = ex2_approx(s0[i]); s1[i] = ex2_approx(s1[i]); }
      if (MASK) {
        const int c = kk * 16 + 2 * t4;
        if (c >= len) { s0[0] = 0.f; s0[2] = 0.f; }
        if (c + 1 >= len) { s0[1] = 0.f; s0[3] = 0.f; }
        if (c + 8 >= len) { s1[0] = 0.f; s1[2] = 0.f; }
        if (c + 9 >= len) { s1[1] = 0.f; s1[3] = 0.f; }
      }
      uint32_t pa[4];
      pa[0] = pack16<F16>(s0[0], s0[1]);
      pa[1] = pack16<F16>(s0[2], s0[3]);
      pa[2] = pack16<F16>(s1[0], s1[1]);
      pa[3] = pack16<F16>(s1[2], s1[3]);
      mma16_16816<F16>(o0, pa, vf[0], vf[1]);
      mma16_16816<F16>(o1, pa, vf[2], vf[3]);
      mma16_16816<F16>(ol, pa, ONES, ONES);
    };
#pragma unroll 3
    for (int kk = 0; kk < nkk - 1; ++kk) pv_block(kk, std::false_type{});
    pv_block(nkk - 1, std::true_type{});
    const float i0 = rcp_approx(ol[0]), i1 = rcp_approx(ol[2]);
    const int r0 = row0 + g, r1 = row0 + g + 8;
    if (r0 < len) {
      bf16* op = ctx + (int64_t)(off + r0) * D + head * DH + 2 * t4;
      *reinterpret_cast<uint32_t*>(op) = pack16<F16>(o0[0] * i0, o0[1] * i0);
      *reinterpret_cast<uint32_t*>(op + 8) = pack16<F16>(o1[0] * i0, o1[1] * i0);
    }
    if (r1 < len) {
      bf16* op = ctx + (int64_t)(off + r1) * D + head * DH + 2 * t4;
      *reinterpret_cast<uint32_t*>(op) = pack16<F16>(o0[2] * i1, o0[3] * i1);
      *reinterpret_cast<uint32_t*>(op + 8) = pack16<F16>(o1[2] * i1, o1[3] * i1);
    }
  }
}

// Per-(chunk, head) attention for equal-length sequences of <= 160 rows with the head's q | k | v slice fetched by TMA.
// In k_attention_bf16_short every warp-level load touches 32 different rows (768-byte stride), i.e. 32 L1 tag lookups
// per instruction: an ablation (9 of 10 key blocks removed) still took 31 of the kernel's 41 us -- the per-CTA
// prologue, serialised in the LSU, was the bottleneck, not the exp2 / HMMA work.  TMA does the strided gather without
// the LSU, but the SM's TMA unit serves about one box ROW per 2.7 cycles whatever its width (three [160 x 32 B]
// boxes per CTA, 23 CTAs per SM = 30k cycles = the 17 us "empty kernel" time of the first TMA version).  The bf16
// qkv buffer is therefore head-interleaved (QKV_HEAD_STRIDE): ONE [160 rows x 128 B] box holds q | k | v of the head
// (+ 32 B of the next head, unused; zero-filled past column 384), 128B-swizzled so ldmatrix reads it conflict-free.
// Rows past the sequence inside a box belong to the next chunk: as keys they are masked (last block), as queries
// they are computed and never stored; past the end of the tensor TMA fills zeros.
template <bool F16>
__global__ void __launch_bounds__(160, 6) k_attention_bf16_tma(const __grid_constant__ CUtensorMap tmQKV, bf16* __restrict__ ctx,
                                                                int seq_len) {
  __shared__ __align__(1024) bf16 Ts[160 * 64];      // row r: 16-byte chunk c at c ^ (r & 7); chunks 0,1 = q, 2,3 = k, 4,5 = v
  __shared__ __align__(8) uint64_t bar;
  __shared__ float s_kmax[5];
  pdl_trigger();
  const int len = seq_len;
  const int head = blockIdx.y;
  const int64_t off = (int64_t)blockIdx.x * seq_len;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQKV);
    mbar_init(&bar, 1);
    fence_barrier_init();
    pdl_wait();                                      // qkv comes from the previous kernel in the stream
    mbar_arrive_expect_tx(&bar, 160 * 128);
    tma_load_2d(Ts, &tmQKV, &bar, head * QKV_HEAD_STRIDE, (int)off);
  }
  __syncthreads();                                   // barrier initialised
  mbar_wait(&bar, 0);
  {  // largest squared key norm, one key row per thread.  Rows past the sequence (the next chunk's rows inside the
     // box) are excluded: the result of a chunk must not depend on what follows it in the buffer.
    const int r7 = threadIdx.x & 7;
    const uint4 ka = *reinterpret_cast<const uint4*>(Ts + threadIdx.x * 64 + ((2 ^ r7) << 3));
    const uint4 kb = *reinterpret_cast<const uint4*>(Ts + threadIdx.x * 64 + ((3 ^ r7) << 3));
    const uint32_t kw[8] = {ka.x, ka.y, ka.z, ka.w, kb.x, kb.y, kb.z, kb.w};
    float kn2 = sqnorm16_h<F16>(kw);
    if ((int)threadIdx.x >= len) kn2 = 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kn2 = fmaxf(kn2, __shfl_xor_sync(0xffffffffu, kn2, o));
    if (lane == 0) s_kmax[warp] = kn2;
  }
  __syncthreads();
  const float kmax2 = fmaxf(fmaxf(fmaxf(s_kmax[0], s_kmax[1]), fmaxf(s_kmax[2], s_kmax[3])), s_kmax[4]);
  const int nkk = (len + 15) >> 4;
  constexpr uint32_t ONES = F16 ? 0x3C003C00u : 0x3F803F80u;
  // ldmatrix lane addressing in the [rows x 128 B] tile with the 128B swizzle: matrix m = lane / 8, row lane % 8;
  // every row base below is a multiple of 8, so the swizzle term is lane % 8
  const int lm = lane >> 3, lr = lane & 7;
  const bf16* k_lane = Ts + ((lm >> 1) * 8 + lr) * 64 + (((2 + (lm & 1)) ^ lr) << 3);   // m0/m1 = dh halves of keys 0-7, m2/m3 of keys 8-15
  const bf16* v_lane = Ts + ((lm & 1) * 8 + lr) * 64 + (((4 + (lm >> 1)) ^ lr) << 3);   // m0/m1 = key halves for dh 0-7, m2/m3 for dh 8-15
  const bf16* q_lane = Ts + ((lm & 1) * 8 + lr) * 64 + (((lm >> 1) ^ lr) << 3);         // a0..a3 of the m16k16 A fragment
#pragma unroll 1
  for (int mt = 0; mt < 2; ++mt) {
    const int row0 = warp * 32 + mt * 16;
    if (row0 >= len) break;                            // warp-uniform
    uint32_t qa[4];
    ldmatrix_x4(qa, q_lane + row0 * 64);
    // softmax stabiliser from the Cauchy-Schwarz bound (see k_attention_bf16_short); exact row maxima as fallback
    float qn0, qn1;
    {
      qn0 = sqnorm4_h<F16>(qa[0], qa[2]);               // row g:     a0 (k 0-1 of this lane's quad), a2 (k 8-9)
      qn1 = sqnorm4_h<F16>(qa[1], qa[3]);               // row g + 8: a1, a3
      qn0 += __shfl_xor_sync(0xffffffffu, qn0, 1); qn0 += __shfl_xor_sync(0xffffffffu, qn0, 2);
      qn1 += __shfl_xor_sync(0xffffffffu, qn1, 1); qn1 += __shfl_xor_sync(0xffffffffu, qn1, 2);
    }
    // margin: the packed-bf16 squared norms may each come out up to 1.6 % low (8 roundings of 2^-9), MUFU.SQRT 2 ulp
    float mx0 = sqrt_approx(qn0 * kmax2) * 1.02f, mx1 = sqrt_approx(qn1 * kmax2) * 1.02f;
    const bool big = !((row0 + g < len ? mx0 : 0.f) * (0.5f / QK_PRESCALE) < 80.f) || !((row0 + g + 8 < len ? mx1 : 0.f) * (0.5f / QK_PRESCALE) < 80.f);   // query rows past the sequence do not vote
    if (__any_sync(0xffffffffu, big)) {                         // warp-uniform; also taken for NaN / inf inputs
      mx0 = -INFINITY; mx1 = -INFINITY;
#pragma unroll 1
      for (int kk = 0; kk < nkk; ++kk) {
        uint32_t kf[4];
        ldmatrix_x4(kf, k_lane + kk * 16 * 64);
        float s0[4], s1[4];
        mma16_16816_z<F16>(s0, qa, kf[0], kf[1]);
        mma16_16816_z<F16>(s1, qa, kf[2], kf[3]);
        const int c = kk * 16 + 2 * t4;
        if (c >= len) { s0[0] = -INFINITY; s0[2] = -INFINITY; }
        if (c + 1 >= len) { s0[1] = -INFINITY; s0[3] = -INFINITY; }
        if (c + 8 >= len) { s1[0] = -INFINITY; s1[2] = -INFINITY; }
        if (c + 9 >= len) { s1[1] = -INFINITY; s1[3] = -INFINITY; }
        mx0 = fmaxf(mx0, fmaxf(fmaxf(s0[0], s0[1]), fmaxf(s1[0], s1[1])));
        mx1 = fmaxf(mx1, fmaxf(fmaxf(s0[2], s0[3]), fmaxf(s1[2], s1[3])));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    }
    const float nb0 = -mx0, nb1 = -mx1;               // k is pre-scaled (QK_PRESCALE): scores are base-2 exponents
    float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f}, ol[4] = {0.f, 0.f, 0.f, 0.f};
    auto pv_block = [&](int kk, auto mask_c) {
      constexpr bool MASK = decltype(mask_c)::value;
      uint32_t kf[4], vf[4];
      ldmatrix_x4(kf, k_lane + kk * 16 * 64);
      ldmatrix_x4_trans(vf, v_lane + kk * 16 * 64);
      float s0[4], s1[4];
      mma16_16816_c<F16>(s0, qa, kf[0], kf[1], nb0, nb1);   // s - stabiliser, straight out of the tensor core
      mma16_16816_c<F16>(s1, qa, kf[2], kf[3], nb0, nb1);
#pragma unroll
      for (int i = 0; i < 4; ++i) { s0[i] = ex2_approx(s0[i]); s1[i] = ex2_approx(s1[i]); }
      if (MASK) {
        const int c = kk * 16 + 2 * t4;
        if (c >= len) { s0[0] = 0.f; s0[2] = 0.f; }
        if (c + 1 >= len) { s0[1] = 0.f; s0[3] = 0.f; }
        if (c + 8 >= len) { s1[0] = 0.f; s1[2] = 0.f; }
        if (c + 9 >= len) { s1[1] = 0.f; s1[3] = 0.f; }
      }
      uint32_t pa[4];
      pa[0] = pack16<F16>(s0[0], s0[1]);
      pa[1] = pack16<F16>(s0[2], s0[3]);
      pa[2] = pack16<F16>(s1[0], s1[1]);
      pa[3] = pack16<F16>(s1[2], s1[3]);
      mma16_16816<F16>(o0, pa, vf[0], vf[1]);
      mma16_16816<F16>(o1, pa, vf[2], vf[3]);
      mma16_16816<F16>(ol, pa, ONES, ONES);              // row sums of the bf16-rounded P (every column is the row sum)
    };
#pragma unroll 3
    for (int kk = 0; kk < nkk - 1; ++kk) pv_block(kk, std::false_type{});
    pv_block(nkk - 1, std::true_type{});
    const float i0 = rcp_approx(ol[0]), i1 = rcp_approx(ol[2]);   // 1 ulp; the result is rounded to bf16 next
    const int r0 = row0 + g, r1 = row0 + g + 8;
    if (r0 < len) {
      bf16* op = ctx + (off + r0) * D + head * DH + 2 * t4;
      *reinterpret_cast<uint32_t*>(op) = pack16<F16>(o0[0] * i0, o0[1] * i0);
      *reinterpret_cast<uint32_t*>(op + 8) = pack16<F16>(o1[0] * i0, o1[1] * i0);
    }
    if (r1 < len) {
      bf16* op = ctx + (off + r1) * D + head * DH + 2 * t4;
      *reinterpret_cast<uint32_t*>(op) = pack16<F16>(o0[2] * i1, o0[3] * i1);
      *reinterpret_cast<uint32_t*>(op + 8) = pack16<F16>(o1[2] * i1, o1[3] * i1);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Attention for the tf32 mode (fp32 q / k / v in torch's [q128 | k128 | v128] layout, fp32 ctx out) on the tensor cores
// at fp32-class accuracy: every operand is split x = hi + lo with hi = bf16(x), lo = bf16(x - hi) (16 mantissa bits
// together) and every product takes three bf16 HMMAs, a.b ~= a_hi.b_hi + a_hi.b_lo + a_lo.b_hi (relative error
// ~2^-16, far below the tf32 GEMMs around it).  12 HMMAs per 16 x 16 block against k_attention_f32's 512 FMAs per
// thread-row; softmax in fp32 with exact fp32 row sums.  Sequences of at most 160 rows (every intra chunk); longer
// ones keep the FMA kernel.  Optionally rounds ctx to tf32 (the out-projection GEMM's operand format), which saves
// the separate rounding pass.
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16(a, b);
  const float2 h = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hi));
  lo = pack_bf16(a - h.x, b - h.y);
}

__global__ void __launch_bounds__(160, 3)
k_attention_split16(const float* __restrict__ qkv, float* __restrict__ ctx, int seq_len, const int* __restrict__ seq_off, int round_out) {
  constexpr int KMAX = 160, WARPS = 5, MT = 2, THREADS = 160;
  __shared__ __align__(16) bf16 Kh[KMAX * AS_ROW];
  __shared__ __align__(16) bf16 Kl[KMAX * AS_ROW];
  __shared__ __align__(16) bf16 Vh[KMAX * AS_ROW];
  __shared__ __align__(16) bf16 Vl[KMAX * AS_ROW];
  __shared__ float s_kmax[WARPS];
  const int seq = blockIdx.x;
  int off, len;
  if (seq_off != nullptr) {
    off = seq_off[seq];
    len = seq_off[seq + 1] - off;
  } else {
    off = seq * seq_len;
    len = seq_len;
  }
  if (len <= 0) return;
  const int head = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const float* base = qkv + (int64_t)off * (3 * D) + head * DH;
  {  // stage K / V rows as hi / lo bf16 (rows past the sequence up to the next multiple of 16 are zero); max squared key norm
    float kn2max = 0.f;
    const int kp = (len + 15) & ~15;
    for (int key = threadIdx.x; key < kp; key += THREADS) {
      float kv[16], vv[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) { kv[i] = 0.f; vv[i] = 0.f; }
      if (key < len) {
        const float4* pk = reinterpret_cast<const float4*>(base + (int64_t)key * (3 * D) + D);
        const float4* pv = reinterpret_cast<const float4*>(base + (int64_t)key * (3 * D) + 2 * D);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 a = __ldg(pk + i), b = __ldg(pv + i);
          kv[4 * i] = a.x; kv[4 * i + 1] = a.y; kv[4 * i + 2] = a.z; kv[4 * i + 3] = a.w;
          vv[4 * i] = b.x; vv[4 * i + 1] = b.y; vv[4 * i + 2] = b.z; vv[4 * i + 3] = b.w;
        }
      }
      uint32_t kh[8], kl[8], vh[8], vl[8];
      float kn2 = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        split_bf16x2(kv[2 * i], kv[2 * i + 1], kh[i], kl[i]);
        split_bf16x2(vv[2 * i], vv[2 * i + 1], vh[i], vl[i]);
        kn2 = fmaf(kv[2 * i], kv[2 * i], kn2);
        kn2 = fmaf(kv[2 * i + 1], kv[2 * i + 1], kn2);
      }
      uint4* d;
      d = reinterpret_cast<uint4*>(Kh + key * AS_ROW); d[0] = make_uint4(kh[0], kh[1], kh[2], kh[3]); d[1] = make_uint4(kh[4], kh[5], kh[6], kh[7]);
      d = reinterpret_cast<uint4*>(Kl + key * AS_ROW); d[0] = make_uint4(kl[0], kl[1], kl[2], kl[3]); d[1] = make_uint4(kl[4], kl[5], kl[6], kl[7]);
      d = reinterpret_cast<uint4*>(Vh + key * AS_ROW); d[0] = make_uint4(vh[0], vh[1], vh[2], vh[3]); d[1] = make_uint4(vh[4], vh[5], vh[6], vh[7]);
      d = reinterpret_cast<uint4*>(Vl + key * AS_ROW); d[0] = make_uint4(vl[0], vl[1], vl[2], vl[3]); d[1] = make_uint4(vl[4], vl[5], vl[6], vl[7]);
      kn2max = fmaxf(kn2max, kn2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kn2max = fmaxf(kn2max, __shfl_xor_sync(0xffffffffu, kn2max, o));
    if (lane == 0) s_kmax[warp] = kn2max;
  }
  __syncthreads();
  float kmax2 = s_kmax[0];
#pragma unroll
  for (int i = 1; i < WARPS; ++i) kmax2 = fmaxf(kmax2, s_kmax[i]);
  const int nkk = (len + 15) >> 4;
  constexpr float SC = 0.25f * 1.4426950408889634f;   // 1/sqrt(16) * log2(e)
  const int lm = lane >> 3, lr = lane & 7;
  const int k_off = ((lm >> 1) * 8 + lr) * AS_ROW + (lm & 1) * 8;   // K: m0/m1 = dh halves of keys 0-7, m2/m3 of keys 8-15
  const int v_off = ((lm & 1) * 8 + lr) * AS_ROW + (lm >> 1) * 8;   // V^T: m0/m1 = key halves for dh 0-7, m2/m3 for dh 8-15
#pragma unroll 1
  for (int mt = 0; mt < MT; ++mt) {
    const int row0 = warp * (16 * MT) + mt * 16;
    if (row0 >= len) break;                            // warp-uniform
    // Q fragment of the m16k16 A operand: a0 = (row g, k 2t4..), a1 = (row g+8, k 2t4..), a2 = (row g, k 8+2t4..), a3 = (row g+8, ..)
    uint32_t qh[4], ql[4];
    float qn0 = 0.f, qn1 = 0.f;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int r = row0 + g + hh * 8;
      float2 a = make_float2(0.f, 0.f), b = a;
      if (r < len) {
        const float* p = base + (int64_t)r * (3 * D);
        a = __ldg(reinterpret_cast<const float2*>(p + 2 * t4));
        b = __ldg(reinterpret_cast<const float2*>(p + 8 + 2 * t4));
      }
      split_bf16x2(a.x, a.y, qh[hh], ql[hh]);
      split_bf16x2(b.x, b.y, qh[hh + 2], ql[hh + 2]);
      const float n = a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y;
      if (hh == 0) qn0 = n; else qn1 = n;
    }
    qn0 += __shfl_xor_sync(0xffffffffu, qn0, 1); qn0 += __shfl_xor_sync(0xffffffffu, qn0, 2);
    qn1 += __shfl_xor_sync(0xffffffffu, qn1, 1); qn1 += __shfl_xor_sync(0xffffffffu, qn1, 2);
    // three-term score block: s = q_hi.k_hi + q_hi.k_lo + q_lo.k_hi for keys [16 kk, 16 kk + 16)
    auto scores = [&](int kk, float (&s0)[4], float (&s1)[4]) {
      uint32_t kh[4], kl[4];
      ldmatrix_x4(kh, Kh + k_off + kk * 16 * AS_ROW);
      ldmatrix_x4(kl, Kl + k_off + kk * 16 * AS_ROW);
      mma_bf16_16816_z(s0, qh, kl[0], kl[1]);
      mma_bf16_16816_z(s1, qh, kl[2], kl[3]);
      mma_bf16_16816(s0, ql, kh[0], kh[1]);
      mma_bf16_16816(s1, ql, kh[2], kh[3]);
      mma_bf16_16816(s0, qh, kh[0], kh[1]);
      mma_bf16_16816(s1, qh, kh[2], kh[3]);
    };
    // softmax stabiliser from the Cauchy-Schwarz bound (see k_attention_bf16_short); exact row maxima as fallback
    float mx0 = sqrtf(qn0 * kmax2) * 1.0001f, mx1 = sqrtf(qn1 * kmax2) * 1.0001f;
    const bool big = !(mx0 * 0.5f < 80.f) || !(mx1 * 0.5f < 80.f);
    if (__any_sync(0xffffffffu, big)) {
      mx0 = -INFINITY; mx1 = -INFINITY;
#pragma unroll 1
      for (int kk = 0; kk < nkk; ++kk) {
        float s0[4], s1[4];
        scores(kk, s0, s1);
        const int c = kk * 16 + 2 * t4;
        if (c >= len) { s0[0] = -INFINITY; s0[2] = -INFINITY; }
        if (c + 1 >= len) { s0[1] = -INFINITY; s0[3] = -INFINITY; }
        if (c + 8 >= len) { s1[0] = -INFINITY; s1[2] = -INFINITY; }
        if (c + 9 >= len) { s1[1] = -INFINITY; s1[3] = -INFINITY; }
        mx0 = fmaxf(mx0, fmaxf(fmaxf(s0[0], s0[1]), fmaxf(s1[0], s1[1])));
        mx1 = fmaxf(mx1, fmaxf(fmaxf(s0[2], s0[3]), fmaxf(s1[2], s1[3])));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    }
    const float nb0 = -mx0 * SC, nb1 = -mx1 * SC;
    float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
    float l0 = 0.f, l1 = 0.f;
#pragma unroll 1
    for (int kk = 0; kk < nkk; ++kk) {
      float s0[4], s1[4];
      scores(kk, s0, s1);
      s0[0] = ex2_approx(fmaf(s0[0], SC, nb0)); s0[1] = ex2_approx(fmaf(s0[1], SC, nb0));
      s0[2] = ex2_approx(fmaf(s0[2], SC, nb1)); s0[3] = ex2_approx(fmaf(s0[3], SC, nb1));
      s1[0] = ex2_approx(fmaf(s1[0], SC, nb0)); s1[1] = ex2_approx(fmaf(s1[1], SC, nb0));
      s1[2] = ex2_approx(fmaf(s1[2], SC, nb1)); s1[3] = ex2_approx(fmaf(s1[3], SC, nb1));
      if (kk == nkk - 1) {                             // only the last block can hold keys past the sequence
        const int c = kk * 16 + 2 * t4;
        if (c >= len) { s0[0] = 0.f; s0[2] = 0.f; }
        if (c + 1 >= len) { s0[1] = 0.f; s0[3] = 0.f; }
        if (c + 8 >= len) { s1[0] = 0.f; s1[2] = 0.f; }
        if (c + 9 >= len) { s1[1] = 0.f; s1[3] = 0.f; }
      }
      l0 += (s0[0] + s0[1]) + (s1[0] + s1[1]);
      l1 += (s0[2] + s0[3]) + (s1[2] + s1[3]);
      uint32_t ph[4], pl[4];
      split_bf16x2(s0[0], s0[1], ph[0], pl[0]);
      split_bf16x2(s0[2], s0[3], ph[1], pl[1]);
      split_bf16x2(s1[0], s1[1], ph[2], pl[2]);
      split_bf16x2(s1[2], s1[3], ph[3], pl[3]);
      uint32_t vh[4], vl[4];
      ldmatrix_x4_trans(vh, Vh + v_off + kk * 16 * AS_ROW);
      ldmatrix_x4_trans(vl, Vl + v_off + kk * 16 * AS_ROW);
      mma_bf16_16816(o0, ph, vl[0], vl[1]);
      mma_bf16_16816(o1, ph, vl[2], vl[3]);
      mma_bf16_16816(o0, pl, vh[0], vh[1]);
      mma_bf16_16816(o1, pl, vh[2], vh[3]);
      mma_bf16_16816(o0, ph, vh[0], vh[1]);
      mma_bf16_16816(o1, ph, vh[2], vh[3]);
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    float r[8] = {o0[0] * i0, o0[1] * i0, o1[0] * i0, o1[1] * i0, o0[2] * i1, o0[3] * i1, o1[2] * i1, o1[3] * i1};
    if (round_out) {
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = round_tf32(r[i]);
    }
    const int r0 = row0 + g, r1 = row0 + g + 8;
    if (r0 < len) {
      float* op = ctx + (int64_t)(off + r0) * D + head * DH + 2 * t4;
      *reinterpret_cast<float2*>(op) = make_float2(r[0], r[1]);
      *reinterpret_cast<float2*>(op + 8) = make_float2(r[2], r[3]);
    }
    if (r1 < len) {
      float* op = ctx + (int64_t)(off + r1) * D + head * DH + 2 * t4;
      *reinterpret_cast<float2*>(op) = make_float2(r[4], r[5]);
      *reinterpret_cast<float2*>(op + 8) = make_float2(r[6], r[7]);
    }
  }
}

// tf32-mode attention: tensor cores for sequences <= 160 rows, the FMA kernel otherwise.  `rounded` tells the caller
// whether ctx already is in tf32.
static int launch_attention_tf32mode(ResepHandle* h, const float* qkv, float* ctx, int n_seq, int seq_len, const int* seq_off,
                                     const int* tile_seq, const int* tile_q0, int n_tiles, int max_len, cudaStream_t st, bool* rounded) {
  static const bool use_split = !(getenv("RESEP_ATTN_SPLIT16") && getenv("RESEP_ATTN_SPLIT16")[0] == '0');
  const int longest = seq_off == nullptr ? seq_len : max_len;
  *rounded = false;
  if (!use_split || longest > 160) return launch_attention_f32(h, qkv, ctx, n_seq, seq_len, seq_off, tile_seq, tile_q0, n_tiles, st);
  if (n_seq == 0) return RESEP_OK;
  ProfScope prof_scope(h, "k_attention_split16", st);
  k_attention_split16<<<dim3((unsigned)n_seq, NH), 160, 0, st>>>(qkv, ctx, seq_len, seq_off, 1);
  RESEP_LAUNCH_CHECK(h, "k_attention_split16");
  *rounded = true;
  return RESEP_OK;
}

static int launch_attention_bf16(ResepHandle* h, const bf16* qkv, bf16* ctx, int n_seq, int seq_len, const int* seq_off,
                                 const int* tile_seq, const int* tile_q0, int n_tiles128, int max_len, cudaStream_t st, bool intra) {
  // The intra blocks (every sequence is a 150-row chunk) run on tcgen05.  The memory transformer's sequences never do,
  // whatever their length: which of ITS kernels a sequence gets depends on the batch composition, and those are
  // bit-identical to one another by construction -- the tcgen05 kernel's arithmetic is not.
  static const bool use_tc = attn_tc_default() ? !(getenv("RESEP_ATTN_TC") && getenv("RESEP_ATTN_TC")[0] == '0')
                                               : (getenv("RESEP_ATTN_TC") && getenv("RESEP_ATTN_TC")[0] == '1');
  if (intra && use_tc && tile_seq == nullptr && seq_len == CHUNK && (int64_t)n_seq * seq_len < 2000000000LL)
    return launch_attn_tc(h, qkv, ctx, n_seq, st);
  const int longest = tile_seq == nullptr ? seq_len : max_len;
  const bool f16 = h->fmt16 != 0;
  ProfScope prof_scope(h, longest <= 160 ? (tile_seq == nullptr ? "k_attention_bf16_tma" : "k_attention_bf16_short") : longest <= 448 ? "k_attention_bf16_mid" : "k_attention_bf16", st);
  // ragged case: the plan's tile list is cut in 128-row tiles for the fp32 kernel; this kernel covers
  // 160 rows per CTA, so a 128-row tile list still covers every row (rows 128..159 of a tile repeat work
  // of the next tile with identical results).
  if (tile_seq == nullptr) {
    if (n_seq == 0) return RESEP_OK;
    static const bool use_tma = !(getenv("RESEP_ATTN_TMA") && getenv("RESEP_ATTN_TMA")[0] == '0');
    if (seq_len <= 160 && use_tma && (int64_t)n_seq * seq_len < 2000000000LL) {
      CUtensorMap tmQKV;
      int rc = make_tmap_head(h, &tmQKV, qkv, (int64_t)n_seq * seq_len, 3 * D, 160);
      if (rc) return rc;
      RESEP_CUDA(h, launch_pdl(f16 ? k_attention_bf16_tma<true> : k_attention_bf16_tma<false>, dim3((unsigned)n_seq, NH), dim3(160), 0, st, tmQKV, ctx, seq_len));
    } else if (seq_len <= 160) {
      RESEP_CUDA(h, launch_pdl(f16 ? k_attention_bf16_short<160, 5, 2, true> : k_attention_bf16_short<160, 5, 2, false>, dim3((unsigned)n_seq, NH), dim3(160), 0, st, qkv, ctx, seq_len,
                               (const int*)nullptr, 1));
    } else if (seq_len <= 448) {
      const int qt = (seq_len + 63) / 64;
      RESEP_CUDA(h, launch_pdl(f16 ? k_attention_bf16_short<448, 4, 1, true> : k_attention_bf16_short<448, 4, 1, false>, dim3((unsigned)(n_seq * qt), NH), dim3(128), 0, st, qkv, ctx, seq_len,
                               (const int*)nullptr, qt));
    } else {
      const int tps = (seq_len + AKT - 1) / AKT;
      if (f16) k_attention_bf16<true><<<dim3((unsigned)(n_seq * tps), NH), 160, 0, st>>>(qkv, ctx, seq_len, nullptr, nullptr, nullptr);
      else k_attention_bf16<false><<<dim3((unsigned)(n_seq * tps), NH), 160, 0, st>>>(qkv, ctx, seq_len, nullptr, nullptr, nullptr);
    }
  } else if (max_len <= 160) {
    if (n_seq == 0) return RESEP_OK;
    RESEP_CUDA(h, launch_pdl(f16 ? k_attention_bf16_short<160, 5, 2, true> : k_attention_bf16_short<160, 5, 2, false>, dim3((unsigned)n_seq, NH), dim3(160), 0, st, qkv, ctx, 0, seq_off, 1));
  } else if (max_len <= 448) {
    if (n_seq == 0) return RESEP_OK;
    const int qt = (max_len + 63) / 64;
    RESEP_CUDA(h, launch_pdl(f16 ? k_attention_bf16_short<448, 4, 1, true> : k_attention_bf16_short<448, 4, 1, false>, dim3((unsigned)(n_seq * qt), NH), dim3(128), 0, st, qkv, ctx, 0, seq_off, qt));
  } else {
    if (n_tiles128 == 0) return RESEP_OK;
    if (f16) k_attention_bf16<true><<<dim3((unsigned)n_tiles128, NH), 160, 0, st>>>(qkv, ctx, 0, seq_off, tile_seq, tile_q0);
    else k_attention_bf16<false><<<dim3((unsigned)n_tiles128, NH), 160, 0, st>>>(qkv, ctx, 0, seq_off, tile_seq, tile_q0);
  }
  RESEP_LAUNCH_CHECK(h, "k_attention_bf16");
  return RESEP_OK;
}

// ------------------------------------------------------------------------------------------------
// small helpers: dtype conversion for the test hook
template <typename OutT>
__global__ void k_convert(const float* __restrict__ x, OutT* __restrict__ y, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if constexpr (sizeof(OutT) == 2) y[i] = __float2bfloat16_rn(x[i]);
  else y[i] = round_tf32(x[i]);
}

__global__ void k_split_tf32(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float h = round_tf32(w[i]);
  hi[i] = h;
  lo[i] = round_tf32(w[i] - h);
}

__global__ void k_split_w16(const float* __restrict__ w, bf16* __restrict__ hi, bf16* __restrict__ lo, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bf16 h = __float2bfloat16_rn(w[i]);
  hi[i] = h;
  lo[i] = __float2bfloat16_rn(w[i] - __bfloat162float(h));
}

int tc_linear_test(ResepHandle* h, const float* A, const float* W, const float* bias, float* out, int64_t M, int N, int K,
                   bool relu, int precision, cudaStream_t st) {
  const bool is16 = precision >= RESEP_PREC_BF16;   // 2: the handle's bf16 weight mode; test-only codes 3: bf16(W) only, 4: hi+lo
  const size_t esz = is16 ? 2 : 4;
  void *a2 = nullptr, *w2 = nullptr, *w3 = nullptr;
  RESEP_CUDA(h, cudaMalloc(&a2, (size_t)M * K * esz));
  RESEP_CUDA(h, cudaMalloc(&w2, (size_t)N * K * esz));
  RESEP_CUDA(h, cudaMalloc(&w3, (size_t)N * K * esz));
  int rc;
  if (is16) {
    k_convert<bf16><<<(unsigned)((M * K + 255) / 256), 256, 0, st>>>(A, (bf16*)a2, M * K);
    k_split_w16<<<(unsigned)(((int64_t)N * K + 255) / 256), 256, 0, st>>>(W, (bf16*)w2, (bf16*)w3, (int64_t)N * K);
    h->launches += 2;
    rc = gemm_bf16<EPI_STORE_F32>(h, precision == 2 ? h->w16_mode : precision - 3, (bf16*)a2, (bf16*)w2, (bf16*)w3, bias, out, M, N, K,
                                  relu, st);
  } else {
    k_convert<float><<<(unsigned)((M * K + 255) / 256), 256, 0, st>>>(A, (float*)a2, M * K);
    k_split_tf32<<<(unsigned)(((int64_t)N * K + 255) / 256), 256, 0, st>>>(W, (float*)w2, (float*)w3, (int64_t)N * K);
    h->launches += 2;
    rc = launch_gemm_tc<float, EPI_STORE_F32, true>(h, (float*)a2, (float*)w2, bias, out, M, N, K, relu, st, (float*)w3);
  }
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(a2);
  cudaFree(w2);
  cudaFree(w3);
  if (rc) return rc;
  if (e != cudaSuccess) return set_err(h, RESEP_ECUDA, std::string("tc_linear_test: ") + cudaGetErrorString(e));
  return RESEP_OK;
}

// ------------------------------------------------------------------------------------------------
// One transformer layer with tensor-core GEMMs.
//   bf16 mode: y / qkv / ctx / hid are bf16 (half the bytes of the fp32 scratch regions they live in)
//   tf32 mode: fp32 buffers holding tf32-rounded values; attention uses the fp32 kernel
bool g_serial_launches = false;

// Profiling aid behind resep_layer_kernel_repeat: ONE of the layer's three fused 16-bit kernels, `reps` times back to
// back on the buffers a previous tc_run_layer call left behind.
int tc_repeat_layer_kernel(ResepHandle* h, const LayerDev& lw, int which, float* o, int64_t rows, int n_seq, int seq_len,
                           float* qkv, float* ctx, int reps, bool pdl, cudaStream_t st, bool intra) {
  bf16 *qb = reinterpret_cast<bf16*>(qkv), *cb = reinterpret_cast<bf16*>(ctx);
  g_serial_launches = !pdl;
  int rc = RESEP_OK;
  for (int i = 0; i < reps && rc == RESEP_OK; ++i)
    rc = which == 0 ? launch_qkv2_tc(h, lw, o, qb, rows, st)
       : which == 1 ? launch_attention_bf16(h, qb, cb, n_seq, seq_len, nullptr, nullptr, nullptr, 0, seq_len, st, intra)
                    : launch_post2_tc(h, lw, cb, o, rows, st);
  g_serial_launches = false;
  return rc;
}

int tc_run_layer(ResepHandle* h, const LayerDev& lw, float* o, int64_t rows, int n_seq, int seq_len, const int* seq_off,
                 const int* tile_seq, const int* tile_q0, int n_tiles, int max_seq_len, float* y, float* qkv, float* ctx, float* hid,
                 int precision, cudaStream_t st, bool intra) {
  int rc;
  if (precision == RESEP_PREC_BF16) {
    bf16 *yb = reinterpret_cast<bf16*>(y), *qb = reinterpret_cast<bf16*>(qkv), *cb = reinterpret_cast<bf16*>(ctx),
         *hb = reinterpret_cast<bf16*>(hid);
    static const bool fused = !(getenv("RESEP_FUSED") && getenv("RESEP_FUSED")[0] == '0');
    if (h->fmt16) {   // the fp16 mode exists as instances of the CTA-pair kernels and the attention kernels only
      if ((rc = launch_qkv2_tc(h, lw, o, qb, rows, st))) return rc;
      if ((rc = launch_attention_bf16(h, qb, cb, n_seq, seq_len, seq_off, tile_seq, tile_q0, n_tiles, max_seq_len, st, intra))) return rc;
      return launch_post2_tc(h, lw, cb, o, rows, st);
    }
    if (fused) {
      static const bool qkv2 = !(getenv("RESEP_QKV2") && getenv("RESEP_QKV2")[0] == '0');
      if ((rc = qkv2 ? launch_qkv2_tc(h, lw, o, qb, rows, st) : launch_qkv_tc(h, lw, o, qb, rows, st))) return rc;   // norm1 + in-projection in one kernel
    } else {
      if ((rc = launch_layernorm<bf16>(h, o, lw.norm1_w, lw.norm1_b, yb, rows, st))) return rc;
      if ((rc = gemm_bf16<EPI_STORE_BF16>(h, h->w16_mode, yb, lw.in_w_bf, lw.in_w_bl, lw.in_b_hi, qb, rows, 3 * D, D, false, st))) return rc;
    }
    if ((rc = launch_attention_bf16(h, qb, cb, n_seq, seq_len, seq_off, tile_seq, tile_q0, n_tiles, max_seq_len, st, intra))) return rc;
    static const bool post2 = !(getenv("RESEP_POST2") && getenv("RESEP_POST2")[0] == '0');
    if (fused) return post2 ? launch_post2_tc(h, lw, cb, o, rows, st) : launch_post_tc(h, lw, cb, o, rows, st);   // out-proj + LN2 + FFN in one kernel
    if ((rc = gemm_bf16<EPI_RESID_F32>(h, h->w16_mode, cb, lw.out_w_bf, lw.out_w_bl, lw.out_b, o, rows, D, D, false, st))) return rc;
    if ((rc = launch_layernorm<bf16>(h, o, lw.norm2_w, lw.norm2_b, yb, rows, st))) return rc;
    if ((rc = gemm_bf16<EPI_STORE_BF16>(h, h->w16_mode == 1, yb, lw.f1_w_bf, lw.f1_w_bl, lw.f1_b, hb, rows, FFN, D, true, st))) return rc;
    if ((rc = gemm_bf16<EPI_RESID_F32>(h, h->w16_mode == 1, hb, lw.f2_w_bf, lw.f2_w_bl, lw.f2_b, o, rows, D, FFN, false, st))) return rc;
    return RESEP_OK;
  }
  // tf32: operands are fp32 in memory; the tensor core reads the top 19 bits, so every producer rounds
  // to tf32 (cvt.rna) first and the weights were rounded on upload.
  if ((rc = launch_layernorm_tf32(h, o, lw.norm1_w, lw.norm1_b, y, rows, st))) return rc;
  if ((rc = launch_gemm_tc<float, EPI_STORE_F32, true>(h, y, lw.in_w_tf, lw.in_b, qkv, rows, 3 * D, D, false, st, lw.in_w_lo))) return rc;
  bool ctx_rounded = false;
  if ((rc = launch_attention_tf32mode(h, qkv, ctx, n_seq, seq_len, seq_off, tile_seq, tile_q0, n_tiles, max_seq_len, st, &ctx_rounded))) return rc;
  if (!ctx_rounded && (rc = launch_round_tf32(h, ctx, rows * D, st))) return rc;
  if ((rc = launch_gemm_tc<float, EPI_RESID_F32, true>(h, ctx, lw.out_w_tf, lw.out_b, o, rows, D, D, false, st, lw.out_w_lo))) return rc;
  if ((rc = launch_layernorm_tf32(h, o, lw.norm2_w, lw.norm2_b, y, rows, st))) return rc;
  if ((rc = launch_gemm_tc<float, EPI_STORE_TF32, true>(h, y, lw.f1_w_tf, lw.f1_b, hid, rows, FFN, D, true, st, lw.f1_w_lo))) return rc;
  if ((rc = launch_gemm_tc<float, EPI_RESID_F32, true>(h, hid, lw.f2_w_tf, lw.f2_b, o, rows, D, FFN, false, st, lw.f2_w_lo))) return rc;
  return RESEP_OK;
}

int tc_run_mask(ResepHandle* h, const float* a, float* y_scratch, float* mask, int64_t M, int precision,
                cudaStream_t st, bool prelu_done) {
  int rc;
  if (precision == RESEP_PREC_BF16) {
    bf16* yb = reinterpret_cast<bf16*>(y_scratch);
    if (!prelu_done && (rc = launch_prelu_t<bf16>(h, a, h->w.prelu_a, yb, M * D, st))) return rc;
    if (h->fmt16 && !prelu_done) return set_err(h, RESEP_EINVAL, "fp16 mode: the mask GEMM expects the block epilogue's fp16 PReLU output");
    return gemm_bf16<EPI_STORE_F32>(h, h->w16_mode, yb, h->fmt16 ? h->w.fc_w_h[0] : h->w.fc_w_bf, h->fmt16 ? h->w.fc_w_h[1] : h->w.fc_w_bl,
                                   h->w.fc_b, mask, M, NSPK * D, D, true, st);
  }
  if ((rc = launch_prelu(h, a, h->w.prelu_a, y_scratch, M * D, st))) return rc;
  if ((rc = launch_round_tf32(h, y_scratch, M * D, st))) return rc;
  return launch_gemm_tc<float, EPI_STORE_F32, true>(h, y_scratch, h->w.fc_w_tf, h->w.fc_b, mask, M, NSPK * D, D, true, st,
                                                  h->w.fc_w_lo);
}

}  // namespace resep
