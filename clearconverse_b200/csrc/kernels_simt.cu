// fp32 FMA kernels of the RE-SepFormer path: the HBM-bound ends of the pipeline (encoder,
// prologue, LayerNorm, gLN epilogue, decoder) used by every precision mode, plus fp32 GEMM and
// attention kernels that make up the reference-grade RESEP_PREC_FP32 mode against which the
// tcgen05 modes are checked on the device.
//
// Upstream arithmetic each kernel restates (speechbrain, see oracle/resepformer_oracle.py):
//   encoder      dual_path.Encoder.forward           relu(conv1d(1->128, k16, s8, no bias))
//   prologue     SBTransformerBlock_wnormandskip     x + pos_enc(x); pipeline `output + hc`
//   layernorm    sb.nnet.normalization.LayerNorm     eps 1e-6
//   attention    nn.MultiheadAttention (8 heads x 16), no masks, softmax over all keys
//   epilogue     TransformerEncoder.norm + dual_path.GlobalLayerNorm + skip; `output.mean(1)`
//   decoder      mask apply + dual_path.Decoder (ConvTranspose1d 128->1, k16, s8) + pad/crop
#include <cuda_fp16.h>

#include "resep_internal.cuh"

namespace resep {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------
// Encoder (dual_path.Encoder: relu(conv1d k16 s8), no bias) writing token-major, chunk-padded rows.  One CTA = 30
// consecutive rows of one chunk (5 CTAs per chunk); a WARP produces a row: lane l holds the 16 taps of filters
// 4 l .. 4 l + 3 in registers, reads the row's 16 samples as four broadcast 16-byte loads from the staged window and
// stores one float4 -- a 512-byte row per store instruction.  Rows past the item's last frame (the `rest` zero padding
// of _padfeature) are written as zeros.  With `o` != null the first transformer block's prologue is fused in:
// o = x0 + pe[frame in chunk] (hc is zero for the first block, so its skip input is x0 itself): the separate
// k_block_prologue pass over the same 33 MB (config 2) disappears.
constexpr int ENC_ROWS = 50;            // rows per CTA (A/B on config 2: 30 rows 37.3 us, 50 rows 33.0 us, 75 rows 37.5 us per launch)
__global__ void __launch_bounds__(D) k_encoder_chunked(const float* __restrict__ mix, const float* __restrict__ enc_w,
                                                       const int64_t* __restrict__ item_off,
                                                       const int64_t* __restrict__ item_len,
                                                       const int* __restrict__ item_L, const int* __restrict__ chunk_item,
                                                       const int* __restrict__ chunk_frame0, float* __restrict__ x0,
                                                       float* __restrict__ o, const float* __restrict__ pe) {
  __shared__ __align__(16) float s[ENC_ROWS * STRIDE + KSZ];
  const int chunk = blockIdx.x / (CHUNK / ENC_ROWS);
  const int r0 = (blockIdx.x % (CHUNK / ENC_ROWS)) * ENC_ROWS;
  const int item = chunk_item[chunk];
  const int frame0 = chunk_frame0[chunk] + r0;
  const int L = item_L[item];
  const int64_t T = item_len[item];
  const float* src = mix + item_off[item];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < ENC_ROWS * STRIDE + KSZ; i += D) {
    const int64_t t = (int64_t)frame0 * STRIDE + i;
    s[i] = (t < T) ? src[t] : 0.f;
  }
  float w[4][KSZ];
#pragma unroll
  for (int f = 0; f < 4; ++f)
#pragma unroll
    for (int k = 0; k < KSZ; k += 4) {
      const float4 t4 = *reinterpret_cast<const float4*>(enc_w + (4 * lane + f) * KSZ + k);
      w[f][k] = t4.x; w[f][k + 1] = t4.y; w[f][k + 2] = t4.z; w[f][k + 3] = t4.w;
    }
  __syncthreads();
  float4* dst = reinterpret_cast<float4*>(x0 + ((int64_t)chunk * CHUNK + r0) * D) + lane;
  float4* dsto = o != nullptr ? reinterpret_cast<float4*>(o + ((int64_t)chunk * CHUNK + r0) * D) + lane : nullptr;
  const float4* pe4 = reinterpret_cast<const float4*>(pe + (int64_t)r0 * D) + lane;       // position in the chunk = row in the chunk
  for (int j = warp; j < ENC_ROWS; j += D / 32) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (frame0 + j < L) {
      float x[KSZ];
#pragma unroll
      for (int k = 0; k < KSZ; k += 4) {
        const float4 t4 = *reinterpret_cast<const float4*>(&s[j * STRIDE + k]);
        x[k] = t4.x; x[k + 1] = t4.y; x[k + 2] = t4.z; x[k + 3] = t4.w;
      }
#pragma unroll
      for (int f = 0; f < 4; ++f) {
#pragma unroll
        for (int k = 0; k < KSZ; ++k) acc[f] = fmaf(w[f][k], x[k], acc[f]);     // same order as the one-filter-per-thread form
        acc[f] = fmaxf(acc[f], 0.f);
      }
    }
    dst[(int64_t)j * (D / 4)] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    if (dsto != nullptr) {
      const float4 e = pe4[(int64_t)j * (D / 4)];
      dsto[(int64_t)j * (D / 4)] = make_float4(acc[0] + e.x, acc[1] + e.y, acc[2] + e.z, acc[3] + e.w);
    }
  }
}

int launch_encoder_chunked(ResepHandle* h, const float* mix, const Plan& p, float* x0, cudaStream_t st, float* o_fused) {
  ProfScope prof_scope_69(h, "k_encoder_chunked", st);
  k_encoder_chunked<<<(unsigned)(p.n_chunks * (CHUNK / ENC_ROWS)), D, 0, st>>>(
      mix, h->w.enc_w, p.d_item_off, p.d_item_len, p.d_item_L, p.d_chunk_item, p.d_chunk_frame0, x0, o_fused, h->w.pe);
  RESEP_LAUNCH_CHECK(h, "k_encoder_chunked");
  return RESEP_OK;
}

__global__ void __launch_bounds__(D) k_encoder_single(const float* __restrict__ mix, const float* __restrict__ enc_w,
                                                      int64_t T, int L, float* __restrict__ out) {
  __shared__ float s[ENC_ROWS * STRIDE + KSZ];
  const int frame0 = blockIdx.x * ENC_ROWS;
  const int n = threadIdx.x;
  for (int i = n; i < ENC_ROWS * STRIDE + KSZ - STRIDE; i += D) {
    int64_t t = (int64_t)frame0 * STRIDE + i;
    s[i] = (t < T) ? mix[t] : 0.f;
  }
  float w[KSZ];
#pragma unroll
  for (int k = 0; k < KSZ; ++k) w[k] = enc_w[n * KSZ + k];
  __syncthreads();
  for (int j = 0; j < ENC_ROWS && frame0 + j < L; ++j) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < KSZ; ++k) acc = fmaf(w[k], s[j * STRIDE + k], acc);
    out[((int64_t)frame0 + j) * D + n] = fmaxf(acc, 0.f);
  }
}

int launch_encoder_single(ResepHandle* h, const float* mix, int64_t T, float* tokens, cudaStream_t st) {
  int L = (int)((T - KSZ) / STRIDE + 1);
  ProfScope prof_scope_98(h, "k_encoder_single", st);
  k_encoder_single<<<(L + ENC_ROWS - 1) / ENC_ROWS, D, 0, st>>>(mix, h->w.enc_w, T, L, tokens);
  RESEP_LAUNCH_CHECK(h, "k_encoder_single");
  return RESEP_OK;
}

// ------------------------------------------------------------------------------------------
// Block prologue: (xin = xprev + hc[row / seq_len])?; o = xin + pe[pos(row)].  One float4 per thread.
__global__ void k_block_prologue(const float* __restrict__ xprev, const float* __restrict__ hc, float* xin,
                                 float* __restrict__ o, int64_t n4, const float* __restrict__ pe,
                                 const int* __restrict__ pos, int seq_len) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int64_t row = i / (D / 4);
  const int c4 = (int)(i % (D / 4));
  float4 v = reinterpret_cast<const float4*>(xprev)[i];
  if (hc != nullptr) {
    float4 a = reinterpret_cast<const float4*>(hc)[(row / seq_len) * (D / 4) + c4];
    v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    reinterpret_cast<float4*>(xin)[i] = v;
  }
  const int p = pos ? pos[row] : (int)(row % seq_len);
  float4 e = reinterpret_cast<const float4*>(pe)[(int64_t)p * (D / 4) + c4];
  v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
  reinterpret_cast<float4*>(o)[i] = v;
}

int launch_block_prologue(ResepHandle* h, const float* xprev, const float* hc, float* xin, float* o, int64_t rows,
                          const int* pos, int seq_len, cudaStream_t st) {
  int64_t n4 = rows * (D / 4);
  if (n4 == 0) return RESEP_OK;
  ProfScope prof_scope_128(h, "k_block_prologue", st);
  k_block_prologue<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(xprev, hc, xin, o, n4, h->w.pe, pos, seq_len);
  RESEP_LAUNCH_CHECK(h, "k_block_prologue");
  return RESEP_OK;
}

// ------------------------------------------------------------------------------------------
// LayerNorm over rows of 128: one warp per row, one float4 per lane, two-pass variance.
__device__ __forceinline__ float4 ln_row(float4 v, float4 w, float4 b) {
  float mean = warp_sum(v.x + v.y + v.z + v.w) * (1.f / D);
  float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
  float var = warp_sum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.f / D);
  float rstd = rsqrtf(var + LN_EPS);
  return make_float4(dx * rstd * w.x + b.x, dy * rstd * w.y + b.y, dz * rstd * w.z + b.z, dw * rstd * w.w + b.w);
}

__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(bf16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

__device__ __forceinline__ void store4(__half* p, float4 v) {
  const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<const uint32_t*>(&a);
  u.y = *reinterpret_cast<const uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
struct tf32_t { float v; };   // tag type: fp32 storage holding tf32-rounded values
__device__ __forceinline__ void store4(tf32_t* p, float4 v) {
  *reinterpret_cast<float4*>(p) = make_float4(rna_tf32(v.x), rna_tf32(v.y), rna_tf32(v.z), rna_tf32(v.w));
}

template <typename OutT>
__global__ void __launch_bounds__(256) k_layernorm(const float* __restrict__ x, const float* __restrict__ w,
                                                   const float* __restrict__ b, OutT* __restrict__ y, int64_t rows) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float4 v = reinterpret_cast<const float4*>(x + row * D)[lane];
  float4 r = ln_row(v, reinterpret_cast<const float4*>(w)[lane], reinterpret_cast<const float4*>(b)[lane]);
  store4(y + row * D + lane * 4, r);
}

template <typename OutT>
int launch_layernorm(ResepHandle* h, const float* x, const float* w, const float* b, OutT* y, int64_t rows,
                     cudaStream_t st) {
  if (rows == 0) return RESEP_OK;
  ProfScope prof_scope_177(h, "k_layernorm", st);
  k_layernorm<OutT><<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, w, b, y, rows);
  RESEP_LAUNCH_CHECK(h, "k_layernorm");
  return RESEP_OK;
}
template int launch_layernorm<float>(ResepHandle*, const float*, const float*, const float*, float*, int64_t, cudaStream_t);
template int launch_layernorm<bf16>(ResepHandle*, const float*, const float*, const float*, bf16*, int64_t, cudaStream_t);
int launch_layernorm_tf32(ResepHandle* h, const float* x, const float* w, const float* b, float* y, int64_t rows,
                          cudaStream_t st) {
  return launch_layernorm<tf32_t>(h, x, w, b, reinterpret_cast<tf32_t*>(y), rows, st);
}

__global__ void k_round_tf32(float* x, int64_t n4) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = reinterpret_cast<float4*>(x)[i];
  reinterpret_cast<float4*>(x)[i] = make_float4(rna_tf32(v.x), rna_tf32(v.y), rna_tf32(v.z), rna_tf32(v.w));
}
int launch_round_tf32(ResepHandle* h, float* x, int64_t n, cudaStream_t st) {
  const int64_t n4 = n / 4;
  if (n4 == 0) return RESEP_OK;
  ProfScope prof_scope_197(h, "k_round_tf32", st);
  k_round_tf32<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(x, n4);
  RESEP_LAUNCH_CHECK(h, "k_round_tf32");
  return RESEP_OK;
}

// ------------------------------------------------------------------------------------------
// fp32 GEMM  C[M,N] = A[M,K] . W[N,K]^T + bias (+relu) (+residual).  64x64x16 tiles, 256 threads,
// 4x4 outputs per thread.  N % 64 == 0, K % 16 == 0 (true for every matrix of this model).
constexpr int GB = 64, GK = 16;
__global__ void __launch_bounds__(256) k_gemm_f32(const float* __restrict__ A, const float* __restrict__ W,
                                                  const float* __restrict__ bias, const float* residual, float* C,
                                                  int64_t M, int N, int K, int relu) {
  __shared__ __align__(16) float As[GK][GB + 4];
  __shared__ __align__(16) float Ws[GK][GB + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * GB;
  const int n0 = blockIdx.y * GB;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  float acc[4][4] = {};
  const bool arow_ok = (m0 + lrow) < M;
  const float* ap = A + (m0 + lrow) * K + lk;
  const float* wp = W + (int64_t)(n0 + lrow) * K + lk;
  for (int k0 = 0; k0 < K; k0 += GK) {
    float4 a = arow_ok ? *reinterpret_cast<const float4*>(ap + k0) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 w = *reinterpret_cast<const float4*>(wp + k0);
    As[lk + 0][lrow] = a.x; As[lk + 1][lrow] = a.y; As[lk + 2][lrow] = a.z; As[lk + 3][lrow] = a.w;
    Ws[lk + 0][lrow] = w.x; Ws[lk + 1][lrow] = w.y; Ws[lk + 2][lrow] = w.z; Ws[lk + 3][lrow] = w.w;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 wv = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w};
      const float ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], ww[j], acc[i][j]);
    }
    __syncthreads();
  }
  float4 bv = *reinterpret_cast<const float4*>(bias + n0 + tx * 4);
  const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      o[j] = acc[i][j] + bb[j];
      if (relu) o[j] = fmaxf(o[j], 0.f);
    }
    float* cp = C + m * N + n0 + tx * 4;
    if (residual != nullptr) {
      float4 r = *reinterpret_cast<const float4*>(residual + m * N + n0 + tx * 4);
      o[0] += r.x; o[1] += r.y; o[2] += r.z; o[3] += r.w;
    }
    *reinterpret_cast<float4*>(cp) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

int launch_gemm_f32(ResepHandle* h, const float* A, const float* W, const float* bias, const float* residual, float* C,
                    int64_t M, int N, int K, bool relu, cudaStream_t st) {
  if (M == 0) return RESEP_OK;
  if (N % GB != 0 || K % GK != 0) return set_err(h, RESEP_EINVAL, "gemm_f32: N % 64 or K % 16 != 0");
  dim3 grid((unsigned)((M + GB - 1) / GB), (unsigned)(N / GB));
  ProfScope prof_scope_265(h, "k_gemm_f32", st);
  k_gemm_f32<<<grid, 256, 0, st>>>(A, W, bias, residual, C, M, N, K, relu ? 1 : 0);
  RESEP_LAUNCH_CHECK(h, "k_gemm_f32");
  return RESEP_OK;
}

// ------------------------------------------------------------------------------------------
// fp32 attention, any sequence length.  One CTA = one (query tile of QT rows, head) of one
// sequence; one thread per query row; K/V of the head are streamed through shared memory in
// tiles of 128 keys with a running (max, sum) softmax.  q is scaled by 1/sqrt(16) first, as
// nn.MultiheadAttention does.
constexpr int ATT_KT = 128;
template <int QT>
__global__ void __launch_bounds__(QT) k_attention_f32(const float* __restrict__ qkv, float* __restrict__ ctx,
                                                      int seq_len, const int* __restrict__ seq_off,
                                                      const int* __restrict__ tile_seq,
                                                      const int* __restrict__ tile_q0) {
  __shared__ __align__(16) float Ks[ATT_KT][DH];
  __shared__ __align__(16) float Vs[ATT_KT][DH];
  int seq, q0, off, len;
  if (tile_seq != nullptr) {
    seq = tile_seq[blockIdx.x];
    q0 = tile_q0[blockIdx.x];
    off = seq_off[seq];
    len = seq_off[seq + 1] - off;
  } else {
    const int tps = (seq_len + QT - 1) / QT;
    seq = blockIdx.x / tps;
    q0 = (blockIdx.x % tps) * QT;
    off = seq * seq_len;
    len = seq_len;
  }
  const int head = blockIdx.y;
  const int qi = q0 + threadIdx.x;
  const bool active = qi < len;
  float q[DH], acc[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) acc[d] = 0.f;
  if (active) {
    const float4* qp = reinterpret_cast<const float4*>(qkv + (int64_t)(off + qi) * (3 * D) + head * DH);
#pragma unroll
    for (int d4 = 0; d4 < DH / 4; ++d4) {
      float4 v = qp[d4];
      q[d4 * 4 + 0] = v.x * 0.25f; q[d4 * 4 + 1] = v.y * 0.25f; q[d4 * 4 + 2] = v.z * 0.25f; q[d4 * 4 + 3] = v.w * 0.25f;
    }
  }
  float mrun = -INFINITY, lrun = 0.f;
  for (int kt = 0; kt < len; kt += ATT_KT) {
    const int nk = min(ATT_KT, len - kt);
    __syncthreads();
    for (int i = threadIdx.x; i < nk * (DH / 4); i += QT) {
      const int j = i / (DH / 4), d4 = i % (DH / 4);
      const float* base = qkv + (int64_t)(off + kt + j) * (3 * D) + head * DH + d4 * 4;
      *reinterpret_cast<float4*>(&Ks[j][d4 * 4]) = *reinterpret_cast<const float4*>(base + D);
      *reinterpret_cast<float4*>(&Vs[j][d4 * 4]) = *reinterpret_cast<const float4*>(base + 2 * D);
    }
    __syncthreads();
    if (active) {
      float mt = -INFINITY;
      for (int j = 0; j < nk; ++j) {
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < DH; ++d) s = fmaf(q[d], Ks[j][d], s);
        mt = fmaxf(mt, s);
      }
      const float mnew = fmaxf(mrun, mt);
      const float sc = __expf(mrun - mnew);   // exp(-inf) == 0 on the first tile
      lrun *= sc;
#pragma unroll
      for (int d = 0; d < DH; ++d) acc[d] *= sc;
      for (int j = 0; j < nk; ++j) {
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < DH; ++d) s = fmaf(q[d], Ks[j][d], s);
        const float pj = expf(s - mnew);
        lrun += pj;
#pragma unroll
        for (int d = 0; d < DH; ++d) acc[d] = fmaf(pj, Vs[j][d], acc[d]);
      }
      mrun = mnew;
    }
  }
  if (active) {
    const float inv = 1.f / lrun;
    float4* op = reinterpret_cast<float4*>(ctx + (int64_t)(off + qi) * D + head * DH);
#pragma unroll
    for (int d4 = 0; d4 < DH / 4; ++d4)
      op[d4] = make_float4(acc[d4 * 4] * inv, acc[d4 * 4 + 1] * inv, acc[d4 * 4 + 2] * inv, acc[d4 * 4 + 3] * inv);
  }
}

int launch_attention_f32(ResepHandle* h, const float* qkv, float* ctx, int n_seq, int seq_len, const int* seq_off,
                         const int* tile_seq, const int* tile_q0, int n_tiles, cudaStream_t st) {
  ProfScope prof_scope(h, "k_attention_f32", st);
  if (tile_seq == nullptr) {
    if (n_seq == 0) return RESEP_OK;
    if (seq_len <= 160) {
      k_attention_f32<160><<<dim3((unsigned)n_seq, NH), 160, 0, st>>>(qkv, ctx, seq_len, nullptr, nullptr, nullptr);
    } else {
      const int tps = (seq_len + 127) / 128;
      k_attention_f32<128><<<dim3((unsigned)(n_seq * tps), NH), 128, 0, st>>>(qkv, ctx, seq_len, nullptr, nullptr, nullptr);
    }
  } else {
    if (n_tiles == 0) return RESEP_OK;
    k_attention_f32<128><<<dim3((unsigned)n_tiles, NH), 128, 0, st>>>(qkv, ctx, 0, seq_off, tile_seq, tile_q0);
  }
  RESEP_LAUNCH_CHECK(h, "k_attention_f32");
  return RESEP_OK;
}

// ------------------------------------------------------------------------------------------
// Block epilogue, one CTA per sequence: y = LayerNorm(o) (TransformerEncoder.norm) written back
// to o; gLN statistics over all len*128 values of the sequence (fp64 accumulation of sum and
// sum of squares); out = gln_w * (y - mu) * rstd + gln_b + xin; optional mean over the rows
// (the chunk summary `output.mean(1)`).  `out` may alias `xin`.
template <int NWARP>
__global__ void __launch_bounds__(32 * NWARP) k_block_epilogue(float* o, const float* __restrict__ fn_w,
                                                        const float* __restrict__ fn_b, const float* __restrict__ gln_w,
                                                        const float* __restrict__ gln_b, const float* xin, float* out,
                                                        float* __restrict__ seq_mean, int seq_len,
                                                        const int* __restrict__ seq_off) {
  __shared__ double red[2][NWARP];
  __shared__ float stat[2];
  __shared__ __align__(16) float colsum[NWARP][D];
  const int seq = blockIdx.x;
  int off, len;
  if (seq_off != nullptr) {
    off = seq_off[seq];
    len = seq_off[seq + 1] - off;
  } else {
    off = seq * seq_len;
    len = seq_len;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float4 w4 = reinterpret_cast<const float4*>(fn_w)[lane], b4 = reinterpret_cast<const float4*>(fn_b)[lane];
  // rows warp, warp + NWARP, ...: RB of them in flight at a time (one row at a time the single CTA of the coupled
  // memory sequence spent 22 us on 432 rows, all of it load latency)
  constexpr int RB = NWARP >= 32 ? 4 : 8;            // (64 registers per thread at 1,024 threads)
  double s1 = 0.0, s2 = 0.0;
  for (int rb = warp; rb < len; rb += NWARP * RB) {
    float4 xr[RB];
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const int r = rb + i * NWARP;
      if (r < len) xr[i] = reinterpret_cast<const float4*>(o + (int64_t)(off + r) * D)[lane];
    }
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const int r = rb + i * NWARP;
      if (r < len) {
        const float4 y = ln_row(xr[i], w4, b4);
        reinterpret_cast<float4*>(o + (int64_t)(off + r) * D)[lane] = y;
        s1 += (double)y.x + (double)y.y + (double)y.z + (double)y.w;
        s2 += (double)y.x * y.x + (double)y.y * y.y + (double)y.z * y.z + (double)y.w * y.w;
      }
    }
  }
  s1 = warp_sum_d(s1);
  s2 = warp_sum_d(s2);
  if (lane == 0) { red[0][warp] = s1; red[1][warp] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < NWARP; ++i) { a += red[0][i]; b += red[1][i]; }
    const double n = (double)len * D;
    const double mu = a / n;
    double var = b / n - mu * mu;
    if (var < 0.0) var = 0.0;
    stat[0] = (float)mu;
    stat[1] = (float)(1.0 / sqrt(var + (double)GLN_EPS));
  }
  __syncthreads();
  const float mu = stat[0], rstd = stat[1];
  const float4 g4 = reinterpret_cast<const float4*>(gln_w)[lane], h4 = reinterpret_cast<const float4*>(gln_b)[lane];
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int rb = warp; rb < len; rb += NWARP * RB) {
    float4 yr[RB], xr[RB];
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const int r = rb + i * NWARP;
      if (r < len) {
        const int64_t idx = (int64_t)(off + r) * (D / 4) + lane;
        yr[i] = reinterpret_cast<const float4*>(o)[idx];
        xr[i] = reinterpret_cast<const float4*>(xin)[idx];
      }
    }
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const int r = rb + i * NWARP;
      if (r >= len) continue;
      const int64_t idx = (int64_t)(off + r) * (D / 4) + lane;
      const float4 y = yr[i], x = xr[i];
      float4 v;
      v.x = g4.x * (y.x - mu) * rstd + h4.x + x.x;
      v.y = g4.y * (y.y - mu) * rstd + h4.y + x.y;
      v.z = g4.z * (y.z - mu) * rstd + h4.z + x.z;
      v.w = g4.w * (y.w - mu) * rstd + h4.w + x.w;
      reinterpret_cast<float4*>(out)[idx] = v;
      cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w;
    }
  }
  if (seq_mean != nullptr) {
    *reinterpret_cast<float4*>(&colsum[warp][lane * 4]) = cs;
    __syncthreads();
    if (threadIdx.x < D) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < NWARP; ++i) t += colsum[i][threadIdx.x];
      seq_mean[(int64_t)seq * D + threadIdx.x] = t / (float)len;
    }
  }
}

// Same, for equal-length sequences of at most 150 rows (every intra-chunk block): the LayerNorm'ed chunk stays in
// shared memory between the statistics pass and the normalisation pass (76.8 KB, two CTAs per SM), so the kernel
// moves 3 x 512 B per row (o in, xin in, out) instead of 5 x.  Optionally also writes PReLU(out) in bf16 -- the A
// operand of the output_fc GEMM (resepformer.py `output_fc = Sequential(PReLU(), Conv1d(.., 1))`) -- saving the
// separate PReLU kernel's pass over the activations.
constexpr int EPI_T = 512;
__global__ void __launch_bounds__(EPI_T, 2) k_block_epilogue_chunk(const float* __restrict__ o, const float* __restrict__ fn_w,
                                                                   const float* __restrict__ fn_b, const float* __restrict__ gln_w,
                                                                   const float* __restrict__ gln_b, const float* xin, float* out,
                                                                   float* __restrict__ seq_mean, int seq_len,
                                                                   bf16* __restrict__ prelu_out, const float* __restrict__ prelu_a, int f16) {
  extern __shared__ __align__(16) float ys[];            // [seq_len][128]
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // a PDL successor (k_maskdec_tc) may set up meanwhile
  __shared__ double red[2][EPI_T / 32];
  __shared__ float stat[2];
  __shared__ __align__(16) float colsum[EPI_T / 32][D];
  constexpr int NW_ = EPI_T / 32;
  const int seq = blockIdx.x;
  const int64_t off = (int64_t)seq * seq_len;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float4 w4 = reinterpret_cast<const float4*>(fn_w)[lane], b4 = reinterpret_cast<const float4*>(fn_b)[lane];
  // Each warp owns rows warp, warp + 16, ...: up to RPW of them.  All their loads are issued before the first one is
  // used (a row at a time the kernel was latency-bound at 15 us per CTA), and the skip-connection rows are fetched
  // while the chunk statistics are being reduced.
  constexpr int RPW = (CHUNK + NW_ - 1) / NW_;
  double s1 = 0.0, s2 = 0.0;
  float4 xr[RPW];
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const int r = warp + i * NW_;
    if (r < seq_len) xr[i] = reinterpret_cast<const float4*>(o + (off + r) * D)[lane];
  }
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const int r = warp + i * NW_;
    if (r < seq_len) {
      const float4 y = ln_row(xr[i], w4, b4);
      reinterpret_cast<float4*>(ys + r * D)[lane] = y;
      s1 += (double)y.x + (double)y.y + (double)y.z + (double)y.w;
      s2 += (double)y.x * y.x + (double)y.y * y.y + (double)y.z * y.z + (double)y.w * y.w;
    }
  }
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const int r = warp + i * NW_;
    if (r < seq_len) xr[i] = reinterpret_cast<const float4*>(xin)[(off + r) * (D / 4) + lane];
  }
  s1 = warp_sum_d(s1);
  s2 = warp_sum_d(s2);
  if (lane == 0) { red[0][warp] = s1; red[1][warp] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < NW_; ++i) { a += red[0][i]; b += red[1][i]; }
    const double n = (double)seq_len * D;
    const double mu = a / n;
    double var = b / n - mu * mu;
    if (var < 0.0) var = 0.0;
    stat[0] = (float)mu;
    stat[1] = (float)(1.0 / sqrt(var + (double)GLN_EPS));
  }
  __syncthreads();
  const float mu = stat[0], rstd = stat[1];
  const float4 g4 = reinterpret_cast<const float4*>(gln_w)[lane], h4 = reinterpret_cast<const float4*>(gln_b)[lane];
  const float slope = prelu_out != nullptr ? prelu_a[0] : 0.f;
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const int r = warp + i * NW_;
    if (r >= seq_len) continue;
    const int64_t idx = (off + r) * (D / 4) + lane;
    const float4 y = reinterpret_cast<const float4*>(ys + r * D)[lane];
    const float4 x = xr[i];
    float4 v;
    v.x = g4.x * (y.x - mu) * rstd + h4.x + x.x;
    v.y = g4.y * (y.y - mu) * rstd + h4.y + x.y;
    v.z = g4.z * (y.z - mu) * rstd + h4.z + x.z;
    v.w = g4.w * (y.w - mu) * rstd + h4.w + x.w;
    reinterpret_cast<float4*>(out)[idx] = v;
    cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w;
    if (prelu_out != nullptr) {
      float4 pz;
      pz.x = v.x >= 0.f ? v.x : slope * v.x; pz.y = v.y >= 0.f ? v.y : slope * v.y;
      pz.z = v.z >= 0.f ? v.z : slope * v.z; pz.w = v.w >= 0.f ? v.w : slope * v.w;
      if (f16) store4(reinterpret_cast<__half*>(prelu_out) + idx * 4, pz);   // (fp16 mode: the 16-bit container holds IEEE halves)
      else store4(prelu_out + idx * 4, pz);
    }
  }
  if (seq_mean != nullptr) {
    *reinterpret_cast<float4*>(&colsum[warp][lane * 4]) = cs;
    __syncthreads();
    if (threadIdx.x < D) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < NW_; ++i) t += colsum[i][threadIdx.x];
      seq_mean[(int64_t)seq * D + threadIdx.x] = t / (float)seq_len;
    }
  }
}

int launch_block_epilogue(ResepHandle* h, float* o, const float* fn_w, const float* fn_b, const float* gln_w,
                          const float* gln_b, const float* xin, float* out, float* seq_mean, int n_seq, int seq_len,
                          const int* seq_off, cudaStream_t st, bf16* prelu_out, const float* prelu_a) {
  if (n_seq == 0) return RESEP_OK;
  ProfScope prof_scope_451(h, "k_block_epilogue", st);
  if (seq_off == nullptr && seq_len <= CHUNK) {
    const size_t smem = (size_t)seq_len * D * sizeof(float);
    RESEP_CUDA(h, cudaFuncSetAttribute(k_block_epilogue_chunk, cudaFuncAttributeMaxDynamicSharedMemorySize, CHUNK * D * (int)sizeof(float)));
    k_block_epilogue_chunk<<<(unsigned)n_seq, EPI_T, smem, st>>>(o, fn_w, fn_b, gln_w, gln_b, xin, out, seq_mean, seq_len, prelu_out, prelu_a, h->fmt16);
    RESEP_LAUNCH_CHECK(h, "k_block_epilogue_chunk");
    return RESEP_OK;
  }
  // few long sequences (the coupled memory transformer is ONE sequence of all chunk summaries): 32 warps per CTA
  if (n_seq < 2 * h->sm_count)
    k_block_epilogue<32><<<(unsigned)n_seq, 1024, 0, st>>>(o, fn_w, fn_b, gln_w, gln_b, xin, out, seq_mean, seq_len, seq_off);
  else
    k_block_epilogue<8><<<(unsigned)n_seq, 256, 0, st>>>(o, fn_w, fn_b, gln_w, gln_b, xin, out, seq_mean, seq_len, seq_off);
  RESEP_LAUNCH_CHECK(h, "k_block_epilogue");
  if (prelu_out != nullptr) {   // long sequences: PReLU as its own pass
    int64_t rows = 0;
    if (seq_off == nullptr) rows = (int64_t)n_seq * seq_len;
    else return set_err(h, RESEP_EINVAL, "block epilogue: fused PReLU needs equal-length sequences");
    if (h->fmt16) return launch_prelu_t<__half>(h, out, prelu_a, reinterpret_cast<__half*>(prelu_out), rows * D, st);
    return launch_prelu_t<bf16>(h, out, prelu_a, prelu_out, rows * D, st);
  }
  return RESEP_OK;
}

// ------------------------------------------------------------------------------------------
// PReLU with one shared slope (output_fc.0), fp32 or bf16 output for the mask GEMM operand.
template <typename OutT>
__global__ void k_prelu(const float* __restrict__ x, const float* __restrict__ a, OutT* __restrict__ y, int64_t n4) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float s = a[0];
  float4 v = reinterpret_cast<const float4*>(x)[i];
  v.x = v.x >= 0.f ? v.x : s * v.x;
  v.y = v.y >= 0.f ? v.y : s * v.y;
  v.z = v.z >= 0.f ? v.z : s * v.z;
  v.w = v.w >= 0.f ? v.w : s * v.w;
  store4(y + i * 4, v);
}

template <typename OutT>
int launch_prelu_t(ResepHandle* h, const float* x, const float* a, OutT* y, int64_t n, cudaStream_t st);
template <typename OutT>
int launch_prelu_t(ResepHandle* h, const float* x, const float* a, OutT* y, int64_t n, cudaStream_t st) {
  int64_t n4 = n / 4;
  if (n4 == 0) return RESEP_OK;
  ProfScope prof_scope_475(h, "k_prelu", st);
  k_prelu<OutT><<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(x, a, y, n4);
  RESEP_LAUNCH_CHECK(h, "k_prelu");
  return RESEP_OK;
}
template int launch_prelu_t<float>(ResepHandle*, const float*, const float*, float*, int64_t, cudaStream_t);
template int launch_prelu_t<bf16>(ResepHandle*, const float*, const float*, bf16*, int64_t, cudaStream_t);
template int launch_prelu_t<__half>(ResepHandle*, const float*, const float*, __half*, int64_t, cudaStream_t);
int launch_prelu(ResepHandle* h, const float* x, const float* a, float* y, int64_t n, cudaStream_t st) {
  return launch_prelu_t<float>(h, x, a, y, n, st);
}

// ------------------------------------------------------------------------------------------
// Mask apply + decoder + pad/crop.  One CTA = 32 output "slots" (8 samples each) of one item.
// Slot q receives taps 0..7 of frame q and taps 8..15 of frame q-1 (ConvTranspose1d k16 s8), so
// the CTA stages h[l][s][n] = mask[row_l][2n+s] * x0[row_l][n] for frames q0-1 .. q0+31 in shared
// memory, forms the 33 x 2 x 16 frame products against dec_w, and writes est[t][0..1] as float2.
// Samples past T_est get no contribution and come out as exact zeros (upstream's F.pad).
constexpr int DEC_SLOTS = 31;                       // 32 frames -> 64 (frame, speaker) rows x 4 tap quads = 256 threads
constexpr int DEC_ROWS = (DEC_SLOTS + 1) * NSPK;   // 64 (frame, speaker) rows of the frame-product GEMM
constexpr int DEC_HP = DEC_ROWS + 1;               // pitch of the transposed h tile (odd: conflict-free column writes)
__global__ void __launch_bounds__(256) k_decoder(const float* __restrict__ mask, const float* __restrict__ x0,
                                                 const float* __restrict__ dec_w, const int64_t* __restrict__ item_off,
                                                 const int64_t* __restrict__ item_len, const int* __restrict__ item_L,
                                                 const int* __restrict__ item_row0, const int* __restrict__ tile_item,
                                                 const int* __restrict__ tile_slot0, float* __restrict__ est) {
  // fr[(j, s)][k] = sum_n h[(j, s)][n] * dec_w[n][k] is a [66 x 128] . [128 x 16] product.  h is staged transposed
  // (hs[n][row]) so that the 8 rows a warp works on are adjacent words, and each thread owns a 1 x 4 block of fr:
  // one LDS.32 + one LDS.128 per 4 FMAs.
  extern __shared__ __align__(16) float dsm[];
  float* hs = dsm;                                   // [128][DEC_HP]
  float* ws = hs + D * DEC_HP;                       // [128][16]
  float* fr = ws + D * KSZ;                          // [66][16]
  const int item = tile_item[blockIdx.x];
  const int q0 = tile_slot0[blockIdx.x];
  const int L = item_L[item];
  const int64_t T = item_len[item];
  const int64_t row0 = item_row0[item];
  for (int i = threadIdx.x; i < D * KSZ / 4; i += 256) reinterpret_cast<float4*>(ws)[i] = reinterpret_cast<const float4*>(dec_w)[i];
  {  // 32 frames x 128 filters = 16 (frame, filter) pairs per thread: all loads issued before any is used
    const int n = threadIdx.x & (D - 1), jb = threadIdx.x >> 7;       // frames jb, jb + 2, ..., jb + 30
    float xv[16];
    float2 mk[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int l = q0 - 1 + jb + 2 * u;
      const bool ok = l >= 0 && l < L;
      xv[u] = ok ? __ldg(x0 + (row0 + l) * D + n) : 0.f;
      mk[u] = ok ? __ldg(reinterpret_cast<const float2*>(mask + (row0 + l) * (NSPK * D) + 2 * n)) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int j = jb + 2 * u;
      hs[n * DEC_HP + 2 * j] = mk[u].x * xv[u];
      hs[n * DEC_HP + 2 * j + 1] = mk[u].y * xv[u];
    }
  }
  __syncthreads();
  {
    const int kq = threadIdx.x & 3, row = threadIdx.x >> 2;          // 64 rows x 4 tap quads
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int n = 0; n < D; ++n) {
      const float hv = hs[n * DEC_HP + row];
      const float4 w4 = *reinterpret_cast<const float4*>(ws + n * KSZ + 4 * kq);
      acc.x = fmaf(hv, w4.x, acc.x); acc.y = fmaf(hv, w4.y, acc.y);
      acc.z = fmaf(hv, w4.z, acc.z); acc.w = fmaf(hv, w4.w, acc.w);
    }
    *reinterpret_cast<float4*>(fr + row * KSZ + 4 * kq) = acc;
  }
  __syncthreads();
  float* dst = est + 2 * item_off[item];
  for (int i = threadIdx.x; i < DEC_SLOTS * STRIDE; i += 256) {
    const int qs = i / STRIDE, r = i % STRIDE;
    const int64_t t = (int64_t)(q0 + qs) * STRIDE + r;
    if (t >= T) continue;
    // slot q0+qs: frame (q0+qs) is local index qs+1, frame (q0+qs-1) is local index qs
    float2 v;
    v.x = fr[((qs + 1) * NSPK + 0) * KSZ + r] + fr[(qs * NSPK + 0) * KSZ + r + STRIDE];
    v.y = fr[((qs + 1) * NSPK + 1) * KSZ + r] + fr[(qs * NSPK + 1) * KSZ + r + STRIDE];
    *reinterpret_cast<float2*>(dst + 2 * t) = v;
  }
}

int launch_decoder(ResepHandle* h, const float* mask, const float* x0, const Plan& p, float* est, cudaStream_t st) {
  if (p.n_dec_tiles == 0) return RESEP_OK;
  const size_t smem = (D * DEC_HP + D * KSZ + DEC_ROWS * KSZ) * sizeof(float);
  static_assert((D * DEC_HP + D * KSZ + DEC_ROWS * KSZ) * sizeof(float) <= 48 * 1024, "decoder smem");
  ProfScope prof_scope_547(h, "k_decoder", st);
  k_decoder<<<(unsigned)p.n_dec_tiles, 256, smem, st>>>(mask, x0, h->w.dec_w, p.d_item_off, p.d_item_len, p.d_item_L,
                                                       p.d_item_row0, p.d_dec_tile_item, p.d_dec_tile_slot0, est);
  RESEP_LAUNCH_CHECK(h, "k_decoder");
  return RESEP_OK;
}

// ------------------------------------------------------------------------------------------
// Per-source peak normalisation s / (max|s| + 1e-8), the first thing the caller does with every separated source
// (api.py:1082): two passes over est on the device instead of a device->host round trip per source.
constexpr int PEAK_TILE = 4096;   // samples per CTA
__global__ void __launch_bounds__(256) k_peak_max(const float* __restrict__ est, const int64_t* __restrict__ item_off,
                                                  const int64_t* __restrict__ item_len, unsigned* __restrict__ peaks) {
  const int item = blockIdx.y;
  const int64_t T = item_len[item];
  const int64_t t0 = (int64_t)blockIdx.x * PEAK_TILE;
  if (t0 >= T) return;
  const float2* src = reinterpret_cast<const float2*>(est) + item_off[item];
  const int64_t t1 = t0 + PEAK_TILE < T ? t0 + PEAK_TILE : T;
  float m0 = 0.f, m1 = 0.f;
  for (int64_t t = t0 + threadIdx.x; t < t1; t += 256) {
    const float2 v = src[t];
    m0 = fmaxf(m0, fabsf(v.x));
    m1 = fmaxf(m1, fabsf(v.y));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, o));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
  }
  // non-negative floats order like their bit patterns
  if ((threadIdx.x & 31) == 0) {
    atomicMax(peaks + 2 * item, __float_as_uint(m0));
    atomicMax(peaks + 2 * item + 1, __float_as_uint(m1));
  }
}

__global__ void __launch_bounds__(256) k_peak_scale(float* __restrict__ est, const int64_t* __restrict__ item_off,
                                                    const int64_t* __restrict__ item_len, const float* __restrict__ peaks) {
  const int item = blockIdx.y;
  const int64_t T = item_len[item];
  const int64_t t0 = (int64_t)blockIdx.x * PEAK_TILE;
  if (t0 >= T) return;
  float2* dst = reinterpret_cast<float2*>(est) + item_off[item];
  const int64_t t1 = t0 + PEAK_TILE < T ? t0 + PEAK_TILE : T;
  const float d0 = peaks[2 * item] + 1e-8f, d1 = peaks[2 * item + 1] + 1e-8f;
  for (int64_t t = t0 + threadIdx.x; t < t1; t += 256) {
    float2 v = dst[t];
    v.x = v.x / d0;                                   // IEEE division: bit-identical to the caller's torch expression
    v.y = v.y / d1;
    dst[t] = v;
  }
}

int launch_peak_normalize(ResepHandle* h, float* est, const Plan& p, int64_t max_len, float* peaks, cudaStream_t st) {
  if (p.B == 0 || max_len <= 0) return RESEP_OK;
  RESEP_CUDA(h, cudaMemsetAsync(peaks, 0, sizeof(float) * 2 * p.B, st));
  const dim3 grid((unsigned)((max_len + PEAK_TILE - 1) / PEAK_TILE), (unsigned)p.B);
  {
    ProfScope prof_scope(h, "k_peak_max", st);
    k_peak_max<<<grid, 256, 0, st>>>(est, p.d_item_off, p.d_item_len, reinterpret_cast<unsigned*>(peaks));
    RESEP_LAUNCH_CHECK(h, "k_peak_max");
  }
  ProfScope prof_scope(h, "k_peak_scale", st);
  k_peak_scale<<<grid, 256, 0, st>>>(est, p.d_item_off, p.d_item_len, peaks);
  RESEP_LAUNCH_CHECK(h, "k_peak_scale");
  return RESEP_OK;
}

// ------------------------------------------------------------------------------------------
// Polyphase FIR resampler (windowed-sinc taps computed by the host layer: torchaudio.functional.resample's
// definition).  The product feeds 16 kHz audio to a separator trained at 8 kHz (api.py:115 vs the model's
// sample_rate 8000) without resampling; these two passes let a caller run the model at its native rate and get the
// sources back at the file's rate without leaving the device.
//   y[row][m * up + p][c] = sum_j taps[p][j] * xpad[row][m * down + j][c],  xpad[i] = x[i - width] (0 outside)
// x: [rows][n_in][ch], y: [rows][n_out][ch] (ch = 1 for mixtures, 2 for separated sources); taps in shared memory.
__global__ void __launch_bounds__(256) k_resample_fir(const float* __restrict__ x, int64_t n_in, float* __restrict__ y, int64_t n_out,
                                                      int ch, int down, int up, const float* __restrict__ taps, int ktaps, int width) {
  extern __shared__ float s_taps[];
  const bool in_smem = (size_t)up * ktaps * sizeof(float) <= 48 * 1024;   // else (e.g. 44.1 kHz -> 8 kHz: 80 x 509) from L1/L2
  if (in_smem) {
    for (int i = threadIdx.x; i < up * ktaps; i += 256) s_taps[i] = taps[i];
    __syncthreads();
  }
  const int row = blockIdx.y;
  const float* xr = x + (int64_t)row * n_in * ch;
  float* yr = y + (int64_t)row * n_out * ch;
  const int64_t total = n_out * ch;
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (int64_t)gridDim.x * 256) {
    const int64_t o = e / ch;
    const int c = (int)(e - o * ch);
    const int64_t m = o / up;
    const int p = (int)(o - m * up);
    const int64_t i0 = m * down - width;
    const float* tp = (in_smem ? s_taps : taps) + p * ktaps;
    float acc = 0.f;
    for (int j = 0; j < ktaps; ++j) {
      const int64_t i = i0 + j;
      if (i >= 0 && i < n_in) acc = fmaf(tp[j], xr[i * ch + c], acc);
    }
    yr[e] = acc;
  }
}

int launch_resample_fir(ResepHandle* h, const float* x, int rows, int64_t n_in, float* y, int64_t n_out, int ch, int down, int up,
                        const float* taps, int ktaps, int width, cudaStream_t st) {
  if (rows <= 0 || n_out <= 0) return RESEP_OK;
  const size_t tap_bytes = (size_t)up * ktaps * sizeof(float);
  const size_t smem = tap_bytes <= 48 * 1024 ? tap_bytes : 0;
  int64_t blocks = (n_out * ch + 255) / 256;
  if (blocks > 4096) blocks = 4096;
  ProfScope prof_scope(h, "k_resample_fir", st);
  k_resample_fir<<<dim3((unsigned)blocks, (unsigned)rows), 256, smem, st>>>(x, n_in, y, n_out, ch, down, up,
                                                                                                       taps, ktaps, width);
  RESEP_LAUNCH_CHECK(h, "k_resample_fir");
  return RESEP_OK;
}

}  // namespace resep
