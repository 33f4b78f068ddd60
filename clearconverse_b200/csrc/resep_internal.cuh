// Internal declarations shared by the translation units of libresep_b200.so.
// Architecture constants are those of speechbrain/resepformer-wsj02mix.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "resep_b200.h"

namespace resep {

constexpr int D = 128;       // d_model == encoder filters
constexpr int KSZ = 16;      // encoder / decoder kernel
constexpr int STRIDE = 8;
constexpr int CHUNK = 150;   // segment_size K
// internal batch mode of resep_forward_span's inner spans: one item, no padding chunk when L % 150 == 0
constexpr int RESEP_BATCH_SPAN_EXACT = 2;
constexpr int NH = 8;
constexpr int DH = 16;
// bf16 qkv buffer: row = [head 0: q16 k16 v16 | head 1: ... ] (the fp32 / tf32 paths keep torch's [q128 | k128 | v128])
constexpr int QKV_HEAD_STRIDE = 3 * DH, QKV_K = DH, QKV_V = 2 * DH;
// ... and its k slice is pre-multiplied (in the fp32 in-projection weights / bias, before the bf16 split) by
// 1/sqrt(DH) * log2(e): q . k is then the softmax exponent in base 2 and the attention kernels need no scaling FFMA
constexpr float QK_PRESCALE = 0.25f * 1.4426950408889634f;
constexpr int FFN = 1024;
constexpr int NL = 8;
constexpr int NSPK = 2;
constexpr float LN_EPS = 1e-6f;
constexpr float GLN_EPS = 1.1920928955078125e-07f;  // torch.finfo(float32).eps

typedef __nv_bfloat16 bf16;

struct LayerDev {
  const float *norm1_w, *norm1_b, *in_w, *in_b, *out_w, *out_b, *norm2_w, *norm2_b, *f1_w, *f1_b, *f2_w, *f2_b;
  const bf16 *in_w_bf, *out_w_bf, *f1_w_bf, *f2_w_bf;       // bf16 copies (RESEP_PREC_BF16); in_w_bf / in_w_bl rows head-interleaved
  const float* in_b_hi;                                     // in_b in the same head-interleaved order
  const bf16 *in_w_bl, *out_w_bl, *f1_w_bl, *f2_w_bl;       // bf16(W - bf16(W)): low part for the split-weight mode
  // the same four weights as IEEE fp16 hi + lo (RESEP_PREC_FP16; stored behind the 16-bit container type `bf16`):
  // [0] = fp16(W), [1] = fp16(W - fp16(W)); in_w rows head-interleaved and key rows pre-scaled like in_w_bf
  const bf16 *in_w_h[2], *out_w_h[2], *f1_w_h[2], *f2_w_h[2];
  const float *in_w_tf, *out_w_tf, *f1_w_tf, *f2_w_tf;      // tf32-rounded fp32 copies (RESEP_PREC_TF32): hi part
  const float *in_w_lo, *out_w_lo, *f1_w_lo, *f2_w_lo;      // tf32(W - hi): the TF32 mode runs W = hi + lo
  // HOST copy of out_b[128], norm2_w[128], norm2_b[128], f2_b[128], f1_b[1024]: k_post2_tc takes them as kernel
  // parameters (constant bank) so that its epilogues do not spend shared-memory bandwidth on broadcast loads
  const float* h_post_par;
  const float* h_in_b;      // HOST copy of in_b_hi[384], norm1_w[128], norm1_b[128] (k_qkv2_tc kernel parameters)
  // k_post2_tc's FFN1 with norm2's scale and shift folded in: W1g[h][k] = W1[h][k] * g2[k], b1g = b1 + W1 . be2, so
  // that its LayerNorm epilogue only normalises (no parameter loads on the tile-boundary critical path).
  // bf16 hi / lo, fp16 hi / lo, and the HOST copy of b1g[1024]
  const bf16 *f1g_w_bf, *f1g_w_bl, *f1g_w_h[2];
  const float* h_b1g;
  const float* post_par;    // DEVICE: out_b[128] | f2_b[128] | b1g[1024], one bulk copy into k_post2_tc's shared memory
};
struct BlockDev {
  LayerDev layers[NL];
  const float *fn_w, *fn_b, *gln_w, *gln_b;
};
struct WeightsDev {
  const float *enc_w, *dec_w, *prelu_a, *fc_w, *fc_b, *pe;
  const bf16 *fc_w_bf, *fc_w_bl;
  const bf16* fc_w_h[2];    // fp16 hi / lo
  const float *fc_w_tf, *fc_w_lo;
  int64_t pe_rows;
  BlockDev blk[3];  // 0 = seg_model[0], 1 = seg_model[1], 2 = mem_model[0]
};

// Per-call shape tables ("plan"), cached on the handle by shape signature.
struct Plan {
  std::vector<int64_t> key;
  int B = 0;
  int64_t n_chunks = 0;   // sum_i S_i
  int64_t M = 0;          // 150 * n_chunks intra tokens
  int n_mem_seq = 0;      // 1 (coupled) or B (independent)
  int max_mem_len = 0;
  int n_mem_tiles = 0;    // attention query tiles over the memory sequences
  int n_dec_tiles = 0;
  // k_maskdec_tc (fused output_fc + decoder) needs a token row for every output slot of every item (8 * 150 * S >= T:
  // false only when L % 150 == 149 and T % 8 != 0) and sample offsets that fit an int
  bool maskdec_ok = false;
  // one device block holding every table below, and the pinned host block it is uploaded from (both come from the
  // handle's size-class pools and go back there on eviction: no cudaMalloc / cudaFree / stream sync per new shape)
  void* dev = nullptr;
  void* host = nullptr;
  size_t dev_bytes = 0;     // bytes used
  size_t cap_bytes = 0;     // size class of the blocks
  const int64_t* d_item_off = nullptr;    // [B] sample offset of item in mix
  const int64_t* d_item_len = nullptr;    // [B] T
  const int* d_item_L = nullptr;          // [B] frames
  const int* d_item_row0 = nullptr;       // [B] first token row of the item (150 * first chunk)
  const int* d_chunk_item = nullptr;      // [n_chunks]
  const int* d_chunk_frame0 = nullptr;    // [n_chunks] frame index of the chunk's first row
  const int* d_mem_pos = nullptr;         // [n_chunks] position of the chunk in its memory sequence
  const int* d_mem_seq_off = nullptr;     // [n_mem_seq+1]
  const int* d_mem_tile_seq = nullptr;    // [n_mem_tiles]
  const int* d_mem_tile_q0 = nullptr;     // [n_mem_tiles]
  const int* d_dec_tile_item = nullptr;   // [n_dec_tiles]
  const int* d_dec_tile_slot0 = nullptr;  // [n_dec_tiles]
  uint64_t last_use = 0;
};

}  // namespace resep

struct ResepHandle {
  int device = 0;
  int sm_count = 148;
  std::string err;
  void* arena = nullptr;  // all weights, one allocation
  std::vector<float> host_par;   // per layer: the 1,536 floats LayerDev::h_post_par points at + 384 for h_in_b
  size_t arena_bytes = 0;
  resep::WeightsDev w;
  std::vector<resep::Plan*> plans;
  uint64_t tick = 0;
  int64_t launches = 0;
  // RESEP_PREC_BF16 weight operands (DESIGN.md "precision modes"): 2 = in-proj / out-proj / output_fc as bf16 hi + lo
  // (two MMAs per K-slice), FFN weights as one rounded bf16 (default); 1 = every weight hi + lo; 0 = bf16(W) only
  int w16_mode = 2;
  // what the 16-bit kernels of the CURRENT call compute in: 0 = bf16 (RESEP_PREC_BF16), 1 = fp16 (RESEP_PREC_FP16, every
  // weight hi + lo).  Set by the forward entry points from the precision argument.
  int fmt16 = 0;
  // the fp16 mode's setting.  Default 1: every weight as fp16 hi + lo.  Measured worst max-abs over four weight seeds x four
  // shapes / config-2 throughput: this 7.3e-4 / 33.5 k; FFN1 single, the rest hi + lo (RESEP_W16F=ffn2) 7.5e-4 / 36.5 k -- but 1.09e-3
  // at the full config-2 batch, outside the 1e-3 contract; FFN2 single (ffn1) 9.4e-4 / 36.2 k; both FFN single (mixed) 8.7e-4 / 39.9 k
  int w16_mode_fp16 = 1;
  int w16_mode_bf16 = 2;    // the bf16 mode's weight-operand setting (RESEP_W16), restored when a bf16 call follows an fp16 one
  bool prof_on = false;   // resep_profile(): bracket every launch with CUDA events
  struct ProfRec { cudaEvent_t a, b; const char* name; };
  std::vector<ProfRec> prof;
  // CUDA graphs of whole forward passes, keyed by (plan, buffers, mode): the ~80 launches of a forward cost more
  // host time than the kernels take on the device at small batch sizes (api.py calls with B = 1)
  struct GraphRec {
    std::vector<uint64_t> key;
    cudaGraphExec_t exec = nullptr;
    int64_t launches = 0;
    uint64_t last_use = 0;
    bool seen_only = true;   // first sighting runs eagerly (plans, attributes and occupancy queries are set up outside capture)
    const resep::Plan* plan = nullptr;   // the captured kernels hold pointers into this plan's device tables
  };
  std::vector<GraphRec> graphs;
  cudaStream_t cap_stream = nullptr;   // capture happens here (the caller's stream may be the legacy default stream)
  int use_graphs = 1;                  // RESEP_GRAPH=0 disables
  int graph_failures = 0;              // capture / instantiate failures of THIS handle; three switch its graphs off
  struct PoolBlock { void* dev; void* host; size_t cap; };
  std::vector<PoolBlock> plan_pool;    // free (device, pinned host) block pairs of evicted plans, reused by size class
  bool tc_ready = false;  // tensor maps for the tcgen05 path built
  void* tc_state = nullptr;
};

namespace resep {

// ---------------------------------------------------------------- error plumbing
int set_err(ResepHandle* h, int code, const std::string& msg);
#define RESEP_CUDA(h, expr)                                                                    \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return ::resep::set_err(h, RESEP_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
  } while (0)
// RAII: in profiling mode records an event before and after the launch(es) issued during its lifetime.
struct ProfScope {
  ResepHandle* h;
  cudaStream_t st;
  int idx = -1;
  ProfScope(ResepHandle* h_, const char* name, cudaStream_t st_) : h(h_), st(st_) {
    if (!h->prof_on) return;
    ResepHandle::ProfRec r{nullptr, nullptr, name};
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    cudaEventRecord(r.a, st);
    h->prof.push_back(r);
    idx = (int)h->prof.size() - 1;
  }
  ~ProfScope() {
    if (idx >= 0) cudaEventRecord(h->prof[idx].b, st);
  }
};
#define RESEP_LAUNCH_CHECK(h, name)                                                            \
  do {                                                                                         \
    (h)->launches++;                                                                           \
    cudaError_t e__ = cudaGetLastError();                                                      \
    if (e__ != cudaSuccess)                                                                    \
      return ::resep::set_err(h, RESEP_ECUDA, std::string("launch ") + name + ": " + cudaGetErrorString(e__)); \
  } while (0)

// ---------------------------------------------------------------- fp32 kernels (kernels_simt.cu)
// o_fused != null: also writes o = x0 + pe[row in chunk], the prologue of the first transformer block
int launch_encoder_chunked(ResepHandle* h, const float* mix, const Plan& p, float* x0, cudaStream_t st, float* o_fused = nullptr);
int launch_encoder_single(ResepHandle* h, const float* mix, int64_t T, float* tokens, cudaStream_t st);
// o = xin + pe[pos]; if hc != null first xin = xprev + hc[row / rows_per_hc] (written to xin)
int launch_block_prologue(ResepHandle* h, const float* xprev, const float* hc, float* xin, float* o, int64_t rows,
                          const int* pos, int seq_len, cudaStream_t st);
template <typename OutT>
int launch_layernorm(ResepHandle* h, const float* x, const float* w, const float* b, OutT* y, int64_t rows,
                     cudaStream_t st);
// C = A.W^T + bias (relu?) (+ residual, may alias C)
int launch_gemm_f32(ResepHandle* h, const float* A, const float* W, const float* bias, const float* residual, float* C,
                    int64_t M, int N, int K, bool relu, cudaStream_t st);
// generic fp32 attention over sequences; qkv [rows,384] -> ctx [rows,128]
int launch_attention_f32(ResepHandle* h, const float* qkv, float* ctx, int n_seq, int seq_len, const int* seq_off,
                         const int* tile_seq, const int* tile_q0, int n_tiles, cudaStream_t st);
// final LayerNorm + gLN per sequence + skip (+ per-sequence column mean)
// prelu_out (optional, bf16 [rows,128]): additionally PReLU(out; prelu_a) for the output_fc GEMM
int launch_block_epilogue(ResepHandle* h, float* o, const float* fn_w, const float* fn_b, const float* gln_w,
                          const float* gln_b, const float* xin, float* out, float* seq_mean, int n_seq, int seq_len,
                          const int* seq_off, cudaStream_t st, bf16* prelu_out = nullptr, const float* prelu_a = nullptr);
// y = round_to_tf32(LayerNorm(x)); x[i] = round_to_tf32(x[i]) in place
int launch_layernorm_tf32(ResepHandle* h, const float* x, const float* w, const float* b, float* y, int64_t rows,
                          cudaStream_t st);
int launch_round_tf32(ResepHandle* h, float* x, int64_t n, cudaStream_t st);
int launch_prelu(ResepHandle* h, const float* x, const float* a, float* y, int64_t n, cudaStream_t st);
template <typename OutT>
int launch_prelu_t(ResepHandle* h, const float* x, const float* a, OutT* y, int64_t n, cudaStream_t st);
// mask [M,256] (already relu'd), x0 [M,128] -> est
int launch_decoder(ResepHandle* h, const float* mask, const float* x0, const Plan& p, float* est, cudaStream_t st);
int launch_resample_fir(ResepHandle* h, const float* x, int rows, int64_t n_in, float* y, int64_t n_out, int ch, int down, int up,
                        const float* taps, int ktaps, int width, cudaStream_t st);
// est[item][t][spk] /= max_t |est[item][t][spk]| + 1e-8 (api.py:1082); peaks: device float[2 * B], receives the maxima
int launch_peak_normalize(ResepHandle* h, float* est, const Plan& p, int64_t max_len, float* peaks, cudaStream_t st);
// bf16 mode: output_fc + ReLU mask + feature product + decoder in one kernel (kernels_maskdec.cu); needs p.maskdec_ok
int launch_maskdec(ResepHandle* h, const bf16* prelu, const float* x0, const Plan& p, float* est, cudaStream_t st);

// ---------------------------------------------------------------- tensor-core kernels (kernels_tc.cu)
int tc_init(ResepHandle* h);
void tc_destroy(ResepHandle* h);

}  // namespace resep
