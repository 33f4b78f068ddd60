/*
 * resep_b200.h -- C ABI of libresep_b200.so: the RE-SepFormer overlap-separation forward
 * pass as hand-written sm_100a CUDA kernels.
 *
 * The reference (Yotsuei/ClearConverse) has no native code and no FFI: its hot path is the
 * Python call  `separated = self.separator.separate_batch(subsegment)`
 * (/root/reference/back/api.py:1077) on an object built by
 * `SepformerSeparation.from_hparams(...)` (api.py:713-717) whose arithmetic lives in the
 * un-vendored dependency `speechbrain`.  This header is the boundary a maintainer binds
 * (ctypes) to replace that object; each entry point names the
 * reference / upstream interface it stands in for.
 *
 * Conventions: every function returns 0 on success or a negative RESEP_E* code; no C++
 * exception crosses the ABI; the message for the last failure on a handle is
 * resep_last_error(h) (resep_last_error(NULL) for failures of resep_create itself).
 * A handle is thread-compatible (one caller at a time), bound to one CUDA device, and stays
 * usable after an error.  All device work is stream-ordered on the `stream` argument
 * (a cudaStream_t passed as void*); the caller owns every buffer it passes in.
 * There is no CPU fallback: without a CUDA device resep_create fails.
 */
#ifndef RESEP_B200_H_
#define RESEP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RESEP_ABI_VERSION 1

/* status codes */
#define RESEP_OK            0
#define RESEP_EINVAL       -1   /* bad argument: null pointer, B<=0, unsupported config ...          */
#define RESEP_ESHORT       -2   /* an item is shorter than the encoder kernel (T < 16): upstream's
                                   Conv1d raises; api.py:1107 turns it into "[Processing error]"   */
#define RESEP_ECUDA        -3   /* CUDA runtime / driver error (message has the cudaError string)   */
#define RESEP_EWORKSPACE   -4   /* workspace smaller than resep_workspace_bytes() asked for         */
#define RESEP_EPOS         -5   /* a sequence is longer than the positional-encoding table given    */
#define RESEP_ENODEVICE    -6   /* no CUDA device / not an sm_100 device                            */

/* arithmetic of the GEMMs and attention (everything else -- residual stream, LayerNorm, gLN,
 * softmax, encoder, decoder -- is fp32 in every mode) */
#define RESEP_PREC_FP32     0   /* fp32 FMA kernels (no tensor cores): reference-grade parity path  */
#define RESEP_PREC_TF32     1   /* tcgen05 kind::tf32, fp32 accumulate in TMEM                      */
#define RESEP_PREC_BF16     2   /* tcgen05 kind::f16 (bf16 operands), fp32 accumulate in TMEM       */
#define RESEP_PREC_FP16     3   /* the same fused kernels with IEEE fp16 operands (11-bit significand
                                   = tf32's precision), every weight as fp16 hi + lo: the fast
                                   fp32-tolerance mode (max-abs <= 1e-3)                            */

/* batch semantics of the memory (inter-chunk) transformer, SURVEY.md section 8e */
#define RESEP_BATCH_COUPLED      0  /* upstream's literal B>1 behaviour: chunks of all items form ONE
                                       sequence (resepformer.py `hc.unsqueeze(0)`)                  */
#define RESEP_BATCH_INDEPENDENT  1  /* per item == looping B=1 calls, what api.py:1073-1077 does    */

/* Architecture of speechbrain/resepformer-wsj02mix (hyperparams.yaml).  The kernels are
 * specialised for exactly these values; resep_create rejects anything else. */
typedef struct ResepConfig {
  int32_t abi_version;   /* RESEP_ABI_VERSION */
  int32_t n_filters;     /* 128  encoder out_channels == d_model            */
  int32_t kernel_size;   /* 16   encoder / decoder kernel                   */
  int32_t stride;        /* 8                                               */
  int32_t segment_size;  /* 150  chunk length K                             */
  int32_t n_heads;       /* 8                                               */
  int32_t d_ffn;         /* 1024                                            */
  int32_t n_layers;      /* 8    transformer layers per block               */
  int32_t n_blocks;      /* 2    masknet `layer` (n_blocks-1 memory blocks) */
  int32_t n_spks;        /* 2                                               */
} ResepConfig;

/* One pre-norm TransformerEncoderLayer (speechbrain Transformer.py).  HOST pointers, fp32,
 * PyTorch layouts (Linear weight = [out_features, in_features], row-major). */
typedef struct ResepLayerWeights {
  const float* norm1_w;     /* [128]        norm1.norm.weight                */
  const float* norm1_b;     /* [128]                                         */
  const float* in_proj_w;   /* [384,128]    self_att.att.in_proj_weight      */
  const float* in_proj_b;   /* [384]                                         */
  const float* out_proj_w;  /* [128,128]    self_att.att.out_proj.weight     */
  const float* out_proj_b;  /* [128]                                         */
  const float* norm2_w;     /* [128]        norm2.norm.weight                */
  const float* norm2_b;     /* [128]                                         */
  const float* ffn1_w;      /* [1024,128]   pos_ffn.ffn.0.weight             */
  const float* ffn1_b;      /* [1024]                                        */
  const float* ffn2_w;      /* [128,1024]   pos_ffn.ffn.3.weight             */
  const float* ffn2_b;      /* [128]                                         */
} ResepLayerWeights;

/* One SBTransformerBlock_wnormandskip (speechbrain resepformer.py). */
typedef struct ResepBlockWeights {
  ResepLayerWeights layers[8];
  const float* final_norm_w;  /* [128]  mdl.norm.norm.weight                 */
  const float* final_norm_b;  /* [128]                                       */
  const float* gln_w;         /* [128]  norm.weight ([128,1] upstream)       */
  const float* gln_b;         /* [128]  norm.bias                            */
} ResepBlockWeights;

/* Everything encoder.ckpt / masknet.ckpt / decoder.ckpt hold (api.py:729). */
typedef struct ResepWeights {
  const float* enc_w;         /* [128,16]   encoder conv1d.weight [128,1,16]            */
  const float* dec_w;         /* [128,16]   decoder weight [128,1,16]                   */
  const float* prelu_a;       /* [1]        model.output_fc.0.weight                    */
  const float* fc_w;          /* [256,128]  model.output_fc.1.weight [256,128,1]        */
  const float* fc_b;          /* [256]                                                  */
  const float* pe;            /* [pe_rows,128] pos_enc.pe rows 0..pe_rows-1 (a stored buffer
                                 upstream; identical for all three blocks)              */
  int64_t      pe_rows;       /* >= 150; a memory sequence longer than this -> RESEP_EPOS */
  ResepBlockWeights seg[2];   /* model.seg_model.{0,1}                                  */
  ResepBlockWeights mem[1];   /* model.mem_model.0                                      */
} ResepWeights;

typedef struct ResepHandle ResepHandle;

/* Replaces SepformerSeparation.from_hparams(source, savedir, run_opts={"device"})
 * (api.py:713-717): copies the weights to `device`, builds the bf16 / packed copies. */
int resep_create(const ResepConfig* cfg, const ResepWeights* w, int device, ResepHandle** out);

/* Re-uploads weights into an existing handle (a corrected form of api.py:738-745). */
int resep_load_weights(ResepHandle* h, const ResepWeights* w);

int resep_destroy(ResepHandle* h);

const char* resep_last_error(const ResepHandle* h);

/* Bytes of device scratch resep_forward needs for B items of `item_len[i]` samples. */
int resep_workspace_bytes(ResepHandle* h, int B, const int64_t* item_len, int precision, size_t* bytes);

/* Replaces SepformerSeparation.separate_batch(mix) (api.py:1077; upstream
 * speechbrain/inference/separation.py).
 *   mix          DEVICE fp32; item i occupies mix[item_off[i] .. item_off[i]+item_len[i])
 *   item_off/len HOST int64[B]        (a dense [B,T] batch is off[i]=i*T, len[i]=T)
 *   est          DEVICE fp32; item i, sample t, speaker s at est[2*item_off[i] + 2*t + s]
 *                (a dense batch is exactly upstream's contiguous [B,T,2])
 * Trailing samples past T_est = 8*(L-1)+16 are written as exact zeros (upstream's F.pad). */
int resep_forward(ResepHandle* h, const float* mix, const int64_t* item_off, const int64_t* item_len, int B,
                  float* est, void* workspace, size_t workspace_bytes, int precision, int batch_mode,
                  void* stream);

/* One long recording split by chunks across several GPUs (optional row): the intra blocks are per chunk,
 * so rank r runs them on its contiguous chunk range; the only exchange is the chunk summaries.
 *   phase 1  encoder + seg_model[0] on the span            -> chunk_means [n,128]   (all-gather these, in span order)
 *   resep_memory_block on the FULL sequence of summaries    -> hc [S,128]           (every rank, redundantly)
 *   phase 2  seg_model[1](out + hc rows of the span) + mask + decoder -> est [span_len, 2]
 * A span that ends inside the recording (`inner` = 1) must hold a whole number of chunks plus the decoder's 8-sample
 * tail: span_len = 1200 * n + 8; it gets no padding chunk, and its last 8 output samples are the upper filter taps of
 * its last frame, to be ADDED to the first 8 samples of the next span's output (overlap-add across the cut).  The last
 * span (`inner` = 0) is an ordinary item.  Both phases must be given the same workspace (resep_workspace_bytes for one
 * item of span_len samples); results are bit-identical to the unsplit resep_forward. */
typedef struct ResepSpanCtl {
  int phase;              /* 1 or 2                                                        */
  int inner;              /* 1: the span ends on a chunk boundary inside the recording     */
  float* chunk_means;     /* phase 1 out: DEVICE fp32 [n_chunks of the span, 128]          */
  const float* hc;        /* phase 2 in:  DEVICE fp32 [n_chunks of the span, 128]          */
} ResepSpanCtl;
int resep_forward_span(ResepHandle* h, const float* mix, int64_t span_len, float* est, void* workspace,
                       size_t workspace_bytes, int precision, void* stream, const ResepSpanCtl* ctl);
/* mem_model[0] on one sequence of chunk summaries: chunk_means, hc DEVICE fp32 [n_chunks,128] (may alias). */
int resep_memory_workspace_bytes(ResepHandle* h, int n_chunks, size_t* bytes);
int resep_memory_block(ResepHandle* h, const float* chunk_means, float* hc, int n_chunks, void* workspace,
                       size_t workspace_bytes, int precision, void* stream);

/* Polyphase FIR resampling on the device, for running the 8 kHz separator on the product's 16 kHz audio
 * (api.py:115 feeds 16 kHz to a model whose sample_rate is 8000).  Same arithmetic as
 * torchaudio.functional.resample: x zero-padded by `width` on the left,
 *   y[row][m*up + p][c] = sum_j taps[p*ktaps + j] * xpad[row][m*down + j][c]
 *   x  DEVICE fp32 [rows][n_in][channels]    y  DEVICE fp32 [rows][n_out][channels]
 *   taps DEVICE fp32 [up][ktaps] (clearconverse_b200.separation.resample_taps builds torchaudio's windowed sinc) */
int resep_resample_fir(ResepHandle* h, const float* x, int rows, int64_t n_in, float* y, int64_t n_out, int channels,
                       int down, int up, const float* taps, int ktaps, int width, void* stream);

/* Replaces the caller's first step on every separated source, api.py:1082
 * `source / (source.abs().max() + 1e-8)`, on the device and for all items at once:
 *   est[item][t][spk] /= max_t |est[item][t][spk]| + 1e-8   in place, layout as resep_forward's est
 *   peaks        DEVICE fp32[2*B], receives the maxima (item-major, speaker-minor)
 * IEEE division, so the result equals the torch expression bit for bit.  NaN samples are ignored by the maximum. */
int resep_peak_normalize(ResepHandle* h, float* est, const int64_t* item_off, const int64_t* item_len, int B,
                         float* peaks, void* stream);

/* Same as resep_forward, additionally copying intermediates for per-kernel parity tests.
 * Any pointer may be NULL.  Token-major fp32, chunk-padded: M = 150 * sum_i S_i rows of 128. */
typedef struct ResepDebugOut {
  float* enc;         /* [M,128]  encoder features, chunk-padded (pad rows are zero)      */
  float* seg0;        /* [M,128]  output of seg_model[0]                                  */
  float* chunk_mean;  /* [M/150,128] chunk summaries fed to mem_model[0]                  */
  float* mem0;        /* [M/150,128] output of mem_model[0]                               */
  float* seg1;        /* [M,128]  output of seg_model[1]                                  */
} ResepDebugOut;

int resep_forward_debug(ResepHandle* h, const float* mix, const int64_t* item_off, const int64_t* item_len,
                        int B, float* est, void* workspace, size_t workspace_bytes, int precision,
                        int batch_mode, void* stream, const ResepDebugOut* dbg);

/* Kernels launched by this handle since creation (bench.py's "gpu_launches" evidence). */
int64_t resep_launch_count(const ResepHandle* h);

/* Per-kernel device timing for bench.py's roofline object: resep_profile(h,1) makes every kernel
 * launch of this handle be bracketed by CUDA events on its launch stream; resep_profile_report
 * synchronises, sums the event durations per kernel name, writes a JSON object
 * {"kernel": {"ms": total_ms, "launches": n}, ...} into buf (NUL-terminated, truncated to cap)
 * and switches profiling off.  Not for use inside a timed region (event records perturb it). */
int resep_profile(ResepHandle* h, int enable);
int resep_profile_report(ResepHandle* h, char* buf, size_t cap);

/* ---- per-kernel entry points (unit parity tests; DEVICE pointers, stream-ordered) ------- */

/* dual_path.Encoder: relu(conv1d k16 s8) of ONE item -> token-major [L,128]. */
int resep_encoder_fwd(ResepHandle* h, const float* mix, int64_t T, float* tokens_out, void* stream);

/* One TransformerEncoderLayer over `n_seq` equal-length sequences of `seq_len` rows:
 * x [n_seq*seq_len,128] fp32 updated in place. block: 0,1 = seg_model[i]; 2 = mem_model[0]. */
int resep_layer_fwd(ResepHandle* h, int block, int layer, float* x, int n_seq, int seq_len,
                    void* workspace, size_t workspace_bytes, int precision, void* stream);

/* Measurement aid: launches ONE of the three fused kernels of that layer (bf16 / fp16 modes) `reps` times back to
 * back on the buffers a preceding resep_layer_fwd call with the same arguments left in `workspace`, so that a pair of
 * events around the call times the kernel itself, without per-launch event records or host gaps.
 * which: 0 = LayerNorm + in-projection (k_qkv2_tc), 1 = attention, 2 = out-proj + LayerNorm + FFN (k_post2_tc).
 * pdl: 0 = every launch waits for its predecessor to drain (a launch's duration includes its own set-up and tail),
 * 1 = programmatic dependent launch as in the product path.  x is overwritten; the values are not meaningful. */
int resep_layer_kernel_repeat(ResepHandle* h, int block, int layer, int which, float* x, int n_seq, int seq_len,
                              void* workspace, size_t workspace_bytes, int precision, int reps, int pdl, void* stream);

/* out[M,N] = A[M,K] . W[N,K]^T + bias, through the GEMM kernel of the given precision
 * (W, bias: DEVICE fp32).  For testing the tcgen05 path against the fp32 path. */
int resep_linear_fwd(ResepHandle* h, const float* A, const float* W, const float* bias, float* out,
                     int64_t M, int N, int K, int relu, int precision, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* RESEP_B200_H_ */
