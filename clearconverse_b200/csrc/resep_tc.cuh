// Tensor-core (tcgen05 / TMEM / TMA) side of libresep_b200.so; implemented in kernels_tc.cu.
#pragma once
#include "resep_internal.cuh"

namespace resep {

// One TransformerEncoderLayer on o [rows,128] (fp32 residual stream, updated in place) with the
// GEMMs and attention on tensor cores.  y/qkv/ctx/hid are scratch regions of the workspace.
int tc_repeat_layer_kernel(ResepHandle* h, const LayerDev& lw, int which, float* o, int64_t rows, int n_seq, int seq_len,
                           float* qkv, float* ctx, int reps, bool pdl, cudaStream_t st, bool intra);
int tc_run_layer(ResepHandle* h, const LayerDev& lw, float* o, int64_t rows, int n_seq, int seq_len, const int* seq_off,
                 const int* tile_seq, const int* tile_q0, int n_tiles, int max_seq_len, float* y, float* qkv, float* ctx, float* hid,
                 int precision, cudaStream_t st, bool intra = false);

// Intra-chunk attention on tcgen05 / TMEM (kernels_attn.cu), bf16 mode: ctx[n_chunks * 150, 128] from the head-interleaved
// bf16 in-projection buffer.  The default for the intra blocks; RESEP_ATTN_TC=0 selects the mma.sync kernel.
int launch_attn_tc(ResepHandle* h, const bf16* qkv, bf16* ctx, int n_chunks, cudaStream_t st);
// whether the intra blocks take k_attn_tc unless RESEP_ATTN_TC says otherwise (the faster of the two measured kernels)
bool attn_tc_default();

// Fused post-attention half of a layer (kernels_layer.cu), bf16 mode:
//   o <- o' + W2 relu(W1 LN2(o') + b1) + b2,  o' = o + ctx . Wo^T + bo      (o fp32 in place, ctx bf16)
int launch_post_tc(ResepHandle* h, const LayerDev& lw, const bf16* ctx, float* o, int64_t rows, cudaStream_t st);

// The same on CTA pairs (tcgen05 cta_group::2, kernels_post2.cu): the default; RESEP_POST2=0 selects the 1-CTA kernel
int launch_post2_tc(ResepHandle* h, const LayerDev& lw, const bf16* ctx, float* o, int64_t rows, cudaStream_t st);

// Fused norm1 + in-projection (kernels_layer.cu), bf16 mode: qkv[rows,384] = LN1(o) . Win^T + bin
int launch_qkv_tc(ResepHandle* h, const LayerDev& lw, const float* o, bf16* qkv, int64_t rows, cudaStream_t st);

// The same on CTA pairs with resident weights (kernels_qkv2.cu): the default; RESEP_QKV2=0 selects the 1-CTA kernel
int launch_qkv2_tc(ResepHandle* h, const LayerDev& lw, const float* o, bf16* qkv, int64_t rows, cudaStream_t st);

extern long long* g_attn_trace;   // development aid: clock trace buffer of k_attn_tc (null unless RESEP_TRACE is set)
extern long long* g_post_trace;   // development aid: clock trace buffer of k_post_tc (null unless RESEP_TRACE is set)

// output_fc: mask[M,256] = relu(prelu(a) . fc_w^T + fc_b)   (fp32 out)
// prelu_done: y_scratch already holds PReLU(a) in bf16 (written by the block epilogue)
int tc_run_mask(ResepHandle* h, const float* a, float* y_scratch, float* mask, int64_t M, int precision, cudaStream_t st,
                bool prelu_done = false);

// out = A . W^T + bias for arbitrary DEVICE fp32 W (test hook for the GEMM kernel)
int tc_linear_test(ResepHandle* h, const float* A, const float* W, const float* bias, float* out, int64_t M, int N, int K,
                   bool relu, int precision, cudaStream_t st);

}  // namespace resep
