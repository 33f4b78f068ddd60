// Tensor-core (tcgen05 / TMEM / TMA) side of libresep_b200.so; implemented in kernels_tc.cu.
#pragma once
#include "resep_internal.cuh"

namespace resep {

// One TransformerEncoderLayer on o [rows,128] (fp32 residual stream, updated in place) with the
// GEMMs and attention on tensor cores.  y/qkv/ctx/hid are scratch regions of the workspace.
int tc_run_layer(ResepHandle* h, const LayerDev& lw, float* o, int64_t rows, int n_seq, int seq_len, const int* seq_off,
                 const int* tile_seq, const int* tile_q0, int n_tiles, float* y, float* qkv, float* ctx, float* hid,
                 int precision, cudaStream_t st);

// output_fc: mask[M,256] = relu(prelu(a) . fc_w^T + fc_b)   (fp32 out)
int tc_run_mask(ResepHandle* h, const float* a, float* y_scratch, float* mask, int64_t M, int precision, cudaStream_t st);

// out = A . W^T + bias for arbitrary DEVICE fp32 W (test hook for the GEMM kernel)
int tc_linear_test(ResepHandle* h, const float* A, const float* W, const float* bias, float* out, int64_t M, int N, int K,
                   bool relu, int precision, cudaStream_t st);

}  // namespace resep
