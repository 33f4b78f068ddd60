// Shared by the tensor-core translation units: PTX wrappers + TMA tensor-map construction.
#pragma once
#include <cuda.h>

#include <cstdlib>
#include <string>

#include "ptx_sm100.cuh"
#include "resep_tc.cuh"

namespace resep {

// driver entry point for cuTensorMapEncodeTiled (resolved at run time: no link dependency on libcuda)
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
extern PFN_encodeTiled g_encode;

// 2-D row-major [rows, cols] tensor, box = [box_rows, 128 bytes of columns], 128B swizzle, OOB reads as 0.
template <typename T>
static int make_tmap(ResepHandle* h, CUtensorMap* m, const T* base, int64_t rows, int cols, int box_rows) {
  const CUtensorMapDataType dt = sizeof(T) == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)cols * sizeof(T)};
  cuuint32_t box[2] = {(cuuint32_t)(128 / sizeof(T)), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, dt, 2, const_cast<T*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(h, RESEP_ECUDA, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
  return RESEP_OK;
}

// 2-D row-major [rows, cols] bf16 tensor, box = [box_rows, 16 columns = 32 bytes], 32B swizzle, OOB reads as 0:
// one (sequence, head) slice of q, k or v out of the packed [rows, 384] in-projection output.
static int make_tmap_head(ResepHandle* h, CUtensorMap* m, const bf16* base, int64_t rows, int cols, int box_rows) {
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)cols * sizeof(bf16)};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};   // q | k | v of one head (48 columns) + 16 unused
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(base), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(h, RESEP_ECUDA, "cuTensorMapEncodeTiled (head slice) failed: " + std::to_string((int)r));
  return RESEP_OK;
}

// Launch with programmatic dependent launch: the kernel may start while its predecessor in the stream drains; it
// calls ptx::pdl_wait() before touching anything the predecessor produced.  RESEP_PDL=0 launches it serialised.
extern bool g_serial_launches;   // profiling aid (resep_layer_kernel_repeat): launch without the PDL attribute
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  static const bool pdl = !(getenv("RESEP_PDL") && getenv("RESEP_PDL")[0] == '0');
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = pdl && !g_serial_launches ? 1 : 0;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

}  // namespace resep
