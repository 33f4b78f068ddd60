#include "resep_tc.cuh"
namespace resep {
int tc_init(ResepHandle* h) { return set_err(h, RESEP_EINVAL, "tensor-core path not built yet"); }
void tc_destroy(ResepHandle*) {}
int tc_run_layer(ResepHandle* h, const LayerDev&, float*, int64_t, int, int, const int*, const int*, const int*, int, float*, float*, float*, float*, int, cudaStream_t) { return set_err(h, RESEP_EINVAL, "tensor-core path not built yet"); }
int tc_run_mask(ResepHandle* h, const float*, float*, float*, int64_t, int, cudaStream_t) { return set_err(h, RESEP_EINVAL, "tensor-core path not built yet"); }
int tc_linear_test(ResepHandle* h, const float*, const float*, const float*, float*, int64_t, int, int, bool, int, cudaStream_t) { return set_err(h, RESEP_EINVAL, "tensor-core path not built yet"); }
}
