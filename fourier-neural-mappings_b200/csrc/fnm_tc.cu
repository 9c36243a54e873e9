// tcgen05 / TMA kernels for the heavy spatial passes.  (placeholder: the SIMT path serves every
// shape until these kernels are validated on hardware.)
#include "fnm_stages.cuh"

namespace fnm {

int tc_available() { return 0; }
bool tc_ydft_supported(int64_t, int, int, int) { return false; }
int tc_ydft(const float*, const float*, float*, int64_t, int, int, int, int, int, cudaStream_t) {
    return fail(FNM_ENOTSUP, "tc_ydft: not built");
}
bool tc_pointwise_ysyn_supported(int64_t, int, int, int, int) { return false; }
int tc_pointwise_ysyn(const float*, int, const float*, int, const float*, const float*, const float*, const float*,
                      float*, int64_t, int, int, int, int, int, cudaStream_t) {
    return fail(FNM_ENOTSUP, "tc_pointwise_ysyn: not built");
}

}  // namespace fnm
