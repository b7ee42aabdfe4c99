// Tensor-core convolution path (placeholder until the tcgen05 kernels land: every layer falls through to the
// CUDA-core kernels, so a bf16 plan already runs end to end with bf16 storage and fp32 accumulation).
#include "conv_tc.cuh"

namespace hpfg {

int tc_plan_init(hpfg_unet_plan *) { return HPFG_OK; }
void tc_plan_free(hpfg_unet_plan *) {}
int tc_pack_all(hpfg_unet_plan *, const float *, cudaStream_t) { return HPFG_OK; }
int tc_fprop(hpfg_unet_plan *, int, const void *, void *, LoadXform, float *, int *, bool *done, cudaStream_t) {
    *done = false;
    return HPFG_OK;
}
int tc_fprop_1x1(hpfg_unet_plan *, int, const void *, void *, LoadXform, const float *, bool *done, cudaStream_t) {
    *done = false;
    return HPFG_OK;
}
int tc_dgrad(hpfg_unet_plan *, int, const void *, void *, bool *done, cudaStream_t) {
    *done = false;
    return HPFG_OK;
}
int tc_wgrad(hpfg_unet_plan *, int, const void *, LoadXform, const void *, float *, float *, int, bool *done, cudaStream_t) {
    *done = false;
    return HPFG_OK;
}

}  // namespace hpfg
