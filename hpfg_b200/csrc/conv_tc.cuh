// Tensor-core (tcgen05 / TMEM / TMA) implicit-GEMM convolution path for bf16 plans: interface used by the
// plan's schedules.  Every entry point sets *done = false (and returns HPFG_OK) when the layer shape is not
// handled by the tensor-core kernels, in which case the caller runs the CUDA-core kernel instead.
#pragma once
#include "unet_plan.cuh"

namespace hpfg {

int tc_plan_init(hpfg_unet_plan *p);
void tc_plan_free(hpfg_unet_plan *p);
// bf16 weight packing for all tensor-core layers (one pass per optimiser step)
int tc_pack_all(hpfg_unet_plan *p, const float *params, cudaStream_t s);
// 3x3 fprop: in (bf16 NHWC, transformed on load by xf) -> out (bf16 NHWC, bias-free); stats partials [*P][2*Cout]
int tc_fprop(hpfg_unet_plan *p, int conv, const void *in, void *out, LoadXform xf, float *stats, int *P, bool *done,
             cudaStream_t s);
int tc_fprop_1x1(hpfg_unet_plan *p, int conv, const void *in, void *out, LoadXform xf, const float *bias, bool *done,
                 cudaStream_t s);
int tc_dgrad(hpfg_unet_plan *p, int conv, const void *dout, void *din, bool *done, cudaStream_t s);
int tc_wgrad(hpfg_unet_plan *p, int conv, const void *in, LoadXform xf, const void *dout, float *dw_oihw, float *dbias,
             int accumulate, bool *done, cudaStream_t s);

}  // namespace hpfg
