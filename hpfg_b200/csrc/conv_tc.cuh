// Tensor-core (tcgen05 / TMEM / TMA) implicit-GEMM convolution path for bf16 plans: interface used by the
// plan's schedules.  Every entry point sets *done = false (and returns HPFG_OK) when the layer shape is not
// handled by the tensor-core kernels, in which case the caller runs the CUDA-core kernel instead.
#pragma once
#include "unet_plan.cuh"

namespace hpfg {

int tc_plan_init(hpfg_unet_plan *p);
void tc_plan_free(hpfg_unet_plan *p);
// bf16 weight packing for all tensor-core layers (one pass per optimiser step)
int tc_pack_all(hpfg_unet_plan *p, const float *params, cudaStream_t s, bool need_dgrad = true);
// 3x3 fprop: in (bf16 NHWC, transformed on load by xf) -> out (bf16 NHWC, bias-free); stats partials [*P][2*Cout]
int tc_fprop(hpfg_unet_plan *p, int conv, const void *in, void *out, LoadXform xf, float *stats, int *P, bool *done,
             cudaStream_t s);
int tc_fprop_1x1(hpfg_unet_plan *p, int conv, const void *in, void *out, LoadXform xf, const float *bias, bool *done,
                 cudaStream_t s);
// out_conv: bf16 NHWC in (transformed on load) -> fp32 NCHW logits (+bias), Cout padded to 16 inside the kernel
int tc_fprop_logits(hpfg_unet_plan *p, int conv, const void *in, LoadXform xf, const float *bias, float *logits_nchw, cudaStream_t s);
int tc_dgrad(hpfg_unet_plan *p, int conv, const void *dout, void *din, bool *done, cudaStream_t s);
int tc_wgrad(hpfg_unet_plan *p, int conv, const void *in, LoadXform xf, const void *dout, float *dw_oihw, float *dbias,
             int accumulate, bool *done, cudaStream_t s);

// tensor-core wgrad (wgrad_tc.cu)
int64_t tc_wgrad_scratch_floats(int N, int H, int W, int Cin, int Cout, int KS);
int tc_wgrad_run(int ks, int N, int H, int W, int Cin, int Cout, int cin_real, int cout_real, const void *x, LoadXform xf, const void *dy, float *scratch,
                 int64_t scratch_floats, float *dw_oihw, float *dbias, int accumulate, cudaStream_t s);

// fp32 NCHW [N,C,H,W] -> bf16 NHWC with the channel count padded to 16 (zeros): network input and dlogits
int pad_to_nhwc16(const float *src_nchw, void *dst_bf16_nhwc16, int N, int C, int H, int W, cudaStream_t s);

}  // namespace hpfg
