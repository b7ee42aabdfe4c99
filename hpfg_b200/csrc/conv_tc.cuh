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
// Optional BatchNorm-backward fusion of a data-gradient / weight-gradient launch (bf16 plans):
//   in_raw  != nullptr: the gradient operand is built on load from TWO tensors, draw = sc*g + kb*raw + kd (g = the tensor passed
//                       as dout, raw = the saved conv output of the BatchNorm layer the gradient enters through);
//   out_raw != nullptr (dgrad only): the output is the gradient wrt the ACTIVATED output of a BatchNorm layer; the epilogue
//                       stores g = dact * leaky'(bn(raw)) * dropout' instead and writes the partial sums (sum g | sum g*raw).
struct TcBwdFuse {
    const void *in_raw = nullptr;
    const float *sc = nullptr, *kb = nullptr, *kd = nullptr;
    const void *out_raw = nullptr;
    const float *gs_scale = nullptr, *gs_shift = nullptr;
    const uint32_t *gs_dropbits = nullptr;
    float gs_inv_keep = 1.f;
};
int tc_dgrad(hpfg_unet_plan *p, int conv, const void *dout, void *din, bool *done, cudaStream_t s, const TcBwdFuse *fuse = nullptr,
             float *stats = nullptr, int *P = nullptr);
// Optional stream split of a weight gradient: partials into `scratch`, `ev_partials` recorded behind the tensor-core kernel on its
// stream, the fixed-order reduction on `reduce_stream` (waits for ev_partials, records ev_reduced).  Default: everything on one stream.
struct TcWgradStreams {
    float *scratch = nullptr;
    cudaStream_t reduce_stream = nullptr;
    cudaEvent_t ev_partials = nullptr, ev_reduced = nullptr;
};
int tc_wgrad(hpfg_unet_plan *p, int conv, const void *in, LoadXform xf, const void *dout, float *dw_oihw, float *dbias,
             int accumulate, bool *done, cudaStream_t s, const TcBwdFuse *fuse = nullptr, const TcWgradStreams *ws = nullptr);

// tensor-core wgrad (wgrad_tc.cu)
int64_t tc_wgrad_scratch_floats(int N, int H, int W, int Cin, int Cout, int KS);
int tc_wgrad_run(int ks, int N, int H, int W, int Cin, int Cout, int cin_real, int cout_real, const void *x, LoadXform xf, const void *dy, float *scratch,
                 int64_t scratch_floats, float *dw_oihw, float *dbias, int accumulate, cudaStream_t s, const TcBwdFuse *fuse = nullptr,
                 const TcWgradStreams *ws = nullptr);

// fp32 NCHW [N,C,H,W] -> bf16 NHWC with the channel count padded to 16 (zeros): network input and dlogits
int pad_to_nhwc16(const float *src_nchw, void *dst_bf16_nhwc16, int N, int C, int H, int W, cudaStream_t s);

}  // namespace hpfg
