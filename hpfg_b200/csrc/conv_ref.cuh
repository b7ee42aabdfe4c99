// CUDA-core direct convolution kernels (fprop / dgrad / wgrad) with fp32 FMA accumulation.
// They serve (a) the whole fp32 check path (HPFG_PREC_FP32, the 1e-5 parity bar) and (b) the two layers of
// the bf16 path whose GEMM N or K is too narrow for tcgen05 (in_conv.0 with K = 9*Cin <= 27 and out_conv with
// N = num_classes <= 8), reading/writing the API's NCHW fp32 tensors directly through strided views.
#pragma once
#include "common.cuh"
#include "glue.cuh"

namespace hpfg {

// 4-D tensor view with element strides (NHWC contiguous: sc = 1; NCHW: sw = 1).
struct TView {
    void *p;
    int64_t sn, sh, sw, sc;
};
inline TView nhwc_view(void *p, int H, int W, int C) {
    return TView{p, (int64_t)H * W * C, (int64_t)W * C, (int64_t)C, 1};
}
inline TView nchw_view(void *p, int C, int H, int W) {
    return TView{p, (int64_t)C * H * W, (int64_t)W, 1, (int64_t)H * W};
}

// optional per-input-channel transform applied while loading the A operand:
//   v = dropout(leaky(v*scale[c] + shift[c]))            (BatchNorm + LeakyReLU + Dropout of the producer)
struct LoadXform {
    const float *scale, *shift;   // nullptr -> identity
    DropSpec drop;                // bits indexed by the NHWC element index of the INPUT tensor
};

// out[n,y,x,co] = sum_{r,s,ci} xform(in)[n,y+r-pad,x+s-pad,ci] * wpk[(r*KS+s)][ci][co]  (+ bias[co])
// stats (optional): per pixel-tile partial sums [tiles][2*Cout] of the bias-free output (sum | sum sq).
template <typename TI, typename TO>
int conv_ref_fprop(TView in, TView out, const float *wpk, const float *bias, int N, int H, int W, int Cin, int Cout,
                   int KS, LoadXform xf, float *stats, cudaStream_t s);
int conv_ref_num_tiles(int N, int H, int W);

// dwpk[(r*KS+s)][ci][co] = sum_{n,y,x} xform(in)[n,y+r-pad,x+s-pad,ci] * dout[n,y,x,co];  dbias[co] = sum dout.
// Written through split partials (scratch) and a fixed-order reduction straight into the flat OIHW gradient.
template <typename TI, typename TD>
int conv_ref_wgrad(TView in, TView dout, int N, int H, int W, int Cin, int Cout, int KS, LoadXform xf, float *scratch,
                   int64_t scratch_floats, float *dw_oihw, float *dbias, int accumulate, cudaStream_t s);
int64_t conv_ref_wgrad_scratch_floats(int N, int H, int W, int Cin, int Cout, int KS);

// weight packing: OIHW fp32 -> fprop [tap][ci][co] and dgrad [tap'][co][ci] (tap' = flipped tap), fp32
int pack_weights_ref(const float *w_oihw, float *wpk_fprop, float *wpk_dgrad, int Cin, int Cout, int KS, cudaStream_t s);

}  // namespace hpfg
