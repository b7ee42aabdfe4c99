// Declarations of the layout-generic "glue" kernels shared by the fp32 check path and the bf16 tensor-core
// path: BatchNorm statistics finalisation / backward, activated max-pool, bilinear-upsample + concat gather
// and their adjoints, dropout bit-mask generation.  All activations are NHWC, T = float|bf16.
#pragma once
#include "common.cuh"

namespace hpfg {

// Per-BatchNorm device state owned by the plan (all fp32, C entries each).
struct BnState {
    float *scale, *shift;      // fused affine: y = x*scale + shift  (scale = gamma*invstd, shift = beta - mean*scale)
    float *mean, *invstd;      // batch statistics of the bias-free conv output (saved for backward)
    float *c1, *c2;            // backward: mean(g), mean(g*xhat)
    float *kb, *kd;            // backward, folded: draw = scale*g + kb*raw + kd  (kb = -scale*c2*invstd, kd = -scale*c1 - kb*mean)
};

struct DropSpec {
    const uint32_t *bits;      // NHWC bit-packed keep mask (bit = element index & 31), or nullptr = no dropout
    float inv_keep;            // 1/(1-p)
};

// stats partials: [P][2*C] floats (sum | sum of squares); finalize -> BnState (+ running stats when training)
int bn_finalize(const float *partials, int P, int C, int64_t count, const float *gamma, const float *beta,
                const float *conv_bias, float *running_mean, float *running_var, int64_t *counter, int training,
                BnState st, cudaStream_t s, double *sync_sums = nullptr);   // sync_sums (4*256 doubles): statistics over all ranks
// eval mode: scale/shift from the running statistics (conv bias folded in)
int bn_eval_affine(int C, const float *gamma, const float *beta, const float *conv_bias, const float *running_mean,
                   const float *running_var, BnState st, cudaStream_t s);

template <typename T>
int pool_act(const T *raw, T *pooled, int N, int H, int W, int C, BnState bn, cudaStream_t s);
template <typename T>
int upcat(const T *raw_skip, BnState bn_skip, const T *low, T *cat, int N, int h, int w, int F, cudaStream_t s);

// BatchNorm (+LeakyReLU +dropout) backward.  dact: grad wrt dropout(leaky(bn(raw))).  draw: grad wrt raw.
template <typename T>
int bn_bwd(const T *dact, const T *raw, T *draw, int64_t M, int C, BnState bn, DropSpec drop, float *partials,
           int max_partials, float *dgamma, float *dbeta, int accumulate, cudaStream_t s, double *sync_sums = nullptr);

// grad wrt an encoder feature = skip half of dcat (+) un-pooled grad of the pooled tensor
template <typename T>
int skip_pool_bwd(const T *dcat, const T *dpooled, const T *raw, BnState bn, T *dact, int N, int H, int W, int F,
                  cudaStream_t s);
// The same, fused with BatchNorm-backward pass 0 of that feature's own BatchNorm: stores g = dact * leaky'(bn(raw)) and
// writes the partial sums (sum g | sum g*raw) as [*P][2*F] rows; bn_bwd_reduce turns them into c1/c2/kb/kd + dgamma/dbeta.
template <typename T>
int skip_pool_bwd_gstat(const T *dcat, const T *dpooled, const T *raw, BnState bn, T *g_out, int N, int H, int W, int F,
                        float *partials, int max_partials, int *P, cudaStream_t s);
template <typename T>
int bn_bwd_from_g(const T *g, const T *raw, T *draw, int64_t M, int C, BnState bn, cudaStream_t s);
// BatchNorm-backward reduction of [P][2*C] partials (sum g | sum g*raw) -> bn.c1/c2/kb/kd, dgamma, dbeta (count = N*H*W)
int bn_bwd_reduce(const float *partials, int P, int C, int64_t count, BnState bn, float *dgamma, float *dbeta, int accumulate,
                  cudaStream_t s);
// adjoint of the bilinear x2 (align_corners) upsample: dcat[..., F:2F] at (2h,2w) -> dlow at (h,w)
template <typename T>
int up_bwd(const T *dcat, T *dlow, int N, int h, int w, int F, cudaStream_t s);

// dropout keep-mask bits (NHWC order) from a user NCHW uint8 mask or from Philox(seed, offset, NCHW index)
int dropout_bits(uint32_t *bits, const uint8_t *mask_nchw, int N, int H, int W, int C, float p, uint64_t seed,
                 uint64_t offset, cudaStream_t s);

// the same for up to 8 tensors in one launch (tensor i uses Philox offset + i)
int dropout_bits_multi(int n_jobs, uint32_t *const *bits, const uint8_t *const *masks_nchw, int N, const int *H, const int *W,
                       const int *C, const float *p, uint64_t seed, uint64_t offset, cudaStream_t s,
                       const uint64_t *offset_dev = nullptr);   // offset_dev: device scalar overriding offset

// NHWC (T) -> NCHW fp32 copy with optional per-channel bias (debug taps)
template <typename T>
int nhwc_to_nchw_f32(const T *src, float *dst, int N, int H, int W, int C, const float *bias, cudaStream_t s);

// activated ConvBlock output leaky(bn(raw)) as fp32 NCHW (feature tap for callers outside the plan), and its adjoint side:
// an external fp32 NCHW gradient added into a T NHWC gradient buffer
template <typename T>
int act_nhwc_to_nchw_f32(const T *raw, BnState bn, float *dst, int N, int H, int W, int C, cudaStream_t s);
template <typename T>
int add_nchw_f32_to_nhwc(T *dst, const float *src, int N, int H, int W, int C, cudaStream_t s);

}  // namespace hpfg
