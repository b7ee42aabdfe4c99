// Glue kernels (see glue.cuh).  All HBM-bound element-wise / gather / per-channel-reduction work:
// 16-byte vector accesses along the contiguous channel dimension, grids sized from the SM count.
// Reference semantics: nn.BatchNorm2d / nn.LeakyReLU / nn.Dropout / nn.MaxPool2d(2) /
// nn.Upsample(scale_factor=2, bilinear, align_corners=True) / torch.cat as wired in model/unet.py:12-58.
#include "glue.cuh"

#include <algorithm>

namespace hpfg {

template <typename T> struct Vec;
template <> struct Vec<float> {
    static constexpr int N = 4;
    __device__ static void load(const float *p, float (&v)[4]) {
        const float4 r = *reinterpret_cast<const float4 *>(p);
        v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
    }
    __device__ static void store(float *p, const float (&v)[4]) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct Vec<bf16> {
    static constexpr int N = 8;
    __device__ static void load(const bf16 *p, float (&v)[8]) {
        const uint4 r = *reinterpret_cast<const uint4 *>(p);
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    __device__ static void store(bf16 *p, const float (&v)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<const uint32_t *>(&h);
        }
        *reinterpret_cast<uint4 *>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
};

static int ew_grid(int64_t work_items, int threads = 256) {
    int64_t b = (work_items + threads - 1) / threads;
    const int64_t cap = (int64_t)kNumSMs * 16;
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
}

// ------------------------------------------------------------------------------------------ bn_finalize
__global__ void __launch_bounds__(256) bn_finalize_kernel(const float *__restrict__ partials, int P, int C,
                                                          double inv_count, double unbias,
                                                          const float *__restrict__ gamma,
                                                          const float *__restrict__ beta,
                                                          const float *__restrict__ conv_bias, float *running_mean,
                                                          float *running_var, int64_t *counter, int training,
                                                          BnState st) {
    pdl_prologue();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= C) return;
    const int c = warp;
    double s = 0.0, q = 0.0;
    for (int p = lane; p < P; p += 32) {
        s += (double)partials[(int64_t)p * 2 * C + c];
        q += (double)partials[(int64_t)p * 2 * C + C + c];
    }
    s = warp_sum(s);
    q = warp_sum(q);
    if (lane != 0) return;
    const double mean = s * inv_count;
    double var = q * inv_count - mean * mean;
    if (var < 0.0) var = 0.0;
    const double invstd = 1.0 / sqrt(var + (double)kBnEps);
    st.scale[c] = (float)((double)gamma[c] * invstd);
    st.shift[c] = (float)((double)beta[c] - mean * (double)gamma[c] * invstd);
    st.mean[c] = (float)mean;
    st.invstd[c] = (float)invstd;
    if (training) {   // running estimates track the biased conv output (bias is left out of the stored tensor)
        const float b = conv_bias ? conv_bias[c] : 0.f;
        running_mean[c] = (1.f - kBnMomentum) * running_mean[c] + kBnMomentum * ((float)mean + b);
        running_var[c] = (1.f - kBnMomentum) * running_var[c] + kBnMomentum * (float)(var * unbias);
        if (c == 0 && counter) *counter += 1;
    }
}

// ---- exact-global mode: partial rows -> fp64 sums[2C] (one launch), all-reduce of those sums over the ranks (host hook), then the
// same finalisation from the summed values with the global element count.
__global__ void __launch_bounds__(256) bn_sum_partials_kernel(const float *__restrict__ partials, int P, int C2, double *__restrict__ sums,
                                                              double *__restrict__ copy) {
    pdl_prologue();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= C2) return;
    double s = 0.0;
    for (int p = lane; p < P; p += 32) s += (double)partials[(int64_t)p * C2 + warp];
    s = warp_sum(s);
    if (lane == 0) {
        sums[warp] = s;
        if (copy) copy[warp] = s;
    }
}

__global__ void __launch_bounds__(256) bn_finalize_sums_kernel(const double *__restrict__ sums, int C, double inv_count, double unbias,
                                                               const float *__restrict__ gamma, const float *__restrict__ beta,
                                                               const float *__restrict__ conv_bias, float *running_mean,
                                                               float *running_var, int64_t *counter, int training, BnState st) {
    pdl_prologue();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double mean = sums[c] * inv_count;
    double var = sums[C + c] * inv_count - mean * mean;
    if (var < 0.0) var = 0.0;
    const double invstd = 1.0 / sqrt(var + (double)kBnEps);
    st.scale[c] = (float)((double)gamma[c] * invstd);
    st.shift[c] = (float)((double)beta[c] - mean * (double)gamma[c] * invstd);
    st.mean[c] = (float)mean;
    st.invstd[c] = (float)invstd;
    if (training) {
        const float b = conv_bias ? conv_bias[c] : 0.f;
        running_mean[c] = (1.f - kBnMomentum) * running_mean[c] + kBnMomentum * ((float)mean + b);
        running_var[c] = (1.f - kBnMomentum) * running_var[c] + kBnMomentum * (float)(var * unbias);
        if (c == 0 && counter) *counter += 1;
    }
}

// backward: c1 / c2 / kb / kd from the GLOBAL sums (sums_g), dgamma / dbeta from this rank's LOCAL sums (the gradient all-reduce adds
// the ranks' contributions afterwards)
__global__ void __launch_bounds__(256) bn_bwd_finalize_sums_kernel(const double *__restrict__ sums_l, const double *__restrict__ sums_g, int C,
                                                                   double inv_count, BnState bn, float *dgamma, float *dbeta, int accumulate) {
    pdl_prologue();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double invstd = (double)bn.invstd[c], mean = (double)bn.mean[c];
    const double sg = sums_g[c], qg = invstd * (sums_g[C + c] - mean * sg);
    const double sl = sums_l[c], ql = invstd * (sums_l[C + c] - mean * sl);
    bn.c1[c] = (float)(sg * inv_count);
    bn.c2[c] = (float)(qg * inv_count);
    const float sc = bn.scale[c];
    const float kb = -sc * bn.c2[c] * bn.invstd[c];
    bn.kb[c] = kb;
    bn.kd[c] = -sc * bn.c1[c] - kb * bn.mean[c];
    if (accumulate) { dgamma[c] += (float)ql; dbeta[c] += (float)sl; }
    else { dgamma[c] = (float)ql; dbeta[c] = (float)sl; }
}

int bn_finalize(const float *partials, int P, int C, int64_t count, const float *gamma, const float *beta,
                const float *conv_bias, float *running_mean, float *running_var, int64_t *counter, int training,
                BnState st, cudaStream_t s, double *sync_sums) {
    ProfScope _prof(PROF_GLUE, s);
    if (sync_sums && training) {
        const int64_t gcount = count * sync_world();
        const double unbias_g = gcount > 1 ? (double)gcount / (double)(gcount - 1) : 1.0;
        HPFG_CUDA_CHECK(launch_plain(bn_sum_partials_kernel, ceil_div(2 * C * 32, 256), 256, 0, s, partials, P, 2 * C, sync_sums, (double *)nullptr));
        HPFG_LAUNCH_CHECK();
        HPFG_RETURN_IF(sync_allreduce(sync_sums, 2 * C, true, s));
        HPFG_CUDA_CHECK(launch_plain(bn_finalize_sums_kernel, ceil_div(C, 256), 256, 0, s, (const double *)sync_sums, C, 1.0 / (double)gcount, unbias_g,
                                   gamma, beta, conv_bias, running_mean, running_var, counter, training, st));
        HPFG_LAUNCH_CHECK();
        return HPFG_OK;
    }
    const double unbias = count > 1 ? (double)count / (double)(count - 1) : 1.0;
    HPFG_CUDA_CHECK(launch_pdl(bn_finalize_kernel, ceil_div(C * 32, 256), 256, 0, s, partials, P, C, 1.0 / (double)count, unbias, gamma,
                                                             beta, conv_bias, running_mean, running_var, counter,
                                                             training, st));
    HPFG_LAUNCH_CHECK();
    return HPFG_OK;
}

__global__ void bn_eval_affine_kernel(int C, const float *gamma, const float *beta, const float *conv_bias,
                                      const float *rm, const float *rv, BnState st) {
    pdl_prologue();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float invstd = 1.f / sqrtf(rv[c] + kBnEps);
    const float b = conv_bias ? conv_bias[c] : 0.f;
    st.scale[c] = gamma[c] * invstd;
    st.shift[c] = beta[c] - (rm[c] - b) * gamma[c] * invstd;   // stored tensor excludes the conv bias
    st.mean[c] = rm[c] - b;
    st.invstd[c] = invstd;
}

int bn_eval_affine(int C, const float *gamma, const float *beta, const float *conv_bias, const float *running_mean,
                   const float *running_var, BnState st, cudaStream_t s) {
    ProfScope _prof(PROF_GLUE, s);
    HPFG_CUDA_CHECK(launch_pdl(bn_eval_affine_kernel, ceil_div(C, 128), 128, 0, s, C, gamma, beta, conv_bias, running_mean, running_var, st));
    HPFG_LAUNCH_CHECK();
    return HPFG_OK;
}

// --------------------------------------------------------------------------------------------- pool_act
template <typename T>
__global__ void __launch_bounds__(256) pool_act_kernel(const T *__restrict__ raw, T *__restrict__ pooled, int N, int H,
                                                       int W, int C, BnState bn, FastDiv dCV, FastDiv dWo, FastDiv dHo) {
    pdl_prologue();
    constexpr int V = Vec<T>::N;
    const int Ho = H >> 1, Wo = W >> 1;
    const uint32_t total = (uint32_t)N * Ho * Wo * dCV.d;
    // the grid stride (256 * gridDim) is a multiple of CV, so a thread keeps one channel vector: its affine lives in registers
    const uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t cv = i0 - fast_div(i0, dCV) * dCV.d;
    float sc[V], sh[V];
#pragma unroll
    for (int k = 0; k < V; ++k) { sc[k] = bn.scale[cv * V + k]; sh[k] = bn.shift[cv * V + k]; }
    for (uint32_t i = i0; i < total; i += gridDim.x * blockDim.x) {
        uint32_t pix = fast_div(i, dCV), wo, ho, n;
        fast_divmod(pix, dWo, pix, wo);
        fast_divmod(pix, dHo, n, ho);
        float best[V];
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            const int h = 2 * ho + (d >> 1), w = 2 * wo + (d & 1);
            float v[V];
            Vec<T>::load(raw + (((int64_t)n * H + h) * W + w) * C + cv * V, v);
#pragma unroll
            for (int k = 0; k < V; ++k) {
                const float a = leaky(fmaf(v[k], sc[k], sh[k]));
                best[k] = d == 0 ? a : fmaxf(best[k], a);
            }
        }
        Vec<T>::store(pooled + (((int64_t)n * Ho + ho) * Wo + wo) * C + cv * V, best);
    }
}

template <typename T>
int pool_act(const T *raw, T *pooled, int N, int H, int W, int C, BnState bn, cudaStream_t s) {
    ProfScope _prof(PROF_GLUE, s);
    const int64_t total = (int64_t)N * (H / 2) * (W / 2) * (C / Vec<T>::N);
    HPFG_REQUIRE(total < (1ll << 31), "pool_act: tensor too large for 32-bit indexing");
    HPFG_CUDA_CHECK(launch_pdl(pool_act_kernel<T>, ew_grid(total), 256, 0, s, raw, pooled, N, H, W, C, bn, make_fastdiv(C / Vec<T>::N), make_fastdiv(W / 2), make_fastdiv(H / 2)));
    HPFG_LAUNCH_CHECK();
    return HPFG_OK;
}

// ------------------------------------------------------------------------------------------------ upcat
// bilinear source coordinates exactly as ATen's upsample_bilinear2d (align_corners=True):
// scale = (in-1)/(out-1) in fp32, src = scale*dst, i0 = (int)src, lambda1 = src - i0.
__device__ __forceinline__ void bilinear_src(int dst, float scale, int in_size, int &i0, int &i1, float &l0, float &l1) {
    const float src = scale * (float)dst;
    i0 = (int)src;
    if (i0 > in_size - 1) i0 = in_size - 1;
    i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
    l1 = src - (float)i0;
    l0 = 1.f - l1;
}

// blockIdx.y = 0: skip half (encoder feature = leaky(bn(raw))), 1: bilinearly upsampled half -- the two code paths never
// share a warp (with one grid a warp held both kinds of channel vectors and executed both paths: 244 instructions per
// 16-byte item, issue-bound at 20 % of HBM bandwidth in ncu).
template <typename T>
__global__ void __launch_bounds__(256) upcat_kernel(const T *__restrict__ raw_skip, BnState bn, const T *__restrict__ low,
                                                    T *__restrict__ cat, int N, int h, int w, int F, float sh_, float sw_,
                                                    FastDiv dFV, FastDiv dW, FastDiv dH) {
    pdl_prologue();
    constexpr int V = Vec<T>::N;
    const int H = 2 * h, W = 2 * w;
    const uint32_t total = (uint32_t)N * H * W * dFV.d;
    // the grid stride is a multiple of FV: a thread keeps one channel vector
    const uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t cv = i0 - fast_div(i0, dFV) * dFV.d;
    if (blockIdx.y == 0) {
        float sc[V], sf[V];
#pragma unroll
        for (int k = 0; k < V; ++k) { sc[k] = bn.scale[cv * V + k]; sf[k] = bn.shift[cv * V + k]; }
        for (uint32_t i = i0; i < total; i += gridDim.x * blockDim.x) {
            const uint32_t pix = fast_div(i, dFV);
            float v[V], out[V];
            Vec<T>::load(raw_skip + (int64_t)pix * F + cv * V, v);
#pragma unroll
            for (int k = 0; k < V; ++k) out[k] = leaky(fmaf(v[k], sc[k], sf[k]));
            Vec<T>::store(cat + (int64_t)pix * (2 * F) + cv * V, out);
        }
    } else {
        for (uint32_t i = i0; i < total; i += gridDim.x * blockDim.x) {
            uint32_t pix = fast_div(i, dFV), x, y, n;
            const uint32_t opix = pix;
            fast_divmod(pix, dW, pix, x);
            fast_divmod(pix, dH, n, y);
            int y0, y1, x0, x1;
            float ly0, ly1, lx0, lx1;
            bilinear_src(y, sh_, h, y0, y1, ly0, ly1);
            bilinear_src(x, sw_, w, x0, x1, lx0, lx1);
            float a[V], b[V], c[V], d[V], out[V];
            const T *base = low + (int64_t)n * h * w * F + cv * V;
            Vec<T>::load(base + ((int64_t)y0 * w + x0) * F, a);
            Vec<T>::load(base + ((int64_t)y0 * w + x1) * F, b);
            Vec<T>::load(base + ((int64_t)y1 * w + x0) * F, c);
            Vec<T>::load(base + ((int64_t)y1 * w + x1) * F, d);
#pragma unroll
            for (int k = 0; k < V; ++k) out[k] = ly0 * (lx0 * a[k] + lx1 * b[k]) + ly1 * (lx0 * c[k] + lx1 * d[k]);
            Vec<T>::store(cat + (int64_t)opix * (2 * F) + F + cv * V, out);
        }
    }
}

template <typename T>
int upcat(const T *raw_skip, BnState bn_skip, const T *low, T *cat, int N, int h, int w, int F, cudaStream_t s) {
    ProfScope _prof(PROF_GLUE, s);
    const int64_t total = (int64_t)N * 4 * h * w * (F / Vec<T>::N);     // per half
    const float sh_ = (2 * h > 1) ? (float)(h - 1) / (float)(2 * h - 1) : 0.f;
    const float sw_ = (2 * w > 1) ? (float)(w - 1) / (float)(2 * w - 1) : 0.f;
    HPFG_REQUIRE(total < (1ll << 31), "upcat: tensor too large for 32-bit indexing");
    HPFG_CUDA_CHECK(launch_pdl(upcat_kernel<T>, dim3((unsigned)ew_grid(total), 2), 256, 0, s, raw_skip, bn_skip, low, cat, N, h, w, F, sh_, sw_, make_fastdiv(F / Vec<T>::N), make_fastdiv(2 * w), make_fastdiv(2 * h)));
    HPFG_LAUNCH_CHECK();
    return HPFG_OK;
}

// ----------------------------------------------------------------------------------------------- bn_bwd
// g = dact * dropout' * leaky'(z);  xhat = (raw-mean)*invstd.
// MODE 0: reduce sum(g), sum(g*raw) into partials (the finalize kernel turns the raw moment into sum(g*xhat)).
// MODE 1: draw = scale*(g - c1 - xhat*c2) = scale*g + B*raw + D with B = -scale*c2*invstd, D = -scale*c1 - B*mean.
// MODE 2: as MODE 1 when the first tensor already holds g (written by skip_pool_bwd_gstat): no activation / dropout math.
// Few per-channel constants on purpose: 64 registers per thread keep four CTAs per SM resident (ncu: at 84 registers
// the kernel ran two CTAs per SM and reached 33-41 % of HBM bandwidth).
#ifndef HPFG_GLUE_MINBLOCKS
#define HPFG_GLUE_MINBLOCKS 4
#endif
template <typename T, int MODE>
__global__ void __launch_bounds__(256, HPFG_GLUE_MINBLOCKS) bn_bwd_kernel(const T *__restrict__ dact, const T *__restrict__ raw,
                                                        T *__restrict__ draw, int64_t M, int C, BnState bn, DropSpec drop,
                                                        float *__restrict__ partials) {
    pdl_prologue();
    constexpr int V = Vec<T>::N;
    __shared__ float red[2 * 256 * V];
    const int CV = C / V, R = 256 / CV;
    const int cv = threadIdx.x % CV, r = threadIdx.x / CV;
    const bool active = r < R;
    float sc[V], sh[V], kb[V], kd[V];        // MODE 0: kb/kd are the running sums
#pragma unroll
    for (int k = 0; k < V; ++k) {
        const int c = cv * V + k;
        sc[k] = bn.scale[c]; sh[k] = bn.shift[c];
        kb[k] = kd[k] = 0.f;
        if (MODE >= 1) {
            kb[k] = -sc[k] * bn.c2[c] * bn.invstd[c];
            kd[k] = -sc[k] * bn.c1[c] - kb[k] * bn.mean[c];
        }
    }
    const int64_t rows_per_block = (M + gridDim.x - 1) / gridDim.x;
    const int64_t p0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t p1 = (p0 + rows_per_block < M) ? p0 + rows_per_block : M;
    if (active)
#pragma unroll 2
        for (int64_t p = p0 + r; p < p1; p += R) {
            float d[V], x[V];
            Vec<T>::load(dact + p * C + cv * V, d);
            Vec<T>::load(raw + p * C + cv * V, x);
            uint32_t mbits = 0xffffffffu;
            if (drop.bits) {
                const int64_t e = p * C + cv * V;          // V divides 32, so the V bits sit in one word
                mbits = drop.bits[e >> 5] >> (e & 31);
            }
#pragma unroll
            for (int k = 0; k < V; ++k) {
                float g = d[k];
                if (MODE != 2) {
                    const float z = fmaf(x[k], sc[k], sh[k]);
                    g = d[k] * leaky_grad(z);
                    if (drop.bits) g = ((mbits >> k) & 1u) ? g * drop.inv_keep : 0.f;
                }
                if (MODE == 0) { kb[k] += g; kd[k] = fmaf(g, x[k], kd[k]); }
                else d[k] = fmaf(sc[k], g, fmaf(kb[k], x[k], kd[k]));
            }
            if (MODE >= 1) Vec<T>::store(draw + p * C + cv * V, d);
        }
    if (MODE == 0) {
        // deterministic cross-row reduction through shared memory: red[which][r][c]
        if (active) {
#pragma unroll
            for (int k = 0; k < V; ++k) {
                red[(0 * 256 + r * CV + cv) * V + k] = kb[k];
                red[(1 * 256 + r * CV + cv) * V + k] = kd[k];
            }
        }
        __syncthreads();
        for (int o = threadIdx.x; o < 2 * C; o += 256) {
            const int which = o / C, c = o % C;
            float s = 0.f;
            for (int rr = 0; rr < R; ++rr) s += red[(which * 256 + rr * CV + c / V) * V + (c % V)];
            partials[(int64_t)blockIdx.x * 2 * C + o] = s;
        }
    }
}

__global__ void __launch_bounds__(256) bn_bwd_finalize_kernel(const float *__restrict__ partials, int P, int C,
                                                              double inv_count, BnState bn, float *dgamma, float *dbeta,
                                                              int accumulate) {
    pdl_prologue();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= C) return;
    const int c = warp;
    double s = 0.0, q = 0.0;
    for (int p = lane; p < P; p += 32) {
        s += (double)partials[(int64_t)p * 2 * C + c];
        q += (double)partials[(int64_t)p * 2 * C + C + c];
    }
    s = warp_sum(s);
    q = warp_sum(q);
    if (lane != 0) return;
    q = (double)bn.invstd[c] * (q - (double)bn.mean[c] * s);      // partials hold sum(g*raw): xhat = (raw - mean) * invstd
    bn.c1[c] = (float)(s * inv_count);
    bn.c2[c] = (float)(q * inv_count);
    {   // folded constants of pass 1 (computed as bn_bwd_kernel<1> computes them, in fp32, from the rounded c1 / c2)
        const float sc = bn.scale[c];
        const float kb = -sc * bn.c2[c] * bn.invstd[c];
        bn.kb[c] = kb;
        bn.kd[c] = -sc * bn.c1[c] - kb * bn.mean[c];
    }
    if (accumulate) { dgamma[c] += (float)q; dbeta[c] += (float)s; }
    else { dgamma[c] = (float)q; dbeta[c] = (float)s; }
}

template <typename T>
int bn_bwd(const T *dact, const T *raw, T *draw, int64_t M, int C, BnState bn, DropSpec drop, float *partials,
           int max_partials, float *dgamma, float *dbeta, int accumulate, cudaStream_t s, double *sync_sums) {
    ProfScope _prof(PROF_GLUE, s);
    int P = (int)((M + 63) / 64);
    if (P > kNumSMs * 4) P = kNumSMs * 4;
    if (P > max_partials) P = max_partials;
    HPFG_CUDA_CHECK(launch_pdl(bn_bwd_kernel<T, 0>, P, 256, 0, s, dact, raw, draw, M, C, bn, drop, partials));
    HPFG_LAUNCH_CHECK();
    if (sync_sums) {     // exact-global mode: local sums -> copy -> all-reduce -> finalize from (local, global)
        double *loc = sync_sums, *glob = sync_sums + 2 * 256;
        HPFG_CUDA_CHECK(launch_plain(bn_sum_partials_kernel, ceil_div(2 * C * 32, 256), 256, 0, s, (const float *)partials, P, 2 * C, loc, glob));
        HPFG_LAUNCH_CHECK();
        HPFG_RETURN_IF(sync_allreduce(glob, 2 * C, true, s));
        HPFG_CUDA_CHECK(launch_plain(bn_bwd_finalize_sums_kernel, ceil_div(C, 256), 256, 0, s, (const double *)loc, (const double *)glob, C,
                                   1.0 / ((double)M * sync_world()), bn, dgamma, dbeta, accumulate));
        HPFG_LAUNCH_CHECK();
    } else {
        HPFG_CUDA_CHECK(launch_pdl(bn_bwd_finalize_kernel, ceil_div(C * 32, 256), 256, 0, s, partials, P, C, 1.0 / (double)M, bn, dgamma, dbeta,
                                   accumulate));
        HPFG_LAUNCH_CHECK();
    }
    int P2 = (int)((M + 63) / 64);
    if (P2 > kNumSMs * 8) P2 = kNumSMs * 8;
    HPFG_CUDA_CHECK(launch_pdl(bn_bwd_kernel<T, 1>, P2, 256, 0, s, dact, raw, draw, M, C, bn, drop, nullptr));
    HPFG_LAUNCH_CHECK();
    return HPFG_OK;
}

int bn_bwd_reduce(const float *partials, int P, int C, int64_t count, BnState bn, float *dgamma, float *dbeta, int accumulate,
                  cudaStream_t s) {
    ProfScope _prof(PROF_GLUE, s);
    HPFG_CUDA_CHECK(launch_pdl(bn_bwd_finalize_kernel, ceil_div(C * 32, 256), 256, 0, s, partials, P, C, 1.0 / (double)count, bn, dgamma, dbeta,
                               accumulate));
    HPFG_LAUNCH_CHECK();
    return HPFG_OK;
}

// ------------------------------------------------------------------------------------------ skip_pool_bwd
// GSTAT: store g = dact * leaky'(bn(raw)) instead of dact and reduce (sum g | sum g*raw) per channel into one partial row per
// block (BatchNorm-backward pass 0 of the feature's own BatchNorm, fused: the kernel reads raw anyway).
template <typename T, bool GSTAT>
__global__ void __launch_bounds__(256, GSTAT ? 2 : 4) skip_pool_bwd_kernel(const T *__restrict__ dcat, const T *__restrict__ dpooled,
                                                            const T *__restrict__ raw, BnState bn, T *__restrict__ dact,
                                                            int N, int H, int W, int F, FastDiv dFV, FastDiv dWo, FastDiv dHo,
                                                            float *__restrict__ partials) {
    pdl_prologue();
    constexpr int V = Vec<T>::N;
    __shared__ float red[GSTAT ? 2 * 256 * V : 1];
    float s1[GSTAT ? V : 1], s2[GSTAT ? V : 1];
    if (GSTAT) {
#pragma unroll
        for (int k = 0; k < V; ++k) s1[k] = s2[k] = 0.f;
    }
    const int Ho = H >> 1, Wo = W >> 1;
    const uint32_t total = (uint32_t)N * Ho * Wo * dFV.d;
    const uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t cv = i0 - fast_div(i0, dFV) * dFV.d;      // fixed per thread (grid stride is a multiple of FV)
    float sc[V], sh[V];
#pragma unroll
    for (int k = 0; k < V; ++k) { sc[k] = bn.scale[cv * V + k]; sh[k] = bn.shift[cv * V + k]; }
    for (uint32_t i = i0; i < total; i += gridDim.x * blockDim.x) {
        uint32_t pix = fast_div(i, dFV), wo, ho, n;
        fast_divmod(pix, dWo, pix, wo);
        fast_divmod(pix, dHo, n, ho);
        float best[V], dp[V];
        int arg[V];
#pragma unroll
        for (int k = 0; k < V; ++k) { dp[k] = 0.f; arg[k] = 0; best[k] = 0.f; }
        if (dpooled) Vec<T>::load(dpooled + (((int64_t)n * Ho + ho) * Wo + wo) * F + cv * V, dp);
        float rawv[GSTAT ? 4 : 1][V];
#pragma unroll
        for (int d = 0; d < 4; ++d) {   // first maximum in (h,w) scan order, as max_pool2d_with_indices
            const int h = 2 * ho + (d >> 1), w = 2 * wo + (d & 1);
            float v[V];
            Vec<T>::load(raw + (((int64_t)n * H + h) * W + w) * F + cv * V, v);
#pragma unroll
            for (int k = 0; k < V; ++k) {
                const float a = leaky(fmaf(v[k], sc[k], sh[k]));
                if (d == 0 || a > best[k]) { best[k] = a; arg[k] = d; }
                if (GSTAT) rawv[d][k] = v[k];
            }
        }
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            const int h = 2 * ho + (d >> 1), w = 2 * wo + (d & 1);
            const int64_t p = ((int64_t)n * H + h) * W + w;
            float o[V];
            if (dcat) Vec<T>::load(dcat + p * (2 * F) + cv * V, o);
            else {
#pragma unroll
                for (int k = 0; k < V; ++k) o[k] = 0.f;
            }
            if (dpooled) {
#pragma unroll
                for (int k = 0; k < V; ++k) if (arg[k] == d) o[k] += dp[k];
            }
            if (GSTAT) {
#pragma unroll
                for (int k = 0; k < V; ++k) {
                    const float x = rawv[d][k];
                    const float g = o[k] * leaky_grad(fmaf(x, sc[k], sh[k]));
                    o[k] = g;
                    s1[k] += g;
                    s2[k] = fmaf(g, x, s2[k]);
                }
            }
            Vec<T>::store(dact + p * F + cv * V, o);
        }
    }
    if (GSTAT) {
        // deterministic block reduction: thread t owns channel vector cv = t % FV (the grid stride is a multiple of FV)
        const int FV = (int)dFV.d, R = 256 / FV, r = threadIdx.x / FV, c0 = threadIdx.x % FV;
#pragma unroll
        for (int k = 0; k < V; ++k) {
            red[threadIdx.x * V + k] = s1[k];
            red[(256 + threadIdx.x) * V + k] = s2[k];
        }
        (void)r; (void)c0;
        __syncthreads();
        for (int o = threadIdx.x; o < 2 * F; o += 256) {
            const int which = o / F, c = o % F;
            float s = 0.f;
            for (int rr = 0; rr < R; ++rr) s += red[(which * 256 + rr * FV + c / V) * V + (c % V)];
            partials[(int64_t)blockIdx.x * 2 * F + o] = s;
        }
    }
}

template <typename T>
int skip_pool_bwd(const T *dcat, const T *dpooled, const T *raw, BnState bn, T *dact, int N, int H, int W, int F,
                  cudaStream_t s) {
    ProfScope _prof(PROF_GLUE, s);
    const int64_t total = (int64_t)N * (H / 2) * (W / 2) * (F / Vec<T>::N);
    HPFG_REQUIRE(total < (1ll << 31), "skip_pool_bwd: tensor too large for 32-bit indexing");
    HPFG_CUDA_CHECK(launch_pdl(skip_pool_bwd_kernel<T, false>, ew_grid(total), 256, 0, s, dcat, dpooled, raw, bn, dact, N, H, W, F, make_fastdiv(F / Vec<T>::N), make_fastdiv(W / 2), make_fastdiv(H / 2), (float *)nullptr));
    HPFG_LAUNCH_CHECK();
    return HPFG_OK;
}

// pass 1 alone on a tensor that already holds g (after skip_pool_bwd_gstat + bn_bwd_reduce): draw = scale*g + kb*raw + kd
template <typename T>
int bn_bwd_from_g(const T *g, const T *raw, T *draw, int64_t M, int C, BnState bn, cudaStream_t s) {
    ProfScope _prof(PROF_GLUE, s);
    int P2 = (int)((M + 63) / 64);
    if (P2 > kNumSMs * 8) P2 = kNumSMs * 8;
    DropSpec none{nullptr, 1.f};
    HPFG_CUDA_CHECK(launch_pdl(bn_bwd_kernel<T, 2>, P2, 256, 0, s, g, raw, draw, M, C, bn, none, (float *)nullptr));
    HPFG_LAUNCH_CHECK();
    return HPFG_OK;
}

template <typename T>
int skip_pool_bwd_gstat(const T *dcat, const T *dpooled, const T *raw, BnState bn, T *g_out, int N, int H, int W, int F,
                        float *partials, int max_partials, int *P, cudaStream_t s) {
    ProfScope _prof(PROF_GLUE, s);
    const int64_t total = (int64_t)N * (H / 2) * (W / 2) * (F / Vec<T>::N);
    HPFG_REQUIRE(total < (1ll << 31), "skip_pool_bwd: tensor too large for 32-bit indexing");
    HPFG_REQUIRE(256 % (F / Vec<T>::N) == 0, "skip_pool_bwd_gstat: channel vectors must divide the block");
    int grid = ew_grid(total);
    grid = std::min(grid, std::min(kNumSMs * 4, max_partials));       // one partial row per block
    HPFG_CUDA_CHECK(launch_pdl(skip_pool_bwd_kernel<T, true>, grid, 256, 0, s, dcat, dpooled, raw, bn, g_out, N, H, W, F, make_fastdiv(F / Vec<T>::N), make_fastdiv(W / 2), make_fastdiv(H / 2), partials));
    HPFG_LAUNCH_CHECK();
    *P = grid;
    return HPFG_OK;
}

// ------------------------------------------------------------------------------------------------ up_bwd
template <typename T>
__global__ void __launch_bounds__(256) up_bwd_kernel(const T *__restrict__ dcat, T *__restrict__ dlow, int N, int h, int w,
                                                     int F, float sh_, float sw_, FastDiv dFV, FastDiv dw, FastDiv dh) {
    pdl_prologue();
    constexpr int V = Vec<T>::N;
    const int H = 2 * h, W = 2 * w;
    const uint32_t total = (uint32_t)N * h * w * dFV.d;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        uint32_t cv, pix, xx, yy, n;
        fast_divmod(i, dFV, pix, cv);
        fast_divmod(pix, dw, pix, xx);
        fast_divmod(pix, dh, n, yy);
        const int x = (int)xx, y = (int)yy;
        float acc[V];
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] = 0.f;
        // destination rows / columns whose bilinear footprint can touch source row y / column x: dst in [2y-2, 2y+3];
        // per-axis weights first (12 coordinate evaluations instead of 6 + 36)
        float wy[6], wx[6];
#pragma unroll
        for (int t = 0; t < 6; ++t) {
            const int Y = 2 * y - 2 + t, X = 2 * x - 2 + t;
            wy[t] = wx[t] = 0.f;
            if (Y >= 0 && Y < H) {
                int y0, y1; float ly0, ly1;
                bilinear_src(Y, sh_, h, y0, y1, ly0, ly1);
                if (y0 == y) wy[t] += ly0;
                if (y1 == y) wy[t] += ly1;
            }
            if (X >= 0 && X < W) {
                int x0, x1; float lx0, lx1;
                bilinear_src(X, sw_, w, x0, x1, lx0, lx1);
                if (x0 == x) wx[t] += lx0;
                if (x1 == x) wx[t] += lx1;
            }
        }
#pragma unroll
        for (int ty = 0; ty < 6; ++ty) {
            if (wy[ty] == 0.f) continue;
            const T *row = dcat + (((int64_t)n * H + (2 * y - 2 + ty)) * W + (2 * x - 2)) * (2 * F) + F + cv * V;
#pragma unroll
            for (int tx = 0; tx < 6; ++tx) {
                if (wx[tx] == 0.f) continue;
                float g[V];
                Vec<T>::load(row + (int64_t)tx * (2 * F), g);
                const float wgt = wy[ty] * wx[tx];
#pragma unroll
                for (int k = 0; k < V; ++k) acc[k] = fmaf(wgt, g[k], acc[k]);
            }
        }
        Vec<T>::store(dlow + (((int64_t)n * h + y) * w + x) * F + cv * V, acc);
    }
}

template <typename T>
int up_bwd(const T *dcat, T *dlow, int N, int h, int w, int F, cudaStream_t s) {
    ProfScope _prof(PROF_GLUE, s);
    const int64_t total = (int64_t)N * h * w * (F / Vec<T>::N);
    const float sh_ = (2 * h > 1) ? (float)(h - 1) / (float)(2 * h - 1) : 0.f;
    const float sw_ = (2 * w > 1) ? (float)(w - 1) / (float)(2 * w - 1) : 0.f;
    HPFG_REQUIRE(total < (1ll << 31), "up_bwd: tensor too large for 32-bit indexing");
    HPFG_CUDA_CHECK(launch_pdl(up_bwd_kernel<T>, ew_grid(total), 256, 0, s, dcat, dlow, N, h, w, F, sh_, sw_, make_fastdiv(F / Vec<T>::N), make_fastdiv(w), make_fastdiv(h)));
    HPFG_LAUNCH_CHECK();
    return HPFG_OK;
}

// ------------------------------------------------------------------------------------------ dropout_bits
// Keep-mask bits in NHWC element order for up to 8 tensors in ONE launch (blockIdx.y = tensor).  Library stream: one
// Philox4x32-10 call per 8 consecutive NHWC elements (counter = element index / 8, key = seed, counter words 2-3 =
// offset + tensor); each 32-bit output word gives two 16-bit uniforms, keep <=> u16 >= p * 65536 (Bernoulli(1-p) to
// within 8e-6, as nn.Dropout).  User masks (parity harness) arrive as NCHW uint8 and are gathered into the same
// bit layout.
struct DropJob {
    uint32_t *bits;
    const uint8_t *mask;      // NCHW uint8 keep mask or nullptr
    int H, W, C;
    uint32_t thresh;          // p * 65536
    long long total, n_words;
};
struct DropJobs {
    int n;
    DropJob j[8];
};

__global__ void __launch_bounds__(256) dropout_bits_kernel(const DropJobs J, int N, uint64_t seed, uint64_t offset,
                                                           const uint64_t *offset_dev) {
    pdl_prologue();
    const DropJob &D = J.j[blockIdx.y];
    const uint64_t off = (offset_dev ? *offset_dev : offset) + blockIdx.y;   // device scalar: CUDA-graph replays
    for (long long wi = (long long)blockIdx.x * blockDim.x + threadIdx.x; wi < D.n_words; wi += (long long)gridDim.x * blockDim.x) {
        uint32_t word = 0;
        if (D.mask) {
            for (int b = 0; b < 32; ++b) {
                const long long e = wi * 32 + b;
                if (e >= D.total) break;
                const int c = (int)(e % D.C);
                long long pix = e / D.C;
                const int x = (int)(pix % D.W);
                pix /= D.W;
                const int y = (int)(pix % D.H), n = (int)(pix / D.H);
                word |= (D.mask[(((long long)n * D.C + c) * D.H + y) * D.W + x] != 0 ? 1u : 0u) << b;
            }
        } else {
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
                const uint64_t ctr = (uint64_t)wi * 4 + g8;
                const uint4 r = philox4x32_10(make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)),
                                              make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)off, (uint32_t)(off >> 32)));
                const uint32_t v[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    word |= ((v[q] & 0xffffu) >= D.thresh ? 1u : 0u) << (8 * g8 + 2 * q);
                    word |= ((v[q] >> 16) >= D.thresh ? 1u : 0u) << (8 * g8 + 2 * q + 1);
                }
            }
        }
        D.bits[wi] = word;
    }
}

int dropout_bits_multi(int n_jobs, uint32_t *const *bits, const uint8_t *const *masks_nchw, int N, const int *H, const int *W,
                       const int *C, const float *p, uint64_t seed, uint64_t offset, cudaStream_t s, const uint64_t *offset_dev) {
    ProfScope _prof(PROF_GLUE, s);
    HPFG_REQUIRE(n_jobs >= 1 && n_jobs <= 8, "dropout_bits_multi: 1..8 tensors per launch");
    DropJobs J{};
    J.n = n_jobs;
    long long max_words = 1;
    for (int i = 0; i < n_jobs; ++i) {
        DropJob &D = J.j[i];
        D.bits = bits[i];
        D.mask = masks_nchw ? masks_nchw[i] : nullptr;
        D.H = H[i]; D.W = W[i]; D.C = C[i];
        D.thresh = (uint32_t)(p[i] * 65536.0f + 0.5f);
        D.total = (long long)N * H[i] * W[i] * C[i];
        D.n_words = (D.total + 31) / 32;
        max_words = std::max(max_words, D.n_words);
    }
    HPFG_CUDA_CHECK(launch_pdl(dropout_bits_kernel, dim3((unsigned)ew_grid(max_words), (unsigned)n_jobs), 256, 0, s, J, N, seed, offset, offset_dev));
    HPFG_LAUNCH_CHECK();
    return HPFG_OK;
}

int dropout_bits(uint32_t *bits, const uint8_t *mask_nchw, int N, int H, int W, int C, float p, uint64_t seed,
                 uint64_t offset, cudaStream_t s) {
    return dropout_bits_multi(1, &bits, mask_nchw ? &mask_nchw : nullptr, N, &H, &W, &C, &p, seed, offset, s, nullptr);
}

// -------------------------------------------------------------------------------------- nhwc_to_nchw_f32
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T *__restrict__ src, float *__restrict__ dst, int N, int H, int W, int C,
                                    const float *__restrict__ bias) {
    pdl_prologue();
    const int64_t total = (int64_t)N * C * H * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        int64_t t = i / W;
        const int y = (int)(t % H);
        t /= H;
        const int c = (int)(t % C), n = (int)(t / C);
        dst[i] = to_f32(src[(((int64_t)n * H + y) * W + x) * C + c]) + (bias ? bias[c] : 0.f);
    }
}

template <typename T>
int nhwc_to_nchw_f32(const T *src, float *dst, int N, int H, int W, int C, const float *bias, cudaStream_t s) {
    ProfScope _prof(PROF_GLUE, s);
    HPFG_CUDA_CHECK(launch_pdl(nhwc_to_nchw_kernel<T>, ew_grid((int64_t)N * C * H * W), 256, 0, s, src, dst, N, H, W, C, bias));
    HPFG_LAUNCH_CHECK();
    return HPFG_OK;
}

// ------------------------------------------------------- activated feature tap / external gradient (UNet_Plus necks)
// dst[n,c,y,x] (fp32 NCHW) = leaky(raw[n,y,x,c]*scale[c] + shift[c]): the ConvBlock output the consumers' loaders would
// build on the fly, materialised for a caller outside the plan (model/unet.py:196,201 feature[-1]).
template <typename T>
__global__ void act_nhwc_to_nchw_kernel(const T *__restrict__ raw, const float *__restrict__ scale,
                                        const float *__restrict__ shift, float *__restrict__ dst, int N, int HW, int C) {
    pdl_prologue();
    __shared__ float tile[32][33];      // 32 pixels x 32 channels, transposed through shared memory (both sides coalesced)
    const int n = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 256 threads: 8 rows per pass
    for (int r = ty; r < 32; r += 8) {
        const int pix = p0 + r, c = c0 + tx;
        float v = 0.f;
        if (pix < HW && c < C) v = leaky(to_f32(raw[((int64_t)n * HW + pix) * C + c]) * scale[c] + shift[c]);
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r, pix = p0 + tx;
        if (pix < HW && c < C) dst[((int64_t)n * C + c) * HW + pix] = tile[tx][r];
    }
}

template <typename T>
int act_nhwc_to_nchw_f32(const T *raw, BnState bn, float *dst, int N, int H, int W, int C, cudaStream_t s) {
    ProfScope _prof(PROF_GLUE, s);
    const int HW = H * W;
    HPFG_CUDA_CHECK(launch_pdl(act_nhwc_to_nchw_kernel<T>, dim3((HW + 31) / 32, (C + 31) / 32, N), 256, 0, s, raw, bn.scale,
                               bn.shift, dst, N, HW, C));
    HPFG_LAUNCH_CHECK();
    return HPFG_OK;
}

// dst[n,y,x,c] (T NHWC) += src[n,c,y,x] (fp32 NCHW): an external gradient joins the data-gradient chain.
template <typename T>
__global__ void add_nchw_to_nhwc_kernel(T *__restrict__ dst, const float *__restrict__ src, int N, int HW, int C) {
    pdl_prologue();
    __shared__ float tile[32][33];
    const int n = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r, pix = p0 + tx;
        tile[r][tx] = (pix < HW && c < C) ? src[((int64_t)n * C + c) * HW + pix] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int pix = p0 + r, c = c0 + tx;
        if (pix < HW && c < C) {
            T *d = dst + ((int64_t)n * HW + pix) * C + c;
            *d = from_f32<T>(to_f32(*d) + tile[tx][r]);
        }
    }
}

template <typename T>
int add_nchw_f32_to_nhwc(T *dst, const float *src, int N, int H, int W, int C, cudaStream_t s) {
    ProfScope _prof(PROF_GLUE, s);
    const int HW = H * W;
    HPFG_CUDA_CHECK(launch_pdl(add_nchw_to_nhwc_kernel<T>, dim3((HW + 31) / 32, (C + 31) / 32, N), 256, 0, s, dst, src, N, HW, C));
    HPFG_LAUNCH_CHECK();
    return HPFG_OK;
}

#define INSTANTIATE(T)                                                                                              \
    template int pool_act<T>(const T *, T *, int, int, int, int, BnState, cudaStream_t);                            \
    template int upcat<T>(const T *, BnState, const T *, T *, int, int, int, int, cudaStream_t);                    \
    template int bn_bwd<T>(const T *, const T *, T *, int64_t, int, BnState, DropSpec, float *, int, float *,       \
                           float *, int, cudaStream_t, double *);                                                   \
    template int skip_pool_bwd<T>(const T *, const T *, const T *, BnState, T *, int, int, int, int, cudaStream_t); \
    template int bn_bwd_from_g<T>(const T *, const T *, T *, int64_t, int, BnState, cudaStream_t);                   \
    template int skip_pool_bwd_gstat<T>(const T *, const T *, const T *, BnState, T *, int, int, int, int, float *, int, int *, cudaStream_t); \
    template int up_bwd<T>(const T *, T *, int, int, int, int, cudaStream_t);                                       \
    template int act_nhwc_to_nchw_f32<T>(const T *, BnState, float *, int, int, int, int, cudaStream_t);           \
    template int add_nchw_f32_to_nhwc<T>(T *, const float *, int, int, int, int, cudaStream_t);                      \
    template int nhwc_to_nchw_f32<T>(const T *, float *, int, int, int, int, const float *, cudaStream_t);
INSTANTIATE(float)
INSTANTIATE(bf16)

}  // namespace hpfg

// ---- layer-isolated test hook for the bf16 glue kernels (include/hpfg_b200.h: hpfg_glue_debug) -----------------------------
using namespace hpfg;
extern "C" int hpfg_glue_debug(int op, int N, int H, int W, int C, const void *a, const void *b, const void *c, const float *scale,
                               const float *shift, const float *mean, const float *invstd, const uint8_t *keep_mask_nchw, float p_drop,
                               void *out_bf16, float *out_f32, void *stream) {
    HPFG_REQUIRE(op >= 0 && op <= 5 && a && out_bf16, "hpfg_glue_debug: bad arguments");
    HPFG_REQUIRE(C % 16 == 0 && C <= 256, "hpfg_glue_debug: C must be a multiple of 16, at most 256");
    cudaStream_t s = (cudaStream_t)stream;
    float *mem = nullptr, *partials = nullptr;
    uint32_t *bits = nullptr;
    const int max_partials = kNumSMs * 8;
    HPFG_CUDA_CHECK(cudaMalloc(&mem, 8 * 256 * 4));
    HPFG_CUDA_CHECK(cudaMalloc(&partials, (size_t)max_partials * 2 * 256 * 4));
    HPFG_CUDA_CHECK(cudaMemsetAsync(mem, 0, 8 * 256 * 4, s));
    BnState st{mem, mem + 256, mem + 512, mem + 768, mem + 1024, mem + 1280, mem + 1536, mem + 1792};
    if (scale) HPFG_CUDA_CHECK(cudaMemcpyAsync(st.scale, scale, C * 4, cudaMemcpyDeviceToDevice, s));
    if (shift) HPFG_CUDA_CHECK(cudaMemcpyAsync(st.shift, shift, C * 4, cudaMemcpyDeviceToDevice, s));
    if (mean) HPFG_CUDA_CHECK(cudaMemcpyAsync(st.mean, mean, C * 4, cudaMemcpyDeviceToDevice, s));
    if (invstd) HPFG_CUDA_CHECK(cudaMemcpyAsync(st.invstd, invstd, C * 4, cudaMemcpyDeviceToDevice, s));
    if (keep_mask_nchw) {
        HPFG_CUDA_CHECK(cudaMalloc(&bits, ((size_t)N * H * W * C + 31) / 32 * 4));
        HPFG_RETURN_IF(dropout_bits(bits, keep_mask_nchw, N, H, W, C, p_drop, 0, 0, s));
    }
    int rc = HPFG_OK;
    const bf16 *A = (const bf16 *)a, *B = (const bf16 *)b, *Cc = (const bf16 *)c;
    bf16 *O = (bf16 *)out_bf16;
    switch (op) {
        case 0: rc = pool_act<bf16>(A, O, N, H, W, C, st, s); break;                       // a = raw [N,H,W,C] -> [N,H/2,W/2,C]
        case 1: rc = upcat<bf16>(A, st, B, O, N, H, W, C, s); break;                       // a = skip raw [N,2H,2W,C], b = low [N,H,W,C] -> cat [N,2H,2W,2C]
        case 2: {                                                                          // a = dact, b = raw [N,H,W,C] -> draw; out_f32 = dgamma | dbeta
            DropSpec ds{bits, bits ? 1.f / (1.f - p_drop) : 1.f};
            rc = bn_bwd<bf16>(A, B, O, (int64_t)N * H * W, C, st, ds, partials, max_partials, out_f32, out_f32 + C, 0, s, nullptr);
            break;
        }
        case 3: rc = skip_pool_bwd<bf16>(A, B, Cc, st, O, N, H, W, C, s); break;           // a = dcat [N,H,W,2C], b = dpooled [N,H/2,W/2,C], c = raw [N,H,W,C]
        case 4: rc = up_bwd<bf16>(A, O, N, H, W, C, s); break;                             // a = dcat [N,2H,2W,2C] -> dlow [N,H,W,C]
        case 5: {                                                                          // as 3, fused with BatchNorm-backward pass 0: out = g; out_f32 = dgamma | dbeta | kb | kd
            int P = 0;
            rc = skip_pool_bwd_gstat<bf16>(A, B, Cc, st, O, N, H, W, C, partials, max_partials, &P, s);
            if (rc == HPFG_OK) rc = bn_bwd_reduce(partials, P, C, (int64_t)N * H * W, st, out_f32, out_f32 + C, 0, s);
            if (rc == HPFG_OK) {
                cudaMemcpyAsync(out_f32 + 2 * C, st.kb, C * 4, cudaMemcpyDeviceToDevice, s);
                cudaMemcpyAsync(out_f32 + 3 * C, st.kd, C * 4, cudaMemcpyDeviceToDevice, s);
            }
            break;
        }
    }
    cudaStreamSynchronize(s);
    cudaFree(mem);
    cudaFree(partials);
    if (bits) cudaFree(bits);
    if (rc == HPFG_OK) HPFG_CUDA_CHECK(cudaGetLastError());
    return rc;
}
