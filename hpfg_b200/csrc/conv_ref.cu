// CUDA-core direct convolution (see conv_ref.cuh).  Shared-memory tiled, fp32 FMA, deterministic reductions.
// Follows nn.Conv2d(kernel_size=3, padding=1) / nn.Conv2d(kernel_size=1) as used in model/unet.py:18,22,50,99.
#include "conv_ref.cuh"

namespace hpfg {

constexpr int kTile = 8;    // 8x8 output pixels per block
constexpr int kCK = 8;      // input channels per shared-memory stage
constexpr int kBN = 64;     // output channels per block

__device__ __forceinline__ float load_xf(const void *base, int64_t off, bool is_bf16) {
    return is_bf16 ? __bfloat162float(reinterpret_cast<const bf16 *>(base)[off])
                   : reinterpret_cast<const float *>(base)[off];
}

// Stage a (kTile+KS-1)^2 x kCK input patch (zero padded, producer's BN+LeakyReLU+dropout applied) in smem.
template <typename TI, int KS>
__device__ __forceinline__ void stage_input(float (*in_s)[(kTile + KS - 1) * (kTile + KS - 1) + 1], const TView &in,
                                            int n, int y0, int x0, int ci0, int H, int W, int Cin, const LoadXform &xf) {
    constexpr int TS = kTile + KS - 1, PAD = KS / 2;
    const TI *src = reinterpret_cast<const TI *>(in.p);
    for (int e = threadIdx.x; e < TS * TS * kCK; e += blockDim.x) {
        const int ci = e % kCK, pix = e / kCK;
        const int gy = y0 + pix / TS - PAD, gx = x0 + pix % TS - PAD, c = ci0 + ci;
        float v = 0.f;
        if (c < Cin && gy >= 0 && gy < H && gx >= 0 && gx < W) {
            v = to_f32(src[n * in.sn + gy * in.sh + gx * in.sw + c * in.sc]);
            if (xf.scale) v = leaky(fmaf(v, xf.scale[c], xf.shift[c]));
            if (xf.drop.bits) {
                const int64_t idx = (((int64_t)n * H + gy) * W + gx) * Cin + c;
                v = ((xf.drop.bits[idx >> 5] >> (idx & 31)) & 1u) ? v * xf.drop.inv_keep : 0.f;
            }
        }
        in_s[ci][pix] = v;
    }
}

template <typename TI, typename TO, int KS>
__global__ void __launch_bounds__(256) conv_ref_fprop_kernel(TView in, TView out, const float *__restrict__ wpk,
                                                             const float *__restrict__ bias, int N, int H, int W, int Cin,
                                                             int Cout, LoadXform xf, float *__restrict__ stats) {
    pdl_prologue();
    constexpr int TS = kTile + KS - 1, KK = KS * KS;
    __shared__ float in_s[kCK][TS * TS + 1];
    __shared__ float w_s[KK][kCK][kBN];
    __shared__ float red[2][16][kBN];
    const int tiles_w = (W + kTile - 1) / kTile, tiles_h = (H + kTile - 1) / kTile;
    int t = blockIdx.x;
    const int tx0 = (t % tiles_w) * kTile;
    t /= tiles_w;
    const int ty0 = (t % tiles_h) * kTile, n = t / tiles_h;
    const int co0 = blockIdx.y * kBN;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int py = (ty * 4) / kTile, px0 = (ty * 4) % kTile;   // this thread's 4 pixels: row py, cols px0..px0+3

    for (int ci0 = 0; ci0 < Cin; ci0 += kCK) {
        stage_input<TI, KS>(in_s, in, n, ty0, tx0, ci0, H, W, Cin, xf);
        for (int e = threadIdx.x; e < KK * kCK * kBN; e += blockDim.x) {
            const int co = e % kBN, ci = (e / kBN) % kCK, tap = e / (kBN * kCK);
            w_s[tap][ci][co] = (ci0 + ci < Cin && co0 + co < Cout) ? wpk[((int64_t)tap * Cin + ci0 + ci) * Cout + co0 + co] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int tap = 0; tap < KK; ++tap) {
            const int r = tap / KS, s = tap % KS;
#pragma unroll
            for (int ci = 0; ci < kCK; ++ci) {
                float a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = in_s[ci][(py + r) * TS + px0 + i + s];
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = w_s[tap][ci][tx + 16 * j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
        }
        __syncthreads();
    }

    TO *dst = reinterpret_cast<TO *>(out.p);
    float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    const int gy = ty0 + py;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gx = tx0 + px0 + i;
        if (gy < H && gx < W) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int co = co0 + tx + 16 * j;
                if (co < Cout) {
                    const float v = acc[i][j];
                    s1[j] += v;
                    s2[j] += v * v;
                    dst[n * out.sn + gy * out.sh + gx * out.sw + co * out.sc] = from_f32<TO>(v + (bias ? bias[co] : 0.f));
                }
            }
        }
    }
    if (stats) {   // fixed-order block reduction -> one partial row per pixel tile
#pragma unroll
        for (int j = 0; j < 4; ++j) { red[0][ty][tx + 16 * j] = s1[j]; red[1][ty][tx + 16 * j] = s2[j]; }
        __syncthreads();
        if (threadIdx.x < 2 * kBN) {
            const int which = threadIdx.x / kBN, c = threadIdx.x % kBN;
            if (co0 + c < Cout) {
                float s = 0.f;
#pragma unroll
                for (int k = 0; k < 16; ++k) s += red[which][k][c];
                stats[(int64_t)blockIdx.x * 2 * Cout + which * Cout + co0 + c] = s;
            }
        }
    }
}

int conv_ref_num_tiles(int N, int H, int W) { return N * ((H + kTile - 1) / kTile) * ((W + kTile - 1) / kTile); }

template <typename TI, typename TO>
int conv_ref_fprop(TView in, TView out, const float *wpk, const float *bias, int N, int H, int W, int Cin, int Cout,
                   int KS, LoadXform xf, float *stats, cudaStream_t s) {
    ProfScope _prof(PROF_CONV_CUDA, s);
    dim3 grid(conv_ref_num_tiles(N, H, W), (Cout + kBN - 1) / kBN);
    if (KS == 3)
        HPFG_CUDA_CHECK(launch_pdl(conv_ref_fprop_kernel<TI, TO, 3>, grid, 256, 0, s, in, out, wpk, bias, N, H, W, Cin, Cout, xf, stats));
    else if (KS == 1)
        HPFG_CUDA_CHECK(launch_pdl(conv_ref_fprop_kernel<TI, TO, 1>, grid, 256, 0, s, in, out, wpk, bias, N, H, W, Cin, Cout, xf, stats));
    else {
        set_error("conv_ref_fprop: kernel size must be 1 or 3");
        return HPFG_ERR_UNSUPPORTED;
    }
    HPFG_LAUNCH_CHECK();
    return HPFG_OK;
}

// ------------------------------------------------------------------------------------------------ wgrad
template <typename TI, typename TD, int KS>
__global__ void __launch_bounds__(256) conv_ref_wgrad_kernel(TView in, TView dout, int N, int H, int W, int Cin, int Cout,
                                                             LoadXform xf, float *__restrict__ scratch, int co_chunks) {
    pdl_prologue();
    constexpr int TS = kTile + KS - 1, KK = KS * KS;
    __shared__ float in_s[kCK][TS * TS + 1];
    __shared__ float d_s[kTile * kTile][kBN];
    const int ci0 = (blockIdx.x / co_chunks) * kCK, co0 = (blockIdx.x % co_chunks) * kBN;
    const int ci = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_w = (W + kTile - 1) / kTile, tiles_h = (H + kTile - 1) / kTile;
    const int num_tiles = N * tiles_h * tiles_w;
    const TD *dsrc = reinterpret_cast<const TD *>(dout.p);
    float acc[KK][2], accb[2] = {0.f, 0.f};
#pragma unroll
    for (int k = 0; k < KK; ++k) acc[k][0] = acc[k][1] = 0.f;

    for (int t = blockIdx.y; t < num_tiles; t += gridDim.y) {
        int tt = t;
        const int x0 = (tt % tiles_w) * kTile;
        tt /= tiles_w;
        const int y0 = (tt % tiles_h) * kTile, n = tt / tiles_h;
        stage_input<TI, KS>(in_s, in, n, y0, x0, ci0, H, W, Cin, xf);
        for (int e = threadIdx.x; e < kTile * kTile * kBN; e += blockDim.x) {
            const int co = e % kBN, pix = e / kBN;
            const int gy = y0 + pix / kTile, gx = x0 + pix % kTile;
            float v = 0.f;
            if (gy < H && gx < W && co0 + co < Cout)
                v = to_f32(dsrc[n * dout.sn + gy * dout.sh + gx * dout.sw + (co0 + co) * dout.sc]);
            d_s[pix][co] = v;
        }
        __syncthreads();
#pragma unroll 4
        for (int pix = 0; pix < kTile * kTile; ++pix) {
            const int py = pix / kTile, px = pix % kTile;
            const float d0 = d_s[pix][lane], d1 = d_s[pix][lane + 32];
            accb[0] += d0;
            accb[1] += d1;
#pragma unroll
            for (int tap = 0; tap < KK; ++tap) {
                const float a = in_s[ci][(py + tap / KS) * TS + px + tap % KS];
                acc[tap][0] = fmaf(a, d0, acc[tap][0]);
                acc[tap][1] = fmaf(a, d1, acc[tap][1]);
            }
        }
        __syncthreads();
    }
    float *dst = scratch + (int64_t)blockIdx.y * ((int64_t)KK * Cin * Cout + Cout);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int co = co0 + lane + 32 * h;
        if (co >= Cout) continue;
        if (ci0 + ci < Cin)
#pragma unroll
            for (int tap = 0; tap < KK; ++tap) dst[((int64_t)tap * Cin + ci0 + ci) * Cout + co] = acc[tap][h];
        if (ci0 == 0 && ci == 0) dst[(int64_t)KK * Cin * Cout + co] = accb[h];
    }
}

// sum the S split partials in a fixed order and scatter into the flat OIHW gradient (+ bias gradient)
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float *__restrict__ scratch, int S, int Cin, int Cout, int KK,
                                                           float *__restrict__ dw_oihw, float *__restrict__ dbias,
                                                           int accumulate) {
    pdl_prologue();
    const int64_t per = (int64_t)KK * Cin * Cout + Cout;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < per; e += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < S; ++k) s += scratch[(int64_t)k * per + e];
        if (e < (int64_t)KK * Cin * Cout) {
            const int co = (int)(e % Cout), ci = (int)((e / Cout) % Cin), tap = (int)(e / ((int64_t)Cout * Cin));
            float *d = dw_oihw + ((int64_t)co * Cin + ci) * KK + tap;
            *d = accumulate ? *d + s : s;
        } else if (dbias) {
            float *d = dbias + (e - (int64_t)KK * Cin * Cout);
            *d = accumulate ? *d + s : s;
        }
    }
}

static int wgrad_splits(int N, int H, int W, int Cin, int Cout) {
    const int chunks = ((Cin + kCK - 1) / kCK) * ((Cout + kBN - 1) / kBN);
    int S = (2 * kNumSMs + chunks - 1) / chunks;
    const int tiles = conv_ref_num_tiles(N, H, W);
    if (S > 64) S = 64;
    if (S > tiles) S = tiles;
    if (S < 1) S = 1;
    return S;
}

int64_t conv_ref_wgrad_scratch_floats(int N, int H, int W, int Cin, int Cout, int KS) {
    return (int64_t)wgrad_splits(N, H, W, Cin, Cout) * ((int64_t)KS * KS * Cin * Cout + Cout);
}

template <typename TI, typename TD>
int conv_ref_wgrad(TView in, TView dout, int N, int H, int W, int Cin, int Cout, int KS, LoadXform xf, float *scratch,
                   int64_t scratch_floats, float *dw_oihw, float *dbias, int accumulate, cudaStream_t s) {
    ProfScope _prof(PROF_WGRAD, s);
    const int S = wgrad_splits(N, H, W, Cin, Cout);
    HPFG_REQUIRE(conv_ref_wgrad_scratch_floats(N, H, W, Cin, Cout, KS) <= scratch_floats, "conv_ref_wgrad: scratch too small");
    const int co_chunks = (Cout + kBN - 1) / kBN;
    dim3 grid(((Cin + kCK - 1) / kCK) * co_chunks, S);
    if (KS == 3)
        HPFG_CUDA_CHECK(launch_pdl(conv_ref_wgrad_kernel<TI, TD, 3>, grid, 256, 0, s, in, dout, N, H, W, Cin, Cout, xf, scratch, co_chunks));
    else if (KS == 1)
        HPFG_CUDA_CHECK(launch_pdl(conv_ref_wgrad_kernel<TI, TD, 1>, grid, 256, 0, s, in, dout, N, H, W, Cin, Cout, xf, scratch, co_chunks));
    else {
        set_error("conv_ref_wgrad: kernel size must be 1 or 3");
        return HPFG_ERR_UNSUPPORTED;
    }
    HPFG_LAUNCH_CHECK();
    const int64_t per = (int64_t)KS * KS * Cin * Cout + Cout;
    int blocks = (int)((per + 255) / 256);
    if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
    HPFG_CUDA_CHECK(launch_pdl(wgrad_reduce_kernel, blocks, 256, 0, s, scratch, S, Cin, Cout, KS * KS, dw_oihw, dbias, accumulate));
    HPFG_LAUNCH_CHECK();
    return HPFG_OK;
}

// ---------------------------------------------------------------------------------------- weight packing
__global__ void pack_weights_ref_kernel(const float *__restrict__ w, float *__restrict__ wf, float *__restrict__ wd, int Cin,
                                        int Cout, int KS) {
    pdl_prologue();
    const int KK = KS * KS;
    const int64_t total = (int64_t)Cout * Cin * KK;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int tap = (int)(e % KK), ci = (int)((e / KK) % Cin), co = (int)(e / ((int64_t)KK * Cin));
        const float v = w[e];
        wf[((int64_t)tap * Cin + ci) * Cout + co] = v;
        if (wd) wd[((int64_t)(KK - 1 - tap) * Cout + co) * Cin + ci] = v;   // rot180 + transpose for dgrad
    }
}

int pack_weights_ref(const float *w_oihw, float *wpk_fprop, float *wpk_dgrad, int Cin, int Cout, int KS, cudaStream_t s) {
    ProfScope _prof(PROF_PACK, s);
    const int64_t total = (int64_t)Cout * Cin * KS * KS;
    int blocks = (int)((total + 255) / 256);
    if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
    HPFG_CUDA_CHECK(launch_pdl(pack_weights_ref_kernel, blocks, 256, 0, s, w_oihw, wpk_fprop, wpk_dgrad, Cin, Cout, KS));
    HPFG_LAUNCH_CHECK();
    return HPFG_OK;
}

#define INST_F(TI, TO)                                                                                                 \
    template int conv_ref_fprop<TI, TO>(TView, TView, const float *, const float *, int, int, int, int, int, int, LoadXform, \
                                        float *, cudaStream_t);
#define INST_W(TI, TD)                                                                                                 \
    template int conv_ref_wgrad<TI, TD>(TView, TView, int, int, int, int, int, int, LoadXform, float *, int64_t, float *, \
                                        float *, int, cudaStream_t);
INST_F(float, float)
INST_F(float, bf16)
INST_F(bf16, float)
INST_F(bf16, bf16)
INST_W(float, float)
INST_W(float, bf16)
INST_W(bf16, float)
INST_W(bf16, bf16)

}  // namespace hpfg
