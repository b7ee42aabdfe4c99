// Tensor-core implicit-GEMM convolution for the bf16 path (sm_100a): tcgen05.mma with TMEM accumulators,
// TMA-staged NHWC bf16 halo tiles, warp-specialised persistent CTAs.
//
//   D[128 pixels, BN out-channels] = sum over taps (r,s) and input-channel steps of
//        A_tap[128 pixels, 16 ch] (smem, K-major)  x  W_tap[BN, 16 ch] (smem, K-major)
//
// One CTA tile = 16 rows x 8 columns of output pixels of one image.  The (16+2)x(8+2) input halo of a channel
// chunk is fetched ONCE by TMA through a 5-D "chunked" tensor map (8 ch, W, H, C/8, N: the chunk dimension has a
// 16-byte stride), so the bytes land directly in the UMMA operand order [chunk of 8 ch][halo pixel][8 ch];
// out-of-bounds = zero = the conv padding.  When the producer's BatchNorm + LeakyReLU + dropout are fused into
// this consumer's loader (model/unet.py:17-25), eight transform warps apply them in place.  The nine taps are nine
// UMMA descriptors into that ONE staged tile: start address shifted by (r*10+s)*16 bytes, SBO = one halo row.
// Input bytes therefore cross L2->SMEM once per tile instead of nine times.
// Epilogue (4 warps): tcgen05.ld -> per-channel sum / sum-of-squares partials for train-mode BatchNorm
// (shuffle butterfly) -> bf16 NHWC store.  fprop, dgrad (flipped/transposed packed weights) and the 1x1
// convolutions of the up-blocks all run through this kernel.
#include "conv_tc.cuh"

#include <algorithm>
#include <cstdlib>
#include <map>
#include <tuple>

#include "tc_common.cuh"

namespace hpfg {

// ------------------------------------------------------------------------------------------ configuration
constexpr int kXfThreads = 256;            // loader-transform threads (warps 4-11)
constexpr int kTcThreads = 128 + kXfThreads + 128;   // + warps 0-3 (TMA, MMA, TMEM alloc, idle) + 4 epilogue warps
constexpr int kMaxStages = 12;
constexpr int kSmemBudget = 200 * 1024;

template <int KS, int KC, int BN, bool RES, int MT>
struct TcCfg {
    static constexpr int PAD = KS / 2, KK = KS * KS;
    // one pipeline stage = MT side-by-side 16x8-pixel UMMA tiles (16 rows x 8*MT columns) sharing one halo fetch:
    // per-stage fixed costs (barrier round trips, tile bookkeeping, proxy fences) are amortised over MT*128 pixels
    static constexpr int TWP = kTW * MT;
    static constexpr int HH = kTH + KS - 1, HW = TWP + KS - 1;     // halo tile
    static constexpr int NPIX = HH * HW;
    static constexpr int NCH = KC / 8;                             // 16-byte channel chunks per stage
    // operand tile [chunk][halo pixel][8 ch], written in exactly this order by TMA (5-D chunked tensor map)
    static constexpr int CH_STRIDE = NPIX * 16;
    static constexpr int OP_BYTES = NCH * CH_STRIDE;
    static constexpr int B_TAP_BYTES = KC * BN * 2;                // [KC/8][BN][8]
    static constexpr int B_BYTES = KK * B_TAP_BYTES;
    static constexpr int al(int v) { return (v + 127) / 128 * 128; }
    static constexpr int OFF_B = al(OP_BYTES);
    // RES: the whole packed weight of the layer (one k-chunk, one n-block) stays resident in shared memory
    static constexpr int STAGE_BYTES = OFF_B + (RES ? 0 : al(B_BYTES));
    static constexpr int RESB_BYTES = RES ? al(B_BYTES) : 0;
    static constexpr int FIXED_BYTES = 1024 /*barriers*/ + 2 * 256 * 4 /*scale,shift*/ + 4 * 2 * BN * 4 /*stat partials*/;
    static constexpr int STAGES_RAW = (kSmemBudget - FIXED_BYTES - RESB_BYTES) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > kMaxStages ? kMaxStages : STAGES_RAW;
    static constexpr int SMEM_BYTES = RESB_BYTES + STAGES * STAGE_BYTES + FIXED_BYTES + 1024 /*alignment slack*/;
    static constexpr int ACC_COLS = MT * BN;                       // accumulator columns per TMEM buffer
    static constexpr int TMEM_COLS = (2 * ACC_COLS <= 32) ? 32 : (2 * ACC_COLS <= 64) ? 64 : (2 * ACC_COLS <= 128) ? 128 : (2 * ACC_COLS <= 256) ? 256 : 512;
    static_assert(STAGES >= 2, "need at least a double buffer");
    static_assert(2 * ACC_COLS <= 512, "accumulators exceed TMEM");
};

struct TcConvParams {
    const bf16 *bpk;          // packed weights: [n_block][k_chunk][tap][KC/8][BN][8]
    bf16 *out;                // NHWC [N,H,W,Cout]
    const float *bias;        // Cout floats or nullptr
    const float *scale, *shift;   // per input channel (producer's fused BN affine) or nullptr = identity
    const uint8_t *dropbits;  // producer's dropout keep bits, NHWC bit order, or nullptr
    float inv_keep;
    float *stats;             // [gridDim.x][2*Cout] per-CTA partial sums (sum | sum of squares) or nullptr
    float *out_nchw;          // if set: write fp32 NCHW [N,out_c_real,H,W] (+bias) instead of bf16 NHWC (out_conv logits)
    int out_c_real;
    int N, H, W, Cin, Cout, tiles_h, tiles_w, m_tiles, n_blocks, k_chunks;
    int dbg;                  // bottleneck-isolation switches (env HPFG_TC_DBG, profiles/layer_bench.py only): 1 no MMA, 2 no stores, 4 no stats, 8 no TMA, 16 no transform
};

// Reduce 16 per-lane values over the 32 lanes of a warp with 16 shuffles (recursive halving); afterwards every
// lane holds the full column sum of column col16(lane).
__device__ __forceinline__ int col16(int lane) { return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1); }
__device__ __forceinline__ float butterfly16(const float (&v)[16], int lane) {
    float a[8], b[4], c[2], d;
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float send = h16 ? v[i] : v[i + 8], keep = h16 ? v[i + 8] : v[i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = h8 ? a[i] : a[i + 4], keep = h8 ? a[i + 4] : a[i];
        b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = h4 ? b[i] : b[i + 2], keep = h4 ? b[i + 2] : b[i];
        c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    {
        const float send = h2 ? c[0] : c[1], keep = h2 ? c[1] : c[0];
        d = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    return d + __shfl_xor_sync(0xffffffffu, d, 1);
}

// warps: 0 TMA producer, 1 MMA issuer, 2 TMEM alloc, 3 idle, 4-11 transform (only when a loader transform is
// fused), 12-15 epilogue.  Per-tile instruction counts of the single-thread roles are kept minimal: the MMA thread
// patches precomputed descriptor words with compile-time offsets and never computes tile coordinates.
template <int KS, int KC, int BN, bool RES, int MT>
__global__ void __launch_bounds__(kTcThreads, 1) tc_conv_kernel(const __grid_constant__ CUtensorMap tmA, const TcConvParams P) {
    using C = TcCfg<KS, KC, BN, RES, MT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *res_b = smem;                                         // resident weights (RES only)
    uint8_t *stage_base = smem + C::RESB_BYTES;
    uint8_t *fixed = stage_base + C::STAGES * C::STAGE_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(fixed);          // full[S] xf[S] empty[S] tfull[2] tempty[2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(fixed + 512);
    float *s_scale = reinterpret_cast<float *>(fixed + 1024);
    float *s_shift = s_scale + 256;
    float *s_part = s_shift + 256;                                 // [4 warps][2*BN]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_full = ptx::smem_u32(bars), bar_xf = bar_full + 8 * C::STAGES, bar_empty = bar_xf + 8 * C::STAGES;
    const uint32_t bar_tfull = bar_empty + 8 * C::STAGES, bar_tempty = bar_tfull + 16;
    const uint32_t stage_u32 = ptx::smem_u32(stage_base);

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            ptx::mbar_init(bar_full + 8 * s, 1);
            ptx::mbar_init(bar_xf + 8 * s, kXfThreads);
            ptx::mbar_init(bar_empty + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(bar_tfull + 8 * a, 1);
            ptx::mbar_init(bar_tempty + 8 * a, 128);
        }
        ptx::fence_barrier_init();
        ptx::prefetch_tensormap(&tmA);
    }
    if (warp == 2) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), C::TMEM_COLS);
    if (P.scale)
        for (int i = threadIdx.x; i < P.Cin; i += blockDim.x) { s_scale[i] = P.scale[i]; s_shift[i] = P.shift[i]; }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // this CTA's work items: fixed n-block, m-tiles mt0, mt0+mstep, ... (n_blocks divides the grid)
    const int total_work = P.m_tiles * P.n_blocks;
    const int nb = blockIdx.x % P.n_blocks, mt0 = blockIdx.x / P.n_blocks, mstep = gridDim.x / P.n_blocks;
    const int n_work = (total_work - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const bool xform = P.scale != nullptr && !(P.dbg & 16);

    if (warp == 0) {
        // ================================================================= TMA producer (warp-uniform, one lane issues)
        {
            TileIter ti;
            ti.init(mt0, mstep, P.tiles_h, P.tiles_w);
            int stage = 0, phase = 0;
            const uint32_t op_bytes = (P.dbg & 8) ? 0u : (uint32_t)C::OP_BYTES;
            const bf16 *bsrc = P.bpk + (size_t)nb * P.k_chunks * (C::B_BYTES / 2);
            for (int it = 0; it < n_work; ++it) {
                const int h0 = ti.th * kTH - C::PAD, w0 = ti.tw * C::TWP - C::PAD;
                for (int kc = 0; kc < P.k_chunks; ++kc) {
                    ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1, 1);
                    const uint32_t sb = stage_u32 + stage * C::STAGE_BYTES, fb = bar_full + 8 * stage;
                    if (ptx::elect_one()) {
                        if (RES) {       // weights ride along with the first tile only and stay resident
                            ptx::mbar_expect_tx(fb, op_bytes + (it == 0 ? C::B_BYTES : 0));
                            if (it == 0) ptx::bulk_load(ptx::smem_u32(res_b), P.bpk, C::B_BYTES, fb);
                        } else {
                            ptx::mbar_expect_tx(fb, op_bytes + C::B_BYTES);
                            ptx::bulk_load(sb + C::OFF_B, bsrc + (size_t)kc * (C::B_BYTES / 2), C::B_BYTES, fb);
                        }
                        if (!(P.dbg & 8))
                        ptx::tma_load_5d(sb, &tmA, fb, 0, w0, h0, kc * C::NCH, ti.n_img);
                    }
                    __syncwarp();
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                ti.next(P.tiles_h, P.tiles_w);
            }
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer (warp-uniform, one lane issues)
        {
            constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, BN, 0, 0);
            // descriptor words: hi = SBO | version, lo = (addr >> 4) | LBO << 16; per-MMA offsets are compile-time adds
            constexpr uint32_t a_hi = (uint32_t)((C::HW * 16) >> 4) | (1u << 14), b_hi = (uint32_t)(128 >> 4) | (1u << 14);
            const uint32_t a_lo0 = ((stage_u32 >> 4) & 0x3FFFu) | ((uint32_t)(C::CH_STRIDE >> 4) << 16);
            const uint32_t b_lo0 = (((RES ? ptx::smem_u32(res_b) : stage_u32 + C::OFF_B) >> 4) & 0x3FFFu) | ((uint32_t)((BN * 16) >> 4) << 16);
            int stage = 0, phase = 0;
            for (int it = 0; it < n_work; ++it) {
                const int acc = it & 1, acc_phase = (it >> 1) & 1;
                ptx::mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1, 2);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * C::ACC_COLS;
                for (int kc = 0; kc < P.k_chunks; ++kc) {
                    ptx::mbar_wait(bar_full + 8 * stage, phase, 3);
                    if (xform) ptx::mbar_wait(bar_xf + 8 * stage, phase, 4);
                    ptx::tc_fence_after();
                    const uint32_t a_lo = a_lo0 + stage * (C::STAGE_BYTES >> 4);
                    const uint32_t b_lo = RES ? b_lo0 : b_lo0 + stage * (C::STAGE_BYTES >> 4);
                    if (ptx::elect_one()) {
                    if (!(P.dbg & 1))
#pragma unroll
                    for (int j = 0; j < MT; ++j) {                 // UMMA tile j = output columns 8j..8j+7 of the stage
#pragma unroll
                    for (int tap = 0; tap < C::KK; ++tap) {
#pragma unroll
                        for (int kk = 0; kk < KC / 16; ++kk) {
                            const uint32_t ao = (uint32_t)((2 * kk * C::CH_STRIDE + ((tap / KS) * C::HW + (tap % KS) + j * kTW) * 16) >> 4);
                            const uint32_t bo = (uint32_t)((tap * C::B_TAP_BYTES + 2 * kk * BN * 16) >> 4);
                            const uint64_t ad = ((uint64_t)a_hi << 32) | (uint64_t)(a_lo + ao);
                            const uint64_t bd = ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + bo);
                            ptx::umma_bf16(d_tmem + j * BN, ad, bd, idesc, (kc | tap | kk) != 0);
                        }
                    }
                    }
                    ptx::umma_commit(bar_empty + 8 * stage);      // smem slot reusable once these MMAs retire
                    if (kc == P.k_chunks - 1) ptx::umma_commit(bar_tfull + 8 * acc);   // accumulator complete -> epilogue
                    }
                    __syncwarp();
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp >= 4 && warp < 4 + kXfThreads / 32) {
        // ================================================================= loader-transform warps (in place)
        if (xform) {
            const int t = threadIdx.x - 128;
            TileIter ti;
            ti.init(mt0, mstep, P.tiles_h, P.tiles_w);
            int stage = 0, phase = 0;
            for (int it = 0; it < n_work; ++it) {
                const int h0 = ti.th * kTH - C::PAD, w0 = ti.tw * C::TWP - C::PAD;
                const size_t img_px = (size_t)ti.n_img * P.H;
                for (int kc = 0; kc < P.k_chunks; ++kc) {
                    ptx::mbar_wait(bar_full + 8 * stage, phase, 5);
                    const uint32_t op = stage_u32 + stage * C::STAGE_BYTES;
                    for (int i = t; i < C::NPIX * C::NCH; i += kXfThreads) {
                        const int c = i / C::NPIX, p = i % C::NPIX;
                        const int gh = h0 + p / C::HW, gw = w0 + p % C::HW;
                        uint4 v = make_uint4(0u, 0u, 0u, 0u);     // conv zero padding applies AFTER the activation
                        if (gh >= 0 && gh < P.H && gw >= 0 && gw < P.W) {
                            v = ptx::lds128(op + i * 16);
                            float f[8];
                            unpack8(v, f);
                            const int ch = kc * KC + c * 8;
                            uint32_t keep = 0xffu;
                            if (P.dropbits) keep = P.dropbits[(((img_px + gh) * P.W + gw) * P.Cin + ch) >> 3];
                            const float4 sc0 = *reinterpret_cast<const float4 *>(s_scale + ch), sc1 = *reinterpret_cast<const float4 *>(s_scale + ch + 4);
                            const float4 sh0 = *reinterpret_cast<const float4 *>(s_shift + ch), sh1 = *reinterpret_cast<const float4 *>(s_shift + ch + 4);
                            const float scl[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
                            const float shf[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                float a = fmaf(f[j], scl[j], shf[j]);
                                a = a > 0.f ? a : kLeakySlope * a;
                                if (P.dropbits) a = ((keep >> j) & 1u) ? a * P.inv_keep : 0.f;
                                f[j] = a;
                            }
                            v = make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
                        }
                        ptx::sts128(op + i * 16, v);
                    }
                    ptx::fence_proxy_async_smem();
                    ptx::mbar_arrive(bar_xf + 8 * stage);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                ti.next(P.tiles_h, P.tiles_w);
            }
        }
    } else if (warp >= 4 + kXfThreads / 32) {
        // ================================================================= epilogue warps
        const int q = warp & 3;                      // TMEM lane quarter this warp may read
        const int m = q * 32 + lane;                 // output pixel within the tile
        const int et = threadIdx.x - (128 + kXfThreads);
        // BatchNorm statistics: every lane keeps the running sum of column col16(lane) of each 16-column group over
        // ALL tiles of this CTA -> one partial row per CTA
        float run1[BN / 16], run2[BN / 16];
#pragma unroll
        for (int gidx = 0; gidx < BN / 16; ++gidx) run1[gidx] = run2[gidx] = 0.f;
        TileIter ti;
        ti.init(mt0, mstep, P.tiles_h, P.tiles_w);
        for (int it = 0; it < n_work; ++it) {
            const int acc = it & 1, acc_phase = (it >> 1) & 1;
            ptx::mbar_wait(bar_tfull + 8 * acc, acc_phase, 6);
            ptx::tc_fence_after();
#pragma unroll
            for (int j = 0; j < MT; ++j) {
            const int gh = ti.th * kTH + m / kTW, gw = ti.tw * C::TWP + j * kTW + m % kTW;
            const bool valid = gh < P.H && gw < P.W;
            bf16 *orow = P.out + (((size_t)ti.n_img * P.H + gh) * P.W + gw) * P.Cout + nb * BN;
#pragma unroll
            for (int gidx = 0; gidx < BN / 16; ++gidx) {
                const int n0 = gidx * 16;
                uint32_t r[16];
                ptx::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + acc * C::ACC_COLS + j * BN + n0, r);
                ptx::tmem_ld_wait();
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = valid ? __uint_as_float(r[j]) : 0.f;
                if (P.stats && !(P.dbg & 4)) {
                    float sq[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) sq[j] = v[j] * v[j];
                    run1[gidx] += butterfly16(v, lane);
                    run2[gidx] += butterfly16(sq, lane);
                }
                if (P.dbg & 2) continue;
                if (valid && P.out_nchw) {
                    float *o = P.out_nchw + ((size_t)ti.n_img * P.out_c_real * P.H + gh) * P.W + gw;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (nb * BN + n0 + j < P.out_c_real)
                            o[(size_t)(nb * BN + n0 + j) * P.H * P.W] = v[j] + (P.bias ? P.bias[nb * BN + n0 + j] : 0.f);
                } else if (valid) {
                    if (P.bias) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] += P.bias[nb * BN + n0 + j];
                    }
                    uint4 lo = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
                    uint4 hi = make_uint4(pack2(v[8], v[9]), pack2(v[10], v[11]), pack2(v[12], v[13]), pack2(v[14], v[15]));
                    *reinterpret_cast<uint4 *>(orow + n0) = lo;
                    *reinterpret_cast<uint4 *>(orow + n0 + 8) = hi;
                }
            }
            }
            ptx::tc_fence_before();
            ptx::mbar_arrive(bar_tempty + 8 * acc);               // accumulator buffer free for the MMA warp
            ti.next(P.tiles_h, P.tiles_w);
        }
        if (P.stats) {    // once per CTA: combine the four epilogue warps, write this CTA's partial row (zeros elsewhere)
            if ((lane & 1) == 0) {
#pragma unroll
                for (int gidx = 0; gidx < BN / 16; ++gidx) {
                    s_part[q * 2 * BN + gidx * 16 + col16(lane)] = run1[gidx];
                    s_part[q * 2 * BN + BN + gidx * 16 + col16(lane)] = run2[gidx];
                }
            }
            ptx::named_bar_sync(1, 128);
            for (int i = et; i < 2 * P.Cout; i += 128) {
                const int which = i / P.Cout, c = i % P.Cout, n = c - nb * BN;
                float sum = 0.f;
                if (n >= 0 && n < BN && n_work > 0)
                    sum = s_part[which * BN + n] + s_part[2 * BN + which * BN + n] + s_part[4 * BN + which * BN + n] +
                          s_part[6 * BN + which * BN + n];
                P.stats[(size_t)blockIdx.x * 2 * P.Cout + i] = sum;
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------ weight packing
struct PackEntry {
    long long dst_begin;      // element offset into the packed buffer
    long long w_off;          // element offset of the OIHW fp32 weight in the flat parameter buffer
    int cin, cout, kk, KC, BN, dgrad;   // cin/cout: channel counts padded to multiples of 16 (layout)
    int cin_real, cout_real;            // true tensor dims (source indexing; padded entries are zero)
};
constexpr int kMaxPack = 48;
struct PackTable {
    int n;
    long long total;
    PackEntry e[kMaxPack];
};

__global__ void __launch_bounds__(256) tc_pack_kernel(const float *__restrict__ params, bf16 *__restrict__ dst, const PackTable T) {
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < T.total; g += (long long)gridDim.x * blockDim.x) {
        int lo = 0;
        while (lo + 1 < T.n && T.e[lo + 1].dst_begin <= g) ++lo;
        const PackEntry &E = T.e[lo];
        long long e = g - E.dst_begin;
        const int cin_v = E.dgrad ? E.cout : E.cin;
        const int k_chunks = cin_v / E.KC;
        const int k = (int)(e % 8);
        e /= 8;
        const int n = (int)(e % E.BN);
        e /= E.BN;
        const int k8 = (int)(e % (E.KC / 8));
        e /= (E.KC / 8);
        const int tap = (int)(e % E.kk);
        e /= E.kk;
        const int kc = (int)(e % k_chunks);
        const int nb = (int)(e / k_chunks);
        const int co_v = nb * E.BN + n, ci_v = kc * E.KC + k8 * 8 + k;
        float w = 0.f;
        if (!E.dgrad) {
            if (co_v < E.cout_real && ci_v < E.cin_real) w = params[E.w_off + ((long long)co_v * E.cin_real + ci_v) * E.kk + tap];
        } else if (ci_v < E.cout_real && co_v < E.cin_real) {       // rot180 + transpose
            w = params[E.w_off + ((long long)ci_v * E.cin_real + co_v) * E.kk + (E.kk - 1 - tap)];
        }
        dst[g] = __float2bfloat16_rn(w);
    }
}

// ------------------------------------------------------------------------------------------ host side
static void pick_cfg(int cin_v, int cout_v, int &KC, int &BN) {
    BN = std::min(cout_v, 128);
    KC = (cin_v == 16 || BN == 128) ? 16 : 32;
}

template <int KS, int KC, int BN, bool RES, int MT>
static int launch_cfg3(const CUtensorMap &map, const TcConvParams &P, cudaStream_t s) {
    using C = TcCfg<KS, KC, BN, RES, MT>;
    static bool attr_set = false;
    if (!attr_set) {
        HPFG_CUDA_CHECK(cudaFuncSetAttribute(tc_conv_kernel<KS, KC, BN, RES, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr_set = true;
    }
    const int total = P.m_tiles * P.n_blocks;
    const int grid = std::min(total, kNumSMs);
    tc_conv_kernel<KS, KC, BN, RES, MT><<<grid, kTcThreads, C::SMEM_BYTES, s>>>(map, P);
    HPFG_LAUNCH_CHECK();
    return HPFG_OK;
}
template <int KS, int KC, int BN, bool RES>
static int launch_cfg2(int mt, const CUtensorMap &map, const TcConvParams &P, cudaStream_t s) {
    if constexpr (KS == 3 && BN <= 32) {      // wide stages only where per-tile overheads dominate (few channels, large images)
        if (mt == 4) return launch_cfg3<KS, KC, BN, RES, 4>(map, P, s);
        if (mt == 2) return launch_cfg3<KS, KC, BN, RES, 2>(map, P, s);
    }
    return launch_cfg3<KS, KC, BN, RES, 1>(map, P, s);
}
template <int KS, int KC, int BN>
static int launch_cfg(int mt, const CUtensorMap &map, const TcConvParams &P, cudaStream_t s) {
    if constexpr (BN <= 64) {     // resident weights whenever the layer's whole weight is one (k-chunk, n-block) stage
        if (P.k_chunks == 1 && P.n_blocks == 1) return launch_cfg2<KS, KC, BN, true>(mt, map, P, s);
    }
    return launch_cfg2<KS, KC, BN, false>(mt, map, P, s);
}

template <int KS>
static int launch_ks(int KC, int BN, int mt, const CUtensorMap &map, const TcConvParams &P, cudaStream_t s) {
    if (KC == 16 && BN == 16) return launch_cfg<KS, 16, 16>(mt, map, P, s);
    if (KC == 16 && BN == 32) return launch_cfg<KS, 16, 32>(mt, map, P, s);
    if (KC == 16 && BN == 128) return launch_cfg<KS, 16, 128>(mt, map, P, s);
    if (KC == 32 && BN == 16) return launch_cfg<KS, 32, 16>(mt, map, P, s);
    if (KC == 32 && BN == 32) return launch_cfg<KS, 32, 32>(mt, map, P, s);
    if (KC == 32 && BN == 64) return launch_cfg<KS, 32, 64>(mt, map, P, s);
    set_error("tc conv: no kernel for KC=" + std::to_string(KC) + " BN=" + std::to_string(BN));
    return HPFG_ERR_UNSUPPORTED;
}

// Run one convolution (conv-view channels cin_v -> cout_v) on the tensor cores.
static int tc_run(int ks, int N, int H, int W, int cin_v, int cout_v, const void *in, void *out, const bf16 *bpk,
                  const float *bias, LoadXform xf, float *stats, int *P_out, cudaStream_t s, float *out_nchw = nullptr,
                  int out_c_real = 0) {
    ProfScope _prof(PROF_CONV_TC, s);
    int KC, BN;
    pick_cfg(cin_v, cout_v, KC, BN);
    // stage width: 4 / 2 UMMA tiles side by side when the image is wide enough to keep every SM busy with wide stages
    int mt = 1;
    if (ks == 3 && BN <= 32) {
        const int th = (H + kTH - 1) / kTH;
        if (W % (4 * kTW) == 0 && N * th * (W / (4 * kTW)) >= 2 * kNumSMs) mt = 4;
        else if (W % (2 * kTW) == 0 && N * th * (W / (2 * kTW)) >= 2 * kNumSMs) mt = 2;
    }
    CUtensorMap map;
    HPFG_RETURN_IF(make_map_chunked(&map, in, N, H, W, cin_v, KC / 8, kTW * mt + ks - 1, kTH + ks - 1));
    TcConvParams P{};
    P.bpk = bpk; P.out = (bf16 *)out; P.bias = bias;
    P.scale = xf.scale; P.shift = xf.shift;
    P.dropbits = reinterpret_cast<const uint8_t *>(xf.drop.bits); P.inv_keep = xf.drop.inv_keep;
    P.stats = stats;
    P.out_nchw = out_nchw; P.out_c_real = out_c_real;
    P.N = N; P.H = H; P.W = W; P.Cin = cin_v; P.Cout = cout_v;
    P.tiles_h = (H + kTH - 1) / kTH; P.tiles_w = (W + kTW * mt - 1) / (kTW * mt);
    P.m_tiles = N * P.tiles_h * P.tiles_w; P.n_blocks = cout_v / BN; P.k_chunks = cin_v / KC;
    { static const char *e = getenv("HPFG_TC_DBG"); P.dbg = e ? atoi(e) : 0; }
    if (P_out) *P_out = std::min(P.m_tiles * P.n_blocks, kNumSMs);
    return ks == 3 ? launch_ks<3>(KC, BN, mt, map, P, s) : launch_ks<1>(KC, BN, mt, map, P, s);
}

// micro-benchmark entry (wgrad_tc.cu: hpfg_conv_tc_bench): packs once per call (cheap) and launches one convolution
int tc_run_bench(int op, int ks, int N, int H, int W, int cin, int cout, const void *in, void *out, const float *w, const float *scale,
                 const float *shift, float *stats, cudaStream_t s) {
    static bf16 *packed = nullptr;
    static long long packed_n = 0;
    const long long n = (long long)cin * cout * ks * ks;
    if (packed_n < n) {
        if (packed) cudaFree(packed);
        if (cudaMalloc(&packed, (size_t)n * 2) != cudaSuccess) return HPFG_ERR_CUDA;
        packed_n = n;
        cudaMemsetAsync(packed, 0, (size_t)n * 2, s);
    }
    (void)w;
    const int cin_v = op ? cout : cin, cout_v = op ? cin : cout;
    LoadXform xf{};
    xf.scale = scale; xf.shift = shift; xf.drop.bits = nullptr; xf.drop.inv_keep = 1.f;
    return tc_run(ks, N, H, W, cin_v, cout_v, in, out, packed, nullptr, xf, stats, nullptr, s);
}

struct TcPlanState {
    bf16 *packed = nullptr;
    PackTable table;
    long long f_off[kNumConv], d_off[kNumConv];   // -1 = layer not on the tensor-core path
};

static int pad16(int c) { return (c + 15) / 16 * 16; }

int tc_plan_init(hpfg_unet_plan *p) {
    auto *st = new TcPlanState();
    p->tc = st;
    st->table.n = 0;
    long long off = 0;
    for (int i = 0; i < kNumConv; ++i) {
        const ConvLayer &cv = p->d.convs[i];
        st->f_off[i] = st->d_off[i] = -1;
        const int cinp = pad16(cv.cin), coutp = pad16(cv.cout);
        const long long n = (long long)cinp * coutp * cv.ks * cv.ks;
        for (int dg = 0; dg < 2; ++dg) {
            if (dg && i == 0) continue;                      // the network input needs no data gradient
            int KC, BN;
            pick_cfg(dg ? coutp : cinp, dg ? cinp : coutp, KC, BN);
            PackEntry &E = st->table.e[st->table.n++];
            E.dst_begin = off; E.w_off = cv.w_off; E.cin = cinp; E.cout = coutp; E.kk = cv.ks * cv.ks;
            E.cin_real = cv.cin; E.cout_real = cv.cout;
            E.KC = KC; E.BN = BN; E.dgrad = dg;
            (dg ? st->d_off[i] : st->f_off[i]) = off;
            off += n;
        }
    }
    st->table.total = off;
    // stats partial rows: one per 16x8 tile -- never more than the 8x8-tile count the workspace was sized for
    if (cudaMalloc(&st->packed, (size_t)off * sizeof(bf16)) != cudaSuccess) {
        set_error("tc_plan_init: cudaMalloc failed");
        cudaGetLastError();
        return HPFG_ERR_CUDA;
    }
    if (!get_encode()) {
        set_error("tc_plan_init: cuTensorMapEncodeTiled not available from the driver");
        return HPFG_ERR_CUDA;
    }
    return HPFG_OK;
}

void tc_plan_free(hpfg_unet_plan *p) {
    auto *st = reinterpret_cast<TcPlanState *>(p->tc);
    if (!st) return;
    if (st->packed) cudaFree(st->packed);
    delete st;
    p->tc = nullptr;
}

int tc_pack_all(hpfg_unet_plan *p, const float *params, cudaStream_t s) {
    ProfScope _prof(PROF_PACK, s);
    auto *st = reinterpret_cast<TcPlanState *>(p->tc);
    const int blocks = (int)std::min<long long>((st->table.total + 255) / 256, (long long)kNumSMs * 8);
    tc_pack_kernel<<<blocks, 256, 0, s>>>(params, st->packed, st->table);
    HPFG_LAUNCH_CHECK();
    return HPFG_OK;
}

int tc_fprop(hpfg_unet_plan *p, int conv, const void *in, void *out, LoadXform xf, float *stats, int *P, bool *done,
             cudaStream_t s) {
    auto *st = reinterpret_cast<TcPlanState *>(p->tc);
    const ConvLayer &cv = p->d.convs[conv];
    *done = false;
    if (st->f_off[conv] < 0) return HPFG_OK;
    HPFG_RETURN_IF(tc_run(cv.ks, p->N, cv.H, cv.W, pad16(cv.cin), cv.cout, in, out, st->packed + st->f_off[conv], nullptr, xf, stats, P, s));
    *done = true;
    return HPFG_OK;
}

int tc_fprop_1x1(hpfg_unet_plan *p, int conv, const void *in, void *out, LoadXform xf, const float *bias, bool *done,
                 cudaStream_t s) {
    auto *st = reinterpret_cast<TcPlanState *>(p->tc);
    const ConvLayer &cv = p->d.convs[conv];
    *done = false;
    if (st->f_off[conv] < 0) return HPFG_OK;
    HPFG_RETURN_IF(tc_run(cv.ks, p->N, cv.H, cv.W, cv.cin, cv.cout, in, out, st->packed + st->f_off[conv], bias, xf, nullptr, nullptr, s));
    *done = true;
    return HPFG_OK;
}

int tc_fprop_logits(hpfg_unet_plan *p, int conv, const void *in, LoadXform xf, const float *bias, float *logits_nchw, cudaStream_t s) {
    auto *st = reinterpret_cast<TcPlanState *>(p->tc);
    const ConvLayer &cv = p->d.convs[conv];
    return tc_run(cv.ks, p->N, cv.H, cv.W, cv.cin, pad16(cv.cout), in, nullptr, st->packed + st->f_off[conv], bias, xf, nullptr, nullptr, s,
                  logits_nchw, cv.cout);
}

int tc_dgrad(hpfg_unet_plan *p, int conv, const void *dout, void *din, bool *done, cudaStream_t s) {
    auto *st = reinterpret_cast<TcPlanState *>(p->tc);
    const ConvLayer &cv = p->d.convs[conv];
    *done = false;
    if (st->d_off[conv] < 0) return HPFG_OK;
    const LoadXform none{};
    HPFG_RETURN_IF(tc_run(cv.ks, p->N, cv.H, cv.W, pad16(cv.cout), cv.cin, dout, din, st->packed + st->d_off[conv], nullptr, none, nullptr, nullptr, s));
    *done = true;
    return HPFG_OK;
}

int tc_wgrad(hpfg_unet_plan *p, int conv, const void *in, LoadXform xf, const void *dout, float *dw_oihw, float *dbias,
             int accumulate, bool *done, cudaStream_t s) {
    const ConvLayer &cv = p->d.convs[conv];
    *done = false;
    HPFG_RETURN_IF(tc_wgrad_run(cv.ks, p->N, cv.H, cv.W, pad16(cv.cin), pad16(cv.cout), cv.cin, cv.cout, in, xf, dout, p->wscratch,
                                p->wscratch_floats, dw_oihw, dbias, accumulate, s));
    *done = true;
    return HPFG_OK;
}

}  // namespace hpfg

// ---- layer-isolated test hook ----------------------------------------------------------------------------
using namespace hpfg;

__global__ void tc_debug_reduce_stats(const float *partials, int P, int C2, float *out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C2) return;
    double s = 0.0;
    for (int p = 0; p < P; ++p) s += (double)partials[(size_t)p * C2 + c];
    out[c] = (float)s;
}

extern "C" int hpfg_conv_tc_debug(int op, int N, int H, int W, int cin, int cout, int ks, const void *in_bf16_nhwc,
                                  const float *w_oihw, const float *bias, const float *scale, const float *shift,
                                  void *out_bf16_nhwc, float *stats_out, void *stream) {
    HPFG_REQUIRE(op == 0 || op == 1, "hpfg_conv_tc_debug: op must be 0 (fprop) or 1 (dgrad)");
    HPFG_REQUIRE(cin % 16 == 0 && cout % 16 == 0 && (ks == 1 || ks == 3), "hpfg_conv_tc_debug: unsupported shape");
    cudaStream_t s = (cudaStream_t)stream;
    const int cin_v = op ? cout : cin, cout_v = op ? cin : cout;
    PackTable T{};
    T.n = 1;
    T.total = (long long)cin * cout * ks * ks;
    pick_cfg(cin_v, cout_v, T.e[0].KC, T.e[0].BN);
    T.e[0].dst_begin = 0; T.e[0].w_off = 0; T.e[0].cin = cin; T.e[0].cout = cout; T.e[0].kk = ks * ks; T.e[0].dgrad = op;
    T.e[0].cin_real = cin; T.e[0].cout_real = cout;
    bf16 *packed = nullptr;
    float *partials = nullptr;
    const int m_tiles = N * ((H + kTH - 1) / kTH) * ((W + kTW - 1) / kTW);
    HPFG_CUDA_CHECK(cudaMalloc(&packed, (size_t)T.total * 2));
    HPFG_CUDA_CHECK(cudaMalloc(&partials, (size_t)m_tiles * 2 * cout_v * 4));
    tc_pack_kernel<<<(int)std::min<long long>((T.total + 255) / 256, 1024), 256, 0, s>>>(w_oihw, packed, T);
    HPFG_LAUNCH_CHECK();
    LoadXform xf{};
    xf.scale = scale; xf.shift = shift; xf.drop.bits = nullptr; xf.drop.inv_keep = 1.f;
    int P = 0;
    int rc = tc_run(ks, N, H, W, cin_v, cout_v, in_bf16_nhwc, out_bf16_nhwc, packed, bias, xf, stats_out ? partials : nullptr, &P, s);
    if (rc == HPFG_OK && stats_out) {
        tc_debug_reduce_stats<<<(2 * cout_v + 127) / 128, 128, 0, s>>>(partials, P, 2 * cout_v, stats_out);
        ++g_launch_count;
    }
    cudaStreamSynchronize(s);
    cudaFree(packed);
    cudaFree(partials);
    if (rc == HPFG_OK) HPFG_CUDA_CHECK(cudaGetLastError());
    return rc;
}
