// Tensor-core weight gradient for the bf16 path (sm_100a):
//
//   dW[tap=(r,s)][ci][co] = sum over pixels  X_act[pixel + (r,s)][ci] * dY[pixel][co]
//
// with the pixel index as the reduction (K) dimension of tcgen05.mma and BOTH operands MN-major (pixels are the slow
// dimension of NHWC tiles).  To fill the M dimension (min 64) the KS column shifts of a tap row are stacked along M:
//
//   D_r[m = s*NB + ci][n = co]  +=  A_r[m][k = pixel] * B[n][k],     A_r = [X shifted by (r,0) | (r,1) | (r,2)],  B = dY
//
// i.e. KS MMAs (M = KS*NB padded to 64/128, N = COB) per 16-pixel K step instead of KS*KS MMAs with 3/4 of M wasted
// (an M=64, N=16 MMA costs 23 cycles of shared-memory operand fetch however many of its rows are useful).
// Staging per 16x8-pixel tile: TMA fetches the dY tile and the (16+2)x(8+2) halo of the layer input once, both
// directly in [8-channel chunk][pixel][8 ch] order (chunked 5-D maps).  The transform warps read the halo, apply the
// producer's BatchNorm+LeakyReLU+dropout (and restore the conv zero padding), and write the KS column-shifted copies
// [s][chunk][halo row][8 cols][8 ch]: with the copies contiguous, "8-row group g = s*(NB/8)+chunk" is an affine
// address (SBO = one copy-chunk), so ONE descriptor spans all KS shifts; the tap row r is a start-address offset.
// They also accumulate the bias gradient from dY.  The pixel dimension is split over persistent CTAs (split-K);
// each CTA keeps its accumulators in TMEM across all its tiles and writes one fp32 partial, and a fixed-order
// reduction kernel sums the partials straight into the flat OIHW gradient (deterministic).
#include <algorithm>
#include <cstdlib>

#include "conv_ref.cuh"
#include "conv_tc.cuh"
#include "tc_common.cuh"

namespace hpfg {

constexpr int kWgXfThreads = 256;   // transform threads (warps 4-11): loader transform of X into shifted copies + bias-gradient sums
constexpr int kWgThreads = 128 + kWgXfThreads + 128;   // warps 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-11 transform, 12-15 epilogue
#ifndef HPFG_WG_SMEM_KB
#define HPFG_WG_SMEM_KB 198      // leaves room for the streaming BatchNorm kernels' static shared memory next to a weight-gradient CTA
#endif
constexpr int kWgSmemBudget = HPFG_WG_SMEM_KB * 1024;

// MT: tile width multiplier (tile = 16 rows x 8*MT columns).  The TMA unit retires roughly one tensor load per ~500
// cycles however small its box (tests/probes/tma_rate_probe.cu), so the 16-channel layers use wide tiles: two loads
// then feed 256 instead of 128 pixels.
// TWO: the dY operand is built by the transform warps from two staged tiles, draw = sc*g + kb*raw + kd (BatchNorm backward of
// the layer's own BatchNorm folded into this kernel: the raw gradient tensor is never materialised).
template <int KS, int NB, int COB, int MT, bool TWO = false>
struct WgCfg {
    static constexpr int PAD = KS / 2, KK = KS * KS;
    static constexpr int TWP = kTW * MT, TPIX = kTH * TWP;                    // tile width / pixels
    static constexpr int HH = kTH + KS - 1, HW = TWP + KS - 1, NPIX_X = HH * HW;
    static constexpr int XR_CHS = NPIX_X * 16, XR_OP = (NB / 8) * XR_CHS;     // raw halo tile, TMA target: [chunk][halo pixel][8 ch]
    static constexpr int ROWP = TWP * 16;                                     // bytes per tile row of one chunk
    static constexpr int XC = HH * ROWP;                                      // one chunk of one shifted copy: [halo row][TWP cols][8 ch]
    static constexpr int X3_OP = KS * (NB / 8) * XC;                          // A operand: [shift s][chunk][halo row][TWP cols][8 ch]
    static constexpr int D_CHS = TPIX * 16, D_OP = (COB / 8) * D_CHS;         // dY tile, TMA target and B operand: [chunk][pixel][8 ch]
    static constexpr int al(int v) { return (v + 127) / 128 * 128; }
    // + one more 8-row group after the copies holding the constant (1,0,...,0) per pixel: its first accumulator row is
    // sum_pixels dY = the bias gradient, computed by the tensor core instead of an extra shared-memory pass over dY
    static constexpr int OFF_DOP = 0, OFF_DRAW = al(D_OP), OFF_XR = OFF_DRAW + (TWO ? al(D_OP) : 0), OFF_X3 = OFF_XR + al(XR_OP),
                         OFF_ONE = OFF_X3 + X3_OP;
    static constexpr int STAGE_BYTES = al(OFF_ONE + XC);
    static constexpr int MROWS = KS * NB;                                     // useful accumulator rows
    static constexpr int UM = MROWS + 8 <= 64 ? 64 : 128;                     // UMMA M (useful rows + the ones group)
    // the A descriptor spans UM/8 row groups; the groups past MROWS read whatever follows the copies in shared memory
    // (their D rows are never stored) -- keep those reads inside the allocation
    static constexpr int TAIL_PAD = al((UM / 8 - MROWS / 8 - 1) * XC);
    static constexpr int FIXED_BYTES = 1024 + 5 * 256 * 4;
    static constexpr int STAGES_RAW = (kWgSmemBudget - FIXED_BYTES - TAIL_PAD) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + TAIL_PAD + FIXED_BYTES + 1024;
    static constexpr int ACC = KS * COB;                                      // accumulator columns: D_r at column r*COB
    static constexpr int TMEM_COLS = (ACC <= 32) ? 32 : (ACC <= 64) ? 64 : (ACC <= 128) ? 128 : (ACC <= 256) ? 256 : 512;
    static_assert(STAGES >= 2, "need at least a double buffer");
    static_assert(ACC <= 512 && MROWS + 8 <= 128, "accumulators exceed TMEM");
    static_assert(X3_OP % 128 == 0, "the ones group must start at the copies' group pitch");
};

struct WgParams {
    const float *scale, *shift;     // producer transform of X (nullptr = identity)
    const float *dsc, *dkb, *dkd;   // TWO: BatchNorm-backward constants per output channel (draw = dsc*g + dkb*raw + dkd)
    const uint8_t *dropbits;
    float inv_keep;
    float *scratch;                 // [S][KK*Cin*Cout + Cout] fp32 partials
    int N, H, W, Cin, Cout, tiles_h, tiles_w, m_tiles, ci_blocks, co_blocks, S;
    long long *trace;               // optional (profiles/ only): CTA 0 records clock64 per role and tile, [role][32]
};
#define WG_TRACE(role, idx) do { if (P.trace && blockIdx.x == 0 && lane == 0 && (idx) < 32) P.trace[(role) * 32 + (idx)] = clock64(); } while (0)

#ifndef HPFG_WG_MINBLOCKS
#define HPFG_WG_MINBLOCKS 2      // at most 64 registers per thread (the kernels need 52-64): two 256-thread glue CTAs of the
                                 // main-stream BatchNorm / gather chain then fit next to a weight-gradient CTA and the two streams of
                                 // the backward pass really overlap (+2.2 % step throughput, profiles/r02_ab_wgrad_coresidency.txt)
#endif
template <int KS, int NB, int COB, int MT, bool TWO>
__global__ void __launch_bounds__(kWgThreads, HPFG_WG_MINBLOCKS) tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                 const __grid_constant__ CUtensorMap tmD,
                                                                 const __grid_constant__ CUtensorMap tmR, const WgParams P) {
    pdl_launch_dependents();
    using C = WgCfg<KS, NB, COB, MT, TWO>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *fixed = smem + C::STAGES * C::STAGE_BYTES + C::TAIL_PAD;
    uint64_t *bars = reinterpret_cast<uint64_t *>(fixed);          // full[S] xf[S] empty[S] done
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(fixed + 512);
    float *s_scale = reinterpret_cast<float *>(fixed + 1024);
    float *s_shift = s_scale + 256;
    float *s_dsc = s_shift + 256, *s_dkb = s_dsc + 256, *s_dkd = s_dkb + 256;

    // broadcast from lane 0: the warp index is warp-uniform for the compiler (role code stays in the uniform datapath)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t bar_full = ptx::smem_u32(bars), bar_xf = bar_full + 8 * C::STAGES, bar_empty = bar_xf + 8 * C::STAGES;
    const uint32_t bar_done = bar_empty + 8 * C::STAGES;
    const uint32_t smem_u32 = ptx::smem_u32(smem);

    // work assignment: block pair (co block, ci block) x pixel split
    const int n_bp = P.ci_blocks * P.co_blocks;
    const int bp = blockIdx.x % n_bp, split = blockIdx.x / n_bp;
    const int cib = bp % P.ci_blocks, cob = bp / P.ci_blocks;
    const int ci0 = cib * NB, co0 = cob * COB;
    const int n_work = (P.m_tiles - split + P.S - 1) / P.S;        // tiles split, split+S, ...
    const bool xform = P.scale != nullptr, want_bias = cib == 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            ptx::mbar_init(bar_full + 8 * s, 1);
            ptx::mbar_init(bar_xf + 8 * s, kWgXfThreads / 32);
            ptx::mbar_init(bar_empty + 8 * s, 1);
        }
        ptx::mbar_init(bar_done, 1);
        ptx::fence_barrier_init();
        ptx::prefetch_tensormap(&tmX);
        ptx::prefetch_tensormap(&tmD);
        if (TWO) ptx::prefetch_tensormap(&tmR);
    }
    if (warp == 2) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), C::TMEM_COLS);
    for (int i = threadIdx.x; i < C::STAGES * (C::XC / 16); i += blockDim.x)      // the constant ones group of every stage
        ptx::sts128(smem_u32 + (i / (C::XC / 16)) * C::STAGE_BYTES + C::OFF_ONE + (i % (C::XC / 16)) * 16, make_uint4(0x3f80u, 0u, 0u, 0u));
    ptx::fence_proxy_async_smem();
    pdl_wait();
    if (P.scale)
        for (int i = threadIdx.x; i < P.Cin; i += blockDim.x) { s_scale[i] = P.scale[i]; s_shift[i] = P.shift[i]; }
    if (TWO)
        for (int i = threadIdx.x; i < COB; i += blockDim.x) {
            const int co = co0 + i;
            const bool ok = co < P.Cout;
            s_dsc[i] = ok ? P.dsc[co] : 0.f; s_dkb[i] = ok ? P.dkb[co] : 0.f; s_dkd[i] = ok ? P.dkd[co] : 0.f;
        }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (P.trace && blockIdx.x == 0 && threadIdx.x == 0) P.trace[5 * 32 + 2] = clock64();

    if (warp == 0) {             // ==================================================== TMA producer (warp-uniform)
        TileIter ti;
        ti.init(split, P.S, P.tiles_h, P.tiles_w);
        int stage = 0, phase = 0;
#pragma unroll 1
        for (int it = 0; it < n_work; ++it) {
            const int h0 = ti.th * kTH, w0 = ti.tw * C::TWP;
            ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1, 11);
            WG_TRACE(0, it);
            const uint32_t sb = smem_u32 + stage * C::STAGE_BYTES;
            if (ptx::elect_one()) {
                ptx::mbar_expect_tx(bar_full + 8 * stage, C::D_OP * (TWO ? 2 : 1) + C::XR_OP);
                ptx::tma_load_5d(sb + C::OFF_DOP, &tmD, bar_full + 8 * stage, 0, w0, h0, co0 / 8, ti.n_img);
                if (TWO) ptx::tma_load_5d(sb + C::OFF_DRAW, &tmR, bar_full + 8 * stage, 0, w0, h0, co0 / 8, ti.n_img);
                ptx::tma_load_5d(sb + C::OFF_XR, &tmX, bar_full + 8 * stage, 0, w0 - C::PAD, h0 - C::PAD, ci0 / 8, ti.n_img);
            }
            __syncwarp();
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            ti.next(P.tiles_h, P.tiles_w);
        }
    } else if (warp == 1) {      // ==================================================== MMA issuer (warp-uniform)
        constexpr uint32_t idesc = ptx::umma_idesc_bf16(C::UM, COB, 1, 1);   // both operands MN-major
        // descriptor words: hi = SBO | version, lo = (addr >> 4) | LBO << 16; per-MMA offsets are compile-time adds
        // A = shifted X copies: M groups (8 rows = one (shift, chunk)) SBO = copy-chunk stride, K groups (8 pixels of one tile row) LBO = row pitch
        // B = dY: N groups (8 co) SBO = chunk stride, K groups (8 pixels of one tile row) LBO = row pitch
        constexpr uint32_t a_hi = (uint32_t)(C::XC >> 4) | (1u << 14), b_hi = (uint32_t)(C::D_CHS >> 4) | (1u << 14);
        const uint32_t a_lo0 = (((smem_u32 + C::OFF_X3) >> 4) & 0x3FFFu) | ((uint32_t)(C::ROWP >> 4) << 16);
        const uint32_t b_lo0 = (((smem_u32 + C::OFF_DOP) >> 4) & 0x3FFFu) | ((uint32_t)(C::ROWP >> 4) << 16);
        int stage = 0, phase = 0;
#pragma unroll 1
        for (int it = 0; it < n_work; ++it) {
            ptx::mbar_wait(bar_xf + 8 * stage, phase, 14);         // transform warps arrive after the TMA data was consumed and the copies written
            ptx::tc_fence_after();
            WG_TRACE(3, it);
            const uint32_t a_lo = a_lo0 + stage * (C::STAGE_BYTES >> 4), b_lo = b_lo0 + stage * (C::STAGE_BYTES >> 4);
            if (ptx::elect_one()) {
#pragma unroll 2
                for (int j = 0; j < 8 * MT; ++j) {       // K step = 16 pixels = tile rows 2*jr, 2*jr+1 x columns 8*jc .. 8*jc+7
                    const int jr = j / MT, jc = j % MT;
                    const uint64_t bd = ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + (uint32_t)((2 * jr * C::ROWP + jc * 128) >> 4));
#pragma unroll
                    for (int r = 0; r < KS; ++r) {       // tap row: start at halo row 2*jr + r of every copy
                        const uint64_t ad = ((uint64_t)a_hi << 32) | (uint64_t)(a_lo + (uint32_t)(((2 * jr + r) * C::ROWP + jc * 128) >> 4));
                        ptx::umma_bf16(tmem_base + r * COB, ad, bd, idesc, (it == 0 && j == 0) ? 0u : 1u);
                    }
                }
                ptx::umma_commit(bar_empty + 8 * stage);
                if (it == n_work - 1) ptx::umma_commit(bar_done);
            }
            __syncwarp();
            WG_TRACE(4, it);
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp >= 4 && warp < 4 + kWgXfThreads / 32) {
        // ================================================================================ transform warps
        const int t = threadIdx.x - 128;
        // thread -> one 8-channel chunk for the whole kernel (its affine lives in registers) and every XPSTEP-th halo pixel
        // (two chunks: lanes alternate chunks and stay bank-conflict free; four chunks would conflict 2-way, so there a warp
        // pair owns a chunk: consecutive lanes read consecutive pixels)
        constexpr int NCHK = NB / 8, XPSTEP = kWgXfThreads / NCHK;
        const int xc = NCHK == 2 ? t % NCHK : t / XPSTEP, xp0 = NCHK == 2 ? t / NCHK : t % XPSTEP, xch = ci0 + xc * 8;
        float scl[8], shf[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { scl[k] = xform ? s_scale[xch + k] : 1.f; shf[k] = xform ? s_shift[xch + k] : 0.f; }
        TileIter ti;
        ti.init(split, P.S, P.tiles_h, P.tiles_w);
        int stage = 0, phase = 0;
#pragma unroll 1
        for (int it = 0; it < n_work; ++it) {
            const int h0 = ti.th * kTH - C::PAD, w0 = ti.tw * C::TWP - C::PAD;
            const size_t img_px = (size_t)ti.n_img * P.H;
            ptx::mbar_wait(bar_full + 8 * stage, phase, 15);
            if (warp == 4) WG_TRACE(1, it);
            const uint32_t sb = smem_u32 + stage * C::STAGE_BYTES;
            for (int p = xp0; p < C::NPIX_X; p += XPSTEP) {
                const int hr = p / C::HW, hc = p % C::HW;
                uint4 v = ptx::lds128(sb + C::OFF_XR + (xc * C::NPIX_X + p) * 16);
                if (xform) {
                    const int gh = h0 + hr, gw = w0 + hc;
                    if (gh >= 0 && gh < P.H && gw >= 0 && gw < P.W) {
                        float f[8];
                        unpack8(v, f);
                        uint32_t keep = 0xffu;
                        if (P.dropbits) keep = P.dropbits[(((img_px + gh) * P.W + gw) * P.Cin + xch) >> 3];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            float a = fmaf(f[k], scl[k], shf[k]);
                            a = fmaxf(a, kLeakySlope * a);
                            if (P.dropbits) a = ((keep >> k) & 1u) ? a * P.inv_keep : 0.f;
                            f[k] = a;
                        }
                        v = make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
                    } else {
                        v = make_uint4(0u, 0u, 0u, 0u);        // conv zero padding applies AFTER the activation
                    }
                }
                // halo column hc lands in copy s at column hc - s
                const uint32_t dst = sb + C::OFF_X3 + xc * C::XC + (hr * C::TWP + hc) * 16;
#pragma unroll
                for (int s = 0; s < KS; ++s)
                    if (hc - s >= 0 && hc - s < C::TWP) ptx::sts128(dst + s * (NB / 8) * C::XC - s * 16, v);
            }
            if constexpr (TWO) {
                // dY operand in place: [chunk of 8 co][tile pixel][8 co] <- dsc*g + dkb*raw + dkd, zero outside the image
                // (a warp works on one chunk at a time: consecutive lanes = consecutive pixels, conflict-free 128-bit accesses)
                constexpr int NCD = COB / 8, GW = NCD >= 8 ? 1 : 8 / NCD;          // warps sharing one chunk
                const int xw = warp - 4;
                const bool tile_border = h0 + C::PAD + kTH > P.H || w0 + C::PAD + C::TWP > P.W;
#pragma unroll 1
                for (int dc = NCD >= 8 ? xw : xw % NCD; dc < NCD; dc += 8) {
                    float a[8], b[8], d[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) { a[k] = s_dsc[dc * 8 + k]; b[k] = s_dkb[dc * 8 + k]; d[k] = s_dkd[dc * 8 + k]; }
#pragma unroll 2
                    for (int px = (NCD >= 8 ? 0 : (xw / NCD) * 32) + lane; px < C::TPIX; px += 32 * GW) {
                        const uint32_t addr = sb + C::OFF_DOP + (dc * C::TPIX + px) * 16;
                        float g[8], r[8];
                        unpack8(ptx::lds128(addr), g);
                        unpack8(ptx::lds128(addr + C::OFF_DRAW), r);
#pragma unroll
                        for (int k = 0; k < 8; ++k) g[k] = fmaf(a[k], g[k], fmaf(b[k], r[k], d[k]));
                        uint4 v = make_uint4(pack2(g[0], g[1]), pack2(g[2], g[3]), pack2(g[4], g[5]), pack2(g[6], g[7]));
                        if (tile_border) {
                            const int gh = h0 + C::PAD + px / C::TWP, gw = w0 + C::PAD + px % C::TWP;
                            if (gh >= P.H || gw >= P.W) v = make_uint4(0u, 0u, 0u, 0u);
                        }
                        ptx::sts128(addr, v);
                    }
                }
            }
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar_xf + 8 * stage);
            if (warp == 4) WG_TRACE(2, it);
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            ti.next(P.tiles_h, P.tiles_w);
        }
    } else if (warp >= 4 + kWgXfThreads / 32) {
        // ================================================================================ epilogue (once)
        // accumulator row m -> TMEM lane: M=128: lane = m; M=64: lane = (m/16)*32 + m%16 (measured, tests/probes)
        const int q = warp & 3;
        const int row = C::UM == 128 ? q * 32 + lane : (lane < 16 ? q * 16 + lane : C::MROWS);
        const int s_shift_idx = row / NB, ci = ci0 + row % NB;          // row = s*NB + ci
        if (n_work > 0) {
            ptx::mbar_wait(bar_done, 0, 16);
            ptx::tc_fence_after();
            if (q == 0) WG_TRACE(5, 0);
            float *dst = P.scratch + (size_t)split * ((size_t)C::KK * P.Cin * P.Cout + P.Cout);
            if (want_bias) {       // accumulator row MROWS of D_0 = ones x dY = sum of dY over this CTA's pixels
                const bool mine = C::UM == 128 ? (q * 32 + lane == C::MROWS) : (lane < 16 && q * 16 + lane == C::MROWS);
#pragma unroll 1
                for (int n0 = 0; n0 < COB; n0 += 16) {
                    uint32_t v[16];
                    ptx::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + n0, v);
                    ptx::tmem_ld_wait();
                    if (mine) {
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj)
                            if (co0 + n0 + jj < P.Cout) dst[(size_t)C::KK * P.Cin * P.Cout + co0 + n0 + jj] = __uint_as_float(v[jj]);
                    }
                }
            }
#pragma unroll 1
            for (int r = 0; r < KS; ++r) {
#pragma unroll 1
                for (int n0 = 0; n0 < COB; n0 += 16) {
                    uint32_t v[16];
                    ptx::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + r * COB + n0, v);
                    ptx::tmem_ld_wait();
                    if (row < C::MROWS && ci < P.Cin) {
                        float *o = dst + ((size_t)(r * KS + s_shift_idx) * P.Cin + ci) * P.Cout + co0 + n0;
#pragma unroll
                        for (int jj = 0; jj < 16; jj += 4)
                            if (co0 + n0 + jj < P.Cout)
                                *reinterpret_cast<float4 *>(o + jj) = make_float4(__uint_as_float(v[jj]), __uint_as_float(v[jj + 1]),
                                                                                  __uint_as_float(v[jj + 2]), __uint_as_float(v[jj + 3]));
                    }
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (P.trace && blockIdx.x == 0 && threadIdx.x == 0) P.trace[5 * 32 + 1] = clock64();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// fixed-order sum of the split partials, scattered into the flat OIHW gradient (+ bias gradient).
// Block = 32 consecutive elements x 8 split groups: group g sums splits g, g+8, ... (independent, coalesced 128-byte
// loads), then the 8 group sums are added in a fixed order -> deterministic, and ~8x more loads in flight than one
// thread per element walking all splits.
__global__ void __launch_bounds__(256) tc_wgrad_reduce_kernel(const float *__restrict__ scratch, int S, int Cin, int Cout, int KK,
                                                              int cin_real, int cout_real, float *__restrict__ dw_oihw,
                                                              float *__restrict__ dbias, int accumulate) {
    pdl_prologue();
    __shared__ float part[8][32];
    const int64_t per = (int64_t)KK * Cin * Cout + Cout;
    const int ex = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int64_t e = (int64_t)blockIdx.x * 32 + ex;
    float acc = 0.f;
    if (e < per) {
#pragma unroll 4
        for (int k = g; k < S; k += 8) acc += scratch[(int64_t)k * per + e];
    }
    part[g][ex] = acc;
    __syncthreads();
    if (g != 0 || e >= per) return;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += part[k][ex];
    if (e < (int64_t)KK * Cin * Cout) {
        const int co = (int)(e % Cout), ci = (int)((e / Cout) % Cin), tap = (int)(e / ((int64_t)Cout * Cin));
        if (co < cout_real && ci < cin_real) {           // padded channels carry no parameter
            float *d = dw_oihw + ((int64_t)co * cin_real + ci) * KK + tap;
            *d = accumulate ? *d + s : s;
        }
    } else if (dbias && e - (int64_t)KK * Cin * Cout < cout_real) {
        float *d = dbias + (e - (int64_t)KK * Cin * Cout);
        *d = accumulate ? *d + s : s;
    }
}

static void wg_shape(int ks, int N, int H, int W, int Cin, int Cout, bool two, int &NB, int &COB, int &MT, int &ci_blocks, int &co_blocks, int &S,
                     int &m_tiles) {
    NB = Cin == 16 ? 16 : 32;
    COB = std::min(Cout, two ? 64 : 128);      // two-source dY: a second 16 KB tile per 64 output channels must fit next to three pipeline stages
    ci_blocks = Cin / NB;
    co_blocks = Cout / COB;
    // wide tiles for the 16-channel layers on large images (TMA-issue bound otherwise; with 32 input channels the
    // wider stage leaves too few pipeline stages in shared memory and measures slower)
    MT = (ks == 3 && NB == 16 && COB <= 32 && W % (2 * kTW) == 0 && (int64_t)N * ((H + kTH - 1) / kTH) * (W / (2 * kTW)) >= 4 * kNumSMs) ? 2 : 1;
    m_tiles = N * ((H + kTH - 1) / kTH) * ((W + kTW * MT - 1) / (kTW * MT));
    S = std::max(1, std::min(tc_cta_cap(2) / (ci_blocks * co_blocks), m_tiles));
}

int64_t tc_wgrad_scratch_floats(int N, int H, int W, int Cin, int Cout, int KS) {
    int64_t need = 0;
    for (int two = 0; two < 2; ++two) {
        int NB, COB, MT, cib, cob, S, mt;
        wg_shape(KS, N, H, W, Cin, Cout, two != 0, NB, COB, MT, cib, cob, S, mt);
        need = std::max(need, (int64_t)S * ((int64_t)KS * KS * Cin * Cout + Cout));
    }
    return need;
}

template <int KS, int NB, int COB, int MT, bool TWO>
static int wg_launch(const CUtensorMap &mx, const CUtensorMap &md, const CUtensorMap &mr, const WgParams &P, cudaStream_t s) {
    using C = WgCfg<KS, NB, COB, MT, TWO>;
    static bool attr_set = false;
    if (!attr_set) {
        HPFG_CUDA_CHECK(cudaFuncSetAttribute(tc_wgrad_kernel<KS, NB, COB, MT, TWO>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr_set = true;
    }
    const int grid = P.ci_blocks * P.co_blocks * P.S;
    HPFG_CUDA_CHECK(launch_pdl(tc_wgrad_kernel<KS, NB, COB, MT, TWO>, grid, kWgThreads, C::SMEM_BYTES, s, mx, md, mr, P));
    HPFG_LAUNCH_CHECK();
    return HPFG_OK;
}

template <int KS>
static int wg_dispatch(int NB, int COB, int MT, bool two, const CUtensorMap &mx, const CUtensorMap &md, const CUtensorMap &mr, const WgParams &P,
                       cudaStream_t s) {
#define HPFG_WG_CASE(nb, cob)                                                               \
    if (NB == nb && COB == cob) {                                                           \
        if constexpr (KS == 3 && cob <= 64) {       /* only 3x3 convolutions are followed by a BatchNorm */ \
            if (two) {                                                                      \
                if constexpr (nb == 16 && cob <= 32) {                                      \
                    if (MT == 2) return wg_launch<KS, nb, cob, 2, true>(mx, md, mr, P, s);  \
                }                                                                           \
                return wg_launch<KS, nb, cob, 1, true>(mx, md, mr, P, s);                   \
            }                                                                               \
        }                                                                                   \
        if (two) break;                                                                     \
        if constexpr (KS == 3 && nb == 16 && cob <= 32) {                                   \
            if (MT == 2) return wg_launch<KS, nb, cob, 2, false>(mx, md, mr, P, s);         \
        }                                                                                   \
        return wg_launch<KS, nb, cob, 1, false>(mx, md, mr, P, s);                          \
    }
    do {
    HPFG_WG_CASE(16, 16) HPFG_WG_CASE(16, 32) HPFG_WG_CASE(16, 64) HPFG_WG_CASE(16, 128)
    HPFG_WG_CASE(32, 16) HPFG_WG_CASE(32, 32) HPFG_WG_CASE(32, 64) HPFG_WG_CASE(32, 128)
#undef HPFG_WG_CASE
    } while (0);
    set_error("tc wgrad: no kernel for NB=" + std::to_string(NB) + " COB=" + std::to_string(COB) + (two ? " (two-source dY)" : ""));
    return HPFG_ERR_UNSUPPORTED;
}

int tc_wgrad_run(int ks, int N, int H, int W, int Cin, int Cout, int cin_real, int cout_real, const void *x, LoadXform xf, const void *dy, float *scratch,
                 int64_t scratch_floats, float *dw_oihw, float *dbias, int accumulate, cudaStream_t s, const TcBwdFuse *fuse,
                 const TcWgradStreams *ws) {
    ProfScope _prof(PROF_WGRAD_TC, s);
    int NB, COB, MT;
    WgParams P{};
    const bool two = fuse && fuse->in_raw;
    wg_shape(ks, N, H, W, Cin, Cout, two, NB, COB, MT, P.ci_blocks, P.co_blocks, P.S, P.m_tiles);
    HPFG_REQUIRE(tc_wgrad_scratch_floats(N, H, W, Cin, Cout, ks) <= scratch_floats, "tc_wgrad: scratch too small");
    CUtensorMap mx, md, mr;
    HPFG_RETURN_IF(make_map_chunked(&mx, x, N, H, W, Cin, NB / 8, kTW * MT + ks - 1, kTH + ks - 1));
    HPFG_RETURN_IF(make_map_chunked(&md, dy, N, H, W, Cout, COB / 8, kTW * MT, kTH));
    mr = md;
    if (two) {
        HPFG_RETURN_IF(make_map_chunked(&mr, fuse->in_raw, N, H, W, Cout, COB / 8, kTW * MT, kTH));
        P.dsc = fuse->sc; P.dkb = fuse->kb; P.dkd = fuse->kd;
    }
    P.scale = xf.scale; P.shift = xf.shift;
    P.dropbits = reinterpret_cast<const uint8_t *>(xf.drop.bits); P.inv_keep = xf.drop.inv_keep;
    P.scratch = scratch;
    P.N = N; P.H = H; P.W = W; P.Cin = Cin; P.Cout = Cout;
    P.tiles_h = (H + kTH - 1) / kTH; P.tiles_w = (W + kTW * MT - 1) / (kTW * MT);
    static long long *trace_buf = nullptr;
    const bool want_trace = getenv("HPFG_WG_TRACE") != nullptr;
    if (want_trace && !trace_buf) cudaMalloc(&trace_buf, 6 * 32 * 8);
    if (want_trace) cudaMemsetAsync(trace_buf, 0, 6 * 32 * 8, s);
    P.trace = want_trace ? trace_buf : nullptr;
    HPFG_RETURN_IF(ks == 3 ? wg_dispatch<3>(NB, COB, MT, two, mx, md, mr, P, s) : wg_dispatch<1>(NB, COB, MT, two, mx, md, mr, P, s));
    if (want_trace && getenv("HPFG_WG_TRACE_DUMP")) {
        long long h[6 * 32];
        cudaStreamSynchronize(s);
        cudaMemcpy(h, trace_buf, sizeof(h), cudaMemcpyDeviceToHost);
        const long long t0 = h[5 * 32 + 2];
        const char *names[5] = {"tma:slot-free", "xf:tile-landed", "xf:done", "mma:operands-ready", "mma:issued"};
        printf("wgrad trace (CTA 0, cycles since setup): NB=%d COB=%d MT=%d S=%d grid=%d; epilogue starts %lld, kernel ends %lld\n", NB, COB, MT, P.S,
               P.ci_blocks * P.co_blocks * P.S, h[5 * 32] - t0, h[5 * 32 + 1] - t0);
        for (int r = 0; r < 5; ++r) {
            printf("%-20s", names[r]);
            for (int i = 0; i < 12; ++i) printf(" %6lld", h[r * 32 + i] ? h[r * 32 + i] - t0 : -1);
            printf("\n");
        }
        fflush(stdout);
    }
    const int64_t per = (int64_t)ks * ks * Cin * Cout + Cout;
    const int blocks = (int)((per + 31) / 32);
    if (ws && ws->reduce_stream) {     // reduction on its own stream, ordered by events (plain launch: its predecessor is on another stream)
        HPFG_CUDA_CHECK(cudaEventRecord(ws->ev_partials, s));
        HPFG_CUDA_CHECK(cudaStreamWaitEvent(ws->reduce_stream, ws->ev_partials, 0));
        HPFG_CUDA_CHECK(launch_plain(tc_wgrad_reduce_kernel, blocks, 256, 0, ws->reduce_stream, (const float *)scratch, P.S, Cin, Cout, ks * ks, cin_real,
                                     cout_real, dw_oihw, dbias, accumulate));
        HPFG_LAUNCH_CHECK();
        HPFG_CUDA_CHECK(cudaEventRecord(ws->ev_reduced, ws->reduce_stream));
        return HPFG_OK;
    }
    HPFG_CUDA_CHECK(launch_pdl(tc_wgrad_reduce_kernel, blocks, 256, 0, s, scratch, P.S, Cin, Cout, ks * ks, cin_real, cout_real, dw_oihw, dbias, accumulate));
    HPFG_LAUNCH_CHECK();
    return HPFG_OK;
}

}  // namespace hpfg

namespace hpfg {
// one thread per pixel: only the C real channels are loaded (coalesced along the pixel index of each NCHW plane), the
// rest of the 16-channel bf16 NHWC pixel is zero (ncu: the predicated 16-channel version was issue-bound, 220
// instructions per pixel)
template <int CMAX>
__global__ void __launch_bounds__(256) pad_to_nhwc16_kernel(const float *__restrict__ src, uint4 *__restrict__ dst, int N, int C, int H, int W, FastDiv dHW) {
    pdl_prologue();
    const uint32_t HW = dHW.d;
    const uint32_t total = (uint32_t)N * HW;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        uint32_t n, pix;
        fast_divmod(i, dHW, n, pix);
        const float *sp = src + (size_t)n * C * HW + pix;
        float v[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] = 0.f;
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
            if (c < C) v[c] = __ldg(sp + (size_t)c * HW);
        dst[2 * (size_t)i] = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
        dst[2 * (size_t)i + 1] = make_uint4(pack2(v[8], v[9]), pack2(v[10], v[11]), pack2(v[12], v[13]), pack2(v[14], v[15]));
    }
}
int pad_to_nhwc16(const float *src_nchw, void *dst, int N, int C, int H, int W, cudaStream_t s) {
    ProfScope _prof(PROF_GLUE, s);
    const int64_t total = (int64_t)N * H * W;
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)kNumSMs * 16);
    if (C <= 4)
        HPFG_CUDA_CHECK(launch_pdl(pad_to_nhwc16_kernel<4>, blocks, 256, 0, s, src_nchw, reinterpret_cast<uint4 *>(dst), N, C, H, W, make_fastdiv((uint32_t)(H * W))));
    else
        HPFG_CUDA_CHECK(launch_pdl(pad_to_nhwc16_kernel<16>, blocks, 256, 0, s, src_nchw, reinterpret_cast<uint4 *>(dst), N, C, H, W, make_fastdiv((uint32_t)(H * W))));
    HPFG_LAUNCH_CHECK();
    return HPFG_OK;
}
}  // namespace hpfg

// ---- layer-isolated test hook -------------------------------------------------------------------------------
using namespace hpfg;
extern "C" int hpfg_wgrad_tc_debug(int N, int H, int W, int cin, int cout, int ks, const void *x_bf16_nhwc,
                                   const void *dy_bf16_nhwc, const float *scale, const float *shift, float *dw_oihw,
                                   float *dbias, void *stream) {
    HPFG_REQUIRE(cin % 16 == 0 && cout % 16 == 0 && (ks == 1 || ks == 3), "hpfg_wgrad_tc_debug: unsupported shape");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t nf = tc_wgrad_scratch_floats(N, H, W, cin, cout, ks);
    float *scratch = nullptr;
    HPFG_CUDA_CHECK(cudaMalloc(&scratch, (size_t)nf * 4));
    LoadXform xf{};
    xf.scale = scale; xf.shift = shift; xf.drop.bits = nullptr; xf.drop.inv_keep = 1.f;
    const int rc = tc_wgrad_run(ks, N, H, W, cin, cout, cin, cout, x_bf16_nhwc, xf, dy_bf16_nhwc, scratch, nf, dw_oihw, dbias, 0, s, nullptr);
    cudaStreamSynchronize(s);
    cudaFree(scratch);
    if (rc == HPFG_OK) HPFG_CUDA_CHECK(cudaGetLastError());
    return rc;
}

extern "C" int hpfg_wgrad_tc_fused_debug(int N, int H, int W, int cin, int cout, int ks, const void *x_bf16_nhwc, const void *g_bf16_nhwc,
                                         const void *raw_bf16_nhwc, const float *sc, const float *kb, const float *kd, const float *scale,
                                         const float *shift, float *dw_oihw, float *dbias, void *stream) {
    HPFG_REQUIRE(cin % 16 == 0 && cout % 16 == 0 && ks == 3, "hpfg_wgrad_tc_fused_debug: unsupported shape");
    HPFG_REQUIRE(x_bf16_nhwc && g_bf16_nhwc && raw_bf16_nhwc && sc && kb && kd, "hpfg_wgrad_tc_fused_debug: null argument");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t nf = tc_wgrad_scratch_floats(N, H, W, cin, cout, ks);
    float *scratch = nullptr;
    HPFG_CUDA_CHECK(cudaMalloc(&scratch, (size_t)nf * 4));
    LoadXform xf{};
    xf.scale = scale; xf.shift = shift; xf.drop.bits = nullptr; xf.drop.inv_keep = 1.f;
    TcBwdFuse f;
    f.in_raw = raw_bf16_nhwc; f.sc = sc; f.kb = kb; f.kd = kd;
    const int rc = tc_wgrad_run(ks, N, H, W, cin, cout, cin, cout, x_bf16_nhwc, xf, g_bf16_nhwc, scratch, nf, dw_oihw, dbias, 0, s, &f);
    cudaStreamSynchronize(s);
    cudaFree(scratch);
    if (rc == HPFG_OK) HPFG_CUDA_CHECK(cudaGetLastError());
    return rc;
}

// ---- layer micro-benchmark hook (profiles/layer_bench.py): times `iters` back-to-back launches of ONE tensor-core
// convolution (op 0 fprop with fused loader + BN-stat epilogue, 1 dgrad, 2 wgrad) on internally allocated
// buffers with CUDA events on `stream`; returns the average milliseconds per launch.
namespace hpfg {
int tc_run_bench(int op, int ks, int N, int H, int W, int cin, int cout, const void *in, void *out, const float *w, const float *scale,
                 const float *shift, float *stats, cudaStream_t s, int fuse_mode, const void *aux);   // conv_tc.cu
}
__global__ void fill_pattern_bf16(__nv_bfloat16 *p, size_t n, float scale) {
    pdl_prologue();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = __float2bfloat16(scale * (float)((int)((i * 2654435761u) >> 24) - 128) / 128.f);
}
extern "C" int hpfg_conv_tc_bench(int op, int N, int H, int W, int cin, int cout, int ks, int iters, float *ms_out_host, void *stream) {
    // op 3 / 4 / 5: data gradient with the two-source loader / the GSTAT epilogue / both; op 6: weight gradient with the two-source dY
    HPFG_REQUIRE(op >= 0 && op <= 6 && iters > 0 && ms_out_host, "hpfg_conv_tc_bench: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t px = (size_t)N * H * W;
    __nv_bfloat16 *a = nullptr, *b = nullptr, *aux = nullptr;
    float *w = nullptr, *sc = nullptr, *stats = nullptr, *scratch = nullptr, *dw = nullptr;
    const int64_t nf = tc_wgrad_scratch_floats(N, H, W, cin, cout, ks);
    HPFG_CUDA_CHECK(cudaMalloc(&a, px * std::max(cin, cout) * 2));
    HPFG_CUDA_CHECK(cudaMalloc(&b, px * std::max(cin, cout) * 2));
    HPFG_CUDA_CHECK(cudaMalloc(&aux, px * std::max(cin, cout) * 2));
    HPFG_CUDA_CHECK(cudaMemsetAsync(aux, 0, px * std::max(cin, cout) * 2, s));
    HPFG_CUDA_CHECK(cudaMalloc(&w, (size_t)cin * cout * ks * ks * 4));
    HPFG_CUDA_CHECK(cudaMalloc(&sc, 2 * 256 * 4));
    HPFG_CUDA_CHECK(cudaMalloc(&stats, (size_t)kNumSMs * 2 * 256 * 4 * 4));
    HPFG_CUDA_CHECK(cudaMalloc(&scratch, (size_t)nf * 4));
    HPFG_CUDA_CHECK(cudaMalloc(&dw, ((size_t)cin * cout * ks * ks + cout) * 4));
    HPFG_CUDA_CHECK(launch_pdl(fill_pattern_bf16, 1024, 256, 0, s, a, px * std::max(cin, cout), 1.f));
    HPFG_CUDA_CHECK(launch_pdl(fill_pattern_bf16, 1024, 256, 0, s, b, px * std::max(cin, cout), 1.f));
    HPFG_CUDA_CHECK(cudaMemsetAsync(w, 0, (size_t)cin * cout * ks * ks * 4, s));
    HPFG_CUDA_CHECK(cudaMemsetAsync(sc, 0, 2 * 256 * 4, s));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    LoadXform xf{};
    xf.scale = sc; xf.shift = sc + 256; xf.drop.bits = nullptr; xf.drop.inv_keep = 1.f;
    int rc = HPFG_OK;
    for (int i = -2; i < iters && rc == HPFG_OK; ++i) {
        if (i == 0) cudaEventRecord(e0, s);
        if (op == 2 || op == 6) {
            TcBwdFuse f;
            f.in_raw = aux; f.sc = sc; f.kb = sc; f.kd = sc;
            rc = tc_wgrad_run(ks, N, H, W, cin, cout, cin, cout, a, xf, b, scratch, nf, dw, dw + (size_t)cin * cout * ks * ks, 0, s, op == 6 ? &f : nullptr);
        } else if (op >= 3) {
            rc = tc_run_bench(1, ks, N, H, W, cin, cout, a, b, w, sc, sc + 256, stats, s, op - 2, aux);
        } else {
            rc = tc_run_bench(op, ks, N, H, W, cin, cout, a, b, w, op == 0 ? sc : nullptr, op == 0 ? sc + 256 : nullptr, op == 0 ? stats : nullptr, s, 0, nullptr);
        }
    }
    cudaEventRecord(e1, s);
    cudaStreamSynchronize(s);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    *ms_out_host = ms / iters;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(a); cudaFree(b); cudaFree(aux); cudaFree(w); cudaFree(sc); cudaFree(stats); cudaFree(scratch); cudaFree(dw);
    if (rc == HPFG_OK) HPFG_CUDA_CHECK(cudaGetLastError());
    return rc;
}
