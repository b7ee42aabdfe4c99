// Tensor-core implicit-GEMM convolution kernel for the bf16 path (sm_100a): tcgen05.mma with TMEM accumulators,
// TMA-staged NHWC bf16 halo tiles, warp-specialised persistent CTAs.  (Host side: conv_tc.cu.)
//
//   D[128 pixels, BN out-channels] = sum over taps (r,s) and input-channel steps of
//        A_tap[128 pixels, 16 ch] (smem, K-major)  x  W_tap[BN, 16 ch] (smem, K-major)
//
// One pipeline stage = MT side-by-side UMMA tiles of 16 rows x 8 columns of output pixels of one image.  The
// (16+2)x(8*MT+2) input halo of a channel chunk is fetched ONCE by TMA through a 5-D "chunked" tensor map
// (8 ch, W, H, C/8, N: the chunk dimension has a 16-byte stride), so the bytes land directly in the UMMA operand
// order [chunk of 8 ch][halo pixel][8 ch]; out-of-bounds = zero = the conv padding.  When the producer's
// BatchNorm + LeakyReLU + dropout are fused into this consumer's loader (model/unet.py:17-25), eight transform
// warps apply them in place (XF = 1: affine + LeakyReLU, XF = 2: + dropout keep bits).  The nine taps are nine UMMA
// descriptors into that ONE staged tile: start address shifted by (r*HW+s)*16 bytes, SBO = one halo row.
// Epilogue (4 warps): tcgen05.ld -> per-channel sum / sum-of-squares for train-mode BatchNorm -> bf16 NHWC store
// (or fp32 NCHW logits + bias when NCHW).  fprop, dgrad (flipped/transposed packed weights) and the 1x1
// convolutions of the up-blocks all run through this kernel.
//
// Code size matters here: five roles execute disjoint code concurrently on one SM, so everything that is not a
// compile-time necessity is a rolled loop (ncu: the single MMA-issuing lane was stalled on instruction fetch
// when its 36 MMAs per stage were fully unrolled next to a 54 KB epilogue).
#pragma once
#include "tc_common.cuh"

namespace hpfg {

// 16 warps: 0-3 TMA producer / MMA issuer / TMEM allocator / idle, then the loader-transform warps, then the epilogue warps.
// Narrow layers (BN <= 32) are transform-heavy (612-pixel halo tiles): 8 transform + 4 epilogue warps.  Wide layers (BN >= 64)
// have small halo tiles but a 64..128-column epilogue with BatchNorm statistics that was the pipeline bottleneck with four
// warps (per-role trace: 5.5 k cycles per 128 x 128 tile): 4 transform + 8 epilogue warps, two per TMEM lane quarter, each
// taking half of the columns.
constexpr int kTcThreads = 512;
constexpr int kMaxStages = 12;
#ifndef HPFG_TC_SMEM_KB
#define HPFG_TC_SMEM_KB 216
#endif
#ifndef HPFG_TC_LB_THREADS
#define HPFG_TC_LB_THREADS 512      // launch bound used for register allocation (640 -> at most 96 registers per thread)
#endif
constexpr int kSmemBudget = HPFG_TC_SMEM_KB * 1024;

// Loader transforms (XF): 0 none; 1 BatchNorm affine + LeakyReLU of the producer; 2 = 1 + dropout keep bits;
// 3 = BatchNorm BACKWARD of the layer this data gradient enters through: the operand is built from TWO staged tiles,
//     draw = sc*g + kb*raw + kd  (g = dact * leaky' * dropout', raw = the layer's saved conv output; glue.cuh BnState).
// Epilogues (EPI): 0 bf16 NHWC store (+ BatchNorm statistics partials when P.stats); 1 fp32 NCHW logits + bias;
// 2 = "GSTAT": the accumulator is the gradient wrt the ACTIVATED output of a BatchNorm layer; the epilogue reads that
//     layer's raw tensor at the same pixel, stores g = dact * leaky'(bn(raw)) * dropout' and accumulates the two
//     BatchNorm-backward sums (sum g | sum g*raw) into the statistics partials -- bn_bwd pass 0 never runs.
template <int KS, int KC, int BN, bool RES, int MT, bool TWO = false>
struct TcCfg {
    static constexpr int PAD = KS / 2, KK = KS * KS;
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
    static constexpr int OFF_R = OFF_B + (RES ? 0 : al(B_BYTES));      // second source tile (TWO: the raw tensor of XF = 3)
    static constexpr int STAGE_BYTES = OFF_R + (TWO ? al(OP_BYTES) : 0);
    static constexpr int RESB_BYTES = RES ? al(B_BYTES) : 0;
    static constexpr int FIXED_BYTES = 1024 /*barriers*/ + 5 * 256 * 4 /*scale,shift,kd | epilogue scale,shift*/ + 4 * 2 * BN * 4 /*stat partials*/;
    static constexpr int STAGES_RAW = (kSmemBudget - FIXED_BYTES - RESB_BYTES) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > kMaxStages ? kMaxStages : STAGES_RAW;
    static constexpr int SMEM_BYTES = RESB_BYTES + STAGES * STAGE_BYTES + FIXED_BYTES + 1024 /*alignment slack*/;
    static constexpr int ACC_COLS = MT * BN;                       // accumulator columns per TMEM buffer
    static constexpr int NACC = (4 * ACC_COLS <= 512) ? 4 : 2;      // accumulator buffers: MMA runs up to NACC-1 stages ahead of the epilogue
    static constexpr int XFW = BN >= 64 ? 4 : 8;                    // loader-transform warps
    static constexpr int EPW = 12 - XFW;                            // epilogue warps (EPW / 4 per TMEM lane quarter)
    static constexpr int TMEM_COLS = (NACC * ACC_COLS <= 32) ? 32 : (NACC * ACC_COLS <= 64) ? 64 : (NACC * ACC_COLS <= 128) ? 128 : (NACC * ACC_COLS <= 256) ? 256 : 512;
    static_assert(STAGES >= 2, "need at least a double buffer");
    static_assert(NACC * ACC_COLS <= 512, "accumulators exceed TMEM");
};

struct TcConvParams {
    const bf16 *bpk;          // packed weights: [n_block][k_chunk][tap][KC/8][BN][8]
    bf16 *out;                // NHWC [N,H,W,Cout]
    const float *bias;        // Cout floats or nullptr
    const float *scale, *shift;   // per input channel (producer's fused BN affine); required when XF > 0.  XF == 3: sc, kb
    const float *kd;          // XF == 3: the third BatchNorm-backward constant per input channel
    const uint8_t *dropbits;  // producer's dropout keep bits, NHWC bit order; required when XF == 2
    // EPI == 2 (GSTAT): the BatchNorm layer whose activated-output gradient this kernel produces
    const bf16 *gs_raw;       // its raw conv output, NHWC [N,H,W,Cout]
    const float *gs_scale, *gs_shift;   // its fused forward affine (z = raw*scale + shift)
    const uint16_t *gs_dropbits;        // its dropout keep bits (16 channels per halfword) or nullptr
    float gs_inv_keep;
    float inv_keep;
    float *stats;             // [gridDim.x][2*Cout] per-CTA partial sums (sum | sum of squares) or nullptr
    float *out_nchw;          // NCHW kernels: fp32 [N,out_c_real,H,W] (+bias) instead of bf16 NHWC (out_conv logits)
    int out_c_real;
    int N, H, W, Cin, Cout, tiles_h, tiles_w, m_tiles, n_blocks, k_chunks;
    int max_ctas;             // grid cap (a multiple of n_blocks): persistent CTAs of this launch
    long long *trace;         // optional (micro-benchmark only): CTA 0 records clock64 per role and stage, [role][64]
    int dbg;                  // bottleneck-isolation switches (env HPFG_TC_DBG, profiles/ only): 1 no MMA, 2 no stores, 4 no stats, 8 no TMA
};

// Reduce 16 per-lane values over the 32 lanes of a warp with 16 shuffles (recursive halving); afterwards every
// lane holds the full column sum of column col16(lane).
__device__ __forceinline__ int col16(int lane) { return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1); }
__device__ __forceinline__ float butterfly16(const float *v, int lane) {
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

#define HPFG_TRACE(role, idx) do { if (P.trace && blockIdx.x == 0 && lane == 0 && (idx) < 64) P.trace[(role) * 64 + (idx)] = clock64(); } while (0)

// MMA issuer role (warp 1; warp-uniform, one elected lane issues).  A separate non-inlined function so that ptxas
// allocates its (uniform) registers independently of the other roles: inlined, the 2 x 36 smem descriptors of a
// stage were spilled / recycled through two uniform-register pairs, and every rewrite of a pair stalled until the
// previous UTCHMMA using it had been dispatched (measured 50 instead of 39 cycles per MMA).
template <int KS, int KC, int BN, bool RES, int MT, int XF>
__device__ __forceinline__ void tc_mma_role(uint32_t bar_full, uint32_t bar_xf, uint32_t bar_empty, uint32_t bar_tfull, uint32_t bar_tempty,
                                         uint32_t stage_u32, uint32_t res_u32, uint32_t tmem_base, int n_work, int k_chunks, int dbg,
                                         long long *trace) {
    using C = TcCfg<KS, KC, BN, RES, MT, XF == 3>;
    const int lane = threadIdx.x & 31;
    {
        constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, BN, 0, 0);
        // descriptor words: hi = SBO | version, lo = (addr >> 4) | LBO << 16; per-MMA offsets are compile-time adds
        constexpr uint32_t a_hi = (uint32_t)((C::HW * 16) >> 4) | (1u << 14), b_hi = (uint32_t)(128 >> 4) | (1u << 14);
        const uint32_t a_lo0 = ((stage_u32 >> 4) & 0x3FFFu) | ((uint32_t)(C::CH_STRIDE >> 4) << 16);
        const uint32_t b_lo0 = (((RES ? res_u32 : stage_u32 + C::OFF_B) >> 4) & 0x3FFFu) | ((uint32_t)((BN * 16) >> 4) << 16);
        // Barrier polls issued after a stage's MMAs would queue behind them in the warp's in-order MIO queue (measured:
        // ~400 cycles per poll), so the NEXT stage's barriers are polled before this stage's MMAs are pushed; the
        // blocking waits only run when that early poll failed.
        int stage = 0, phase = 0, kc = 0, it = 0;
        const int total = n_work * k_chunks;
        bool ready = false;
#pragma unroll 1
        for (int f = 0; f < total; ++f) {
            const int acc = it % C::NACC;
            if (!ready) {
                if (kc == 0) ptx::mbar_wait(bar_tempty + 8 * acc, ((it / C::NACC) & 1) ^ 1, 2);
                ptx::mbar_wait(bar_full + 8 * stage, phase, 3);
                if (XF > 0) ptx::mbar_wait(bar_xf + 8 * stage, phase, 4);
            }
            ptx::tc_fence_after();
            if (trace && blockIdx.x == 0 && lane == 0 && f < 64) trace[1 * 64 + f] = clock64();
            int nstage = stage + 1, nphase = phase, nkc = kc + 1, nit = it;
            if (nstage == C::STAGES) { nstage = 0; nphase ^= 1; }
            if (nkc == k_chunks) { nkc = 0; ++nit; }
            {
                bool ok = f + 1 < total && ptx::mbar_try_wait(bar_full + 8 * nstage, nphase);
                if (XF > 0) ok = ok && ptx::mbar_try_wait(bar_xf + 8 * nstage, nphase);
                if (nkc == 0) ok = ok && ptx::mbar_try_wait(bar_tempty + 8 * (nit % C::NACC), ((nit / C::NACC) & 1) ^ 1);
                ready = __all_sync(0xffffffffu, ok);
            }
            const uint32_t d_tmem = tmem_base + acc * C::ACC_COLS;
            const uint32_t a_lo = a_lo0 + stage * (C::STAGE_BYTES >> 4);
            const uint32_t b_lo = RES ? b_lo0 : b_lo0 + stage * (C::STAGE_BYTES >> 4);
            if (ptx::elect_one()) {
                if (!(dbg & 1)) {
                    // fully unrolled: a rolled loop stalls ~75 cycles at every back-edge until the queued UTCHMMAs have
                    // consumed their uniform-register operands (measured 47 instead of 39 cycles per MMA)
#pragma unroll 1
                    for (int j = 0; j < MT; ++j) {             // UMMA tile j = output columns 8j..8j+7 of the stage
                        const uint32_t a_j = a_lo + (uint32_t)j * (uint32_t)((kTW * 16) >> 4), d_j = d_tmem + j * BN;
#pragma unroll
                        for (int tap = 0; tap < C::KK; ++tap) {
#pragma unroll
                            for (int kk = 0; kk < KC / 16; ++kk) {
                                const uint32_t ao = (uint32_t)((2 * kk * C::CH_STRIDE + ((tap / KS) * C::HW + (tap % KS)) * 16) >> 4);
                                const uint32_t bo = (uint32_t)((tap * C::B_TAP_BYTES + 2 * kk * BN * 16) >> 4);
                                const uint64_t ad = ((uint64_t)a_hi << 32) | (uint64_t)(a_j + ao);
                                const uint64_t bd = ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + bo);
                                ptx::umma_bf16(d_j, ad, bd, idesc, (kc | tap | kk) != 0);
                            }
                        }
                    }
                }
                ptx::umma_commit(bar_empty + 8 * stage);      // smem slot reusable once these MMAs retire
                if (nkc == 0) ptx::umma_commit(bar_tfull + 8 * acc);   // accumulator complete -> epilogue
            }
            __syncwarp();
            if (trace && blockIdx.x == 0 && lane == 0 && f < 64) trace[2 * 64 + f] = clock64();
            stage = nstage; phase = nphase; kc = nkc; it = nit;
        }
    }
}

// warps: 0 TMA producer, 1 MMA issuer, 2 TMEM alloc, 3 idle, 4-11 transform (XF > 0 only), 12-15 epilogue.
template <int KS, int KC, int BN, bool RES, int MT, int XF, int EPI>
__global__ void __launch_bounds__(HPFG_TC_LB_THREADS, 1) tc_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmR,
                                                                const TcConvParams P) {
    using C = TcCfg<KS, KC, BN, RES, MT, XF == 3>;
    constexpr bool NCHW = EPI == 1, GSTAT = EPI == 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *res_b = smem;                                         // resident weights (RES only)
    uint8_t *stage_base = smem + C::RESB_BYTES;
    uint8_t *fixed = stage_base + C::STAGES * C::STAGE_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(fixed);          // full[S] xf[S] empty[S] tfull[NACC] tempty[NACC]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(fixed + 512);
    float *s_scale = reinterpret_cast<float *>(fixed + 1024);
    float *s_shift = s_scale + 256;
    float *s_kd = s_shift + 256;
    float *s_gsc = s_kd + 256;                                     // GSTAT: forward affine of the output-side BatchNorm
    float *s_gsh = s_gsc + 256;
    float *s_part = s_gsh + 256;                                   // [4 warps][2*BN]

    pdl_launch_dependents();
    // broadcast from lane 0: tells the compiler the warp index is warp-uniform, so role branches and everything
    // loop-carried inside them (stage counters, descriptor bases) can live in the uniform datapath
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t bar_full = ptx::smem_u32(bars), bar_xf = bar_full + 8 * C::STAGES, bar_empty = bar_xf + 8 * C::STAGES;
    const uint32_t bar_tfull = bar_empty + 8 * C::STAGES, bar_tempty = bar_tfull + 8 * C::NACC;
    const uint32_t stage_u32 = ptx::smem_u32(stage_base);

    // this CTA's work items: fixed n-block, m-tiles mt0, mt0+mstep, ... (n_blocks divides the grid)
    const int total_work = P.m_tiles * P.n_blocks;
    const int nb = blockIdx.x % P.n_blocks, mt0 = blockIdx.x / P.n_blocks, mstep = gridDim.x / P.n_blocks;
    const int n_work = (total_work - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    // The packed weights were written several launches ago (tc_pack_kernel at the start of the pass), so the first
    // ring of weight stages is requested BEFORE griddepcontrol.wait: their latency overlaps the previous kernel's tail.
    const int total_flat = n_work * P.k_chunks;
    const int pre = RES ? (total_flat > 0 ? 1 : 0) : (total_flat < C::STAGES ? total_flat : C::STAGES);
    const uint32_t op_bytes = (P.dbg & 8) ? 0u : (uint32_t)(C::OP_BYTES * (XF == 3 ? 2 : 1));

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            ptx::mbar_init(bar_full + 8 * s, 1);
            ptx::mbar_init(bar_xf + 8 * s, C::XFW);
            ptx::mbar_init(bar_empty + 8 * s, 1);
        }
        for (int a = 0; a < C::NACC; ++a) {
            ptx::mbar_init(bar_tfull + 8 * a, 1);
            ptx::mbar_init(bar_tempty + 8 * a, C::EPW);
        }
        ptx::fence_barrier_init();
        ptx::prefetch_tensormap(&tmA);
        if (XF == 3) ptx::prefetch_tensormap(&tmR);
        const bf16 *bsrc0 = P.bpk + (size_t)nb * P.k_chunks * (C::B_BYTES / 2);
        for (int f = 0; f < pre; ++f) {
            const uint32_t fb = bar_full + 8 * f;
            ptx::mbar_expect_tx(fb, op_bytes + C::B_BYTES);
            if (RES) ptx::bulk_load(ptx::smem_u32(res_b), P.bpk, C::B_BYTES, fb);
            else ptx::bulk_load(stage_u32 + f * C::STAGE_BYTES + C::OFF_B, bsrc0 + (size_t)(f % P.k_chunks) * (C::B_BYTES / 2), C::B_BYTES, fb);
        }
    }
    if (warp == 2) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), C::TMEM_COLS);
    pdl_wait();       // everything above is private setup; below reads what the previous kernel wrote
    if (XF > 0)
        for (int i = threadIdx.x; i < P.Cin; i += blockDim.x) {
            s_scale[i] = P.scale[i];
            s_shift[i] = P.shift[i];
            if (XF == 3) s_kd[i] = P.kd[i];
        }
    if (GSTAT)
        for (int i = threadIdx.x; i < P.Cout; i += blockDim.x) { s_gsc[i] = P.gs_scale[i]; s_gsh[i] = P.gs_shift[i]; }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (P.trace && threadIdx.x == 0 && blockIdx.x == 0) P.trace[5 * 64] = clock64();


    if (warp == 0) {
        // ================================================================= TMA producer (warp-uniform, one lane issues)
        TileIter ti;
        ti.init(mt0, mstep, P.tiles_h, P.tiles_w);
        int stage = 0, phase = 0, f = 0;
        const bf16 *bsrc = P.bpk + (size_t)nb * P.k_chunks * (C::B_BYTES / 2);
#pragma unroll 1
        for (int it = 0; it < n_work; ++it) {
            const int h0 = ti.th * kTH - C::PAD, w0 = ti.tw * C::TWP - C::PAD;
#pragma unroll 1
            for (int kc = 0; kc < P.k_chunks; ++kc, ++f) {
                ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1, 1);
                HPFG_TRACE(0, it);
                const uint32_t sb = stage_u32 + stage * C::STAGE_BYTES, fb = bar_full + 8 * stage;
                if (ptx::elect_one()) {
                    if (f >= pre) {      // (the first `pre` stages already carry their expect_tx and weight load)
                        if (RES) {       // resident weights: requested once, before the loop
                            ptx::mbar_expect_tx(fb, op_bytes);
                        } else {
                            ptx::mbar_expect_tx(fb, op_bytes + C::B_BYTES);
                            ptx::bulk_load(sb + C::OFF_B, bsrc + (size_t)kc * (C::B_BYTES / 2), C::B_BYTES, fb);
                        }
                    }
                    if (!(P.dbg & 8)) {
                        ptx::tma_load_5d(sb, &tmA, fb, 0, w0, h0, kc * C::NCH, ti.n_img);
                        if (XF == 3) ptx::tma_load_5d(sb + C::OFF_R, &tmR, fb, 0, w0, h0, kc * C::NCH, ti.n_img);
                    }
                }
                __syncwarp();
                if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
            ti.next(P.tiles_h, P.tiles_w);
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer (separate function: own register allocation)
        tc_mma_role<KS, KC, BN, RES, MT, XF>(bar_full, bar_xf, bar_empty, bar_tfull, bar_tempty, stage_u32, ptx::smem_u32(res_b),
                                             tmem_base, n_work, P.k_chunks, P.dbg, P.trace);
    } else if (warp >= 4 && warp < 4 + C::XFW) {
        // ================================================================= loader-transform warps (in place)
        if (XF > 0) {
            // warp -> one 8-channel chunk (scale/shift live in registers), lanes -> 32 consecutive halo pixels
            // (512 contiguous bytes: conflict-free 128-bit shared accesses)
            constexpr int G = C::XFW / C::NCH, STEP = 32 * G, ITERS = (C::NPIX + STEP - 1) / STEP;
            static_assert(G >= 1, "more channel chunks per stage than transform warps");
            const int xw = warp - 4, c = xw % C::NCH, p0 = (xw / C::NCH) * 32 + lane;
            // halo coordinates of this thread's pixels are the same for every tile: (row << 8) | col
            int hrc[ITERS];
#pragma unroll
            for (int k = 0; k < ITERS; ++k) {
                const int p = p0 + k * STEP;
                hrc[k] = ((p / C::HW) << 8) | (p % C::HW);
            }
            TileIter ti;
            ti.init(mt0, mstep, P.tiles_h, P.tiles_w);
            int stage = 0, phase = 0;
#pragma unroll 1
            for (int it = 0; it < n_work; ++it) {
                const int h0 = ti.th * kTH - C::PAD, w0 = ti.tw * C::TWP - C::PAD;
                // only border tiles contain out-of-image halo pixels (conv zero padding applies AFTER the activation)
                const bool border = h0 < 0 || w0 < 0 || h0 + C::HH > P.H || w0 + C::HW > P.W;
                const size_t img_row0 = (size_t)ti.n_img * P.H;
#pragma unroll 1
                for (int kc = 0; kc < P.k_chunks; ++kc) {
                    const int ch = kc * KC + c * 8;
                    float scl[8], shf[8], kdv[XF == 3 ? 8 : 1];
                    {
                        const float4 a0 = *reinterpret_cast<const float4 *>(s_scale + ch), a1 = *reinterpret_cast<const float4 *>(s_scale + ch + 4);
                        const float4 b0 = *reinterpret_cast<const float4 *>(s_shift + ch), b1 = *reinterpret_cast<const float4 *>(s_shift + ch + 4);
                        scl[0] = a0.x; scl[1] = a0.y; scl[2] = a0.z; scl[3] = a0.w; scl[4] = a1.x; scl[5] = a1.y; scl[6] = a1.z; scl[7] = a1.w;
                        shf[0] = b0.x; shf[1] = b0.y; shf[2] = b0.z; shf[3] = b0.w; shf[4] = b1.x; shf[5] = b1.y; shf[6] = b1.z; shf[7] = b1.w;
                        if (XF == 3) {
                            const float4 d0 = *reinterpret_cast<const float4 *>(s_kd + ch), d1 = *reinterpret_cast<const float4 *>(s_kd + ch + 4);
                            kdv[0] = d0.x; kdv[1] = d0.y; kdv[2] = d0.z; kdv[3] = d0.w; kdv[4] = d1.x; kdv[5] = d1.y; kdv[6] = d1.z; kdv[7] = d1.w;
                        }
                    }
                    // dropout keep bytes come from global memory: fetch them all BEFORE waiting for the stage, so their latency
                    // hides behind the TMA wait instead of sitting inside the per-item dependency chain
                    uint32_t keepv[XF == 2 ? ITERS : 1];
                    if (XF == 2) {
#pragma unroll
                        for (int k = 0; k < ITERS; ++k) {
                            const int gh = h0 + (hrc[k] >> 8), gw = w0 + (hrc[k] & 255);
                            const bool ok = ((k + 1) * STEP <= C::NPIX || p0 + k * STEP < C::NPIX) &&
                                            (unsigned)gh < (unsigned)P.H && (unsigned)gw < (unsigned)P.W;
                            keepv[k] = ok ? P.dropbits[((img_row0 + gh) * P.W + gw) * (size_t)(P.Cin >> 3) + (ch >> 3)] : 0xffu;
                        }
                    }
                    ptx::mbar_wait(bar_full + 8 * stage, phase, 5);
                    const uint32_t base = stage_u32 + stage * C::STAGE_BYTES + c * C::CH_STRIDE + p0 * 16;
#pragma unroll
                    for (int k = 0; k < ITERS; ++k) {
                        if ((k + 1) * STEP <= C::NPIX || p0 + k * STEP < C::NPIX) {
                            const uint32_t addr = base + k * STEP * 16;
                            const int gh = h0 + (hrc[k] >> 8), gw = w0 + (hrc[k] & 255);
                            const bool inside = !border || ((unsigned)gh < (unsigned)P.H && (unsigned)gw < (unsigned)P.W);
                            const uint32_t keep = XF == 2 ? keepv[k] : 0xffu;
                            float f[8];
                            unpack8(ptx::lds128(addr), f);
                            if (XF == 3) {
                                float r[8];
                                unpack8(ptx::lds128(addr + C::OFF_R), r);
#pragma unroll
                                for (int j = 0; j < 8; ++j) f[j] = fmaf(scl[j], f[j], fmaf(shf[j], r[j], kdv[j]));
                            } else {
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    float a = fmaf(f[j], scl[j], shf[j]);
                                    a = fmaxf(a, kLeakySlope * a);
                                    if (XF == 2) a = ((keep >> j) & 1u) ? a * P.inv_keep : 0.f;
                                    f[j] = a;
                                }
                            }
                            uint4 v = make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
                            if (!inside) v = make_uint4(0u, 0u, 0u, 0u);
                            ptx::sts128(addr, v);
                        }
                    }
                    ptx::fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(bar_xf + 8 * stage);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                ti.next(P.tiles_h, P.tiles_w);
            }
        }
    } else if (warp >= 4 + C::XFW) {
        // ================================================================= epilogue warps
        const int q = warp & 3;                      // TMEM lane quarter this warp may read
        const int m = q * 32 + lane;                 // output pixel within the tile
        const int mr = m / kTW, mc = m % kTW;
        const int et = threadIdx.x - (128 + C::XFW * 32);
        const int half = (warp - (4 + C::XFW)) >> 2;     // EPW == 8: which half of the n-block's channels this warp owns
        constexpr int NG = BN / 16;
        constexpr int NGW = NG / (C::EPW / 4);           // 16-channel groups per tile handled by one warp
        constexpr int CW = NGW * 16;                     // channels per warp
        // Per-channel sums (train-mode BatchNorm statistics: sum x | sum x^2; GSTAT: sum g | sum g*raw).
        // BN <= 32: lane-private running sums over ALL pixels this lane ever sees (one FADD + one FFMA per value), reduced
        // across lanes once per CTA.  BN >= 64: per-tile shuffle butterfly into one running value per 16-column group
        // (register budget; a shared-memory transpose of the 32 x 16 block measured SLOWER: 7.1 k vs 5.5 k cycles per
        // 128 x 128 tile, profiles/README.md).  Either way: one partial row per CTA, fixed summation order.
        constexpr bool LANE_STATS = CW <= 32;
        constexpr int NRUN = LANE_STATS ? CW : NGW;
        float run1[NRUN], run2[NRUN];
#pragma unroll
        for (int i = 0; i < NRUN; ++i) run1[i] = run2[i] = 0.f;
        const bool want_stats = !NCHW && P.stats != nullptr && !(P.dbg & 4);    // the logits layer has no BatchNorm
        const bool no_store = (P.dbg & 2) != 0;
        // NCHW (logits) kernels: BN == 16, one n-block; the bias of the real output channels lives in registers
        float bias_r[NCHW ? 16 : 1];
        const size_t plane = (size_t)P.H * P.W;
        if constexpr (NCHW) {
#pragma unroll
            for (int i = 0; i < 16; ++i) bias_r[i] = (P.bias && i < P.out_c_real) ? P.bias[i] : 0.f;
        }
        // store one pixel's 16 consecutive output channels starting at channel c0 of this n-block
        auto store16 = [&](const float *v, size_t pix, int n_img, int gh, int gw, int c0) {
            if constexpr (NCHW) {
                float *o = P.out_nchw + ((size_t)n_img * P.out_c_real * P.H + gh) * P.W + gw;
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (i < P.out_c_real) o[(size_t)i * plane] = v[i] + bias_r[i];
            } else {
                bf16 *orow = P.out + pix * P.Cout + nb * BN + c0;
                float b[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) b[i] = v[i] + ((!GSTAT && P.bias) ? P.bias[nb * BN + c0 + i] : 0.f);
                *reinterpret_cast<uint4 *>(orow) = make_uint4(pack2(b[0], b[1]), pack2(b[2], b[3]), pack2(b[4], b[5]), pack2(b[6], b[7]));
                *reinterpret_cast<uint4 *>(orow + 8) = make_uint4(pack2(b[8], b[9]), pack2(b[10], b[11]), pack2(b[12], b[13]), pack2(b[14], b[15]));
            }
        };
        // One 16-channel group of one pixel: accumulator values v (fp32) -> the stored values (in place) and the two summed
        // quantities a | b.  GSTAT needs the output-side raw tensor (and dropout bits) of the same pixel / channels.
        auto transform16 = [&](float *v, float *b2, bool valid, const uint4 &r0, const uint4 &r1, uint32_t keep16, int c0) {
            if constexpr (GSTAT) {
                float xa[8], xb[8], x[16];
                unpack8(r0, xa);
                unpack8(r1, xb);
#pragma unroll
                for (int i = 0; i < 8; ++i) { x[i] = xa[i]; x[8 + i] = xb[i]; }
                const float *gsc = s_gsc + nb * BN + c0, *gsh = s_gsh + nb * BN + c0;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float z = fmaf(x[i], gsc[i], gsh[i]);
                    float g = z > 0.f ? v[i] : kLeakySlope * v[i];
                    if (P.gs_dropbits) g = ((keep16 >> i) & 1u) ? g * P.gs_inv_keep : 0.f;
                    g = valid ? g : 0.f;
                    v[i] = g;
                    b2[i] = g * x[i];
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    v[i] = valid ? v[i] : 0.f;
                    b2[i] = v[i] * v[i];
                }
            }
        };
        // groups of one work item: grp -> (UMMA tile j = grp / NG, channel group gi = grp % NG); batches of GB groups share one
        // tcgen05.wait::ld (register budget: GB * 16 accumulator registers + the running sums)
        constexpr int TG = MT * NGW;
        constexpr int GB0 = LANE_STATS ? (CW == 16 ? (GSTAT ? 2 : 4) : (GSTAT ? 1 : 2)) : (GSTAT ? 2 : 4);
        constexpr int GB = GB0 < TG ? GB0 : TG;
        TileIter ti;
        ti.init(mt0, mstep, P.tiles_h, P.tiles_w);
#pragma unroll 1
        for (int it = 0; it < n_work; ++it) {
            const int acc = it % C::NACC, acc_phase = (it / C::NACC) & 1;
            const int gh = ti.th * kTH + mr, gw0 = ti.tw * C::TWP + mc;
            const size_t pix0 = ((size_t)ti.n_img * P.H + gh) * P.W + gw0;
            // GSTAT: the first batch's raw values are requested before the accumulator wait (their latency hides behind it)
            uint4 rw[GSTAT ? GB : 1][2];
            uint32_t kp[GSTAT ? GB : 1];
            auto load_raw = [&](int b0) {
                if constexpr (GSTAT) {
#pragma unroll
                    for (int t = 0; t < GB; ++t) {
                        const int grp = b0 + t, j = grp / NGW, c0 = (half * NGW + grp % NGW) * 16;
                        const bool valid = gh < P.H && gw0 + j * kTW < P.W;
                        const size_t e = (pix0 + j * kTW) * P.Cout + nb * BN + c0;
                        rw[t][0] = rw[t][1] = make_uint4(0u, 0u, 0u, 0u);
                        kp[t] = 0xffffu;
                        if (valid) {
                            rw[t][0] = *reinterpret_cast<const uint4 *>(P.gs_raw + e);
                            rw[t][1] = *reinterpret_cast<const uint4 *>(P.gs_raw + e + 8);
                            if (P.gs_dropbits) kp[t] = P.gs_dropbits[e >> 4];
                        }
                    }
                }
            };
            load_raw(0);
            ptx::mbar_wait(bar_tfull + 8 * acc, acc_phase, 6);
            ptx::tc_fence_after();
            if (q == 0) HPFG_TRACE(3, it);
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * C::ACC_COLS;
#pragma unroll
            for (int b0 = 0; b0 < TG; b0 += GB) {
                uint32_t r[GB][16];
#pragma unroll
                for (int t = 0; t < GB; ++t)
                    ptx::tmem_ld16(taddr + ((b0 + t) / NGW) * BN + (half * NGW + (b0 + t) % NGW) * 16, r[t]);   // tile j, channel group gi
                if (b0 > 0) load_raw(b0);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int t = 0; t < GB; ++t) {
                    constexpr int dummy = 0;
                    (void)dummy;
                    const int grp = b0 + t, j = grp / NGW, gl = grp % NGW, gi = half * NGW + gl;
                    const int gw = gw0 + j * kTW;
                    const bool valid = gh < P.H && gw < P.W;
                    float v[16], b2[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[t][i]);
                    transform16(v, b2, valid, rw[GSTAT ? t : 0][0], rw[GSTAT ? t : 0][1], kp[GSTAT ? t : 0], gi * 16);
                    if (want_stats) {
                        if constexpr (LANE_STATS) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) { run1[gl * 16 + i] += v[i]; run2[gl * 16 + i] += b2[i]; }
                        } else {
                            run1[gl] += butterfly16(v, lane);
                            run2[gl] += butterfly16(b2, lane);
                        }
                    }
                    if (valid && !no_store) store16(v, pix0 + j * kTW, ti.n_img, gh, gw, gi * 16);
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar_tempty + 8 * acc);   // accumulator buffer free for the MMA warp
            if (q == 0) HPFG_TRACE(4, it);
            ti.next(P.tiles_h, P.tiles_w);
        }
        if (!NCHW && P.stats) {    // once per CTA: combine lanes and the epilogue warps, write this CTA's partial row
            float *mine = s_part + (warp - (4 + C::XFW)) * 2 * CW;          // [EPW warps][sum(CW) | second sum(CW)]
#pragma unroll
            for (int gl = 0; gl < NGW; ++gl) {
                float c1, c2;
                if constexpr (LANE_STATS) {
                    c1 = butterfly16(run1 + gl * 16, lane);
                    c2 = butterfly16(run2 + gl * 16, lane);
                } else {
                    c1 = run1[gl];
                    c2 = run2[gl];
                }
                if ((lane & 1) == 0) {
                    mine[gl * 16 + col16(lane)] = c1;
                    mine[CW + gl * 16 + col16(lane)] = c2;
                }
            }
            ptx::named_bar_sync(1, C::EPW * 32);
            for (int i = et; i < 2 * P.Cout; i += C::EPW * 32) {
                const int which = i / P.Cout, c = i % P.Cout, n = c - nb * BN;
                float sum = 0.f;
                if (n >= 0 && n < BN && n_work > 0) {
                    const float *src = s_part + (n / CW) * 4 * 2 * CW + which * CW + n % CW;     // the four lane quarters of that half
                    sum = src[0] + src[2 * CW] + src[4 * CW] + src[6 * CW];
                }
                P.stats[(size_t)blockIdx.x * 2 * P.Cout + i] = sum;
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (P.trace && threadIdx.x == 0 && blockIdx.x == 0) P.trace[5 * 64 + 1] = clock64();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------ launch dispatch
// One function per (KS, KC, BN), explicitly instantiated in conv_tc_inst*.cu so the kernel variants compile in
// parallel.  mt: stage width (1/2/4 UMMA tiles); xf: loader transform (0..3); epi: epilogue (0..2).
template <int KS, int KC, int BN>
int tc_launch(int mt, int xf, int epi, const CUtensorMap &map, const CUtensorMap &map2, const TcConvParams &P, cudaStream_t s);

template <int KS, int KC, int BN, bool RES, int MT, int XF, int EPI>
static int tc_launch_one(const CUtensorMap &map, const CUtensorMap &map2, const TcConvParams &P, cudaStream_t s) {
    using C = TcCfg<KS, KC, BN, RES, MT, XF == 3>;
    static bool attr_set = false;
    if (!attr_set) {
        HPFG_CUDA_CHECK(cudaFuncSetAttribute(tc_conv_kernel<KS, KC, BN, RES, MT, XF, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr_set = true;
    }
    const int total = P.m_tiles * P.n_blocks;
    const int grid = total < P.max_ctas ? total : P.max_ctas;
    HPFG_CUDA_CHECK(launch_pdl(tc_conv_kernel<KS, KC, BN, RES, MT, XF, EPI>, grid, kTcThreads, C::SMEM_BYTES, s, map, map2, P));
    HPFG_LAUNCH_CHECK();
    return HPFG_OK;
}

// (xf, epi) pairs that exist: forward (0..2, 0), logits (1, 1), data gradients (0 | 3, 0 | 2)
template <int KS, int KC, int BN, bool RES, int MT>
static int tc_launch_xf(int xf, int epi, const CUtensorMap &map, const CUtensorMap &map2, const TcConvParams &P, cudaStream_t s) {
    if constexpr (KS == 3 && KC == 16 && BN == 16 && RES) {     // out_conv: fp32 NCHW logits epilogue
        if (epi == 1 && xf == 1) return tc_launch_one<KS, KC, BN, RES, MT, 1, 1>(map, map2, P, s);
    }
    if (epi == 0) {
        if constexpr (KS == 3) {                                   // dropout only follows the first conv of an encoder ConvBlock
            if (xf == 2) return tc_launch_one<KS, KC, BN, RES, MT, 2, 0>(map, map2, P, s);
            if (xf == 3) return tc_launch_one<KS, KC, BN, RES, MT, 3, 0>(map, map2, P, s);
        }
        if (xf == 1) return tc_launch_one<KS, KC, BN, RES, MT, 1, 0>(map, map2, P, s);
        if (xf == 0) return tc_launch_one<KS, KC, BN, RES, MT, 0, 0>(map, map2, P, s);
    } else if (epi == 2) {
        if constexpr (KS == 3) {
            if (xf == 3) return tc_launch_one<KS, KC, BN, RES, MT, 3, 2>(map, map2, P, s);
        }
        if (xf == 0) return tc_launch_one<KS, KC, BN, RES, MT, 0, 2>(map, map2, P, s);
    }
    set_error("tc conv: unsupported loader transform / epilogue combination xf=" + std::to_string(xf) + " epi=" + std::to_string(epi));
    return HPFG_ERR_UNSUPPORTED;
}

template <int KS, int KC, int BN, bool RES>
static int tc_launch_mt(int mt, int xf, int epi, const CUtensorMap &map, const CUtensorMap &map2, const TcConvParams &P, cudaStream_t s) {
    if constexpr (KS == 3 && BN <= 32) {      // wide stages only where per-tile overheads dominate (few channels, large images)
        if (mt == 4) return tc_launch_xf<KS, KC, BN, RES, 4>(xf, epi, map, map2, P, s);
        if (mt == 2) return tc_launch_xf<KS, KC, BN, RES, 2>(xf, epi, map, map2, P, s);
    }
    return tc_launch_xf<KS, KC, BN, RES, 1>(xf, epi, map, map2, P, s);
}

template <int KS, int KC, int BN>
int tc_launch(int mt, int xf, int epi, const CUtensorMap &map, const CUtensorMap &map2, const TcConvParams &P, cudaStream_t s) {
    if constexpr (BN <= 64) {     // resident weights whenever the layer's whole weight is one (k-chunk, n-block) stage
        if (P.k_chunks == 1 && P.n_blocks == 1) return tc_launch_mt<KS, KC, BN, true>(mt, xf, epi, map, map2, P, s);
    }
    return tc_launch_mt<KS, KC, BN, false>(mt, xf, epi, map, map2, P, s);
}

}  // namespace hpfg
