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

#include "conv_tc_kernel.cuh"

namespace hpfg {

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

// blockIdx.y = table entry (one layer, fprop or dgrad order); one thread packs 8 consecutive k (one 16-byte store)
__global__ void __launch_bounds__(256) tc_pack_kernel(const float *__restrict__ params, bf16 *__restrict__ dst, const PackTable T, int skip_dgrad) {
    pdl_prologue();
    const PackEntry &E = T.e[blockIdx.y];
    if (skip_dgrad && E.dgrad) return;        // forward-only plans (the teacher) never read the transposed weights
    const unsigned cin_v = E.dgrad ? E.cout : E.cin, cout_v = E.dgrad ? E.cin : E.cout;
    const unsigned groups = cin_v * cout_v * E.kk / 8;          // 8-element groups of this entry
    const unsigned k8s = E.KC / 8, k_chunks = cin_v / E.KC;
    for (unsigned g = blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += gridDim.x * blockDim.x) {
        unsigned e = g;
        const unsigned n = e % E.BN;   e /= E.BN;
        const unsigned k8 = e % k8s;   e /= k8s;
        const unsigned tap = e % E.kk; e /= E.kk;
        const unsigned kc = e % k_chunks, nb = e / k_chunks;
        const unsigned co_v = nb * E.BN + n, ci0 = kc * E.KC + k8 * 8;
        float w[8];
#pragma unroll
        for (unsigned k = 0; k < 8; ++k) {
            const unsigned ci_v = ci0 + k;
            w[k] = 0.f;
            if (!E.dgrad) {
                if (co_v < (unsigned)E.cout_real && ci_v < (unsigned)E.cin_real) w[k] = params[E.w_off + ((long long)co_v * E.cin_real + ci_v) * E.kk + tap];
            } else if (ci_v < (unsigned)E.cout_real && co_v < (unsigned)E.cin_real) {       // rot180 + transpose
                w[k] = params[E.w_off + ((long long)ci_v * E.cin_real + co_v) * E.kk + (E.kk - 1 - tap)];
            }
        }
        *reinterpret_cast<uint4 *>(dst + E.dst_begin + (long long)g * 8) = make_uint4(pack2(w[0], w[1]), pack2(w[2], w[3]), pack2(w[4], w[5]), pack2(w[6], w[7]));
    }
}

static dim3 pack_grid(const PackTable &T) {
    long long mx = 1;
    for (int i = 0; i < T.n; ++i) {
        const long long n = (i + 1 < T.n ? T.e[i + 1].dst_begin : T.total) - T.e[i].dst_begin;
        mx = std::max(mx, n);
    }
    return dim3((unsigned)std::min<long long>((mx / 8 + 255) / 256, 64), (unsigned)T.n);
}

// ------------------------------------------------------------------------------------------ host side
static void pick_cfg(int cin_v, int cout_v, int &KC, int &BN) {
    BN = std::min(cout_v, 128);
    KC = (cin_v == 16 || BN == 128) ? 16 : 32;
}

// explicit instantiations live in conv_tc_inst*.cu
static int launch_ks(int ks, int KC, int BN, int mt, int xf, int epi, const CUtensorMap &map, const CUtensorMap &map2, const TcConvParams &P,
                     cudaStream_t s) {
#define HPFG_TC_CASE(kc, bn)                                                        \
    if (KC == kc && BN == bn)                                                       \
        return ks == 3 ? tc_launch<3, kc, bn>(mt, xf, epi, map, map2, P, s) : tc_launch<1, kc, bn>(mt, xf, epi, map, map2, P, s);
    HPFG_TC_CASE(16, 16) HPFG_TC_CASE(16, 32) HPFG_TC_CASE(16, 128) HPFG_TC_CASE(32, 16) HPFG_TC_CASE(32, 32) HPFG_TC_CASE(32, 64)
#undef HPFG_TC_CASE
    set_error("tc conv: no kernel for KC=" + std::to_string(KC) + " BN=" + std::to_string(BN));
    return HPFG_ERR_UNSUPPORTED;
}

static long long *g_tc_trace = nullptr;
static int g_tc_fwd_cap = 0;   // > 0: persistent-CTA budget of the forward convolutions of the plan being run (hpfg_unet_plan_set_forward_ctas)
static int g_tc_kind = 0;   // 0 forward conv, 1 data gradient (set by the entry points around tc_run: single host thread per plan)
static int g_tc_dbg = 0;   // bottleneck-isolation switches, set only by the micro-benchmark entry (never by the product path)

// Run one convolution (conv-view channels cin_v -> cout_v) on the tensor cores.
static int tc_run(int ks, int N, int H, int W, int cin_v, int cout_v, const void *in, void *out, const bf16 *bpk,
                  const float *bias, LoadXform xf, float *stats, int *P_out, cudaStream_t s, float *out_nchw = nullptr,
                  int out_c_real = 0, const TcBwdFuse *fuse = nullptr) {
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
    CUtensorMap map, map2;
    HPFG_RETURN_IF(make_map_chunked(&map, in, N, H, W, cin_v, KC / 8, kTW * mt + ks - 1, kTH + ks - 1));
    map2 = map;
    const bool two = fuse && fuse->in_raw, gstat = fuse && fuse->out_raw;
    if (two) HPFG_RETURN_IF(make_map_chunked(&map2, fuse->in_raw, N, H, W, cin_v, KC / 8, kTW * mt + ks - 1, kTH + ks - 1));
    TcConvParams P{};
    P.bpk = bpk; P.out = (bf16 *)out; P.bias = bias;
    P.scale = xf.scale; P.shift = xf.shift;
    P.dropbits = reinterpret_cast<const uint8_t *>(xf.drop.bits); P.inv_keep = xf.drop.inv_keep;
    P.stats = stats;
    P.out_nchw = out_nchw; P.out_c_real = out_c_real;
    P.N = N; P.H = H; P.W = W; P.Cin = cin_v; P.Cout = cout_v;
    P.tiles_h = (H + kTH - 1) / kTH; P.tiles_w = (W + kTW * mt - 1) / (kTW * mt);
    P.m_tiles = N * P.tiles_h * P.tiles_w; P.n_blocks = cout_v / BN; P.k_chunks = cin_v / KC;
    P.dbg = g_tc_dbg;
    P.trace = g_tc_trace;
    P.max_ctas = tc_cta_cap(g_tc_kind);
    if (g_tc_kind == 0 && g_tc_fwd_cap > 0) P.max_ctas = std::min(P.max_ctas, g_tc_fwd_cap);
    if (P_out) *P_out = std::min(P.m_tiles * P.n_blocks, P.max_ctas);
    int xfm = xf.scale ? (xf.drop.bits ? 2 : 1) : 0;
    if (two) {          // BatchNorm backward in the loader: draw = sc*g + kb*raw + kd
        xfm = 3;
        P.scale = fuse->sc; P.shift = fuse->kb; P.kd = fuse->kd;
    }
    if (gstat) {
        P.gs_raw = (const bf16 *)fuse->out_raw; P.gs_scale = fuse->gs_scale; P.gs_shift = fuse->gs_shift;
        P.gs_dropbits = reinterpret_cast<const uint16_t *>(fuse->gs_dropbits); P.gs_inv_keep = fuse->gs_inv_keep;
    }
    return launch_ks(ks, KC, BN, mt, xfm, out_nchw != nullptr ? 1 : (gstat ? 2 : 0), map, map2, P, s);
}

// micro-benchmark entry (wgrad_tc.cu: hpfg_conv_tc_bench): packs once per call (cheap) and launches one convolution
int tc_run_bench(int op, int ks, int N, int H, int W, int cin, int cout, const void *in, void *out, const float *w, const float *scale,
                 const float *shift, float *stats, cudaStream_t s, int fuse_mode, const void *aux) {
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
    { const char *e = getenv("HPFG_TC_DBG"); g_tc_dbg = e ? atoi(e) : 0; }
    g_tc_fwd_cap = 0;
    const int cin_v = op ? cout : cin, cout_v = op ? cin : cout;
    LoadXform xf{};
    xf.scale = scale; xf.shift = shift; xf.drop.bits = nullptr; xf.drop.inv_keep = 1.f;
    const bool want_trace = getenv("HPFG_TC_TRACE") != nullptr;
    if (want_trace && !g_tc_trace) cudaMalloc(&g_tc_trace, 6 * 64 * 8);
    if (!want_trace) g_tc_trace = nullptr;
    TcBwdFuse f;        // fuse_mode (data gradients only): bit 0 two-source loader, bit 1 GSTAT epilogue; aux = a second bf16 tensor
    if (fuse_mode & 1) { f.in_raw = aux; f.sc = scale; f.kb = scale; f.kd = scale; }
    if (fuse_mode & 2) { f.out_raw = aux; f.gs_scale = scale; f.gs_shift = scale; }
    if (fuse_mode) { xf.scale = nullptr; xf.shift = nullptr; }
    const int rc = tc_run(ks, N, H, W, cin_v, cout_v, in, out, packed, nullptr, xf, stats, nullptr, s, nullptr, 0, fuse_mode ? &f : nullptr);
    g_tc_dbg = 0;
    if (want_trace && getenv("HPFG_TC_TRACE_DUMP")) {       // print CTA 0's per-role timestamps (cycles since the CTA's setup barrier)
        long long h[6 * 64];
        cudaStreamSynchronize(s);
        cudaMemcpy(h, g_tc_trace, sizeof(h), cudaMemcpyDeviceToHost);
        const char *names[5] = {"tma:slot-free", "mma:operands-ready", "mma:issued", "epi:acc-ready", "epi:done"};
        printf("trace: kernel body %lld cycles\n", h[5 * 64 + 1] - h[5 * 64]);
        for (int r = 0; r < 5; ++r) {
            printf("%-20s", names[r]);
            for (int i = 0; i < 24; ++i) printf(" %6lld", h[r * 64 + i] - h[5 * 64]);
            printf("\n");
        }
        fflush(stdout);
    }
    g_tc_trace = nullptr;
    return rc;
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

int tc_pack_all(hpfg_unet_plan *p, const float *params, cudaStream_t s, bool need_dgrad) {
    ProfScope _prof(PROF_PACK, s);
    auto *st = reinterpret_cast<TcPlanState *>(p->tc);
    HPFG_CUDA_CHECK(launch_pdl(tc_pack_kernel, pack_grid(st->table), 256, 0, s, params, st->packed, st->table, need_dgrad ? 0 : 1));
    HPFG_LAUNCH_CHECK();
    return HPFG_OK;
}

int tc_fprop(hpfg_unet_plan *p, int conv, const void *in, void *out, LoadXform xf, float *stats, int *P, bool *done,
             cudaStream_t s) {
    g_tc_fwd_cap = p->fwd_ctas & ~1;
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
    g_tc_fwd_cap = p->fwd_ctas & ~1;
    auto *st = reinterpret_cast<TcPlanState *>(p->tc);
    const ConvLayer &cv = p->d.convs[conv];
    *done = false;
    if (st->f_off[conv] < 0) return HPFG_OK;
    HPFG_RETURN_IF(tc_run(cv.ks, p->N, cv.H, cv.W, cv.cin, cv.cout, in, out, st->packed + st->f_off[conv], bias, xf, nullptr, nullptr, s));
    *done = true;
    return HPFG_OK;
}

int tc_fprop_logits(hpfg_unet_plan *p, int conv, const void *in, LoadXform xf, const float *bias, float *logits_nchw, cudaStream_t s) {
    g_tc_fwd_cap = p->fwd_ctas & ~1;
    auto *st = reinterpret_cast<TcPlanState *>(p->tc);
    const ConvLayer &cv = p->d.convs[conv];
    return tc_run(cv.ks, p->N, cv.H, cv.W, cv.cin, pad16(cv.cout), in, nullptr, st->packed + st->f_off[conv], bias, xf, nullptr, nullptr, s,
                  logits_nchw, cv.cout);
}

int tc_dgrad(hpfg_unet_plan *p, int conv, const void *dout, void *din, bool *done, cudaStream_t s, const TcBwdFuse *fuse, float *stats,
             int *P) {
    auto *st = reinterpret_cast<TcPlanState *>(p->tc);
    const ConvLayer &cv = p->d.convs[conv];
    *done = false;
    if (st->d_off[conv] < 0) return HPFG_OK;
    const LoadXform none{};
    g_tc_kind = 1;
    const int rc = tc_run(cv.ks, p->N, cv.H, cv.W, pad16(cv.cout), cv.cin, dout, din, st->packed + st->d_off[conv], nullptr, none,
                          (fuse && fuse->out_raw) ? stats : nullptr, P, s, nullptr, 0, fuse);
    g_tc_kind = 0;
    HPFG_RETURN_IF(rc);
    *done = true;
    return HPFG_OK;
}

int tc_wgrad(hpfg_unet_plan *p, int conv, const void *in, LoadXform xf, const void *dout, float *dw_oihw, float *dbias,
             int accumulate, bool *done, cudaStream_t s, const TcBwdFuse *fuse, const TcWgradStreams *ws) {
    const ConvLayer &cv = p->d.convs[conv];
    *done = false;
    float *scratch = (ws && ws->scratch) ? ws->scratch : p->wscratch;
    HPFG_RETURN_IF(tc_wgrad_run(cv.ks, p->N, cv.H, cv.W, pad16(cv.cin), pad16(cv.cout), cv.cin, cv.cout, in, xf, dout, scratch,
                                p->wscratch_floats, dw_oihw, dbias, accumulate, s, fuse, ws));
    *done = true;
    return HPFG_OK;
}

}  // namespace hpfg

// ---- layer-isolated test hook ----------------------------------------------------------------------------
using namespace hpfg;

__global__ void tc_debug_reduce_stats(const float *partials, int P, int C2, float *out) {
    pdl_prologue();
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
    g_tc_fwd_cap = 0;
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
    HPFG_CUDA_CHECK(launch_pdl(tc_pack_kernel, pack_grid(T), 256, 0, s, w_oihw, packed, T, 0));
    HPFG_LAUNCH_CHECK();
    // the conv kernel requests its first weight stages before griddepcontrol.wait (safe in the network, where packing is
    // at least two launches back); here packing is the direct predecessor, so finish it first
    HPFG_CUDA_CHECK(cudaStreamSynchronize(s));
    LoadXform xf{};
    xf.scale = scale; xf.shift = shift; xf.drop.bits = nullptr; xf.drop.inv_keep = 1.f;
    int P = 0;
    int rc = tc_run(ks, N, H, W, cin_v, cout_v, in_bf16_nhwc, out_bf16_nhwc, packed, bias, xf, stats_out ? partials : nullptr, &P, s);
    if (rc == HPFG_OK && stats_out) {
        HPFG_CUDA_CHECK(launch_pdl(tc_debug_reduce_stats, (2 * cout_v + 127) / 128, 128, 0, s, partials, P, 2 * cout_v, stats_out));
        ++g_launch_count;
    }
    cudaStreamSynchronize(s);
    cudaFree(packed);
    cudaFree(partials);
    if (rc == HPFG_OK) HPFG_CUDA_CHECK(cudaGetLastError());
    return rc;
}

// Layer-isolated hook for the BatchNorm-backward fusions of the data-gradient kernel (see include/hpfg_b200.h).
extern "C" int hpfg_dgrad_tc_fused_debug(int N, int H, int W, int cin, int cout, int ks, const void *g_in, const void *raw_in,
                                         const float *sc, const float *kb, const float *kd, const float *w_oihw, const void *raw_out,
                                         const float *gs_scale, const float *gs_shift, const uint8_t *gs_keep_mask_nchw, float gs_p_drop,
                                         void *out_bf16_nhwc, float *stats_out, void *stream) {
    HPFG_REQUIRE(cin % 16 == 0 && cout % 16 == 0 && (ks == 1 || ks == 3), "hpfg_dgrad_tc_fused_debug: unsupported shape");
    HPFG_REQUIRE(g_in && w_oihw && out_bf16_nhwc, "hpfg_dgrad_tc_fused_debug: null argument");
    g_tc_fwd_cap = 0;
    cudaStream_t s = (cudaStream_t)stream;
    PackTable T{};
    T.n = 1;
    T.total = (long long)cin * cout * ks * ks;
    pick_cfg(cout, cin, T.e[0].KC, T.e[0].BN);
    T.e[0].dst_begin = 0; T.e[0].w_off = 0; T.e[0].cin = cin; T.e[0].cout = cout; T.e[0].kk = ks * ks; T.e[0].dgrad = 1;
    T.e[0].cin_real = cin; T.e[0].cout_real = cout;
    bf16 *packed = nullptr;
    float *partials = nullptr;
    uint32_t *bits = nullptr;
    HPFG_CUDA_CHECK(cudaMalloc(&packed, (size_t)T.total * 2));
    HPFG_CUDA_CHECK(cudaMalloc(&partials, (size_t)kNumSMs * 2 * cin * 4));
    HPFG_CUDA_CHECK(launch_pdl(tc_pack_kernel, pack_grid(T), 256, 0, s, w_oihw, packed, T, 0));
    HPFG_LAUNCH_CHECK();
    if (gs_keep_mask_nchw) {
        HPFG_CUDA_CHECK(cudaMalloc(&bits, ((size_t)N * H * W * cin + 31) / 32 * 4));
        HPFG_RETURN_IF(dropout_bits(bits, gs_keep_mask_nchw, N, H, W, cin, gs_p_drop, 0, 0, s));
    }
    HPFG_CUDA_CHECK(cudaStreamSynchronize(s));
    TcBwdFuse f;
    if (raw_in) { f.in_raw = raw_in; f.sc = sc; f.kb = kb; f.kd = kd; }
    if (raw_out) {
        f.out_raw = raw_out; f.gs_scale = gs_scale; f.gs_shift = gs_shift;
        f.gs_dropbits = bits; f.gs_inv_keep = bits ? 1.f / (1.f - gs_p_drop) : 1.f;
    }
    const LoadXform none{};
    int P = 0;
    int rc = tc_run(ks, N, H, W, cout, cin, g_in, out_bf16_nhwc, packed, nullptr, none, raw_out ? partials : nullptr, &P, s, nullptr, 0, &f);
    if (rc == HPFG_OK && stats_out && raw_out) {
        HPFG_CUDA_CHECK(launch_pdl(tc_debug_reduce_stats, (2 * cin + 127) / 128, 128, 0, s, partials, P, 2 * cin, stats_out));
        ++g_launch_count;
    }
    cudaStreamSynchronize(s);
    cudaFree(packed);
    cudaFree(partials);
    if (bits) cudaFree(bits);
    if (rc == HPFG_OK) HPFG_CUDA_CHECK(cudaGetLastError());
    return rc;
}
