// UNet plan: network description, workspace, forward/backward schedules and the C ABI entry points for them.
// Wiring follows model/unet.py: Encoder :61-82, Decoder :85-117, ConvBlock :12-28, DownBlock :31-42, UpBlock :45-58.
#include "unet_plan.cuh"

#include <cstdlib>
#include <cstring>

#include "conv_tc.cuh"

namespace hpfg {

static thread_local std::string g_error;
int64_t g_launch_count = 0;
void set_error(const std::string &msg) { g_error = msg; }

typedef void (*hpfg_allreduce_fn)(void *, void *, int64_t, int, void *);
static hpfg_allreduce_fn g_sync_fn = nullptr;
static void *g_sync_ctx = nullptr;
static int g_sync_world = 1;
bool g_loss_global_sums = false;
int sync_world() { return g_sync_world; }
int sync_allreduce(void *device_ptr, int64_t count, bool is_double, cudaStream_t s) {
    if (g_sync_world <= 1) return HPFG_OK;
    if (!g_sync_fn) {
        set_error("exact-global mode needs an all-reduce hook (hpfg_set_allreduce_hook)");
        return HPFG_ERR_INVALID;
    }
    g_sync_fn(g_sync_ctx, device_ptr, count, is_double ? 1 : 0, (void *)s);
    return HPFG_OK;
}

int tc_cta_cap(int kind) {
    static int caps[3] = {-1, -1, -1};
    if (caps[0] < 0) {
        const char *names[3] = {"HPFG_CTAS_FWD", "HPFG_CTAS_DGRAD", "HPFG_CTAS_WGRAD"};
        const int defaults[3] = {kNumSMs, kNumSMs, kNumSMs};
        for (int i = 0; i < 3; ++i) {
            const char *e = getenv(names[i]);
            int v = e ? atoi(e) : defaults[i];
            if (v < 2 || v > kNumSMs) v = kNumSMs;
            caps[i] = v & ~1;               // even: the two n-blocks of a wide layer alternate over the grid
        }
    }
    return caps[kind];
}

bool g_prof_on = false;
struct ProfRec { int cat; cudaEvent_t a, b; };
static std::vector<ProfRec> g_prof;
void prof_push(int cat, cudaStream_t s, bool begin) {
    if (begin) {
        ProfRec r{cat, nullptr, nullptr};
        cudaEventCreate(&r.a);
        cudaEventCreate(&r.b);
        cudaEventRecord(r.a, s);
        g_prof.push_back(r);
    } else {
        for (int i = (int)g_prof.size() - 1; i >= 0; --i)
            if (g_prof[i].cat == cat) { cudaEventRecord(g_prof[i].b, s); break; }
    }
}

void describe_unet(int in_ch, int n_cls, int H, int W, UNetDesc &d) {
    d.in_ch = in_ch;
    d.n_cls = n_cls;
    d.convs.clear();
    d.bns.clear();
    int64_t off = 0, run = 0;
    int pi = 0;
    auto add_param = [&](int64_t n) {
        d.offsets[pi] = off;
        d.sizes[pi] = n;
        ++pi;
        const int64_t o = off;
        off += n;
        return o;
    };
    auto add_conv = [&](const std::string &name, int cin, int cout, int ks, int h, int w, bool with_bn) {
        ConvLayer c;
        c.name = name;
        c.cin = cin; c.cout = cout; c.ks = ks; c.H = h; c.W = w;
        c.w_off = add_param((int64_t)cout * cin * ks * ks);
        c.b_off = add_param(cout);
        c.bn = -1;
        if (with_bn) {
            BnLayer b;
            b.C = cout;
            b.g_off = add_param(cout);
            b.b_off = add_param(cout);
            b.run_off = run;
            run += 2 * cout;
            b.conv = (int)d.convs.size();
            b.H = h; b.W = w;
            c.bn = (int)d.bns.size();
            d.bns.push_back(b);
        }
        d.convs.push_back(c);
    };
    auto add_block = [&](const std::string &prefix, int cin, int cout, int h, int w) {
        add_conv(prefix + ".conv_conv.0", cin, cout, 3, h, w, true);
        add_conv(prefix + ".conv_conv.4", cout, cout, 3, h, w, true);
    };
    int64_t bounds[5] = {0, 0, 0, 0, 0};
    add_block("encoder.in_conv", in_ch, kFt[0], H, W);
    for (int l = 1; l < 5; ++l) {
        if (l == 3) bounds[1] = off;      // last (exposed) bucket = in_conv .. down2 only: 72 k parameters
        add_block("encoder.down" + std::to_string(l) + ".maxpool_conv.1", kFt[l - 1], kFt[l], H >> l, W >> l);
    }
    for (int j = 1; j < 5; ++j) {
        const int lvl = 4 - j;   // skip level; output resolution
        if (j == 1) bounds[2] = off;
        if (j == 3) bounds[3] = off;
        const std::string p = "decoder.up" + std::to_string(j);
        add_conv(p + ".conv1x1", kFt[lvl + 1], kFt[lvl], 1, H >> (lvl + 1), W >> (lvl + 1), false);
        add_block(p + ".conv", 2 * kFt[lvl], kFt[lvl], H >> lvl, W >> lvl);
    }
    add_conv("decoder.out_conv", kFt[0], n_cls, 3, H, W, false);
    bounds[4] = off;
    d.n_params = off;
    d.n_bn_floats = run;
    // buckets in completion order (tail first)
    for (int b = 0; b <= kNumBuckets; ++b) d.bucket_begin[b] = bounds[kNumBuckets - b];
}

static int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

struct Carver {
    char *base;
    int64_t off = 0;
    template <typename P> void take(P *&ptr, int64_t bytes) {
        ptr = base ? reinterpret_cast<P *>(base + off) : nullptr;
        off += align_up(bytes, 256);
    }
};

static void carve(hpfg_unet_plan *p, char *base, int64_t &total) {
    Carver c{base};
    const int64_t N = p->N, H = p->H, W = p->W;
    const int64_t e = (int64_t)p->elt;
    for (auto &b : p->d.bns) c.take(b.raw, N * b.H * b.W * b.C * e);
    for (int l = 1; l < 5; ++l) c.take(p->pooled[l], N * (H >> l) * (W >> l) * kFt[l - 1] * e);
    for (int j = 1; j < 5; ++j) {
        const int lvl = 4 - j;
        c.take(p->low[j], N * (H >> (lvl + 1)) * (W >> (lvl + 1)) * kFt[lvl] * e);
        c.take(p->cat[j], N * (H >> lvl) * (W >> lvl) * 2 * kFt[lvl] * e);
        c.take(p->dcat[j], N * (H >> lvl) * (W >> lvl) * 2 * kFt[lvl] * e);
    }
    for (int k = 0; k < 5; ++k) c.take(p->g[k], N * H * W * 16 * e);
    if (p->precision == HPFG_PREC_BF16) {
        c.take(p->xpad, N * H * W * 16 * 2);
        c.take(p->dlpad, N * H * W * 16 * 2);
    }
    for (int l = 0; l < 5; ++l) c.take(p->dropbits[l], (N * (H >> l) * (W >> l) * kFt[l] + 31) / 32 * 4);
    // BN statistic partials: one row of 2*C floats per 8x8 pixel tile (CUDA-core path) / per CTA tile (tensor path)
    int64_t sf = 0, wf = 0;
    for (auto &b : p->d.bns) {
        const int64_t tiles = conv_ref_num_tiles((int)N, b.H, b.W);
        sf = std::max<int64_t>(sf, tiles * 2 * b.C);
    }
    sf = std::max<int64_t>(sf, (int64_t)kNumSMs * 8 * 2 * 256);   // bn_bwd partials
    p->stats_floats = sf;
    c.take(p->stats, sf * 4);
    for (auto &cv : p->d.convs)
    {
        wf = std::max(wf, conv_ref_wgrad_scratch_floats((int)N, cv.H, cv.W, cv.cin, cv.cout, cv.ks));
        if (p->precision == HPFG_PREC_BF16)
            wf = std::max(wf, tc_wgrad_scratch_floats((int)N, cv.H, cv.W, (cv.cin + 15) / 16 * 16, (cv.cout + 15) / 16 * 16, cv.ks));
    }
    p->wscratch_floats = wf;
    c.take(p->wscratch, wf * 4);
    if (p->precision == HPFG_PREC_BF16) c.take(p->wscratch2, wf * 4);
    int64_t bnf = 0;
    for (auto &b : p->d.bns) bnf += 8 * (int64_t)align_up(b.C, 64);
    c.take(p->bnmem, bnf * 4);
    c.take(p->sync_sums, 4 * 256 * 8);      // exact-global mode: [local | global] x (2 x 256) fp64 sums
    if (base) {
        float *q = p->bnmem;
        for (auto &b : p->d.bns) {
            const int64_t s = align_up(b.C, 64);
            b.st = BnState{q, q + s, q + 2 * s, q + 3 * s, q + 4 * s, q + 5 * s, q + 6 * s, q + 7 * s};
            q += 8 * s;
        }
    }
    for (auto &cv : p->d.convs) {
        const int64_t n = (int64_t)cv.cout * cv.cin * cv.ks * cv.ks;
        c.take(cv.wf, n * 4);
        c.take(cv.wd, n * 4);
    }
    total = c.off;
}

// ------------------------------------------------------------------------------------------------------
template <typename T>
static int forward_impl(hpfg_unet_plan *p, const float *params, float *bn_running, int64_t *bn_counters, const float *x,
                        float *logits, int training, int no_dropout, int save, uint64_t seed, uint64_t offset,
                        const uint8_t *const *masks, cudaStream_t s, const uint64_t *offset_dev = nullptr) {
    UNetDesc &d = p->d;
    const int N = p->N, H = p->H, W = p->W;
    const bool use_drop = training && !no_dropout;
    const bool tc = p->precision == HPFG_PREC_BF16;

    // per-step weight preparation (parameters change every optimiser step)
    if (!tc)
        for (auto &cv : d.convs)
            HPFG_RETURN_IF(pack_weights_ref(params + cv.w_off, cv.wf, cv.wd, cv.cin, cv.cout, cv.ks, s));
    if (tc) HPFG_RETURN_IF(tc_pack_all(p, params, s, save != 0 && training != 0));
    bool drop_on_side = false;
    if (use_drop) {      // all five encoder keep-masks in one launch, on the plan's side stream: nothing reads them before the second
                         // convolution of in_conv, so the 30 us launch overlaps weight packing, input padding and the first convolution
        int hs[5], wsz[5], cs[5];
        float ps[5];
        for (int l = 0; l < 5; ++l) { hs[l] = H >> l; wsz[l] = W >> l; cs[l] = kFt[l]; ps[l] = kEncDropout[l]; }
        static const bool side_ok = !(getenv("HPFG_DROP_SIDE") && getenv("HPFG_DROP_SIDE")[0] == '0');      // A/B switch (profiles/)
        drop_on_side = side_ok && !g_prof_on;
        cudaStream_t ds = drop_on_side ? p->side : s;
        if (drop_on_side) {
            HPFG_CUDA_CHECK(cudaEventRecord(p->ev_ready, s));
            HPFG_CUDA_CHECK(cudaStreamWaitEvent(p->side, p->ev_ready, 0));
            // (the input's bf16 NHWC16 copy on this stream as well measured SLOWER: 3.18-3.23 vs 3.14 ms, profiles/README.md)
        }
        HPFG_RETURN_IF(dropout_bits_multi(5, p->dropbits, masks, N, hs, wsz, cs, ps, seed, offset, ds, offset_dev));
        if (drop_on_side) HPFG_CUDA_CHECK(cudaEventRecord(p->ev_join, p->side));
    }

    // conv + (train: statistics -> finalize | eval: running-stat affine)
    auto conv_bn = [&](int ci, TView in, bool in_is_f32, LoadXform xf) -> int {
        ConvLayer &cv = d.convs[ci];
        BnLayer &b = d.bns[cv.bn];
        TView out = nhwc_view(b.raw, cv.H, cv.W, cv.cout);
        float *stats = training ? p->stats : nullptr;
        int P = conv_ref_num_tiles(N, cv.H, cv.W);
        bool done = false;
        if (tc && !in_is_f32) HPFG_RETURN_IF(tc_fprop(p, ci, in.p, b.raw, xf, stats, &P, &done, s));
        if (!done) {
            if (in_is_f32)
                HPFG_RETURN_IF((conv_ref_fprop<float, T>(in, out, cv.wf, nullptr, N, cv.H, cv.W, cv.cin, cv.cout, cv.ks,
                                                         xf, stats, s)));
            else
                HPFG_RETURN_IF((conv_ref_fprop<T, T>(in, out, cv.wf, nullptr, N, cv.H, cv.W, cv.cin, cv.cout, cv.ks, xf,
                                                     stats, s)));
        }
        if (training)
            return bn_finalize(p->stats, P, b.C, (int64_t)N * cv.H * cv.W, params + b.g_off, params + b.b_off,
                               params + cv.b_off, bn_running + b.run_off, bn_running + b.run_off + b.C,
                               bn_counters ? bn_counters + cv.bn : nullptr, 1, b.st, s, p->sync_bn ? p->sync_sums : nullptr);
        return bn_eval_affine(b.C, params + b.g_off, params + b.b_off, params + cv.b_off, bn_running + b.run_off,
                              bn_running + b.run_off + b.C, b.st, s);
    };
    auto xf_of = [&](int bn, const uint32_t *bits, float p_drop) {
        LoadXform xf{};
        xf.scale = d.bns[bn].st.scale;
        xf.shift = d.bns[bn].st.shift;
        xf.drop.bits = bits;
        xf.drop.inv_keep = bits ? 1.f / (1.f - p_drop) : 1.f;
        return xf;
    };
    const LoadXform none{};

    // ---- encoder (model/unet.py:76-82)
    for (int l = 0; l < 5; ++l) {
        const int h = H >> l, w = W >> l, c0 = 2 * l, c1 = 2 * l + 1;
        if (l == 0 && tc) {      // network input: fp32 NCHW -> bf16 NHWC padded to 16 channels, then a regular tensor-core layer
            HPFG_RETURN_IF(pad_to_nhwc16(x, p->xpad, N, p->in_ch, H, W, s));
            HPFG_RETURN_IF(conv_bn(c0, nhwc_view(p->xpad, H, W, 16), false, none));
        } else if (l == 0)
            HPFG_RETURN_IF(conv_bn(c0, nchw_view(const_cast<float *>(x), p->in_ch, H, W), true, none));
        else
            HPFG_RETURN_IF(conv_bn(c0, nhwc_view(p->pooled[l], h, w, kFt[l - 1]), false, none));
        if (l == 0 && drop_on_side) HPFG_CUDA_CHECK(cudaStreamWaitEvent(s, p->ev_join, 0));     // keep-masks ready
        HPFG_RETURN_IF(conv_bn(c1, nhwc_view(d.bns[c0].raw, h, w, kFt[l]), false,
                               xf_of(c0, use_drop ? p->dropbits[l] : nullptr, kEncDropout[l])));
        if (l < 4)
            HPFG_RETURN_IF(pool_act<T>((const T *)d.bns[c1].raw, (T *)p->pooled[l + 1], N, h, w, kFt[l], d.bns[c1].st, s));
    }
    // ---- decoder (model/unet.py:101-117)
    int prev_bn = 9;
    for (int j = 1; j < 5; ++j) {
        const int lvl = 4 - j, F = kFt[lvl], hl = H >> (lvl + 1), wl = W >> (lvl + 1);
        const int c1x1 = 10 + 3 * (j - 1), cA = c1x1 + 1, cB = c1x1 + 2;
        ConvLayer &c11 = d.convs[c1x1];
        {
            bool done = false;
            LoadXform xf = xf_of(prev_bn, nullptr, 0.f);
            if (tc) HPFG_RETURN_IF(tc_fprop_1x1(p, c1x1, d.bns[prev_bn].raw, p->low[j], xf, params + c11.b_off, &done, s));
            if (!done)
                HPFG_RETURN_IF((conv_ref_fprop<T, T>(nhwc_view(d.bns[prev_bn].raw, hl, wl, 2 * F), nhwc_view(p->low[j], hl, wl, F),
                                                     c11.wf, params + c11.b_off, N, hl, wl, 2 * F, F, 1, xf, nullptr, s)));
        }
        const int skip_bn = 2 * lvl + 1;
        HPFG_RETURN_IF(upcat<T>((const T *)d.bns[skip_bn].raw, d.bns[skip_bn].st, (const T *)p->low[j], (T *)p->cat[j], N, hl,
                                wl, F, s));
        HPFG_RETURN_IF(conv_bn(cA, nhwc_view(p->cat[j], 2 * hl, 2 * wl, 2 * F), false, none));
        HPFG_RETURN_IF(conv_bn(cB, nhwc_view(d.bns[d.convs[cA].bn].raw, 2 * hl, 2 * wl, F), false,
                               xf_of(d.convs[cA].bn, nullptr, 0.f)));
        prev_bn = d.convs[cB].bn;
    }
    // ---- out_conv (model/unet.py:99,116): 3x3 16 -> num_classes with bias, logits fp32 NCHW
    ConvLayer &oc = d.convs[22];
    if (tc)
        HPFG_RETURN_IF(tc_fprop_logits(p, 22, d.bns[17].raw, xf_of(17, nullptr, 0.f), params + oc.b_off, logits, s));
    else
    HPFG_RETURN_IF((conv_ref_fprop<T, float>(nhwc_view(d.bns[17].raw, H, W, 16), nchw_view(logits, p->n_cls, H, W), oc.wf,
                                             params + oc.b_off, N, H, W, 16, p->n_cls, 3, xf_of(17, nullptr, 0.f), nullptr,
                                             s)));
    p->saved = save != 0 && training != 0;
    p->saved_dropout = use_drop;
    p->saved_x = x;
    return HPFG_OK;
}

// HPFG_BWD_FUSE=1 selects the bf16 backward with BatchNorm backward folded into the tensor-core kernels (backward_fused below).
// It is NOT the default: measured on B200 it is slower than the three streaming bn_bwd launches it removes (3.62 vs 3.39 ms
// per Mean-Teacher step, profiles/README.md) -- the fusion moves bytes from 5 TB/s streaming kernels into 3 TB/s conv kernels.
static bool bwd_fused_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("HPFG_BWD_FUSE");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v != 0;
}

// ------------------------------------------------------------------------------------------------------
// bf16 plans: backward with BatchNorm backward folded into the tensor-core kernels (north_star: "the backward pass is fused
// the same way").  Per BatchNorm layer X the unfused schedule ran  bn_bwd<0> -> finalize -> bn_bwd<1> -> {wgrad, dgrad};  here
//   * the kernel that PRODUCES the gradient wrt X's activated output (a data-gradient conv, or skip_pool_bwd for the encoder
//     features) stores g = dact * leaky' * dropout' and writes the partial sums (sum g | sum g*raw)       [GSTAT epilogue]
//   * bn_bwd_reduce turns them into c1/c2 + the folded constants kb/kd (+ dgamma/dbeta)                   [one tiny launch]
//   * the kernels that CONSUME the raw gradient (dgrad and wgrad of X's conv) build it on load from g and raw:
//     draw = scale*g + kb*raw + kd                                                                         [two-source loaders]
// so the raw-gradient tensor never exists and two full tensor passes + one launch per BatchNorm leave the chain.
// Gradient buffers rotate through four slots; a slot is rewritten only after the side-stream weight gradient that read it.
static int backward_fused(hpfg_unet_plan *p, const float *params, const float *dlogits, float *grads, int acc, cudaStream_t s,
                          const float *dbottleneck) {
    using T = bf16;
    UNetDesc &d = p->d;
    const int N = p->N, H = p->H, W = p->W;
    const LoadXform none{};
    auto xf_of = [&](int bn, const uint32_t *bits, float p_drop) {
        LoadXform xf{};
        xf.scale = d.bns[bn].st.scale;
        xf.shift = d.bns[bn].st.shift;
        xf.drop.bits = bits;
        xf.drop.inv_keep = bits ? 1.f / (1.f - p_drop) : 1.f;
        return xf;
    };
    cudaStream_t side = g_prof_on ? s : p->side;
    void *slot[4] = {p->g[0], p->g[1], p->g[2], p->g[4]};
    bool busy[4] = {false, false, false, false};
    int cur = -1;
    auto acquire = [&](void *&out) -> int {      // next gradient slot; waits (stream-level) for its last side-stream reader
        cur = (cur + 1) & 3;
        if (busy[cur]) HPFG_CUDA_CHECK(cudaStreamWaitEvent(s, p->ev_slot[cur], 0));
        busy[cur] = false;
        out = slot[cur];
        return HPFG_OK;
    };
    auto fuse_in = [&](int bn) {                 // two-source loader constants of BatchNorm `bn`
        TcBwdFuse f;
        f.in_raw = d.bns[bn].raw;
        f.sc = d.bns[bn].st.scale; f.kb = d.bns[bn].st.kb; f.kd = d.bns[bn].st.kd;
        return f;
    };
    auto fuse_out = [&](TcBwdFuse &f, int bn, const uint32_t *bits, float p_drop) {
        f.out_raw = d.bns[bn].raw;
        f.gs_scale = d.bns[bn].st.scale; f.gs_shift = d.bns[bn].st.shift;
        f.gs_dropbits = bits; f.gs_inv_keep = bits ? 1.f / (1.f - p_drop) : 1.f;
    };
    // weight gradient of conv `ci` on the side stream; dy = g of the conv's own BatchNorm (fuse != null) or a plain gradient
    auto wgrad = [&](int ci, void *in, LoadXform xf, void *dy, const TcBwdFuse *fuse) -> int {
        ConvLayer &cv = d.convs[ci];
        HPFG_CUDA_CHECK(cudaEventRecord(p->ev_ready, s));
        HPFG_CUDA_CHECK(cudaStreamWaitEvent(side, p->ev_ready, 0));
        bool done = false;
        HPFG_RETURN_IF(tc_wgrad(p, ci, in, xf, dy, grads + cv.w_off, grads + cv.b_off, acc, &done, side, fuse));
        for (int k = 0; k < 4; ++k)
            if (dy == slot[k]) {
                HPFG_CUDA_CHECK(cudaEventRecord(p->ev_slot[k], side));
                busy[k] = true;
            }
        return HPFG_OK;
    };
    auto join_side = [&]() -> int {
        HPFG_CUDA_CHECK(cudaEventRecord(p->ev_join, side));
        HPFG_CUDA_CHECK(cudaStreamWaitEvent(s, p->ev_join, 0));
        return HPFG_OK;
    };
    // data gradient of conv `ci`; gs_bn >= 0: the output is the gradient wrt BatchNorm gs_bn's activated output -> GSTAT + reduce
    auto dgrad = [&](int ci, void *dy, const TcBwdFuse *fin, void *din, int gs_bn, const uint32_t *bits, float p_drop) -> int {
        TcBwdFuse f = fin ? *fin : TcBwdFuse{};
        int P = 0;
        if (gs_bn >= 0) fuse_out(f, gs_bn, bits, p_drop);
        bool done = false;
        HPFG_RETURN_IF(tc_dgrad(p, ci, dy, din, &done, s, (fin || gs_bn >= 0) ? &f : nullptr, p->stats, &P));
        HPFG_REQUIRE(done, "backward_fused: layer not on the tensor-core path");
        if (gs_bn >= 0) {
            BnLayer &bl = d.bns[gs_bn];
            HPFG_RETURN_IF(bn_bwd_reduce(p->stats, P, bl.C, (int64_t)N * bl.H * bl.W, bl.st, grads + bl.g_off, grads + bl.b_off, acc, s));
        }
        return HPFG_OK;
    };

    // ---- out_conv: dlogits fp32 NCHW -> bf16 NHWC16; its data gradient is the gradient wrt up4's activated output (BN 17)
    HPFG_RETURN_IF(pad_to_nhwc16(dlogits, p->dlpad, N, p->n_cls, H, W, s));
    HPFG_RETURN_IF(wgrad(22, d.bns[17].raw, xf_of(17, nullptr, 0.f), p->dlpad, nullptr));
    void *a = nullptr, *c = nullptr, *b = nullptr;
    HPFG_RETURN_IF(acquire(a));
    HPFG_RETURN_IF(dgrad(22, p->dlpad, nullptr, a, 17, nullptr, 0.f));
    // ---- decoder, up4 .. up1 (a = g of the block's second BatchNorm)
    for (int j = 4; j >= 1; --j) {
        const int lvl = 4 - j, c1x1 = 10 + 3 * (j - 1), cA = c1x1 + 1, cB = c1x1 + 2;
        const int bA = d.convs[cA].bn, bB = d.convs[cB].bn;
        TcBwdFuse fB = fuse_in(bB), fA = fuse_in(bA);
        HPFG_RETURN_IF(wgrad(cB, d.bns[bA].raw, xf_of(bA, nullptr, 0.f), a, &fB));
        HPFG_RETURN_IF(acquire(c));
        HPFG_RETURN_IF(dgrad(cB, a, &fB, c, bA, nullptr, 0.f));                     // -> g(A) in c
        HPFG_RETURN_IF(wgrad(cA, p->cat[j], none, c, &fA));
        HPFG_RETURN_IF(dgrad(cA, c, &fA, p->dcat[j], -1, nullptr, 0.f));             // -> dcat (skip | upsampled)
        const int F = kFt[lvl], hl = H >> (lvl + 1), wl = W >> (lvl + 1);
        HPFG_RETURN_IF(acquire(b));
        HPFG_RETURN_IF(up_bwd<T>((const T *)p->dcat[j], (T *)b, N, hl, wl, F, s));   // -> dlow in b
        const int prev_bn = (j == 1) ? 9 : d.convs[cB - 3].bn;
        HPFG_RETURN_IF(wgrad(c1x1, d.bns[prev_bn].raw, xf_of(prev_bn, nullptr, 0.f), b, nullptr));
        HPFG_RETURN_IF(acquire(a));
        const bool plain = (j == 1 && dbottleneck != nullptr);      // an external gradient joins first: BatchNorm 9 goes the unfused way
        HPFG_RETURN_IF(dgrad(c1x1, b, nullptr, a, plain ? -1 : prev_bn, nullptr, 0.f));   // -> g(prev) (or dact(prev)) in a
        if (j == 3 || j == 1) {
            HPFG_RETURN_IF(join_side());
            HPFG_CUDA_CHECK(cudaEventRecord(p->bucket_ev[j == 3 ? 0 : 1], s));
        }
    }
    bool a_is_dact = false;
    if (dbottleneck) {
        HPFG_RETURN_IF(add_nchw_f32_to_nhwc<T>((T *)a, dbottleneck, N, H >> 4, W >> 4, kFt[4], s));
        a_is_dact = true;
    }
    // ---- encoder, down4 .. in_conv
    void *dpooled = nullptr;
    for (int l = 4; l >= 0; --l) {
        const int cA = 2 * l, cB = 2 * l + 1, bA = cA, bB = cB, h = H >> l, w = W >> l;
        const uint32_t *bits = p->saved_dropout ? p->dropbits[l] : nullptr;
        BnLayer &blB = d.bns[bB];
        if (l < 4) {  // g of the encoder feature's BatchNorm from: skip half of dcat + un-pooled grad from the level below
            HPFG_RETURN_IF(acquire(a));
            int P = 0;
            HPFG_RETURN_IF(skip_pool_bwd_gstat<T>((const T *)p->dcat[4 - l], (const T *)dpooled, (const T *)blB.raw, blB.st, (T *)a, N, h, w,
                                                  kFt[l], p->stats, (int)(p->stats_floats / (2 * blB.C)), &P, s));
            HPFG_RETURN_IF(bn_bwd_reduce(p->stats, P, blB.C, (int64_t)N * h * w, blB.st, grads + blB.g_off, grads + blB.b_off, acc, s));
        } else if (a_is_dact) {   // UNet_Plus: dact(9) with the neck's gradient added -> the two-pass BatchNorm backward, materialised
            HPFG_RETURN_IF(acquire(c));
            DropSpec ds{nullptr, 1.f};
            HPFG_RETURN_IF(bn_bwd<T>((const T *)a, (const T *)blB.raw, (T *)c, (int64_t)N * h * w, blB.C, blB.st, ds, p->stats,
                                     (int)(p->stats_floats / (2 * blB.C)), grads + blB.g_off, grads + blB.b_off, acc, s));
            a = c;
        }
        const bool fusedB = !(l == 4 && a_is_dact);
        TcBwdFuse fB = fuse_in(bB), fA = fuse_in(bA);
        HPFG_RETURN_IF(wgrad(cB, d.bns[bA].raw, xf_of(bA, bits, kEncDropout[l]), a, fusedB ? &fB : nullptr));
        HPFG_RETURN_IF(acquire(c));
        HPFG_RETURN_IF(dgrad(cB, a, fusedB ? &fB : nullptr, c, bA, bits, kEncDropout[l]));   // -> g(A) in c
        if (l == 0) {
            HPFG_RETURN_IF(wgrad(0, p->xpad, none, c, &fA));
        } else {
            HPFG_RETURN_IF(wgrad(cA, p->pooled[l], none, c, &fA));
            HPFG_RETURN_IF(dgrad(cA, c, &fA, p->g[3], -1, nullptr, 0.f));            // dpooled for level l-1
            dpooled = p->g[3];
        }
        if (l == 3 || l == 0) {
            HPFG_RETURN_IF(join_side());
            HPFG_CUDA_CHECK(cudaEventRecord(p->bucket_ev[l == 3 ? 2 : 3], s));
        }
    }
    return HPFG_OK;
}

template <typename T>
static int backward_impl(hpfg_unet_plan *p, const float *params, const float *dlogits, float *grads, int acc,
                         cudaStream_t s, const float *dbottleneck = nullptr) {
    UNetDesc &d = p->d;
    const int N = p->N, H = p->H, W = p->W;
    const bool tc = p->precision == HPFG_PREC_BF16;
    const LoadXform none{};
    auto xf_of = [&](int bn, const uint32_t *bits, float p_drop) {
        LoadXform xf{};
        xf.scale = d.bns[bn].st.scale;
        xf.shift = d.bns[bn].st.shift;
        xf.drop.bits = bits;
        xf.drop.inv_keep = bits ? 1.f / (1.f - p_drop) : 1.f;
        return xf;
    };
    // Weight gradients run on the plan's side stream, concurrently with the data-gradient chain: wgrad(layer) and
    // dgrad(layer) both consume the layer's raw gradient and are independent of each other.  The raw gradient lives
    // in one of two alternating buffers; before a buffer is rewritten the main stream waits for the wgrad that read it.
    cudaStream_t side = g_prof_on ? s : p->side;      // per-category timing (bench.py's roofline leg) wants serialized kernels
    p->wg_count = 0;
    p->reduced_pending[0] = p->reduced_pending[1] = false;      // (events of an earlier pass / capture are never waited on)
    void *dr[2] = {p->g[1], p->g[4]};
    bool dr_busy[2] = {false, false};
    int bi = 1;
    void *b = nullptr;
    auto next_b = [&]() -> int {
        bi ^= 1;
        b = dr[bi];
        if (dr_busy[bi]) HPFG_CUDA_CHECK(cudaStreamWaitEvent(s, p->ev_done[bi], 0));
        dr_busy[bi] = false;
        return HPFG_OK;
    };
    // weight gradient of conv `ci`: in (T NHWC, optionally transformed on load) x dout (T NHWC)
    // Order of submission per layer: the raw gradient is ready -> mark_ready() records the event on the main stream, the DATA
    // gradient is enqueued first (it is on the critical path and should get the SMs first), then the weight gradient on the side
    // stream, which waits for the event only (not for the data gradient) and fills in next to the following BatchNorm kernels.
    static const bool dgrad_first = !(getenv("HPFG_DGRAD_FIRST") && getenv("HPFG_DGRAD_FIRST")[0] == '0');      // A/B switch (profiles/)
    bool ready_marked = false;
    auto mark_ready = [&]() -> int {
        HPFG_CUDA_CHECK(cudaEventRecord(p->ev_ready, s));
        ready_marked = true;
        return HPFG_OK;
    };
    auto wgrad = [&](int ci, void *in, LoadXform xf, void *dout) -> int {
        ConvLayer &cv = d.convs[ci];
        if (!ready_marked) HPFG_CUDA_CHECK(cudaEventRecord(p->ev_ready, s));
        ready_marked = false;
        HPFG_CUDA_CHECK(cudaStreamWaitEvent(side, p->ev_ready, 0));
        bool done = false;
        if (tc) {
            // split-K partials go to one of two scratch halves; the fixed-order reduction runs on a third stream, so the next
            // weight gradient does not queue behind it
            static const bool split = !(getenv("HPFG_WG_REDUCE_STREAM") && getenv("HPFG_WG_REDUCE_STREAM")[0] == '0');   // A/B switch (profiles/)
            TcWgradStreams ws;
            const int k = p->wg_count++ & 1;
            if (split && !g_prof_on) {
                ws.scratch = k ? p->wscratch2 : p->wscratch;
                ws.reduce_stream = p->side2;
                ws.ev_partials = p->ev_partials[k];
                ws.ev_reduced = p->ev_reduced[k];
                if (p->reduced_pending[k]) HPFG_CUDA_CHECK(cudaStreamWaitEvent(side, p->ev_reduced[k], 0));   // reduce(n-2) has read this half
                p->reduced_pending[k] = true;
            }
            HPFG_RETURN_IF(tc_wgrad(p, ci, in, xf, dout, grads + cv.w_off, grads + cv.b_off, acc, &done, side, nullptr, &ws));
        }
        if (!done)
            HPFG_RETURN_IF((conv_ref_wgrad<T, T>(nhwc_view(in, cv.H, cv.W, cv.cin), nhwc_view(dout, cv.H, cv.W, cv.cout), N, cv.H,
                                                 cv.W, cv.cin, cv.cout, cv.ks, xf, p->wscratch, p->wscratch_floats,
                                                 grads + cv.w_off, grads + cv.b_off, acc, side)));
        for (int k = 0; k < 2; ++k)
            if (dout == dr[k]) {
                HPFG_CUDA_CHECK(cudaEventRecord(p->ev_done[k], side));
                dr_busy[k] = true;
            }
        return HPFG_OK;
    };
    auto join_side = [&]() -> int {     // main stream waits for every weight gradient (and its reduction) enqueued so far
        HPFG_CUDA_CHECK(cudaEventRecord(p->ev_join, side));
        HPFG_CUDA_CHECK(cudaStreamWaitEvent(s, p->ev_join, 0));
        for (int k = 0; k < 2; ++k)
            if (p->reduced_pending[k]) {
                HPFG_CUDA_CHECK(cudaStreamWaitEvent(s, p->ev_reduced[k], 0));
            }
        return HPFG_OK;
    };
    // data gradient of conv `ci`: din[.., cin] = conv(dout[.., cout], flipped weights)
    auto dgrad = [&](int ci, void *dout, void *din) -> int {
        ConvLayer &cv = d.convs[ci];
        bool done = false;
        if (tc) HPFG_RETURN_IF(tc_dgrad(p, ci, dout, din, &done, s));
        if (done) return HPFG_OK;
        return conv_ref_fprop<T, T>(nhwc_view(dout, cv.H, cv.W, cv.cout), nhwc_view(din, cv.H, cv.W, cv.cin), cv.wd, nullptr,
                                    N, cv.H, cv.W, cv.cout, cv.cin, cv.ks, none, nullptr, s);
    };
    auto bnb = [&](int bn, void *dact, void *draw, const uint32_t *bits, float p_drop) -> int {
        BnLayer &bl = d.bns[bn];
        DropSpec ds{bits, bits ? 1.f / (1.f - p_drop) : 1.f};
        return bn_bwd<T>((const T *)dact, (const T *)bl.raw, (T *)draw, (int64_t)N * bl.H * bl.W, bl.C, bl.st, ds, p->stats,
                         (int)(p->stats_floats / (2 * bl.C)), grads + bl.g_off, grads + bl.b_off, acc, s,
                         p->sync_bn ? p->sync_sums : nullptr);
    };

    // ---- out_conv
    if (tc) {   // dlogits: fp32 NCHW -> bf16 NHWC16, then regular tensor-core wgrad / dgrad with Cout padded to 16
        HPFG_RETURN_IF(pad_to_nhwc16(dlogits, p->dlpad, N, p->n_cls, H, W, s));
        if (dgrad_first) HPFG_RETURN_IF(mark_ready());
        if (dgrad_first) HPFG_RETURN_IF(dgrad(22, p->dlpad, p->g[0]));
        HPFG_RETURN_IF(wgrad(22, d.bns[17].raw, xf_of(17, nullptr, 0.f), p->dlpad));
        if (!dgrad_first) HPFG_RETURN_IF(dgrad(22, p->dlpad, p->g[0]));
    } else {
        ConvLayer &oc = d.convs[22];       // (all weight gradients share wscratch, so this one is ordered on the side stream too)
        HPFG_CUDA_CHECK(cudaEventRecord(p->ev_ready, s));
        HPFG_CUDA_CHECK(cudaStreamWaitEvent(side, p->ev_ready, 0));
        HPFG_RETURN_IF((conv_ref_wgrad<T, float>(nhwc_view(d.bns[17].raw, H, W, 16),
                                                 nchw_view(const_cast<float *>(dlogits), p->n_cls, H, W), N, H, W, 16, p->n_cls,
                                                 3, xf_of(17, nullptr, 0.f), p->wscratch, p->wscratch_floats,
                                                 grads + oc.w_off, grads + oc.b_off, acc, side)));
        HPFG_RETURN_IF((conv_ref_fprop<float, T>(nchw_view(const_cast<float *>(dlogits), p->n_cls, H, W),
                                                 nhwc_view(p->g[0], H, W, 16), oc.wd, nullptr, N, H, W, p->n_cls, 16, 3, none,
                                                 nullptr, s)));
    }
    // gradient scratch rotation: `a` holds the grad wrt the activated output of the block being processed
    void *a = p->g[0], *c = p->g[2];
    // ---- decoder, up4 .. up1
    for (int j = 4; j >= 1; --j) {
        const int lvl = 4 - j, c1x1 = 10 + 3 * (j - 1), cA = c1x1 + 1, cB = c1x1 + 2;
        const int bA = d.convs[cA].bn, bB = d.convs[cB].bn;
        HPFG_RETURN_IF(next_b());
        HPFG_RETURN_IF(bnb(bB, a, b, nullptr, 0.f));                                // a -> draw(B) in b
        if (dgrad_first) HPFG_RETURN_IF(mark_ready());
        if (dgrad_first) HPFG_RETURN_IF(dgrad(cB, b, c));                           // -> dact(A) in c
        HPFG_RETURN_IF(wgrad(cB, d.bns[bA].raw, xf_of(bA, nullptr, 0.f), b));
        if (!dgrad_first) HPFG_RETURN_IF(dgrad(cB, b, c));
        HPFG_RETURN_IF(next_b());
        HPFG_RETURN_IF(bnb(bA, c, b, nullptr, 0.f));                                // -> draw(A) in b
        if (dgrad_first) HPFG_RETURN_IF(mark_ready());
        if (dgrad_first) HPFG_RETURN_IF(dgrad(cA, b, p->dcat[j]));                  // -> dcat (skip | upsampled)
        HPFG_RETURN_IF(wgrad(cA, p->cat[j], none, b));
        if (!dgrad_first) HPFG_RETURN_IF(dgrad(cA, b, p->dcat[j]));
        const int F = kFt[lvl], hl = H >> (lvl + 1), wl = W >> (lvl + 1);
        HPFG_RETURN_IF(next_b());
        HPFG_RETURN_IF(up_bwd<T>((const T *)p->dcat[j], (T *)b, N, hl, wl, F, s));  // -> dlow in b
        const int prev_bn = (j == 1) ? 9 : d.convs[cB - 3].bn;
        if (dgrad_first) HPFG_RETURN_IF(mark_ready());
        if (dgrad_first) HPFG_RETURN_IF(dgrad(c1x1, b, c));                         // -> dact(prev) in c
        HPFG_RETURN_IF(wgrad(c1x1, d.bns[prev_bn].raw, xf_of(prev_bn, nullptr, 0.f), b));
        if (!dgrad_first) HPFG_RETURN_IF(dgrad(c1x1, b, c));
        std::swap(a, c);
        if (j == 3 || j == 1) {
            HPFG_RETURN_IF(join_side());
            HPFG_CUDA_CHECK(cudaEventRecord(p->bucket_ev[j == 3 ? 0 : 1], s));
        }
    }
    // external gradient wrt the activated bottleneck feature (UNet_Plus projection neck, model/unet.py:201): joins the
    // decoder's gradient before the down4 block is differentiated
    if (dbottleneck) HPFG_RETURN_IF(add_nchw_f32_to_nhwc<T>((T *)a, dbottleneck, N, H >> 4, W >> 4, kFt[4], s));
    // ---- encoder, down4 .. in_conv
    void *dpooled = nullptr;
    for (int l = 4; l >= 0; --l) {
        const int cA = 2 * l, cB = 2 * l + 1, bA = cA, bB = cB, h = H >> l, w = W >> l;
        const uint32_t *bits = p->saved_dropout ? p->dropbits[l] : nullptr;
        // grad wrt the encoder feature = skip half of dcat + un-pooled grad from the level below.  bf16 plans: the same kernel
        // already reads the feature's raw tensor, so it also does BatchNorm-backward pass 0 (stores g, writes the two sums):
        // one full tensor pass and one launch less per encoder level than skip_pool_bwd -> bn_bwd<0>.
        static const bool gstat_on = !(getenv("HPFG_SKIP_GSTAT") && getenv("HPFG_SKIP_GSTAT")[0] == '0');     // A/B switch (profiles/)
        const bool gstat = l < 4 && tc && !p->sync_bn && gstat_on;
        if (gstat) {
            BnLayer &bl = d.bns[bB];
            int P = 0;
            HPFG_RETURN_IF(skip_pool_bwd_gstat<T>((const T *)p->dcat[4 - l], (const T *)dpooled, (const T *)bl.raw, bl.st, (T *)a, N, h, w,
                                                  kFt[l], p->stats, (int)(p->stats_floats / (2 * bl.C)), &P, s));
            HPFG_RETURN_IF(bn_bwd_reduce(p->stats, P, bl.C, (int64_t)N * h * w, bl.st, grads + bl.g_off, grads + bl.b_off, acc, s));
            HPFG_RETURN_IF(next_b());
            HPFG_RETURN_IF(bn_bwd_from_g<T>((const T *)a, (const T *)bl.raw, (T *)b, (int64_t)N * h * w, bl.C, bl.st, s));
        } else {
            if (l < 4)
                HPFG_RETURN_IF(skip_pool_bwd<T>((const T *)p->dcat[4 - l], (const T *)dpooled, (const T *)d.bns[bB].raw, d.bns[bB].st,
                                                (T *)a, N, h, w, kFt[l], s));
            HPFG_RETURN_IF(next_b());
            HPFG_RETURN_IF(bnb(bB, a, b, nullptr, 0.f));                            // draw(B) in b
        }
        if (dgrad_first) HPFG_RETURN_IF(mark_ready());
        if (dgrad_first) HPFG_RETURN_IF(dgrad(cB, b, c));                           // dact(A) in c
        HPFG_RETURN_IF(wgrad(cB, d.bns[bA].raw, xf_of(bA, bits, kEncDropout[l]), b));
        if (!dgrad_first) HPFG_RETURN_IF(dgrad(cB, b, c));
        HPFG_RETURN_IF(next_b());
        HPFG_RETURN_IF(bnb(bA, c, b, bits, kEncDropout[l]));                        // draw(A) in b
        if (l == 0 && tc) {
            HPFG_RETURN_IF(wgrad(0, p->xpad, none, b));
        } else if (l == 0) {
            ConvLayer &cv = d.convs[0];
            HPFG_CUDA_CHECK(cudaEventRecord(p->ev_ready, s));
            HPFG_CUDA_CHECK(cudaStreamWaitEvent(side, p->ev_ready, 0));
            HPFG_RETURN_IF((conv_ref_wgrad<float, T>(nchw_view(const_cast<float *>(p->saved_x), p->in_ch, H, W),
                                                     nhwc_view(b, H, W, 16), N, H, W, p->in_ch, 16, 3, none, p->wscratch,
                                                     p->wscratch_floats, grads + cv.w_off, grads + cv.b_off, acc, side)));
        } else {
            if (dgrad_first) HPFG_RETURN_IF(mark_ready());
            if (dgrad_first) HPFG_RETURN_IF(dgrad(cA, b, p->g[3]));                 // dpooled for level l-1
            HPFG_RETURN_IF(wgrad(cA, p->pooled[l], none, b));
            if (!dgrad_first) HPFG_RETURN_IF(dgrad(cA, b, p->g[3]));
            dpooled = p->g[3];
        }
        if (l == 3 || l == 0) {
            HPFG_RETURN_IF(join_side());
            HPFG_CUDA_CHECK(cudaEventRecord(p->bucket_ev[l == 3 ? 2 : 3], s));
        }
    }
    return HPFG_OK;
}

}  // namespace hpfg

using namespace hpfg;

extern "C" const char *hpfg_last_error(void) { return g_error.c_str(); }
extern "C" int hpfg_version(void) { return 100; }
extern "C" int64_t hpfg_launch_count(void) { return g_launch_count; }

extern "C" int hpfg_profile_begin(void) {
    for (auto &r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof.clear();
    g_prof_on = true;
    return HPFG_OK;
}

extern "C" int hpfg_profile_end(double *ms_per_category_host, int64_t *calls_per_category_host) {
    g_prof_on = false;
    HPFG_CUDA_CHECK(cudaDeviceSynchronize());
    for (int c = 0; c < PROF_NCAT; ++c) {
        if (ms_per_category_host) ms_per_category_host[c] = 0.0;
        if (calls_per_category_host) calls_per_category_host[c] = 0;
    }
    for (auto &r : g_prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            if (ms_per_category_host) ms_per_category_host[r.cat] += ms;
            if (calls_per_category_host) calls_per_category_host[r.cat] += 1;
        }
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    cudaGetLastError();
    g_prof.clear();
    return HPFG_OK;
}

extern "C" int hpfg_unet_param_layout(int in_channels, int num_classes, int64_t *offsets_host, int64_t *sizes_host,
                                      int64_t *total_host) {
    HPFG_REQUIRE(in_channels >= 1 && num_classes >= 1, "hpfg_unet_param_layout: bad channel counts");
    UNetDesc d;
    describe_unet(in_channels, num_classes, 16, 16, d);
    if (offsets_host) std::memcpy(offsets_host, d.offsets, sizeof(d.offsets));
    if (sizes_host) std::memcpy(sizes_host, d.sizes, sizeof(d.sizes));
    if (total_host) *total_host = d.n_params;
    return HPFG_OK;
}

extern "C" int hpfg_unet_bn_layout(int in_channels, int num_classes, int64_t *bn_offsets_host, int64_t *bn_channels_host,
                                   int64_t *total_host) {
    HPFG_REQUIRE(in_channels >= 1 && num_classes >= 1, "hpfg_unet_bn_layout: bad channel counts");
    UNetDesc d;
    describe_unet(in_channels, num_classes, 16, 16, d);
    for (int i = 0; i < HPFG_NUM_BN; ++i) {
        if (bn_offsets_host) bn_offsets_host[i] = d.bns[i].run_off;
        if (bn_channels_host) bn_channels_host[i] = d.bns[i].C;
    }
    if (total_host) *total_host = d.n_bn_floats;
    return HPFG_OK;
}

extern "C" int hpfg_unet_plan_create(int batch, int in_channels, int num_classes, int height, int width, int precision,
                                     hpfg_unet_plan_t *plan_out) {
    HPFG_REQUIRE(plan_out, "hpfg_unet_plan_create: plan_out is null");
    HPFG_REQUIRE(batch >= 1 && in_channels >= 1 && in_channels <= 16, "hpfg_unet_plan_create: bad batch / in_channels");
    HPFG_REQUIRE(num_classes >= 1 && num_classes <= 64, "hpfg_unet_plan_create: bad num_classes");
    HPFG_REQUIRE(height >= 16 && width >= 16 && height % 16 == 0 && width % 16 == 0,
                 "hpfg_unet_plan_create: H and W must be multiples of 16 (four 2x2 poolings)");
    HPFG_REQUIRE(precision == HPFG_PREC_FP32 || precision == HPFG_PREC_BF16, "hpfg_unet_plan_create: unknown precision");
    int dev = 0;
    HPFG_CUDA_CHECK(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    HPFG_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) {
        set_error("hpfg_b200 is built for sm_100a only; device is sm_" + std::to_string(prop.major * 10 + prop.minor));
        return HPFG_ERR_UNSUPPORTED;
    }
    auto *p = new hpfg_unet_plan();
    p->N = batch; p->in_ch = in_channels; p->n_cls = num_classes; p->H = height; p->W = width; p->precision = precision;
    p->elt = precision == HPFG_PREC_FP32 ? 4 : 2;
    p->bwd_fusion = precision == HPFG_PREC_BF16 && bwd_fused_enabled();
    describe_unet(in_channels, num_classes, height, width, p->d);
    int64_t total = 0;
    carve(p, nullptr, total);
    if (cudaMalloc(&p->ws, (size_t)total) != cudaSuccess) {
        set_error("hpfg_unet_plan_create: cudaMalloc of " + std::to_string(total) + " bytes failed");
        cudaGetLastError();
        delete p;
        return HPFG_ERR_CUDA;
    }
    p->ws_bytes = total;
    carve(p, p->ws, total);
    for (int b = 0; b < kNumBuckets; ++b) HPFG_CUDA_CHECK(cudaEventCreateWithFlags(&p->bucket_ev[b], cudaEventDisableTiming));
    {   // the weight-gradient stream; HPFG_SIDE_PRIO=1 (profiles/ A/B) creates it at the highest priority
        int least = 0, greatest = 0;
        HPFG_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        const char *e = getenv("HPFG_SIDE_PRIO");
        HPFG_CUDA_CHECK(cudaStreamCreateWithPriority(&p->side, cudaStreamNonBlocking, (e && e[0] == '1') ? greatest : least));
    }
    HPFG_CUDA_CHECK(cudaEventCreateWithFlags(&p->ev_ready, cudaEventDisableTiming));
    for (int k = 0; k < 2; ++k) HPFG_CUDA_CHECK(cudaEventCreateWithFlags(&p->ev_done[k], cudaEventDisableTiming));
    for (int k = 0; k < 4; ++k) HPFG_CUDA_CHECK(cudaEventCreateWithFlags(&p->ev_slot[k], cudaEventDisableTiming));
    HPFG_CUDA_CHECK(cudaStreamCreateWithFlags(&p->side2, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) {
        HPFG_CUDA_CHECK(cudaEventCreateWithFlags(&p->ev_partials[k], cudaEventDisableTiming));
        HPFG_CUDA_CHECK(cudaEventCreateWithFlags(&p->ev_reduced[k], cudaEventDisableTiming));
    }
    HPFG_CUDA_CHECK(cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));
    if (precision == HPFG_PREC_BF16) {
        const int rc = tc_plan_init(p);
        if (rc != HPFG_OK) {
            hpfg_unet_plan_destroy(p);
            return rc;
        }
    }
    *plan_out = p;
    return HPFG_OK;
}

extern "C" int hpfg_unet_plan_destroy(hpfg_unet_plan_t p) {
    if (!p) return HPFG_OK;
    cudaDeviceSynchronize();
    if (p->precision == HPFG_PREC_BF16) tc_plan_free(p);
    for (int b = 0; b < kNumBuckets; ++b)
        if (p->bucket_ev[b]) cudaEventDestroy(p->bucket_ev[b]);
    if (p->ev_ready) cudaEventDestroy(p->ev_ready);
    if (p->ev_join) cudaEventDestroy(p->ev_join);
    for (int k = 0; k < 2; ++k)
        if (p->ev_done[k]) cudaEventDestroy(p->ev_done[k]);
    for (int k = 0; k < 4; ++k)
        if (p->ev_slot[k]) cudaEventDestroy(p->ev_slot[k]);
    if (p->side) cudaStreamDestroy(p->side);
    if (p->side2) cudaStreamDestroy(p->side2);
    for (int k = 0; k < 2; ++k) {
        if (p->ev_partials[k]) cudaEventDestroy(p->ev_partials[k]);
        if (p->ev_reduced[k]) cudaEventDestroy(p->ev_reduced[k]);
    }
    if (p->ws) cudaFree(p->ws);
    delete p;
    return HPFG_OK;
}

extern "C" int64_t hpfg_unet_plan_workspace_bytes(hpfg_unet_plan_t p) { return p ? p->ws_bytes : 0; }

extern "C" int hpfg_set_allreduce_hook(void (*fn)(void *, void *, int64_t, int, void *), void *ctx, int world_size) {
    HPFG_REQUIRE(world_size >= 1 && (fn || world_size == 1), "hpfg_set_allreduce_hook: a hook is required when world_size > 1");
    g_sync_fn = fn;
    g_sync_ctx = ctx;
    g_sync_world = world_size;
    return HPFG_OK;
}

extern "C" int hpfg_ssl_loss_set_global_sums(int enabled) {
    g_loss_global_sums = enabled != 0;
    return HPFG_OK;
}

extern "C" int hpfg_unet_plan_set_sync_bn(hpfg_unet_plan_t p, int enabled) {
    HPFG_REQUIRE(p, "hpfg_unet_plan_set_sync_bn: null plan");
    p->sync_bn = enabled != 0;
    return HPFG_OK;
}

extern "C" int hpfg_unet_plan_set_forward_ctas(hpfg_unet_plan_t p, int ctas) {
    HPFG_REQUIRE(p && ctas >= 0 && ctas <= kNumSMs, "hpfg_unet_plan_set_forward_ctas: 0 (no cap) .. 148");
    p->fwd_ctas = ctas;
    return HPFG_OK;
}

extern "C" int hpfg_unet_plan_set_bwd_fusion(hpfg_unet_plan_t p, int enabled) {
    HPFG_REQUIRE(p, "hpfg_unet_plan_set_bwd_fusion: null plan");
    HPFG_REQUIRE(!enabled || p->precision == HPFG_PREC_BF16, "hpfg_unet_plan_set_bwd_fusion: only bf16 plans have the fused backward");
    p->bwd_fusion = enabled != 0;
    return HPFG_OK;
}

extern "C" int hpfg_unet_forward(hpfg_unet_plan_t p, const float *params, float *bn_running, int64_t *bn_counters,
                                 const float *x, float *logits, int training, int no_dropout, int save_for_backward,
                                 uint64_t dropout_seed, uint64_t dropout_offset, const uint8_t *const *dropout_masks_host,
                                 void *stream) {
    HPFG_REQUIRE(p && params && bn_running && x && logits, "hpfg_unet_forward: null argument");
    cudaStream_t s = (cudaStream_t)stream;
    if (p->precision == HPFG_PREC_FP32)
        return forward_impl<float>(p, params, bn_running, bn_counters, x, logits, training, no_dropout, save_for_backward,
                                   dropout_seed, dropout_offset, dropout_masks_host, s);
    return forward_impl<bf16>(p, params, bn_running, bn_counters, x, logits, training, no_dropout, save_for_backward,
                              dropout_seed, dropout_offset, dropout_masks_host, s);
}

extern "C" int hpfg_unet_backward(hpfg_unet_plan_t p, const float *params, const float *dlogits, float *grads,
                                  int accumulate, void *stream) {
    HPFG_REQUIRE(p && params && dlogits && grads, "hpfg_unet_backward: null argument");
    HPFG_REQUIRE(p->saved, "hpfg_unet_backward: no training forward with save_for_backward on this plan");
    cudaStream_t s = (cudaStream_t)stream;
    if (p->precision == HPFG_PREC_FP32) return backward_impl<float>(p, params, dlogits, grads, accumulate, s);
    if (p->bwd_fusion && !p->sync_bn) return backward_fused(p, params, dlogits, grads, accumulate, s, nullptr);
    return backward_impl<bf16>(p, params, dlogits, grads, accumulate, s);
}

extern "C" int hpfg_unet_backward_ex(hpfg_unet_plan_t p, const float *params, const float *dlogits,
                                     const float *dbottleneck, float *grads, int accumulate, void *stream) {
    HPFG_REQUIRE(p && params && dlogits && grads, "hpfg_unet_backward_ex: null argument");
    HPFG_REQUIRE(p->saved, "hpfg_unet_backward_ex: no training forward with save_for_backward on this plan");
    cudaStream_t s = (cudaStream_t)stream;
    if (p->precision == HPFG_PREC_FP32) return backward_impl<float>(p, params, dlogits, grads, accumulate, s, dbottleneck);
    if (p->bwd_fusion && !p->sync_bn) return backward_fused(p, params, dlogits, grads, accumulate, s, dbottleneck);
    return backward_impl<bf16>(p, params, dlogits, grads, accumulate, s, dbottleneck);
}

extern "C" int hpfg_unet_bottleneck(hpfg_unet_plan_t p, float *feature_nchw, void *stream) {
    HPFG_REQUIRE(p && feature_nchw, "hpfg_unet_bottleneck: null argument");
    HPFG_REQUIRE(p->saved_x != nullptr, "hpfg_unet_bottleneck: no forward has run on this plan");
    cudaStream_t s = (cudaStream_t)stream;
    BnLayer &b = p->d.bns[9];
    if (p->precision == HPFG_PREC_FP32)
        return act_nhwc_to_nchw_f32<float>((const float *)b.raw, b.st, feature_nchw, p->N, p->H >> 4, p->W >> 4, kFt[4], s);
    return act_nhwc_to_nchw_f32<bf16>((const bf16 *)b.raw, b.st, feature_nchw, p->N, p->H >> 4, p->W >> 4, kFt[4], s);
}

extern "C" int hpfg_unet_num_buckets(hpfg_unet_plan_t p) { return p ? kNumBuckets : 0; }

extern "C" int hpfg_unet_bucket_range(hpfg_unet_plan_t p, int bucket, int64_t *offset_host, int64_t *count_host) {
    HPFG_REQUIRE(p && bucket >= 0 && bucket < kNumBuckets, "hpfg_unet_bucket_range: bad bucket");
    const int64_t lo = p->d.bucket_begin[bucket + 1], hi = p->d.bucket_begin[bucket];
    if (offset_host) *offset_host = lo;
    if (count_host) *count_host = hi - lo;
    return HPFG_OK;
}

extern "C" int hpfg_unet_bucket_layout(int in_channels, int num_classes, int64_t *offsets_host, int64_t *counts_host) {
    HPFG_REQUIRE(in_channels >= 1 && num_classes >= 1 && offsets_host && counts_host, "hpfg_unet_bucket_layout: bad arguments");
    UNetDesc d;
    describe_unet(in_channels, num_classes, 16, 16, d);
    for (int b = 0; b < kNumBuckets; ++b) {
        offsets_host[b] = d.bucket_begin[b + 1];
        counts_host[b] = d.bucket_begin[b] - d.bucket_begin[b + 1];
    }
    return HPFG_OK;
}

extern "C" int hpfg_unet_bucket_wait(hpfg_unet_plan_t p, int bucket, void *comm_stream) {
    HPFG_REQUIRE(p && bucket >= 0 && bucket < kNumBuckets, "hpfg_unet_bucket_wait: bad bucket");
    HPFG_CUDA_CHECK(cudaStreamWaitEvent((cudaStream_t)comm_stream, p->bucket_ev[bucket], 0));
    return HPFG_OK;
}

extern "C" int hpfg_unet_debug_tap(hpfg_unet_plan_t p, const char *name, float *out_nchw, int64_t capacity, void *stream) {
    HPFG_REQUIRE(p && name && out_nchw, "hpfg_unet_debug_tap: null argument");
    cudaStream_t s = (cudaStream_t)stream;
    const std::string nm(name);
    const bool f32 = p->precision == HPFG_PREC_FP32;
    auto copy = [&](void *src, int h, int w, int c, const float *bias) -> int {
        HPFG_REQUIRE((int64_t)p->N * h * w * c <= capacity, "hpfg_unet_debug_tap: output buffer too small");
        return f32 ? nhwc_to_nchw_f32<float>((const float *)src, out_nchw, p->N, h, w, c, bias, s)
                   : nhwc_to_nchw_f32<bf16>((const bf16 *)src, out_nchw, p->N, h, w, c, bias, s);
    };
    for (auto &cv : p->d.convs)
        if (cv.name == nm && cv.bn >= 0) return copy(p->d.bns[cv.bn].raw, cv.H, cv.W, cv.cout, nullptr);
    for (int j = 1; j < 5; ++j) {
        const int lvl = 4 - j;
        const std::string pre = "decoder.up" + std::to_string(j);
        if (nm == pre + ".conv1x1") return copy(p->low[j], p->H >> (lvl + 1), p->W >> (lvl + 1), kFt[lvl], nullptr);
        if (nm == pre + ".cat") return copy(p->cat[j], p->H >> lvl, p->W >> lvl, 2 * kFt[lvl], nullptr);
    }
    for (int l = 1; l < 5; ++l)
        if (nm == "pooled" + std::to_string(l)) return copy(p->pooled[l], p->H >> l, p->W >> l, kFt[l - 1], nullptr);
    set_error("hpfg_unet_debug_tap: unknown tap '" + nm + "'");
    return HPFG_ERR_INVALID;
}

extern "C" int hpfg_unet_forward_dv(hpfg_unet_plan_t p, const float *params, float *bn_running, int64_t *bn_counters,
                                    const float *x, float *logits, int training, int no_dropout, int save_for_backward,
                                    uint64_t dropout_seed, const uint64_t *dropout_offset_dev,
                                    const uint8_t *const *dropout_masks_host, void *stream) {
    HPFG_REQUIRE(p && params && bn_running && x && logits && dropout_offset_dev, "hpfg_unet_forward_dv: null argument");
    cudaStream_t s = (cudaStream_t)stream;
    if (p->precision == HPFG_PREC_FP32)
        return forward_impl<float>(p, params, bn_running, bn_counters, x, logits, training, no_dropout, save_for_backward,
                                   dropout_seed, 0, dropout_masks_host, s, dropout_offset_dev);
    return forward_impl<bf16>(p, params, bn_running, bn_counters, x, logits, training, no_dropout, save_for_backward,
                              dropout_seed, 0, dropout_masks_host, s, dropout_offset_dev);
}
