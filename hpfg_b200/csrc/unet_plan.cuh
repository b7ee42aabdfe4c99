// Host-side plan for one UNet instance at one batch shape: network description, workspace carve-up and the
// forward / backward kernel schedules (model/unet.py:61-117,155-175 wiring).
#pragma once
#include <string>
#include <vector>

#include "common.cuh"
#include "conv_ref.cuh"
#include "glue.cuh"

namespace hpfg {

constexpr int kFt[5] = {16, 32, 64, 128, 256};                  // model/unet.py:160
constexpr float kEncDropout[5] = {0.05f, 0.1f, 0.2f, 0.3f, 0.5f};   // model/unet.py:161
constexpr int kNumConv = 23;
constexpr int kNumParams = 82;
constexpr int kNumBuckets = 4;

struct ConvLayer {
    std::string name;
    int cin, cout, ks, H, W;
    int64_t w_off, b_off;        // element offsets into the flat parameter buffer
    int bn;                      // index of the BatchNorm that follows, or -1
    float *wf = nullptr;         // fp32 packed [tap][ci][co]              (CUDA-core fprop)
    float *wd = nullptr;         // fp32 packed [tap'][co][ci], flipped    (CUDA-core dgrad)
};

struct BnLayer {
    int C;
    int64_t g_off, b_off;        // gamma / beta offsets in the flat parameter buffer
    int64_t run_off;             // running_mean offset in bn_running (running_var at +C)
    int conv;                    // producing conv layer
    BnState st;
    void *raw = nullptr;         // [N,H,W,C] bias-free conv output (T)
    int H, W;
};

struct UNetDesc {
    int in_ch, n_cls;
    std::vector<ConvLayer> convs;   // 23, registration order
    std::vector<BnLayer> bns;       // 18, registration order
    int64_t n_params = 0, n_bn_floats = 0;
    int64_t offsets[kNumParams], sizes[kNumParams];
    int64_t bucket_begin[kNumBuckets + 1];   // element offsets, bucket b = [begin[b+1], begin[b]) counted from the tail
};
void describe_unet(int in_ch, int n_cls, int H, int W, UNetDesc &d);

}  // namespace hpfg

struct hpfg_unet_plan {
    int N, in_ch, n_cls, H, W, precision;
    hpfg::UNetDesc d;
    size_t elt;                       // sizeof(T)
    char *ws = nullptr;               // one allocation
    int64_t ws_bytes = 0;
    // activations / scratch (T unless noted)
    void *pooled[5] = {};             // [1..4]
    void *low[5] = {};                // [1..4] conv1x1 outputs
    void *cat[5] = {};                // [1..4]
    void *dcat[5] = {};               // [1..4]
    void *g[5] = {};                  // gradient scratch, N*H*W*16 elements each ([1] and [4]: alternating raw-gradient buffers)
    void *xpad = nullptr, *dlpad = nullptr;   // bf16 NHWC16 copies of the network input / dlogits (bf16 plans)
    uint32_t *dropbits[5] = {};
    float *stats = nullptr;           // BN statistics partials
    int64_t stats_floats = 0;
    float *wscratch = nullptr;        // wgrad split partials (bf16 plans: two halves, used alternately)
    float *wscratch2 = nullptr;
    int64_t wscratch_floats = 0;
    float *bnmem = nullptr;           // BnState arrays
    bool saved = false, saved_dropout = false;
    bool sync_bn = false;             // BatchNorm statistics (forward and backward) over the batch of ALL ranks (hpfg_unet_plan_set_sync_bn)
    double *sync_sums = nullptr;
    int fwd_ctas = 0;                 // > 0: cap on the persistent CTAs of the forward convolutions (two concurrent forwards share the SMs)
    bool bwd_fusion = false;          // bf16: BatchNorm backward folded into the dgrad / wgrad kernels (unet_plan.cu: backward_fused)
    const float *saved_x = nullptr;
    cudaEvent_t bucket_ev[hpfg::kNumBuckets] = {};
    // weight gradients run on a side stream, concurrently with the data-gradient chain of the same layer
    cudaStream_t side = nullptr;
    cudaStream_t side2 = nullptr;     // the split-K reductions of the weight gradients (tiny kernels: off the weight-gradient stream's chain)
    cudaEvent_t ev_partials[2] = {}, ev_reduced[2] = {};
    bool reduced_pending[2] = {false, false};
    int wg_count = 0;
    cudaEvent_t ev_ready = nullptr, ev_join = nullptr, ev_done[2] = {}, ev_slot[4] = {};
    void *tc = nullptr;               // tensor-core path state (conv_tc.cu), bf16 plans only
};
