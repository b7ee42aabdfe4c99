// Shared device/host helpers for the hpfg_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <utility>

#include "../../include/hpfg_b200.h"

namespace hpfg {

// ---- error plumbing (no exceptions across the C ABI) --------------------------------------------------
void set_error(const std::string &msg);
extern int64_t g_launch_count;

#define HPFG_CUDA_CHECK(expr)                                                                      \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            ::hpfg::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                              ":" + std::to_string(__LINE__) + ")");                               \
            return HPFG_ERR_CUDA;                                                                  \
        }                                                                                          \
    } while (0)

#define HPFG_RETURN_IF(expr)                \
    do {                                    \
        int _rc = (expr);                   \
        if (_rc != HPFG_OK) return _rc;     \
    } while (0)

#define HPFG_REQUIRE(cond, msg)                     \
    do {                                            \
        if (!(cond)) {                              \
            ::hpfg::set_error(std::string(msg));    \
            return HPFG_ERR_INVALID;                \
        }                                           \
    } while (0)

// counts launches and surfaces launch-configuration errors immediately
#define HPFG_LAUNCH_CHECK()                 \
    do {                                    \
        ++::hpfg::g_launch_count;           \
        HPFG_CUDA_CHECK(cudaGetLastError()); \
    } while (0)

constexpr float kLeakySlope = 0.01f;   // nn.LeakyReLU() default (model/unet.py:20,24)
constexpr float kBnEps = 1e-5f;        // nn.BatchNorm2d defaults
constexpr float kBnMomentum = 0.1f;
constexpr int kNumSMs = 148;

// ---- dtype helpers ------------------------------------------------------------------------------------
using bf16 = __nv_bfloat16;

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float leaky(float v) { return v > 0.f ? v : kLeakySlope * v; }
__device__ __forceinline__ float leaky_grad(float v) { return v > 0.f ? 1.f : kLeakySlope; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- Philox4x32-10 (counter-based; keyed by (seed), counter = (element_index/4, offset)) ----------------
__device__ __forceinline__ uint4 philox4x32_10(uint2 key, uint4 ctr) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------------------
// Every kernel of this library is launched with cudaLaunchAttributeProgrammaticStreamSerialization and runs
// pdl_prologue() (or launch_dependents early + wait after its private setup) before it touches global memory the
// previous kernel may still be writing: the next kernel's launch latency and prologue overlap this kernel's tail.
// griddepcontrol.wait returns only when the prerequisite grid has completed and flushed, so correctness never
// depends on where launch_dependents sits.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
    pdl_launch_dependents();
    pdl_wait();
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// Plain (fully stream-ordered) launch: used for the kernel that follows a host-enqueued all-reduce in exact-global mode -- a
// programmatic dependent launch there may start on the trigger of the kernel BEFORE the collective's event wait.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_plain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cfg.attrs = nullptr;
    cfg.numAttrs = 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ---- division by a runtime constant without the (very slow) 64-bit integer divide ------------------------------
// q = (umulhi(n, m) + n) >> s, exact for n < 2^31 (all element counts here are checked against that on the host).
struct FastDiv {
    uint32_t d, m, s;
};
inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.d = d;
    f.s = 0;
    while ((1u << f.s) < d) ++f.s;
    f.m = (uint32_t)(((uint64_t(1) << 32) * ((uint64_t(1) << f.s) - d)) / d + 1);
    return f;
}
__device__ __forceinline__ uint32_t fast_div(uint32_t n, const FastDiv &f) { return (__umulhi(n, f.m) + n) >> f.s; }
__device__ __forceinline__ void fast_divmod(uint32_t n, const FastDiv &f, uint32_t &q, uint32_t &r) {
    q = fast_div(n, f);
    r = n - q * f.d;
}

// Persistent-CTA budget of the tensor-core kernels per kind (0 forward convs, 1 data gradients, 2 weight gradients).  A kernel
// with one 200 KB CTA per SM on all 148 SMs leaves no room for the kernels of the other stream; capping the grids lets the two
// streams of a step (student | teacher forward, dgrad chain | weight gradients) run on disjoint SMs at the same time.
int tc_cta_cap(int kind);

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- exact-global data-parallel mode (SURVEY 8e caveats 1-3): the library has no communicator of its own; the host registers a
// sum-all-reduce hook (hpfg_set_allreduce_hook) that is called, in stream order, on the few small device buffers that carry
// batch-wide sums: BatchNorm statistics (forward and backward) and the loss accumulators.
int sync_world();
int sync_allreduce(void *device_ptr, int64_t count, bool is_double, cudaStream_t s);   // HPFG_OK when no hook is registered and world == 1
extern bool g_loss_global_sums;

// ---- optional per-category device timing (bench.py's roofline leg): CUDA events around each host launcher ----
enum ProfCat { PROF_CONV_TC = 0, PROF_CONV_CUDA = 1, PROF_WGRAD = 2, PROF_GLUE = 3, PROF_LOSS = 4, PROF_OPTIM = 5, PROF_PACK = 6, PROF_WGRAD_TC = 7, PROF_NCAT = 8 };
extern bool g_prof_on;
void prof_push(int cat, cudaStream_t s, bool begin);
struct ProfScope {
    int cat;
    cudaStream_t s;
    ProfScope(int c, cudaStream_t st) : cat(c), s(st) { if (g_prof_on) prof_push(cat, s, true); }
    ~ProfScope() { if (g_prof_on) prof_push(cat, s, false); }
};

}  // namespace hpfg
