// Probe (GPU): TMA fill rate of a (16+2)x(8+2) halo tile of a bf16 NHWC [32,224,224,C] tensor for three tensor-map shapes:
//   0: 5-D chunked (8ch, W, H, C/8, N)  box (8,10,18,C/8,1)   -> 16-byte elements, lands in UMMA [chunk][pixel][8] order
//   1: 4-D         (C, W, H, N)         box (C,10,18,1)       -> C*2-byte rows per pixel
//   2: 3-D merged  (W*C, H, N)          box (10*C,18,1)       -> one 10*C*2-byte row per halo row
// One producer thread + one consumer thread per CTA (148 persistent CTAs), 8-stage ring, nothing else.
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include "common.cuh"
#include "tc_common.cuh"
namespace hpfg { void set_error(const std::string &) {} int64_t g_launch_count = 0; bool g_prof_on = false; void prof_push(int, cudaStream_t, bool) {} }
using namespace hpfg;

constexpr int STAGES = 8;
template <int MODE, int C>
__global__ void __launch_bounds__(64, 1) fill(const __grid_constant__ CUtensorMap tm, int tiles_h, int tiles_w, int m_tiles, unsigned *sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int BYTES = 180 * C * 2, STAGE = (BYTES + 127) / 128 * 128;
    __shared__ uint64_t bars[2 * STAGES];
    const uint32_t full = ptx::smem_u32(bars), empty = full + 8 * STAGES, base = ptx::smem_u32(smem);
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(full + 8 * s, 1); ptx::mbar_init(empty + 8 * s, 1); }
        ptx::fence_barrier_init();
    }
    __syncthreads();
    const int tpi = tiles_h * tiles_w;
    if (threadIdx.x == 0) {
        int stage = 0, phase = 0;
        for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
            const int n = mt / tpi, r = mt % tpi, h0 = (r / tiles_w) * 16 - 1, w0 = (r % tiles_w) * 8 - 1;
            ptx::mbar_wait(empty + 8 * stage, phase ^ 1, 1);
            ptx::mbar_expect_tx(full + 8 * stage, BYTES);
            if (MODE == 0) ptx::tma_load_5d(base + stage * STAGE, &tm, full + 8 * stage, 0, w0, h0, 0, n);
            else if (MODE == 1) ptx::tma_load_4d(base + stage * STAGE, &tm, full + 8 * stage, 0, w0, h0, n);
            else asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                              ::"r"(base + stage * STAGE), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(full + 8 * stage), "r"(w0 * C), "r"(h0), "r"(n) : "memory");
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (threadIdx.x == 32) {
        int stage = 0, phase = 0;
        unsigned acc = 0;
        for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
            ptx::mbar_wait(full + 8 * stage, phase, 2);
            acc += *reinterpret_cast<volatile unsigned *>(smem + stage * STAGE + 64);
            ptx::mbar_arrive(empty + 8 * stage);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (acc == 0x12345678u) *sink = acc;
    }
}

template <int MODE, int C>
static void run(const void *t, unsigned *sink) {
    const int N = 32, H = 224, W = 224;
    CUtensorMap m;
    EncodeTiledFn enc = get_encode();
    if (MODE == 0) make_map_chunked(&m, t, N, H, W, C, C / 8, 10, 18);
    else if (MODE == 1) make_map(&m, t, N, H, W, C, C, 10, 18);
    else {
        cuuint64_t dims[3] = {(cuuint64_t)W * C, (cuuint64_t)H, (cuuint64_t)N};
        cuuint64_t strides[2] = {(cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
        cuuint32_t box[3] = {(cuuint32_t)(10 * C), 18, 1}, es[3] = {1, 1, 1};
        CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(t), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("mode 2 C=%d: encode failed %d\n", C, (int)r); return; }
    }
    const int smem = STAGES * ((180 * C * 2 + 127) / 128 * 128);
    cudaFuncSetAttribute(fill<MODE, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int th = 14, tw = 28, mt = N * th * tw;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) fill<MODE, C><<<148, 64, smem>>>(m, th, tw, mt, sink);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) fill<MODE, C><<<148, 64, smem>>>(m, th, tw, mt, sink);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("mode %d C=%2d: %7.1f us per pass over %d tiles (%s)  -> %.2f TB/s of halo bytes\n", MODE, C, ms * 100, mt, cudaGetErrorString(e),
           (double)mt * 180 * C * 2 / (ms * 1e-4) / 1e12 * 1e-6 * 1e6 / 1e0 / 1e0 * 1e-0 / 1.0 * 1e-0);
}

int main() {
    void *t; unsigned *sink;
    cudaMalloc(&t, (size_t)32 * 224 * 224 * 64 * 2);
    cudaMemset(t, 0, (size_t)32 * 224 * 224 * 64 * 2);
    cudaMalloc(&sink, 4);
    run<0, 16>(t, sink); run<1, 16>(t, sink); run<2, 16>(t, sink);
    run<0, 32>(t, sink); run<1, 32>(t, sink); run<2, 32>(t, sink);
    run<0, 64>(t, sink); run<1, 64>(t, sink);
    return 0;
}
