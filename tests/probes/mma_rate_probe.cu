// Probe (GPU): cycles per tcgen05.mma (cta_group::1, kind::f16, K=16, SWIZZLE_NONE smem operands) for several
// (M, N, operand-major) shapes: 512 back-to-back MMAs issued by one elected lane, timed with clock64 around
// issue + commit + mbarrier wait.
#include <cstdio>
#include <cuda_bf16.h>
#include "tc_ptx.cuh"
using namespace hpfg;

template <int M, int N, int MAJOR, int ROT>
__global__ void probe(long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0;
    if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_barrier_init(); }
    if (threadIdx.x < 32) ptx::tmem_alloc(ptx::smem_u32(&tslot), 512);
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tslot;
    if (threadIdx.x < 32) {
        const uint32_t idesc = ptx::umma_idesc_bf16(M, N, MAJOR, MAJOR);
        const uint32_t a = ptx::smem_u32(smem), b = a;   // operand values are irrelevant; keep every read inside the 64 KB
        // K-major: LBO = rows*16 (K halves), SBO = 128;  MN-major: LBO = 128 (K groups), SBO = 2048 (MN groups)
        const uint64_t ad = MAJOR ? ptx::umma_desc(a, 128, 2048) : ptx::umma_desc(a, M * 16, 128);
        const uint64_t bd = MAJOR ? ptx::umma_desc(b, 128, 1024) : ptx::umma_desc(b, N * 16, 128);
        long long t0 = clock64();
        if (ptx::elect_one()) {
#pragma unroll 8
            for (int i = 0; i < 512; ++i) ptx::umma_bf16(tmem + (ROT ? (i % (512 / N > 8 ? 8 : 512 / N)) * N : 0), ad, bd, idesc, 1);
            ptx::umma_commit(ptx::smem_u32(&bar));
        }
        __syncwarp();
        ptx::mbar_wait(ptx::smem_u32(&bar), 0, 99);
        long long t1 = clock64();
        if (threadIdx.x == 0) *out = t1 - t0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) ptx::tmem_dealloc(tmem, 512);
}

template <int M, int N, int MAJOR, int ROT>
static void run(long long *d) {
    cudaFuncSetAttribute(probe<M, N, MAJOR, ROT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    long long h = 0;
    for (int i = 0; i < 2; ++i) probe<M, N, MAJOR, ROT><<<1, 128, 64 * 1024>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("M=%3d N=%3d %s-major %s: %6.1f cycles / MMA  (%s)\n", M, N, MAJOR ? "MN" : "K ", ROT ? "rotating accumulators" : "same accumulator     ", (double)h / 512, cudaGetErrorString(e));
}

int main() {
    long long *d;
    cudaMalloc(&d, 8);
    run<128, 16, 0, 0>(d); run<128, 16, 0, 1>(d); run<128, 32, 0, 1>(d); run<128, 64, 0, 1>(d); run<128, 128, 0, 1>(d); run<128, 256, 0, 1>(d);
    run<64, 16, 0, 1>(d); run<64, 64, 0, 1>(d); run<64, 128, 0, 1>(d); run<64, 256, 0, 1>(d);
    run<128, 16, 1, 1>(d); run<64, 16, 1, 1>(d); run<64, 32, 1, 1>(d); run<64, 48, 1, 1>(d); run<64, 144, 1, 1>(d); run<64, 256, 1, 1>(d); run<128, 144, 1, 1>(d);
    return 0;
}
