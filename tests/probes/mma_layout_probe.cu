// Probe (GPU): cycles per tcgen05.mma (M=128, K=16, bf16, SWIZZLE_NONE K-major A) as a function of the A-operand
// descriptor geometry used by the implicit-GEMM conv: start-address offset (tap shift), SBO (halo row pitch) and LBO
// (channel-chunk stride).  Answers: do 16-byte-shifted / non-128-byte-pitched core matrices slow the operand fetch?
#include <cstdio>
#include <cuda_bf16.h>
#include "tc_ptx.cuh"
using namespace hpfg;

template <int N>
__global__ void probe(long long *out, uint32_t a_off, uint32_t a_lbo, uint32_t a_sbo, int mn_major) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0;
    if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_barrier_init(); }
    if (threadIdx.x < 32) ptx::tmem_alloc(ptx::smem_u32(&tslot), 512);
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tslot;
    if (threadIdx.x < 32) {
        const uint32_t idesc = ptx::umma_idesc_bf16(128, N, mn_major, 0);
        const uint32_t a = ptx::smem_u32(smem) + a_off, b = ptx::smem_u32(smem) + 190 * 1024;
        const uint64_t ad = ptx::umma_desc(a, a_lbo, a_sbo);
        const uint64_t bd = ptx::umma_desc(b, N * 16, 128);
        long long t0 = clock64();
        if (ptx::elect_one()) {
#pragma unroll 8
            for (int i = 0; i < 512; ++i) ptx::umma_bf16(tmem + (i % 8) * N, ad, bd, idesc, 1);
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

template <int N>
static void run(long long *d, const char *what, uint32_t off, uint32_t lbo, uint32_t sbo, int mn = 0) {
    cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    long long h = 0;
    for (int i = 0; i < 2; ++i) probe<N><<<1, 128, 200 * 1024>>>(d, off, lbo, sbo, mn);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("N=%3d %-44s off=%5u lbo=%6u sbo=%5u : %6.1f cycles / MMA (%s)\n", N, what, off, lbo, sbo, (double)h / 512, cudaGetErrorString(e));
}

int main() {
    long long *d;
    cudaMalloc(&d, 8);
    run<16>(d, "canonical dense", 0, 2048, 128);
    run<16>(d, "canonical dense, shifted 16B", 16, 2048, 128);
    run<16>(d, "canonical dense, shifted 64B", 64, 2048, 128);
    run<16>(d, "conv MT=4 pitch 34px, tap (0,0)", 0, 9792, 544);
    run<16>(d, "conv MT=4 pitch 34px, tap (0,1)", 16, 9792, 544);
    run<16>(d, "conv MT=4 pitch 34px, tap (1,0)", 544, 9792, 544);
    run<16>(d, "conv MT=4 pitch 34px, tap (1,1)", 560, 9792, 544);
    run<16>(d, "conv pitch 40px (640B), tap (0,0)", 0, 11520, 640);
    run<16>(d, "conv pitch 40px (640B), tap (0,1)", 16, 11520, 640);
    run<16>(d, "conv pitch 40px (640B), tap (0,2)", 32, 11520, 640);
    run<16>(d, "conv pitch 40px (640B), tap (1,1)", 656, 11520, 640);
    run<16>(d, "conv pitch 36px (576B), tap (0,0)", 0, 10368, 576);
    run<16>(d, "conv pitch 36px (576B), tap (0,1)", 16, 10368, 576);
    run<16>(d, "conv MT=1 pitch 10px (160B), tap (0,0)", 0, 2880, 160);
    run<16>(d, "conv MT=1 pitch 10px (160B), tap (1,1)", 176, 2880, 160);
    run<16>(d, "conv pitch 16px (256B), tap (0,0)", 0, 4608, 256);
    run<16>(d, "conv pitch 16px (256B), tap (1,1)", 272, 4608, 256);
    run<32>(d, "canonical dense", 0, 2048, 128);
    run<32>(d, "conv MT=4 pitch 34px, tap (1,1)", 560, 9792, 544);
    run<32>(d, "conv pitch 40px, tap (1,1)", 656, 11520, 640);
    run<64>(d, "canonical dense", 0, 2048, 128);
    run<64>(d, "conv MT=1 pitch 10px, tap (1,1)", 176, 2880, 160);
    run<64>(d, "conv pitch 16px, tap (1,1)", 272, 4608, 256);
    run<128>(d, "canonical dense", 0, 2048, 128);
    run<128>(d, "conv MT=1 pitch 10px, tap (1,1)", 176, 2880, 160);
    run<128>(d, "conv pitch 16px, tap (1,1)", 272, 4608, 256);
    // MN-major A (wgrad dY^T: LBO=128 K-groups, SBO = chunk stride) for reference
    run<16>(d, "wgrad A MN-major dense", 0, 128, 2048, 1);
    return 0;
}
