// Probe (GPU): tcgen05.mma rate when consecutive MMAs read DIFFERENT A tiles (the conv's 9 tap-shifted windows x MT tiles)
// versus the same A tile every time.  M=128, K=16, bf16, SWIZZLE_NONE K-major, conv MT=4 geometry (pitch 34 px, chunk
// stride 9792 B).  MODE 0: same descriptor; 1: 9 taps of one tile; 2: 9 taps x 4 tiles (kernel order); 3: as 2 but B also varies per tap.
#include <cstdio>
#include <cuda_bf16.h>
#include "tc_ptx.cuh"
using namespace hpfg;

template <int N, int MODE>
__global__ void probe(long long *out, int acc0, int tcols, int commit_every) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar2[2];
    uint64_t &bar = bar2[0];
    __shared__ uint32_t tslot;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0x3c003c00u + i;
    if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::mbar_init(ptx::smem_u32(&bar) + 8, 1); ptx::fence_barrier_init(); }
    if (threadIdx.x < 32) ptx::tmem_alloc(ptx::smem_u32(&tslot), tcols);
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tslot;
    if (threadIdx.x < 32) {
        constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, N, 0, 0);
        const uint32_t a = ptx::smem_u32(smem), b = ptx::smem_u32(smem) + 150 * 1024;
        long long t0 = clock64();
        if (ptx::elect_one()) {
#pragma unroll 1
            for (int rep = 0; rep < 16; ++rep) {
#pragma unroll 1
                for (int j = 0; j < 4; ++j) {
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        uint32_t ao = 0, bo = 0;
                        if (MODE >= 1) ao = ((tap / 3) * 34 + tap % 3) * 16;
                        if (MODE >= 2) ao += j * 128 + (rep % 4) * 19584;
                        if (MODE >= 3) bo = tap * 16 * N * 2;
                        ptx::umma_bf16(tmem + j * N, ptx::umma_desc(a + ao, 9792, 544), ptx::umma_desc(b + bo, N * 16, 128), idesc, (acc0 && tap == 0) ? 0u : 1u);
                    }
                }
                if (commit_every) ptx::umma_commit(ptx::smem_u32(&bar) + 8);
            }
            ptx::umma_commit(ptx::smem_u32(&bar));
        }
        __syncwarp();
        ptx::mbar_wait(ptx::smem_u32(&bar), 0, 99);
        long long t1 = clock64();
        if (threadIdx.x == 0) *out = t1 - t0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) ptx::tmem_dealloc(tmem, tcols);
}

template <int N, int MODE>
static void run(long long *d, int grid, int threads = 128, int acc0 = 0, int tcols = 512, int commit_every = 0) {
    cudaFuncSetAttribute(probe<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    long long h = 0;
    for (int i = 0; i < 2; ++i) probe<N, MODE><<<grid, threads, 200 * 1024>>>(d, acc0, tcols, commit_every);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("N=%3d mode %d grid %3d threads %d acc0 %d tcols %d commit %d: %6.1f cycles / MMA (%s)\n", N, MODE, grid, threads, acc0, tcols, commit_every, (double)h / 576, cudaGetErrorString(e));
}

int main() {
    long long *d;
    cudaMalloc(&d, 8);
    run<16, 3>(d, 148, 512); run<16, 3>(d, 148, 128, 1); run<16, 3>(d, 148, 128, 0, 256); run<16, 3>(d, 148, 128, 0, 512, 1); run<16, 3>(d, 148, 512, 1, 256, 1);
    run<16, 0>(d, 1); run<16, 1>(d, 1); run<16, 2>(d, 1); run<16, 3>(d, 1); run<16, 3>(d, 148);
    run<32, 0>(d, 1); run<32, 3>(d, 1);
    run<64, 0>(d, 1); run<64, 3>(d, 1);
    return 0;
}
