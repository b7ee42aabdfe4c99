// Probe (GPU): TMEM lane layout of the accumulator for tcgen05.mma cta_group::1 with M = 64.
// A[m][0] = m + 1 (other k zero), B[n][0] = n + 1  =>  D[m][n] = (m+1)(n+1).  Dump 128 lanes x 16 columns.
#include <cstdio>
#include <cuda_bf16.h>
#include "tc_ptx.cuh"
using namespace hpfg;

__global__ void probe(float *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *A = smem, *B = smem + 8192;            // K-major, SWIZZLE_NONE: [k8][row][8]: LBO = rows*16, SBO = 128
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    for (int i = threadIdx.x; i < 16384 / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0;
    __syncthreads();
    if (threadIdx.x < 128) *(__nv_bfloat16 *)(A + threadIdx.x * 16) = __float2bfloat16((float)(threadIdx.x + 1));   // k = 0
    if (threadIdx.x < 16) *(__nv_bfloat16 *)(B + threadIdx.x * 16) = __float2bfloat16((float)(threadIdx.x + 1));
    ptx::fence_proxy_async_smem();
    if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_barrier_init(); }
    if (threadIdx.x < 32) ptx::tmem_alloc(ptx::smem_u32(&tslot), 32);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tslot;
    if (threadIdx.x == 0) {
        // clear all 128 lanes with an M=128 MMA on a zero A (rows 128.. of smem region beyond are zero: use B region offset)
        const uint64_t az = ptx::umma_desc(ptx::smem_u32(smem) + 12288, 128 * 16, 128);   // zero area
        const uint64_t bd = ptx::umma_desc(ptx::smem_u32(B), 16 * 16, 128);
        ptx::umma_bf16(tmem, az, bd, ptx::umma_idesc_bf16(128, 16, 0, 0), 0);
        const uint64_t ad = ptx::umma_desc(ptx::smem_u32(A), 128 * 16, 128);              // rows 0..63 used by M=64
        ptx::umma_bf16(tmem, ad, bd, ptx::umma_idesc_bf16(64, 16, 0, 0), 0);
        ptx::umma_commit(ptx::smem_u32(&bar));
    }
    if (threadIdx.x < 128) {
        ptx::mbar_wait(ptx::smem_u32(&bar), 0, 99);
        ptx::tc_fence_after();
        uint32_t r[16];
        ptx::tmem_ld16(tmem + ((uint32_t)((threadIdx.x / 32) * 32) << 16), r);
        ptx::tmem_ld_wait();
        for (int j = 0; j < 16; ++j) out[threadIdx.x * 16 + j] = __uint_as_float(r[j]);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) ptx::tmem_dealloc(tmem, 32);
}

int main() {
    float *d, h[128 * 16];
    cudaMalloc(&d, sizeof(h));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
    probe<<<1, 128, 32 * 1024>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("lane: D[lane][0] (row = value-1), D[lane][1]/2\n");
    for (int l = 0; l < 128; ++l) printf("%d:%g,%g%s", l, h[l * 16], h[l * 16 + 1] / 2, (l % 8 == 7) ? "\n" : "  ");
    return 0;
}
