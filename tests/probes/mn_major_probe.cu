// Probe (GPU): which of {LBO,SBO} is the MN-group stride and which the K-group stride for MN-major,
// SWIZZLE_NONE tcgen05 operands?  One CTA, one MMA M=128,N=16,K=16 (plus K=32 via two MMAs), exact small ints.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I hpfg_b200/csrc tests/probes/mn_major_probe.cu -o /tmp/mn_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include "tc_ptx.cuh"
using namespace hpfg;

constexpr int M = 128, N = 16, K = 32;
constexpr int A_MG = 2064, A_KG = 128;      // byte strides used to BUILD the A image: M-group (8 rows), K-group (8 k)
constexpr int B_NG = 2896, B_KG = 160;      // same for B (N-group, K-group)

__host__ __device__ inline float aval(int m, int k) { return (float)(((m * 7 + k * 3) % 5) - 2); }
__host__ __device__ inline float bval(int n, int k) { return (float)(((n * 5 + k) % 7) - 3); }

__global__ void probe(float *out, int variant) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *A = smem, *B = smem + 40 * 1024;
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    for (int i = threadIdx.x; i < 60 * 1024 / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0;
    __syncthreads();
    // element (m,k): unit (mu = m/8, kg = k/8, ki = k%8), 8 m's contiguous inside the 16-byte unit
    for (int i = threadIdx.x; i < M * K; i += blockDim.x) {
        const int m = i % M, k = i / M;
        *(__nv_bfloat16 *)(A + (m / 8) * A_MG + (k / 8) * A_KG + (k % 8) * 16 + (m % 8) * 2) = __float2bfloat16(aval(m, k));
    }
    for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
        const int n = i % N, k = i / N;
        *(__nv_bfloat16 *)(B + (n / 8) * B_NG + (k / 8) * B_KG + (k % 8) * 16 + (n % 8) * 2) = __float2bfloat16(bval(n, k));
    }
    ptx::fence_proxy_async_smem();
    if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_barrier_init(); }
    if (threadIdx.x < 32) ptx::tmem_alloc(ptx::smem_u32(&tslot), 32);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tslot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = ptx::umma_idesc_bf16(M, N, 1, 1);
        for (int ks = 0; ks < K / 16; ++ks) {
            // K step of 16 = two K-groups; next K step starts 2 K-groups further
            const uint32_t a0 = ptx::smem_u32(A) + ks * 2 * A_KG, b0 = ptx::smem_u32(B) + ks * 2 * B_KG;
            uint64_t ad, bd;
            if (variant & 1) ad = ptx::umma_desc(a0, A_MG, A_KG); else ad = ptx::umma_desc(a0, A_KG, A_MG);   // (lbo, sbo)
            if (variant & 2) bd = ptx::umma_desc(b0, B_NG, B_KG); else bd = ptx::umma_desc(b0, B_KG, B_NG);
            ptx::umma_bf16(tmem, ad, bd, idesc, ks > 0);
        }
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
    float *d, h[M * N];
    cudaMalloc(&d, sizeof(h));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int v = 0; v < 4; ++v) {
        probe<<<1, 128, 64 * 1024>>>(d, v);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", v, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        int ok = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                float ref = 0;
                for (int k = 0; k < K; ++k) ref += aval(m, k) * bval(n, k);
                ok += (h[m * N + n] == ref);
            }
        printf("variant %d (A: %s, B: %s): %d / %d exact\n", v, (v & 1) ? "lbo=MNgroup,sbo=Kgroup" : "lbo=Kgroup,sbo=MNgroup",
               (v & 2) ? "lbo=MNgroup,sbo=Kgroup" : "lbo=Kgroup,sbo=MNgroup", ok, M * N);
        if (v == 0) { printf("  D[0][0..7] ="); for (int n = 0; n < 8; ++n) printf(" %g", h[n]); printf("\n"); }
    }
    return 0;
}
