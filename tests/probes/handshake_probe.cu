// Probe (GPU): cycles per pipeline stage of the bare mbarrier handshake skeleton of tc_conv_kernel (no TMA, no MMA,
// no TMEM): producer -> full -> "MMA" warp -> (empty, tfull) -> 4 epilogue warps -> tempty -> "MMA" warp.
// Variants: V=0 every lane of a role polls the barrier (as the kernel does); V=1 lane 0 polls, then __syncwarp.
#include <cstdio>
#include "tc_ptx.cuh"
using namespace hpfg;

template <int V, int STAGES, int NACC>
__global__ void __launch_bounds__(512, 1) probe(long long *out, int n_work, int epi_work) {
    __shared__ uint64_t bars[3 * 16 + 16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_full = ptx::smem_u32(bars), bar_empty = bar_full + 8 * STAGES, bar_tfull = bar_empty + 8 * STAGES, bar_tempty = bar_tfull + 8 * NACC;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(bar_full + 8 * s, 1); ptx::mbar_init(bar_empty + 8 * s, 1); }
        for (int a = 0; a < NACC; ++a) { ptx::mbar_init(bar_tfull + 8 * a, 1); ptx::mbar_init(bar_tempty + 8 * a, 4); }
        ptx::fence_barrier_init();
    }
    __syncthreads();
    auto wait = [&](uint32_t bar, uint32_t parity) {
        if (V == 0) ptx::mbar_wait(bar, parity, 1);
        else { if (lane == 0) ptx::mbar_wait(bar, parity, 1); __syncwarp(); }
    };
    const long long t0 = clock64();
    if (warp == 0) {
        int stage = 0, phase = 0;
        for (int it = 0; it < n_work; ++it) {
            wait(bar_empty + 8 * stage, phase ^ 1);
            if (ptx::elect_one()) ptx::mbar_arrive(bar_full + 8 * stage);
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        int stage = 0, phase = 0, acc = 0, aphase = 0;
        for (int it = 0; it < n_work; ++it) {
            wait(bar_tempty + 8 * acc, aphase ^ 1);
            wait(bar_full + 8 * stage, phase);
            if (ptx::elect_one()) { ptx::mbar_arrive(bar_empty + 8 * stage); ptx::mbar_arrive(bar_tfull + 8 * acc); }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            if (++acc == NACC) { acc = 0; aphase ^= 1; }
        }
    } else if (warp >= 12) {
        int acc = 0, aphase = 0;
        float x = (float)lane;
        for (int it = 0; it < n_work; ++it) {
            wait(bar_tfull + 8 * acc, aphase);
            for (int k = 0; k < epi_work; ++k) x = x * 1.0001f + 0.5f;      // dependent chain = epilogue latency stand-in (4 cyc each)
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar_tempty + 8 * acc);
            if (++acc == NACC) { acc = 0; aphase ^= 1; }
        }
        if (x == 12345.f) out[1] = 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) out[0] = clock64() - t0;
}

template <int V, int STAGES, int NACC>
static void run(long long *d, int epi) {
    long long h = 0;
    const int n = 2000;
    for (int i = 0; i < 2; ++i) probe<V, STAGES, NACC><<<148, 512>>>(d, n, epi);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("V=%d stages=%2d nacc=%d epilogue chain %4d cycles: %7.1f cycles / stage (%s)\n", V, STAGES, NACC, epi * 4, (double)h / n, cudaGetErrorString(e));
}

int main() {
    long long *d;
    cudaMalloc(&d, 16);
    run<0, 10, 2>(d, 0); run<1, 10, 2>(d, 0);
    run<0, 10, 2>(d, 100); run<1, 10, 2>(d, 100);
    run<0, 10, 2>(d, 300); run<1, 10, 2>(d, 300);
    run<0, 10, 4>(d, 100); run<1, 10, 4>(d, 100);
    run<0, 10, 4>(d, 300); run<1, 10, 4>(d, 300);
    run<1, 10, 8>(d, 300);
    return 0;
}
