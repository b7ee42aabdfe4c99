// Helpers shared by the tensor-core kernels (conv_tc.cu: fprop/dgrad, wgrad_tc.cu: wgrad).
#pragma once
#include <string>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace hpfg {

constexpr int kTH = 16, kTW = 8;            // CTA pixel tile (rows x cols) = 128 pixels = one UMMA M (fprop) / 8 K-steps (wgrad)

__device__ __forceinline__ void unpack8(const uint4 &v, float (&f)[8]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&h);
}

// Walks the m-tiles this CTA owns (mt0, mt0+step, ...) without per-tile integer divisions.
struct TileIter {
    int n_img, th, tw, dn, dh, dw;
    __device__ __forceinline__ void init(int mt, int step, int tiles_h, int tiles_w) {
        const int tpi = tiles_h * tiles_w;
        n_img = mt / tpi;
        int r = mt % tpi;
        th = r / tiles_w;
        tw = r % tiles_w;
        dn = step / tpi;
        r = step % tpi;
        dh = r / tiles_w;
        dw = r % tiles_w;
    }
    __device__ __forceinline__ void next(int tiles_h, int tiles_w) {
        tw += dw;
        if (tw >= tiles_w) { tw -= tiles_w; ++th; }
        th += dh;
        if (th >= tiles_h) { th -= tiles_h; ++n_img; }
        n_img += dn;
    }
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 4-D tiled map over a bf16 NHWC tensor: dims (C, W, H, N), box (kc, box_w, box_h, 1), zero fill out of bounds
inline int make_map(CUtensorMap *m, const void *base, int N, int H, int W, int Cc, int kc, int box_w, int box_h) {
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return HPFG_ERR_CUDA;
    }
    cuuint64_t dims[4] = {(cuuint64_t)Cc, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)Cc * 2, (cuuint64_t)W * Cc * 2, (cuuint64_t)H * W * Cc * 2};
    cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
        return HPFG_ERR_CUDA;
    }
    return HPFG_OK;
}


// 5-D tiled map that makes TMA itself produce the UMMA operand layout [8-channel chunk][h][w][8 ch]:
// dims (8 ch, W, H, C/8 chunks, N) with the chunk dimension strided by 16 bytes, box (8, box_w, box_h, chunks, 1).
inline int make_map_chunked(CUtensorMap *m, const void *base, int N, int H, int W, int Cc, int chunks, int box_w, int box_h) {
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return HPFG_ERR_CUDA;
    }
    cuuint64_t dims[5] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(Cc / 8), (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)Cc * 2, (cuuint64_t)W * Cc * 2, 16, (cuuint64_t)H * W * Cc * 2};
    cuuint32_t box[5] = {8, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)chunks, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (chunked 5-D) failed with code " + std::to_string((int)r));
        return HPFG_ERR_CUDA;
    }
    return HPFG_OK;
}

}  // namespace hpfg
