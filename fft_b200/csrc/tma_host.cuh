// tma_host.cuh -- host side of the TMA tensor copies: encode a 3-D tensor map (batch, rows, cols) of 8-byte elements.
// cuTensorMapEncodeTiled is a driver entry point; it is fetched through the runtime so libcuda is not a link dependency.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace ssfft {

typedef CUresult (*ssfft_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                          const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                          CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline ssfft_encode_tiled_fn tensor_map_encoder() {
    static ssfft_encode_tiled_fn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            p = nullptr;
        }
        return (ssfft_encode_tiled_fn)p;
    }();
    return fn;
}

// `base` viewed as [batch][rows][cols] elements of 8 bytes (one fp32 complex), copied in boxes of [1][box_rows][box_cols]
inline bool encode_tensor_map_3d(CUtensorMap *tm, const void *base, long long batch, int rows, int cols, int box_rows,
                                 int box_cols) {
    ssfft_encode_tiled_fn enc = tensor_map_encoder();
    if (!enc || (reinterpret_cast<uintptr_t>(base) & 15u) || batch <= 0 || batch > 0x7fffffffLL) return false;
    if (box_rows > 256 || box_cols > 256 || (box_cols * 8) % 16) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
    const cuuint64_t strides[2] = {(cuuint64_t)cols * 8, (cuuint64_t)cols * rows * 8};
    const cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void *>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace ssfft
