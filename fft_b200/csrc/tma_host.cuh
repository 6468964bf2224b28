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

// `base` viewed as [batch][rows][cols] complex elements of elem_bytes (8: fp32, 16: fp64), copied in boxes of
// [1][box_rows][box_cols].  The map itself counts 8-byte words (a fp64 element is two): the kernel scales its column
// coordinate the same way.
inline bool encode_tensor_map_3d(CUtensorMap *tm, const void *base, long long batch, int rows, int cols, int box_rows,
                                 int box_cols, int elem_bytes = 8) {
    ssfft_encode_tiled_fn enc = tensor_map_encoder();
    const int w = elem_bytes / 8;  // words per element
    if (!enc || (reinterpret_cast<uintptr_t>(base) & 15u) || batch <= 0 || batch > 0x7fffffffLL || (w != 1 && w != 2)) return false;
    if (box_rows > 256 || box_cols * w > 256 || (box_cols * elem_bytes) % 16) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)cols * w, (cuuint64_t)rows, (cuuint64_t)batch};
    const cuuint64_t strides[2] = {(cuuint64_t)cols * elem_bytes, (cuuint64_t)cols * rows * elem_bytes};
    const cuuint32_t box[3] = {(cuuint32_t)(box_cols * w), (cuuint32_t)box_rows, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void *>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace ssfft
