// cluster_launch.cuh -- host launchers for the cluster-resident four-step kernels (cluster.cuh).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "cluster.cuh"

namespace ssfft {

template <typename Cfg, int KIND>
int cluster_prepare() {
    auto kern = cluster_fft_kernel<Cfg, KIND>;
    if (Cfg::smem_bytes > 48 * 1024 &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes) != cudaSuccess)
        return -1;
    if (Cfg::C > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)
        return -1;
    return 0;
}

template <typename Cfg, int KIND>
int cluster_max_clusters() {
    if (cluster_prepare<Cfg, KIND>() != 0) { cudaGetLastError(); return -1; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(Cfg::C * 1024);
    cfg.blockDim = dim3(Cfg::THREADS);
    cfg.dynamicSmemBytes = Cfg::smem_bytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = Cfg::C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, cluster_fft_kernel<Cfg, KIND>, &cfg) != cudaSuccess) { cudaGetLastError(); return -1; }
    return n;
}

// Tensor map of the input seen as (batch, N1, N2) elements of 8 bytes (16 for fp64), box = one CTA's [N1][CT1] tile.
// cuTensorMapEncodeTiled is a driver entry point: fetched through the runtime so libcuda is not a link dependency.
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
template <typename Cfg>
bool make_input_tensor_map(CUtensorMap *tm, const void *in, long long batch) {
    using T = typename Cfg::T;
    if (sizeof(cx<T>) != 8) return false;  // 8-byte elements only (fp32 complex)
    ssfft_encode_tiled_fn enc = tensor_map_encoder();
    if (!enc || (reinterpret_cast<uintptr_t>(in) & 15u) || batch <= 0 || batch > 0x7fffffffLL) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)Cfg::N2, (cuuint64_t)Cfg::N1, (cuuint64_t)batch};
    const cuuint64_t strides[2] = {(cuuint64_t)Cfg::N2 * sizeof(cx<T>), (cuuint64_t)Cfg::N * sizeof(cx<T>)};
    const cuuint32_t box[3] = {(cuuint32_t)Cfg::CT1, (cuuint32_t)Cfg::N1, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void *>(in), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// one persistent launch: min(batch, co-resident clusters) clusters loop over the transforms
template <typename Cfg, int KIND>
int launch_cluster(const void *params, int max_clusters, cudaStream_t s) {
    using T = typename Cfg::T;
    ClusterParams<T> q = *reinterpret_cast<const ClusterParams<T> *>(params);
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    q.use_tma = (KIND != CL_C2R) && q.use_tma && make_input_tensor_map<Cfg>(&tmap, q.in, q.batch);
    long long clusters = q.batch < max_clusters ? q.batch : max_clusters;
    if (clusters <= 0) return 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * Cfg::C));
    cfg.blockDim = dim3(Cfg::THREADS);
    cfg.dynamicSmemBytes = Cfg::smem_bytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = Cfg::C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, cluster_fft_kernel<Cfg, KIND>, q, tmap) == cudaSuccess ? 0 : 2;
}

template <typename Cfg>
ClusterEntry make_cluster_entry(const char *name, unsigned kinds_default, unsigned kinds_all) {
    const char *all = getenv("SSFFT_DSMEM_ALL");
    const unsigned kinds = (all && all[0] == '1') ? kinds_all : kinds_default;
    ClusterEntry e;
    e.prec = sizeof(typename Cfg::T) == 4 ? 0 : 1;
    e.n1 = Cfg::N1; e.n2 = Cfg::N2; e.csize = Cfg::C; e.ra0 = Cfg::RA0; e.rb0 = Cfg::RB0;
    e.name = name;
    e.smem_bytes = Cfg::smem_bytes;
    e.kinds = kinds;
    e.launch[0] = &launch_cluster<Cfg, CL_C2C>;
    e.launch[1] = &launch_cluster<Cfg, CL_R2C>;
    e.launch[2] = &launch_cluster<Cfg, CL_C2R>;
    e.max_clusters[0] = &cluster_max_clusters<Cfg, CL_C2C>;
    e.max_clusters[1] = &cluster_max_clusters<Cfg, CL_R2C>;
    e.max_clusters[2] = &cluster_max_clusters<Cfg, CL_C2R>;
    return e;
}

}  // namespace ssfft
