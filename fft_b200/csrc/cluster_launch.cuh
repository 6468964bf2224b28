// cluster_launch.cuh -- host launchers for the cluster-resident four-step kernels (cluster.cuh).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "cluster.cuh"
#include "tma_host.cuh"

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

// Tensor map of the input seen as (batch, N1, N2) elements of 8 bytes, box = one CTA's [N1][CT1] tile (tma_host.cuh).
template <typename Cfg>
bool make_input_tensor_map(CUtensorMap *tm, const void *in, long long batch) {
    if (sizeof(cx<typename Cfg::T>) != 8) return false;  // 8-byte elements only (fp32 complex)
    return encode_tensor_map_3d(tm, in, batch, Cfg::N1, Cfg::N2, Cfg::N1, Cfg::CT1);
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
