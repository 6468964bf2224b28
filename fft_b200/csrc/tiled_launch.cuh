// tiled_launch.cuh -- host launchers for the four-step tile kernels (tiled.cuh).
#pragma once
#include <cstdlib>

#include <cstring>
#include <mutex>

#include "tiled.cuh"
#include "tma_host.cuh"

namespace ssfft {

template <typename Cfg, int FLAVOR>
int launch_tile(const void *params, cudaStream_t s) {
    using T = typename Cfg::T;
    const TileParams<T> &p = *reinterpret_cast<const TileParams<T> *>(params);
    static int ready_mask = 0;
    static int resident[32] = {0};
    static std::mutex setup_mutex;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 2;
    if (dev < 0 || dev >= 32) return 1;
    std::unique_lock<std::mutex> lock(setup_mutex);
    if (!(ready_mask & (1 << dev))) {
        if (Cfg::smem_bytes > 48 * 1024 &&
            cudaFuncSetAttribute(tile_fft_kernel<Cfg, FLAVOR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)Cfg::smem_bytes) != cudaSuccess)
            return 2;
        int per_sm = 0, sms = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tile_fft_kernel<Cfg, FLAVOR>, Cfg::THREADS,
                                                          Cfg::smem_bytes) != cudaSuccess)
            return 2;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        resident[dev] = (per_sm < 1 ? 1 : per_sm) * sms;
        ready_mask |= 1 << dev;
    }
    const int resident_dev = resident[dev];
    lock.unlock();
    const int width = (FLAVOR == TILE_A_C2C)   ? p.n2
                      : (FLAVOR == TILE_B_C2C) ? p.n1
                      : (FLAVOR == TILE_A_R2C || FLAVOR == TILE_A_C2R) ? p.n2 / 2
                                               : p.n1 / 2 + 1;
    const long long items = p.batch * ((width + Cfg::CT - 1) / Cfg::CT);
    if (items <= 0) return 0;
    long long grid = items;
    const long long cap = (long long)resident_dev * fused_waves();
    if (cap > 0 && grid > cap) grid = cap;
    tile_fft_kernel<Cfg, FLAVOR><<<(unsigned)grid, dim3(Cfg::CT, Cfg::TX), Cfg::smem_bytes, s>>>(p);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

template <typename Cfg>
TileEntry make_tile_entry(const char *name) {
    TileEntry e;
    e.prec = sizeof(typename Cfg::T) == 4 ? 0 : 1;
    e.len = Cfg::L;
    e.name = name;
    e.tw_total = Cfg::tw_total;
    e.np = Cfg::NP;
    e.ct = Cfg::CT;
    for (int i = 0; i < 3; ++i) e.radix[i] = Cfg::radix(i);
    e.launch[TILE_A_C2C] = &launch_tile<Cfg, TILE_A_C2C>;
    e.launch[TILE_B_C2C] = &launch_tile<Cfg, TILE_B_C2C>;
    e.launch[TILE_A_R2C] = &launch_tile<Cfg, TILE_A_R2C>;
    e.launch[TILE_B_R2C] = &launch_tile<Cfg, TILE_B_R2C>;
    e.launch[TILE_B_C2R] = &launch_tile<Cfg, TILE_B_C2R>;
    e.launch[TILE_A_C2R] = &launch_tile<Cfg, TILE_A_C2R>;
    return e;
}

template <typename CfgA, typename CfgB, int KIND>
int fourstep_max_clusters(int cluster_size) {
    constexpr size_t smem = fourstep_smem_bytes<CfgA, CfgB>();
    auto kern = fourstep_cluster_kernel<CfgA, CfgB, KIND>;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return -1;
    if (cluster_size > 8 &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)
        return -1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cluster_size * 1024);
    cfg.blockDim = dim3(CfgA::THREADS);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_size; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return -1; }
    return n;
}

template <typename CfgA, typename CfgB, int KIND>
int launch_fourstep(const void *params, int max_clusters, cudaStream_t s) {
    using T = typename CfgA::T;
    const FourStepParams<T> &q = *reinterpret_cast<const FourStepParams<T> *>(params);
    constexpr size_t smem = fourstep_smem_bytes<CfgA, CfgB>();
    const int csize = q.cluster_size > 0 ? q.cluster_size : fourstep_cluster_size();
    // max_clusters = groups * group_clusters (plan time); never start more groups than there are transforms
    const int G = q.group_clusters > 1 ? q.group_clusters : 1;
    long long groups = max_clusters / G;
    if (q.batch < groups) groups = q.batch;
    long long clusters = groups * G;
    if (clusters <= 0) return 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * csize));
    cfg.blockDim = dim3(CfgA::THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[3];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (G > 1) {  // the software barrier needs every CTA of the grid resident at once
        attr[1].id = cudaLaunchAttributeCooperative;
        attr[1].val.cooperative = 1;
        cfg.numAttrs = 2;
    }
    // Optional (SSFFT_L2_PERSIST=1): pin the scratch in L2 with a persisting access-policy window.  Measured
    // SLOWER on B200 (65536: 48 % -> 28 % of roofline), so it is off by default; consumed scratch lines are
    // dropped with discard.global.L2 inside the kernel instead.
    static int persist_state = 0;  // 0 = unknown, 1 = enabled, -1 = unavailable / disabled
    static size_t max_window = 0;
    static std::mutex persist_mutex;
    std::unique_lock<std::mutex> plock(persist_mutex);
    if (persist_state == 0) {
        const char *e = getenv("SSFFT_L2_PERSIST");
        int dev = 0, max_persist = 0, max_win = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
        cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, dev);
        if (!(e && e[0] == '1') || max_persist <= 0 || max_win <= 0 ||
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist) != cudaSuccess) {
            cudaGetLastError();
            persist_state = -1;
        } else {
            persist_state = 1;
            max_window = (size_t)max_win;
        }
    }
    plock.unlock();
    if (persist_state == 1 && G == 1) {
        size_t bytes = (size_t)2 * (size_t)clusters * (size_t)q.scratch_per * sizeof(cx<T>);
        if (bytes > max_window) bytes = max_window;
        attr[1].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[1].val.accessPolicyWindow.base_ptr = (void *)q.scratch;
        attr[1].val.accessPolicyWindow.num_bytes = bytes;
        attr[1].val.accessPolicyWindow.hitRatio = 1.0f;
        attr[1].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[1].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cfg.numAttrs = 2;
    }
    // staged variants: stage-1 tiles come through a tensor map of the user input, [batch][N1][W] complex
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    FourStepParams<T> qq = q;
    qq.use_tma = 0;
    if (SSFFT_FOURSTEP_TMA && fourstep_staged<CfgA, CfgB>() && KIND != 2) {
        const int W = KIND == 0 ? CfgB::L : CfgB::L / 2;
        if (!encode_tensor_map_3d(&tmap, q.in, q.batch, CfgA::L, W, CfgA::L, CfgA::CT)) return 3;  // caller falls back
        qq.use_tma = 1;
    }
    return cudaLaunchKernelEx(&cfg, fourstep_cluster_kernel<CfgA, CfgB, KIND>, qq, tmap) == cudaSuccess ? 0 : 2;
}

template <typename CfgA, typename CfgB>
FourStepEntry make_fourstep_entry(const char *name) {
    FourStepEntry e;
    e.prec = sizeof(typename CfgA::T) == 4 ? 0 : 1;
    e.n1 = CfgA::L; e.n2 = CfgB::L; e.name = name;
    e.threads = CfgA::THREADS;
    e.smem_bytes = fourstep_smem_bytes<CfgA, CfgB>();
    e.launch[0] = &launch_fourstep<CfgA, CfgB, 0>;
    e.launch[1] = &launch_fourstep<CfgA, CfgB, 1>;
    e.launch[2] = &launch_fourstep<CfgA, CfgB, 2>;
    e.max_clusters[0] = &fourstep_max_clusters<CfgA, CfgB, 0>;
    e.max_clusters[1] = &fourstep_max_clusters<CfgA, CfgB, 1>;
    e.max_clusters[2] = &fourstep_max_clusters<CfgA, CfgB, 2>;
    // tiles per transform of each stage (same formulas as the kernel)
    e.tiles[0][0] = (tile_width<TILE_A_C2C>(CfgA::L, CfgB::L) + CfgA::CT - 1) / CfgA::CT;
    e.tiles[0][1] = (tile_width<TILE_B_C2C>(CfgA::L, CfgB::L) + CfgB::CT - 1) / CfgB::CT;
    e.tiles[1][0] = (tile_width<TILE_A_R2C>(CfgA::L, CfgB::L) + CfgA::CT - 1) / CfgA::CT;
    e.tiles[1][1] = (tile_width<TILE_B_R2C>(CfgA::L, CfgB::L) + CfgB::CT - 1) / CfgB::CT;
    e.tiles[2][0] = (tile_width<TILE_B_C2R>(CfgA::L, CfgB::L) + CfgB::CT - 1) / CfgB::CT;
    e.tiles[2][1] = (tile_width<TILE_A_C2R>(CfgA::L, CfgB::L) + CfgA::CT - 1) / CfgA::CT;
    return e;
}

#define SSFFT_TILE(T, L, R0, R1, R2, TX, CT, MINB) \
    make_tile_entry<TileCfg<T, L, R0, R1, R2, TX, CT, MINB>>(#T "_tile" #L "_" #R0 "x" #R1 "x" #R2 "_ct" #CT)

}  // namespace ssfft
