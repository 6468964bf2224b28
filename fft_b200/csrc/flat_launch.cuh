// flat_launch.cuh -- host launchers for the ticket-queue four-step kernels (flat.cuh).
#pragma once
#include <cstring>
#include <mutex>

#include "flat.cuh"
#include "tma_host.cuh"

namespace ssfft {

template <typename CfgA, typename CfgB, int INV, int NSTAGE, int MINB, bool INPLACE, int KIND = 0>
int flat_max_ctas() {
    using Lay = FlatLayout<CfgA, CfgB, NSTAGE, INPLACE, KIND>;
    auto kern = fourstep_flat_kernel<CfgA, CfgB, INV, NSTAGE, MINB, INPLACE, KIND>;
    static int cached[32] = {0};
    static std::mutex m;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 32) return -1;
    std::lock_guard<std::mutex> lock(m);
    if (cached[dev]) return cached[dev];
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lay::smem_bytes) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    int per_sm = 0, sms = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, CfgA::THREADS + kFlatHelpers, Lay::smem_bytes) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cached[dev] = per_sm * sms > 0 ? per_sm * sms : -1;
    return cached[dev];
}

// returns 0 on success, 2 on a launch error, 3 when the input cannot be described by a tensor map (caller falls back)
template <typename CfgA, typename CfgB, int INV, int NSTAGE, int MINB, bool INPLACE, int KIND = 0>
int launch_flat(const void *params, int ctas, cudaStream_t s) {
    using T = typename CfgA::T;
    using Lay = FlatLayout<CfgA, CfgB, NSTAGE, INPLACE, KIND>;
    const FlatParams<T> &q = *reinterpret_cast<const FlatParams<T> *>(params);
    if (q.batch <= 0) return 0;
    if (flat_max_ctas<CfgA, CfgB, INV, NSTAGE, MINB, INPLACE, KIND>() < 1) return 2;  // also sets the shared-memory attribute
    CUtensorMap tmap, tmap2;
    memset(&tmap, 0, sizeof(tmap));
    memset(&tmap2, 0, sizeof(tmap2));
    constexpr int kBoxRows = flat_box_rows(CfgA::L);
    constexpr int kBoxCols = KIND == 2 ? CfgA::CT / 2 : CfgA::CT;
    if (!encode_tensor_map_3d(&tmap, q.in, q.batch, CfgA::L, CfgB::L, kBoxRows, kBoxCols, (int)sizeof(cx<T>))) return 3;
    if (KIND == 2 && !encode_tensor_map_3d(&tmap2, q.in, q.batch, CfgA::L, CfgB::L, kBoxRows, CfgA::CT / 2 + 2, (int)sizeof(cx<T>))) return 3;
    fourstep_flat_kernel<CfgA, CfgB, INV, NSTAGE, MINB, INPLACE, KIND><<<(unsigned)ctas, CfgA::THREADS + kFlatHelpers, Lay::smem_bytes, s>>>(q, tmap, tmap2);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

// REAL: which RealFFT kernels of length 2 N1 N2 the entry carries -- bit 0 forward, bit 1 inverse (default: both for the
// entries with a separate exchange buffer).  The first entry of a size that carries a direction is that direction's default.
template <typename CfgA, typename CfgB, int NSTAGE, int MINB, bool INPLACE = true, int REAL = INPLACE ? 0 : 3>
FlatEntry make_flat_entry(const char *name) {
    using Lay = FlatLayout<CfgA, CfgB, NSTAGE, INPLACE>;
    FlatEntry e;
    e.prec = sizeof(typename CfgA::T) == 4 ? 0 : 1;
    e.n1 = CfgA::L; e.n2 = CfgB::L; e.name = name;
    for (int i = 0; i < 3; ++i) e.ra[i] = CfgA::radix(i);
    e.na_passes = CfgA::NP; e.cta = CfgA::CT; e.ctb = CfgB::CT;
    e.threads = CfgA::THREADS + kFlatHelpers; e.nstage = NSTAGE; e.minb = MINB; e.inplace = INPLACE ? 1 : 0;
    e.smem_bytes = Lay::smem_bytes;
    e.tile_b_tw = CfgB::tw_total;
    e.nb_passes = CfgB::NP;
    for (int i = 0; i < 3; ++i) e.rb[i] = CfgB::radix(i);
    e.launch[0] = &launch_flat<CfgA, CfgB, 0, NSTAGE, MINB, INPLACE>;
    e.launch[1] = &launch_flat<CfgA, CfgB, 1, NSTAGE, MINB, INPLACE>;
    e.max_ctas[0] = &flat_max_ctas<CfgA, CfgB, 0, NSTAGE, MINB, INPLACE>;
    e.max_ctas[1] = &flat_max_ctas<CfgA, CfgB, 1, NSTAGE, MINB, INPLACE>;
    e.launch_real[0] = e.launch_real[1] = nullptr;
    e.max_ctas_real[0] = e.max_ctas_real[1] = nullptr;
    constexpr bool real_ok = CfgA::CT % 2 == 0 && CfgB::CT % 2 == 0 && CfgA::prod(CfgA::NP - 1) % (CfgB::CT / 2) == 0;
    static_assert(REAL == 0 || real_ok, "these tiles cannot carry the real flavours");
    if constexpr ((REAL & 1) != 0) {
        e.launch_real[0] = &launch_flat<CfgA, CfgB, 0, NSTAGE, MINB, INPLACE, 1>;
        e.max_ctas_real[0] = &flat_max_ctas<CfgA, CfgB, 0, NSTAGE, MINB, INPLACE, 1>;
    }
    if constexpr ((REAL & 2) != 0) {
        e.launch_real[1] = &launch_flat<CfgA, CfgB, 1, NSTAGE, MINB, INPLACE, 2>;
        e.max_ctas_real[1] = &flat_max_ctas<CfgA, CfgB, 1, NSTAGE, MINB, INPLACE, 2>;
    }
    return e;
}

}  // namespace ssfft
