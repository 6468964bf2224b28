// fused_launch.cuh -- host launcher shared by the fused-kernel translation units.
#pragma once
#include <mutex>

#include "fused.cuh"

namespace ssfft {

// grid = min(#transform groups, SMs * resident CTAs per SM * waves); CTAs loop over groups.
template <typename Cfg>
int launch_cfg(const void *tw, const void *in, void *out, long long batch, int inverse, int mode, const void *rtw,
               cudaStream_t s) {
    using T = typename Cfg::T;
    static int ready_mask = 0;      // per-device one-time setup
    static int resident[64] = {0};  // SMs * CTAs/SM per device
    static std::mutex setup_mutex;  // plans may be executed from several host threads
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 2;
    if (dev < 0 || dev >= 32) return 1;
    std::unique_lock<std::mutex> lock(setup_mutex);
    if (!(ready_mask & (1 << dev))) {
        if (Cfg::smem_bytes > 48 * 1024 &&
            (cudaFuncSetAttribute(fused_fft_kernel<Cfg, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)Cfg::smem_bytes) != cudaSuccess ||
             cudaFuncSetAttribute(fused_fft_kernel<Cfg, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)Cfg::smem_bytes) != cudaSuccess))
            return 2;
        int per_sm = 0, sms = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused_fft_kernel<Cfg, false>, Cfg::TX * Cfg::FPB,
                                                          Cfg::smem_bytes) != cudaSuccess)
            return 2;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (per_sm < 1) per_sm = 1;
        resident[dev] = per_sm * sms;
        ready_mask |= 1 << dev;
    }
    const int resident_dev = resident[dev];
    lock.unlock();
    const long long groups = (batch + Cfg::FPB - 1) / Cfg::FPB;
    if (groups <= 0) return 0;
    long long grid = groups;
    const long long cap = (long long)resident_dev * fused_waves();
    if (cap > 0 && grid > cap) grid = cap;
    dim3 block(Cfg::TX, Cfg::FPB);
    if (mode >= FUSED_R2C_MOD)
        fused_fft_kernel<Cfg, true><<<(unsigned)grid, block, Cfg::smem_bytes, s>>>(
            (const cx<T> *)in, (cx<T> *)out, (const cx<T> *)tw, (const cx<T> *)rtw, batch, inverse, mode, FusedIo<T>{});
    else
        fused_fft_kernel<Cfg, false><<<(unsigned)grid, block, Cfg::smem_bytes, s>>>(
            (const cx<T> *)in, (cx<T> *)out, (const cx<T> *)tw, (const cx<T> *)rtw, batch, inverse, mode, FusedIo<T>{});
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

// Extended I/O (ssfft_exec_*_ex): the NoStaging twin of the configuration behind strided / overlapping loads and stores
// with fused multipliers.  `io` points at a host FusedIo<T>.
// X: the configuration that runs (NoStaging<Cfg>, or the column configuration ColumnCfg<Cfg>)
template <typename X>
int launch_ex_as(const void *tw, const void *in, void *out, long long batch, int inverse, int mode, const void *rtw,
                 const void *io, cudaStream_t s) {
    using T = typename X::T;
    static int ready_mask = 0;
    static int resident[64] = {0};
    static std::mutex setup_mutex;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 2;
    if (dev < 0 || dev >= 32) return 1;
    std::unique_lock<std::mutex> lock(setup_mutex);
    if (!(ready_mask & (1 << dev))) {
        if (X::smem_bytes > 48 * 1024 &&
            cudaFuncSetAttribute(fused_fft_kernel<X, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)X::smem_bytes) != cudaSuccess)
            return 2;
        int per_sm = 0, sms = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused_fft_kernel<X, false, true>, X::TX * X::FPB,
                                                          X::smem_bytes) != cudaSuccess)
            return 2;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (per_sm < 1) per_sm = 1;
        resident[dev] = per_sm * sms;
        ready_mask |= 1 << dev;
    }
    const int resident_dev = resident[dev];
    lock.unlock();
    const long long groups = (batch + X::FPB - 1) / X::FPB;
    if (groups <= 0) return 0;
    long long grid = groups;
    const long long cap = (long long)resident_dev * fused_waves();
    if (cap > 0 && grid > cap) grid = cap;
    dim3 block(X::TX, X::FPB);
    fused_fft_kernel<X, false, true><<<(unsigned)grid, block, X::smem_bytes, s>>>(
        (const cx<T> *)in, (cx<T> *)out, (const cx<T> *)tw, (const cx<T> *)rtw, batch, inverse, mode,
        *static_cast<const FusedIo<T> *>(io));
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

template <typename Cfg>
int launch_cfg_ex(const void *tw, const void *in, void *out, long long batch, int inverse, int mode, const void *rtw,
                  const void *io, cudaStream_t s) {
    return launch_ex_as<NoStaging<Cfg>>(tw, in, out, batch, inverse, mode, rtw, io, s);
}

template <typename Cfg>
FusedEntry make_entry(const char *name) {
    FusedEntry e;
    e.prec = sizeof(typename Cfg::T) == 4 ? 0 : 1;
    e.n = Cfg::N;
    e.name = name;
    e.tw_total = Cfg::tw_total;
    e.np = Cfg::NP;
    for (int i = 0; i < 4; ++i) e.radix[i] = Cfg::radix(i);
    e.real_only = 0;
    e.launch = &launch_cfg<Cfg>;
    e.launch_ex = nullptr;
    e.launch_ex_cols = nullptr;
#ifndef SSFFT_NO_EX_KERNELS  // tools/kbench.cu times hundreds of plain variants: no extended-I/O instantiations there
    e.launch_ex = &launch_cfg_ex<Cfg>;
    if constexpr (column_fpb<Cfg>() > 0) e.launch_ex_cols = &launch_ex_as<ColumnCfg<Cfg>>;
#endif
    return e;
}

#define SSFFT_FUSED(T, N, R0, R1, R2, R3, TX, FPB, MINB) \
    make_entry<FusedCfg<T, N, R0, R1, R2, R3, TX, FPB, MINB>>(#T "_" #N "_" #R0 "x" #R1 "x" #R2 "x" #R3)
// general form: PADS = shared-memory padding shift (one extra element per 2^PADS; 31 = none), PF = TMA prefetch.
// The padding per size comes from an offline bank-conflict model of the exchange.
#define SSFFT_FUSED_X(T, N, R0, R1, R2, R3, TX, FPB, MINB, PADS, PF) \
    make_entry<FusedCfg<T, N, R0, R1, R2, R3, TX, FPB, MINB, PADS, PF>>(#T "_" #N "_" #R0 "x" #R1 "x" #R2 "x" #R3 "_p" #PADS "_tma" #PF)
// same, with the TMA bulk-copy prefetch of the next transform group (cp.async.bulk + mbarrier)
#define SSFFT_FUSED_PF(T, N, R0, R1, R2, R3, TX, FPB, MINB) \
    make_entry<FusedCfg<T, N, R0, R1, R2, R3, TX, FPB, MINB, 4, 1>>(#T "_" #N "_" #R0 "x" #R1 "x" #R2 "x" #R3 "_tma")

template <typename Cfg>
FusedEntry make_real_entry(const char *name) {
    FusedEntry e = make_entry<Cfg>(name);
    e.real_only = 1;
    return e;
}
// entry used for RealFFT plans only (R2C epilogue / C2R on-the-fly gather have their own optimum: tools/kbench.cu real)
#define SSFFT_FUSED_REAL(T, N, R0, R1, R2, R3, TX, FPB, MINB, PADS, PF) \
    make_real_entry<FusedCfg<T, N, R0, R1, R2, R3, TX, FPB, MINB, PADS, PF>>(#T "_" #N "_" #R0 "x" #R1 "x" #R2 "x" #R3 "_p" #PADS "_tma" #PF "_real")

}  // namespace ssfft
