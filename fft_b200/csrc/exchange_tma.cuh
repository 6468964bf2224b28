// exchange_tma.cuh -- the fused exchange of the distributed four-step (transpose + twiddle + all-to-all + placement,
// real_kernels.cuh: exchange_transpose_kernel) driven by the TMA engine instead of by thousands of resident warps.
//
// exchange_transpose_kernel keeps NVLink 93 % busy (ncu, profiles/exchange_transpose_r02x_*) -- but with 64 warps per SM,
// every one of them waiting on its own 8-byte loads and stores (49 % of stall samples on the long scoreboard).  Beside a
// transform kernel it gets a quarter of those warps and a quarter of the bandwidth, which is why chunked phases bought
// nothing.  Here one CTA of 128 threads per SM is enough: tiles [64 rows][32 columns] come in as ONE tensor copy each
// (ring of three), the threads transpose them through shared memory (multiplying by W_N^((row0 + r) c) where the phase
// needs it), and the 32 columns of a tile leave as 32 bulk stores of 512 bytes straight into the peer's HBM.  The copies in
// flight belong to the TMA unit, not to warps, so the kernel keeps its rate with 4 warps and 50 KB of shared memory per SM.
//
// RESULT (profiles/bench_dist_tma_r02y.txt, 2^30 over 2 GPUs, parity green): correct, but SLOWER than the kernel it was
// meant to replace -- 14.9 ms per transform with 2 CTAs per SM (16.2 with 1, 14.8 with 4) against 13.9 ms -- and the exchange
// of a chunk still does not hide behind the transform of the next one (14.5 ms at best with 4 chunks).  A CTA's tiles go to
// the local and to the remote device in turn; when the remote stores queue behind NVLink the CTA's local tiles wait with
// them, which 64 independent warps per SM do not.  Off by default (SSFFT_EXCHANGE_TMA=1 selects it); kept as the record
// of the experiment.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "fused.cuh"
#include "real_kernels.cuh"
#include "tma_host.cuh"

namespace ssfft {

#if defined(__CUDACC__) && !defined(SSFFT_EMUL)

template <typename T>
struct ExchangeTmaCfg {
    static constexpr int TR = 64;                        // rows per tile = elements per bulk store (512 B fp32, 1 KiB fp64)
    static constexpr int TC = sizeof(T) == 4 ? 32 : 16;  // columns per tile
    static constexpr int THREADS = 128, NIN = 3, NOUT = 2;
    static constexpr int PITCH = TR + 2;                 // output rows 16 bytes apart from a power of two: 2-way conflicts only
    static constexpr size_t kIn = (size_t)TR * TC * sizeof(cx<T>), kOut = (size_t)TC * PITCH * sizeof(cx<T>);
    static constexpr size_t smem_bytes = NIN * kIn + NOUT * kOut + 64;
};

__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <typename T>
__global__ void __launch_bounds__(ExchangeTmaCfg<T>::THREADS)
exchange_tma_kernel(const __grid_constant__ CUtensorMap tmap, PeerPtrs<T> dst, long long rows, long long cols, long long blk,
                    long long dst_pitch, long long dst_col0, long long row0, unsigned long long n_total, int conj_tw) {
    using C = ExchangeTmaCfg<T>;
    constexpr int TR = C::TR, TC = C::TC, G = C::THREADS / TC, PER = TR / G, PITCH = C::PITCH;
    extern __shared__ __align__(128) unsigned char smem[];
    cx<T> *in_buf = reinterpret_cast<cx<T> *>(smem);
    cx<T> *out_buf = reinterpret_cast<cx<T> *>(smem + C::NIN * C::kIn);
    unsigned long long *full = reinterpret_cast<unsigned long long *>(smem + C::NIN * C::kIn + C::NOUT * C::kOut);
    const int tid = threadIdx.x, c = tid % TC, rg = tid / TC;
    const long long col_tiles = cols / TC, tiles = (rows / TR) * col_tiles;
    if (tid == 0)
        for (int s = 0; s < C::NIN; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    auto issue = [&](long long t, int s) {  // tile t -> input slot s (thread 0)
        const long long rb = t / col_tiles, cb = t - rb * col_tiles;
        mbar_expect_tx(&full[s], (unsigned)C::kIn);
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                smem_u32(in_buf + (size_t)s * TR * TC)),
            "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"((int)(cb * TC) * (int)(sizeof(cx<T>) / 8)), "r"((int)(rb * TR)), "r"(0),
            "r"(smem_u32(&full[s]))
            : "memory");
    };
    if (tid == 0)
        for (int s = 0; s < C::NIN; ++s) {
            const long long t = (long long)blockIdx.x + (long long)s * gridDim.x;
            if (t < tiles) issue(t, s);
        }
    long long it = 0;
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
        const int s = (int)(it % C::NIN), o = (int)(it % C::NOUT);
        const long long rb = t / col_tiles, cb = t - rb * col_tiles;
        const long long r0 = rb * TR, c0 = cb * TC;
        // the output buffer about to be rewritten: its stores (two tiles ago) must have read it
        if (tid < TC) bulk_wait_read<C::NOUT - 1>();
        double wr = 1.0, wi = 0.0, sr = 1.0, si = 0.0;
        if (n_total) {
            const unsigned long long cc = (unsigned long long)(c0 + c);
            const unsigned long long q0 = (unsigned long long)(((unsigned __int128)(unsigned long long)(row0 + r0 + rg) * cc) % n_total);
            const unsigned long long qs = (unsigned long long)(((unsigned __int128)(unsigned long long)G * cc) % n_total);
            sincospi(-2.0 * (double)q0 / (double)n_total, &wi, &wr);
            sincospi(-2.0 * (double)qs / (double)n_total, &si, &sr);
            if (conj_tw) { wi = -wi; si = -si; }
        }
        __syncthreads();  // everybody is past the previous tile; the wait above has been done by the storing lanes
        mbar_wait(&full[s], (unsigned)((it / C::NIN) & 1));
        const cx<T> *in = in_buf + (size_t)s * TR * TC;
        cx<T> *out = out_buf + (size_t)o * TC * PITCH;
#pragma unroll 4
        for (int i = 0; i < PER; ++i) {
            const int r = rg + G * i;
            cx<T> v = in[r * TC + c];
            if (n_total) {
                v = cmul(v, mk<T>((T)wr, (T)wi));
                const double nr = wr * sr - wi * si, ni = wr * si + wi * sr;  // G rows further
                wr = nr; wi = ni;
            }
            out[c * PITCH + r] = v;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // these stores -> the bulk copies below
        __syncthreads();
        if (tid < TC) {  // lane = column of the tile: one run of TR elements into the peer (or own) HBM
            const long long cc = c0 + tid, j = cc / blk, cl = cc - j * blk;
            bulk_s2g(dst.p[j] + cl * dst_pitch + dst_col0 + r0, out + tid * PITCH, (unsigned)(TR * sizeof(cx<T>)));
            bulk_commit();
        }
        if (tid == 0) {  // the input slot is free again: fetch the tile three rounds ahead
            const long long tn = t + (long long)C::NIN * gridDim.x;
            if (tn < tiles) issue(tn, s);
        }
    }
    if (tid < TC) bulk_wait_all();  // the stores have reached their destination before the kernel ends
}

// 0 launched; 3: shape / alignment not covered (the caller uses exchange_transpose_kernel)
template <typename T>
int launch_exchange_tma(const void *src, void *const *dst_ptrs, int world, size_t rows, size_t cols, size_t dst_pitch, size_t dst_col0,
                        size_t row0, unsigned long long n_total, int inverse, int max_ctas, cudaStream_t stream) {
    using C = ExchangeTmaCfg<T>;
    const size_t blk = cols / (size_t)world;
    if (rows % C::TR || cols % C::TC || blk % C::TC || (dst_pitch * sizeof(cx<T>)) % 16 || (dst_col0 * sizeof(cx<T>)) % 16) return 3;
    for (int i = 0; i < world; ++i)
        if (reinterpret_cast<uintptr_t>(dst_ptrs[i]) & 15u) return 3;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    if (!encode_tensor_map_3d(&tmap, src, 1, (int)rows, (int)cols, C::TR, C::TC, (int)sizeof(cx<T>))) return 3;
    static bool attr_set[64] = {false};  // per device (function attributes are per context)
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); return 3; }
    if (!attr_set[dev]) {
        if (cudaFuncSetAttribute(exchange_tma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_bytes) != cudaSuccess) {
            cudaGetLastError();
            return 3;
        }
        attr_set[dev] = true;
    }
    PeerPtrs<T> pp;
    for (int i = 0; i < kMaxPeers; ++i) pp.p[i] = i < world ? (cx<T> *)dst_ptrs[i] : nullptr;
    long long tiles = (long long)(rows / C::TR) * (long long)(cols / C::TC);
    long long ctas = tiles < max_ctas ? tiles : max_ctas;
    if (ctas < 1) return 0;
    exchange_tma_kernel<T><<<(unsigned)ctas, C::THREADS, C::smem_bytes, stream>>>(tmap, pp, (long long)rows, (long long)cols, (long long)blk,
                                                                                     (long long)dst_pitch, (long long)dst_col0, (long long)row0,
                                                                                     n_total, inverse);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

#endif

}  // namespace ssfft
