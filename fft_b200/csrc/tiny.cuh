// tiny.cuh -- transforms of a handful of points (N <= 24): one THREAD per transform.
//
// The reference's benchmark starts at N = 1 (benchmark/benchmark.h:27-52: 1, 2, 3, 4, 6, 8, 9, 12, 16, 18, 24, ...); below
// the single-pass kernels (N >= 64) those lengths ran through the pass interpreter at 17-60 % of the HBM roofline.  Here a
// CTA moves 256 consecutive transforms through shared memory with fully coalesced accesses, and every thread transforms one
// of them in registers with the compile-time codelet Dft<N> (codelets.cuh) -- no twiddle table, no exchange between threads.
#pragma once
#include <cuda_runtime.h>

#include "codelets.cuh"
#include "cplx.cuh"

namespace ssfft {

constexpr int kTinyMax = 24;
constexpr bool tiny_supported(size_t n) { return n >= 1 && (n <= 16 || n == 18 || n == 20 || n == 24); }

#ifdef __CUDACC__

template <typename T, int N>
__global__ void __launch_bounds__(256) tiny_fft_kernel(const cx<T> *__restrict__ in, cx<T> *__restrict__ out, long long batch, int inverse) {
    constexpr int PER = 256, PITCH = N | 1;  // odd pitch: a thread walking its own transform meets no bank conflict
    extern __shared__ __align__(16) unsigned char tiny_smem[];
    cx<T> *sm = reinterpret_cast<cx<T> *>(tiny_smem);
    const int tid = threadIdx.x;
    for (long long b0 = (long long)blockIdx.x * PER; b0 < batch; b0 += (long long)gridDim.x * PER) {
        const long long left = batch - b0;
        const int cnt = left < PER ? (int)left : PER;
        const cx<T> *src = in + b0 * N;
        for (int i = tid; i < cnt * N; i += PER) sm[(i / N) * PITCH + (i % N)] = src[i];
        __syncthreads();
        if (tid < cnt) {
            cx<T> v[N];
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const cx<T> x = sm[tid * PITCH + j];
                v[j] = inverse ? cswap(x) : x;
            }
            Dft<N>::run(v);
#pragma unroll
            for (int j = 0; j < N; ++j) sm[tid * PITCH + j] = inverse ? cswap(v[j]) : v[j];
        }
        __syncthreads();
        cx<T> *dst = out + b0 * N;
        for (int i = tid; i < cnt * N; i += PER) dst[i] = sm[(i / N) * PITCH + (i % N)];
        __syncthreads();
    }
}

template <typename T, int N>
int launch_tiny_n(const void *in, void *out, long long batch, int inverse, cudaStream_t s) {
    constexpr size_t smem = (size_t)256 * (N | 1) * sizeof(cx<T>);
    static bool attr_set[64] = {false};
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); return 2; }
    if (smem > 48 * 1024 && !attr_set[dev]) {
        if (cudaFuncSetAttribute(tiny_fft_kernel<T, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 2; }
        attr_set[dev] = true;
    }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long blocks = (batch + 255) / 256;
    if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
    tiny_fft_kernel<T, N><<<(unsigned)blocks, 256, smem, s>>>((const cx<T> *)in, (cx<T> *)out, batch, inverse);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

template <typename T>
int launch_tiny(size_t n, const void *in, void *out, long long batch, int inverse, cudaStream_t s) {
    switch (n) {
#define SSFFT_TINY_CASE(K) case K: return launch_tiny_n<T, K>(in, out, batch, inverse, s);
        SSFFT_TINY_CASE(1) SSFFT_TINY_CASE(2) SSFFT_TINY_CASE(3) SSFFT_TINY_CASE(4) SSFFT_TINY_CASE(5) SSFFT_TINY_CASE(6)
        SSFFT_TINY_CASE(7) SSFFT_TINY_CASE(8) SSFFT_TINY_CASE(9) SSFFT_TINY_CASE(10) SSFFT_TINY_CASE(11) SSFFT_TINY_CASE(12)
        SSFFT_TINY_CASE(13) SSFFT_TINY_CASE(14) SSFFT_TINY_CASE(15) SSFFT_TINY_CASE(16) SSFFT_TINY_CASE(18) SSFFT_TINY_CASE(20)
        SSFFT_TINY_CASE(24)
#undef SSFFT_TINY_CASE
    }
    return 3;
}

#endif

}  // namespace ssfft
