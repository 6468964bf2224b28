// bluestein.cuh -- chirp-z (Bluestein) transform for lengths whose largest prime factor does not fit one CTA.
//
// The reference handles ANY length: trial division leaves a large prime as one factor (signalsmith-fft.h:146-150) and
// fftStepGeneric runs it as an O(p^2) loop (:187-215).  On the GPU a prime p above the shared-memory limit of the
// generic pass interpreter (about 13.6 K points in fp32, 6.8 K in fp64) used to be rejected.  Bluestein's identity
//      n k = (n^2 + k^2 - (k - n)^2) / 2
// turns the length-N DFT into a convolution:  X[k] = c[k] * sum_n (x[n] c[n]) * conj(c[k - n]),  c[n] = exp(-i pi n^2 / N),
// which is computed with the fast power-of-two kernels of this library at length M = 2^ceil(log2(2N - 1)):
//      a = zero-padded x * c   ->  FFT_M  ->  * FFT_M(conj c, wrapped)  ->  IFFT_M  ->  * c / M.
// O(M log M) instead of O(N p); the phases n^2 mod 2N are exact integers, roots are evaluated in long double.
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "cplx.cuh"

namespace ssfft {

#ifdef __CUDACC__
// a[b][j] = x[b][j] * c[j] for j < n, 0 for n <= j < m.  The inverse transform swaps re / im on the way in and out.
template <typename T>
__global__ void bluestein_pre_kernel(const cx<T> *__restrict__ in, cx<T> *__restrict__ a, const cx<T> *__restrict__ chirp,
                                     long long n, long long m, long long batch, int inverse) {
    for (long long b = blockIdx.y; b < batch; b += gridDim.y)
        for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (long long)gridDim.x * blockDim.x) {
            cx<T> v = mk<T>((T)0, (T)0);
            if (j < n) {
                const cx<T> x = in[b * n + j];
                v = cmul(inverse ? cswap(x) : x, chirp[j]);
            }
            a[b * m + j] = v;
        }
}
template <typename T>
__global__ void bluestein_mul_kernel(cx<T> *__restrict__ a, const cx<T> *__restrict__ filt, long long m, long long batch) {
    for (long long b = blockIdx.y; b < batch; b += gridDim.y)
        for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (long long)gridDim.x * blockDim.x)
            a[b * m + j] = cmul(a[b * m + j], filt[j]);
}
// out[b][k] = y[b][k] * c[k] / m for k < n
template <typename T>
__global__ void bluestein_post_kernel(const cx<T> *__restrict__ y, cx<T> *__restrict__ out, const cx<T> *__restrict__ chirp,
                                      long long n, long long m, long long batch, T scale, int inverse) {
    for (long long b = blockIdx.y; b < batch; b += gridDim.y)
        for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
            cx<T> v = cmul(y[b * m + k], chirp[k]);
            v = mk<T>(v.x * scale, v.y * scale);
            out[b * n + k] = inverse ? cswap(v) : v;
        }
}
#endif

inline size_t bluestein_length(size_t n) {
    size_t m = 1;
    while (m < 2 * n - 1) m <<= 1;
    return m;
}
// chirp[j] = exp(-i pi j^2 / n), j < n (phase j^2 mod 2n, exact);  wrapped[j] = conj(chirp[|j|]) at j and m - j
template <typename T>
inline void fill_bluestein_tables(std::vector<T> &chirp, std::vector<T> &wrapped, size_t n, size_t m) {
    chirp.assign(2 * n, (T)0);
    wrapped.assign(2 * m, (T)0);
    const long double pi = 3.14159265358979323846264338327950288L;
    for (size_t j = 0; j < n; ++j) {
        const unsigned __int128 sq = (unsigned __int128)j * j;
        const unsigned long long q = (unsigned long long)(sq % (2 * (unsigned __int128)n));
        const long double a = pi * (long double)q / (long double)n;
        const T c = (T)cosl(a), s = (T)sinl(a);
        chirp[2 * j] = c; chirp[2 * j + 1] = -s;
        wrapped[2 * j] = c; wrapped[2 * j + 1] = s;
        if (j) { wrapped[2 * (m - j)] = c; wrapped[2 * (m - j) + 1] = s; }
    }
}

}  // namespace ssfft
