// composite.cuh -- N = R * M with a small radix R (2 ... 16) around one of the fast plans of length M.
//
// The reference factorises EVERY length into radix-2/3/4/generic steps (signalsmith-fft.h:99-150) and its benchmark walks
// 2^k * {1, 3, 9} up to 2^24 (benchmark/benchmark.h:27-52).  The specialised kernels of this library cover 2^k up to 2^20
// and 2^k * {3, 9} up to 9216; the lengths above used to fall to the pass interpreter (generic x generic, 7-15 % of the
// HBM roofline).  This plan is one decimation-in-time step of radix R around the best plan of length M = N / R:
//
//   pre   y[k1][c]  = W_N^(c k1) * sum_r x[r M + c] W_R^(r k1)        (radix_pass_kernel: R strided rows in, R rows out,
//                                                                      every access coalesced along c; the twiddle is the
//                                                                      k1-th power of ONE table value W_N^c)
//   inner Z[k1][.]  = FFT_M(y[k1][.])                                  (any plan of the library, R * batch transforms, in place)
//   post  X[k1 + R k2] = Z[k1][k2]                                     (interleave_kernel: R rows in, one contiguous run out,
//                                                                      staged through shared memory)
//
// Three trips through HBM instead of one: 15-25 % of the roofline where the pass interpreter reached 7-9 % (measured,
// profiles/sweep_r02n_f32.txt).  Keeping y / Z in L2 between the launches (short trips) was measured slower than streaming.
// The next step for these lengths is the ticket-queue four-step with a 3 * 2^k / 9 * 2^k row stage (two trips, one in L2).
// Inverse transforms swap re / im on the way in and out, the same identity the other kernels use.
#pragma once
#include <cuda_runtime.h>

#include <type_traits>
#include <vector>

#include "codelets.cuh"
#include "cplx.cuh"

namespace ssfft {

#ifdef __CUDACC__

// w[k] = g^k for k < R (w[0] unused), products at most ~log2(R) deep
template <int R, typename T>
__device__ __forceinline__ void small_powers(cx<T> (&w)[R], cx<T> g) {
    if constexpr (R > 1) w[1] = g;
#pragma unroll
    for (int k = 2; k < R; ++k) w[k] = cmul(w[k / 2], w[k - k / 2]);
}

// one column per thread, 8-byte accesses in fp32: for buffers that are not 16-byte aligned or an odd inner length
template <typename T, int R>
__global__ void __launch_bounds__(256) radix_pass_scalar_kernel(const cx<T> *__restrict__ in, cx<T> *__restrict__ out, const cx<T> *__restrict__ tw,
                                                                 long long m, long long batch, int inverse) {
    const long long total = m * batch;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / m, c = i - b * m;
        const cx<T> *src = in + b * m * R + c;
        cx<T> v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const cx<T> x = src[(long long)r * m];
            v[r] = inverse ? cswap(x) : x;
        }
        Dft<R>::run(v);
        cx<T> w[R];
        small_powers<R>(w, tw[c]);
        cx<T> *dst = out + b * m * R + c;
        dst[0] = v[0];
#pragma unroll
        for (int k = 1; k < R; ++k) dst[(long long)k * m] = cmul(v[k], w[k]);
    }
}

// V columns per thread: V * sizeof(cx<T>) = 16 bytes per access (two fp32 columns, one fp64 column)
template <typename T, int R>
__global__ void __launch_bounds__(256) radix_pass_kernel(const cx<T> *__restrict__ in, cx<T> *__restrict__ out, const cx<T> *__restrict__ tw,
                                                          long long m, long long batch, int inverse) {
    constexpr int V = sizeof(T) == 4 ? 2 : 1;
    using Vec = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
    const long long mv = m / V, total = mv * batch;  // m is even whenever V = 2 (checked by the launcher)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / mv, c = (i - b * mv) * V;
        const cx<T> *src = in + b * m * R + c;
        Vec raw[R];
#pragma unroll
        for (int r = 0; r < R; ++r) raw[r] = __ldcs(reinterpret_cast<const Vec *>(src + (long long)r * m));
        const Vec twv = *reinterpret_cast<const Vec *>(tw + c);
        cx<T> *dst = out + b * m * R + c;
        Vec res[R];
#pragma unroll
        for (int col = 0; col < V; ++col) {
            cx<T> v[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                cx<T> x;
                if constexpr (V == 2) x = col == 0 ? mk<T>(raw[r].x, raw[r].y) : mk<T>(raw[r].z, raw[r].w);
                else x = mk<T>(raw[r].x, raw[r].y);
                v[r] = inverse ? cswap(x) : x;
            }
            Dft<R>::run(v);
            cx<T> g;
            if constexpr (V == 2) g = col == 0 ? mk<T>(twv.x, twv.y) : mk<T>(twv.z, twv.w);
            else g = mk<T>(twv.x, twv.y);
            cx<T> w[R];
            small_powers<R>(w, g);
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const cx<T> y = k ? cmul(v[k], w[k]) : v[0];
                if constexpr (V == 2) {
                    if (col == 0) { res[k].x = y.x; res[k].y = y.y; }
                    else { res[k].z = y.x; res[k].w = y.y; }
                } else {
                    res[k].x = y.x; res[k].y = y.y;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < R; ++k) *reinterpret_cast<Vec *>(dst + (long long)k * m) = res[k];
    }
}

// out[b][k1 + R k2] = in[b][k1 M + k2].  Every WARP moves tiles of 32 k2 x R through its own slice of shared memory (no
// block-wide barrier: the loads of one warp run under the stores of the others; 32 k2 in fp32, 16 in fp64) so that both sides are whole, aligned
// runs; with an even R and a 16-byte aligned output a lane stores two neighbouring elements at once.
template <typename T, int R>
__global__ void __launch_bounds__(256) interleave_kernel(const cx<T> *__restrict__ in, cx<T> *__restrict__ out, long long m,
                                                          long long batch, int inverse, int vec_ok) {
    constexpr int W = sizeof(T) == 8 ? 16 : 32, PITCH = W + 1, WARPS = 8;  // k2 per tile (fp64: 35 KB of shared memory at R = 16)
    __shared__ cx<T> sm_all[WARPS * R * PITCH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    cx<T> *sm = sm_all + warp * R * PITCH;
    const long long tiles_per = (m + W - 1) / W, tiles = tiles_per * batch;
    for (long long t = (long long)blockIdx.x * WARPS + warp; t < tiles; t += (long long)gridDim.x * WARPS) {
        const long long b = t / tiles_per, k20 = (t - b * tiles_per) * W;
        const int span = (int)(m - k20 < W ? m - k20 : W);
        const cx<T> *src = in + b * m * R + k20;
        if (lane < span) {
            cx<T> v[R];
#pragma unroll
            for (int k1 = 0; k1 < R; ++k1) v[k1] = src[(long long)k1 * m + lane];
#pragma unroll
            for (int k1 = 0; k1 < R; ++k1) sm[k1 * PITCH + lane] = inverse ? cswap(v[k1]) : v[k1];
        }
        __syncwarp();
        cx<T> *dst = out + b * m * R + k20 * R;
        const int n_out = span * R;
        if constexpr (sizeof(T) == 4 && R % 2 == 0) {
            if (vec_ok) {
                for (int j = 2 * lane; j < n_out; j += 64) {
                    const cx<T> a = sm[(j % R) * PITCH + j / R], c = sm[((j + 1) % R) * PITCH + (j + 1) / R];
                    *reinterpret_cast<float4 *>(dst + j) = make_float4(a.x, a.y, c.x, c.y);
                }
            } else {
                for (int j = lane; j < n_out; j += 32) dst[j] = sm[(j % R) * PITCH + j / R];
            }
        } else {
            for (int j = lane; j < n_out; j += 32) dst[j] = sm[(j % R) * PITCH + j / R];
        }
        __syncwarp();
    }
}

#endif  // __CUDACC__

// tw[c] = W_N^c, c < m (interleaved re, im), N = r * m
template <typename T>
inline void fill_composite_twiddles(std::vector<T> &h, size_t r, size_t m) {
    h.assign(2 * m, (T)0);
    const long double pi2 = 2.0L * 3.14159265358979323846264338327950288L;
    const long double n = (long double)r * (long double)m;
    for (size_t c = 0; c < m; ++c) {
        const long double a = pi2 * (long double)c / n;
        h[2 * c] = (T)cosl(a);
        h[2 * c + 1] = (T)(-sinl(a));
    }
}

}  // namespace ssfft
