// real_kernels.cuh -- stand-alone RealFFT pre/post-twiddle passes and small utility kernels.
//
// RealFFT<V>::fft post-twiddle  (signalsmith-fft.h:459-472) and RealFFT<V>::ifft pre-twiddle (:478-492),
// plus the ModifiedRealFFT rotations (:450-452, :497-498).  These separate passes are used when the
// complex core runs through the generic / four-step path; the fused single-pass kernels (fused.cuh) do
// the same arithmetic in their prologue / epilogue without the extra HBM round trip.
#pragma once
#include "cplx.cuh"

namespace ssfft {

// Forward post-twiddle for one pair (i, ci): Z -> spectrum bins.  tw = twiddlesMinusI[i].
template <typename T>
SSFFT_HD void r2c_pair(cx<T> zi, cx<T> zc, cx<T> tw, cx<T> &oi, cx<T> &oc) {
    const T half = (T)0.5;
    cx<T> odd = mk<T>((zi.x + zc.x) * half, (zi.y - zc.y) * half);    // (Zi + conj Zc)/2   :466
    cx<T> even_i = mk<T>((zi.x - zc.x) * half, (zi.y + zc.y) * half); // (Zi - conj Zc)/2   :467
    cx<T> rot = cmul(even_i, tw);                                     // :468
    oi = odd + rot;                                                   // :470
    oc = cconj(odd - rot);                                            // :471
}
// Inverse pre-twiddle for one pair (no 1/2: result is scaled by N overall, :486-491)
template <typename T>
SSFFT_HD void c2r_pair(cx<T> v, cx<T> v2, cx<T> tw, cx<T> &bi, cx<T> &bc) {
    cx<T> odd = mk<T>(v.x + v2.x, v.y - v2.y);   // v + conj v2
    cx<T> rot = mk<T>(v.x - v2.x, v.y + v2.y);   // v - conj v2
    cx<T> even_i = cmulc(rot, tw);               // * conj(tw)   :488
    bi = odd + even_i;
    bc = cconj(odd - even_i);
}

#ifdef __CUDACC__
// In place on `data` (batch x h complex): Z -> packed half spectrum.
template <typename T>
__global__ void r2c_post_kernel(cx<T> *__restrict__ data, const cx<T> *__restrict__ tw, long long h, long long batch,
                                int modified) {
    const long long per = h / 2 + 1;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= per * batch) return;
    const long long b = idx / per, i = idx - b * per;
    cx<T> *z = data + b * h;
    if (!modified) {
        if (i == 0) {
            cx<T> z0 = z[0];
            z[0] = mk<T>(z0.x + z0.y, z0.x - z0.y);  // DC in .re, Nyquist in .im  (:459-462)
            return;
        }
        const long long ci = h - i;
        cx<T> oi, oc;
        r2c_pair(z[i], z[ci], tw[i], oi, oc);
        z[i] = oi;
        z[ci] = oc;  // when i == ci the second write wins, as in the reference
    } else {
        const long long ci = h - 1 - i;
        if (ci < i) return;
        cx<T> oi, oc;
        r2c_pair(z[i], z[ci], tw[i], oi, oc);
        z[i] = oi;
        z[ci] = oc;
    }
}

// Out of place: packed half spectrum (batch x h) -> pre-twiddled complex buffer (batch x h).
template <typename T>
__global__ void c2r_pre_kernel(const cx<T> *__restrict__ in, cx<T> *__restrict__ out, const cx<T> *__restrict__ tw,
                               long long h, long long batch, int modified) {
    const long long per = h / 2 + 1;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= per * batch) return;
    const long long b = idx / per, i = idx - b * per;
    const cx<T> *x = in + b * h;
    cx<T> *y = out + b * h;
    if (!modified) {
        if (i == 0) {
            cx<T> v = x[0];
            y[0] = mk<T>(v.x + v.y, v.x - v.y);  // :478-481
            return;
        }
        const long long ci = h - i;
        cx<T> bi, bc;
        c2r_pair(x[i], x[ci], tw[i], bi, bc);
        y[i] = bi;
        y[ci] = bc;
    } else {
        const long long ci = h - 1 - i;
        if (ci < i) return;
        cx<T> bi, bc;
        c2r_pair(x[i], x[ci], tw[i], bi, bc);
        y[i] = bi;
        y[ci] = bc;
    }
}

// dst[b][i] = src[b][i] * rot[i]   (conj_rot: * conj(rot[i]))   -- ModifiedRealFFT rotations
template <typename T>
__global__ void rotate_kernel(cx<T> *__restrict__ dst, const cx<T> *__restrict__ src, const cx<T> *__restrict__ rot,
                              long long h, long long total, int conj_rot) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const cx<T> r = rot[idx % h];
    const cx<T> v = src[idx];
    dst[idx] = conj_rot ? cmulc(v, r) : cmul(v, r);
}

// ---- unfused form of the extended I/O (ssfft_exec_*_ex) for plans without a fused EX kernel ----
// dst[b * dst_dist + e * dst_stride] = src[b * src_dist + e * src_stride] * mul[b * mul_dist + e],  e < len, b < batch.
// E is the element type of the side: T on the real side of a real plan, cx<T> elsewhere.  mul_kind 1: real table,
// 2: complex table (E = cx<T> only); packed0: the row is a RealFFT half spectrum, element 0 = (DC, Nyquist) is
// multiplied component by component.  Consecutive threads take consecutive e: the contiguous side is coalesced.
template <typename T>
SSFFT_HD T ex_apply_mul_real(T v, const T *mul, int mul_kind, long long at) { return mul_kind == 1 ? v * mul[at] : v; }
template <typename T>
SSFFT_HD cx<T> ex_apply_mul_cx(cx<T> v, const T *mul, int mul_kind, long long at, bool first_packed) {
    if (mul_kind == 1) return mk<T>(v.x * mul[at], v.y * mul[at]);
    if (mul_kind == 2) {
        const cx<T> w = reinterpret_cast<const cx<T> *>(mul)[at];
        return first_packed ? mk<T>(v.x * w.x, v.y * w.y) : cmul(v, w);
    }
    return v;
}
#ifdef __CUDACC__
template <typename T, bool REAL_SIDE>
__global__ void ex_copy_kernel(const void *__restrict__ src_, void *__restrict__ dst_, long long len, long long batch,
                               long long src_dist, long long src_stride, long long dst_dist, long long dst_stride,
                               const T *__restrict__ mul, int mul_kind, long long mul_dist, int packed0) {
    const long long total = len * batch;
    // (b, e) advance incrementally: a 64-bit division per element would make this copy instruction-bound
    const long long step = (long long)gridDim.x * blockDim.x, step_b = step / len, step_e = step - step_b * len;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long b = idx / len, e = idx - b * len;
#pragma unroll 4
    for (; idx < total; idx += step, b += step_b, e += step_e) {
        if (e >= len) { e -= len; ++b; }
        if constexpr (REAL_SIDE) {
            const T *src = static_cast<const T *>(src_);
            T *dst = static_cast<T *>(dst_);
            dst[b * dst_dist + e * dst_stride] = ex_apply_mul_real<T>(src[b * src_dist + e * src_stride], mul, mul_kind, b * mul_dist + e);
        } else {
            const cx<T> *src = static_cast<const cx<T> *>(src_);
            cx<T> *dst = static_cast<cx<T> *>(dst_);
            dst[b * dst_dist + e * dst_stride] =
                ex_apply_mul_cx<T>(src[b * src_dist + e * src_stride], mul, mul_kind, b * mul_dist + e, packed0 && e == 0);
        }
    }
}
#endif

// ---- local building blocks of the distributed four-step (fft_b200/dist.py) ----
// out[b][c][r] = in[b][r][c] * W_N^((row0 + r) * c)^(+-1)   (twiddle optional: n_total == 0 -> plain transpose)
// 32x32 tiles through shared memory, both sides coalesced.  The twiddle phase is evaluated in double
// (sincospi) from the exact integer (row*col) mod N, so it is correct for N up to 2^40 in both precisions.
// Each CTA walks kRowTiles consecutive 32x32 tiles down the rows.  A thread always sees the same column c and rows
// that advance by 8, so its twiddle is W^((row0+r) c) with r stepping by 8: one exact sincospi for the start, one
// for the step W^(8c), then a rotation recurrence in DOUBLE (<= 4*kRowTiles steps, error ~1e-15) -- 16x fewer
// sincospi calls than one per element (measured: 7.6 ms -> plain-transpose speed for 4 GiB).
constexpr int kRowTiles = 8;
template <typename T>
__global__ void transpose_twiddle_kernel(const cx<T> *__restrict__ in, cx<T> *__restrict__ out, long long rows,
                                         long long cols, long long row0, unsigned long long n_total, int conj_tw) {
    __shared__ cx<T> tile[32][33];
    const long long b = blockIdx.z;
    const long long rbase = (long long)blockIdx.y * 32 * kRowTiles, c0 = (long long)blockIdx.x * 32;
    const cx<T> *src = in + b * rows * cols;
    cx<T> *dst = out + b * rows * cols;
    const long long c = c0 + threadIdx.x;
    double wr = 1.0, wi = 0.0, sr = 1.0, si = 0.0;  // current twiddle and the per-8-rows step
    if (n_total && c < cols) {
        const unsigned long long q0 = (unsigned long long)(((unsigned __int128)(unsigned long long)(row0 + rbase + threadIdx.y) *
                                                            (unsigned long long)c) % n_total);
        const unsigned long long qs = (unsigned long long)(((unsigned __int128)8ull * (unsigned long long)c) % n_total);
        sincospi(-2.0 * (double)q0 / (double)n_total, &wi, &wr);
        sincospi(-2.0 * (double)qs / (double)n_total, &si, &sr);
        if (conj_tw) { wi = -wi; si = -si; }
    }
    for (int rt = 0; rt < kRowTiles; ++rt) {
        const long long r0 = rbase + 32 * rt;
        if (r0 >= rows) break;
        for (int i = threadIdx.y; i < 32; i += 8) {
            const long long r = r0 + i;
            if (r < rows && c < cols) {
                cx<T> v = src[r * cols + c];
                if (n_total) v = cmul(v, mk<T>((T)wr, (T)wi));
                tile[i][threadIdx.x] = v;
            }
            const double nr = wr * sr - wi * si, ni = wr * si + wi * sr;  // advance 8 rows
            wr = nr; wi = ni;
        }
        __syncthreads();
        for (int i = threadIdx.y; i < 32; i += 8) {
            const long long cc = c0 + i, r = r0 + threadIdx.x;
            if (r < rows && cc < cols) dst[cc * rows + r] = tile[threadIdx.x][i];
        }
        __syncthreads();
    }
}

// ---- fused exchange of the distributed four-step: transpose (+ twiddle) + all-to-all + placement in ONE kernel.
// The local matrix src[rows][cols] is split into `world` column blocks of `blk` columns; block j goes, transposed,
// straight into rank j's HBM over NVLink (peer pointers from CUDA IPC) at dst_j[cl * dst_pitch + dst_col0 + r]:
// exactly where the next local FFT wants it, so no pack, no NCCL staging and no unpack pass exist.
// Stores to a peer are 256-byte runs (32 consecutive r); the twiddle uses the same double recurrence as above.
constexpr int kMaxPeers = 16;
template <typename T>
struct PeerPtrs {
    cx<T> *p[kMaxPeers];
};
template <typename T>
__global__ void exchange_transpose_kernel(const cx<T> *__restrict__ src, PeerPtrs<T> dst, long long rows, long long cols,
                                          long long blk, long long dst_pitch, long long dst_col0, long long row0,
                                          unsigned long long n_total, int conj_tw) {
    __shared__ cx<T> tile[32][33];
    const long long rbase = (long long)blockIdx.y * 32 * kRowTiles, c0 = (long long)blockIdx.x * 32;
    const long long c = c0 + threadIdx.x;
    double wr = 1.0, wi = 0.0, sr = 1.0, si = 0.0;
    if (n_total && c < cols) {
        const unsigned long long q0 = (unsigned long long)(((unsigned __int128)(unsigned long long)(row0 + rbase + threadIdx.y) *
                                                            (unsigned long long)c) % n_total);
        const unsigned long long qs = (unsigned long long)(((unsigned __int128)8ull * (unsigned long long)c) % n_total);
        sincospi(-2.0 * (double)q0 / (double)n_total, &wi, &wr);
        sincospi(-2.0 * (double)qs / (double)n_total, &si, &sr);
        if (conj_tw) { wi = -wi; si = -si; }
    }
    for (int rt = 0; rt < kRowTiles; ++rt) {
        const long long r0 = rbase + 32 * rt;
        if (r0 >= rows) break;
        for (int i = threadIdx.y; i < 32; i += 8) {
            const long long r = r0 + i;
            if (r < rows && c < cols) {
                cx<T> v = src[r * cols + c];
                if (n_total) v = cmul(v, mk<T>((T)wr, (T)wi));
                tile[i][threadIdx.x] = v;
            }
            const double nr = wr * sr - wi * si, ni = wr * si + wi * sr;
            wr = nr; wi = ni;
        }
        __syncthreads();
        for (int i = threadIdx.y; i < 32; i += 8) {
            const long long cc = c0 + i, r = r0 + threadIdx.x;
            if (r < rows && cc < cols) {
                const long long j = cc / blk, cl = cc - j * blk;
                dst.p[j][cl * dst_pitch + dst_col0 + r] = tile[threadIdx.x][i];  // peer (or own) HBM
            }
        }
        __syncthreads();
    }
}

// out[b][a][c] = in[a][b][c]: swaps the two outer dimensions, moving contiguous runs of `run` elements
template <typename T>
__global__ void permute102_kernel(const cx<T> *__restrict__ in, cx<T> *__restrict__ out, long long A, long long B,
                                  long long run) {
    // one (a, b) run per blockIdx.y slot, 16-byte copies along the run when it is aligned
    const long long pairs = A * B;
    for (long long ab = blockIdx.y; ab < pairs; ab += gridDim.y) {
        const long long a = ab % A, b = ab / A;  // enumerate the OUTPUT order [b][a]
        const cx<T> *s = in + (a * B + b) * run;
        cx<T> *d = out + ab * run;
        if (sizeof(cx<T>) == 8 && (run % 2 == 0)) {
            const float4 *s4 = reinterpret_cast<const float4 *>(s);
            float4 *d4 = reinterpret_cast<float4 *>(d);
            for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < run / 2; i += (long long)gridDim.x * blockDim.x)
                d4[i] = s4[i];
        } else {
            for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < run; i += (long long)gridDim.x * blockDim.x)
                d[i] = s[i];
        }
    }
}

// counter-based uniform [-0.5, 0.5) generator, twin of oracle_fill_uniform_* 
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
template <typename T>
__global__ void fill_uniform_kernel(T *__restrict__ dst, unsigned long long count, unsigned long long seed,
                                    unsigned long long first) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        unsigned long long h = splitmix64((seed << 40) + first + i);
        if (sizeof(T) == 4) dst[i] = (T)((float)(h >> 40) * (1.0f / 16777216.0f) - 0.5f);
        else dst[i] = (T)((double)(h >> 11) * (1.0 / 9007199254740992.0) - 0.5);
    }
}
#endif

}  // namespace ssfft
