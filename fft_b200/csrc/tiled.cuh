// tiled.cuh -- four-step (N = N1 * N2) tile kernels for transforms too large for one CTA.
//
// GPU analogue of the reference's cache-blocking branch (addPlanSteps, signalsmith-fft.h:130-133): when a
// transform does not fit on chip it is split into N2 column FFTs of length N1, a twiddle W_N^(n2*k1), and
// N1 row FFTs of length N2 whose output is stored transposed, X[k1 + N1*k2].  The intermediate lives in a
// scratch buffer small enough to stay in the 126 MB L2, so HBM sees each input and output byte once.
//
// A CTA owns a TILE of CT adjacent columns (or rows): threadIdx.x = lane within the tile (so every global
// access is a 128-byte segment and every shared-memory access is conflict-free), threadIdx.y = butterfly
// thread.  Shared layout [idx][lane] with pitch CT+1 (lets the row kernels transpose on the way in/out).
//
// Real transforms (RealFFT<V>::fft/ifft, :446-502) use a REAL four-step instead of the reference's
// "N/2 complex + post-twiddle" trick, because that trick pairs bin k with bin N/2-k which live in different
// CTAs here.  Two adjacent real columns are packed into one complex column FFT and separated in shared
// memory; only rows k1 <= N1/2 are kept, and each row FFT emits its own bins (k2 < N2/2) plus, conjugated,
// the bins of row N1-k1 (k2 >= N2/2).  Same outputs (bins 0..N/2-1, (DC, Nyquist) packed in bin 0), no
// cross-CTA exchange, half the traffic of a complex transform.
//
//   (scratch = tile-major, see tm_off)
//   flavor      FFT length  tile over              reads                      writes
//   A_C2C       N1          n2 (columns)           x[n1][n2]                  Y[k1][n2] * W_N^(n2 k1)   (scratch)
//   B_C2C       N2          k1 (rows)              Y[k1][n2]                  X[k1 + N1 k2]
//   A_R2C       N1          n2/2 (column pairs)    real x as complex pairs    Y[k1<=N1/2][n2] * W       (scratch)
//   B_R2C       N2          k1 <= N1/2             Y[k1][n2]                  packed half spectrum
//   B_C2R       N2          k1 <= N1/2             packed half spectrum       U[k1][n2] * conj W        (scratch)
//   A_C2R       N1          n2/2 (column pairs)    U (Hermitian-extended)     real x as complex pairs
#pragma once
#include <cuda.h>  // CUtensorMap (types only)
#include <cuda_runtime.h>

#include <type_traits>

#include "codelets.cuh"
#include "fused.cuh"

// Same double buffering, but the tile is brought by the TMA engine: ONE tensor copy (stage 1: box [N1][CT] of the
// (batch, N1, N2) input) or ONE bulk copy (stage 2: the contiguous tile-major scratch block) per tile, signalled
// through an mbarrier -- no per-thread copy instructions, no registers.  Staging layout is dense ([idx][CT]).
// MEASURED on B200 (profiles/sweep_fourstep_tma_r01_float32.json): no gain for complex (65536: 50.1 % vs 50.5 %,
// 32768: 45.5 % both) and a loss for real 65536 (3.45 -> 4.03 ms per forward+inverse): with 3 CTAs/SM the loads of one
// CTA already overlap the butterflies of the others, the kernels are bound by issue slots and dependent latency.
// Off by default; parity tests pass with it on.
#ifndef SSFFT_FOURSTEP_TMA
#define SSFFT_FOURSTEP_TMA 0
#endif

namespace ssfft {

enum { TILE_A_C2C = 0, TILE_B_C2C = 1, TILE_A_R2C = 2, TILE_B_R2C = 3, TILE_B_C2R = 4, TILE_A_C2R = 5 };

template <typename T_, int L_, int R0_, int R1_, int R2_, int TX_, int CT_, int MINB_>
struct TileCfg {
    using T = T_;
    static constexpr int L = L_, TX = TX_, CT = CT_, MINB = MINB_;
    static constexpr int NP = (R2_ > 1) ? 3 : (R1_ > 1) ? 2 : 1;
    static constexpr int E = L / TX;
    static constexpr int PITCH = CT + 1;
    static constexpr int THREADS = TX * CT;
    __host__ __device__ static constexpr int radix(int i) { return i == 0 ? R0_ : i == 1 ? R1_ : R2_; }
    __host__ __device__ static constexpr int prod(int i) { return i == 0 ? 1 : i == 1 ? R0_ : R0_ * R1_; }
    __host__ __device__ static constexpr int mnext(int i) { return L / (prod(i) * radix(i)); }
    __host__ __device__ static constexpr int tw_off(int i) {
        int o = 0;
        for (int k = 0; k < i; ++k) o += (radix(k) - 1) * mnext(k);
        return o;
    }
    static constexpr int tw_total = tw_off(NP - 1);
    static constexpr size_t smem_bytes = (size_t)L * PITCH * sizeof(cx<T>);
    static_assert(R0_ * R1_ * R2_ == L_, "radices must multiply to L");
    static_assert(E % R0_ == 0 && E % R1_ == 0 && E % R2_ == 0, "E must be a multiple of every radix");
    // (the kernels of tiled.cuh / cluster.cuh are only instantiated for powers of two; flat.cuh also runs row stages of
    // length 3 * 2^k, whose radices are not powers of two)
};

template <typename T>
struct TileParams {
    const cx<T> *in;
    cx<T> *out;
    const cx<T> *tw;   // per-pass twiddles of the length-L transform ([r][m'] per pass, as fused.cuh)
    const cx<T> *tw4;  // four-step twiddles W_N^(n2*k1) laid out [k1][n2]
    int n1, n2;        // N = n1 * n2 (REAL length for the real flavors)
    long long batch;
    long long in_stride, out_stride;  // cx elements between consecutive transforms
    int inverse;       // C2C only: swap re/im on the way in (A) and out (B)
    int ctb_log2;      // log2 of the stage-2 tile width: scratch / tw4 are tile-major in blocks of 2^ctb_log2 rows
};

#ifdef __CUDACC__

template <typename T>
__device__ __forceinline__ cx<T> ld_plain(const cx<T> *p) {
    using V = typename vec2<T>::type;
    V v = *reinterpret_cast<const V *>(p);
    return mk<T>(v.x, v.y);
}
template <typename T>
__device__ __forceinline__ void st_plain(cx<T> *p, cx<T> v) {
    using V = typename vec2<T>::type;
    V w; w.x = v.x; w.y = v.y;
    *reinterpret_cast<V *>(p) = w;
}

// scratch is rewritten by other SMs between uses: read it through L2 (L1 is not coherent)
template <typename T>
__device__ __forceinline__ cx<T> ld_l2(const cx<T> *p) {
    using V = typename vec2<T>::type;
    V v = __ldcg(reinterpret_cast<const V *>(p));
    return mk<T>(v.x, v.y);
}

template <int FLAVOR>
__host__ __device__ constexpr int tile_width(int n1, int n2) {
    return (FLAVOR == TILE_A_C2C)   ? n2
           : (FLAVOR == TILE_B_C2C) ? n1
           : (FLAVOR == TILE_A_R2C || FLAVOR == TILE_A_C2R) ? n2 / 2
                                    : n1 / 2 + 1;
}

// Scratch (and four-step twiddle table) layout, "tile-major": element (k1, n2) lives at
//     (k1 / CTB) * (CTB * n2) + n2 * CTB + (k1 % CTB)            CTB = lanes of the stage-2 (row) kernel
// i.e. the [n2][CTB] image a stage-2 CTA wants is one contiguous block, and stage 1 can write 128-byte runs
// of consecutive k1.  No shared-memory transposition is needed on either side.
__device__ __forceinline__ long long tm_off(int k1, int n2idx, int n2, int ctb_log2) {
    const int ctb = 1 << ctb_log2;
    return (long long)(k1 >> ctb_log2) * ((long long)ctb * n2) + (long long)n2idx * ctb + (k1 & (ctb - 1));
}


// Where the raw element `idx` of lane `lane` of a first-pass-from-global flavor lives, and what has to be done
// to it after loading (conjugation / zeroing of the packed bin / re-im swap).  Shared by the direct gather and
// by the cp.async prefetch into the staging buffer.
template <int FLAVOR, int L, typename T>
__device__ __forceinline__ const cx<T> *tile_src(const cx<T> *gin, int lane, int idx, int n1, int n2, int ctb) {
    if constexpr (FLAVOR == TILE_A_C2C) return gin + (long long)idx * n2 + lane;
    else if constexpr (FLAVOR == TILE_A_R2C) return gin + (long long)idx * (n2 / 2) + lane;
    else if constexpr (FLAVOR == TILE_B_C2C || FLAVOR == TILE_B_R2C) return gin + tm_off(lane, idx, n2, ctb);
    else {  // TILE_B_C2R: row k1 = lane, element k2 = idx of the Hermitian-extended packed spectrum
        const int k1 = lane, k2 = idx;
        if (k2 < L / 2) return gin + k1 + (long long)n1 * k2;
        if (k1 == 0) return (k2 == L / 2) ? gin : gin + (long long)n1 * (L - k2);
        return gin + (n1 - k1) + (long long)n1 * (L - 1 - k2);
    }
}
template <int FLAVOR, int L, typename T>
__device__ __forceinline__ cx<T> tile_fix(cx<T> x, int lane, int idx, int inverse) {
    if constexpr (FLAVOR == TILE_A_C2C) return inverse ? cswap(x) : x;
    else if constexpr (FLAVOR == TILE_B_C2R) {
        const int k1 = lane, k2 = idx;
        if (k2 < L / 2) {
            if (k1 == 0 && k2 == 0) x.y = (T)0;                      // bin 0 packs (DC, Nyquist)
        } else if (k1 == 0) {
            x = (k2 == L / 2) ? mk<T>(x.y, (T)0) : cconj(x);
        } else {
            x = cconj(x);
        }
        return cswap(x);                                             // inverse transform via the swap trick
    } else return x;
}

__device__ __forceinline__ void cp_async8(void *dst_smem, const void *src_gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)),
                 "l"(src_gmem)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Asynchronously copy the raw first-pass inputs of one tile into the staging buffer `st` ([idx][lane], pitch
// CT+1): LDGSTS, no registers, completes in the background while the previous tile is being transformed.
template <typename Cfg, int FLAVOR, int N1C = 0, int N2C = 0, int CTBLOG = -1>
__device__ __forceinline__ void tile_prefetch(const TileParams<typename Cfg::T> &p, const cx<typename Cfg::T> *gin,
                                              int lane0, cx<typename Cfg::T> *st) {
    using T = typename Cfg::T;
    static_assert(sizeof(cx<T>) == 8, "cp.async staging is implemented for fp32 tiles");
    constexpr int L = Cfg::L, TX = Cfg::TX, CT = Cfg::CT, E = Cfg::E, PITCH = Cfg::PITCH;
    constexpr int R = Cfg::radix(0), NR = L / R, U = E / R;
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
    const int c = tid % CT, t = tid / CT;
    const int n1 = N1C ? N1C : p.n1, n2 = N2C ? N2C : p.n2, ctb = CTBLOG >= 0 ? CTBLOG : p.ctb_log2;
    const int lane = lane0 + c;
    if (lane < tile_width<FLAVOR>(n1, n2)) {
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const int idx = t + TX * u + NR * j;
                cp_async8(st + idx * PITCH + c, tile_src<FLAVOR, L>(gin, lane, idx, n1, n2, ctb));
            }
    }
    cp_async_commit();
}

// One tile of one transform: lanes lane0 .. lane0+CT-1 of the tiled dimension.  gin/gout point at the
// transform (user buffer or scratch, depending on the flavor).  Every thread of the CTA must call this.
// N1C / N2C / CTBLOG: compile-time four-step dimensions (0 / -1 = take them from `p` at run time).  With them
// fixed every global address is base + immediate, which roughly halves the instruction count of a tile.
// First-pass inputs of one tile into a caller-owned register array (same element order as tile_body's gather).
// Issued one tile ahead, the loads stay in flight while the previous tile is transformed (register double buffer).
template <typename Cfg, int FLAVOR, int N1C = 0, int N2C = 0, int CTBLOG = -1>
__device__ __forceinline__ void tile_load(const TileParams<typename Cfg::T> &p, const cx<typename Cfg::T> *gin, int lane0,
                                          cx<typename Cfg::T> (&v)[Cfg::E]) {
    using T = typename Cfg::T;
    constexpr int L = Cfg::L, TX = Cfg::TX, CT = Cfg::CT, E = Cfg::E;
    constexpr int R = Cfg::radix(0), NR = L / R, U = E / R;
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
    const int c = tid % CT, t = tid / CT;
    const int n1 = N1C ? N1C : p.n1, n2 = N2C ? N2C : p.n2, ctb = CTBLOG >= 0 ? CTBLOG : p.ctb_log2;
    const int lane = lane0 + c;
    if (lane < tile_width<FLAVOR>(n1, n2)) {
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const int idx = t + TX * u + NR * j;
                const cx<T> *src = tile_src<FLAVOR, L>(gin, lane, idx, n1, n2, ctb);
                cx<T> x;
                if constexpr (FLAVOR == TILE_B_C2C || FLAVOR == TILE_B_R2C) x = ld_l2(src);
                else x = ld_stream(src);
                v[u * R + j] = tile_fix<FLAVOR, L>(x, lane, idx, p.inverse);
            }
    }
}

struct NoHook {
    __device__ __forceinline__ void operator()() const {}
};
// STAGED: the first-pass inputs were prefetched into `st` by tile_prefetch (the caller has waited for them);
// `hook` runs right after the first block-wide barrier, i.e. as soon as `st` may be overwritten again.
template <typename Cfg, int FLAVOR, int N1C = 0, int N2C = 0, int CTBLOG = -1, bool STAGED = false, typename Hook = NoHook,
          bool PRE = false>
__device__ __forceinline__ void tile_body(const TileParams<typename Cfg::T> &p, const cx<typename Cfg::T> *gin,
                                          cx<typename Cfg::T> *gout, int lane0, cx<typename Cfg::T> *sm,
                                          const cx<typename Cfg::T> *st = nullptr, Hook hook = Hook(),
                                          const cx<typename Cfg::T> *pre = nullptr /* E preloaded inputs, or null */) {
    using T = typename Cfg::T;
    constexpr int L = Cfg::L, TX = Cfg::TX, CT = Cfg::CT, E = Cfg::E, NP = Cfg::NP, PITCH = Cfg::PITCH;
    constexpr int THREADS = Cfg::THREADS;
    constexpr bool kFirstFromSmem = (FLAVOR == TILE_A_C2R);
    constexpr bool kLastToSmem = (FLAVOR == TILE_A_R2C);
    // stage-1 C2C: the last pass runs with lanes along the butterfly index so its stores are k1-contiguous
    constexpr bool kTransposedLast = (FLAVOR == TILE_A_C2C);
    static_assert(!kTransposedLast || NP >= 2, "transposed last pass needs a shared-memory exchange before it");
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
    const int c = tid % CT, t = tid / CT;      // lanes along the tiled dimension (global loads coalesce over c)
    const int c2 = tid / TX, t2 = tid % TX;    // transposed mapping: lanes along the butterfly index
    const int n1 = N1C ? N1C : p.n1, n2 = N2C ? N2C : p.n2, ctb = CTBLOG >= 0 ? CTBLOG : p.ctb_log2;
    const int width = tile_width<FLAVOR>(n1, n2);
    {
        const int lane = lane0 + c;  // column (A flavors) or row k1 (B flavors)
        const bool live = lane < width;
        cx<T> v[E];

        if constexpr (FLAVOR == TILE_A_C2R) {
            // G[k1] = U[k1][2c'] + i U[k1][2c'+1], Hermitian-extended to k1 > N1/2; scratch holds swap(U).
            // Lanes run along k1 so the tile-major scratch is read in 128-byte runs.
            constexpr int KH = L / 2;
            for (int e = tid; e < KH * CT; e += THREADS) {
                const int k1 = e % KH, cc = e / KH;
                if (lane0 + cc < width) {
                    const int col = 2 * (lane0 + cc);
                    const cx<T> a = cswap(ld_l2(gin + tm_off(k1, col, n2, ctb)));
                    const cx<T> b = cswap(ld_l2(gin + tm_off(k1, col + 1, n2, ctb)));
                    sm[k1 * PITCH + cc] = cswap(mk<T>(a.x - b.y, a.y + b.x));
                    if (k1 > 0) sm[(L - k1) * PITCH + cc] = cswap(mk<T>(a.x + b.y, b.x - a.y));
                }
            }
            if (tid < CT && lane0 + tid < width) {  // k1 = N1/2 (self-conjugate row)
                const int col = 2 * (lane0 + tid);
                const cx<T> a = cswap(ld_l2(gin + tm_off(KH, col, n2, ctb)));
                const cx<T> b = cswap(ld_l2(gin + tm_off(KH, col + 1, n2, ctb)));
                sm[KH * PITCH + tid] = cswap(mk<T>(a.x - b.y, a.y + b.x));
            }
            __syncthreads();
        }

        sfor<0, NP>([&](auto pc) {
            constexpr int ps = decltype(pc)::value;
            constexpr int R = Cfg::radix(ps), P = Cfg::prod(ps), MN = Cfg::mnext(ps), NR = L / R, U = E / R;
            constexpr bool first = (ps == 0), last = (ps == NP - 1);
            constexpr bool tr = last && kTransposedLast;  // this pass uses the transposed thread mapping
            const int tt = tr ? t2 : t, cc = tr ? c2 : c;
            // ---- gather
            if constexpr (first && !kFirstFromSmem) {
                if constexpr (PRE) {  // inputs were loaded one tile ahead by tile_load (register double buffer)
#pragma unroll
                    for (int e = 0; e < E; ++e) v[e] = pre[e];
                } else if (live) {
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int j = 0; j < R; ++j) {
                            const int idx = t + TX * u + NR * j;
                            cx<T> x;
                            if constexpr (STAGED) {
                                x = st[idx * (SSFFT_FOURSTEP_TMA ? CT : PITCH) + c];
                            } else {
                                const cx<T> *src = tile_src<FLAVOR, L>(gin, lane, idx, n1, n2, ctb);
                                if constexpr (FLAVOR == TILE_B_C2C || FLAVOR == TILE_B_R2C) x = ld_l2(src);
                                else x = ld_stream(src);
                            }
                            x = tile_fix<FLAVOR, L>(x, lane, idx, p.inverse);
                            v[u * R + j] = x;
                        }
                }
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int j = 0; j < R; ++j) v[u * R + j] = sm[(tt + TX * u + NR * j) * PITCH + cc];
                __syncthreads();
            }
            // ---- butterflies + inter-pass twiddles (identical for every lane of a tile)
#pragma unroll
            for (int u = 0; u < U; ++u) {
                cx<T> w[R];
#pragma unroll
                for (int j = 0; j < R; ++j) w[j] = v[u * R + j];
                Dft<R>::run(w);
                if constexpr (!last) {
                    const int b = tt + TX * u;
                    const int mp = b / P;
                    const cx<T> *twp = p.tw + Cfg::tw_off(ps) + mp;
#pragma unroll
                    for (int r = 1; r < R; ++r) w[r] = cmul(w[r], ld_table(twp + (r - 1) * MN));
                }
#pragma unroll
                for (int j = 0; j < R; ++j) v[u * R + j] = w[j];
            }
            // ---- scatter
            if constexpr (!last) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int b = tt + TX * u;
                    const int mp = b / P, racc = b - mp * P;
                    const int o = racc + P * R * mp;
#pragma unroll
                    for (int r = 0; r < R; ++r) sm[(o + P * r) * PITCH + cc] = v[u * R + r];
                }
                __syncthreads();
                if constexpr (first) hook();  // every thread has consumed the staging buffer
            } else if constexpr (kLastToSmem) {
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int r = 0; r < R; ++r) sm[(t + TX * u + P * r) * PITCH + c] = v[u * R + r];
                __syncthreads();
            } else {
                const int lane_s = lane0 + cc;  // the lane this thread stores for (differs from `lane` when transposed)
                if (lane_s < width) {
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const int k = tt + TX * u + P * r;  // natural-order output index of this stage
                            cx<T> x = v[u * R + r];
                            if constexpr (FLAVOR == TILE_A_C2C) {        // column n2 = lane_s, row k1 = k
                                const long long o = tm_off(k, lane_s, n2, ctb);
                                st_plain(gout + o, cmul(x, ld_table(p.tw4 + o)));  // scratch: keep it in L2
                            } else if constexpr (FLAVOR == TILE_B_C2R) {  // row k1 = lane_s, column n2 = k
                                const long long o = tm_off(lane_s, k, n2, ctb);
                                st_plain(gout + o, cmul(x, ld_table(p.tw4 + o)));  // swapped domain: plain W
                            } else if constexpr (FLAVOR == TILE_B_C2C) {
                                if (p.inverse) x = cswap(x);
                                st_stream(gout + lane_s + (long long)n1 * k, x);
                            } else if constexpr (FLAVOR == TILE_A_C2R) {
                                st_stream(gout + (long long)k * (n2 / 2) + lane_s, cswap(x));
                            } else {  // TILE_B_R2C: row k1 = lane_s, bin k2 = k
                                const int k1 = lane_s, k2 = k;
                                if (k2 < L / 2) {
                                    if (k1 == 0 && k2 == 0) reinterpret_cast<T *>(gout)[0] = x.x;  // DC
                                    else st_stream(gout + k1 + (long long)n1 * k2, x);
                                } else if (k1 == 0) {
                                    if (k2 == L / 2) reinterpret_cast<T *>(gout)[1] = x.x;  // Nyquist
                                } else if (k1 < n1 / 2) {
                                    st_stream(gout + (n1 - k1) + (long long)n1 * (L - 1 - k2), cconj(x));
                                }
                            }
                        }
                }
            }
        });

        if constexpr (FLAVOR == TILE_A_R2C) {
            // separate the two real columns packed in Z, twiddle, store rows k1 <= N1/2 of the scratch;
            // lanes run along k1 so each store instruction writes 128-byte runs of the tile-major scratch
            constexpr int KH = L / 2;
            const T half = (T)0.5;
            auto emit = [&](int k1, int cc) {
                const cx<T> z = sm[k1 * PITCH + cc], zc = sm[((L - k1) & (L - 1)) * PITCH + cc];
                const cx<T> xe = mk<T>((z.x + zc.x) * half, (z.y - zc.y) * half);
                const cx<T> xo = mk<T>((z.y + zc.y) * half, (zc.x - z.x) * half);
                const int col = 2 * (lane0 + cc);
                const long long o0 = tm_off(k1, col, n2, ctb), o1 = tm_off(k1, col + 1, n2, ctb);
                st_plain(gout + o0, cmul(xe, ld_table(p.tw4 + o0)));
                st_plain(gout + o1, cmul(xo, ld_table(p.tw4 + o1)));
            };
            for (int e = tid; e < KH * CT; e += THREADS) {
                const int k1 = e % KH, cc = e / KH;
                if (lane0 + cc < width) emit(k1, cc);
            }
            if (tid < CT && lane0 + tid < width) emit(KH, tid);
            __syncthreads();
        }
        if constexpr (NP == 1 && FLAVOR != TILE_A_R2C) __syncthreads();
    }
}

// stand-alone launch of one stage over a whole batch (two launches per chunk; used for fp64 and as fallback)
template <typename Cfg, int FLAVOR>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB) tile_fft_kernel(TileParams<typename Cfg::T> p) {
    using T = typename Cfg::T;
    SSFFT_DYNAMIC_SMEM(ssfft_smem);
    cx<T> *sm = reinterpret_cast<cx<T> *>(ssfft_smem);
    const int width = tile_width<FLAVOR>(p.n1, p.n2);
    const int tiles = (width + Cfg::CT - 1) / Cfg::CT;
    const long long items = p.batch * tiles;
    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
        const long long B = item / tiles;
        const int lane0 = (int)(item - B * tiles) * Cfg::CT;
        tile_body<Cfg, FLAVOR>(p, p.in + B * p.in_stride, p.out + B * p.out_stride, lane0, sm);
    }
}

// ---------------------------------------------------------------------------------------------
// Both stages in ONE persistent launch: a thread-block cluster owns a transform.  Stage 1 tiles are
// spread over the cluster's CTAs, a hardware cluster barrier (release/acquire) publishes the scratch,
// stage 2 tiles follow.  The scratch (two slots per cluster, alternating) is a few MB in total and
// clusters are placed on one die, so the intermediate never leaves that die's L2.
//   KIND 0: C2C  (A_C2C on CfgA, then B_C2C on CfgB)
//   KIND 1: R2C  (A_R2C on CfgA, then B_R2C on CfgB)
//   KIND 2: C2R  (B_C2R on CfgB, then A_C2R on CfgA)
// ---------------------------------------------------------------------------------------------
template <typename T>
struct FourStepParams {
    const cx<T> *in;
    cx<T> *out;
    cx<T> *scratch;          // 2 * clusters * scratch_per elements
    const cx<T> *tw_a, *tw_b, *tw4;
    int n1, n2;
    long long batch, user_stride, scratch_per;
    int inverse, ctb_log2;
    int discard;  // drop consumed scratch lines from L2 (discard.global.L2)
    // A GROUP of `group_clusters` clusters shares one transform (tiles are spread over all its CTAs, the two stages
    // are separated by a software barrier on group_ctr[group]).  Large transforms use big groups so that the scratch
    // of all transforms in flight (2 slots per group) stays inside the L2 instead of spilling to HBM.
    int group_clusters;
    unsigned *group_ctr;  // one zero-initialised counter per group
    int cluster_size;     // CTAs per hardware cluster of this launch (host side only)
    int use_tma;          // the tensor map passed next to these parameters is valid (set by the launcher)
};

#ifdef SSFFT_EMUL  // CPU execution of the kernels (tests/host/simt/simt_emul.h): hooks instead of PTX
inline unsigned cluster_ctarank() { return simt::cluster_ctarank(); }
inline unsigned cluster_nctarank() { return simt::cluster_nctarank(); }
inline void cluster_barrier() { simt::cluster_barrier(); }
inline unsigned ld_acquire_gpu(const unsigned *p) { simt::spin_yield(); return *p; }
inline void discard_l2_line(const void *a) { memset(const_cast<void *>(a), 0xff, 128); }  // poison: nobody may read it again
#else
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cluster_nctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void discard_l2_line(const void *a) { asm volatile("discard.global.L2 [%0], 128;" ::"l"(a) : "memory"); }
#endif

// Barrier over the CTAs of a group (all co-resident: the launch is cooperative / sized to residency).  The counter
// only grows; `target` is the value it reaches when every CTA of the group has arrived for this barrier.
// Bounded spin: a scheduling surprise becomes a trap (launch error), never a hung GPU.
__device__ __forceinline__ void group_barrier(unsigned *ctr, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        __threadfence();  // this CTA's scratch stores are visible device-wide before it signals
        atomicAdd(ctr, 1u);
        const long long t0 = clock64();
        for (;;) {
            const unsigned v = ld_acquire_gpu(ctr);
            if (v >= target) break;
            __nanosleep(40);
            if (clock64() - t0 > 4000000000LL) __trap();
        }
    }
    __syncthreads();
}

// Drop consumed scratch lines from L2 WITHOUT writing them back (the slot is fully rewritten before it is
// read again), so the intermediate costs no HBM write traffic.  `first_line` .. +count lines of 128 bytes.
template <typename T>
__device__ __forceinline__ void discard_lines(const cx<T> *base, long long first_line, int count, int stride_lines,
                                              int groups, int tid, int nthreads) {
    // `groups` runs of `count` consecutive lines, run g starting at first_line + g * stride_lines
    for (int i = tid; i < count * groups; i += nthreads) {
        const int g = i / count, l = i - g * count;
        const char *a = reinterpret_cast<const char *>(base) + (first_line + (long long)g * stride_lines + l) * 128;
        discard_l2_line(a);
    }
}

// two tile buffers (exchange + cp.async staging) per CTA, if that still fits the CTAs/SM of the launch bounds.
// MEASURED SLOWER on B200 (65536 C2C: 50 % -> 36 % of roofline; real 65536: 34 % -> 28 %): the 8-byte LDGSTS
// copies and the extra wait + barrier per tile cost more than the latency they hide at 3 CTAs/SM.  Kept behind
// -DSSFFT_FOURSTEP_STAGED=1 for future work (a TMA bulk copy of the contiguous stage-2 block is the next try).
#ifndef SSFFT_FOURSTEP_STAGED
#define SSFFT_FOURSTEP_STAGED 0
#endif
template <typename CfgA, typename CfgB>
__host__ __device__ constexpr bool fourstep_staged() {
    return (SSFFT_FOURSTEP_STAGED || SSFFT_FOURSTEP_TMA) && sizeof(typename CfgA::T) == 4 &&
           2 * (CfgA::smem_bytes > CfgB::smem_bytes ? CfgA::smem_bytes : CfgB::smem_bytes) *
                   (size_t)(CfgA::MINB < CfgB::MINB ? CfgA::MINB : CfgB::MINB) <= 220 * 1024;
}
template <typename CfgA, typename CfgB>
__host__ __device__ constexpr size_t fourstep_smem_bytes() {
    return (CfgA::smem_bytes > CfgB::smem_bytes ? CfgA::smem_bytes : CfgB::smem_bytes) * (fourstep_staged<CfgA, CfgB>() ? 2 : 1);
}

// L2 prefetch hints for the first stage (the one that reads the user buffer from HBM): while a CTA transforms one stage-1 tile it asks for the lines
// of the NEXT stage-1 tile it will work on (same transform, or its first tile of the next transform) with
// prefetch.global.L2 -- one instruction per 128-byte line, no registers, no shared memory, no wait: the loads of the next
// tile then hit L2 instead of HBM.  Motivation: about a sixth of the stall samples of the 65536 kernel sit on the first use of
// the tile loads (profiles/c2c65536_fourstep_cluster_r01b_stalls.txt) and the two schemes that HOLD the prefetched data
// (cp.async staging, register double buffer) lost more occupancy than they hid.  Runs on the CPU emulation
// (tests/test_tiled_emul.py builds it); NOT YET MEASURED on the GPU, hence off by default: NVFLAGS += -DSSFFT_FOURSTEP_L2PF=1.
#ifndef SSFFT_FOURSTEP_L2PF
#define SSFFT_FOURSTEP_L2PF 0
#endif
// every 128-byte line the first pass of stage-1 tile `lane0` will read from the user buffer.  TILE_A_C2C / TILE_A_R2C: row
// idx of the tile is CT consecutive elements.  TILE_B_C2R (packed half spectrum, Hermitian-extended on the fly): the CT
// rows k1 of the tile read CT consecutive elements per k2, ascending for k2 < L/2 and descending (mirrored, offset by
// one element) for k2 >= L/2 -- the addresses of the first, middle and last live lane cover the one or two lines.
template <typename Cfg, int FLAVOR, int N1C, int N2C>
__device__ __forceinline__ void tile_l2_prefetch(const cx<typename Cfg::T> *gin, int lane0, int tid) {
    using T = typename Cfg::T;
    static_assert(FLAVOR == TILE_A_C2C || FLAVOR == TILE_A_R2C || FLAVOR == TILE_B_C2R, "stage-1 tiles that read the user buffer");
    if constexpr (FLAVOR == TILE_B_C2R) {
        constexpr int width = tile_width<FLAVOR>(N1C, N2C);
        const int last = (lane0 + Cfg::CT - 1 < width ? lane0 + Cfg::CT - 1 : width - 1);
        for (int i = tid; i < Cfg::L * 3; i += Cfg::THREADS) {
            const int idx = i / 3, part = i - idx * 3;
            const int lane = part == 0 ? lane0 : part == 1 ? (lane0 + last) / 2 : last;
            prefetch_l2(tile_src<FLAVOR, Cfg::L>(gin, lane, idx, N1C, N2C, 0));
        }
    } else {
        constexpr int kPerLine = 128 / (int)sizeof(cx<T>);
        constexpr int kLines = (Cfg::CT + kPerLine - 1) / kPerLine;  // lines per row of the tile
        for (int i = tid; i < Cfg::L * kLines; i += Cfg::THREADS) {
            const int idx = i / kLines, part = i - idx * kLines;
            prefetch_l2(tile_src<FLAVOR, Cfg::L>(gin, lane0 + part * kPerLine, idx, N1C, N2C, 0));
        }
    }
}

#ifndef SSFFT_FOURSTEP_REGPREFETCH
#define SSFFT_FOURSTEP_REGPREFETCH 0  // measured slower (65536 C2C: 50.6 % -> 43.9 %): 3 CTAs/SM beat 2 CTAs/SM + register prefetch
#endif
template <typename CfgA, typename CfgB>
__host__ __device__ constexpr bool fourstep_regprefetch() {
    return SSFFT_FOURSTEP_REGPREFETCH && !fourstep_staged<CfgA, CfgB>() && sizeof(typename CfgA::T) == 4 && CfgA::E <= 16 &&
           CfgB::E <= 16;
}
template <typename CfgA, typename CfgB>
__host__ __device__ constexpr int fourstep_minb() {
    return fourstep_regprefetch<CfgA, CfgB>() ? 2 : (CfgA::MINB < CfgB::MINB ? CfgA::MINB : CfgB::MINB);
}

template <typename CfgA, typename CfgB, int KIND>
__global__ void __launch_bounds__(CfgA::THREADS, (fourstep_minb<CfgA, CfgB>()))
fourstep_cluster_kernel(FourStepParams<typename CfgA::T> q, const __grid_constant__ CUtensorMap tmap) {
    using T = typename CfgA::T;
    static_assert(CfgA::THREADS == CfgB::THREADS, "both stages must use the same CTA size");
    SSFFT_DYNAMIC_SMEM(ssfft_smem);
    __shared__ __align__(8) unsigned long long stage_bar;
    cx<T> *sm = reinterpret_cast<cx<T> *>(ssfft_smem);
    const int crank = (int)cluster_ctarank(), csize0 = (int)cluster_nctarank();
    const long long cid0 = blockIdx.x / csize0, nclusters0 = gridDim.x / csize0;
    // group view: `rank` of `csize` CTAs work on a transform, `cid` of `nclusters` groups (a group is one cluster when
    // group_clusters == 1)
    const int G = q.group_clusters > 1 ? q.group_clusters : 1;
    const int rank = (int)(cid0 % G) * csize0 + crank, csize = G * csize0;
    const long long cid = cid0 / G, nclusters = nclusters0 / G;
    unsigned bar_target = 0;
    constexpr int F1 = KIND == 0 ? TILE_A_C2C : KIND == 1 ? TILE_A_R2C : TILE_B_C2R;
    constexpr int F2 = KIND == 0 ? TILE_B_C2C : KIND == 1 ? TILE_B_R2C : TILE_A_C2R;
    using Cfg1 = typename std::conditional<KIND == 2, CfgB, CfgA>::type;
    using Cfg2 = typename std::conditional<KIND == 2, CfgA, CfgB>::type;
    TileParams<T> p1, p2;
    p1.in = nullptr; p1.out = nullptr; p1.tw = KIND == 2 ? q.tw_b : q.tw_a; p1.tw4 = q.tw4;
    p1.n1 = q.n1; p1.n2 = q.n2; p1.batch = 1; p1.in_stride = 0; p1.out_stride = 0; p1.inverse = q.inverse;
    p1.ctb_log2 = q.ctb_log2;
    p2 = p1;
    p2.tw = KIND == 2 ? q.tw_a : q.tw_b;
    constexpr int kCtbLog = CfgB::CT == 32 ? 5 : CfgB::CT == 16 ? 4 : CfgB::CT == 8 ? 3 : CfgB::CT == 4 ? 2 : -1;
    static_assert(kCtbLog >= 0, "row-stage tile width must be 4, 8, 16 or 32");
    constexpr int tiles1 = (tile_width<F1>(CfgA::L, CfgB::L) + Cfg1::CT - 1) / Cfg1::CT;
    constexpr int tiles2 = (tile_width<F2>(CfgA::L, CfgB::L) + Cfg2::CT - 1) / Cfg2::CT;
    // cp.async double buffering: the next tile's inputs stream into `st` while the current tile is transformed.
    // Enabled when two tile buffers per CTA still leave room for the CTAs/SM the launch bounds ask for.
    constexpr size_t kTileBytes = CfgA::smem_bytes > CfgB::smem_bytes ? CfgA::smem_bytes : CfgB::smem_bytes;
    constexpr bool kTma = SSFFT_FOURSTEP_TMA != 0;
    // (C2R gathers a Hermitian-extended spectrum and builds its second stage in shared memory: not staged by TMA)
    constexpr bool kStage = fourstep_staged<CfgA, CfgB>() && (!kTma || KIND != 2);
    constexpr bool kStage2 = kStage && (F2 != TILE_A_C2R);  // A_C2R builds its input in shared memory itself
    cx<T> *st = reinterpret_cast<cx<T> *>(ssfft_smem + kTileBytes);
    bool s1_ready = false;  // the first stage-1 tile of the coming transform is already in flight
    unsigned st_par = 0;
    const bool tid0 = (threadIdx.x == 0 && threadIdx.y == 0);
    if constexpr (kStage && kTma) {
        if (tid0) mbar_init(&stage_bar, 1);
        __syncthreads();
    }
    // request the first-pass inputs of a tile into `st`
    auto issue_s1 = [&](long long Bx, int tile) {  // stage-1 tile `tile` of transform Bx (user buffer)
        if constexpr (kTma) {
            if (tid0) {
                constexpr unsigned bytes = (unsigned)(Cfg1::L * Cfg1::CT * sizeof(cx<T>));
                mbar_expect_tx(&stage_bar, bytes);
                asm volatile(
                    "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                        smem_u32(st)),
                    "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(tile * Cfg1::CT), "r"(0), "r"((int)Bx), "r"(smem_u32(&stage_bar))
                    : "memory");
            }
        } else if constexpr (kStage) {
            tile_prefetch<Cfg1, F1, CfgA::L, CfgB::L, kCtbLog>(p1, q.in + Bx * q.user_stride, tile * Cfg1::CT, st);
        }
    };
    auto issue_s2 = [&](const cx<T> *scr_, int tile) {  // stage-2 tile: one contiguous tile-major scratch block
        if constexpr (kTma) {
            if (tid0) {
                constexpr unsigned bytes = (unsigned)(Cfg2::L * Cfg2::CT * sizeof(cx<T>));
                mbar_expect_tx(&stage_bar, bytes);
                asm volatile("fence.proxy.async;" ::: "memory");  // peers' generic-proxy scratch stores -> async-proxy read
                bulk_g2s(st, scr_ + (long long)tile * Cfg2::CT * Cfg2::L, bytes, &stage_bar);
            }
        } else if constexpr (kStage) {
            tile_prefetch<Cfg2, F2, CfgA::L, CfgB::L, kCtbLog>(p2, scr_, tile * Cfg2::CT, st);
        }
    };
    auto stage_wait = [&]() {
        if constexpr (kTma) {
            mbar_wait(&stage_bar, st_par);
            st_par ^= 1u;
        } else {
            cp_async_wait_all();
            __syncthreads();
        }
    };
    // register double buffering (the adopted overlap scheme): next tile's inputs are loaded into a second register
    // set while the current tile is transformed; 2 CTAs/SM at <= 128 registers instead of 3 at 80.
    constexpr bool kRegPf = fourstep_regprefetch<CfgA, CfgB>();
    constexpr bool kRegPf2 = (F2 != TILE_A_C2R);
    [[maybe_unused]] cx<T> vn1[Cfg1::E];
    bool have1 = false;
    int slot = 0;
    for (long long B = cid; B < q.batch; B += nclusters, slot ^= 1) {
        cx<T> *scr = q.scratch + (cid * 2 + slot) * q.scratch_per;
        const cx<T> *uin = q.in + B * q.user_stride;
        cx<T> *uout = q.out + B * q.user_stride;
        const long long Bn = B + nclusters;
        const cx<T> *uin_next = q.in + Bn * q.user_stride;
        auto prefetch_next_s1 = [&]() {
            if constexpr (kStage) {
                if (Bn < q.batch && rank < tiles1) {
                    issue_s1(Bn, rank);
                    s1_ready = true;
                }
            }
        };
        // ---- stage 1
        for (int tile = rank; tile < tiles1; tile += csize) {
            if constexpr (kRegPf) {
                // register double buffer: this tile's inputs were requested one tile ago; request the next one now
                if (!have1) tile_load<Cfg1, F1, CfgA::L, CfgB::L, kCtbLog>(p1, uin, tile * Cfg1::CT, vn1);
                cx<T> vc[Cfg1::E];
#pragma unroll
                for (int e = 0; e < Cfg1::E; ++e) vc[e] = vn1[e];
                const int next = tile + csize;
                have1 = next < tiles1;
                if (have1) tile_load<Cfg1, F1, CfgA::L, CfgB::L, kCtbLog>(p1, uin, next * Cfg1::CT, vn1);
                tile_body<Cfg1, F1, CfgA::L, CfgB::L, kCtbLog, false, NoHook, true>(p1, uin, scr, tile * Cfg1::CT, sm, nullptr,
                                                                                NoHook(), vc);
            } else {
                if constexpr (kStage) {
                    if (!s1_ready) issue_s1(B, tile);
                    s1_ready = false;
                    stage_wait();
                }
                const int next = tile + csize;
                auto hook = [&]() {
                    if constexpr (kStage) {
                        if (next < tiles1) {
                            issue_s1(B, next);
                            s1_ready = true;
                        }
                    }
                };
                if constexpr (SSFFT_FOURSTEP_L2PF != 0) {
                    // hint the next stage-1 tile of this CTA into L2 while this one is transformed
                    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
                    if (next < tiles1) tile_l2_prefetch<Cfg1, F1, CfgA::L, CfgB::L>(uin, next * Cfg1::CT, tid);
                    else if (Bn < q.batch && rank < tiles1) tile_l2_prefetch<Cfg1, F1, CfgA::L, CfgB::L>(uin_next, rank * Cfg1::CT, tid);
                }
                tile_body<Cfg1, F1, CfgA::L, CfgB::L, kCtbLog, kStage>(p1, uin, scr, tile * Cfg1::CT, sm, st, hook);
            }
        }
        // stage-1 stores of every CTA working on this transform are visible (scratch is read through L2: ld.global.cg)
        if (G == 1) {
            cluster_barrier();
        } else {
            bar_target += (unsigned)csize;
            group_barrier(q.group_ctr + cid, bar_target);
        }
        // ---- stage 2 (start rotates so the CTA that gets an extra, ragged tile changes between transforms)
        bool s2_ready = false;
        if constexpr (kStage && !kStage2) prefetch_next_s1();  // stage 2 does not use `st`: overlap all of it
        [[maybe_unused]] cx<T> vn2[Cfg2::E];
        bool have2 = false;
        for (int i = rank; i < tiles2; i += csize) {
            const int tile = (int)((i + B) % tiles2);
            const int inext = i + csize;
            if constexpr (kRegPf && kRegPf2) {
                if (!have2) tile_load<Cfg2, F2, CfgA::L, CfgB::L, kCtbLog>(p2, scr, tile * Cfg2::CT, vn2);
                cx<T> vc[Cfg2::E];
#pragma unroll
                for (int e = 0; e < Cfg2::E; ++e) vc[e] = vn2[e];
                have2 = inext < tiles2;
                if (have2) {
                    tile_load<Cfg2, F2, CfgA::L, CfgB::L, kCtbLog>(p2, scr, (int)((inext + B) % tiles2) * Cfg2::CT, vn2);
                } else if (Bn < q.batch && rank < tiles1) {  // last stage-2 tile: request the next transform's first tile
                    tile_load<Cfg1, F1, CfgA::L, CfgB::L, kCtbLog>(p1, uin_next, rank * Cfg1::CT, vn1);
                    have1 = true;
                }
                tile_body<Cfg2, F2, CfgA::L, CfgB::L, kCtbLog, false, NoHook, true>(p2, scr, uout, tile * Cfg2::CT, sm, nullptr,
                                                                                NoHook(), vc);
            } else {
                if constexpr (kRegPf) {  // stage 2 builds its input itself (A_C2R): still prefetch the next stage-1 tile
                    if (inext >= tiles2 && Bn < q.batch && rank < tiles1) {
                        tile_load<Cfg1, F1, CfgA::L, CfgB::L, kCtbLog>(p1, uin_next, rank * Cfg1::CT, vn1);
                        have1 = true;
                    }
                }
                if constexpr (kStage2) {
                    if (!s2_ready) issue_s2(scr, tile);
                    s2_ready = false;
                    stage_wait();
                }
                auto hook = [&]() {
                    if constexpr (kStage2) {
                        if (inext < tiles2) {
                            issue_s2(scr, (int)((inext + B) % tiles2));
                            s2_ready = true;
                        } else {
                            prefetch_next_s1();
                        }
                    }
                };
                tile_body<Cfg2, F2, CfgA::L, CfgB::L, kCtbLog, kStage2>(p2, scr, uout, tile * Cfg2::CT, sm, st, hook);
            }
            if (q.discard) {
                // every stage-2 tile consumes a disjoint set of 128-byte scratch lines (tile-major layout)
                constexpr int kLinesPerRow = CfgB::CT * (int)sizeof(cx<T>) / 128;  // lines per (block, n2)
                const int tid = threadIdx.x + threadIdx.y * blockDim.x;
                if constexpr (kLinesPerRow >= 1) {
                    if constexpr (KIND == 2) {  // A_C2R: columns 2*lane0 .. +2*CT of every row block
                        constexpr int blocks = (CfgA::L / 2 + CfgB::CT) / CfgB::CT;
                        discard_lines(scr, (long long)tile * 2 * CfgA::CT * kLinesPerRow, 2 * CfgA::CT * kLinesPerRow,
                                      CfgB::L * kLinesPerRow, blocks, tid, CfgA::THREADS);
                    } else {  // B flavors: one contiguous block of CT rows
                        discard_lines(scr, (long long)tile * CfgB::L * kLinesPerRow, CfgB::L * kLinesPerRow, 0, 1, tid,
                                      CfgA::THREADS);
                    }
                }
            }
        }
    }
    if constexpr (kStage && !kTma) cp_async_wait_all();
}

#endif  // __CUDACC__

// ---------------------------------------------------------------------------------------------
// registry of tile kernels: one entry per (precision, L), six flavor launchers each
// ---------------------------------------------------------------------------------------------
struct TileEntry {
    int prec, len;
    const char *name;
    int tw_total, np, radix[3];
    int ct;  // lanes per tile
    int (*launch[6])(const void *params, cudaStream_t s);  // params: TileParams<T>
};
const std::vector<TileEntry> &tile_registry();

// cluster four-step kernels: one entry per (precision, n1, n2), launchers for C2C / R2C / C2R
struct FourStepEntry {
    int prec, n1, n2;
    const char *name;
    int threads;
    size_t smem_bytes;
    // returns 0 on success; *clusters_out (may be null) reports how many clusters the launch used
    int (*launch[3])(const void *params, int max_clusters, cudaStream_t s);
    int (*max_clusters[3])(int cluster_size);  // co-resident clusters on the current device
    int tiles[3][2];  // stage-1 / stage-2 tiles per transform, per kind
};
const std::vector<FourStepEntry> &fourstep_registry();
int fourstep_cluster_size();  // env SSFFT_CLUSTER (default 4)

template <typename T>
inline int find_fourstep(size_t n1, size_t n2) {
    const int prec = sizeof(T) == 4 ? 0 : 1;
    const auto &reg = fourstep_registry();
    for (size_t i = 0; i < reg.size(); ++i)
        if (reg[i].prec == prec && (size_t)reg[i].n1 == n1 && (size_t)reg[i].n2 == n2) return (int)i;
    return -1;
}

template <typename T>
inline int find_tile(size_t len) {
    const int prec = sizeof(T) == 4 ? 0 : 1;
    const auto &reg = tile_registry();
    for (size_t i = 0; i < reg.size(); ++i)
        if (reg[i].prec == prec && (size_t)reg[i].len == len) return (int)i;
    return -1;
}

// Four-step twiddles W_total^(k1 * c), rows k1 < n1 (real transforms: k1 <= n1/2), columns c < n2, in the tile-major
// layout of the scratch (see tm_off): blocks of `ctb` rows, [c][k1 % ctb] inside a block; rows padded to whole blocks.
// Returns the padded row count (scratch elements per transform = rows * n2).
template <typename T>
inline size_t fill_fourstep_twiddles(std::vector<T> &h, size_t total, size_t n1, size_t n2, size_t ctb, bool real) {
    const size_t rows_live = real ? n1 / 2 + 1 : n1;
    const size_t rows = (rows_live + ctb - 1) / ctb * ctb;
    h.assign(2 * rows * n2, (T)0);
    for (size_t k1 = 0; k1 < rows_live; ++k1)
        for (size_t c = 0; c < n2; ++c) {
            unsigned long long q = (unsigned long long)k1 * c % total;
            long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)q / (long double)total;
            const size_t o = (k1 / ctb) * (ctb * n2) + c * ctb + (k1 % ctb);
            h[2 * o] = (T)cosl(a);
            h[2 * o + 1] = (T)(-sinl(a));
        }
    return rows;
}

template <typename T>
inline int build_tile_twiddles(int id, void **d_out) {
    const TileEntry &e = tile_registry()[id];
    std::vector<T> h(2 * (size_t)(e.tw_total > 0 ? e.tw_total : 1));
    size_t o = 0;
    int P = 1;
    for (int p = 0; p + 1 < e.np; ++p) {
        const int R = e.radix[p], MN = e.len / (P * R);
        for (int r = 1; r < R; ++r)
            for (int m = 0; m < MN; ++m) {
                unsigned long long q = (unsigned long long)P * m * r % (unsigned long long)e.len;
                long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)q / (long double)e.len;
                h[2 * o] = (T)cosl(a);
                h[2 * o + 1] = (T)(-sinl(a));
                ++o;
            }
        P *= R;
    }
    if (cudaMalloc(d_out, h.size() * sizeof(T)) != cudaSuccess) return 5;
    if (cudaMemcpy(*d_out, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) return 2;
    return 0;
}

}  // namespace ssfft
