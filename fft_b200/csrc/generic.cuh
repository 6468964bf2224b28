// generic.cuh -- the any-size batched FFT kernel: a Stockham pass interpreter in shared memory.
//
// This is the coverage path: every N the reference accepts (any factorisation, FFT<V>::setPlan
// signalsmith-fft.h:139-185) runs here when no specialised fused kernel exists for it.  It replaces
// permute (:288-293) + the run<inverse> step loop (:295-315) in one launch:
//   * the digit-reversal permutation is folded into the Stockham write index (no table, no gather);
//   * each pass does register butterflies of a compile-time radix (codelets.cuh) or, for primes without
//     a codelet, an O(p^2) sum with PRECOMPUTED roots (the reference recomputes cos/sin per term, :204-205);
//   * data makes one trip HBM -> registers -> shared memory ... -> HBM.
//
// Stockham recursion used by every kernel in this library (N = R * M per pass, P = product of earlier
// radices, butterfly b = racc + P*m', racc < P, m' < M):
//      in  : src[b + (N/R)*j]                      j = 0..R-1        (constant geometry)
//      tw  : W_N^(P*m'*r)                           r = 0..R-1
//      out : dst[racc + P*r + P*R*m']
// First pass reads natural-order input (P = 1), last pass writes natural-order output (M = 1).
#pragma once
#include "codelets.cuh"

namespace ssfft {

constexpr int kMaxPasses = 24;

// shared-memory index padding: one extra element per 16 keeps power-of-two strides off one bank group
SSFFT_HD int spad(int e) { return e + (e >> 4); }

template <typename T>
struct GenericParams {
    int n;                    // transform length handled by this launch
    int npass;
    int radix[kMaxPasses];
    int prod[kMaxPasses];     // P for each pass
    const cx<T> *roots;       // W_n^k, k < n
    // element e of transform t lives at (t / cols) * outer + (t % cols) * inner + e * es   (in cx units)
    long long in_outer, in_inner, in_es;
    long long out_outer, out_inner, out_es;
    int in_cols, out_cols;
    int inverse;              // swap re/im on load and store
    // four-step epilogue: out[k] *= W_M^((t % ep_cols) * k), W_M^q = ep_lo[q & mask] * ep_hi[q >> shift]
    const cx<T> *ep_lo, *ep_hi;
    int ep_shift, ep_cols;
    long long batch;
    int smem_stride;          // cx elements per shared buffer (padded)
    int stage_input;          // first pass has no codelet: copy the transform into shared memory first
};

template <typename T>
struct GlobalSrc {
    const cx<T> *p;
    long long es;
    int swap;
    SSFFT_HD cx<T> load(int e) const {
        cx<T> v = p[(long long)e * es];
        return swap ? cswap(v) : v;
    }
};
template <typename T>
struct SharedSrc {
    const cx<T> *p;
    SSFFT_HD cx<T> load(int e) const { return p[spad(e)]; }
};
template <typename T>
struct SharedDst {
    cx<T> *p;
    SSFFT_HD void store(int e, cx<T> v) const { p[spad(e)] = v; }
};
template <typename T>
struct GlobalDst {
    cx<T> *p;
    long long es;
    int swap;
    const cx<T> *ep_lo, *ep_hi;
    int ep_shift;
    long long ep_col;  // column index multiplying k in the epilogue twiddle
    SSFFT_HD void store(int e, cx<T> v) const {
        if (ep_lo) {
            long long q = ep_col * (long long)e;
            cx<T> w = ep_lo[q & ((1ll << ep_shift) - 1)];
            if (ep_hi) w = cmul(w, ep_hi[q >> ep_shift]);
            v = cmul(v, w);
        }
        p[(long long)e * es] = swap ? cswap(v) : v;
    }
};

// one butterfly of a compile-time radix
template <typename T, int R, typename Src, typename Dst>
SSFFT_HD void pass_codelet(int b, int nr, int P, int mnext, const cx<T> *roots, const Src &src, const Dst &dst) {
    cx<T> v[R];
    sfor<0, R>([&](auto jc) { constexpr int j = decltype(jc)::value; v[j] = src.load(b + nr * j); });
    Dft<R>::run(v);
    const int mp = b / P, racc = b - mp * P;
    if (mnext > 1) {
        const int base = P * mp;
        sfor<1, R>([&](auto rc) { constexpr int r = decltype(rc)::value; v[r] = cmul(v[r], roots[base * r]); });
    }
    const int o = racc + P * R * mp;
    sfor<0, R>([&](auto rc) { constexpr int r = decltype(rc)::value; dst.store(o + P * r, v[r]); });
}

// one OUTPUT of a run-time radix butterfly (work item w = b + nr*r): O(R) per output, O(R^2) per butterfly
template <typename T, typename Src, typename Dst>
SSFFT_HD void pass_anyradix(int w, int R, int nr, int P, int mnext, const cx<T> *roots, const Src &src,
                            const Dst &dst) {
    const int r = w / nr, b = w - r * nr;
    cx<T> sum = src.load(b);
    int q = 0;  // (j*r) mod R
    for (int j = 1; j < R; ++j) {
        q += r;
        if (q >= R) q -= R;
        sum = sum + cmul(src.load(b + nr * j), roots[nr * q]);  // W_R^q = W_n^((n/R) q)
    }
    const int mp = b / P, racc = b - mp * P;
    if (mnext > 1 && r) sum = cmul(sum, roots[P * mp * r]);
    dst.store(racc + P * r + P * R * mp, sum);
}

SSFFT_HD bool radix_has_codelet(int r) {
    return r == 1 || r == 2 || r == 3 || r == 4 || r == 5 || r == 7 || r == 8 || r == 9 || r == 11 || r == 13 ||
           r == 16;
}

// run one pass for the butterflies tx, tx+TX, ... of one transform
template <typename T, typename Src, typename Dst>
SSFFT_HD void run_pass(int R, int n, int P, const cx<T> *roots, const Src &src, const Dst &dst, int tx, int TX) {
    const int nr = n / R, mnext = n / (P * R);
#define SSFFT_CASE(RR)                                                                     \
    case RR:                                                                               \
        for (int b = tx; b < nr; b += TX) pass_codelet<T, RR>(b, nr, P, mnext, roots, src, dst); \
        break;
    switch (R) {
        SSFFT_CASE(1) SSFFT_CASE(2) SSFFT_CASE(3) SSFFT_CASE(4) SSFFT_CASE(5) SSFFT_CASE(7) SSFFT_CASE(8)
        SSFFT_CASE(9) SSFFT_CASE(11) SSFFT_CASE(13) SSFFT_CASE(16)
        default:
            for (int w = tx; w < n; w += TX) pass_anyradix<T>(w, R, nr, P, mnext, roots, src, dst);
    }
#undef SSFFT_CASE
}

#ifdef __CUDACC__
// blockDim = (TX, FPB): threadIdx.x strides over butterflies, threadIdx.y picks the transform in the block.
template <typename T>
__global__ void generic_fft_kernel(const cx<T> *__restrict__ in, cx<T> *__restrict__ out, GenericParams<T> p) {
    SSFFT_DYNAMIC_SMEM(ssfft_smem);
    cx<T> *bufA = reinterpret_cast<cx<T> *>(ssfft_smem) + (size_t)threadIdx.y * 2 * p.smem_stride;
    cx<T> *bufB = bufA + p.smem_stride;
    const int tx = threadIdx.x, TX = blockDim.x;
    const long long t = (long long)blockIdx.x * blockDim.y + threadIdx.y;
    const bool active = t < p.batch;
    const long long tt = active ? t : 0;
    GlobalSrc<T> gsrc{in + (tt / p.in_cols) * p.in_outer + (tt % p.in_cols) * p.in_inner, p.in_es, p.inverse};
    GlobalDst<T> gdst{out + (tt / p.out_cols) * p.out_outer + (tt % p.out_cols) * p.out_inner,
                      p.out_es, p.inverse, p.ep_lo, p.ep_hi, p.ep_shift, (long long)(tt % p.ep_cols)};
    cx<T> *cur = bufA, *nxt = bufB;
    if (p.stage_input) {
        if (active)
            for (int e = tx; e < p.n; e += TX) cur[spad(e)] = gsrc.load(e);
        __syncthreads();
    }
    for (int i = 0; i < p.npass; ++i) {
        const bool first = (i == 0) && !p.stage_input, last = (i == p.npass - 1);
        if (active) {
            if (first && last) run_pass<T>(p.radix[i], p.n, p.prod[i], p.roots, gsrc, gdst, tx, TX);
            else if (first) run_pass<T>(p.radix[i], p.n, p.prod[i], p.roots, gsrc, SharedDst<T>{nxt}, tx, TX);
            else if (last) run_pass<T>(p.radix[i], p.n, p.prod[i], p.roots, SharedSrc<T>{cur}, gdst, tx, TX);
            else run_pass<T>(p.radix[i], p.n, p.prod[i], p.roots, SharedSrc<T>{cur}, SharedDst<T>{nxt}, tx, TX);
        }
        if (!last) {
            __syncthreads();
            cx<T> *s = cur; cur = nxt; nxt = s;
        }
    }
}
#endif  // __CUDACC__

}  // namespace ssfft
