// generic_plan.h -- host-side planning of the generic (any-size) kernel path: how a length is laid out over a CTA,
// the kernel parameters of a launch, and the generic four-step (two generic stages + epilogue twiddle) for lengths that
// do not fit one CTA.  No CUDA calls: ssfft.cu uploads the tables and launches, tests/host/generic_emul.cpp runs the
// same plans through the CPU execution of generic_fft_kernel.
#pragma once
#include <cstring>

#include "generic.cuh"
#include "plan.h"
#include "planner.h"

namespace ssfft {

// largest length one CTA can hold in the generic kernel's two padded shared buffers
inline size_t generic_limit(size_t elem, int smem_max) {
    size_t n = (size_t)smem_max / (2 * elem);
    while (n > 1 && 2 * (size_t)(spad((int)n) + 1) * elem > (size_t)smem_max) --n;
    return n;
}

// Radices, thread / transform geometry and shared-memory footprint of one single-launch transform of length n.
// elem = sizeof(complex<V>).  Returns false when the length cannot run in one CTA (too many passes, too much memory).
inline bool plan_generic_stage(GenericStage &st, size_t n, size_t elem, int smem_max) {
    st.n = (int)n;
    st.radix = choose_radices(n);
    st.prod.clear();
    int P = 1, maxr = 1;
    for (int r : st.radix) { st.prod.push_back(P); P *= r; if (r > maxr) maxr = r; }
    if ((int)st.radix.size() > kMaxPasses) return false;
    st.stage_input = radix_has_codelet(st.radix[0]) ? 0 : 1;
    st.smem_stride = spad((int)n) + 1;
    const size_t per = 2 * (size_t)st.smem_stride * elem;
    if (per > (size_t)smem_max) return false;
    // threads per transform: about one codelet butterfly each, power of two in [1, 256]
    int want = (int)(n / (size_t)(maxr > 16 ? 1 : maxr));
    if (!radix_has_codelet(maxr)) want = (int)n;  // any-radix passes parallelise over outputs
    int tx = 1;
    while (tx < want && tx < 256) tx *= 2;
    st.tx = tx;
    int fpb = 256 / tx;
    if (fpb < 1) fpb = 1;
    while (fpb > 1 && per * (size_t)fpb > (size_t)smem_max / 2) fpb /= 2;  // leave room for 2 CTAs/SM
    st.fpb = fpb;
    st.smem_bytes = per * (size_t)fpb;
    return true;
}

// element e of transform t lives at (t / cols) * outer + (t % cols) * inner + e * es   (cx units), per side
struct GenericLayout {
    long long in_outer, in_inner, in_es;
    int in_cols;
    long long out_outer, out_inner, out_es;
    int out_cols;
};

// kernel parameters of one launch; ep_lo / ep_hi = null: no four-step epilogue twiddle
template <typename T>
inline GenericParams<T> make_generic_params(const GenericStage &st, const void *roots, long long batch, const GenericLayout &L,
                                            int inverse, const void *ep_lo, const void *ep_hi, int ep_shift, int ep_cols) {
    GenericParams<T> p;
    memset(&p, 0, sizeof(p));
    p.n = st.n;
    p.npass = (int)st.radix.size();
    for (int i = 0; i < p.npass; ++i) { p.radix[i] = st.radix[i]; p.prod[i] = st.prod[i]; }
    p.roots = (const cx<T> *)roots;
    p.in_outer = L.in_outer; p.in_inner = L.in_inner; p.in_es = L.in_es; p.in_cols = L.in_cols;
    p.out_outer = L.out_outer; p.out_inner = L.out_inner; p.out_es = L.out_es; p.out_cols = L.out_cols;
    p.inverse = inverse;
    p.ep_lo = (const cx<T> *)ep_lo;
    p.ep_hi = (const cx<T> *)ep_hi;
    p.ep_shift = ep_shift;
    p.ep_cols = ep_cols > 0 ? ep_cols : 1;
    p.batch = batch;
    p.smem_stride = st.smem_stride;
    p.stage_input = st.stage_input;
    return p;
}

// Generic four-step n = n1 * n2 (x[n1][n2] row-major): (1) length-n1 transforms down the n2 columns, multiplied by
// W_n^(c * k1) on store (epilogue), into a scratch; (2) length-n2 transforms along the rows, stored transposed
// X[k1 + n1 * k2].  The epilogue twiddle W_n^q, q < n, is the product of two small tables: q = hi * 2^shift + lo.
struct GenericFourStep {
    size_t n1 = 0, n2 = 0;
    int ep_shift = 0;
    size_t lo_count = 0, hi_count = 0;  // entries of the tables W_n^lo (step 1) and W_n^(hi * 2^shift)
};
inline bool plan_generic_fourstep(GenericFourStep &f, size_t n, size_t limit) {
    if (!choose_split(n, limit, &f.n1, &f.n2)) return false;
    int shift = 0;
    while ((1ull << (2 * shift)) < n) ++shift;
    f.ep_shift = shift;
    f.lo_count = (size_t)1 << shift;
    f.hi_count = (n + f.lo_count - 1) / f.lo_count;
    return true;
}
inline GenericLayout fourstep_col_layout(size_t n, size_t n2) {  // stage 1 over nb * n2 transforms
    return GenericLayout{(long long)n, 1, (long long)n2, (int)n2, (long long)n, 1, (long long)n2, (int)n2};
}
inline GenericLayout fourstep_row_layout(size_t n, size_t n1, size_t n2) {  // stage 2 over nb * n1 transforms
    return GenericLayout{(long long)n, (long long)n2, 1, (int)n1, (long long)n, 1, (long long)n1, (int)n1};
}

}  // namespace ssfft
