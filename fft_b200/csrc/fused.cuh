// fused.cuh -- specialised single-pass batched FFT kernels (compile-time size and radices).
//
// The hot path of the library (BG, X, RF, RI): one launch does
// what the reference does in permute (:288-293) + every fftStepN pass (:187-286) -- and, for real
// transforms, RealFFT's pack / post-twiddle (:446-473) or pre-twiddle / unpack (:475-502) as well.
//
// Data flow per transform:  HBM --coalesced streaming loads--> registers --radix-R0 butterflies + twiddles-->
//   shared memory (Stockham index = digit reversal folded into the exchange) --> registers --radix-R1 ...-->
//   ... --> registers --coalesced streaming stores--> HBM.     Each input byte is read from HBM once and
//   each output byte written once: algorithmic bytes == DRAM traffic (2*N*sizeof(complex) per transform).
//
// A thread owns E = N/TX points.  In pass p (radix R) it runs E/R butterflies b_u = t + TX*u:
//      in  : src[b + (N/R)*j]                (lanes read consecutive addresses: conflict-free, coalesced)
//      tw  : W_N^(P*m'*r) from a per-pass table laid out [r][m'] (HBM-resident, L1/L2-cached)
//      out : dst[racc + P*r + P*R*m']        (padded index keeps power-of-two strides off one bank group)
// Only a forward transform is generated; the inverse swaps re/im on load and store.
//
// Input staging (FusedCfg::PF): 0 = plain streaming loads; 1 = the next group of transforms is fetched by ONE TMA bulk
// copy into a separate staging buffer while the current group is transformed; 2 = the same copy lands in the exchange
// buffer itself as soon as the last pass has gathered (half the shared memory).  Which one a size uses, and its
// radices / threads / CTAs per SM, come from the tools/kbench.cu sweeps kept under profiles/kbench_*_r01.txt.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <vector>

#include "codelets.cuh"
#include "planner.h"
#include "real_kernels.cuh"

namespace ssfft {

// *_MOD: ModifiedRealFFT (RealFFT<V, halfFreqShift>, signalsmith-fft.h:389-395): input rotated by exp(-i pi n/N)
// on load, pairs (i, N/2-1-i), twiddle phase shifted by half a bin, inverse un-rotated on store.  The rotation
// table is stored right behind the real twiddles: rot = rtw + (N_complex/2 + 1).
enum { FUSED_C2C = 0, FUSED_R2C = 1, FUSED_C2R = 2, FUSED_R2C_MOD = 3, FUSED_C2R_MOD = 4 };
enum { FUSED_CONTIG = 0 };

template <typename T_, int N_, int R0_, int R1_, int R2_, int R3_, int TX_, int FPB_, int MINB_, int PADSHIFT_ = 4, int PF_ = 0>
struct FusedCfg {
    using T = T_;
    static constexpr int N = N_, TX = TX_, FPB = FPB_, MINB = MINB_, PADSHIFT = PADSHIFT_;
    // PF = 1: the next group of transforms is prefetched HBM -> shared staging buffer by one TMA bulk copy
    // (cp.async.bulk + mbarrier) while the current group is being transformed.
    // PF = 2: same, but the copy lands IN the exchange buffer (dense, in front of the padded image) as soon as the last
    // pass has gathered its inputs, so no second buffer is needed: half the shared memory of PF = 1 (more CTAs per SM,
    // and prefetch for sizes whose staging buffer would not fit), for a shorter lead time and one more barrier.
    static constexpr int PF = PF_;
    static constexpr int NP = (R3_ > 1) ? 4 : (R2_ > 1) ? 3 : (R1_ > 1) ? 2 : 1;
    __host__ __device__ static constexpr int radix(int i) { return i == 0 ? R0_ : i == 1 ? R1_ : i == 2 ? R2_ : R3_; }
    // butterflies per thread in pass i (rounded up: a pass whose N/R is not a multiple of TX is "ragged",
    // the excess threads idle behind a predicate) and the register array size a thread needs
    __host__ __device__ static constexpr int bfly(int i) { return (N_ / radix(i) + TX_ - 1) / TX_; }
    __host__ __device__ static constexpr int emax() {
        int e = 0;
        for (int i = 0; i < NP; ++i) e = bfly(i) * radix(i) > e ? bfly(i) * radix(i) : e;
        return e;
    }
    static constexpr int E = emax();
    __host__ __device__ static constexpr int prod(int i) { return i == 0 ? 1 : i == 1 ? R0_ : i == 2 ? R0_ * R1_ : R0_ * R1_ * R2_; }
    __host__ __device__ static constexpr int mnext(int i) { return N / (prod(i) * radix(i)); }
    // twiddle table offset (in cx elements) of pass i: passes 0..NP-2 have (R-1)*mnext entries
    __host__ __device__ static constexpr int tw_off(int i) {
        int o = 0;
        for (int k = 0; k < i; ++k) o += (radix(k) - 1) * mnext(k);
        return o;
    }
    static constexpr int tw_total = tw_off(NP - 1);
    __host__ __device__ static constexpr int pad(int e) { return e + (e >> PADSHIFT); }
    static constexpr int SM_STRIDE = pad(N) + 1;  // cx elements of shared memory per transform
    static constexpr size_t xchg_bytes = (((size_t)SM_STRIDE * FPB * sizeof(cx<T>)) + 127) / 128 * 128;
    static constexpr size_t stage_bytes = PF == 1 ? (size_t)N * FPB * sizeof(cx<T>) : 0;
    static constexpr size_t smem_bytes = xchg_bytes + stage_bytes;
    static_assert(!PF || ((size_t)N * FPB * sizeof(cx<T>)) % 16 == 0, "bulk copies need 16-byte multiples: use an even FPB");
    static_assert(!PF || (R1_ > 1), "prefetch variant needs at least two passes");
    static_assert(R0_ * R1_ * R2_ * R3_ == N_, "radices must multiply to N");
    static_assert(TX_ >= 1 && TX_ <= N_, "bad thread count");
};

// Same configuration without TMA staging: the extended-I/O instantiation (EX) stages its input with a cooperative copy
// loop (frames may overlap, be strided or be only 4-byte aligned, none of which a bulk copy of a whole group can express).
template <typename Cfg>
using NoStaging = FusedCfg<typename Cfg::T, Cfg::N, Cfg::radix(0), Cfg::radix(1), Cfg::radix(2), Cfg::radix(3), Cfg::TX, Cfg::FPB,
                           Cfg::MINB, Cfg::PADSHIFT, 0>;

// Column configuration of a size (opt-in experiment SSFFT_EX_COLCFG=1, NOT YET MEASURED): configurations that hold one
// or two transforms per CTA walk the columns of a matrix one 8-byte element per 32-byte sector.  With 4 (or 2) transforms
// per CTA -- same passes and threads per transform, 1 CTA/SM -- the transform-fastest copy loops of the EX instantiation
// move whole sectors.  column_fpb() = 0: the size has no such configuration (already >= 4 per CTA, or it would not fit).
template <typename Cfg>
__host__ __device__ constexpr int column_fpb() {
    if (Cfg::FPB >= 4 || Cfg::N < 2048) return 0;
    for (int f = 4; f > Cfg::FPB; f /= 2) {
        const size_t smem = ((size_t)(Cfg::pad(Cfg::N) + 1) * f * sizeof(cx<typename Cfg::T>) + 127) / 128 * 128;
        if (Cfg::TX * f <= 1024 && smem <= 227 * 1024) return f;
    }
    return 0;
}
template <typename Cfg, int FPBX = column_fpb<Cfg>()>
using ColumnCfg = FusedCfg<typename Cfg::T, Cfg::N, Cfg::radix(0), Cfg::radix(1), Cfg::radix(2), Cfg::radix(3), Cfg::TX,
                           (FPBX > 0 ? FPBX : Cfg::FPB), 1, Cfg::PADSHIFT, 0>;

// Extended I/O of the EX instantiation (ssfft_exec_*_ex, include/ssfft.h): layouts other than "contiguous batch" and
// pointwise multipliers fused into the first load / last store.  "Elements" are reals on the real side of a RealFFT
// (R2C input, C2R output) and complex values everywhere else.
enum { FUSED_MUL_NONE = 0, FUSED_MUL_REAL = 1, FUSED_MUL_COMPLEX = 2 };
template <typename T>
struct FusedIo {
    long long in_dist, in_stride;    // elements between transforms / between samples of one transform
    long long out_dist, out_stride;
    const T *pre, *post;             // multiplier tables (real or interleaved complex), nullptr = none
    long long pre_dist, post_dist;   // elements between the multipliers of consecutive transforms (0: shared by all)
    int pre_kind, post_kind;         // FUSED_MUL_*
    int in_vec, out_vec;             // real side: sample pairs (2i, 2i+1) may be accessed as one aligned vector
    // "lite" sides need no staging through shared memory: the first pass gathers straight from HBM like the plain kernel
    // (unit stride, real window or none -- overlapping frames only move the base pointer), the last pass stores
    // straight from registers (unit stride, no multiplier).  Set by fused_io_finalize().
    int lite_in, lite_out;
};

// Derived flags of an extended-I/O request (host side; the CPU emulation of the kernels uses the same function).
// mode: FUSED_C2C / FUSED_R2C / FUSED_C2R.  Strides and distances must already hold their defaults.
template <typename T>
inline void fused_io_finalize(FusedIo<T> &io, int mode, const void *in, const void *out) {
    const bool in_real = mode == 1 /* FUSED_R2C */, out_real = mode == 2 /* FUSED_C2R */;
    const size_t vec = 2 * sizeof(T);
    // real side: the sample pair (2i, 2i+1) is one aligned vector when every frame starts on an even sample
    io.in_vec = in_real && io.in_stride == 1 && io.in_dist % 2 == 0 && reinterpret_cast<size_t>(in) % vec == 0;
    io.out_vec = out_real && io.out_stride == 1 && io.out_dist % 2 == 0 && reinterpret_cast<size_t>(out) % vec == 0;
    const bool window_ok = io.pre_kind == FUSED_MUL_NONE ||
                           (io.pre_kind == FUSED_MUL_REAL &&
                            (!in_real || (reinterpret_cast<size_t>(io.pre) % vec == 0 && io.pre_dist % 2 == 0)));
    io.lite_in = mode != 2 && window_ok && (in_real ? io.in_vec != 0 : io.in_stride == 1);
    io.lite_out = io.post_kind == FUSED_MUL_NONE && (out_real ? io.out_vec != 0 : io.out_stride == 1);
}

#ifdef __CUDACC__

template <typename T> struct vec2;
template <> struct vec2<float> { using type = float2; };
template <> struct vec2<double> { using type = double2; };

// streaming (evict-first) global accesses: the data is touched exactly once
template <typename T>
__device__ __forceinline__ cx<T> ld_stream(const cx<T> *p) {
    using V = typename vec2<T>::type;
    V v = __ldcs(reinterpret_cast<const V *>(p));
    return mk<T>(v.x, v.y);
}
template <typename T>
__device__ __forceinline__ void st_stream(cx<T> *p, cx<T> v) {
    using V = typename vec2<T>::type;
    V w; w.x = v.x; w.y = v.y;
    __stcs(reinterpret_cast<V *>(p), w);
}
template <typename T>
__device__ __forceinline__ cx<T> ld_table(const cx<T> *p) {
    using V = typename vec2<T>::type;
    V v = __ldg(reinterpret_cast<const V *>(p));
    return mk<T>(v.x, v.y);
}

#ifdef SSFFT_EMUL
// Host emulation of the kernels (tests/host/simt/simt_emul.h, CPU tests only): same call sites, hooks instead of PTX.
inline unsigned smem_u32(const void *p) { return (unsigned)reinterpret_cast<size_t>(p); }
inline void mbar_init(unsigned long long *bar, unsigned count) { simt::mbar_init(bar, count); }
inline void mbar_expect_tx(unsigned long long *bar, unsigned bytes) { simt::mbar_expect_tx(bar, bytes); }
inline void mbar_wait(unsigned long long *bar, unsigned parity) { simt::mbar_wait(bar, parity); }
inline void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar) {
    simt::bulk_g2s(dst_smem, src_gmem, bytes, bar);
}
#else
// ---- TMA bulk copy + mbarrier (sm_90+/sm_100a PTX; SASS: UBLKCP / SYNCS)
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a lost TMA completion must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    const unsigned addr = smem_u32(bar);
    const long long t0 = clock64();
    for (;;) {
        unsigned done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if (clock64() - t0 > 4000000000LL) __trap();  // ~2 s at 1.9 GHz
    }
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
#endif  // SSFFT_EMUL

// prefetch.global.L2 (SASS CCTL.E.PF2): a hint, no register or shared memory is held and nothing waits for it.  The CPU
// emulation READS the address instead, so a hint outside the buffer faults under a sanitizer.
#ifdef SSFFT_EMUL
inline void prefetch_l2(const void *a) { (void)*static_cast<const volatile unsigned char *>(a); }
#else
__device__ __forceinline__ void prefetch_l2(const void *a) { asm volatile("prefetch.global.L2 [%0];" ::"l"(a)); }
#endif
// Opt-in experiment (NVFLAGS += -DSSFFT_FUSED_L2PF=1, off by default, NOT YET MEASURED): at the top of every iteration a
// CTA hints the group of transforms it will need next (PF = 1: the one after, its TMA copy is issued early in this
// iteration) from HBM into L2, one instruction per 128-byte line.  Aimed at the sizes whose staging has a short lead:
// in-place staging (PF = 2: 2048 ... 16384, 2^k * 3 / 9 above 4608) and the kernels without staging.
#ifndef SSFFT_FUSED_L2PF
#define SSFFT_FUSED_L2PF 0
#endif

// ---- extended I/O (EX instantiation): cooperative, rolled copy loops between HBM and the shared-memory image of a
// group of transforms apply the layout and the multipliers; the passes in between are the plain kernel's.
// SIDE: 0 = complex elements, 1 = real samples accessed one by one, 2 = real samples whose pairs (2i, 2i+1) are one
// aligned vector.  KIND: FUSED_MUL_*.  packed: the buffer is a RealFFT half spectrum whose bin 0 holds (DC, Nyquist):
// a complex multiplier acts on its two components separately (DC * DC', Nyquist * Nyquist'), which is what
// multiplying two such spectra means.  SIDE and KIND are compile-time so that each copy loop is branch-free and its
// loads can be batched; ex_dispatch picks the instantiation once per group.
template <int N> struct ic { static constexpr int value = N; };
template <int SIDE, int KIND, typename T>
__device__ __forceinline__ cx<T> ex_mul(cx<T> v, const T *table, long long first, int idx, bool packed) {
    if constexpr (KIND == FUSED_MUL_REAL) {
        if constexpr (SIDE != 0) {
            const T *w = table + first + 2 * (long long)idx;
            return mk<T>(v.x * __ldg(w), v.y * __ldg(w + 1));
        } else {
            const T w = __ldg(table + first + idx);
            return mk<T>(v.x * w, v.y * w);
        }
    } else if constexpr (KIND == FUSED_MUL_COMPLEX) {
        const cx<T> w = ld_table(reinterpret_cast<const cx<T> *>(table) + first + idx);
        return (packed && idx == 0) ? mk<T>(v.x * w.x, v.y * w.y) : cmul(v, w);
    } else {
        return v;
    }
}
template <int SIDE, int KIND, typename T>
__device__ __forceinline__ cx<T> ex_load(const cx<T> *in, const FusedIo<T> &io, long long tr, int idx, bool packed) {
    cx<T> v;
    if constexpr (SIDE == 0) {
        v = ld_stream(in + tr * io.in_dist + (long long)idx * io.in_stride);
    } else {
        const T *x = reinterpret_cast<const T *>(in) + tr * io.in_dist;
        if constexpr (SIDE == 2) v = ld_stream(reinterpret_cast<const cx<T> *>(x) + idx);
        else v = mk<T>(__ldcs(x + (2 * (long long)idx) * io.in_stride), __ldcs(x + (2 * (long long)idx + 1) * io.in_stride));
    }
    return ex_mul<SIDE, KIND>(v, io.pre, tr * io.pre_dist, idx, packed);
}
template <int SIDE, int KIND, typename T>
__device__ __forceinline__ void ex_store(cx<T> *out, const FusedIo<T> &io, long long tr, int k, cx<T> v, bool packed) {
    v = ex_mul<SIDE, KIND>(v, io.post, tr * io.post_dist, k, packed);
    if constexpr (SIDE == 0) {
        st_stream(out + tr * io.out_dist + (long long)k * io.out_stride, v);
    } else {
        T *y = reinterpret_cast<T *>(out) + tr * io.out_dist;
        if constexpr (SIDE == 2) st_stream(reinterpret_cast<cx<T> *>(y) + k, v);
        else { __stcs(y + (2 * (long long)k) * io.out_stride, v.x); __stcs(y + (2 * (long long)k + 1) * io.out_stride, v.y); }
    }
}
// body(ic<SIDE>, ic<KIND>) for the run-time (side, kind); a complex multiplier on a real side is rejected by the host
template <typename F>
__device__ __forceinline__ void ex_dispatch(int side, int kind, F &&body) {
    if (side == 0) {
        if (kind == FUSED_MUL_REAL) body(ic<0>{}, ic<FUSED_MUL_REAL>{});
        else if (kind == FUSED_MUL_COMPLEX) body(ic<0>{}, ic<FUSED_MUL_COMPLEX>{});
        else body(ic<0>{}, ic<FUSED_MUL_NONE>{});
    } else if (side == 1) {
        if (kind == FUSED_MUL_REAL) body(ic<1>{}, ic<FUSED_MUL_REAL>{});
        else body(ic<1>{}, ic<FUSED_MUL_NONE>{});
    } else {
        if (kind == FUSED_MUL_REAL) body(ic<2>{}, ic<FUSED_MUL_REAL>{});
        else body(ic<2>{}, ic<FUSED_MUL_NONE>{});
    }
}

// EX: extended I/O (FusedIo, ssfft_exec_*_ex) -- strided / overlapping layouts and fused multipliers behind the same
// passes; instantiated for NoStaging configurations only, so that the plain kernels carry none of its address
// arithmetic either.  mode: FUSED_C2C (inverse = 0 / 1), FUSED_R2C, FUSED_C2R (inverse = 1).
// MOD: the ModifiedRealFFT flavours live in their own instantiation so that the plain kernels carry none of their
// address arithmetic (with a run-time flag the C2R gather of the unmodified transform lost 7-20 % of roofline).
// `io` is read by the EX instantiation only (plain launches pass FusedIo<T>{}).
template <typename Cfg, bool MOD = false, bool EX = false>
__global__ void __launch_bounds__(Cfg::TX *Cfg::FPB, Cfg::MINB)
fused_fft_kernel(const cx<typename Cfg::T> *__restrict__ in, cx<typename Cfg::T> *__restrict__ out,
                 const cx<typename Cfg::T> *__restrict__ tw, const cx<typename Cfg::T> *__restrict__ rtw,
                 long long batch, int inverse, int mode, FusedIo<typename Cfg::T> io) {
    using T = typename Cfg::T;
    static_assert(!EX || (Cfg::PF == 0 && !MOD), "extended I/O: no TMA staging, unmodified transforms");
    constexpr bool STAGED = (Cfg::PF != 0) || EX;  // the first pass reads its input from a dense shared-memory image
    constexpr int N = Cfg::N, TX = Cfg::TX, FPB = Cfg::FPB, E = Cfg::E, NP = Cfg::NP;
    constexpr bool PF = Cfg::PF != 0;
    SSFFT_DYNAMIC_SMEM(ssfft_smem);
    __shared__ __align__(8) unsigned long long mbar;
    const int t = threadIdx.x, f = threadIdx.y;
    cx<T> *sm = reinterpret_cast<cx<T> *>(ssfft_smem) + (size_t)f * Cfg::SM_STRIDE;
    const cx<T> *stage_all = reinterpret_cast<const cx<T> *>((Cfg::PF == 2 || EX) ? ssfft_smem : ssfft_smem + Cfg::xchg_bytes);
    // EX: images one element apart from a multiple of N, so that a copy loop running over the transforms of a group
    // (column layouts) does not put all of them on the same shared-memory bank
    constexpr int STAGE_STRIDE = EX ? N + 1 : N;
    const cx<T> *stage = stage_all + (size_t)f * STAGE_STRIDE;
    const long long groups = (batch + FPB - 1) / FPB;
    const bool leader = (t == 0 && f == 0);
    constexpr bool mod = MOD;                                             // half-bin-shifted real transform
    const bool is_r2c = mod ? (mode == FUSED_R2C_MOD) : (mode == FUSED_R2C), is_c2r = mod ? (mode == FUSED_C2R_MOD) : (mode == FUSED_C2R);
    const cx<T> *rot = rtw + (N / 2 + 1);                                  // modifiedRotations (:426-432), modified plans only
    unsigned parity = 0;

    // one thread asks the TMA engine for a whole group of transforms (contiguous in HBM).  Bulk copies
    // move multiples of 16 bytes: a ragged last group of an odd-length size falls back to plain loads.
    // (the source must be 16-byte aligned as well: a user pointer offset by one complex element uses plain loads)
    const bool in_aligned = (reinterpret_cast<unsigned long long>(in) & 15ull) == 0;
    auto group_bytes = [&](long long g) -> unsigned {
        const long long first_tr = g * FPB;
        const long long cnt = (batch - first_tr < FPB) ? (batch - first_tr) : FPB;
        return (unsigned)(cnt * N * sizeof(cx<T>));
    };
    auto prefetch = [&](long long g) {
        if (leader && g < groups) {
            const unsigned bytes = group_bytes(g);
            if (bytes % 16 == 0 && in_aligned) {
                mbar_expect_tx(&mbar, bytes);
                bulk_g2s(const_cast<cx<T> *>(stage_all), in + g * FPB * N, bytes, &mbar);
            }
        }
    };
    if constexpr (PF) {
        if (leader) mbar_init(&mbar, 1);
        __syncthreads();
        prefetch(blockIdx.x);
    }

    for (long long g = blockIdx.x; g < groups; g += gridDim.x) {
        const long long tr = g * FPB + f;
        const bool active = tr < batch;
        if constexpr (SSFFT_FUSED_L2PF != 0 && !EX) {
            const long long gp = g + (Cfg::PF == 1 ? 2 : 1) * (long long)gridDim.x;
            if (gp < groups) {
                const unsigned bytes = group_bytes(gp);
                const char *base = reinterpret_cast<const char *>(in + gp * FPB * N);
                for (unsigned off = (unsigned)(f * TX + t) * 128u; off < bytes; off += (unsigned)(TX * FPB) * 128u) prefetch_l2(base + off);
            }
        }
        const cx<T> *gin = in + (active ? tr : 0) * N;
        cx<T> *gout = out + (active ? tr : 0) * N;
        if constexpr (EX) {  // lite sides: the transform's own base pointer (distances count reals on a real side)
            gin = reinterpret_cast<const cx<T> *>(reinterpret_cast<const T *>(in) + (active ? tr : 0) * io.in_dist * (is_r2c ? 1 : 2));
            gout = reinterpret_cast<cx<T> *>(reinterpret_cast<T *>(out) + (active ? tr : 0) * io.out_dist * (is_c2r ? 1 : 2));
        }
        cx<T> v[E];
        if constexpr (PF) {
            const unsigned bytes = group_bytes(g);
            if (bytes % 16 == 0 && in_aligned) {
                mbar_wait(&mbar, parity);  // this group's input has landed in the staging buffer
                parity ^= 1u;
            } else {  // ragged tail: cooperative plain copy into the staging buffer
                cx<T> *st = const_cast<cx<T> *>(stage_all);
                const int n_el = (int)(bytes / sizeof(cx<T>));
                for (int i = f * TX + t; i < n_el; i += TX * FPB) st[i] = ld_stream(in + g * FPB * N + i);
                __syncthreads();
            }
        }
        // EX: transforms of this group, and the cooperative gather of their inputs -- layout and pre-multiplier are
        // applied on the way into the dense shared-memory image the first pass reads (as with PF = 2, minus the TMA)
        const int ex_cnt = (int)((batch - g * FPB < FPB) ? (batch - g * FPB) : FPB);
        bool ex_lite_in = false, ex_lite_out = false;
        if constexpr (EX) { ex_lite_in = io.lite_in != 0; ex_lite_out = io.lite_out != 0; }
        if constexpr (EX) if (!ex_lite_in) {
            cx<T> *st = const_cast<cx<T> *>(stage_all);
            // consecutive threads take consecutive samples of a transform -- or, when the transforms of the group are
            // closer together in memory than the samples of one transform (columns of a matrix), consecutive transforms
            const bool tr_fastest = FPB > 1 && ex_cnt == FPB && io.in_dist < io.in_stride;
            ex_dispatch(is_r2c ? (io.in_vec ? 2 : 1) : 0, io.pre_kind, [&](auto side, auto kind) {
#pragma unroll 8
                for (int i = f * TX + t; i < ex_cnt * N; i += TX * FPB) {
                    const int ff = tr_fastest ? i % FPB : i / N, idx = tr_fastest ? i / FPB : i - ff * N;
                    st[ff * STAGE_STRIDE + idx] = ex_load<decltype(side)::value, decltype(kind)::value>(in, io, g * FPB + ff, idx, is_c2r);
                }
            });
            __syncthreads();
        }
        auto load_in = [&](int idx) -> cx<T> {
            if constexpr (STAGED) return stage[idx];
            else return ld_stream(gin + idx);
        };

        // C2R (RealFFT::ifft :478-492): the pre-twiddle is applied on the fly while gathering pass 0 -- the
        // element buf[i] only needs in[i], in[N-i] and tw[min(i, N-i)], all of which this thread can fetch
        // itself, so no shared-memory round trip is needed (see the gather below).
        sfor<0, NP>([&](auto pc) {
            constexpr int p = decltype(pc)::value;
            constexpr int R = Cfg::radix(p), P = Cfg::prod(p), MN = Cfg::mnext(p), NR = N / R, U = Cfg::bfly(p);
            constexpr bool first = (p == 0), last = (p == NP - 1);
            constexpr bool ragged = (NR % TX) != 0;  // some threads have no butterfly in their last slot
            auto has = [&](int u) { return !ragged || (t + TX * u) < NR; };
            // ---- gather inputs
            if constexpr (first) {
                if (EX && ex_lite_in) {
                    // extended I/O, lite input: gather straight from HBM as the plain kernel does; the window (if any)
                    // is one more L1-resident load per element.  R2C: element idx = samples (2 idx, 2 idx + 1).
                    if constexpr (EX) {
                        if (active) {
                            const T *wtab = io.pre + tr * io.pre_dist;
                            auto gather = [&](auto kind) {
                                constexpr int KD = decltype(kind)::value;
#pragma unroll
                                for (int u = 0; u < U; ++u)
#pragma unroll
                                    for (int j = 0; j < R; ++j)
                                        if (has(u)) {
                                            const int idx = t + TX * u + NR * j;
                                            cx<T> x = ld_stream(gin + idx);
                                            if constexpr (KD == FUSED_MUL_REAL) {
                                                if (is_r2c) {
                                                    const cx<T> w = ld_table(reinterpret_cast<const cx<T> *>(wtab) + idx);
                                                    x = mk<T>(x.x * w.x, x.y * w.y);
                                                } else {
                                                    const T w = __ldg(wtab + idx);
                                                    x = mk<T>(x.x * w, x.y * w);
                                                }
                                            }
                                            v[u * R + j] = inverse ? cswap(x) : x;
                                        }
                            };
                            if (io.pre_kind == FUSED_MUL_REAL) gather(ic<FUSED_MUL_REAL>{});
                            else gather(ic<FUSED_MUL_NONE>{});
                        }
                    }
                } else if (is_c2r) {
                    if (STAGED || active) {
#pragma unroll
                        for (int u = 0; u < U; ++u)
#pragma unroll
                            for (int j = 0; j < R; ++j) {
                                if (!has(u)) continue;
                                const int i = t + TX * u + NR * j;
                                const int ci = mod ? N - 1 - i : (i ? N - i : 0);
                                const bool lo = mod ? (2 * i <= N - 1) : (2 * i <= N);  // first or second element of its pair
                                const cx<T> vi = load_in(i), vc = load_in(ci);
                                const cx<T> w = ld_table(rtw + (lo ? i : ci));
                                cx<T> bi, bc;
                                c2r_pair(lo ? vi : vc, lo ? vc : vi, w, bi, bc);
                                cx<T> x = lo ? bi : bc;
                                if (i == 0 && !mod) x = mk<T>(vi.x + vi.y, vi.x - vi.y);  // (DC, Nyquist) unpack  :478-481
                                v[u * R + j] = cswap(x);
                            }
                    }
                } else if (mod && mode == FUSED_R2C_MOD) {
                    if (STAGED || active) {
#pragma unroll
                        for (int u = 0; u < U; ++u)
#pragma unroll
                            for (int j = 0; j < R; ++j)
                                if (has(u)) {
                                    const int idx = t + TX * u + NR * j;
                                    v[u * R + j] = cmul(load_in(idx), ld_table(rot + idx));  // pre-rotation :450-452
                                }
                    }
                } else if (STAGED || active) {
                    if (inverse) {
#pragma unroll
                        for (int u = 0; u < U; ++u)
#pragma unroll
                            for (int j = 0; j < R; ++j)
                                if (has(u)) v[u * R + j] = cswap(load_in(t + TX * u + NR * j));
                    } else {
#pragma unroll
                        for (int u = 0; u < U; ++u)
#pragma unroll
                            for (int j = 0; j < R; ++j)
                                if (has(u)) v[u * R + j] = load_in(t + TX * u + NR * j);
                    }
                }
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int j = 0; j < R; ++j)
                        if (has(u)) v[u * R + j] = sm[Cfg::pad(t + TX * u + NR * j)];
                __syncthreads();  // everyone has read: the buffer may be overwritten
            }
            if constexpr (first && Cfg::PF == 2) __syncthreads();  // everyone holds its inputs: the dense image may be overwritten
            if constexpr (first && EX) { if (!ex_lite_in) __syncthreads(); }
            if constexpr (last && !first && Cfg::PF == 2) {
                if (!is_r2c) prefetch(g + gridDim.x);  // the exchange buffer is free from here on (R2C: after its epilogue)
            }
            // ---- butterflies + inter-pass twiddles
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (!has(u)) continue;
                cx<T> w[R];
#pragma unroll
                for (int j = 0; j < R; ++j) w[j] = v[u * R + j];
                Dft<R>::run(w);
                if constexpr (!last) {
                    const int b = t + TX * u;
                    const int mp = b / P;
                    const cx<T> *twp = tw + Cfg::tw_off(p) + mp;
#pragma unroll
                    for (int r = 1; r < R; ++r) w[r] = cmul(w[r], ld_table(twp + (r - 1) * MN));
                }
#pragma unroll
                for (int j = 0; j < R; ++j) v[u * R + j] = w[j];
            }
            // ---- scatter outputs
            if constexpr (!last) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int b = t + TX * u;
                    const int mp = b / P, racc = b - mp * P;
                    const int o = racc + P * R * mp;
                    if (has(u)) {
#pragma unroll
                        for (int r = 0; r < R; ++r) sm[Cfg::pad(o + P * r)] = v[u * R + r];
                    }
                }
                __syncthreads();
                if constexpr (Cfg::PF == 1 && first) prefetch(g + gridDim.x);  // every thread has consumed the staging buffer
            } else {
                // last pass: P == N/R, m' == 0, racc == b  ->  natural-order index b + P*r
                bool stored = false;
                if constexpr (EX) if (!ex_lite_out) {
                    stored = true;
                    // results -> shared memory in natural order; ONE cooperative rolled loop per group then applies the
                    // RealFFT post-twiddle (R2C), the post-multiplier and the output layout
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int r = 0; r < R; ++r)
                            if (has(u)) sm[Cfg::pad(t + TX * u + P * r)] = inverse ? cswap(v[u * R + r]) : v[u * R + r];
                    __syncthreads();
                    const cx<T> *img = reinterpret_cast<const cx<T> *>(ssfft_smem);
                    const bool tr_fastest = FPB > 1 && ex_cnt == FPB && io.out_dist < io.out_stride;  // as in the gather
                    if (is_r2c) {
                        constexpr int H2 = N / 2 + 1;  // pairs (k, N-k), k = 0 .. N/2  (RealFFT::fft :459-472)
                        ex_dispatch(0, io.post_kind, [&](auto side, auto kind) {
                            constexpr int SD = decltype(side)::value, KD = decltype(kind)::value;
#pragma unroll 2
                            for (int i = f * TX + t; i < ex_cnt * H2; i += TX * FPB) {
                                const int ff = tr_fastest ? i % FPB : i / H2, k = tr_fastest ? i / FPB : i - ff * H2;
                                const cx<T> *row = img + (size_t)ff * Cfg::SM_STRIDE;
                                const cx<T> zi = row[Cfg::pad(k)], zc = row[Cfg::pad(k ? N - k : 0)];
                                const long long trk = g * FPB + ff;
                                if (k == 0) {
                                    ex_store<SD, KD>(out, io, trk, 0, mk<T>(zi.x + zi.y, zi.x - zi.y), true);  // (DC, Nyquist)
                                } else {
                                    cx<T> oi, oc;
                                    r2c_pair(zi, zc, ld_table(rtw + k), oi, oc);
                                    ex_store<SD, KD>(out, io, trk, k, oi, true);
                                    ex_store<SD, KD>(out, io, trk, N - k, oc, true);  // self-pair: second write wins
                                }
                            }
                        });
                    } else {
                        ex_dispatch(is_c2r ? (io.out_vec ? 2 : 1) : 0, io.post_kind, [&](auto side, auto kind) {
#pragma unroll 4
                            for (int i = f * TX + t; i < ex_cnt * N; i += TX * FPB) {
                                const int ff = tr_fastest ? i % FPB : i / N, k = tr_fastest ? i / FPB : i - ff * N;
                                ex_store<decltype(side)::value, decltype(kind)::value>(out, io, g * FPB + ff, k,
                                                                                       img[(size_t)ff * Cfg::SM_STRIDE + Cfg::pad(k)], false);
                            }
                        });
                    }
                    __syncthreads();  // the next group's gather overwrites the image
                }
                if (stored) {
                    // extended I/O: done above
                } else if (is_r2c) {
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int r = 0; r < R; ++r)
                            if (has(u)) sm[Cfg::pad(t + TX * u + P * r)] = v[u * R + r];
                    __syncthreads();
                    // ---------------- R2C epilogue: post-twiddle pairs (RealFFT::fft :459-472)
                    if (active) {
                        // pairs (i, N-i), i = 0 .. N/2, spread over the threads with a compile-time trip count so
                        // the shared-memory reads and twiddle loads of all iterations are in flight together
                        constexpr int H2 = N / 2 + 1, ITER = (H2 + TX - 1) / TX, CH = 4;  // CH pairs in flight
#pragma unroll
                        for (int u0 = 0; u0 < ITER; u0 += CH) {
                            cx<T> zi[CH], zc[CH], tw_[CH];
#pragma unroll
                            for (int k = 0; k < CH; ++k) {
                                const int i = t + TX * (u0 + k);
                                if (u0 + k < ITER && i < H2 && (!mod || 2 * i <= N - 1)) {
                                    zi[k] = sm[Cfg::pad(i)];
                                    zc[k] = sm[Cfg::pad(mod ? N - 1 - i : (i ? N - i : 0))];
                                    tw_[k] = ld_table(rtw + i);
                                }
                            }
#pragma unroll
                            for (int k = 0; k < CH; ++k) {
                                const int i = t + TX * (u0 + k);
                                if (u0 + k < ITER && i < H2 && (!mod || 2 * i <= N - 1)) {
                                    if (i == 0 && !mod) {
                                        st_stream(gout, mk<T>(zi[k].x + zi[k].y, zi[k].x - zi[k].y));  // (DC, Nyquist) :459-462
                                    } else {
                                        cx<T> oi, oc;
                                        r2c_pair(zi[k], zc[k], tw_[k], oi, oc);
                                        st_stream(gout + i, oi);
                                        st_stream(gout + (mod ? N - 1 - i : N - i), oc);  // self-pair: second write wins
                                    }
                                }
                            }
                        }
                    }
                    __syncthreads();
                    if constexpr (Cfg::PF == 2) prefetch(g + gridDim.x);
                } else if (mod && mode == FUSED_C2R_MOD) {
                    if (active) {
#pragma unroll
                        for (int u = 0; u < U; ++u)
#pragma unroll
                            for (int r = 0; r < R; ++r)
                                if (has(u)) {
                                    const int k = t + TX * u + P * r;
                                    st_stream(gout + k, cmulc(cswap(v[u * R + r]), ld_table(rot + k)));  // un-rotation :497-498
                                }
                    }
                } else if (active) {
                    if (inverse) {
#pragma unroll
                        for (int u = 0; u < U; ++u)
#pragma unroll
                            for (int r = 0; r < R; ++r)
                                if (has(u)) st_stream(gout + t + TX * u + P * r, cswap(v[u * R + r]));
                    } else {
#pragma unroll
                        for (int u = 0; u < U; ++u)
#pragma unroll
                            for (int r = 0; r < R; ++r)
                                if (has(u)) st_stream(gout + t + TX * u + P * r, v[u * R + r]);
                    }
                }
            }
        });
    }
}

#endif  // __CUDACC__

// ---------------------------------------------------------------------------------------------
// registry: which (precision, N) have a specialised kernel, and how to launch it
// ---------------------------------------------------------------------------------------------
struct FusedEntry {
    int prec;      // 0 f32, 1 f64
    int n;
    const char *name;
    int tw_total;  // cx elements
    int radix[4], np;
    int real_only;  // 1: tuned for the R2C / C2R flavours, picked for real plans only (the C2C entry of the size differs)
    int (*launch)(const void *tw, const void *in, void *out, long long batch, int inverse, int mode, const void *rtw,
                  cudaStream_t s);
    // extended I/O (strided / overlapping layouts, fused multipliers): io -> host FusedIo<T>; unmodified transforms only
    int (*launch_ex)(const void *tw, const void *in, void *out, long long batch, int inverse, int mode, const void *rtw,
                     const void *io, cudaStream_t s);
    // same through the column configuration of the size (ColumnCfg); null when it has none
    int (*launch_ex_cols)(const void *tw, const void *in, void *out, long long batch, int inverse, int mode, const void *rtw,
                          const void *io, cudaStream_t s);
};

const std::vector<FusedEntry> &fused_registry();
int fused_waves();  // resident-CTA waves per launch before CTAs loop (env SSFFT_FUSED_WAVES, default 4)

// for_real: the plan is a RealFFT -- prefer an entry tuned for the real flavours when the size has one
template <typename T>
inline int find_fused(size_t n, int for_real) {
    const int prec = sizeof(T) == 4 ? 0 : 1;
    const auto &reg = fused_registry();
    int any = -1;
    for (size_t i = 0; i < reg.size(); ++i)
        if (reg[i].prec == prec && (size_t)reg[i].n == n) {
            if (reg[i].real_only) { if (for_real) return (int)i; }
            else if (any < 0) any = (int)i;
        }
    return any;
}
inline const char *fused_name(int id) { return fused_registry()[id].name; }

// per-pass twiddle tables laid out [r-1][m']: W_N^(P*m'*r)
template <typename T>
inline int build_fused_twiddles(int id, void **d_out) {
    const FusedEntry &e = fused_registry()[id];
    std::vector<T> h(2 * (size_t)(e.tw_total > 0 ? e.tw_total : 1));
    size_t o = 0;
    int P = 1;
    for (int p = 0; p + 1 < e.np; ++p) {
        const int R = e.radix[p], MN = e.n / (P * R);
        for (int r = 1; r < R; ++r)
            for (int m = 0; m < MN; ++m) {
                unsigned long long q = (unsigned long long)P * m * r % (unsigned long long)e.n;
                long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)q / (long double)e.n;
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

template <typename T>
inline int launch_fused(int id, const void *tw, const void *in, void *out, long long batch, int inverse, int mode,
                        const void *rtw, cudaStream_t s, std::atomic<uint64_t> *counter) {
    const FusedEntry &e = fused_registry()[id];
    int rc = e.launch(tw, in, out, batch, inverse, mode, rtw, s);
    if (counter) ++*counter;
    return rc;
}

}  // namespace ssfft
