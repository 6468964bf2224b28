// flat.cuh -- four-step (N = N1 * N2) as ONE persistent, warp-specialised launch fed by a ticket queue.
//
// Same decomposition as tiled.cuh (GPU analogue of the reference's cache-blocking branch, signalsmith-fft.h:130-133):
// N2 column FFTs of length N1, the twiddle W_N^(n2*k1), N1 row FFTs of length N2 stored transposed, the intermediate in an
// L2-resident scratch.  What changes is how the work is scheduled and how the data reaches the butterflies -- the
// round-1 kernels (one thread-block cluster per transform, per-thread loads, an N-entry twiddle table in L2) were
// latency-bound (ncu: long-scoreboard 36 %, cluster barrier 8 %, 16 of 148 SMs stranded by 4-CTA clusters):
//
//  * TICKETS instead of clusters.  The work of a call is a list of tickets
//        phase p = 0, 1, ... :  [tiles1 column tiles of transform p] [tiles2 row tiles of transform p - D]
//    handed out by one atomic counter.  A row tile waits until cnt1[b] says every column tile of its transform is in the
//    scratch; a column tile waits until the previous user of its scratch slot (transform b - NS) has been read
//    (cnt2).  Every dependency of a ticket has a SMALLER ticket number and tickets are only ever held by running CTAs,
//    so the launch needs no co-residency guarantee and cannot deadlock; D phases of distance make the waits free in
//    practice.  No clusters: all 148 SMs work, no cluster barrier, no release fence per transform.
//  * WARP SPECIALISATION.  Warp NC/32 of every CTA is a producer: it takes the tickets, polls the dependencies and has
//    the TMA engine bring the tile into a shared-memory ring (column tiles: ONE cp.async.bulk.tensor box [N1][CT] of the
//    (batch, N1, N2) input; row tiles: ONE cp.async.bulk of the contiguous tile-major scratch block), `full` / `empty`
//    mbarriers per ring slot.  The NC consumer threads never issue a global load: they read the ring, run the
//    butterflies, store.  Completion (cnt1 / cnt2) is signalled by the producer after the consumers' `done` mbarrier.
//  * NO N-ENTRY TWIDDLE TABLE.  With k1 = r0 + R0*r1 (r0 = output digit of the first column pass) the four-step twiddle
//    factors as W_N^(n2*r0) * W_N^(n2*R0*r1).  The first factor merges with the inter-pass twiddle of the column FFT:
//    W_N1^(b*r0) * W_N^(n2*r0) = g^r0 with g = W_N^(b*N2 + n2) -- the powers of ONE number per butterfly, built in
//    registers from g, g^2, g^4, g^8 (two tiny tables: W_N1^(b*2^k) and W_N^(n2*2^k)).  The second factor is a
//    [N2][R1] table whose slice for a tile (CT*R1 values) rides along with the tile's TMA copy.
//  * forward / inverse are separate instantiations (the re/im swap of the inverse costs no select instructions).
#pragma once
#include <cstdio>
#include <cstdlib>

#include "tiled.cuh"

namespace ssfft {

template <typename T>
struct FlatParams {
    const cx<T> *in;
    cx<T> *out;
    cx<T> *scratch;         // nslots * scratch_per elements
    const cx<T> *tw_b;      // row-stage pass twiddles ([r-1][m'] as in tiled.cuh)
    const cx<T> *ga[2], *gb[2];  // per non-last column pass p: W_(N1/P)^(m'*2^k) [N1/(P R)][LOG R]  and  W_(N/P)^(n2*2^k) [N2][LOG R]
    const cx<T> *s4;        // W_N^(n2*P_last*r) [N2][R_last]
    unsigned *ctrl;         // [0] ticket counter; cnt1 = ctrl + 32; cnt2 = cnt1 + cap
    long long batch, user_stride, scratch_per, cap;
    int nslots, delay, discard;
    unsigned long long *stats;  // -DSSFFT_FLAT_STATS builds only: kFlatStats counters per CTA (else unused, null)
};
constexpr int kFlatStats = 16;
constexpr int kFlatHelpers = 64;  // two helper warps per CTA: the TMA producer and the completion signaller
#ifndef SSFFT_FLAT_STATS
#define SSFFT_FLAT_STATS 0
#endif
// -DSSFFT_FLAT_NOCOMPUTE=1 (measurement build): the consumers skip butterflies and twiddles -- same copies, barriers,
// dependency traffic and stores, wrong results: what the schedule and the memory system can do without the arithmetic
#ifndef SSFFT_FLAT_NOCOMPUTE
#define SSFFT_FLAT_NOCOMPUTE 0
#endif
// -DSSFFT_FLAT_NODEPS=1 (measurement build, wrong results): dependencies are never waited for -- the ceiling of the copy /
// store pipeline alone
#ifndef SSFFT_FLAT_NODEPS
#define SSFFT_FLAT_NODEPS 0
#endif

__host__ __device__ constexpr int flat_ilog2(int v) { return v <= 1 ? 0 : 1 + flat_ilog2(v / 2); }
__host__ __device__ constexpr int flat_topbit(int v) { return 1 << flat_ilog2(v); }

// shared-memory map of a CTA (bytes).  INPLACE: a ring slot is also the exchange buffer of the tile it holds (the TMA
// copy lands dense at its start, the passes overwrite it with the padded image, the tile's slice of s4 sits behind it),
// so a ring of two fits three CTAs per SM; otherwise one separate exchange buffer, dense slots and a small ring of s4
// slices.  The row-stage pass table lives in shared memory when it is small (two-pass row tiles), else it is read
// through L1 like the column-stage tables.
template <typename CfgA, typename CfgB, int NSTAGE, bool INPLACE>
struct FlatLayout {
    using T = typename CfgA::T;
    static constexpr size_t al(size_t v) { return (v + 127) / 128 * 128; }
    static constexpr size_t kTileA = (size_t)CfgA::L * CfgA::CT * sizeof(cx<T>), kTileB = (size_t)CfgB::L * CfgB::CT * sizeof(cx<T>);
    static constexpr size_t kExchA = CfgA::smem_bytes, kExchB = CfgB::smem_bytes;
    static constexpr size_t kExch = al(kExchA > kExchB ? kExchA : kExchB);
    static constexpr size_t kSBlk = al((size_t)CfgA::CT * CfgA::radix(CfgA::NP - 1) * sizeof(cx<T>));
    static constexpr size_t kSlot = INPLACE ? kExch + kSBlk : al(kTileA > kTileB ? kTileA : kTileB);
    static constexpr bool kTwBShared = (size_t)CfgB::tw_total * sizeof(cx<T>) <= 2048;
    static constexpr size_t kTwB = kTwBShared ? al((size_t)(CfgB::tw_total > 0 ? CfgB::tw_total : 1) * sizeof(cx<T>)) : 0;
    static constexpr size_t oExch = 0, oSlots = oExch + (INPLACE ? 0 : kExch), oSBlk = oSlots + NSTAGE * kSlot,
                            oTwB = oSBlk + (INPLACE ? 0 : (NSTAGE + 1) * kSBlk), oDesc = oTwB + kTwB,
                            oBars = oDesc + al((size_t)(2 * NSTAGE + 1) * 32);
    static constexpr size_t oSync = oBars + al((size_t)(3 * NSTAGE + 1) * 8);  // producer <-> signaller counters
    static constexpr size_t smem_bytes = oSync + 128;
};

struct FlatDesc {  // what the producer tells the consumers about a ring slot (and itself about an unsignalled item)
    int kind;      // 0 column tile, 1 row tile, 2 end of work
    int tile;
    long long b;
    long long t_issue, pad;  // statistics builds: clock64() when the copy was issued
};
static_assert(sizeof(FlatDesc) == 32, "descriptor slots are 32 bytes");

#ifdef __CUDACC__

#ifdef SSFFT_EMUL
inline void mbar_arrive(unsigned long long *bar) { simt::mbar_arrive(bar); }
inline bool mbar_test(unsigned long long *bar, unsigned parity) { return simt::mbar_test(bar, parity); }
inline void consumer_barrier(int n) { simt::named_barrier(1, (unsigned)n); }
inline void producer_idle() { simt::spin_yield(); }
inline void producer_moved() { simt::state().progress = true; }  // global counters changed: not a deadlock
inline void fence_proxy_async() {}
template <typename T>
inline void tma_tile_3d(cx<T> *dst, const void *, const cx<T> *in, int n1, int n2, int box_rows, int ct, int col0, int row0, long long b,
                        unsigned long long *bar) {
    for (int r = 0; r < box_rows; ++r)
        bulk_g2s(dst + (size_t)r * ct, in + b * (long long)n1 * n2 + (long long)(row0 + r) * n2 + col0, (unsigned)(ct * sizeof(cx<T>)), bar);
}
#else
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(unsigned long long *bar, unsigned parity) {
    unsigned done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void consumer_barrier(int n) { asm volatile("bar.sync 1, %0;" ::"r"(n) : "memory"); }
__device__ __forceinline__ void producer_idle() { __nanosleep(64); }
__device__ __forceinline__ void producer_moved() {}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// box [box_rows][ct] at (col0, row0) of transform b of the (batch, N1, N2) input (tensor map built by the launcher)
template <typename T>
__device__ __forceinline__ void tma_tile_3d(cx<T> *dst, const void *tmap, const cx<T> *, int, int, int, int, int col0, int row0,
                                            long long b, unsigned long long *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<unsigned long long>(tmap)), "r"(col0), "r"(row0), "r"((int)b), "r"(smem_u32(bar))
        : "memory");
}
#endif

// progress counters the two helper threads share through shared memory
__device__ __forceinline__ long long ld_shared_volatile(const long long *p) { return *reinterpret_cast<const volatile long long *>(p); }
__device__ __forceinline__ void st_shared_volatile(long long *p, long long v) { *reinterpret_cast<volatile long long *>(p) = v; }

// two consecutive table entries with one 128-bit access where the type allows it
template <typename T>
__device__ __forceinline__ void ld_pair_global(const cx<T> *p, cx<T> &a, cx<T> &b) {
    if constexpr (sizeof(T) == 4) {
        const float4 f = __ldg(reinterpret_cast<const float4 *>(p));
        a = mk<T>(f.x, f.y); b = mk<T>(f.z, f.w);
    } else {
        a = ld_table(p); b = ld_table(p + 1);
    }
}
template <typename T>
__device__ __forceinline__ void ld_pair_shared(const cx<T> *p, cx<T> &a, cx<T> &b) {
    if constexpr (sizeof(T) == 4) {
        const float4 f = *reinterpret_cast<const float4 *>(p);
        a = mk<T>(f.x, f.y); b = mk<T>(f.z, f.w);
    } else {
        a = p[0]; b = p[1];
    }
}

// w[r] *= g^r for r < R, given pw2[k] = g^(2^k): the exponent is split into a low and a high half so that no power is
// more than two products away from the table values (R = 16: 11 products for 15 powers).
template <int R, typename T>
__device__ __forceinline__ void apply_powers(cx<T> (&w)[R], const cx<T> (&pw2)[flat_ilog2(R) > 0 ? flat_ilog2(R) : 1]) {
    constexpr int LOG = flat_ilog2(R), LO = (LOG + 1) / 2, NLO = 1 << LO, NHI = R >> LO;
    cx<T> lo[NLO], hi[NHI > 0 ? NHI : 1];
    sfor<1, NLO>([&](auto ic_) {
        constexpr int i = decltype(ic_)::value, hb = flat_topbit(i), rest = i - hb;
        if constexpr (rest == 0) lo[i] = pw2[flat_ilog2(hb)];
        else lo[i] = cmul(lo[rest], pw2[flat_ilog2(hb)]);
    });
    sfor<1, NHI>([&](auto ic_) {
        constexpr int i = decltype(ic_)::value, hb = flat_topbit(i), rest = i - hb;
        if constexpr (rest == 0) hi[i] = pw2[LO + flat_ilog2(hb)];
        else hi[i] = cmul(hi[rest], pw2[LO + flat_ilog2(hb)]);
    });
    sfor<1, R>([&](auto rc) {
        constexpr int r = decltype(rc)::value, l = r & (NLO - 1), h = r >> LO;
        if constexpr (l != 0 && h != 0) w[r] = cmul(w[r], cmul(hi[h], lo[l]));
        else if constexpr (l != 0) w[r] = cmul(w[r], lo[l]);
        else w[r] = cmul(w[r], hi[h]);
    });
}

// ---- column tile: CT adjacent columns, length-L FFT each, times W_N^(n2*k1), into the tile-major scratch.
// Pass p (radix R, P = product of the earlier radices, butterfly b: m' = b / P, racc = b % P) writes digit r of
// k1 = sum_p P_p r_p; its inter-pass twiddle W_L^(P m' r) and the factor W_N^(n2 P r) of the four-step twiddle are the
// powers g^r of g = W_(L/P)^(m') * W_(N/P)^(n2) (tables ga[p], gb[p]); the last pass multiplies by s4[n2][r].
template <typename Cfg, int INV, int N2C, int CTBLOG, bool INPLACE, typename T, typename Release>
__device__ __forceinline__ void flat_stage_a(const FlatParams<T> &q, const cx<T> *st, const cx<T> *sb, cx<T> *sm, cx<T> *scr, int lane0,
                                             int tid, Release release) {
    constexpr int L = Cfg::L, TX = Cfg::TX, CT = Cfg::CT, E = Cfg::E, PITCH = Cfg::PITCH, NC = Cfg::THREADS, NP = Cfg::NP;
    static_assert(NP == 2 || NP == 3, "two or three passes per tile");
    constexpr int CTB = 1 << CTBLOG;
    constexpr int RL = Cfg::radix(NP - 1), PL = Cfg::prod(NP - 1);
    static_assert(PL % CTB == 0 || CTB % PL == 0, "last-pass stride and scratch block height must nest");
    static_assert(RL % 2 == 0, "pairs of last-pass twiddles are loaded together");
    const int c = tid % CT, t = tid / CT;    // lanes along the columns (global / ring accesses)
    const int t2 = tid % TX, c2 = tid / TX;  // last pass: lanes along k1 (scratch written in runs of consecutive k1)
    cx<T> v[E];
    sfor<0, NP>([&](auto pc) {
        constexpr int ps = decltype(pc)::value;
        constexpr int R = Cfg::radix(ps), P = Cfg::prod(ps), NR = L / R, U = E / R, LOG = flat_ilog2(R);
        constexpr bool first = ps == 0, last = ps == NP - 1;
        static_assert((1 << LOG) == R && (P & (P - 1)) == 0, "power-of-two radices");
        const int tt = last ? t2 : t, cc = last ? c2 : c;
        if constexpr (first) {
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    const cx<T> x = st[(t + TX * u + NR * j) * CT + c];
                    v[u * R + j] = INV ? cswap(x) : x;
                }
            if constexpr (INPLACE) consumer_barrier(NC);  // the dense image is read: the padded image may overwrite it
            else release();                               // the ring slot may be refilled
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int j = 0; j < R; ++j) v[u * R + j] = sm[(tt + TX * u + NR * j) * PITCH + cc];
            if constexpr (!(INPLACE && last)) consumer_barrier(NC);  // everybody has read the exchange buffer: it may be overwritten
        }
        if constexpr (!last) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int b = tt + TX * u, mp = b / P, racc = b % P;
                cx<T> g[LOG];
                {
                    cx<T> ga[LOG], gb[LOG];
                    const cx<T> *pa = q.ga[ps] + mp * LOG, *pb = q.gb[ps] + (lane0 + cc) * LOG;
                    if constexpr (LOG % 2 == 0) {
#pragma unroll
                        for (int k = 0; k < LOG; k += 2) {
                            ld_pair_global(pa + k, ga[k], ga[k + 1]);
                            ld_pair_global(pb + k, gb[k], gb[k + 1]);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < LOG; ++k) { ga[k] = ld_table(pa + k); gb[k] = ld_table(pb + k); }
                    }
#pragma unroll
                    for (int k = 0; k < LOG; ++k) g[k] = cmul(ga[k], gb[k]);
                }
                cx<T> w[R];
#pragma unroll
                for (int j = 0; j < R; ++j) w[j] = v[u * R + j];
                if constexpr (!SSFFT_FLAT_NOCOMPUTE) {
                    Dft<R>::run(w);
                    apply_powers<R>(w, g);
                } else {
                    w[0] = w[0] + g[0];
                }
                const int o = racc + P * R * mp;
#pragma unroll
                for (int r = 0; r < R; ++r) sm[(o + P * r) * PITCH + cc] = w[r];
            }
            consumer_barrier(NC);
        } else {
            cx<T> s[R];
#pragma unroll
            for (int r = 0; r < R; r += 2) ld_pair_shared(sb + cc * R + r, s[r], s[r + 1]);
            if constexpr (INPLACE) release();  // tile and twiddle slice are in registers: the slot may be refilled
#pragma unroll
            for (int u = 0; u < U; ++u) {
                cx<T> w[R];
#pragma unroll
                for (int j = 0; j < R; ++j) w[j] = v[u * R + j];
                if constexpr (!SSFFT_FLAT_NOCOMPUTE) Dft<R>::run(w);
                const int b = tt + TX * u;  // k1 = b + P * r lives at block k1 / CTB, row k1 % CTB of the tile-major scratch
                if constexpr (P % CTB == 0) {
                    cx<T> *dst = scr + (long long)(b >> CTBLOG) * ((long long)CTB * N2C) + (long long)(lane0 + cc) * CTB + (b & (CTB - 1));
#pragma unroll
                    for (int r = 0; r < R; ++r) st_plain(dst + (long long)r * (P / CTB) * ((long long)CTB * N2C), SSFFT_FLAT_NOCOMPUTE ? w[r] + s[r] : cmul(w[r], s[r]));
                } else {
                    constexpr int Q = CTB / P;
                    cx<T> *dst = scr + (long long)(lane0 + cc) * CTB + b;
#pragma unroll
                    for (int r = 0; r < R; ++r) st_plain(dst + (long long)(r / Q) * ((long long)CTB * N2C) + P * (r % Q), SSFFT_FLAT_NOCOMPUTE ? w[r] + s[r] : cmul(w[r], s[r]));
                }
            }
        }
    });
}

// ---- row tile: CT adjacent rows k1 (one contiguous tile-major scratch block), length-L FFT each, stored transposed
template <typename Cfg, int INV, int N1C, bool INPLACE, bool TWSH, typename T, typename Release>
__device__ __forceinline__ void flat_stage_b(const cx<T> *twb, const cx<T> *st, cx<T> *sm, cx<T> *uout, int lane0, int tid,
                                             Release release) {
    constexpr int L = Cfg::L, TX = Cfg::TX, CT = Cfg::CT, E = Cfg::E, PITCH = Cfg::PITCH, NC = Cfg::THREADS, NP = Cfg::NP;
    static_assert(NP == 2 || NP == 3, "two or three passes per tile");
    const int c = tid % CT, t = tid / CT;
    cx<T> v[E];
    sfor<0, NP>([&](auto pc) {
        constexpr int ps = decltype(pc)::value;
        constexpr int R = Cfg::radix(ps), P = Cfg::prod(ps), NR = L / R, U = E / R, MN = Cfg::mnext(ps);
        constexpr bool first = ps == 0, last = ps == NP - 1;
        if constexpr (first) {
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int j = 0; j < R; ++j) v[u * R + j] = st[(t + TX * u + NR * j) * CT + c];
            if constexpr (INPLACE) consumer_barrier(NC);
            else release();
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int j = 0; j < R; ++j) v[u * R + j] = sm[(t + TX * u + NR * j) * PITCH + c];
            if constexpr (INPLACE && last) release();
            else consumer_barrier(NC);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int b = t + TX * u, mp = b / P, racc = b % P;
            cx<T> w[R];
#pragma unroll
            for (int j = 0; j < R; ++j) w[j] = v[u * R + j];
            if constexpr (!SSFFT_FLAT_NOCOMPUTE) Dft<R>::run(w);
            if constexpr (!last) {
                const cx<T> *twp = twb + Cfg::tw_off(ps) + mp;
                if constexpr (!SSFFT_FLAT_NOCOMPUTE) {
#pragma unroll
                    for (int r = 1; r < R; ++r) w[r] = cmul(w[r], TWSH ? twp[(r - 1) * MN] : ld_table(twp + (r - 1) * MN));
                }
                const int o = racc + P * R * mp;
#pragma unroll
                for (int r = 0; r < R; ++r) sm[(o + P * r) * PITCH + c] = w[r];
            } else {
                cx<T> *dst = uout + (lane0 + c) + (long long)N1C * b;  // k2 = b + P * r
#pragma unroll
                for (int r = 0; r < R; ++r) st_stream(dst + (long long)N1C * P * r, INV ? cswap(w[r]) : w[r]);
            }
        }
        if constexpr (!last) consumer_barrier(NC);
    });
}

// KIND 0: C2C.  (real flavours stay on the cluster kernel of tiled.cuh for now)
template <typename CfgA, typename CfgB, int INV, int NSTAGE, int MINB, bool INPLACE>
__global__ void __launch_bounds__(CfgA::THREADS + kFlatHelpers, MINB)
fourstep_flat_kernel(FlatParams<typename CfgA::T> q, const __grid_constant__ CUtensorMap tmap) {
    using T = typename CfgA::T;
    using Lay = FlatLayout<CfgA, CfgB, NSTAGE, INPLACE>;
    static_assert(CfgA::THREADS == CfgB::THREADS, "both stages must use the same CTA size");
    constexpr int NC = CfgA::THREADS;
    constexpr int N1 = CfgA::L, N2 = CfgB::L;
    constexpr int tiles1 = N2 / CfgA::CT, tiles2 = N1 / CfgB::CT, PT = tiles1 + tiles2;
    static_assert(N2 % CfgA::CT == 0 && N1 % CfgB::CT == 0, "whole tiles");
    constexpr int kCtbLog = flat_ilog2(CfgB::CT);
    static_assert((1 << kCtbLog) == CfgB::CT, "row-stage tile width must be a power of two");
    constexpr int NDONE = NSTAGE + 1;
    SSFFT_DYNAMIC_SMEM(ssfft_smem);
    cx<T> *exch = reinterpret_cast<cx<T> *>(ssfft_smem + Lay::oExch);
    cx<T> *twb_sm = reinterpret_cast<cx<T> *>(ssfft_smem + Lay::oTwB);
    const cx<T> *twb = Lay::kTwBShared ? twb_sm : q.tw_b;
    FlatDesc *desc = reinterpret_cast<FlatDesc *>(ssfft_smem + Lay::oDesc);  // [NSTAGE] ring, then [NDONE] producer history
    unsigned long long *full = reinterpret_cast<unsigned long long *>(ssfft_smem + Lay::oBars);
    unsigned long long *empty = full + NSTAGE, *done = empty + NSTAGE;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NC); }
        for (int s = 0; s < NDONE; ++s) mbar_init(&done[s], NC);
        long long *sync0 = reinterpret_cast<long long *>(ssfft_smem + Lay::oSync);
        sync0[0] = sync0[1] = sync0[2] = 0;
    }
    if constexpr (Lay::kTwBShared)
        for (int i = tid; i < CfgB::tw_total; i += NC + kFlatHelpers) twb_sm[i] = ld_table(q.tw_b + i);
    __syncthreads();

    // sync[0]: items the producer has issued; sync[1]: items the signaller has published; sync[2]: 1 = no more items
    long long *sync = reinterpret_cast<long long *>(ssfft_smem + Lay::oSync);
    FlatDesc *hist = desc + NSTAGE;
    unsigned *cnt1 = q.ctrl + 32, *cnt2 = cnt1 + q.cap;
    if (tid == NC + 32) {
        // ================= signaller (one thread) =================
        // When the consumers have finished an item (done mbarrier) it makes their scratch / output stores visible
        // device-wide (the fence waits for them to drain, ~1 us) and bumps the transform's counter.  Its own thread so
        // that this wait never delays the next copy.
        long long signaled = 0;
        long long t_idle = clock64();
        for (;;) {
            const long long issued = ld_shared_volatile(&sync[0]);
            if (signaled < issued) {
                const int k = (int)(signaled % NDONE);
                if (mbar_test(&done[k], (unsigned)((signaled / NDONE) & 1))) {
                    const FlatDesc h = hist[k];
                    __threadfence();
                    atomicAdd(h.kind == 0 ? &cnt1[h.b] : &cnt2[h.b], 1u);
                    ++signaled;
                    __threadfence_block();
                    st_shared_volatile(&sync[1], signaled);
                    producer_moved();
                    t_idle = clock64();
                    continue;
                }
            } else if (ld_shared_volatile(&sync[2]) != 0 && signaled >= ld_shared_volatile(&sync[0])) {
                break;
            }
            producer_idle();
            if (clock64() - t_idle > 8000000000LL) __trap();
        }
        return;
    }
    if (tid >= NC) {
        // ================= producer (one thread) =================
        if (tid != NC) return;
        const long long total = (q.batch + q.delay) * PT;
        long long issued = 0;
        bool have = false, exhausted = false, ready = false;
        FlatDesc cur{2, 0, 0, 0, 0};
        [[maybe_unused]] unsigned long long st_[kFlatStats] = {0};
        [[maybe_unused]] long long t_have = 0, t_ready = 0;
        long long t_idle = clock64();  // bounded waits: a scheduling surprise becomes a launch error, never a hung GPU
        for (;;) {
            bool moved = false;
            const long long signaled = ld_shared_volatile(&sync[1]);  // items whose history / done slots may be reused
            // a ticket is taken as late as the ring allows (ring of one: right away, its copy must start the moment the
            // slot frees; deeper rings: when a slot is free) -- tickets held idle widen the window of transforms in flight
            // and with it the scratch the schedule needs
            bool want = !have && !exhausted;
            if (want && NSTAGE > 1)
                want = issued - signaled <= NSTAGE &&
                       (issued < NSTAGE || mbar_test(&empty[issued % NSTAGE], (unsigned)(((issued / NSTAGE) - 1) & 1)));
            if (want) {
                [[maybe_unused]] const long long t0 = clock64();
                const long long tk = (long long)atomicAdd(q.ctrl, 1u);
                if constexpr (SSFFT_FLAT_STATS) { t_have = clock64(); st_[3] += t_have - t0; ++st_[4]; }
                ready = false;
                if (tk >= total) {
                    exhausted = true;
                    cur.kind = 2;
                    have = ready = true;
                } else {
                    const long long ph = tk / PT;
                    const int r = (int)(tk - ph * PT);
                    if (r < tiles1) { cur.kind = 0; cur.b = ph; cur.tile = r; }
                    else { cur.kind = 1; cur.b = ph - q.delay; cur.tile = r - tiles1; }
                    have = cur.b >= 0 && cur.b < q.batch;
                }
                moved = true;
            }
            if (have && !ready) {  // dependencies are polled while the ring slot is still busy: the copy starts the moment it frees
                if (SSFFT_FLAT_NODEPS) ready = true;
                else if (cur.kind == 0) ready = cur.b < q.nslots || ld_acquire_gpu(&cnt2[cur.b - q.nslots]) >= (unsigned)tiles2;
                else ready = ld_acquire_gpu(&cnt1[cur.b]) >= (unsigned)tiles1;
                if (ready) moved = true;
                if constexpr (SSFFT_FLAT_STATS) {
                    if (ready) { t_ready = clock64(); st_[cur.kind == 0 ? 5 : 6] += t_ready - t_have; }
                    else ++st_[7];  // polls that found the dependency open
                }
            }
            if (have && ready && issued - signaled <= NSTAGE) {
                const int s = (int)(issued % NSTAGE);
                const bool slot_free = issued < NSTAGE || mbar_test(&empty[s], (unsigned)(((issued / NSTAGE) - 1) & 1));
                if (slot_free) {
                    if (cur.kind == 2) {
                        desc[s] = cur;
                        mbar_arrive(&full[s]);
                        break;
                    }
                    {
                        if constexpr (SSFFT_FLAT_STATS) { cur.t_issue = clock64(); st_[8] += cur.t_issue - t_ready; }
                        desc[s] = cur;
                        hist[issued % NDONE] = cur;
                        cx<T> *slot = reinterpret_cast<cx<T> *>(ssfft_smem + Lay::oSlots + (size_t)s * Lay::kSlot);
                        if (cur.kind == 0) {
                            cx<T> *sblk = INPLACE ? reinterpret_cast<cx<T> *>(reinterpret_cast<unsigned char *>(slot) + Lay::kExch)
                                                  : reinterpret_cast<cx<T> *>(ssfft_smem + Lay::oSBlk + (size_t)(issued % NDONE) * Lay::kSBlk);
                            constexpr unsigned sbytes = (unsigned)(CfgA::CT * CfgA::radix(CfgA::NP - 1) * sizeof(cx<T>));
                            mbar_expect_tx(&full[s], (unsigned)Lay::kTileA + sbytes);
                            constexpr int kBoxRows = N1 > 256 ? 256 : N1;
#pragma unroll
                            for (int r0 = 0; r0 < N1; r0 += kBoxRows)
                                tma_tile_3d<T>(slot + (size_t)r0 * CfgA::CT, &tmap, q.in, N1, N2, kBoxRows, CfgA::CT, cur.tile * CfgA::CT, r0,
                                               cur.b, &full[s]);
                            bulk_g2s(sblk, q.s4 + (long long)cur.tile * CfgA::CT * CfgA::radix(CfgA::NP - 1), sbytes, &full[s]);
                        } else {
                            mbar_expect_tx(&full[s], (unsigned)Lay::kTileB);
                            fence_proxy_async();  // other CTAs' generic-proxy scratch stores -> async-proxy read
                            bulk_g2s(slot, q.scratch + (cur.b % q.nslots) * q.scratch_per + (long long)cur.tile * CfgB::CT * N2,
                                     (unsigned)Lay::kTileB, &full[s]);
                        }
                        ++issued;
                        __threadfence_block();  // the history entry is written before the signaller learns of the item
                        st_shared_volatile(&sync[0], issued);
                        have = false;
                        moved = true;
                    }
                }
            }
            if (moved) {
                producer_moved();
                t_idle = clock64();
            } else {
                producer_idle();
                if (clock64() - t_idle > 8000000000LL) __trap();
            }
        }
        __threadfence_block();
        st_shared_volatile(&sync[2], 1);  // the signaller drains the items still in flight and leaves
        if constexpr (SSFFT_FLAT_STATS)
            if (q.stats)
                for (int i = 3; i < 10; ++i) q.stats[(size_t)blockIdx.x * kFlatStats + i] = st_[i];
        return;
    }

    // ================= consumers =================
    [[maybe_unused]] unsigned long long cs_[kFlatStats] = {0};
    [[maybe_unused]] const long long t_begin = clock64();
    for (long long j = 0;; ++j) {
        const int s = (int)(j % NSTAGE);
        [[maybe_unused]] bool waited = false;
        [[maybe_unused]] long long t0 = 0;
        if constexpr (SSFFT_FLAT_STATS) {
            waited = !mbar_test(&full[s], (unsigned)((j / NSTAGE) & 1));
            t0 = clock64();
        }
        mbar_wait(&full[s], (unsigned)((j / NSTAGE) & 1));
        const FlatDesc d = desc[s];
        if constexpr (SSFFT_FLAT_STATS) {
            const long long t1 = clock64();
            if (d.kind != 2) {
                cs_[1] += t1 - t0;
                ++cs_[2];
                if (waited) { cs_[d.kind == 0 ? 10 : 12] += t1 - d.t_issue; ++cs_[d.kind == 0 ? 11 : 13]; }
            }
        }
        if (d.kind == 2) break;
        cx<T> *st = reinterpret_cast<cx<T> *>(ssfft_smem + Lay::oSlots + (size_t)s * Lay::kSlot);
        cx<T> *xb = INPLACE ? st : exch;  // exchange buffer of this tile
        auto release = [&]() { mbar_arrive(&empty[s]); };
        cx<T> *scr = q.scratch + (d.b % q.nslots) * q.scratch_per;
        if (d.kind == 0) {
            const cx<T> *sblk = INPLACE ? reinterpret_cast<const cx<T> *>(reinterpret_cast<const unsigned char *>(st) + Lay::kExch)
                                        : reinterpret_cast<const cx<T> *>(ssfft_smem + Lay::oSBlk + (size_t)(j % NDONE) * Lay::kSBlk);
            flat_stage_a<CfgA, INV, N2, kCtbLog, INPLACE>(q, st, sblk, xb, scr, d.tile * CfgA::CT, tid, release);
        } else {
            if (q.discard) {  // the block is in shared memory now: drop its lines from L2 without a write-back
                constexpr int kLines = (int)(Lay::kTileB / 128);
                const char *blk = reinterpret_cast<const char *>(scr + (long long)d.tile * CfgB::CT * N2);
                for (int i = tid; i < kLines; i += NC) discard_l2_line(blk + (size_t)i * 128);
            }
            flat_stage_b<CfgB, INV, N1, INPLACE, Lay::kTwBShared>(twb, st, xb, q.out + d.b * q.user_stride, d.tile * CfgB::CT, tid, release);
        }
        mbar_arrive(&done[j % NDONE]);
    }
    if constexpr (SSFFT_FLAT_STATS)
        if (q.stats && tid == 0) {
            unsigned long long *o = q.stats + (size_t)blockIdx.x * kFlatStats;
            o[0] = clock64() - t_begin; o[1] = cs_[1]; o[2] = cs_[2];
            for (int i = 10; i < 14; ++i) o[i] = cs_[i];
        }
}

#endif  // __CUDACC__

// ---------------------------------------------------------------------------------------------
// host side: tables, registry
// ---------------------------------------------------------------------------------------------
struct FlatEntry {
    int prec, n1, n2;
    const char *name;
    int ra[3], na_passes, cta, ctb;  // column-stage radices / lanes, row-stage lanes
    int threads, nstage, minb, inplace;
    size_t smem_bytes;
    int tile_b_tw;  // entries of the row-stage pass table
    int rb[3], nb_passes;
    int (*launch[2])(const void *params, int ctas, cudaStream_t s);  // [inverse]; params: FlatParams<T>; 3 = no tensor map
    int (*max_ctas[2])();                                            // co-resident CTAs on the current device
};
const std::vector<FlatEntry> &flat_registry();

// first registered entry of the size, or the variant named by SSFFT_FLAT_VARIANT="ring,ctas_per_sm[,inplace]"
template <typename T>
inline int find_flat(size_t n1, size_t n2) {
    const int prec = sizeof(T) == 4 ? 0 : 1;
    const auto &reg = flat_registry();
    int want_ring = 0, want_minb = 0, want_inplace = 1;
    if (const char *e = getenv("SSFFT_FLAT_VARIANT")) sscanf(e, "%d,%d,%d", &want_ring, &want_minb, &want_inplace);
    int first = -1;
    for (size_t i = 0; i < reg.size(); ++i)
        if (reg[i].prec == prec && (size_t)reg[i].n1 == n1 && (size_t)reg[i].n2 == n2) {
            if (first < 0) first = (int)i;
            if (reg[i].nstage == want_ring && reg[i].minb == want_minb && reg[i].inplace == want_inplace) return (int)i;
        }
    return first;
}

template <typename T>
inline void flat_root(T *dst, unsigned long long num, unsigned long long den) {
    num %= den;
    const long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)num / (long double)den;
    dst[0] = (T)cosl(a);
    dst[1] = (T)(-sinl(a));
}
// Tables of the column stage (interleaved re, im), radices ra[0 .. passes-1] of the length-n1 transform:
//   per non-last pass p (P = product of the earlier radices, R = ra[p], LOG = log2 R):
//     ga[p][m'][k] = W_(n1/P)^(m' * 2^k), m' < n1 / (P R);      gb[p][c][k] = W_(n/P)^(c * 2^k), c < n2
//   s4[c][r] = W_n^(c * P_last * r), r < R_last
template <typename T>
inline void fill_flat_tables(std::vector<T> (&ga)[2], std::vector<T> (&gb)[2], std::vector<T> &s4, size_t n1, size_t n2, const int *ra,
                             int passes) {
    const size_t n = n1 * n2;
    size_t P = 1;
    for (int p = 0; p + 1 < passes; ++p) {
        const int R = ra[p], lg = flat_ilog2(R);
        const size_t mcount = n1 / (P * R);
        ga[p].assign(2 * mcount * lg, (T)0);
        gb[p].assign(2 * n2 * lg, (T)0);
        for (size_t m = 0; m < mcount; ++m)
            for (int k = 0; k < lg; ++k) flat_root<T>(&ga[p][2 * (m * lg + k)], (unsigned long long)m << k, n1 / P);
        for (size_t c = 0; c < n2; ++c)
            for (int k = 0; k < lg; ++k) flat_root<T>(&gb[p][2 * (c * lg + k)], (unsigned long long)c << k, n / P);
        P *= R;
    }
    const int RL = ra[passes - 1];
    s4.assign(2 * n2 * RL, (T)0);
    for (size_t c = 0; c < n2; ++c)
        for (int r = 0; r < RL; ++r) flat_root<T>(&s4[2 * (c * RL + r)], (unsigned long long)c * P * r, n);
}
// row-stage pass table, same layout as build_tile_twiddles ([r-1][m'] per pass)
template <typename T>
inline void fill_flat_row_twiddles(std::vector<T> &h, int n2, const int *radix, int passes, int tw_total) {
    h.assign(2 * (size_t)(tw_total > 0 ? tw_total : 1), (T)0);
    size_t o = 0;
    int P = 1;
    for (int p = 0; p + 1 < passes; ++p) {
        const int R = radix[p], MN = n2 / (P * R);
        for (int r = 1; r < R; ++r)
            for (int m = 0; m < MN; ++m) { flat_root<T>(&h[2 * o], (unsigned long long)P * m * r, (unsigned long long)n2); ++o; }
        P *= R;
    }
}

}  // namespace ssfft
