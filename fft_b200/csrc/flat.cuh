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
#include <cstring>

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
    // real transforms (RealFFT<V>::fft / ifft, signalsmith-fft.h:446-502, on the complex length M = N1 * N2, N = 2 M):
    //   R2C post-twiddle  -i W_N^k / 2,  k = k1 + N1 k2:   ra[k1] = W_N^k1 (k1 < N1),  rb[k2] = -i W_(2 N2)^k2 / 2 (k2 < N2)
    //   C2R pre-twiddle  conj(-i W_N^n), n = n1 N2 + n2:   ra[n1] = i conj(W_(2 N1)^n1) (n1 < N1),  rb[n2] = conj(W_N^n2) (n2 < N2)
    const cx<T> *ra, *rb;
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
// rows per TMA box of a column tile: the largest divisor of the column length that a box can have (at most 256)
__host__ __device__ constexpr int flat_box_rows(int n1) {
    int r = n1 > 256 ? 256 : n1;
    while (n1 % r) --r;
    return r;
}

// shared-memory map of a CTA (bytes).  INPLACE: a ring slot is also the exchange buffer of the tile it holds (the TMA
// copy lands dense at its start, the passes overwrite it with the padded image, the tile's slice of s4 sits behind it),
// so a ring of two fits three CTAs per SM; otherwise one separate exchange buffer, dense slots and a small ring of s4
// slices.  The row-stage pass table lives in shared memory when it is small (two-pass row tiles), else it is read
// through L1 like the column-stage tables.
template <typename CfgA, typename CfgB, int NSTAGE, bool INPLACE, int KIND = 0>
struct FlatLayout {
    using T = typename CfgA::T;
    static constexpr size_t al(size_t v) { return (v + 127) / 128 * 128; }
    // column tile: dense [L][CT].  C2R (H = CT/2): the low columns [L][H], their partners N2 - n2 in a box [L][H + 2] (a
    // tensor copy must start on a 16-byte boundary = an even column of 8-byte elements, the partners of an even-aligned
    // run start on an odd one: the box starts one column early and is two columns wider), and for tile 0 the self-paired
    // column N2/2 as the first column of a third box [L][H]
    static constexpr size_t kTileA = (size_t)CfgA::L * CfgA::CT * sizeof(cx<T>);
    static constexpr size_t kC2rLow = (size_t)CfgA::L * (CfgA::CT / 2) * sizeof(cx<T>);
    static constexpr size_t kC2rHigh = (size_t)CfgA::L * (CfgA::CT / 2 + 2) * sizeof(cx<T>);
    static constexpr size_t kSlotA = KIND == 2 ? 2 * kC2rLow + kC2rHigh : kTileA;
    // row tile: dense [L][CT]; R2C: two blocks [L][CT/2], the second 64 bytes further
    static constexpr size_t kTileB = (size_t)CfgB::L * CfgB::CT * sizeof(cx<T>);
    static constexpr size_t kSlotB = KIND == 1 ? kTileB + 2 * 8 * sizeof(cx<T>) : kTileB;
    static constexpr size_t kExchA = CfgA::smem_bytes, kExchB = CfgB::smem_bytes;
    static constexpr size_t kExch = al(kExchA > kExchB ? kExchA : kExchB);
    static constexpr size_t kSBlk = KIND == 2 ? 0 : al((size_t)CfgA::CT * CfgA::radix(CfgA::NP - 1) * sizeof(cx<T>));
    static constexpr size_t kLanded = al(kSlotA > kSlotB ? kSlotA : kSlotB);  // what the copies of one tile bring
    static constexpr size_t kImage = kExch > kLanded ? kExch : kLanded;       // in place: landed tile and exchange buffer share it
    static constexpr size_t kSlot = INPLACE ? kImage + kSBlk : kLanded;
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
// a TMA box may start at any column and hang over the edge of the tensor (zero fill): element-wise on the CPU
template <typename T>
inline void tma_tile_3d(cx<T> *dst, const void *, const cx<T> *in, int n1, int n2, int box_rows, int ct, int col0, int row0, long long b,
                        unsigned long long *bar) {
    // what the hardware requires and a CPU does not notice: box start on a 16-byte boundary in global memory, box
    // destination on a 128-byte boundary in shared memory, inner box extent a multiple of 16 bytes
    if ((col0 * sizeof(cx<T>)) % 16 || (ct * sizeof(cx<T>)) % 16 || reinterpret_cast<uintptr_t>(dst) % 128) {
        fprintf(stderr, "tensor copy violates the TMA alignment rules: col0 %d, box cols %d, dst %p\n", col0, ct, (void *)dst);
        abort();
    }
    for (int r = 0; r < box_rows; ++r)
        for (int c = 0; c < ct; ++c) {
            if (col0 + c < n2) simt::tma_copy(dst + (size_t)r * ct + c, in + b * (long long)n1 * n2 + (long long)(row0 + r) * n2 + col0 + c, (unsigned)sizeof(cx<T>), bar);
            else simt::tma_zero(dst + (size_t)r * ct + c, (unsigned)sizeof(cx<T>), bar);
        }
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
// box [box_rows][ct] at (col0, row0) of transform b of the (batch, N1, N2) input (tensor map built by the launcher; its
// elements are 8-byte words: a double-precision complex value is two of them)
template <typename T>
__device__ __forceinline__ void tma_tile_3d(cx<T> *dst, const void *tmap, const cx<T> *, int, int, int, int, int col0, int row0,
                                            long long b, unsigned long long *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<unsigned long long>(tmap)), "r"(col0 * (int)(sizeof(cx<T>) / 8)), "r"(row0), "r"((int)b), "r"(smem_u32(bar))
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
// KIND 0: complex.  KIND 1 (R2C): the same column FFT of the sample pairs, but row k1 of the scratch is stored at position
// rho(k1) (flat_rho) so that the rows the post-twiddle pairs, k1 and N1 - k1, sit in two blocks of CTB rows that one row
// tile loads.  KIND 2 (C2R): the tile is CT/2 low columns and their CT/2 partner columns N2 - n2 (two TMA boxes); the
// pre-twiddle of RealFFT::ifft (:478-492) is applied while the first pass gathers; `tile` is the tile index.
template <typename Cfg, int INV, int N2C, int CTBLOG, bool INPLACE, int KIND, typename T, typename Release>
__device__ __forceinline__ void flat_stage_a(const FlatParams<T> &q, const cx<T> *st, const cx<T> *sb, cx<T> *sm, cx<T> *scr, int tile,
                                             int tid, Release release) {
    const int lane0 = tile * Cfg::CT;
    constexpr int L = Cfg::L, TX = Cfg::TX, CT = Cfg::CT, E = Cfg::E, PITCH = Cfg::PITCH, NC = Cfg::THREADS, NP = Cfg::NP;
    static_assert(NP == 2 || NP == 3, "two or three passes per tile");
    constexpr int CTB = 1 << CTBLOG;
    constexpr int RL = Cfg::radix(NP - 1), PL = Cfg::prod(NP - 1);
    static_assert(PL % CTB == 0 || CTB % PL == 0, "last-pass stride and scratch block height must nest");
    static_assert(RL % 2 == 0, "pairs of last-pass twiddles are loaded together");
    const int c = tid % CT, t = tid / CT;    // lanes along the columns (global / ring accesses)
    const int t2 = tid % TX, c2 = tid / TX;  // last pass: lanes along k1 (scratch written in runs of consecutive k1)
    // column of a lane: contiguous, or (C2R) low half / mirrored high half, lane CT-1 of tile 0 = the self-paired N2/2
    auto col_of = [&](int lane) -> int {
        if constexpr (KIND != 2) return lane0 + lane;
        else {
            constexpr int H = CT / 2;
            if (lane < H) return tile * H + lane;
            if (tile == 0 && lane == CT - 1) return N2C / 2;
            return N2C - tile * H - (CT - 1 - lane);
        }
    };
    cx<T> v[E];
    sfor<0, NP>([&](auto pc) {
        constexpr int ps = decltype(pc)::value;
        constexpr int R = Cfg::radix(ps), P = Cfg::prod(ps), NR = L / R, U = E / R, LOG = flat_ilog2(R);
        constexpr bool first = ps == 0, last = ps == NP - 1;
        // every pass but the last takes its twiddles as powers of one number (tables per bit of the digit): powers of
        // two there; the last pass multiplies by a table row and may carry a factor 3 (column lengths 3 * 2^j)
        static_assert(last || (1 << LOG) == R, "power-of-two radices before the last pass");
        static_assert((P & (P - 1)) == 0, "the digits below the last one are powers of two");
        const int tt = last ? t2 : t, cc = last ? c2 : c;
        if constexpr (first && KIND == 2) {
            // ring slot: low columns [L][H], partner box [L][H + 2] (column N2 - n2 of low column n2 = H tile + i sits at
            // box column H - i), then for tile 0 the box whose first column is N2/2
            constexpr int H = CT / 2, HP = H + 2, LOW = L * H, HIGH = L * HP;
            const bool lane_self0 = tile == 0 && c == 0, lane_selfm = tile == 0 && c == CT - 1;
            auto where = [&](int lane, int &pitch) -> const cx<T> * {
                if (lane < H) { pitch = H; return st + lane; }
                if (tile == 0 && lane == CT - 1) { pitch = H; return st + LOW + HIGH; }
                pitch = HP;
                return st + LOW + (lane - H + 1);
            };
            int own_pitch, par_pitch;
            const cx<T> *own = where(c, own_pitch);
            const cx<T> *par = where(CT - 1 - c, par_pitch);
            if (lane_self0 || lane_selfm) { par = own; par_pitch = own_pitch; }
            const cx<T> cb = ld_table(q.rb + col_of(c));
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    const int n1 = t + TX * u + NR * j;
                    const int n1p = lane_self0 ? (n1 ? L - n1 : 0) : L - 1 - n1;
                    const cx<T> xo = own[n1 * own_pitch], xp = par[n1p * par_pitch];
                    const cx<T> tc = cmul(ld_table(q.ra + n1), cb);  // conj(-i W_N^n)
                    const cx<T> sum = mk<T>(xo.x + xp.x, xo.y - xp.y), dif = mk<T>(xo.x - xp.x, xo.y + xp.y);  // X +- conj X'
                    cx<T> z = sum + cmul(dif, tc);
                    if (lane_self0 && n1 == 0) z = mk<T>(xo.x + xo.y, xo.x - xo.y);  // bin 0 packs (DC, Nyquist) (:478-481)
                    v[u * R + j] = cswap(z);  // inverse transform = forward transform of the swapped data
                }
            if constexpr (INPLACE) consumer_barrier(NC);  // the landed boxes are read: the padded image may overwrite them
            else release();
        } else if constexpr (first) {
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
                    const cx<T> *pa = q.ga[ps] + mp * LOG, *pb = q.gb[ps] + col_of(cc) * LOG;
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
            const int n2 = col_of(cc);
            if constexpr (KIND == 2) {  // scattered columns: the slice of s4 comes through L1 instead of the ring
#pragma unroll
                for (int r = 0; r < R; r += 2) ld_pair_global(q.s4 + (long long)n2 * R + r, s[r], s[r + 1]);
            } else {
#pragma unroll
                for (int r = 0; r < R; r += 2) ld_pair_shared(sb + cc * R + r, s[r], s[r + 1]);
            }
            if constexpr (INPLACE) release();  // tile and twiddle slice are in registers: the slot may be refilled
#pragma unroll
            for (int u = 0; u < U; ++u) {
                cx<T> w[R];
#pragma unroll
                for (int j = 0; j < R; ++j) w[j] = v[u * R + j];
                if constexpr (!SSFFT_FLAT_NOCOMPUTE) Dft<R>::run(w);
                const int b = tt + TX * u;  // k1 = b + P * r lives at block k1 / CTB, row k1 % CTB of the tile-major scratch
                if constexpr (KIND == 1) {
                    // rho(k1): k1 < L/2 stays, k1 > L/2 moves down one row, L/2 goes to the last row (flat_rho).  k1 < L/2
                    // exactly for r < R/2, so both halves are base + immediate; (b, r) = (0, R/2) is the one exception.
                    static_assert(P % CTB == 0 && R % 2 == 0, "scratch blocks must nest in the last-pass stride");
                    const long long blk = (long long)CTB * N2C;
                    cx<T> *lo = scr + (long long)(b >> CTBLOG) * blk + (long long)n2 * CTB + (b & (CTB - 1));
                    const int bm = b - 1;  // arithmetic shift / mask: floor semantics for b = 0
                    cx<T> *hi = scr + (long long)(bm >> CTBLOG) * blk + (long long)n2 * CTB + (bm & (CTB - 1));
                    cx<T> *mid = scr + (long long)((L - 1) >> CTBLOG) * blk + (long long)n2 * CTB + ((L - 1) & (CTB - 1));
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const cx<T> val = cmul(w[r], s[r]);
                        cx<T> *dst = (r < R / 2 ? lo : hi) + (long long)r * (P / CTB) * blk;
                        if (r == R / 2 && b == 0) dst = mid;
                        st_plain(dst, val);
                    }
                } else if constexpr (P % CTB == 0) {
                    cx<T> *dst = scr + (long long)(b >> CTBLOG) * ((long long)CTB * N2C) + (long long)n2 * CTB + (b & (CTB - 1));
#pragma unroll
                    for (int r = 0; r < R; ++r) st_plain(dst + (long long)r * (P / CTB) * ((long long)CTB * N2C), SSFFT_FLAT_NOCOMPUTE ? w[r] + s[r] : cmul(w[r], s[r]));
                } else {
                    constexpr int Q = CTB / P;
                    cx<T> *dst = scr + (long long)n2 * CTB + b;
#pragma unroll
                    for (int r = 0; r < R; ++r) st_plain(dst + (long long)(r / Q) * ((long long)CTB * N2C) + P * (r % Q), SSFFT_FLAT_NOCOMPUTE ? w[r] + s[r] : cmul(w[r], s[r]));
                }
            }
        }
    });
}

// Position of row k1 of the R2C intermediate in the scratch (n1 rows): the post-twiddle pairs row k1 with row n1 - k1, so
// the upper half is stored one row down (block [n1 - HB t - HB, n1 - HB t) then holds exactly the partners of block t's
// rows) and the self-paired row n1/2 takes the slot that frees at the very end.
__host__ __device__ constexpr int flat_rho(int k1, int n1) { return k1 < n1 / 2 ? k1 : k1 == n1 / 2 ? n1 - 1 : k1 - 1; }

// ---- row tile of a real forward transform: HB = CT/2 rows k1 = HB tile + l and their partners n1 - k1 (two scratch
// blocks of HB rows), length-L FFT each, then RealFFT::fft's post-twiddle (:459-472) on the pairs
// (k1, k2) <-> (n1 - k1, L - 1 - k2) through the exchange buffer, stored as bins k = k1 + n1 k2 of the half spectrum.
template <typename Cfg, int N1C, bool INPLACE, bool TWSH, typename T, typename Release>
__device__ __forceinline__ void flat_stage_b_r2c(const FlatParams<T> &q, const cx<T> *twb, const cx<T> *st, cx<T> *sm, cx<T> *uout,
                                                 int tile, int tid, Release release) {
    constexpr int L = Cfg::L, TX = Cfg::TX, CT = Cfg::CT, E = Cfg::E, PITCH = Cfg::PITCH, NC = Cfg::THREADS, NP = Cfg::NP;
    constexpr int H = CT / 2, HALF = L * H + 8;  // ring slot: [2][L][H], the high block 8 elements further (other banks)
    static_assert(NP == 2 || NP == 3, "two or three passes per tile");
    const int c = tid % CT, t = tid / CT;
    cx<T> v[E];
    sfor<0, NP>([&](auto pc) {
        constexpr int ps = decltype(pc)::value;
        constexpr int R = Cfg::radix(ps), P = Cfg::prod(ps), NR = L / R, U = E / R, MN = Cfg::mnext(ps);
        constexpr bool first = ps == 0, last = ps == NP - 1;
        if constexpr (first) {
            const cx<T> *src = st + (c / H) * HALF + (c % H);
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int j = 0; j < R; ++j) v[u * R + j] = src[(t + TX * u + NR * j) * H];
            if constexpr (INPLACE) consumer_barrier(NC);  // both landed blocks are read: the padded image may overwrite them
            else release();
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int j = 0; j < R; ++j) v[u * R + j] = sm[(t + TX * u + NR * j) * PITCH + c];
            consumer_barrier(NC);
        }
        if constexpr (!last) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int b = t + TX * u, mp = b / P, racc = b % P;
                cx<T> w[R];
#pragma unroll
                for (int j = 0; j < R; ++j) w[j] = v[u * R + j];
                Dft<R>::run(w);
                const cx<T> *twp = twb + Cfg::tw_off(ps) + mp;
#pragma unroll
                for (int r = 1; r < R; ++r) w[r] = cmul(w[r], TWSH ? twp[(r - 1) * MN] : ld_table(twp + (r - 1) * MN));
                const int o = racc + P * R * mp;
#pragma unroll
                for (int r = 0; r < R; ++r) sm[(o + P * r) * PITCH + c] = w[r];
            }
            consumer_barrier(NC);
        } else {
            static_assert(R % 2 == 0, "the pairs split the last-pass outputs in halves");
            // Z[k2][lane] into the exchange buffer, k2 = b + P r
#pragma unroll
            for (int u = 0; u < U; ++u) {
                cx<T> w[R];
#pragma unroll
                for (int j = 0; j < R; ++j) w[j] = v[u * R + j];
                Dft<R>::run(w);
                const int b = t + TX * u;
#pragma unroll
                for (int r = 0; r < R; ++r) { sm[(b + P * r) * PITCH + c] = w[r]; v[u * R + r] = w[r]; }
            }
            consumer_barrier(NC);
            // my rows: lane -> k1; pairs (k1, k2 < L/2) <-> (n1 - k1, L - 1 - k2) are handled by the owner of the first element
            const bool row0 = tile == 0 && c == 0, rowm = tile == 0 && c == CT - 1;
            const int k1 = c < H ? tile * H + c : rowm ? N1C / 2 : N1C - tile * H - (CT - 1 - c);
            const int pl = (row0 || rowm) ? c : CT - 1 - c;  // partner lane (rows 0 and n1/2 pair with themselves)
            const cx<T> ta = ld_table(q.ra + k1);
            cx<T> zp[U * (R / 2)];
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int r = 0; r < R / 2; ++r) {
                    const int k2 = t + TX * u + P * r;
                    const int kp = row0 ? (k2 ? L - k2 : 0) : L - 1 - k2;
                    zp[u * (R / 2) + r] = sm[kp * PITCH + pl];
                }
            cx<T> zmid = mk<T>((T)0, (T)0);
            if (row0 && t == 0) zmid = sm[(L / 2) * PITCH];  // bin M/2 pairs with itself
            // partners are in registers: the next tile may overwrite the exchange buffer (in place: the slot may be refilled,
            // and no consumer touches it again before that copy has landed)
            if constexpr (INPLACE) release();
            else consumer_barrier(NC);
            const int kph = k1 ? N1C - k1 : 0;  // partner row (0 for row 0)
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int r = 0; r < R / 2; ++r) {
                    const int k2 = t + TX * u + P * r;
                    const cx<T> zm = v[u * R + r], zq = zp[u * (R / 2) + r];
                    if (row0 && k2 == 0) {  // (DC, Nyquist) packed in bin 0 (:459-462)
                        st_stream(uout, mk<T>(zm.x + zm.y, zm.x - zm.y));
                        continue;
                    }
                    const cx<T> tw = cmul(ta, ld_table(q.rb + k2));  // -i W_N^k / 2
                    const cx<T> sum = mk<T>(zm.x + zq.x, zm.y - zq.y), dif = mk<T>(zm.x - zq.x, zm.y + zq.y);  // Z +- conj Z'
                    const cx<T> rot = cmul(dif, tw), hs = mk<T>((T)0.5 * sum.x, (T)0.5 * sum.y);
                    const int kp = row0 ? L - k2 : L - 1 - k2;
                    st_stream(uout + k1 + (long long)N1C * k2, hs + rot);
                    st_stream(uout + kph + (long long)N1C * kp, mk<T>(hs.x - rot.x, rot.y - hs.y));  // conj(hs - rot)
                }
            if (row0 && t == 0) {  // bin M/2 (k1 = 0, k2 = L/2): its own partner
                const cx<T> tw = cmul(ta, ld_table(q.rb + L / 2));
                const cx<T> sum = mk<T>(zmid.x + zmid.x, (T)0), dif = mk<T>((T)0, zmid.y + zmid.y);
                const cx<T> rot = cmul(dif, tw), hs = mk<T>((T)0.5 * sum.x, (T)0.5 * sum.y);
                st_stream(uout + (long long)N1C * (L / 2), mk<T>(hs.x - rot.x, rot.y - hs.y));
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
// KIND 0: complex (INV = direction).  KIND 1: RealFFT forward on the sample pairs (INV = 0).  KIND 2: RealFFT inverse
// (INV = 1).  tmap: boxes [box rows][CT] (KIND 2: [box rows][CT/2]) of the (batch, N1, N2) input; tmap2 (KIND 2 only):
// boxes [box rows][2], for the self-paired column N2/2.
template <typename CfgA, typename CfgB, int INV, int NSTAGE, int MINB, bool INPLACE, int KIND = 0>
__global__ void __launch_bounds__(CfgA::THREADS + kFlatHelpers, MINB)
fourstep_flat_kernel(FlatParams<typename CfgA::T> q, const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap2) {
    using T = typename CfgA::T;
    using Lay = FlatLayout<CfgA, CfgB, NSTAGE, INPLACE, KIND>;
    static_assert(KIND == 0 || INV == (KIND == 2), "real flavours: fixed direction");
    static_assert(CfgA::THREADS == CfgB::THREADS, "both stages must use the same CTA size");
    constexpr int NC = CfgA::THREADS;
    constexpr int N1 = CfgA::L, N2 = CfgB::L;
    constexpr int tiles1 = N2 / CfgA::CT, tiles2 = N1 / CfgB::CT, PT = tiles1 + tiles2;
    static_assert(N2 % CfgA::CT == 0 && N1 % CfgB::CT == 0, "whole tiles");
    constexpr int kCtb = KIND == 1 ? CfgB::CT / 2 : CfgB::CT;  // rows per scratch block (R2C: half a tile, see flat_rho)
    constexpr int kCtbLog = flat_ilog2(kCtb);
    static_assert((1 << kCtbLog) == kCtb, "row-stage tile width must be a power of two");
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
                        constexpr int kBoxRows = flat_box_rows(N1);
                        if (cur.kind == 0 && KIND == 2) {
                            // low columns [H t, H t + H); their partners [N2 - H t - H + 1, N2 - H t] inside the box of H + 2
                            // columns that starts at the even column N2 - H t - H (for t = 0 its last two columns are outside
                            // the tensor: zero fill, unused); for t = 0 also the box that starts at the self-paired column N2/2
                            constexpr int H = CfgA::CT / 2, HP = H + 2, LOW = N1 * H, HIGH = N1 * HP;
                            mbar_expect_tx(&full[s], (unsigned)(Lay::kC2rLow + Lay::kC2rHigh + (cur.tile == 0 ? Lay::kC2rLow : 0)));
#pragma unroll
                            for (int r0 = 0; r0 < N1; r0 += kBoxRows) {
                                tma_tile_3d<T>(slot + (size_t)r0 * H, &tmap, q.in, N1, N2, kBoxRows, H, cur.tile * H, r0, cur.b, &full[s]);
                                tma_tile_3d<T>(slot + LOW + (size_t)r0 * HP, &tmap2, q.in, N1, N2, kBoxRows, HP, N2 - cur.tile * H - H, r0, cur.b,
                                               &full[s]);
                                if (cur.tile == 0)
                                    tma_tile_3d<T>(slot + LOW + HIGH + (size_t)r0 * H, &tmap, q.in, N1, N2, kBoxRows, H, N2 / 2, r0, cur.b, &full[s]);
                            }
                        } else if (cur.kind == 0) {
                            cx<T> *sblk = INPLACE ? reinterpret_cast<cx<T> *>(reinterpret_cast<unsigned char *>(slot) + Lay::kImage)
                                                  : reinterpret_cast<cx<T> *>(ssfft_smem + Lay::oSBlk + (size_t)(issued % NDONE) * Lay::kSBlk);
                            constexpr unsigned sbytes = (unsigned)(CfgA::CT * CfgA::radix(CfgA::NP - 1) * sizeof(cx<T>));
                            mbar_expect_tx(&full[s], (unsigned)Lay::kTileA + sbytes);
#pragma unroll
                            for (int r0 = 0; r0 < N1; r0 += kBoxRows)
                                tma_tile_3d<T>(slot + (size_t)r0 * CfgA::CT, &tmap, q.in, N1, N2, kBoxRows, CfgA::CT, cur.tile * CfgA::CT, r0,
                                               cur.b, &full[s]);
                            bulk_g2s(sblk, q.s4 + (long long)cur.tile * CfgA::CT * CfgA::radix(CfgA::NP - 1), sbytes, &full[s]);
                        } else if (KIND == 1) {
                            // scratch blocks of H rows: block t (rows H t ...) and block N1/H - 1 - t (their partners, flat_rho)
                            constexpr int H = CfgB::CT / 2, HALF = N2 * H + 8;
                            constexpr unsigned half_bytes = (unsigned)(N2 * H * sizeof(cx<T>));
                            mbar_expect_tx(&full[s], 2 * half_bytes);
                            fence_proxy_async();
                            const cx<T> *scr = q.scratch + (cur.b % q.nslots) * q.scratch_per;
                            bulk_g2s(slot, scr + (long long)cur.tile * H * N2, half_bytes, &full[s]);
                            bulk_g2s(slot + HALF, scr + (long long)(N1 / H - 1 - cur.tile) * H * N2, half_bytes, &full[s]);
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
            const cx<T> *sblk = INPLACE ? reinterpret_cast<const cx<T> *>(reinterpret_cast<const unsigned char *>(st) + Lay::kImage)
                                        : reinterpret_cast<const cx<T> *>(ssfft_smem + Lay::oSBlk + (size_t)(j % NDONE) * Lay::kSBlk);
            flat_stage_a<CfgA, INV, N2, kCtbLog, INPLACE, KIND>(q, st, sblk, xb, scr, d.tile, tid, release);
        } else if constexpr (KIND == 1) {
            if (q.discard) {
                constexpr int H = CfgB::CT / 2, kLines = (int)(N2 * H * sizeof(cx<T>) / 128);
                const char *lo = reinterpret_cast<const char *>(scr + (long long)d.tile * H * N2);
                const char *hi = reinterpret_cast<const char *>(scr + (long long)(N1 / H - 1 - d.tile) * H * N2);
                for (int i = tid; i < 2 * kLines; i += NC) discard_l2_line((i < kLines ? lo : hi - (size_t)kLines * 128) + (size_t)i * 128);
            }
            flat_stage_b_r2c<CfgB, N1, INPLACE, Lay::kTwBShared>(q, twb, st, xb, q.out + d.b * q.user_stride, d.tile, tid, release);
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
    // RealFFT of length 2 n1 n2 on the same tiles (entries with a separate exchange buffer only): [0] forward, [1] inverse
    int (*launch_real[2])(const void *params, int ctas, cudaStream_t s);
    int (*max_ctas_real[2])();
};
const std::vector<FlatEntry> &flat_registry();

// first registered entry of the size (real_dir: -1 complex, 0 RealFFT forward, 1 RealFFT inverse -- the first entry that
// carries that kernel), or the variant named by SSFFT_FLAT_VARIANT="ring,ctas_per_sm[,inplace]" / SSFFT_FLAT_NAME=<part of
// the entry name> (SSFFT_FLAT_NAME_INV for the inverse real transform alone)
template <typename T>
inline int find_flat(size_t n1, size_t n2, int real_dir = -1) {
    const int prec = sizeof(T) == 4 ? 0 : 1;
    const auto &reg = flat_registry();
    int want_ring = 0, want_minb = 0, want_inplace = 1;
    if (const char *e = getenv("SSFFT_FLAT_VARIANT")) sscanf(e, "%d,%d,%d", &want_ring, &want_minb, &want_inplace);
    const char *want_name = getenv("SSFFT_FLAT_NAME");
    if (real_dir == 1 && getenv("SSFFT_FLAT_NAME_INV")) want_name = getenv("SSFFT_FLAT_NAME_INV");
    int first = -1;
    for (size_t i = 0; i < reg.size(); ++i)
        if (reg[i].prec == prec && (size_t)reg[i].n1 == n1 && (size_t)reg[i].n2 == n2 && (real_dir < 0 || reg[i].launch_real[real_dir])) {
            if (want_name && *want_name && strstr(reg[i].name, want_name)) return (int)i;
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
// Real-transform twiddles (complex length M = n1 * n2, real length N = 2 M), see FlatParams::ra / rb
template <typename T>
inline void fill_flat_real_tables(std::vector<T> &ra, std::vector<T> &rb, size_t n1, size_t n2, bool inverse) {
    const size_t big = 2 * n1 * n2;
    ra.assign(2 * n1, (T)0);
    rb.assign(2 * n2, (T)0);
    const long double pi2 = 2.0L * 3.14159265358979323846264338327950288L;
    if (!inverse) {
        for (size_t k1 = 0; k1 < n1; ++k1) flat_root<T>(&ra[2 * k1], k1, big);              // W_N^k1
        for (size_t k2 = 0; k2 < n2; ++k2) {                                                  // -i W_(2 n2)^k2 / 2
            const long double a = pi2 * (long double)k2 / (long double)(2 * n2);
            rb[2 * k2] = (T)(-0.5L * sinl(a));
            rb[2 * k2 + 1] = (T)(-0.5L * cosl(a));
        }
    } else {
        for (size_t j = 0; j < n1; ++j) {                                                     // i conj(W_(2 n1)^j) = i e^{+ia} = (-sin a, cos a)
            const long double a = pi2 * (long double)j / (long double)(2 * n1);
            ra[2 * j] = (T)(-sinl(a));
            ra[2 * j + 1] = (T)cosl(a);
        }
        for (size_t c = 0; c < n2; ++c) {                                                     // conj(W_N^c)
            const long double a = pi2 * (long double)c / (long double)big;
            rb[2 * c] = (T)cosl(a);
            rb[2 * c + 1] = (T)sinl(a);
        }
    }
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
