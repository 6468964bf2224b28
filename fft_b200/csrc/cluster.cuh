// cluster.cuh -- cluster-resident four-step: one transform lives in the shared memory of a thread-block cluster.
//
// For transforms a little too large for one CTA (complex 2^14 .. 2^16, real 2^15 .. 2^17 in fp32) the L2-scratch
// four-step of tiled.cuh pays for the intermediate twice (store + load through L2, plus the cluster-scope
// barrier that publishes it).  Here the N = N1 * N2 points of a transform are spread over the C CTAs of a
// cluster, N/C = 4096 points (16 per thread) each, and the transposition between the column stage and the row
// stage is ONE all-to-all through distributed shared memory (st.shared::cluster): HBM sees each input and output
// byte once and nothing else leaves the SMs.  Same mathematics as the reference's plan (signalsmith-fft.h:93-185,
// cache-blocking rule :130-133) and the same per-pass structure as fused.cuh:
//
//   stage 1  CTA `rank` owns columns n2 in [rank*CT1, (rank+1)*CT1), CT1 = N2/C:
//     A0  x[n1*N2 + n2] --coalesced--> registers, radix-RA0 butterflies, pass twiddles, exchange through bufA
//     A1  radix-RA1 butterflies (threads transposed: lanes run along k1), times W_N^(n2*k1),
//         then every value is stored into bufB of the CTA that owns row k1            <- DSMEM all-to-all
//   stage 2  CTA `rank` owns rows k1 (CT2 = N1/C of them), bufB = [n2][row lane]:
//     B0  radix-RB0 butterflies, pass twiddles, exchange through bufA
//     B1  radix-RB1 butterflies, X[k1 + N1*k2] --runs of CT2 consecutive k1--> HBM
//
// Real transforms (RealFFT<V>::fft / ifft, :446-502) keep the reference's "N/2 complex + twiddle" scheme:
//   R2C  rows are owned in mirror pairs (k1, N1-k1) so bins i and N/2-i land in the same CTA and the
//        post-twiddle (:459-472) is a CTA-local epilogue of B1;
//   C2R  the pre-twiddle (:478-492) is applied on the fly while A0 gathers its inputs (each thread fetches the
//        partner bin itself), the inverse runs as swap-fft-swap.
//
// Synchronisation per transform: ONE relaxed cluster barrier "every bufB is free" (arrive right after A0, wait just
// before the all-to-all, so it hides behind the A1 butterflies), and a per-CTA mbarrier "my bufB is full" that the
// peers' st.async stores complete byte by byte -- no fence, no second cluster-wide barrier.  bufB doubles as the
// landing zone of the NEXT transform's input tile, prefetched by TMA bulk copies while B1 runs.
//
// Every phase is a __host__ __device__ function of (tid, rank) over explicit buffers, so the complete index
// logic is executed on the CPU by tests/host/cluster_emul.cu (no GPU needed) as well as on the device.
#pragma once
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through the runtime, libcuda is not linked)
#include <cuda_runtime.h>

#include <vector>

#include "codelets.cuh"
#include "fused.cuh"
#include "real_kernels.cuh"

namespace ssfft {

enum { CL_C2C = 0, CL_R2C = 1, CL_C2R = 2 };

template <typename T_, int N1_, int RA0_, int RA1_, int N2_, int RB0_, int RB1_, int C_, int MINB_>
struct ClusterCfg {
    using T = T_;
    static constexpr int N1 = N1_, N2 = N2_, C = C_, MINB = MINB_;
    static constexpr int N = N1 * N2, M = N / C, THREADS = 256, E = M / THREADS;
    static constexpr int RA0 = RA0_, RA1 = RA1_, RB0 = RB0_, RB1 = RB1_;
    static constexpr int CT1 = N2 / C, TX1 = THREADS / CT1;  // stage 1: CT1 column lanes x TX1 butterfly threads
    static constexpr int CT2 = N1 / C, TX2 = THREADS / CT2;  // stage 2: CT2 row lanes x TX2 butterfly threads
    // bufA pitch: the transposed gather of A1 (half-warp = TX1 butterfly threads x 16/TX1 columns) must hit 16
    // different 8-byte bank pairs: pitch == 16/TX1 (mod 16)
    static constexpr int PITCH_A = CT1 + (TX1 >= 16 ? 1 : 16 / TX1);
    static constexpr int BUFA = (N1 * PITCH_A > N2 * CT2) ? N1 * PITCH_A : N2 * CT2;
    static constexpr int TWA = (RA0 - 1) * (N1 / RA0), TWB = (RB0 - 1) * (N2 / RB0);
    // bufB lane swizzle: with TX1 == 8 a half-warp of the all-to-all stores 8 lanes of two consecutive rows n2
    static constexpr int SWZ = (TX1 == 8) ? 8 : 0;
    static constexpr size_t smem_bytes = (size_t)(BUFA + M + TWA + TWB) * sizeof(cx<T>);
    static_assert(E == 16, "a thread owns 16 points");
    static_assert(RA0 * RA1 == N1 && RB0 * RB1 == N2, "two passes per stage");
    static_assert(E % RA0 == 0 && E % RA1 == 0 && E % RB0 == 0 && E % RB1 == 0, "radices must divide 16");
    static_assert(CT1 * TX1 == THREADS && CT2 * TX2 == THREADS, "thread mappings");
    static_assert(N1 / TX1 == E && N2 / TX2 == E, "16 points per thread in both stages");
    static_assert(TX1 == 8 || TX1 == 16, "stage-1 butterfly threads");
    static_assert(CT2 % 2 == 0 && CT2 >= 16, "row lanes");
    __host__ __device__ static constexpr int swz(int n2) { return (n2 & 1) ? SWZ : 0; }
    __host__ __device__ static constexpr int bufb_index(int n2, int lane) { return n2 * CT2 + (lane ^ swz(n2)); }
};

// Which rows k1 a CTA owns in stage 2.  C2C / C2R: CT2 consecutive rows.  R2C: CT2/2 mirror pairs -- lane l < h
// holds row p = rank*h + l, lane l + h holds its mirror N1 - p (rows 0 and N1/2, which are their own mirrors,
// share pair 0), so the RealFFT post-twiddle partner of every bin is in the same CTA.
template <typename Cfg, int KIND>
struct ClusterMap {
    static constexpr int H = Cfg::CT2 / 2;
    SSFFT_HD static int row_of(int rank, int lane) {
        if constexpr (KIND == CL_R2C) {
            const int p = rank * H + (lane < H ? lane : lane - H);
            return lane < H ? p : (p == 0 ? Cfg::N1 / 2 : Cfg::N1 - p);
        } else {
            return rank * Cfg::CT2 + lane;
        }
    }
    SSFFT_HD static void row_dest(int k1, int &owner, int &lane) {
        if constexpr (KIND == CL_R2C) {
            if (k1 < Cfg::N1 / 2) { owner = k1 / H; lane = k1 % H; }
            else if (k1 == Cfg::N1 / 2) { owner = 0; lane = H; }
            else { const int p = Cfg::N1 - k1; owner = p / H; lane = H + p % H; }
        } else {
            owner = k1 / Cfg::CT2;
            lane = k1 % Cfg::CT2;
        }
    }
};

template <typename T>
struct ClusterParams {
    const cx<T> *in;
    cx<T> *out;
    const cx<T> *tw_a, *tw_b;  // pass-0 twiddles of the length-N1 / length-N2 transforms, [r-1][m']
    const cx<T> *tw4;          // W_N^(n2*k1) laid out [n2][k1]
    const cx<T> *rtw;          // RealFFT twiddlesMinusI (N/2 + 1 entries, N = complex length); real kinds only
    long long batch;
    int inverse;               // C2C only
    int use_tma;               // input tiles are prefetched through the tensor map passed next to these parameters
};

// ---------------------------------------------------------------------------------------------------------
// phases (host + device).  Env supplies the memory operations: ld_in / st_out (streaming HBM accesses),
// ld_tab (read-only tables), remote_store(owner, element index in bufB, value).
// ---------------------------------------------------------------------------------------------------------
#pragma nv_exec_check_disable
template <typename Cfg, int KIND, bool STAGED, typename Env>
SSFFT_HD void cl_a0(Env &env, int tid, int rank, const cx<typename Cfg::T> *gin, const cx<typename Cfg::T> *stage,
                    const cx<typename Cfg::T> *rtw, int inverse, const cx<typename Cfg::T> *stw, cx<typename Cfg::T> *bufA) {
    using T = typename Cfg::T;
    constexpr int R = Cfg::RA0, NR = Cfg::N1 / R, U = Cfg::E / R, TX = Cfg::TX1, CT = Cfg::CT1, N2 = Cfg::N2;
    constexpr int H = Cfg::N;  // complex length
    const int c = tid % CT, t = tid / CT;
    const int col = rank * CT + c;
    cx<T> v[Cfg::E];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const int i = (t + TX * u + NR * j) * N2 + col;
            cx<T> x;
            if constexpr (KIND == CL_C2R) {
                // RealFFT::ifft pre-twiddle (:478-492) on the fly; then swap for the swap-fft-swap inverse
                const int ci = i ? H - i : 0;
                const bool lo = 2 * i <= H;
                const cx<T> vi = env.ld_in(gin + i), vc = env.ld_in(gin + ci);
                const cx<T> w = env.ld_tab(rtw + (lo ? i : ci));
                cx<T> bi, bc;
                c2r_pair(lo ? vi : vc, lo ? vc : vi, w, bi, bc);
                x = lo ? bi : bc;
                if (i == 0) x = mk<T>(vi.x + vi.y, vi.x - vi.y);  // (DC, Nyquist) unpack  :478-481
                x = cswap(x);
            } else {
                // STAGED: this CTA's [N1][CT1] input tile was prefetched into `stage` (pitch CT1) by bulk copies
                if constexpr (STAGED) x = stage[(t + TX * u + NR * j) * CT + c];
                else x = env.ld_in(gin + i);
                if (KIND == CL_C2C && inverse) x = cswap(x);
            }
            v[u * R + j] = x;
        }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        cx<T> w[R];
#pragma unroll
        for (int j = 0; j < R; ++j) w[j] = v[u * R + j];
        Dft<R>::run(w);
        const int b = t + TX * u;  // P == 1: m' == b
#pragma unroll
        for (int r = 1; r < R; ++r) w[r] = cmul(w[r], stw[(r - 1) * NR + b]);
#pragma unroll
        for (int r = 0; r < R; ++r) bufA[(r + R * b) * Cfg::PITCH_A + c] = w[r];
    }
}

// A1: gather (transposed thread mapping), last butterflies of stage 1, four-step twiddle.  Results stay in v.
#pragma nv_exec_check_disable
template <typename Cfg, int KIND, typename Env>
SSFFT_HD void cl_a1(Env &env, int tid, int rank, const cx<typename Cfg::T> *bufA, const cx<typename Cfg::T> *tw4,
                    cx<typename Cfg::T> (&v)[Cfg::E]) {
    using T = typename Cfg::T;
    constexpr int R = Cfg::RA1, P = Cfg::RA0, NR = Cfg::N1 / R, U = Cfg::E / R, TX = Cfg::TX1;
    const int t2 = tid % TX, c2 = tid / TX;
    const int n2 = rank * Cfg::CT1 + c2;
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int j = 0; j < R; ++j) v[u * R + j] = bufA[(t2 + TX * u + NR * j) * Cfg::PITCH_A + c2];
    const cx<T> *twp = tw4 + (size_t)n2 * Cfg::N1;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        cx<T> w[R];
#pragma unroll
        for (int j = 0; j < R; ++j) w[j] = v[u * R + j];
        Dft<R>::run(w);
        const int b = t2 + TX * u;  // last pass of the length-N1 transform: output index k1 = b + P*r
#pragma unroll
        for (int r = 0; r < R; ++r) v[u * R + r] = cmul(w[r], env.ld_tab(twp + b + P * r));
    }
}

// the all-to-all: value (k1, n2) goes to bufB[n2][lane(k1)] of the CTA that owns row k1
#pragma nv_exec_check_disable
template <typename Cfg, int KIND, typename Env>
SSFFT_HD void cl_a1_scatter(Env &env, int tid, int rank, const cx<typename Cfg::T> (&v)[Cfg::E]) {
    constexpr int R = Cfg::RA1, P = Cfg::RA0, U = Cfg::E / R, TX = Cfg::TX1;
    const int t2 = tid % TX, c2 = tid / TX;
    const int n2 = rank * Cfg::CT1 + c2;
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int k1 = t2 + TX * u + P * r;
            int owner, lane;
            ClusterMap<Cfg, KIND>::row_dest(k1, owner, lane);
            env.remote_store(owner, Cfg::bufb_index(n2, lane), v[u * R + r]);
        }
}

template <typename Cfg>
SSFFT_HD void cl_b0_gather(int tid, const cx<typename Cfg::T> *bufB, cx<typename Cfg::T> (&v)[Cfg::E]) {
    constexpr int R = Cfg::RB0, NR = Cfg::N2 / R, U = Cfg::E / R, TX = Cfg::TX2, CT = Cfg::CT2;
    const int c = tid % CT, t = tid / CT;
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int j = 0; j < R; ++j) v[u * R + j] = bufB[Cfg::bufb_index(t + TX * u + NR * j, c)];
}

template <typename Cfg>
SSFFT_HD void cl_b0_compute(int tid, cx<typename Cfg::T> (&v)[Cfg::E], const cx<typename Cfg::T> *stw,
                            cx<typename Cfg::T> *bufA) {
    using T = typename Cfg::T;
    constexpr int R = Cfg::RB0, NR = Cfg::N2 / R, U = Cfg::E / R, TX = Cfg::TX2, CT = Cfg::CT2;
    const int c = tid % CT, t = tid / CT;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        cx<T> w[R];
#pragma unroll
        for (int j = 0; j < R; ++j) w[j] = v[u * R + j];
        Dft<R>::run(w);
        const int b = t + TX * u;
#pragma unroll
        for (int r = 1; r < R; ++r) w[r] = cmul(w[r], stw[(r - 1) * NR + b]);
#pragma unroll
        for (int r = 0; r < R; ++r) bufA[(r + R * b) * CT + c] = w[r];
    }
}

template <typename Cfg>
SSFFT_HD void cl_b1_gather(int tid, const cx<typename Cfg::T> *bufA, cx<typename Cfg::T> (&v)[Cfg::E]) {
    constexpr int R = Cfg::RB1, NR = Cfg::N2 / R, U = Cfg::E / R, TX = Cfg::TX2, CT = Cfg::CT2;
    const int c = tid % CT, t = tid / CT;
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int j = 0; j < R; ++j) v[u * R + j] = bufA[(t + TX * u + NR * j) * CT + c];
}

// B1 butterflies; C2C / C2R store X[k1 + N1*k2] to HBM, R2C leaves the bins in bufA ([k2][lane]) for the epilogue
#pragma nv_exec_check_disable
template <typename Cfg, int KIND, typename Env>
SSFFT_HD void cl_b1_finish(Env &env, int tid, int rank, cx<typename Cfg::T> (&v)[Cfg::E], int inverse,
                           cx<typename Cfg::T> *gout, cx<typename Cfg::T> *bufA) {
    using T = typename Cfg::T;
    constexpr int R = Cfg::RB1, P = Cfg::RB0, U = Cfg::E / R, TX = Cfg::TX2, CT = Cfg::CT2;
    const int c = tid % CT, t = tid / CT;
    const int k1 = ClusterMap<Cfg, KIND>::row_of(rank, c);
#pragma unroll
    for (int u = 0; u < U; ++u) {
        cx<T> w[R];
#pragma unroll
        for (int j = 0; j < R; ++j) w[j] = v[u * R + j];
        Dft<R>::run(w);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int k2 = t + TX * u + P * r;
            if constexpr (KIND == CL_R2C) {
                v[u * R + r] = w[r];
                bufA[k2 * CT + c] = w[r];
            } else {
                cx<T> x = w[r];
                if (KIND == CL_C2R || inverse) x = cswap(x);
                env.st_out(gout + k1 + (size_t)Cfg::N1 * k2, x);
            }
        }
    }
}

// RealFFT::fft post-twiddle (:459-472): every thread finishes its own 16 bins; the partner bin N/2 - i sits in the
// mirror lane of the same CTA (ClusterMap).  v still holds this thread's bins.
#pragma nv_exec_check_disable
template <typename Cfg, typename Env>
SSFFT_HD void cl_r2c_epilogue(Env &env, int tid, int rank, const cx<typename Cfg::T> (&v)[Cfg::E],
                              const cx<typename Cfg::T> *bufA, const cx<typename Cfg::T> *rtw, cx<typename Cfg::T> *gout) {
    using T = typename Cfg::T;
    constexpr int R = Cfg::RB1, P = Cfg::RB0, U = Cfg::E / R, TX = Cfg::TX2, CT = Cfg::CT2, N1 = Cfg::N1, N2 = Cfg::N2;
    constexpr int H = Cfg::N, HL = CT / 2;
    const int c = tid % CT, t = tid / CT;
    const int k1 = ClusterMap<Cfg, CL_R2C>::row_of(rank, c);
    const bool self = (k1 == 0 || k1 == N1 / 2);
    const int pl = self ? c : (c < HL ? c + HL : c - HL);  // lane of row (N1 - k1) mod N1
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int k2 = t + TX * u + P * r;
            const int i = k1 + N1 * k2;
            const cx<T> z = v[u * R + r];
            cx<T> o;
            if (i == 0) {
                o = mk<T>(z.x + z.y, z.x - z.y);  // DC in .re, Nyquist in .im  (:459-462)
            } else {
                const int k2p = (k1 == 0) ? N2 - k2 : N2 - 1 - k2;
                const cx<T> zc = bufA[k2p * CT + pl];
                const int ci = H - i;
                const bool lo = 2 * i <= H;
                const cx<T> w = env.ld_tab(rtw + (lo ? i : ci));
                cx<T> oi, oc;
                r2c_pair(lo ? z : zc, lo ? zc : z, w, oi, oc);
                o = lo ? oi : oc;
            }
            env.st_out(gout + i, o);
        }
}

#ifdef __CUDACC__

#ifdef SSFFT_EMUL
// CPU execution of the kernel (tests/host/simt/simt_emul.h): hooks instead of PTX.  The all-to-all store lands in the
// peer CTA's shared memory and completes bytes of the peer's "bufB full" mbarrier, as st.async does.
inline unsigned cl_ctarank() { return simt::cluster_ctarank(); }
inline void cl_arrive_release() { simt::cluster_arrive(); }
inline void cl_arrive_relaxed() { simt::cluster_arrive(); }
inline void cl_wait() { simt::cluster_wait(); }
inline void mbar_wait_cluster(unsigned long long *bar, unsigned parity) { simt::mbar_wait(bar, parity); }
template <typename T>
struct ClusterDevEnv {
    size_t bufb_off;            // offset of bufB in the dynamic shared memory (the same in every CTA of the cluster)
    unsigned long long *bar;    // this CTA's "bufB full" mbarrier (the peers' is at the same place)
    cx<T> ld_in(const cx<T> *p) const { return ld_stream(p); }
    void st_out(cx<T> *p, cx<T> v) const { st_stream(p, v); }
    cx<T> ld_tab(const cx<T> *p) const { return ld_table(p); }
    void remote_store(int owner, int idx, cx<T> v) const {
        cx<T> *peer = reinterpret_cast<cx<T> *>(simt::peer_smem((unsigned)owner) + bufb_off) + idx;
        simt::remote_store_tx((unsigned)owner, peer, &v, (unsigned)sizeof(cx<T>), bar);
    }
};
#else
__device__ __forceinline__ unsigned cl_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cl_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cl_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cl_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// wait for the peers' st.async data: acquire at cluster scope; bounded like mbar_wait (a lost signal traps, never hangs)
__device__ __forceinline__ void mbar_wait_cluster(unsigned long long *bar, unsigned parity) {
    const unsigned addr = smem_u32(bar);
    const long long t0 = clock64();
    for (;;) {
        unsigned done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}

// Peers are addressed through the shared::cluster window (mapa); a store carries its own completion signal:
// st.async ... mbarrier::complete_tx::bytes adds the 8 bytes to the DESTINATION CTA's "bufB full" mbarrier, so the
// all-to-all needs neither a fence nor a cluster-wide barrier on the consumer side.
template <typename T>
struct ClusterDevEnv {
    unsigned bufb;  // shared::cta address of this CTA's bufB (same offset in every CTA of the cluster)
    unsigned bar;   // shared::cta address of the "bufB full" mbarrier
    __device__ __forceinline__ cx<T> ld_in(const cx<T> *p) const { return ld_stream(p); }
    __device__ __forceinline__ void st_out(cx<T> *p, cx<T> v) const { st_stream(p, v); }
    __device__ __forceinline__ cx<T> ld_tab(const cx<T> *p) const { return ld_table(p); }
    __device__ __forceinline__ void remote_store(int owner, int idx, cx<T> v) const {
        const unsigned la = bufb + idx * (unsigned)sizeof(cx<T>);
        unsigned ra;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(owner));
        const unsigned rb = ra + (bar - la);  // same CTA window, same offsets: the peer's mbarrier
        if constexpr (sizeof(T) == 4)
            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(ra),
                         "f"(v.x), "f"(v.y), "r"(rb)
                         : "memory");
        else
            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(ra),
                         "d"(v.x), "d"(v.y), "r"(rb)
                         : "memory");
    }
};
#endif  // SSFFT_EMUL

template <typename Cfg, int KIND>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB)
cluster_fft_kernel(ClusterParams<typename Cfg::T> q, const __grid_constant__ CUtensorMap tmap) {
    using T = typename Cfg::T;
    SSFFT_DYNAMIC_SMEM(ssfft_smem);
    __shared__ __align__(8) unsigned long long bars[2];  // [0] "bufB full" (peers' st.async)   [1] input tile (TMA)
    cx<T> *bufA = reinterpret_cast<cx<T> *>(ssfft_smem);
    cx<T> *bufB = bufA + Cfg::BUFA;
    cx<T> *stwA = bufB + Cfg::M;
    cx<T> *stwB = stwA + Cfg::TWA;
    const int tid = threadIdx.x;
    const int rank = (int)cl_ctarank();
    const long long cid = blockIdx.x / Cfg::C, nclusters = gridDim.x / Cfg::C;
    // C2R gathers bin pairs from all over the spectrum: plain loads.  Otherwise ONE TMA tensor copy per transform
    // brings this CTA's [N1][CT1] tile (box of the (batch, N1, N2) tensor) -- no registers, no LSU slots.
    const bool PF = (KIND != CL_C2R) && q.use_tma;
    constexpr unsigned kTileBytes = (unsigned)(Cfg::M * sizeof(cx<T>));
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
    }
    for (int i = tid; i < Cfg::TWA; i += Cfg::THREADS) stwA[i] = ld_table(q.tw_a + i);
    for (int i = tid; i < Cfg::TWB; i += Cfg::THREADS) stwB[i] = ld_table(q.tw_b + i);
    __syncthreads();
    cl_arrive_release();  // every CTA's mbarriers are initialised before a peer may signal them
    cl_wait();
#ifdef SSFFT_EMUL
    ClusterDevEnv<T> env{(size_t)(reinterpret_cast<unsigned char *>(bufB) - ssfft_smem), &bars[0]};
#else
    ClusterDevEnv<T> env{smem_u32(bufB), smem_u32(&bars[0])};
#endif
    auto prefetch = [&](long long Bn) {
        if (Bn < q.batch && tid == 0) {
            mbar_expect_tx(&bars[1], kTileBytes);
#ifdef SSFFT_EMUL
            // the box [N1][CT1] at (x = rank * CT1, y = 0, z = Bn) of the (batch, N1, N2) tensor, one bulk copy per row
            for (int r = 0; r < Cfg::N1; ++r)
                bulk_g2s(bufB + r * Cfg::CT1, q.in + Bn * Cfg::N + (long long)r * Cfg::N2 + rank * Cfg::CT1,
                         (unsigned)(Cfg::CT1 * sizeof(cx<T>)), &bars[1]);
#else
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                    smem_u32(bufB)),
                "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(rank * Cfg::CT1), "r"(0), "r"((int)Bn), "r"(smem_u32(&bars[1]))
                : "memory");
#endif
        }
    };
    if (PF) prefetch(cid);
    unsigned par_in = 0, par_full = 0;
    for (long long B = cid; B < q.batch; B += nclusters) {
        const cx<T> *gin = q.in + B * Cfg::N;
        cx<T> *gout = q.out + B * Cfg::N;
        cx<T> v[Cfg::E];
        if (tid == 0) mbar_expect_tx(&bars[0], kTileBytes);  // this transform's all-to-all: M elements land in my bufB
        if (PF) {
            mbar_wait(&bars[1], par_in);  // my input tile has landed in bufB
            par_in ^= 1u;
            cl_a0<Cfg, KIND, true>(env, tid, rank, gin, bufB, q.rtw, q.inverse, stwA, bufA);
        } else {
            cl_a0<Cfg, KIND, false>(env, tid, rank, gin, bufB, q.rtw, q.inverse, stwA, bufA);
        }
        __syncthreads();      // every thread has consumed its part of bufB (and the previous transform left it long ago)
        cl_arrive_relaxed();  // "my bufB is free": peers may start their all-to-all stores into it
        cl_a1<Cfg, KIND>(env, tid, rank, bufA, q.tw4, v);
        // bufA is rewritten by B0 below.  The "bufB full" mbarrier already orders that after every thread's A1 gather
        // (each thread contributes elements to every CTA of the cluster, its own included), but a block barrier makes
        // the ordering explicit -- and visible to compute-sanitizer's racecheck -- for well under 1 % of the time.
        __syncthreads();
        cl_wait();            // every bufB of the cluster is free
        cl_a1_scatter<Cfg, KIND>(env, tid, rank, v);
        mbar_wait_cluster(&bars[0], par_full);  // all M elements of my rows have arrived
        par_full ^= 1u;
        cl_b0_gather<Cfg>(tid, bufB, v);
        cl_b0_compute<Cfg>(tid, v, stwB, bufA);
        __syncthreads();      // bufB consumed
        if (PF) prefetch(B + nclusters);  // next input tile streams in behind B1 and the stores
        cl_b1_gather<Cfg>(tid, bufA, v);
        __syncthreads();      // bufA may be overwritten (R2C epilogue image / next transform's A0)
        cl_b1_finish<Cfg, KIND>(env, tid, rank, v, q.inverse, gout, bufA);
        if constexpr (KIND == CL_R2C) {
            __syncthreads();
            cl_r2c_epilogue<Cfg>(env, tid, rank, v, bufA, q.rtw, gout);
            __syncthreads();
        }
    }
}

#endif  // __CUDACC__

// ---------------------------------------------------------------------------------------------------------
// host side: tables + registry
// ---------------------------------------------------------------------------------------------------------
// pass-0 twiddles of a two-pass length-L transform with first radix R0: [r-1][m'] = W_L^(m'*r), m' < L/R0
template <typename T>
inline void fill_cluster_pass_twiddles(std::vector<T> &h, int L, int R0) {
    const int MN = L / R0;
    h.assign(2 * (size_t)(R0 - 1) * MN, (T)0);
    size_t o = 0;
    for (int r = 1; r < R0; ++r)
        for (int m = 0; m < MN; ++m, ++o) {
            const long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)((long long)m * r % L) / (long double)L;
            h[2 * o] = (T)cosl(a);
            h[2 * o + 1] = (T)(-sinl(a));
        }
}
// W_N^(n2*k1) laid out [n2][k1]
template <typename T>
inline void fill_cluster_tw4(std::vector<T> &h, int n1, int n2) {
    const long long n = (long long)n1 * n2;
    h.assign(2 * (size_t)n, (T)0);
    for (int c = 0; c < n2; ++c)
        for (int k = 0; k < n1; ++k) {
            const long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)((long long)c * k % n) / (long double)n;
            const size_t o = (size_t)c * n1 + k;
            h[2 * o] = (T)cosl(a);
            h[2 * o + 1] = (T)(-sinl(a));
        }
}

struct ClusterEntry {
    int prec, n1, n2, csize;
    int ra0, rb0;
    const char *name;
    size_t smem_bytes;
    unsigned kinds;  // bit k set: launch[k] is meant to be used (C2C / R2C / C2R prefer different orientations)
    int (*launch[3])(const void *params, int max_clusters, cudaStream_t s);
    int (*max_clusters[3])();  // co-resident clusters on the current device (<= 0: cannot be scheduled)
};
const std::vector<ClusterEntry> &cluster_registry();

// first registered entry for complex length n that serves `kind`
template <typename T>
inline int find_cluster(size_t n, int kind) {
    const int prec = sizeof(T) == 4 ? 0 : 1;
    const auto &reg = cluster_registry();
    for (size_t i = 0; i < reg.size(); ++i)
        if (reg[i].prec == prec && (size_t)reg[i].n1 * reg[i].n2 == n && (reg[i].kinds >> kind & 1u)) return (int)i;
    return -1;
}

}  // namespace ssfft
