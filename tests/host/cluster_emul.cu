// cluster_emul.cu -- runs the cluster-resident four-step (fft_b200/csrc/cluster.cuh) ON THE CPU.
//
// The kernel's phases are __host__ __device__ functions of (tid, rank); this program plays all C CTAs x 256
// threads of a cluster phase by phase (barriers = loop boundaries), with plain arrays for the shared buffers and
// the DSMEM all-to-all, and checks the result against a double-precision DFT.  It validates the complete index
// logic (thread mappings, ownership maps, swizzles, RealFFT pairing) without a GPU; it also counts shared-memory
// bank conflicts of every access pattern under the half-warp model (16 lanes x 8-byte bank pairs).
// Test infrastructure only -- nothing in the product links against it.
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../fft_b200/csrc/cluster.cuh"
#include "../../fft_b200/csrc/planner.h"

using namespace ssfft;
typedef std::complex<double> cd;

static void fft_double(std::vector<cd> &a, bool inverse) {  // iterative radix-2, unnormalised
    const size_t n = a.size();
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        const double ang = 2 * M_PI / (double)len * (inverse ? 1 : -1);
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; ++k) {
                const cd w(std::cos(ang * (double)k), std::sin(ang * (double)k));
                const cd u = a[i + k], v = a[i + k + len / 2] * w;
                a[i + k] = u + v;
                a[i + k + len / 2] = u - v;
            }
    }
}

struct ConflictCounter {
    long long accesses = 0, extra = 0;
    // one warp-wide 8-byte access: element indices of the 32 lanes (-1 = inactive)
    void warp(const int (&idx)[32]) {
        for (int h = 0; h < 2; ++h) {
            int cnt[16] = {0};
            int seen[16][16];
            int worst = 0;
            for (int l = 0; l < 16; ++l) {
                const int e = idx[h * 16 + l];
                if (e < 0) continue;
                const int b = e & 15;
                bool dup = false;
                for (int k = 0; k < cnt[b]; ++k) dup |= (seen[b][k] == e);
                if (!dup) seen[b][cnt[b]++] = e;
                if (cnt[b] > worst) worst = cnt[b];
            }
            if (worst > 0) { ++accesses; extra += worst - 1; }
        }
    }
};

template <typename Cfg>
struct HostEnv {
    using T = typename Cfg::T;
    std::vector<std::vector<cx<T>>> *bufB;
    cx<T> ld_in(const cx<T> *p) const { return *p; }
    void st_out(cx<T> *p, cx<T> v) const { *p = v; }
    cx<T> ld_tab(const cx<T> *p) const { return *p; }
    void remote_store(int owner, int idx, cx<T> v) const { (*bufB)[owner][idx] = v; }
};

template <typename Cfg, int KIND>
static double run_case(const char *name) {
    using T = typename Cfg::T;
    constexpr int C = Cfg::C, TH = Cfg::THREADS, N = Cfg::N, E = Cfg::E;
    std::vector<T> twa, twb, tw4, rtw(2 * (size_t)(N / 2 + 1));
    fill_cluster_pass_twiddles<T>(twa, Cfg::N1, Cfg::RA0);
    fill_cluster_pass_twiddles<T>(twb, Cfg::N2, Cfg::RB0);
    fill_cluster_tw4<T>(tw4, Cfg::N1, Cfg::N2);
    fill_real_twiddles<T>(rtw.data(), 2 * (size_t)N, false);
    const cx<T> *stwA = reinterpret_cast<const cx<T> *>(twa.data());
    const cx<T> *stwB = reinterpret_cast<const cx<T> *>(twb.data());
    const cx<T> *ptw4 = reinterpret_cast<const cx<T> *>(tw4.data());
    const cx<T> *prtw = reinterpret_cast<const cx<T> *>(rtw.data());

    // input + expected output (double)
    std::vector<cx<T>> in(N), out(N, mk<T>((T)777, (T)777));
    std::vector<cd> expect(N);
    srand(12345 + KIND);
    auto rnd = [] { return (double)rand() / RAND_MAX - 0.5; };
    int inverse = 0;
    if (KIND == CL_C2C) {
        std::vector<cd> x(N);
        for (int i = 0; i < N; ++i) { x[i] = cd(rnd(), rnd()); in[i] = mk<T>((T)x[i].real(), (T)x[i].imag()); }
        expect = x;
        fft_double(expect, false);
    } else if (KIND == CL_R2C) {
        std::vector<cd> x(2 * (size_t)N);
        for (int i = 0; i < 2 * N; ++i) x[i] = cd(rnd(), 0);
        for (int i = 0; i < N; ++i) in[i] = mk<T>((T)x[2 * i].real(), (T)x[2 * i + 1].real());
        fft_double(x, false);
        for (int i = 0; i < N; ++i) expect[i] = x[i];
        expect[0] = cd(x[0].real(), x[N].real());  // (DC, Nyquist)
    } else {
        // packed half spectrum of a real signal -> N_real * signal
        std::vector<cd> x(2 * (size_t)N), sp;
        for (int i = 0; i < 2 * N; ++i) x[i] = cd(rnd(), 0);
        sp = x;
        fft_double(sp, false);
        for (int i = 0; i < N; ++i) in[i] = mk<T>((T)sp[i].real(), (T)sp[i].imag());
        in[0] = mk<T>((T)sp[0].real(), (T)sp[N].real());
        for (int i = 0; i < N; ++i) expect[i] = cd(2.0 * N * x[2 * i].real(), 2.0 * N * x[2 * i + 1].real());
    }

    std::vector<std::vector<cx<T>>> bufA(C, std::vector<cx<T>>(Cfg::BUFA)), bufB(C, std::vector<cx<T>>(Cfg::M));
    std::vector<cx<T>> regs((size_t)C * TH * E);
    HostEnv<Cfg> env{&bufB};
    auto V = [&](int rank, int tid) -> cx<T>(&)[E] { return *reinterpret_cast<cx<T>(*)[E]>(&regs[((size_t)rank * TH + tid) * E]); };

    for (int inv = 0; inv <= (KIND == CL_C2C ? 1 : 0); ++inv) {
        inverse = inv;
        if (inv) {  // inverse C2C: expected = unnormalised inverse DFT
            for (int i = 0; i < N; ++i) expect[i] = cd(in[i].x, in[i].y);
            fft_double(expect, true);
        }
        // the kernel's bulk-copy prefetch: row n1 of CTA r's input tile -> bufB[n1*CT1 ...] (C2C / R2C only)
        constexpr bool PF = KIND != CL_C2R;
        if (PF)
            for (int r = 0; r < C; ++r)
                for (int n1 = 0; n1 < Cfg::N1; ++n1)
                    for (int c = 0; c < Cfg::CT1; ++c) bufB[r][n1 * Cfg::CT1 + c] = in[(size_t)n1 * Cfg::N2 + r * Cfg::CT1 + c];
        for (int r = 0; r < C; ++r) for (int t = 0; t < TH; ++t) cl_a0<Cfg, KIND, PF>(env, t, r, in.data(), bufB[r].data(), prtw, inverse, stwA, bufA[r].data());
        for (int r = 0; r < C; ++r) for (int t = 0; t < TH; ++t) cl_a1<Cfg, KIND>(env, t, r, bufA[r].data(), ptw4, V(r, t));
        for (auto &b : bufB) for (auto &e : b) e = mk<T>((T)NAN, (T)NAN);  // every element must be written by the all-to-all
        for (int r = 0; r < C; ++r) for (int t = 0; t < TH; ++t) cl_a1_scatter<Cfg, KIND>(env, t, r, V(r, t));
        for (int r = 0; r < C; ++r) for (int t = 0; t < TH; ++t) cl_b0_gather<Cfg>(t, bufB[r].data(), V(r, t));
        for (int r = 0; r < C; ++r) for (int t = 0; t < TH; ++t) cl_b0_compute<Cfg>(t, V(r, t), stwB, bufA[r].data());
        for (int r = 0; r < C; ++r) for (int t = 0; t < TH; ++t) cl_b1_gather<Cfg>(t, bufA[r].data(), V(r, t));
        for (int r = 0; r < C; ++r) for (int t = 0; t < TH; ++t) cl_b1_finish<Cfg, KIND>(env, t, r, V(r, t), inverse, out.data(), bufA[r].data());
        if (KIND == CL_R2C)
            for (int r = 0; r < C; ++r) for (int t = 0; t < TH; ++t) cl_r2c_epilogue<Cfg>(env, t, r, V(r, t), bufA[r].data(), prtw, out.data());
        double num = 0, den = 0;
        for (int i = 0; i < N; ++i) {
            const cd d = cd(out[i].x, out[i].y) - expect[i];
            num += std::norm(d);
            den += std::norm(expect[i]);
        }
        const double err = std::sqrt(num / den);
        printf("%-28s kind %d inverse %d  relL2 %.3e\n", name, KIND, inv, err);
        if (!(err < 1e-6 * std::log2((double)N))) return err > 0 ? err : 1.0;
    }
    return 0.0;
}

// bank conflicts of every shared-memory access pattern (half-warp model), by replaying the index formulas
template <typename Cfg, int KIND>
static long long conflicts() {
    constexpr int TH = Cfg::THREADS, E = Cfg::E;
    ConflictCounter cc;
    for (int w = 0; w < TH / 32; ++w) {
        int idx[32];
        {   // A0 scatter
            constexpr int R = Cfg::RA0, U = E / R, TX = Cfg::TX1, CT = Cfg::CT1;
            for (int u = 0; u < U; ++u) for (int r = 0; r < R; ++r) {
                for (int l = 0; l < 32; ++l) { const int tid = w * 32 + l, c = tid % CT, t = tid / CT; idx[l] = (r + R * (t + TX * u)) * Cfg::PITCH_A + c; }
                cc.warp(idx);
            }
        }
        {   // A1 gather
            constexpr int R = Cfg::RA1, NR = Cfg::N1 / R, U = E / R, TX = Cfg::TX1;
            for (int u = 0; u < U; ++u) for (int j = 0; j < R; ++j) {
                for (int l = 0; l < 32; ++l) { const int tid = w * 32 + l, t2 = tid % TX, c2 = tid / TX; idx[l] = (t2 + TX * u + NR * j) * Cfg::PITCH_A + c2; }
                cc.warp(idx);
            }
        }
        {   // all-to-all stores: lanes that go to different CTAs cannot conflict -> offset them by owner * large
            constexpr int R = Cfg::RA1, P = Cfg::RA0, U = E / R, TX = Cfg::TX1;
            for (int rank = 0; rank < Cfg::C; ++rank)
                for (int u = 0; u < U; ++u) for (int r = 0; r < R; ++r) {
                    int own[32];
                    for (int l = 0; l < 32; ++l) {
                        const int tid = w * 32 + l, t2 = tid % TX, c2 = tid / TX, n2 = rank * Cfg::CT1 + c2;
                        int lane;
                        ClusterMap<Cfg, KIND>::row_dest(t2 + TX * u + P * r, own[l], lane);
                        idx[l] = Cfg::bufb_index(n2, lane);
                    }
                    for (int o = 0; o < Cfg::C; ++o) {  // per destination CTA
                        int sub[32];
                        bool any = false;
                        for (int l = 0; l < 32; ++l) { sub[l] = own[l] == o ? idx[l] : -1; any |= own[l] == o; }
                        if (any) cc.warp(sub);
                    }
                }
        }
        {   // B0 gather / scatter, B1 gather
            constexpr int R = Cfg::RB0, NR = Cfg::N2 / R, U = E / R, TX = Cfg::TX2, CT = Cfg::CT2;
            for (int u = 0; u < U; ++u) for (int j = 0; j < R; ++j) {
                for (int l = 0; l < 32; ++l) { const int tid = w * 32 + l, c = tid % CT, t = tid / CT; idx[l] = Cfg::bufb_index(t + TX * u + NR * j, c); }
                cc.warp(idx);
                for (int l = 0; l < 32; ++l) { const int tid = w * 32 + l, c = tid % CT, t = tid / CT; idx[l] = (j + R * (t + TX * u)) * CT + c; }
                cc.warp(idx);
            }
            constexpr int R1 = Cfg::RB1, NR1 = Cfg::N2 / R1, U1 = E / R1;
            for (int u = 0; u < U1; ++u) for (int j = 0; j < R1; ++j) {
                for (int l = 0; l < 32; ++l) { const int tid = w * 32 + l, c = tid % CT, t = tid / CT; idx[l] = (t + TX * u + NR1 * j) * CT + c; }
                cc.warp(idx);
            }
        }
    }
    printf("    shared-memory half-warp accesses %lld, extra (conflict) wavefronts %lld\n", cc.accesses, cc.extra);
    return cc.extra;
}

template <typename Cfg>
static int run_cfg(const char *name, unsigned kinds) {
    int bad = 0;
    if (kinds & 1u) { bad += run_case<Cfg, CL_C2C>(name) != 0.0; bad += conflicts<Cfg, CL_C2C>() != 0; }
    if (kinds & 2u) { bad += run_case<Cfg, CL_R2C>(name) != 0.0; bad += conflicts<Cfg, CL_R2C>() != 0; }
    if (kinds & 4u) { bad += run_case<Cfg, CL_C2R>(name) != 0.0; }
    return bad;
}

int main() {
    int bad = 0;
    bad += run_cfg<ClusterCfg<float, 128, 16, 8, 128, 16, 8, 4, 3>>("128x128 c4", 7u);
    bad += run_cfg<ClusterCfg<float, 256, 16, 16, 128, 16, 8, 8, 3>>("256x128 c8", 7u);
    bad += run_cfg<ClusterCfg<float, 128, 16, 8, 256, 16, 16, 8, 3>>("128x256 c8", 7u);
    bad += run_cfg<ClusterCfg<float, 256, 16, 16, 256, 16, 16, 16, 3>>("256x256 c16", 7u);
    printf(bad ? "CLUSTER-EMUL-FAILED (%d)\n" : "CLUSTER-EMUL-OK\n", bad);
    return bad ? 1 : 0;
}
