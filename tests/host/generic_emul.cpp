// generic_emul.cpp -- the any-size path (generic.cuh + generic_plan.h + the RealFFT passes of real_kernels.cuh) executed
// ON THE CPU (TEST INFRASTRUCTURE): the library's own planner decides radices, thread geometry, shared-memory layout and
// the four-step split, g++ compiles the unmodified kernels against tests/host/simt/, every CUDA thread runs as a fiber.
// Sizes: the reference's test sizes (tests/00-fft.cpp:8-16), primes with and without a codelet, mixed radices, and
// lengths beyond one CTA (generic four-step with the two-table epilogue twiddle); complex forward / inverse, RealFFT and
// ModifiedRealFFT forward / inverse through the same sequence of launches as ssfft.cu.  Compared with the oracle.
#include <cmath>
#include <cstdio>
#include <vector>

#include "simt/simt_emul.h"
#include "../../fft_b200/csrc/generic_plan.h"
#include "../../fft_b200/csrc/real_kernels.cuh"
#include "../../oracle/oracle_fft.h"

using namespace ssfft;

static int g_bad = 0, g_runs = 0;
static const int kSmemMax = (int)simt::kSmemBytes;

template <typename T> static void fill(T *dst, size_t count, uint64_t seed) {
    if (sizeof(T) == 4) oracle_fill_uniform_f32((float *)dst, count, seed, 0);
    else oracle_fill_uniform_f64((double *)dst, count, seed, 0);
}
template <typename T>
static double rel_l2(const T *got, const T *want, size_t scalars) {
    long double e = 0, n = 0;
    for (size_t i = 0; i < scalars; ++i) {
        const long double d = (long double)got[i] - (long double)want[i];
        e += d * d;
        n += (long double)want[i] * want[i];
    }
    if (!(e == e)) return 1e30;
    return n > 0 ? (double)sqrtl(e / n) : (double)sqrtl(e);
}

template <typename T>
struct Stage {
    GenericStage st;
    std::vector<T> roots;
    bool ok = false;
    Stage(size_t n) {
        ok = plan_generic_stage(st, n, sizeof(cx<T>), kSmemMax);
        roots.resize(2 * (n ? n : 1));
        fill_roots<T>(roots.data(), n, n, 1);
    }
    bool run(const void *in, void *out, long long batch, const GenericLayout &L, int inverse, const T *ep_lo, const T *ep_hi,
             int ep_shift, int ep_cols) {
        const GenericParams<T> p = make_generic_params<T>(st, roots.data(), batch, L, inverse, ep_lo, ep_hi, ep_shift, ep_cols);
        const unsigned blocks = (unsigned)((batch + st.fpb - 1) / st.fpb);
        return simt::launch(dim3(blocks), dim3((unsigned)st.tx, (unsigned)st.fpb),
                            [&] { generic_fft_kernel<T>((const cx<T> *)in, (cx<T> *)out, p); });
    }
};

// the complex core of ssfft.cu:exec_complex for a plan without a specialised kernel: one launch, or the generic four-step
template <typename T>
struct ComplexCore {
    size_t n;
    bool four_step = false;
    Stage<T> *direct = nullptr, *col = nullptr, *row = nullptr;
    GenericFourStep fs;
    std::vector<T> ep_lo, ep_hi, scratch;
    explicit ComplexCore(size_t n_) : n(n_) {
        const size_t limit = generic_limit(sizeof(cx<T>), kSmemMax);
        if (n <= limit) {
            direct = new Stage<T>(n);
        } else {
            four_step = true;
            if (!plan_generic_fourstep(fs, n, limit)) { printf("FAIL n=%zu: no split\n", n); ++g_bad; return; }
            col = new Stage<T>(fs.n1);
            row = new Stage<T>(fs.n2);
            ep_lo.resize(2 * fs.lo_count); ep_hi.resize(2 * fs.hi_count);
            fill_roots<T>(ep_lo.data(), n, fs.lo_count, 1);
            fill_roots<T>(ep_hi.data(), n, fs.hi_count, fs.lo_count);
        }
    }
    ~ComplexCore() { delete direct; delete col; delete row; }
    bool run(const void *in, void *out, long long batch, int inverse) {
        if (!four_step) {
            const GenericLayout L{(long long)n, 0, 1, 1, (long long)n, 0, 1, 1};
            return direct->run(in, out, batch, L, inverse, nullptr, nullptr, 0, 1);
        }
        scratch.assign(2 * n * (size_t)batch, (T)NAN);
        return col->run(in, scratch.data(), batch * (long long)fs.n2, fourstep_col_layout(n, fs.n2), inverse, ep_lo.data(), ep_hi.data(),
                        fs.ep_shift, (int)fs.n2) &&
               row->run(scratch.data(), out, batch * (long long)fs.n1, fourstep_row_layout(n, fs.n1, fs.n2), inverse, nullptr, nullptr,
                        fs.ep_shift, 1);
    }
};

template <typename T>
static void check(const char *what, size_t n, const T *got, const T *want, size_t scalars, double factor = 1.0) {
    const double lim = factor * (sizeof(T) == 4 ? 1e-6 : 1e-14) * std::log2((double)(n < 2 ? 2 : n));
    const double err = rel_l2(got, want, scalars);
    ++g_runs;
    if (getenv("EMUL_VERBOSE")) printf("  n=%zu %s: relL2 %.3e (limit %.3e)\n", n, what, err, lim);
    if (!(err <= lim)) { ++g_bad; printf("FAIL n=%zu %s: relL2 %.3e > %.3e\n", n, what, err, lim); }
}

// Lengths with a large prime factor p: the reference evaluates the roots of its O(p^2) step on a phase rounded to V
// (`V phase = 2*M_PI*f*i/factor`, signalsmith-fft.h:204), so its own float result is only good to ~1e-4 for p ~ 1000
// and the oracle, which restates it bit for bit, inherits that.  The kernels use exact-phase roots; for such lengths
// they are checked against a long-double DFT on 64 sampled bins per transform instead.
static size_t largest_prime_factor(size_t n) {
    size_t best = 1;
    for (size_t p : factorise(n)) best = p > best ? p : best;
    return best;
}
template <typename T>
static void check_sampled_exact(const char *what, size_t n, long long batch, const T *in, const T *got, int inverse) {
    long double err = 0, nrm = 0;
    const long double tau = 6.283185307179586476925286766559L;
    for (long long b = 0; b < batch; ++b)
        for (int i = 0; i < 64; ++i) {
            const size_t k = ((size_t)i * n / 64 + 7 * (size_t)i) % n;
            long double re = 0, im = 0;
            for (size_t j = 0; j < n; ++j) {
                const long double a = tau * (long double)((unsigned long long)j * k % n) / (long double)n;
                const long double c = cosl(a), sn = inverse ? sinl(a) : -sinl(a);
                const long double xr = in[2 * (b * n + j)], xi = in[2 * (b * n + j) + 1];
                re += xr * c - xi * sn;
                im += xr * sn + xi * c;
            }
            const long double dr = (long double)got[2 * (b * n + k)] - re, di = (long double)got[2 * (b * n + k) + 1] - im;
            err += dr * dr + di * di;
            nrm += re * re + im * im;
        }
    const double lim = (sizeof(T) == 4 ? 1e-6 : 1e-14) * std::log2((double)n), rel = (err == err) ? (double)sqrtl(err / nrm) : 1e30;
    ++g_runs;
    if (getenv("EMUL_VERBOSE")) printf("  n=%zu %s vs exact DFT: relL2 %.3e (limit %.3e)\n", n, what, rel, lim);
    if (!(rel <= lim)) { ++g_bad; printf("FAIL n=%zu %s vs exact DFT: relL2 %.3e > %.3e\n", n, what, rel, lim); }
}

template <typename T>
static void run_complex(size_t n, long long batch) {
    constexpr int prec = sizeof(T) == 4 ? 0 : 1;
    ComplexCore<T> core(n);
    std::vector<T> in(2 * n * batch), out(2 * n * batch), want(2 * n * batch);
    const bool reference_is_inexact = largest_prime_factor(n) > 64;
    for (int inverse = 0; inverse < 2; ++inverse) {
        fill(in.data(), in.size(), 71 + inverse);
        oracle_batch(inverse, prec, n, batch, in.data(), want.data(), 4);
        std::fill(out.begin(), out.end(), (T)NAN);
        if (!core.run(in.data(), out.data(), batch, inverse)) { ++g_bad; printf("FAIL n=%zu: deadlock\n", n); return; }
        if (reference_is_inexact) check_sampled_exact(inverse ? "c2c inverse" : "c2c forward", n, batch, in.data(), out.data(), inverse);
        else check(inverse ? "c2c inverse" : "c2c forward", n, out.data(), want.data(), out.size());
    }
    // in place (the library accepts in == out)
    fill(in.data(), in.size(), 73);
    oracle_batch(0, prec, n, batch, in.data(), want.data(), 4);
    if (!core.four_step) {
        std::vector<T> orig = in;
        core.run(in.data(), in.data(), batch, 0);
        if (reference_is_inexact) check_sampled_exact("c2c in place", n, batch, orig.data(), in.data(), 0);
        else check("c2c in place", n, in.data(), want.data(), in.size());
    }
}

// RealFFT / ModifiedRealFFT of real length nr through the generic complex core and the stand-alone passes, as
// ssfft.cu:exec_r2c_typed / exec_c2r_typed do
template <typename T>
static void run_real(size_t nr, long long batch, int modified) {
    constexpr int prec = sizeof(T) == 4 ? 0 : 1;
    const size_t h = nr / 2;
    ComplexCore<T> core(h);
    std::vector<T> rtw(2 * (nr / 4 + 1)), rot(2 * (h ? h : 1));
    fill_real_twiddles<T>(rtw.data(), nr, modified != 0);
    fill_modified_rotations<T>(rot.data(), nr);
    std::vector<T> in(nr * batch), out(nr * batch, (T)NAN), want(nr * batch);
    auto blocks = [](long long items) { return dim3((unsigned)((items + 255) / 256)); };
    // forward
    fill(in.data(), in.size(), 75 + modified);
    oracle_batch(modified ? 4 : 2, prec, nr, batch, in.data(), want.data(), 4);
    const void *src = in.data();
    if (modified) {
        simt::launch(blocks((long long)h * batch), dim3(256), [&] {
            rotate_kernel<T>((cx<T> *)out.data(), (const cx<T> *)in.data(), (const cx<T> *)rot.data(), (long long)h, (long long)h * batch, 0);
        });
        src = out.data();
    }
    core.run(src, out.data(), batch, 0);
    simt::launch(blocks((long long)(h / 2 + 1) * batch), dim3(256), [&] {
        r2c_post_kernel<T>((cx<T> *)out.data(), (const cx<T> *)rtw.data(), (long long)h, batch, modified);
    });
    check(modified ? "modified r2c" : "r2c", nr, out.data(), want.data(), out.size());
    // inverse
    fill(in.data(), in.size(), 77 + modified);
    oracle_batch(modified ? 5 : 3, prec, nr, batch, in.data(), want.data(), 4);
    std::fill(out.begin(), out.end(), (T)NAN);
    simt::launch(blocks((long long)(h / 2 + 1) * batch), dim3(256), [&] {
        c2r_pre_kernel<T>((const cx<T> *)in.data(), (cx<T> *)out.data(), (const cx<T> *)rtw.data(), (long long)h, batch, modified);
    });
    core.run(out.data(), out.data(), batch, 1);
    if (modified)
        simt::launch(blocks((long long)h * batch), dim3(256), [&] {
            rotate_kernel<T>((cx<T> *)out.data(), (const cx<T> *)out.data(), (const cx<T> *)rot.data(), (long long)h, (long long)h * batch, 1);
        });
    check(modified ? "modified c2r" : "c2r", nr, out.data(), want.data(), out.size(), 2.0);
}

int main(int argc, char **argv) {
    const int part = argc > 1 ? atoi(argv[1]) : 0, parts = argc > 2 ? atoi(argv[2]) : 1;
    // the reference's own test sizes, primes with / without a codelet, mixed radices, one near the single-CTA limit
    const size_t direct[] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 28, 32, 49, 64, 98,
                             100, 128, 243, 256, 289, 360, 1001, 1100, 2310, 4199, 13000, 2 * 1031};
    // beyond one CTA: generic four-step (2^5 5^5, 3 2^15, 2^4 3 5^3 7, 17 2^12, a length whose split is very uneven)
    const size_t split[] = {100000, 98304, 42000, 69632, 27 * 1031};  // the last one: prime factor 1031 (see check_sampled_exact)
    const size_t real[] = {2, 4, 6, 10, 30, 98, 250, 1100, 2002, 200000};
    int k = 0;
    for (size_t n : direct) if (k++ % parts == part) { run_complex<float>(n, n < 64 ? 37 : 5); run_complex<double>(n, n < 64 ? 37 : 3); }
    for (size_t n : split) if (k++ % parts == part) { run_complex<float>(n, 2); run_complex<double>(n, 2); }
    for (size_t n : real) if (k++ % parts == part)
        for (int mod = 0; mod < 2; ++mod) { run_real<float>(n, 3, mod); run_real<double>(n, 3, mod); }
    printf("%d runs, %d failure(s)\n", g_runs, g_bad);
    printf(g_bad ? "GENERIC-EMUL-FAILED\n" : "GENERIC-EMUL-OK\n");
    return g_bad != 0;
}
