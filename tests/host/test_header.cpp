// test_header.cpp -- exercises include/signalsmith-fft.h the way a C++ user of the reference would,
// plus the new batched device-pointer overloads, and checks results against the oracle (checker only).
// Build: g++ -std=c++11 tests/host/test_header.cpp -Iinclude -Ioracle -Lfft_b200 -lssfft -Loracle -loracle
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <list>
#include <vector>

#include "signalsmith-fft.h"
#include "oracle_fft.h"

static int failures = 0;
#define CHECK(cond, ...) do { if (!(cond)) { ++failures; printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } } while (0)

template <typename V> struct Oracle;
template <> struct Oracle<float> {
    static void fft(size_t n, const std::complex<float> *in, std::complex<float> *out, int inv) {
        oracle_plan_f32 *p = oracle_plan_create_f32(n); oracle_fft_f32(p, (const float *)in, (float *)out, inv); oracle_plan_destroy_f32(p);
    }
    static void rfft(size_t n, const float *in, std::complex<float> *out, int mod) {
        oracle_rplan_f32 *p = oracle_rplan_create_f32(n, mod); oracle_rfft_f32(p, in, (float *)out); oracle_rplan_destroy_f32(p);
    }
    static void fill(float *d, size_t c, uint64_t seed) { oracle_fill_uniform_f32(d, c, seed, 0); }
    static double tol(size_t n) { return 1e-6 * std::max(1.0, std::log2((double)n)); }
};
template <> struct Oracle<double> {
    static void fft(size_t n, const std::complex<double> *in, std::complex<double> *out, int inv) {
        oracle_plan_f64 *p = oracle_plan_create_f64(n); oracle_fft_f64(p, (const double *)in, (double *)out, inv); oracle_plan_destroy_f64(p);
    }
    static void rfft(size_t n, const double *in, std::complex<double> *out, int mod) {
        oracle_rplan_f64 *p = oracle_rplan_create_f64(n, mod); oracle_rfft_f64(p, in, (double *)out); oracle_rplan_destroy_f64(p);
    }
    static void fill(double *d, size_t c, uint64_t seed) { oracle_fill_uniform_f64(d, c, seed, 0); }
    static double tol(size_t n) { return 1e-14 * std::max(1.0, std::log2((double)n)); }
};

template <typename V>
double relL2(const std::complex<V> *a, const std::complex<V> *b, size_t n) {
    double num = 0, den = 0;
    for (size_t i = 0; i < n; ++i) { num += std::norm(std::complex<double>(a[i]) - std::complex<double>(b[i])); den += std::norm(std::complex<double>(b[i])); }
    return den > 0 ? std::sqrt(num / den) : std::sqrt(num);
}

template <typename V>
void testType(const char *name) {
    using cplx = std::complex<V>;
    const size_t sizes[] = {1, 2, 3, 5, 8, 12, 49, 256, 1000, 4096, 6000};
    for (size_t n : sizes) {
        std::vector<cplx> in(n), out(n), ref(n), back(n);
        Oracle<V>::fill((V *)in.data(), 2 * n, 3);
        std::vector<cplx> inCopy = in;
        signalsmith::FFT<V> fft(n);
        fft.fft(in, out);                                   // containers
        CHECK(in == inCopy, "%s n=%zu input changed", name, n);
        Oracle<V>::fft(n, in.data(), ref.data(), 0);
        CHECK(relL2(out.data(), ref.data(), n) <= Oracle<V>::tol(n), "%s n=%zu fwd err %g", name, n, relL2(out.data(), ref.data(), n));
        fft.ifft(out.data(), back.begin());                 // pointer in, iterator out
        double e = 0, d = 0;
        for (size_t i = 0; i < n; ++i) { e += std::norm(std::complex<double>(back[i]) - (double)n * std::complex<double>(in[i])); d += std::norm((double)n * std::complex<double>(in[i])); }
        CHECK(std::sqrt(e / d) <= 2 * Oracle<V>::tol(n), "%s n=%zu roundtrip err %g", name, n, std::sqrt(e / d));
    }
    // copies own their plan (reference objects are independent after a copy); setSize re-plans only on change
    signalsmith::FFT<V> a(96), b = a;
    CHECK(b.size() == 96 && a.setSize(96) == 96 && a.setSize(64) == 64 && b.size() == 96, "copy/setSize");
    {
        std::vector<signalsmith::FFT<V>> pool(3, signalsmith::FFT<V>(256));  // the pattern of a thread pool
        std::vector<cplx> in(256), o0(256), o2(256);
        Oracle<V>::fill((V *)in.data(), 512, 77);
        pool[0].fft(in, o0);
        pool[2].fft(in, o2);
        CHECK(o0 == o2 && pool[1].size() == 256, "copies of an FFT object are independent plans");
        signalsmith::FFT<V> moved = std::move(pool[1]);
        moved.fft(in, o2);
        CHECK(o0 == o2, "moved-from plan handle");
    }
    {
        // one transform sharded over logical ranks on device 0 (DistributedFFT, C ABI ssfft_dist_*)
        const size_t dn = 1 << 14, P = 2, per = dn / P;
        std::vector<cplx> x(dn), y(dn), ref(dn);
        Oracle<V>::fill((V *)x.data(), 2 * dn, 78);
        signalsmith::DistributedFFT<V> dfft(dn, std::vector<int>(P, 0));
        void *din[P], *dout[P];
        for (size_t r = 0; r < P; ++r) {
            CHECK(ssfft_malloc(&din[r], per * sizeof(cplx)) == 0 && ssfft_malloc(&dout[r], per * sizeof(cplx)) == 0, "malloc");
            ssfft_memcpy_h2d(din[r], x.data() + r * per, per * sizeof(cplx), nullptr);
        }
        ssfft_stream_synchronize(nullptr);
        dfft.fft((cplx *const *)din, (cplx *const *)dout);
        dfft.synchronize();
        for (size_t r = 0; r < P; ++r) ssfft_memcpy_d2h(y.data() + r * per, dout[r], per * sizeof(cplx), nullptr);
        ssfft_stream_synchronize(nullptr);
        Oracle<V>::fft(dn, x.data(), ref.data(), 0);
        CHECK(relL2(y.data(), ref.data(), dn) <= Oracle<V>::tol(dn), "DistributedFFT err %g", relL2(y.data(), ref.data(), dn));
        for (size_t r = 0; r < P; ++r) { ssfft_free(din[r]); ssfft_free(dout[r]); }
    }
    CHECK(signalsmith::FFT<V>(1000, 1).size() == 1024 && signalsmith::FFT<V>(1000, -1).size() == 768, "fastDirection");
    CHECK(a.setSizeMinimum(1025) == 1152 && a.setSizeMaximum(1025) == 1024, "setSizeMinimum/Maximum");

    // batched device-pointer overloads
    const size_t n = 4096, batch = 19;
    std::vector<cplx> h(n * batch), hOut(n * batch), ref(n);
    Oracle<V>::fill((V *)h.data(), 2 * n * batch, 5);
    void *dIn = nullptr, *dOut = nullptr;
    CHECK(ssfft_malloc(&dIn, n * batch * sizeof(cplx)) == 0 && ssfft_malloc(&dOut, n * batch * sizeof(cplx)) == 0, "malloc");
    ssfft_memcpy_h2d(dIn, h.data(), n * batch * sizeof(cplx), nullptr);
    signalsmith::FFT<V> big(n);
    big.fft((const cplx *)dIn, (cplx *)dOut, batch);
    ssfft_memcpy_d2h(hOut.data(), dOut, n * batch * sizeof(cplx), nullptr);
    ssfft_stream_synchronize(nullptr);
    for (size_t bi : {size_t(0), size_t(7), batch - 1}) {
        Oracle<V>::fft(n, h.data() + bi * n, ref.data(), 0);
        CHECK(relL2(hOut.data() + bi * n, ref.data(), n) <= Oracle<V>::tol(n), "%s device batch %zu", name, bi);
    }
    ssfft_free(dIn); ssfft_free(dOut);

    // RealFFT / ModifiedRealFFT: vector<V> in, vector<complex> out of FULL length (upper half must stay untouched)
    for (size_t nr : {size_t(2), size_t(6), size_t(98), size_t(1024), size_t(1000)}) {
        std::vector<V> x(nr), xb(nr);
        Oracle<V>::fill(x.data(), nr, 9);
        std::vector<cplx> spec(nr, cplx(123, 456)), sref(nr / 2);
        signalsmith::RealFFT<V> r(nr);
        CHECK(r.size() == nr && r.setSize(nr) == nr / 2, "RealFFT size quirk");
        r.fft(x, spec);
        Oracle<V>::rfft(nr, x.data(), sref.data(), 0);
        CHECK(relL2(spec.data(), sref.data(), nr / 2) <= Oracle<V>::tol(nr), "%s real n=%zu", name, nr);
        for (size_t i = nr / 2; i < nr; ++i) CHECK(spec[i] == cplx(123, 456), "upper half touched");
        r.ifft(spec, xb);
        double e = 0, d = 0;
        for (size_t i = 0; i < nr; ++i) { e += std::pow((double)xb[i] - (double)nr * x[i], 2); d += std::pow((double)nr * x[i], 2); }
        CHECK(std::sqrt(e / d) <= 2 * Oracle<V>::tol(nr), "%s real roundtrip n=%zu", name, nr);
        signalsmith::ModifiedRealFFT<V> m(nr);
        std::vector<cplx> mspec(nr / 2), mref(nr / 2);
        m.fft(x, mspec);
        Oracle<V>::rfft(nr, x.data(), mref.data(), 1);
        CHECK(relL2(mspec.data(), mref.data(), nr / 2) <= Oracle<V>::tol(nr), "%s modified real n=%zu", name, nr);
    }
    // extended overloads: an STFT straight out of a signal (overlapping frames, window applied on load) ...
    {
        const size_t nr = 1024, hop = 257, frames = 21, len = (frames - 1) * hop + nr;
        std::vector<V> sig(len), win(nr), frame(nr);
        Oracle<V>::fill(sig.data(), len, 21);
        Oracle<V>::fill(win.data(), nr, 22);
        for (auto &w : win) w += V(1);
        std::vector<cplx> spec(frames * nr / 2), sref(nr / 2);
        void *dSig = nullptr, *dWin = nullptr, *dSpec = nullptr;
        ssfft_malloc(&dSig, len * sizeof(V)); ssfft_malloc(&dWin, nr * sizeof(V)); ssfft_malloc(&dSpec, spec.size() * sizeof(cplx));
        ssfft_memcpy_h2d(dSig, sig.data(), len * sizeof(V), nullptr);
        ssfft_memcpy_h2d(dWin, win.data(), nr * sizeof(V), nullptr);
        ssfft_io io = ssfft_io();
        io.in_dist = (int64_t)hop; io.pre = dWin; io.pre_kind = SSFFT_MUL_REAL;
        signalsmith::RealFFT<V> r(nr);
        r.fft((const V *)dSig, (cplx *)dSpec, frames, io);
        ssfft_memcpy_d2h(spec.data(), dSpec, spec.size() * sizeof(cplx), nullptr);
        ssfft_stream_synchronize(nullptr);
        for (size_t f : {size_t(0), size_t(10), frames - 1}) {
            for (size_t i = 0; i < nr; ++i) frame[i] = sig[f * hop + i] * win[i];
            Oracle<V>::rfft(nr, frame.data(), sref.data(), 0);
            CHECK(relL2(spec.data() + f * nr / 2, sref.data(), nr / 2) <= 2 * Oracle<V>::tol(nr), "%s stft frame %zu", name, f);
        }
        // ... and the column pass of a 2-D transform, in place on a row-major [rows][cols] matrix
        const size_t rows = 256, cols = 24;
        std::vector<cplx> m(rows * cols), mOut(rows * cols), col(rows), cref(rows);
        Oracle<V>::fill((V *)m.data(), 2 * rows * cols, 23);
        void *dM = nullptr;
        ssfft_malloc(&dM, m.size() * sizeof(cplx));
        ssfft_memcpy_h2d(dM, m.data(), m.size() * sizeof(cplx), nullptr);
        ssfft_io cio = ssfft_io();
        cio.in_stride = cio.out_stride = (int64_t)cols; cio.in_dist = cio.out_dist = 1;
        signalsmith::FFT<V> cfft(rows);
        cfft.fft((const cplx *)dM, (cplx *)dM, cols, cio);
        ssfft_memcpy_d2h(mOut.data(), dM, m.size() * sizeof(cplx), nullptr);
        ssfft_stream_synchronize(nullptr);
        for (size_t c : {size_t(0), size_t(11), cols - 1}) {
            for (size_t i = 0; i < rows; ++i) col[i] = m[i * cols + c];
            Oracle<V>::fft(rows, col.data(), cref.data(), 0);
            for (size_t i = 0; i < rows; ++i) col[i] = mOut[i * cols + c];
            CHECK(relL2(col.data(), cref.data(), rows) <= Oracle<V>::tol(rows), "%s column %zu", name, c);
        }
        bool threw = false;
        cio.out_dist = 0; cio.out_stride = 1;  // in place with a different output layout: rejected
        try { cfft.fft((const cplx *)dM, (cplx *)dM, cols, cio); } catch (const std::runtime_error &) { threw = true; }
        CHECK(threw, "invalid extended request must throw");
        ssfft_free(dSig); ssfft_free(dWin); ssfft_free(dSpec); ssfft_free(dM);
    }
    CHECK(signalsmith::RealFFT<V>(7).size() == 6, "odd real size truncates");
    CHECK(signalsmith::RealFFT<V>::sizeMinimum(256) == 258 && signalsmith::RealFFT<V>::sizeMaximum(1000) == 1024, "RealFFT size helper quirks");
}

int main() {
    try {
        testType<float>("float");
        testType<double>("double");
    } catch (const std::exception &e) {
        printf("EXCEPTION: %s\n", e.what());
        return 2;
    }
    printf(failures ? "HEADER-TESTS FAILED (%d)\n" : "HEADER-TESTS OK\n", failures);
    return failures ? 1 : 0;
}
