// ref_shim.cpp -- extern "C" shim around the UNMODIFIED reference header (TEST INFRASTRUCTURE ONLY).
//
// Compiled by oracle/Makefile with -I$(REFERENCE_DIR) so that "signalsmith-fft.h" is the file
// under /root/reference where it lies; the output goes to oracle/_ref/libssfft_ref.so (git-ignored,
// travels to the GPU box).  No reference source is copied into this repository.
//
// Used to (a) pin the C oracle (tests/test_oracle.py), (b) generate tests/golden/ fixtures, and
// (c) time the reference's own CPU path (bench.py cpu_baseline.kind == "reference").
// Flags follow the reference's benchmark build (Makefile:44): -O3 -msse2 -mavx, plus -Wno-narrowing
// because FFT<float> narrows at signalsmith-fft.h:124,205,424,430.
#define SIGNALSMITH_FFT_NAMESPACE signalsmith_ref
#include "signalsmith-fft.h"

#include <chrono>
#include <complex>
#include <cstddef>
#include <thread>
#include <vector>

namespace {

template <typename V>
void c2c_range(size_t n, const V *in, V *out, size_t b0, size_t b1, int inverse, double *seconds,
               int *ready, int *go) {
    signalsmith_ref::FFT<V> fft(n);
    __atomic_add_fetch(ready, 1, __ATOMIC_SEQ_CST);
    while (!__atomic_load_n(go, __ATOMIC_SEQ_CST)) std::this_thread::yield();
    auto t0 = std::chrono::steady_clock::now();
    for (size_t b = b0; b < b1; ++b) {
        const std::complex<V> *src = reinterpret_cast<const std::complex<V> *>(in) + b * n;
        std::complex<V> *dst = reinterpret_cast<std::complex<V> *>(out) + b * n;
        if (inverse) fft.ifft(src, dst);
        else fft.fft(src, dst);
    }
    *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

template <typename V, int FLAGS>
void real_range(size_t n, const V *in, V *out, size_t b0, size_t b1, int inverse, double *seconds,
                int *ready, int *go) {
    signalsmith_ref::RealFFT<V, FLAGS> fft(n);
    const size_t nr = fft.size();
    __atomic_add_fetch(ready, 1, __ATOMIC_SEQ_CST);
    while (!__atomic_load_n(go, __ATOMIC_SEQ_CST)) std::this_thread::yield();
    auto t0 = std::chrono::steady_clock::now();
    for (size_t b = b0; b < b1; ++b) {
        if (!inverse) {
            const V *src = in + b * nr;
            std::complex<V> *dst = reinterpret_cast<std::complex<V> *>(out) + b * (nr / 2);
            fft.fft(src, dst);
        } else {
            const std::complex<V> *src = reinterpret_cast<const std::complex<V> *>(in) + b * (nr / 2);
            V *dst = out + b * nr;
            fft.ifft(src, dst);
        }
    }
    *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

template <typename V>
double batch_typed(int kind, size_t n, size_t batch, const V *in, V *out, int threads) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > batch) threads = batch ? (int)batch : 1;
    std::vector<std::thread> pool;
    std::vector<double> secs(threads, 0.0);
    int ready = 0, go = 0;
    for (int t = 0; t < threads; ++t) {
        size_t b0 = batch * (size_t)t / (size_t)threads, b1 = batch * (size_t)(t + 1) / (size_t)threads;
        switch (kind) {
            case 0: pool.emplace_back(c2c_range<V>, n, in, out, b0, b1, 0, &secs[t], &ready, &go); break;
            case 1: pool.emplace_back(c2c_range<V>, n, in, out, b0, b1, 1, &secs[t], &ready, &go); break;
            case 2: pool.emplace_back(real_range<V, 0>, n, in, out, b0, b1, 0, &secs[t], &ready, &go); break;
            case 3: pool.emplace_back(real_range<V, 0>, n, in, out, b0, b1, 1, &secs[t], &ready, &go); break;
            case 4: pool.emplace_back(real_range<V, 1>, n, in, out, b0, b1, 0, &secs[t], &ready, &go); break;
            case 5: pool.emplace_back(real_range<V, 1>, n, in, out, b0, b1, 1, &secs[t], &ready, &go); break;
        }
    }
    while (__atomic_load_n(&ready, __ATOMIC_SEQ_CST) < threads) std::this_thread::yield();
    __atomic_store_n(&go, 1, __ATOMIC_SEQ_CST);  // plans built; start every thread together
    double worst = 0;
    for (int t = 0; t < threads; ++t) {
        pool[t].join();
        if (secs[t] > worst) worst = secs[t];
    }
    return worst;
}

}  // namespace

extern "C" {

// Same contract as oracle_batch() in oracle_fft.h, executed by the reference's own classes.
double ref_batch(int kind, int prec, size_t n, size_t batch, const void *in, void *out, int threads) {
    if (kind < 0 || kind > 5) return -1.0;
    if (prec == 0) return batch_typed<float>(kind, n, batch, (const float *)in, (float *)out, threads);
    if (prec == 1) return batch_typed<double>(kind, n, batch, (const double *)in, (double *)out, threads);
    return -1.0;
}

size_t ref_fft_size_minimum(size_t n) { return signalsmith_ref::FFT<double>::sizeMinimum(n); }
size_t ref_fft_size_maximum(size_t n) { return signalsmith_ref::FFT<double>::sizeMaximum(n); }
size_t ref_realfft_size_minimum(size_t n) { return signalsmith_ref::RealFFT<double>::sizeMinimum(n); }
size_t ref_realfft_size_maximum(size_t n) { return signalsmith_ref::RealFFT<double>::sizeMaximum(n); }
// RealFFT quirks: setSize() returns N/2, size() returns 2*(N/2)
size_t ref_realfft_setsize_return(size_t n) { signalsmith_ref::RealFFT<double> r(2); return r.setSize(n); }
size_t ref_realfft_size(size_t n) { signalsmith_ref::RealFFT<double> r(n); return r.size(); }
int ref_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
