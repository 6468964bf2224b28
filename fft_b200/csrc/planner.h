// planner.h -- host-side plan builder: factorisation, pass radices, four-step split, twiddle tables.
//
// Replaces FFT<V>::setPlan / addPlanSteps (signalsmith-fft.h:93-185).  Same first step (ascending
// trial-division factorisation, :140-151); the rest is re-designed for the GPU: prime factors are
// grouped into register-sized radices (16/8/4/2, 9/3, 5, 7, 11, 13, anything else "generic"), the
// permutation table disappears (folded into Stockham indexing), twiddles are exact-phase double
// cos/sin rounded once to V (the reference rounds the PHASE to V first, :123), and sizes too big for
// one CTA's shared memory are split N = N1 * N2 (four-step) -- the GPU analogue of the reference's
// 64 KiB cache-blocking rule (:130-133).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <vector>

namespace ssfft {

// FFT<V>::sizeMinimum / sizeMaximum  signalsmith-fft.h:317-348.  Fast sizes are 2^k * {1, 3, 9}.
inline bool fast_small(size_t s) {
    switch (s) {
        case 0: case 1: case 2: case 3: case 4: case 6: case 8: case 9: case 12: case 16: case 18: case 24:
            return true;
        default:
            return false;
    }
}
inline size_t size_minimum(size_t size) {
    size_t power2 = 1;
    while (size >= 32) { size = (size - 1) / 2 + 1; power2 *= 2; }
    while (size < 32 && !fast_small(size)) ++size;
    return power2 * size;
}
inline size_t size_maximum(size_t size) {
    size_t power2 = 1;
    while (size >= 32) { size /= 2; power2 *= 2; }
    while (size > 1 && !fast_small(size)) --size;
    return power2 * size;
}
// RealFFT<V>::sizeMinimum / sizeMaximum  :403-408 -- reproduced including their quirks
inline size_t real_size_minimum(size_t size) { return (size_minimum((size - 1) / 2) + 1) * 2; }
inline size_t real_size_maximum(size_t size) { return size_minimum(size / 2) * 2; }

// ascending prime factors by trial division, remaining prime once f > sqrt(size)  (:140-151)
inline std::vector<size_t> factorise(size_t size) {
    std::vector<size_t> f;
    size_t d = 2;
    while (size > 1) {
        if (size % d == 0) { f.push_back(d); size /= d; }
        else if ((double)d > std::sqrt((double)size)) d = size;
        else ++d;
    }
    return f;
}

// Group prime factors into pass radices for the generic kernel (order = execution order).
inline std::vector<int> choose_radices(size_t n) {
    std::vector<size_t> f = factorise(n);
    int c2 = 0, c3 = 0;
    std::vector<int> rest;  // primes >= 5, ascending
    for (size_t p : f) {
        if (p == 2) ++c2;
        else if (p == 3) ++c3;
        else rest.push_back((int)p);
    }
    std::vector<int> r;
    while (c2 >= 4) { r.push_back(16); c2 -= 4; }
    if (c2 == 3) r.push_back(8);
    if (c2 == 2) r.push_back(4);
    if (c2 == 1) r.push_back(2);
    while (c3 >= 2) { r.push_back(9); c3 -= 2; }
    if (c3 == 1) r.push_back(3);
    for (int p : rest) r.push_back(p);  // 5, 7, 11, 13 have codelets; larger primes run the O(p^2) pass last
    if (r.empty()) r.push_back(1);      // n == 1: a copy
    return r;
}

// roots table W_n^k = exp(-2 pi i k / n), k < n, exact phase in double (octant-reduced), rounded once.
template <typename T>
inline void fill_roots(T *dst_interleaved, size_t n, size_t count, size_t step = 1) {
    // entry k holds W_n^(k*step)
    for (size_t k = 0; k < count; ++k) {
        // reduce k*step mod n exactly, then use symmetry around the octants for accuracy
        unsigned long long q = (unsigned long long)((__uint128_t)k * step % n);
        long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)q / (long double)n;
        dst_interleaved[2 * k] = (T)cosl(a);
        dst_interleaved[2 * k + 1] = (T)(-sinl(a));
    }
}

// RealFFT twiddles  signalsmith-fft.h:420-425: tw[i] = (sin phi, -cos phi), phi = -2 pi (i [+ 0.5]) / N
template <typename T>
inline void fill_real_twiddles(T *dst_interleaved, size_t n_real, bool modified) {
    size_t hh = n_real / 4 + 1;
    for (size_t i = 0; i < hh; ++i) {
        long double phi = -2.0L * 3.14159265358979323846264338327950288L * ((long double)i + (modified ? 0.5L : 0.0L)) /
                          (long double)n_real;
        dst_interleaved[2 * i] = (T)sinl(phi);
        dst_interleaved[2 * i + 1] = (T)(-cosl(phi));
    }
}
// ModifiedRealFFT rotations  :426-432: rot[i] = exp(-2 pi i * i / N), i < N/2
template <typename T>
inline void fill_modified_rotations(T *dst_interleaved, size_t n_real) {
    for (size_t i = 0; i < n_real / 2; ++i) {
        long double phi = -2.0L * 3.14159265358979323846264338327950288L * (long double)i / (long double)n_real;
        dst_interleaved[2 * i] = (T)cosl(phi);
        dst_interleaved[2 * i + 1] = (T)sinl(phi);
    }
}

// Four-step split n = n1 * n2 with both factors <= limit; prefers balanced factors whose own prime
// factors are small.  Returns false when no split fits (a prime factor > limit).
inline bool choose_split(size_t n, size_t limit, size_t *n1_out, size_t *n2_out) {
    std::vector<size_t> f = factorise(n);
    for (size_t p : f)
        if (p > limit) return false;
    // greedy: build n1 from the largest primes down while staying <= sqrt-ish target and <= limit
    double target = std::sqrt((double)n);
    size_t best1 = 0;
    double best_score = 1e300;
    // enumerate divisors (n has few prime factors; divisor count is small for practical sizes)
    std::vector<size_t> primes;
    std::vector<int> expo;
    for (size_t p : f) {
        if (!primes.empty() && primes.back() == p) ++expo.back();
        else { primes.push_back(p); expo.push_back(1); }
    }
    std::vector<size_t> divs{1};
    for (size_t i = 0; i < primes.size(); ++i) {
        size_t cur_n = divs.size(), mult = 1;
        for (int k = 1; k <= expo[i]; ++k) {
            mult *= primes[i];
            for (size_t j = 0; j < cur_n; ++j) divs.push_back(divs[j] * mult);
        }
    }
    for (size_t d : divs) {
        size_t e = n / d;
        if (d > limit || e > limit) continue;
        double score = std::fabs(std::log((double)d / target));
        if (score < best_score) { best_score = score; best1 = d; }
    }
    if (!best1) return false;
    *n1_out = best1;
    *n2_out = n / best1;
    return true;
}

}  // namespace ssfft
