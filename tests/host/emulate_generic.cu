// Host emulation of generic_fft_kernel's pass schedule (no GPU needed): the pass bodies in generic.cuh are
// __host__ __device__, so running them for tx = 0..TX-1 sequentially, pass by pass, is exactly what the
// kernel computes between __syncthreads().  Checked against a long-double naive DFT.
#include "../../fft_b200/csrc/generic.cuh"
#include "../../fft_b200/csrc/planner.h"
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace ssfft;

template <typename T>
double run_case(int n, int inverse) {
    std::vector<int> radices = choose_radices(n);
    std::vector<cx<T>> roots(n), in(n), out(n), bufA(spad(n) + 2), bufB(spad(n) + 2);
    fill_roots<T>((T *)roots.data(), n, n);
    std::vector<std::complex<long double>> x(n);
    for (int i = 0; i < n; ++i) {
        x[i] = {(long double)(rand() / (double)RAND_MAX - 0.5), (long double)(rand() / (double)RAND_MAX - 0.5)};
        in[i] = mk<T>((T)x[i].real(), (T)x[i].imag());
        x[i] = {(long double)in[i].x, (long double)in[i].y};
    }
    const int TX = 7;  // deliberately odd
    int npass = (int)radices.size();
    bool stage = !radix_has_codelet(radices[0]);
    GlobalSrc<T> gsrc{in.data(), 1, inverse};
    GlobalDst<T> gdst{out.data(), 1, inverse, nullptr, nullptr, 0, 0};
    cx<T> *cur = bufA.data(), *nxt = bufB.data();
    if (stage) for (int e = 0; e < n; ++e) cur[spad(e)] = gsrc.load(e);
    int P = 1;
    for (int i = 0; i < npass; ++i) {
        bool first = (i == 0) && !stage, last = (i == npass - 1);
        for (int tx = 0; tx < TX; ++tx) {
            if (first && last) run_pass<T>(radices[i], n, P, roots.data(), gsrc, gdst, tx, TX);
            else if (first) run_pass<T>(radices[i], n, P, roots.data(), gsrc, SharedDst<T>{nxt}, tx, TX);
            else if (last) run_pass<T>(radices[i], n, P, roots.data(), SharedSrc<T>{cur}, gdst, tx, TX);
            else run_pass<T>(radices[i], n, P, roots.data(), SharedSrc<T>{cur}, SharedDst<T>{nxt}, tx, TX);
        }
        if (!last) std::swap(cur, nxt);
        P *= radices[i];
    }
    long double err = 0, nrm = 0;
    const long double tau = 6.283185307179586476925286766559L;
    for (int k = 0; k < n; ++k) {
        std::complex<long double> s = 0;
        for (int j = 0; j < n; ++j) {
            long long q = ((long long)j * k) % n;
            long double a = tau * q / n;
            s += x[j] * std::complex<long double>(cosl(a), inverse ? sinl(a) : -sinl(a));
        }
        err += std::norm(s - std::complex<long double>(out[k].x, out[k].y));
        nrm += std::norm(s);
    }
    return (double)sqrtl(err / nrm);
}

int main() {
    int sizes[] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 28, 32,
                   49, 64, 98, 100, 128, 243, 256, 289, 360, 512, 625, 1000, 1024, 2187, 3125};
    int bad = 0;
    for (int n : sizes)
        for (int inv = 0; inv < 2; ++inv) {
            double ef = run_case<float>(n, inv), ed = run_case<double>(n, inv);
            bool ok = ef < 5e-7 && ed < 1e-15 * (4 + n / 64);
            if (!ok) { ++bad; printf("FAIL "); }
            if (!ok || n >= 1000 || n == 17) printf("n=%d inv=%d f32 %.2e f64 %.2e\n", n, inv, ef, ed);
        }
    size_t a, b;
    size_t big[] = {32768, 65536, 1u << 20, 6000 * 7, 1000000, 3u << 18};
    for (size_t n : big) { bool ok = choose_split(n, 8192, &a, &b); printf("split %zu -> %d %zu x %zu\n", n, ok, a, b); }
    printf(bad ? "FAILED\n" : "ALL OK\n");
    return bad != 0;
}
