// codelets.cuh -- register-resident forward DFTs of compile-time length R (natural order in and out).
//
// These replace the reference's per-step scalar butterflies: fftStep2 (signalsmith-fft.h:217-233),
// fftStep3 (:235-259), fftStep4 (:261-286) and fftStepGeneric (:187-215).  Differences by design:
//   * radix-5 (and 7, 11, 13) get real butterflies with compile-time constants -- the reference runs
//     every prime >= 5 through the O(p^2) generic step with cos/sin in the inner loop (:204-205);
//   * composite radices (8, 9, 16, 25, 27, ...) are built at compile time by Cooley-Tukey recursion
//     in registers so a transform needs fewer shared-memory exchanges;
//   * only the forward transform exists: the inverse is done by swapping re/im on load and store
//     (ifft(x) = swap(fft(swap(x)))), which is free on a GPU.
// Forward sign convention: X[k] = sum_n x[n] exp(-2 pi i n k / R), unnormalised (tests/00-fft.cpp:35-40).
#pragma once
#include "cplx.cuh"

namespace ssfft {

template <int R>
struct Dft;

template <>
struct Dft<1> {
    template <typename T> static SSFFT_HD void run(cx<T> (&)[1]) {}
};

template <>
struct Dft<2> {
    template <typename T> static SSFFT_HD void run(cx<T> (&v)[2]) {
        cx<T> a = v[0], b = v[1];
        v[0] = a + b;
        v[1] = a - b;
    }
};

template <>
struct Dft<4> {
    template <typename T> static SSFFT_HD void run(cx<T> (&v)[4]) {
        cx<T> s02 = v[0] + v[2], d02 = v[0] - v[2];
        cx<T> s13 = v[1] + v[3], d13 = v[1] - v[3];
        v[0] = s02 + s13;
        v[1] = sub_i(d02, d13);  // d02 - i d13
        v[2] = s02 - s13;
        v[3] = add_i(d02, d13);  // d02 + i d13
    }
};

// Odd prime P: pair x[j] with x[P-j]; (P-1)/2 cosine sums and sine sums with compile-time constants.
template <int P>
struct DftOddPrime {
    template <typename T> static SSFFT_HD void run(cx<T> (&v)[P]) {
        constexpr int H = (P - 1) / 2;
        cx<T> tp[H], tm[H];
        sfor<0, H>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            tp[j] = v[j + 1] + v[P - 1 - j];
            tm[j] = v[j + 1] - v[P - 1 - j];
        });
        cx<T> x0 = v[0];
        cx<T> sum = x0;
        sfor<0, H>([&](auto jc) { sum = sum + tp[decltype(jc)::value]; });
        v[0] = sum;
        sfor<1, H + 1>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            cx<T> a = x0, b = mk<T>((T)0, (T)0);
            sfor<0, H>([&](auto jc) {
                constexpr int j = decltype(jc)::value + 1;
                constexpr ct::cs w = ct::cossin2pi((long long)j * k, P);
                constexpr T c = (T)w.c, s = (T)w.s;
                a.x += c * tp[j - 1].x;
                a.y += c * tp[j - 1].y;
                if constexpr (j == 1) {
                    b.x = s * tm[0].x;
                    b.y = s * tm[0].y;
                } else {
                    b.x += s * tm[j - 1].x;
                    b.y += s * tm[j - 1].y;
                }
            });
            v[k] = sub_i(a, b);      // a - i b
            v[P - k] = add_i(a, b);  // a + i b
        });
    }
};

// Composite R = A * B, Cooley-Tukey in registers:  n = A*n2 + n1,  k = B*k1 + k2.
template <int R>
struct Dft {
    template <typename T> static SSFFT_HD void run(cx<T> (&v)[R]) {
        if constexpr (ct::is_prime(R)) {
            DftOddPrime<R>::run(v);
        } else {
            constexpr int A = (R % 4 == 0 && R > 4) ? 4 : ct::smallest_factor(R);
            constexpr int B = R / A;
            cx<T> y[R];
            sfor<0, A>([&](auto n1c) {
                constexpr int n1 = decltype(n1c)::value;
                cx<T> t[B];
                sfor<0, B>([&](auto n2c) { constexpr int n2 = decltype(n2c)::value; t[n2] = v[A * n2 + n1]; });
                Dft<B>::run(t);
                sfor<0, B>([&](auto k2c) {
                    constexpr int k2 = decltype(k2c)::value;
                    y[n1 * B + k2] = mul_root<n1 * k2, R>(t[k2]);
                });
            });
            sfor<0, B>([&](auto k2c) {
                constexpr int k2 = decltype(k2c)::value;
                cx<T> t[A];
                sfor<0, A>([&](auto n1c) { constexpr int n1 = decltype(n1c)::value; t[n1] = y[n1 * B + k2]; });
                Dft<A>::run(t);
                sfor<0, A>([&](auto k1c) { constexpr int k1 = decltype(k1c)::value; v[B * k1 + k2] = t[k1]; });
            });
        }
    }
};

}  // namespace ssfft
