// cplx.cuh -- interleaved complex value type + constexpr trigonometry for compile-time twiddles.
//
// Replaces the scalar helpers of the reference (perf::complexMul / perf::complexAddI,
// signalsmith-fft.h:30-53) for device code.  Layout is the reference's: std::complex<V> is
// (re, im) interleaved, 8 B for float / 16 B for double, so a cx<T>* aliases a std::complex<T>*.
#pragma once
#include <cuda_runtime.h>
#include <type_traits>

#define SSFFT_HD __host__ __device__ __forceinline__

// The dynamic shared memory of a CTA.  Under SSFFT_EMUL (CPU execution of the kernels, tests/host/simt/) several CTAs
// of a cluster are alive at once, so the name is a pointer to the current CTA's buffer instead of one array.
#ifdef SSFFT_EMUL
#define SSFFT_DYNAMIC_SMEM(name) unsigned char *name = simt::dynamic_smem()
#else
#define SSFFT_DYNAMIC_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#endif

namespace ssfft {

template <typename T>
struct __align__(2 * sizeof(T)) cx {
    T x, y;
};

template <typename T> SSFFT_HD cx<T> mk(T x, T y) { cx<T> r; r.x = x; r.y = y; return r; }
template <typename T> SSFFT_HD cx<T> operator+(cx<T> a, cx<T> b) { return mk<T>(a.x + b.x, a.y + b.y); }
template <typename T> SSFFT_HD cx<T> operator-(cx<T> a, cx<T> b) { return mk<T>(a.x - b.x, a.y - b.y); }
// a * b  (4 mul + 2 add, contracted to 2 mul + 2 fma by nvcc)
template <typename T> SSFFT_HD cx<T> cmul(cx<T> a, cx<T> b) {
    return mk<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
template <typename T> SSFFT_HD cx<T> cmulc(cx<T> a, cx<T> b) {
    return mk<T>(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
template <typename T> SSFFT_HD cx<T> cconj(cx<T> a) { return mk<T>(a.x, -a.y); }
template <typename T> SSFFT_HD cx<T> cswap(cx<T> a) { return mk<T>(a.y, a.x); }
// a - i*b  and  a + i*b
template <typename T> SSFFT_HD cx<T> sub_i(cx<T> a, cx<T> b) { return mk<T>(a.x + b.y, a.y - b.x); }
template <typename T> SSFFT_HD cx<T> add_i(cx<T> a, cx<T> b) { return mk<T>(a.x - b.y, a.y + b.x); }

// ---------------------------------------------------------------------------------------------
// Blackwell packed fp32: add / mul / fma.f32x2 (SASS FADD2 / FMUL2 / FFMA2) work on an aligned register pair,
// i.e. on one interleaved complex value, with per-operand swap / negate / broadcast modifiers.  A complex
// add, a +-i rotation-and-add and a broadcast multiply are ONE instruction each instead of two, which is what
// would matter for a kernel bound by issue slots.  MEASURED SLOWER on B200 (tools/ubench_f32x2.cu: FADD2 issues at
// half the rate of FADD, so the fp32 pipe gains nothing, and pairing costs extra MOVs: the headline 4096 kernel went
// 86.5 % -> 81.7 % of the HBM roofline, 2048 90 -> 80 %, 2^20 28 -> 21 %; profiles/sweep_packed_f32x2_r01_float32.json).
// Kept behind -DSSFFT_USE_F32X2 as a recorded experiment; parity tests pass with it on.
// ---------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000) && defined(SSFFT_USE_F32X2)
#define SSFFT_F32X2 1
__device__ __forceinline__ float2 f2(cx<float> a) { return make_float2(a.x, a.y); }
__device__ __forceinline__ cx<float> c2(float2 a) { return mk<float>(a.x, a.y); }
__device__ __forceinline__ cx<float> operator+(cx<float> a, cx<float> b) { return c2(__fadd2_rn(f2(a), f2(b))); }
__device__ __forceinline__ cx<float> operator-(cx<float> a, cx<float> b) { return c2(__fadd2_rn(f2(a), make_float2(-b.x, -b.y))); }
__device__ __forceinline__ cx<float> cmul(cx<float> a, cx<float> b) {
    return c2(__ffma2_rn(make_float2(a.y, a.y), make_float2(-b.y, b.x), __fmul2_rn(make_float2(a.x, a.x), f2(b))));
}
__device__ __forceinline__ cx<float> cmulc(cx<float> a, cx<float> b) {  // a * conj(b)
    return c2(__ffma2_rn(make_float2(a.y, a.y), make_float2(b.y, b.x), __fmul2_rn(make_float2(a.x, a.x), make_float2(b.x, -b.y))));
}
__device__ __forceinline__ cx<float> sub_i(cx<float> a, cx<float> b) { return c2(__fadd2_rn(f2(a), make_float2(b.y, -b.x))); }
__device__ __forceinline__ cx<float> add_i(cx<float> a, cx<float> b) { return c2(__fadd2_rn(f2(a), make_float2(-b.y, b.x))); }
// v * (c - i s)
__device__ __forceinline__ cx<float> mul_cs(cx<float> v, float c, float s) {
    return c2(__ffma2_rn(make_float2(v.y, -v.x), make_float2(s, s), __fmul2_rn(f2(v), make_float2(c, c))));
}
__device__ __forceinline__ cx<float> scale2(cx<float> v, float h) { return c2(__fmul2_rn(f2(v), make_float2(h, h))); }
#endif

// ---------------------------------------------------------------------------------------------
// constexpr cos/sin of 2*pi*num/den, exact octant reduction on the rational, Taylor on [0, pi/4].
// Used only for compile-time butterfly constants (evaluated by the front end, never on the GPU).
// ---------------------------------------------------------------------------------------------
namespace ct {
constexpr double kPi = 3.14159265358979323846264338327950288;

__host__ __device__ constexpr double taylor_sin(double x) {  // |x| <= pi/4
    double x2 = x * x, term = x, sum = x;
    for (int k = 1; k < 14; ++k) {
        term *= -x2 / double((2 * k) * (2 * k + 1));
        sum += term;
    }
    return sum;
}
__host__ __device__ constexpr double taylor_cos(double x) {
    double x2 = x * x, term = 1.0, sum = 1.0;
    for (int k = 1; k < 14; ++k) {
        term *= -x2 / double((2 * k - 1) * (2 * k));
        sum += term;
    }
    return sum;
}
struct cs { double c, s; };
// cos and sin of 2*pi*num/den
__host__ __device__ constexpr cs cossin2pi(long long num, long long den) {
    num %= den;
    if (num < 0) num += den;
    // angle = 2*pi*num/den ; work in eighths of a turn: oct = floor(8*num/den)
    long long oct = (8 * num) / den;
    long long rem = 8 * num - oct * den;  // angle = (oct + rem/den) * pi/4
    double c = 0, s = 0;
    if (oct % 2 == 0) {
        double a = (double)rem / (double)den * (kPi / 4);  // in [0, pi/4)
        c = taylor_cos(a); s = taylor_sin(a);
    } else {
        double a = (double)(den - rem) / (double)den * (kPi / 4);  // pi/4 - frac, in (0, pi/4]
        // cos(pi/4*(1) - a') pattern: angle within octant measured from the next axis
        c = taylor_sin(a); s = taylor_cos(a);
    }
    // (c, s) is for the angle folded into the first quadrant pair of octants {0,1}; rotate by quadrants
    long long quad = (oct / 2) % 4;
    if (rem == 0 && oct % 2 == 0) { c = 1.0; s = 0.0; }  // exact axes
    double cc = c, ss = s;
    if (quad == 1) { cc = -s; ss = c; }
    if (quad == 2) { cc = -c; ss = -s; }
    if (quad == 3) { cc = s; ss = -c; }
    return cs{cc, ss};
}
__host__ __device__ constexpr bool is_prime(int n) {
    if (n < 2) return false;
    for (int d = 2; d * d <= n; ++d)
        if (n % d == 0) return false;
    return true;
}
__host__ __device__ constexpr int smallest_factor(int n) {
    for (int d = 2; d * d <= n; ++d)
        if (n % d == 0) return d;
    return n;
}
}  // namespace ct

// compile-time for-loop with the index available as a constant expression
template <int B, int E, typename F>
SSFFT_HD void sfor(F &&f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        sfor<B + 1, E>(f);
    }
}

// v * exp(-2*pi*i*NUM/DEN) with the constant folded at compile time; trivial cases cost nothing.
template <int NUM_, int DEN, typename T>
SSFFT_HD cx<T> mul_root(cx<T> v) {
    constexpr int NUM = ((NUM_ % DEN) + DEN) % DEN;
    if constexpr (NUM == 0) {
        return v;
    } else if constexpr (4 * NUM == DEN) {  // -i
        return mk<T>(v.y, -v.x);
    } else if constexpr (2 * NUM == DEN) {  // -1
        return mk<T>(-v.x, -v.y);
    } else if constexpr (4 * NUM == 3 * DEN) {  // +i
        return mk<T>(-v.y, v.x);
    } else if constexpr ((8 * NUM) % DEN == 0) {  // odd eighth turns: (+-1 +-i)/sqrt2
        constexpr int e = (8 * NUM) / DEN;        // 1,3,5,7
        constexpr T h = (T)0.70710678118654752440;
#ifdef SSFFT_F32X2
        if constexpr (sizeof(T) == 4) {  // (v -+ i v) * (+-h): one packed add (swap + half negate) and one packed multiply
            if constexpr (e == 1) return scale2(sub_i(v, v), h);        // (x + y, y - x) h
            else if constexpr (e == 3) return scale2(add_i(v, v), -h);  // (x - y, y + x)(-h) = (y - x, -(x + y)) h
            else if constexpr (e == 5) return scale2(sub_i(v, v), -h);  // (-(x + y), x - y) h
            else return scale2(add_i(v, v), h);                          // (x - y, x + y) h
        }
#endif
        // exp(-i*pi/4*e): e=1: (1 - i)h ; e=3: (-1 - i)h ; e=5: (-1 + i)h ; e=7: (1 + i)h
        if constexpr (e == 1) return mk<T>((v.x + v.y) * h, (v.y - v.x) * h);
        else if constexpr (e == 3) return mk<T>((v.y - v.x) * h, -(v.x + v.y) * h);
        else if constexpr (e == 5) return mk<T>(-(v.x + v.y) * h, (v.x - v.y) * h);
        else return mk<T>((v.x - v.y) * h, (v.x + v.y) * h);
    } else {
        constexpr ct::cs w = ct::cossin2pi(NUM, DEN);
        constexpr T c = (T)w.c, s = (T)w.s;  // multiply by (c - i s)
#ifdef SSFFT_F32X2
        if constexpr (sizeof(T) == 4) return mul_cs(v, c, s);
#endif
        return mk<T>(v.x * c + v.y * s, v.y * c - v.x * s);
    }
}

}  // namespace ssfft
