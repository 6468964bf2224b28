// tools/kbench.cu -- times fused-kernel template variants in ONE GPU call (development tool, not shipped).
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr \
//        -I fft_b200/csrc -o tools/kbench.bin tools/kbench.cu && gpurun -- ./tools/kbench.bin tuneN
// B(...) = plain loads, P(...) = TMA prefetch into a staging buffer (PF = 1), P2(...) = in-place staging (PF = 2);
// arguments: type, N, four radices, threads per transform, transforms per CTA, min CTAs/SM, padding shift, waves.
// g_mode selects C2C / R2C / C2R.  Every variant of a size must reproduce the first variant's output (maxdiff column).
// Sections tune2 ... tune11 are the rounds whose outputs are kept under profiles/kbench_*_r01.txt; the earliest
// sections (4096, c4, pow2, real, tune2-4) compile only with -DKBENCH_ALL.
#define SSFFT_NO_EX_KERNELS 1  // plain kernels only (compile time)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>

#include "fused_launch.cuh"

namespace ssfft {
const std::vector<FusedEntry> &fused_registry() { static std::vector<FusedEntry> v; return v; }
static int g_waves = 4;
int fused_waves() { return g_waves; }
}  // namespace ssfft
using namespace ssfft;

template <typename Cfg>
void *make_tw() {
    using T = typename Cfg::T;
    FusedEntry e = make_entry<Cfg>("x");
    std::vector<T> h(2 * (size_t)(e.tw_total > 0 ? e.tw_total : 1));
    size_t o = 0; int P = 1;
    for (int p = 0; p + 1 < e.np; ++p) {
        const int R = e.radix[p], MN = e.n / (P * R);
        for (int r = 1; r < R; ++r)
            for (int m = 0; m < MN; ++m) {
                unsigned long long q = (unsigned long long)P * m * r % (unsigned long long)e.n;
                long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)q / (long double)e.n;
                h[2 * o] = (T)cosl(a); h[2 * o + 1] = (T)(-sinl(a)); ++o;
            }
        P *= R;
    }
    void *d; cudaMalloc(&d, h.size() * sizeof(T));
    cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    return d;
}

__global__ void diff_kernel(const float *a, const float *b, size_t n, unsigned *maxbits) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    float m = 0.f;
    for (; i < n; i += st) m = fmaxf(m, fabsf(a[i] - b[i]));
    atomicMax(maxbits, __float_as_uint(m));
}
static void *g_ref = nullptr; static int g_ref_n = 0; static size_t g_ref_sz = 0;
static unsigned *g_maxbits = nullptr;

static int g_mode = 0;            // FUSED_C2C / FUSED_R2C / FUSED_C2R
static const void *g_rtw = nullptr;  // RealFFT twiddles for the current complex length

template <typename T>
void *make_rtw(int n_complex) {
    std::vector<T> h(2 * (size_t)(n_complex / 2 + 1));
    fill_real_twiddles<T>(h.data(), 2 * (size_t)n_complex, false);
    void *d; cudaMalloc(&d, h.size() * sizeof(T));
    cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    return d;
}

template <typename Cfg>
void bench(const char *name, const void *in, void *out, long long batch, int waves, int iters = 10) {
    using T = typename Cfg::T;
    g_waves = waves;
    void *tw = make_tw<Cfg>();
    void *rtw = g_mode ? make_rtw<T>(Cfg::N) : nullptr;
    g_rtw = rtw;
    const int inv_ = g_mode == 2 ? 1 : 0;
#define launch_cfg_m(tw_, in_, out_, batch_) launch_cfg<Cfg>(tw_, in_, out_, batch_, inv_, g_mode, g_rtw, 0)
    int rc = launch_cfg_m(tw, in, out, batch);
    if (rc || cudaDeviceSynchronize() != cudaSuccess) { printf("%-40s FAILED rc=%d %s\n", name, rc, cudaGetErrorString(cudaGetLastError())); return; }
    for (int i = 0; i < 2; ++i) launch_cfg_m(tw, in, out, batch);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) launch_cfg_m(tw, in, out, batch);
#undef launch_cfg_m
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fused_fft_kernel<Cfg>, Cfg::TX * Cfg::FPB, Cfg::smem_bytes);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, fused_fft_kernel<Cfg>);
    double bytes = 2.0 * batch * Cfg::N * sizeof(cx<T>);
    // correctness guard: every variant of a size must reproduce the first variant's output
    const size_t cmp_floats = (size_t)64 << 20;  // first 256 MiB (f32 view)
    float maxdiff = -1.f;
    if (sizeof(T) == 4) {
        if (!g_ref) { cudaMalloc(&g_ref, cmp_floats * 4); cudaMalloc(&g_maxbits, 4); }
        if (g_ref_n != Cfg::N + 100000 * g_mode || g_ref_sz != sizeof(T)) {
            cudaMemcpy(g_ref, out, cmp_floats * 4, cudaMemcpyDeviceToDevice); g_ref_n = Cfg::N + 100000 * g_mode; g_ref_sz = sizeof(T);
        } else {
            cudaMemset(g_maxbits, 0, 4);
            diff_kernel<<<1184, 256>>>((const float *)out, (const float *)g_ref, cmp_floats, g_maxbits);
            unsigned bits = 0; cudaMemcpy(&bits, g_maxbits, 4, cudaMemcpyDeviceToHost);
            memcpy(&maxdiff, &bits, 4);
        }
    }
    if (rtw) cudaFree(rtw);
    printf("%s %-46s waves=%d  %8.4f ms  %7.1f GB/s  frac %.3f  regs=%d occ=%d smem=%zu  maxdiff=%g\n", g_mode == 0 ? "C2C" : g_mode == 1 ? "R2C" : "C2R", name, waves, ms,
           bytes / ms / 1e6, bytes / ms / 1e6 / 6528.1, fa.numRegs, occ, Cfg::smem_bytes, maxdiff);
    cudaFree(tw);
}

#define B(T, N, R0, R1, R2, R3, TX, FPB, MINB, PADS, W) \
    bench<FusedCfg<T, N, R0, R1, R2, R3, TX, FPB, MINB, PADS>>(#T " " #N " " #R0 "x" #R1 "x" #R2 "x" #R3 " tx" #TX " fpb" #FPB " mb" #MINB " ps" #PADS, in, out, batch_for(N, sizeof(T)), W)
#define P(T, N, R0, R1, R2, R3, TX, FPB, MINB, PADS, W) \
    bench<FusedCfg<T, N, R0, R1, R2, R3, TX, FPB, MINB, PADS, 1>>(#T " " #N " " #R0 "x" #R1 "x" #R2 "x" #R3 " tx" #TX " fpb" #FPB " mb" #MINB " ps" #PADS " PF", in, out, batch_for(N, sizeof(T)), W)

#define P2(T, N, R0, R1, R2, R3, TX, FPB, MINB, PADS, W) \
    bench<FusedCfg<T, N, R0, R1, R2, R3, TX, FPB, MINB, PADS, 2>>(#T " " #N " " #R0 "x" #R1 "x" #R2 "x" #R3 " tx" #TX " fpb" #FPB " mb" #MINB " ps" #PADS " PF2", in, out, batch_for(N, sizeof(T)), W)

static long long batch_for(int n, size_t sz) { return (long long)((2ull << 30) / (2 * sz * n)); }  // 2 GiB of input

int main(int argc, char **argv) {
    setvbuf(stdout, nullptr, _IOLBF, 0);
    const char *which = argc > 1 ? argv[1] : "4096";
    void *in, *out;
    size_t bytes = (size_t)2200 << 20;
    cudaMalloc(&in, bytes); cudaMalloc(&out, bytes);
    cudaMemset(in, 0, bytes);
    // plain copy for reference (what the HBM peak figure is measured with)
    {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int i = 0; i < 2; ++i) cudaMemcpyAsync(out, in, (size_t)2 << 30, cudaMemcpyDeviceToDevice);
        cudaEventRecord(e0);
        for (int i = 0; i < 10; ++i) cudaMemcpyAsync(out, in, (size_t)2 << 30, cudaMemcpyDeviceToDevice);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
        printf("cudaMemcpy D2D 2 GiB: %.4f ms  %.1f GB/s (read+write)\n", ms, 2.0 * (2ull << 30) / ms / 1e6);
    }
    std::string w(which);
    // input must be non-trivial for the correctness guard
    { std::vector<float> h(1 << 20); for (size_t i = 0; i < h.size(); ++i) h[i] = (float)((i * 2654435761u) >> 8 & 0xffff) / 65536.f - 0.5f;
      for (size_t off = 0; off < bytes; off += h.size() * 4) cudaMemcpy((char *)in + off, h.data(), std::min(h.size() * 4, bytes - off), cudaMemcpyHostToDevice); }
#ifdef KBENCH_ALL  // earlier sweeps (results under profiles/kbench_*.txt); compile with -DKBENCH_ALL to rerun them
    if (w == "4096" || w == "all") {
        B(float, 4096, 16, 16, 16, 1, 256, 1, 2, 4, 4);
        B(float, 4096, 16, 16, 16, 1, 256, 1, 3, 4, 4);
        B(float, 4096, 16, 16, 16, 1, 256, 1, 3, 4, 8);
        B(float, 4096, 16, 16, 16, 1, 256, 1, 3, 4, 16);
        B(float, 4096, 16, 16, 16, 1, 256, 1, 4, 4, 4);
        P(float, 4096, 16, 16, 16, 1, 256, 1, 2, 4, 4);
        P(float, 4096, 16, 16, 16, 1, 256, 1, 3, 4, 4);
        P(float, 4096, 16, 16, 16, 1, 256, 1, 3, 4, 1);
        P(float, 4096, 16, 16, 16, 1, 256, 1, 3, 4, 16);
        P(float, 4096, 16, 16, 16, 1, 256, 2, 1, 4, 4);
        P(float, 4096, 8, 8, 8, 8, 512, 1, 3, 4, 4);
        P(float, 4096, 32, 32, 4, 1, 128, 1, 2, 4, 4);
        P(float, 4096, 64, 64, 1, 1, 64, 2, 1, 4, 4);
    }
    if (w == "c4" || w == "all") {
        B(float, 1000, 10, 10, 10, 1, 100, 4, 2, 4, 4);
        B(float, 1000, 10, 10, 10, 1, 100, 4, 3, 4, 4);
        B(float, 1000, 10, 10, 10, 1, 100, 2, 5, 4, 4);
        B(float, 1000, 10, 10, 10, 1, 100, 2, 6, 4, 4);
        P(float, 1000, 10, 10, 10, 1, 100, 2, 4, 4, 4);
        P(float, 1000, 10, 10, 10, 1, 100, 4, 2, 4, 4);
        P(float, 1000, 10, 10, 10, 1, 100, 4, 3, 4, 4);
        P(float, 1000, 20, 10, 5, 1, 50, 4, 4, 4, 4);
        P(float, 1000, 20, 10, 5, 1, 50, 8, 2, 4, 4);
        B(float, 2187, 27, 9, 9, 1, 81, 2, 4, 4, 4);
        P(float, 2187, 27, 9, 9, 1, 81, 2, 4, 4, 4);
        P(float, 2187, 27, 9, 9, 1, 81, 4, 2, 4, 4);
        P(float, 2187, 9, 9, 9, 3, 243, 2, 2, 4, 4);
        B(float, 3125, 25, 25, 5, 1, 125, 1, 4, 4, 4);
        B(float, 3125, 25, 25, 5, 1, 125, 1, 5, 4, 4);
        P(float, 3125, 25, 25, 5, 1, 125, 2, 2, 4, 4);
        P(float, 3125, 25, 25, 5, 1, 125, 2, 3, 4, 4);
        B(float, 6000, 10, 10, 10, 6, 200, 1, 2, 4, 4);
        B(float, 6000, 10, 10, 10, 6, 200, 1, 3, 4, 4);
        P(float, 6000, 10, 10, 10, 6, 200, 1, 2, 4, 4);
        P(float, 6000, 10, 10, 10, 6, 200, 1, 3, 4, 4);
        P(float, 6000, 10, 10, 10, 6, 200, 2, 1, 4, 4);
    }
    if (w == "real") {
        // real transforms of length 2N through the fused kernel of complex length N: R2C epilogue / C2R on-the-fly gather
        for (int mode = 1; mode <= 2; ++mode) {
            g_mode = mode;
            B(float, 128, 16, 8, 1, 1, 8, 32, 2, 4, 4);   // registered
            B(float, 128, 16, 8, 1, 1, 8, 32, 3, 4, 4);
            B(float, 128, 16, 8, 1, 1, 8, 16, 4, 4, 4);
            B(float, 128, 16, 8, 1, 1, 8, 16, 6, 4, 4);
            B(float, 128, 8, 16, 1, 1, 8, 32, 3, 4, 4);
            B(float, 128, 8, 4, 4, 1, 16, 16, 4, 4, 4);
            B(float, 128, 8, 4, 4, 1, 16, 16, 6, 4, 4);
            B(float, 512, 32, 16, 1, 1, 16, 8, 4, 5, 4);  // registered
            P(float, 512, 32, 16, 1, 1, 16, 8, 3, 5, 4);
            B(float, 512, 16, 16, 2, 1, 32, 4, 3, 4, 4);
            B(float, 512, 16, 16, 2, 1, 32, 4, 4, 4, 4);
            B(float, 512, 8, 8, 8, 1, 64, 4, 4, 4, 4);
            B(float, 512, 8, 8, 8, 1, 64, 4, 6, 4, 4);
            P(float, 512, 8, 8, 8, 1, 64, 4, 4, 4, 4);
            B(float, 2048, 16, 16, 8, 1, 128, 1, 6, 4, 4);  // registered
            P(float, 2048, 16, 16, 8, 1, 128, 2, 3, 4, 4);
            P(float, 4096, 16, 16, 16, 1, 256, 1, 2, 4, 4);  // registered
            P(float, 4096, 16, 16, 16, 1, 256, 1, 3, 4, 4);
            B(float, 4096, 16, 16, 16, 1, 256, 1, 3, 4, 4);
            P(float, 8192, 32, 16, 16, 1, 256, 1, 1, 5, 4);  // registered
            B(float, 8192, 32, 16, 16, 1, 256, 1, 1, 5, 4);
            B(float, 16384, 32, 32, 16, 1, 512, 1, 1, 5, 4);  // registered
        }
        g_mode = 0;
    }
    if (w == "tune2") {
        B(float, 16384, 32, 32, 16, 1, 512, 1, 1, 5, 4);   // registered
        B(float, 16384, 16, 32, 32, 1, 512, 1, 1, 5, 4);
        B(float, 16384, 32, 16, 32, 1, 512, 1, 1, 5, 4);
        B(float, 16384, 32, 32, 16, 1, 512, 1, 1, 4, 4);
        B(float, 16384, 16, 16, 8, 8, 1024, 1, 1, 4, 4);
        B(float, 16384, 16, 16, 16, 4, 1024, 1, 1, 4, 4);
        P(float, 8192, 32, 16, 16, 1, 256, 1, 1, 5, 4);    // registered
        P(float, 8192, 16, 16, 32, 1, 256, 1, 1, 5, 4);
        P(float, 8192, 16, 32, 16, 1, 256, 1, 1, 5, 4);
        B(float, 8192, 16, 16, 8, 4, 512, 1, 2, 4, 4);
        P(float, 8192, 16, 16, 8, 4, 512, 1, 1, 4, 4);
        P(float, 6000, 10, 10, 10, 6, 200, 1, 2, 31, 4);   // registered
        P(float, 6000, 24, 25, 10, 1, 250, 1, 2, 31, 4);
        P(float, 6000, 25, 24, 10, 1, 250, 1, 2, 31, 4);
        P(float, 6000, 20, 20, 15, 1, 300, 1, 2, 31, 4);
        P(float, 6000, 15, 20, 20, 1, 300, 1, 2, 31, 4);
        P(float, 6000, 16, 15, 25, 1, 250, 1, 2, 31, 4);
        B(float, 6000, 20, 20, 15, 1, 300, 1, 3, 31, 4);
        B(float, 2187, 27, 9, 9, 1, 81, 3, 2, 31, 4);      // registered
        B(float, 2187, 27, 27, 3, 1, 81, 3, 2, 31, 4);
        B(float, 2187, 9, 9, 27, 1, 81, 3, 2, 31, 4);
        B(float, 2187, 9, 27, 9, 1, 81, 3, 2, 31, 4);
        B(float, 2187, 27, 9, 9, 1, 81, 3, 3, 31, 4);
        B(float, 3125, 25, 25, 5, 1, 125, 1, 5, 31, 4);    // registered
        B(float, 3125, 5, 25, 25, 1, 125, 1, 5, 31, 4);
        B(float, 3125, 25, 5, 25, 1, 125, 1, 5, 31, 4);
        B(float, 3125, 25, 25, 5, 1, 125, 2, 3, 31, 4);
        P(float, 1000, 10, 10, 10, 1, 100, 2, 4, 31, 4);   // registered
        P(float, 1000, 25, 8, 5, 1, 125, 2, 4, 31, 4);
        P(float, 1000, 8, 25, 5, 1, 125, 2, 4, 31, 4);
        P(float, 1000, 10, 10, 10, 1, 100, 4, 3, 31, 4);
    }
    if (w == "tune3") {
        P(float, 6000, 25, 24, 10, 1, 250, 1, 2, 31, 4);
        P(float, 6000, 25, 24, 10, 1, 250, 1, 2, 4, 4);
        P(float, 6000, 25, 24, 10, 1, 250, 1, 2, 5, 4);
        B(float, 6000, 25, 24, 10, 1, 250, 1, 3, 31, 4);
        B(float, 6000, 25, 24, 10, 1, 250, 1, 4, 31, 4);
        P(float, 6000, 25, 12, 20, 1, 300, 1, 2, 31, 4);
        P(float, 6000, 25, 20, 12, 1, 300, 1, 2, 31, 4);
        P(float, 6000, 25, 15, 16, 1, 400, 1, 2, 31, 4);
        B(float, 2187, 9, 9, 27, 1, 81, 3, 2, 31, 4);
        B(float, 2187, 9, 9, 27, 1, 81, 3, 3, 31, 4);
        P(float, 2187, 9, 9, 27, 1, 81, 2, 3, 31, 4);
        B(float, 2187, 9, 9, 27, 1, 81, 2, 4, 31, 4);
        B(double, 6000, 10, 10, 10, 6, 200, 1, 1, 31, 4);   // registered
        B(double, 6000, 25, 24, 10, 1, 250, 1, 1, 31, 4);
        B(double, 6000, 20, 20, 15, 1, 300, 1, 1, 31, 4);
        B(double, 6000, 15, 20, 20, 1, 300, 1, 1, 31, 4);
        P(double, 6000, 10, 10, 10, 6, 200, 1, 1, 31, 4);
        B(double, 6000, 10, 10, 10, 6, 200, 1, 2, 31, 4);
        B(double, 3125, 25, 25, 5, 1, 125, 2, 1, 31, 4);    // registered
        B(double, 3125, 25, 25, 5, 1, 125, 1, 2, 31, 4);
        B(double, 3125, 25, 25, 5, 1, 125, 1, 3, 31, 4);
        B(double, 3125, 5, 25, 25, 1, 125, 1, 3, 31, 4);
        B(double, 2187, 9, 9, 9, 3, 243, 1, 2, 31, 4);      // registered
        B(double, 2187, 9, 9, 27, 1, 81, 3, 1, 31, 4);
        B(double, 2187, 9, 9, 27, 1, 81, 2, 2, 31, 4);
        B(double, 2187, 27, 9, 9, 1, 81, 2, 2, 31, 4);
        B(double, 1000, 10, 10, 10, 1, 100, 2, 2, 31, 4);   // registered
        B(double, 1000, 10, 10, 10, 1, 100, 2, 3, 31, 4);
        B(double, 1000, 10, 10, 10, 1, 100, 2, 4, 31, 4);
        P(double, 1000, 10, 10, 10, 1, 100, 2, 2, 31, 4);
        P(double, 1000, 10, 10, 10, 1, 100, 2, 3, 31, 4);
    }
    if (w == "tune4") {
        P(double, 6000, 25, 24, 10, 1, 250, 1, 1, 31, 4);
        P(double, 6000, 25, 20, 12, 1, 300, 1, 1, 31, 4);
        P(double, 3125, 25, 25, 5, 1, 125, 1, 2, 31, 4);   // odd N: bulk copies of 50000 B (multiple of 16)
        P(double, 2187, 9, 9, 9, 3, 243, 1, 2, 31, 4);
        P(double, 2187, 9, 9, 9, 3, 243, 2, 1, 31, 4);
        B(double, 1024, 8, 8, 4, 4, 128, 2, 3, 3, 4);      // registered is the PF form of this
        P(double, 1024, 8, 8, 4, 4, 128, 2, 3, 3, 4);
        P(double, 1024, 16, 8, 8, 1, 64, 2, 3, 3, 4);
        P(double, 1024, 8, 8, 16, 1, 64, 2, 3, 3, 4);
        P(double, 1024, 16, 16, 4, 1, 64, 2, 3, 3, 4);
        P(double, 1024, 8, 8, 4, 4, 128, 4, 1, 3, 4);
        P(double, 4096, 8, 8, 8, 8, 512, 1, 1, 3, 4);      // registered
        P(double, 4096, 16, 16, 16, 1, 256, 1, 1, 3, 4);
        P(double, 4096, 16, 16, 16, 1, 256, 1, 1, 4, 4);
        P(double, 4096, 16, 16, 4, 4, 512, 1, 1, 3, 4);
        B(double, 8192, 8, 8, 8, 16, 512, 1, 1, 3, 4);
        B(double, 8192, 16, 8, 8, 8, 512, 1, 1, 3, 4);
        B(double, 8192, 8, 8, 8, 16, 1024, 1, 1, 3, 4);
        P(float, 1536, 16, 16, 6, 1, 96, 2, 3, 4, 4);
        B(float, 1536, 16, 16, 6, 1, 96, 2, 3, 4, 4);      // registered
        P(float, 2304, 16, 16, 9, 1, 144, 2, 2, 4, 4);
        B(float, 2304, 16, 16, 9, 1, 144, 1, 3, 4, 4);     // registered
        P(float, 3072, 16, 16, 12, 1, 192, 1, 3, 4, 4);
        B(float, 3072, 16, 16, 12, 1, 192, 1, 3, 4, 4);    // registered
        P(float, 6144, 16, 16, 24, 1, 384, 1, 1, 4, 4);
        B(float, 6144, 16, 16, 24, 1, 384, 1, 1, 4, 4);    // registered
        P(float, 4608, 16, 16, 18, 1, 288, 1, 2, 4, 4);
        B(float, 4608, 16, 16, 18, 1, 288, 1, 2, 4, 4);    // registered
    }
    if (w == "pow2" || w == "all") {
        B(float, 1024, 32, 32, 1, 1, 32, 4, 2, 4, 4);
        B(float, 1024, 32, 32, 1, 1, 32, 4, 3, 4, 4);
        P(float, 1024, 32, 32, 1, 1, 32, 4, 2, 4, 4);
        B(float, 1024, 16, 16, 4, 1, 64, 4, 3, 4, 4);
        B(float, 1024, 16, 16, 4, 1, 64, 2, 6, 4, 4);
        B(float, 2048, 16, 16, 8, 1, 128, 2, 3, 4, 4);
        B(float, 2048, 16, 16, 8, 1, 128, 1, 6, 4, 4);
        P(float, 2048, 16, 16, 8, 1, 128, 2, 3, 4, 4);
        B(float, 512, 8, 8, 8, 1, 64, 4, 4, 4, 4);
        B(float, 512, 8, 8, 8, 1, 64, 4, 6, 4, 4);
        B(float, 512, 32, 16, 1, 1, 16, 8, 4, 4, 4);
        B(float, 8192, 32, 16, 16, 1, 256, 1, 1, 4, 4);
        P(float, 8192, 32, 16, 16, 1, 256, 1, 1, 4, 4);
        P(float, 8192, 16, 8, 8, 8, 512, 1, 1, 4, 4);
        B(float, 8192, 16, 8, 8, 8, 512, 1, 2, 4, 4);
        P(float, 8192, 16, 8, 8, 8, 512, 1, 2, 4, 4);
        B(float, 16384, 32, 32, 16, 1, 512, 1, 1, 4, 4);
        P(float, 16384, 16, 16, 8, 8, 1024, 1, 1, 4, 4);
        B(double, 1024, 8, 8, 4, 4, 128, 2, 2, 4, 4);
        B(double, 1024, 8, 8, 4, 4, 128, 2, 3, 4, 4);
        P(double, 1024, 8, 8, 4, 4, 128, 2, 2, 4, 4);
        P(double, 1024, 8, 8, 4, 4, 128, 2, 3, 4, 4);
        B(double, 4096, 8, 8, 8, 8, 512, 1, 1, 4, 4);
        P(double, 4096, 8, 8, 8, 8, 512, 1, 1, 4, 4);
        P(double, 2048, 8, 8, 8, 4, 256, 1, 2, 4, 4);
        B(double, 2048, 8, 8, 8, 4, 256, 1, 2, 4, 4);
    }
#endif  // KBENCH_ALL
    if (w == "tune11") {
        P2(float, 16384, 32, 32, 16, 1, 512, 1, 1, 5, 4);   // registered
        P2(float, 16384, 32, 32, 16, 1, 512, 1, 1, 4, 4);
        P2(float, 16384, 32, 16, 32, 1, 512, 1, 1, 5, 4);
        P2(float, 16384, 16, 32, 32, 1, 512, 1, 1, 5, 4);
        P2(float, 16384, 32, 32, 16, 1, 512, 1, 1, 5, 16);
        P2(float, 16384, 32, 32, 16, 1, 512, 1, 1, 5, 1);
        P(float, 4096, 16, 16, 16, 1, 256, 1, 2, 4, 2);
        P(float, 4096, 16, 16, 16, 1, 256, 1, 2, 4, 8);
        P(float, 4096, 16, 16, 16, 1, 256, 1, 2, 4, 32);
        P2(float, 8192, 32, 16, 16, 1, 256, 1, 2, 5, 16);
        P2(float, 8192, 32, 16, 16, 1, 256, 1, 2, 5, 1);
        for (int mode = 1; mode <= 2; ++mode) {
            g_mode = mode;
            P(float, 128, 16, 8, 1, 1, 8, 32, 2, 4, 4);     // registered
            P(float, 128, 16, 8, 1, 1, 8, 16, 4, 4, 4);
            P(float, 128, 16, 8, 1, 1, 8, 16, 3, 4, 4);
            P(float, 128, 8, 16, 1, 1, 8, 32, 2, 4, 4);
            P(float, 128, 8, 16, 1, 1, 16, 16, 3, 4, 4);
            P2(float, 128, 16, 8, 1, 1, 8, 16, 6, 4, 4);
            P(float, 64, 8, 8, 1, 1, 8, 32, 2, 4, 4);       // registered
            P(float, 64, 8, 8, 1, 1, 8, 32, 3, 4, 4);
            P(float, 64, 8, 8, 1, 1, 8, 16, 4, 4, 4);
            P2(float, 16384, 32, 32, 16, 1, 512, 1, 1, 5, 4);
            P2(float, 16384, 32, 32, 16, 1, 512, 1, 1, 4, 4);
        }
        g_mode = 0;
    }
    if (w == "tune10") {  // real flavours of the PF = 2 entries, and a few more PF = 2 candidates
        for (int mode = 1; mode <= 2; ++mode) {
            g_mode = mode;
            P(float, 8192, 32, 16, 16, 1, 256, 1, 1, 5, 4);
            P2(float, 8192, 32, 16, 16, 1, 256, 1, 2, 5, 4);
            P(float, 2048, 16, 16, 8, 1, 128, 2, 3, 4, 4);
            P2(float, 2048, 16, 16, 8, 1, 128, 1, 6, 4, 4);
            B(float, 2048, 16, 16, 8, 1, 128, 1, 6, 4, 4);
            P2(float, 4096, 16, 16, 16, 1, 256, 1, 3, 4, 4);
            P2(float, 4096, 16, 16, 16, 1, 256, 1, 4, 4, 4);
            P2(float, 1024, 32, 32, 1, 1, 32, 4, 3, 5, 4);
            P(float, 1024, 32, 32, 1, 1, 32, 4, 3, 5, 4);
            P2(float, 512, 32, 16, 1, 1, 16, 8, 4, 5, 4);
            P(float, 512, 32, 16, 1, 1, 16, 8, 3, 5, 4);
            P2(float, 256, 16, 16, 1, 1, 16, 8, 6, 4, 4);
            P(float, 256, 16, 16, 1, 1, 16, 8, 4, 4, 4);
            P2(float, 128, 16, 8, 1, 1, 8, 32, 3, 4, 4);
            P(float, 128, 16, 8, 1, 1, 8, 32, 2, 4, 4);
            P2(double, 4096, 16, 16, 16, 1, 256, 1, 2, 4, 4);
            P(double, 4096, 16, 16, 16, 1, 256, 1, 1, 4, 4);
            P2(double, 8192, 16, 16, 32, 1, 256, 1, 1, 3, 4);
        }
        g_mode = 0;
        P2(float, 1024, 32, 32, 1, 1, 32, 4, 3, 5, 4);
        P2(float, 1024, 32, 32, 1, 1, 32, 2, 6, 5, 4);
        P2(float, 512, 32, 16, 1, 1, 16, 8, 4, 5, 4);
        P2(float, 512, 32, 16, 1, 1, 16, 4, 6, 5, 4);
        P2(float, 1536, 16, 16, 6, 1, 96, 2, 3, 4, 4);
        P2(float, 1536, 16, 16, 6, 1, 96, 2, 5, 4, 4);
        P2(float, 3072, 16, 16, 12, 1, 192, 1, 3, 4, 4);
        P2(float, 3072, 16, 16, 12, 1, 192, 1, 5, 4, 4);
        P2(float, 2304, 16, 16, 9, 1, 144, 2, 3, 4, 4);
        P2(float, 2304, 16, 16, 9, 1, 144, 1, 5, 4, 4);
        P2(float, 1000, 10, 10, 10, 1, 100, 2, 5, 31, 4);
        P2(float, 1000, 10, 10, 10, 1, 100, 2, 8, 31, 4);
        P2(double, 1024, 8, 8, 16, 1, 64, 2, 4, 3, 4);
        P2(double, 1024, 8, 8, 16, 1, 64, 2, 3, 3, 4);
        P2(double, 1536, 8, 8, 8, 3, 192, 1, 3, 3, 4);
        P2(double, 2304, 8, 8, 4, 9, 288, 1, 3, 3, 4);
        P2(double, 1000, 10, 10, 10, 1, 100, 2, 4, 31, 4);
        P2(double, 6000, 25, 24, 10, 1, 250, 1, 1, 31, 4);
    }
    if (w == "tune9") {  // PF = 2: the prefetch lands in the exchange buffer itself (half the shared memory)
        P(float, 8192, 32, 16, 16, 1, 256, 1, 1, 5, 4);     // registered
        P2(float, 8192, 32, 16, 16, 1, 256, 1, 1, 5, 4);
        P2(float, 8192, 32, 16, 16, 1, 256, 1, 2, 5, 4);
        P2(float, 8192, 16, 16, 32, 1, 512, 1, 1, 5, 4);
        B(float, 16384, 32, 32, 16, 1, 512, 1, 1, 5, 4);    // registered
        P2(float, 16384, 32, 32, 16, 1, 512, 1, 1, 5, 4);
        P2(float, 9216, 32, 16, 18, 1, 288, 1, 1, 5, 4);
        P2(float, 9216, 32, 16, 18, 1, 288, 1, 2, 5, 4);
        P(float, 4096, 16, 16, 16, 1, 256, 1, 2, 4, 4);     // registered
        P2(float, 4096, 16, 16, 16, 1, 256, 1, 2, 4, 4);
        P2(float, 4096, 16, 16, 16, 1, 256, 1, 3, 4, 4);
        P2(float, 4096, 16, 16, 16, 1, 256, 1, 4, 4, 4);
        P2(float, 6144, 16, 16, 24, 1, 384, 1, 1, 4, 4);
        P2(float, 6144, 16, 16, 24, 1, 384, 1, 2, 4, 4);
        P2(float, 6000, 25, 24, 10, 1, 250, 1, 2, 31, 4);
        P2(float, 6000, 25, 24, 10, 1, 250, 1, 3, 31, 4);
        P2(float, 4608, 16, 16, 18, 1, 288, 1, 2, 4, 4);
        P2(float, 4608, 16, 16, 18, 1, 288, 1, 3, 4, 4);
        P2(float, 2048, 16, 16, 8, 1, 128, 2, 3, 4, 4);
        P2(float, 2048, 16, 16, 8, 1, 128, 1, 6, 4, 4);
        P2(float, 2187, 9, 9, 27, 1, 81, 2, 3, 31, 4);
        P2(float, 2187, 9, 9, 27, 1, 81, 2, 4, 31, 4);
        P2(float, 3125, 25, 25, 5, 1, 125, 2, 3, 31, 4);
        P2(double, 4096, 16, 16, 16, 1, 256, 1, 1, 4, 4);
        P2(double, 4096, 16, 16, 16, 1, 256, 1, 2, 4, 4);
        P2(double, 8192, 16, 16, 32, 1, 256, 1, 1, 3, 4);
        P2(double, 2048, 8, 16, 16, 1, 128, 1, 2, 3, 4);
        P2(double, 2048, 8, 16, 16, 1, 128, 1, 3, 3, 4);
        P2(double, 6000, 25, 24, 10, 1, 250, 1, 1, 31, 4);
        P2(double, 6000, 25, 24, 10, 1, 250, 1, 2, 31, 4);
        P2(double, 3072, 16, 16, 12, 1, 192, 1, 1, 3, 4);
        P2(double, 3072, 16, 16, 12, 1, 192, 1, 2, 3, 4);
        P2(double, 3125, 25, 25, 5, 1, 125, 1, 3, 31, 4);
        P2(double, 2187, 9, 9, 9, 3, 243, 1, 3, 31, 4);
        for (int mode = 1; mode <= 2; ++mode) {
            g_mode = mode;
            P(float, 8192, 32, 16, 16, 1, 256, 1, 1, 5, 4);
            P2(float, 8192, 32, 16, 16, 1, 256, 1, 1, 5, 4);
            B(float, 16384, 32, 32, 16, 1, 512, 1, 1, 5, 4);
            P2(float, 16384, 32, 32, 16, 1, 512, 1, 1, 5, 4);
            P2(float, 4096, 16, 16, 16, 1, 256, 1, 3, 4, 4);
        }
        g_mode = 0;
    }
    if (w == "tune8") {  // complex cores of RealFFT 1000 / 6000 (C2C + both real flavours)
        for (int mode = 0; mode <= 2; ++mode) {
            g_mode = mode;
            P(float, 500, 10, 10, 5, 1, 50, 4, 4, 31, 4);
            P(float, 500, 10, 10, 5, 1, 50, 4, 3, 31, 4);
            P(float, 500, 10, 10, 5, 1, 50, 8, 2, 31, 4);
            P(float, 500, 25, 20, 1, 1, 25, 8, 2, 31, 4);
            P(float, 500, 20, 25, 1, 1, 25, 8, 2, 31, 4);
            P(float, 3000, 25, 12, 10, 1, 125, 2, 2, 31, 4);
            P(float, 3000, 25, 12, 10, 1, 125, 1, 4, 31, 4);
            P(float, 3000, 10, 10, 30, 1, 100, 2, 2, 31, 4);
            P(float, 3000, 15, 20, 10, 1, 150, 1, 3, 31, 4);
            P(float, 3000, 25, 24, 5, 1, 125, 2, 2, 31, 4);
            P(double, 500, 10, 10, 5, 1, 50, 4, 3, 31, 4);
            P(double, 500, 10, 10, 5, 1, 50, 4, 2, 31, 4);
            P(double, 500, 25, 20, 1, 1, 25, 4, 2, 31, 4);
            P(double, 3000, 25, 12, 10, 1, 125, 1, 2, 31, 4);
            P(double, 3000, 10, 10, 30, 1, 100, 1, 2, 31, 4);
            P(double, 3000, 25, 24, 5, 1, 125, 1, 2, 31, 4);
        }
        g_mode = 0;
    }
    if (w == "tune7") {
        P(float, 8192, 32, 16, 16, 1, 256, 1, 1, 5, 4);    // registered
        P(float, 8192, 16, 16, 32, 1, 512, 1, 1, 5, 4);    // radix-32 pass ragged over 512 threads
        P(float, 8192, 32, 16, 16, 1, 512, 1, 1, 5, 4);
        P(float, 8192, 16, 32, 16, 1, 512, 1, 1, 5, 4);
        P(float, 8192, 32, 16, 16, 1, 256, 1, 1, 4, 4);
        P(float, 8192, 32, 16, 16, 1, 256, 1, 1, 31, 4);
        P(float, 4096, 16, 16, 16, 1, 256, 1, 2, 4, 4);    // registered
        P(float, 4096, 16, 16, 16, 1, 256, 1, 2, 5, 4);
        P(float, 4096, 16, 16, 16, 1, 256, 1, 2, 31, 4);
        P(float, 1024, 32, 32, 1, 1, 32, 4, 2, 5, 4);      // registered
        P(float, 1024, 32, 32, 1, 1, 32, 4, 3, 5, 4);
        P(float, 1024, 16, 16, 4, 1, 64, 4, 3, 4, 4);
        P(float, 1000, 10, 10, 10, 1, 100, 2, 4, 31, 4);   // registered
        P(float, 1000, 10, 10, 10, 1, 100, 2, 5, 31, 4);
        P(float, 1000, 10, 10, 10, 1, 100, 2, 3, 31, 4);
        P(float, 1000, 25, 40, 1, 1, 40, 4, 2, 31, 4);
        P(float, 2187, 9, 9, 27, 1, 81, 2, 3, 31, 4);      // registered
        P(float, 2187, 27, 81, 1, 1, 81, 2, 2, 31, 4);
        P(float, 2187, 9, 9, 27, 1, 81, 2, 2, 31, 4);
        P(double, 8192, 16, 8, 8, 8, 512, 1, 1, 3, 4);
        B(double, 8192, 16, 16, 32, 1, 512, 1, 1, 3, 4);
        B(double, 8192, 16, 16, 32, 1, 256, 1, 1, 3, 4);
        P(double, 4096, 16, 16, 16, 1, 256, 1, 1, 4, 4);   // registered
        P(double, 4096, 16, 16, 16, 1, 512, 1, 1, 4, 4);
        P(double, 2187, 9, 9, 27, 1, 81, 2, 2, 31, 4);
        P(double, 6000, 25, 24, 10, 1, 250, 1, 1, 31, 4);  // registered
        P(double, 6000, 25, 24, 10, 1, 500, 1, 1, 31, 4);
    }
    if (w == "tune6") {
        P(float, 96, 16, 6, 1, 1, 6, 32, 2, 4, 4);      B(float, 96, 16, 6, 1, 1, 6, 32, 2, 4, 4);
        P(float, 192, 16, 12, 1, 1, 12, 16, 2, 4, 4);   B(float, 192, 16, 12, 1, 1, 12, 16, 2, 4, 4);
        P(float, 144, 16, 9, 1, 1, 9, 16, 2, 4, 4);     B(float, 144, 16, 9, 1, 1, 9, 16, 2, 4, 4);
        P(float, 288, 16, 18, 1, 1, 18, 8, 2, 4, 4);    B(float, 288, 16, 18, 1, 1, 18, 8, 2, 4, 4);
        P(float, 512, 32, 16, 1, 1, 16, 8, 4, 5, 4);    B(float, 512, 32, 16, 1, 1, 16, 8, 4, 5, 4);
        P(float, 512, 32, 16, 1, 1, 16, 8, 3, 5, 4);
        P(float, 2048, 16, 16, 8, 1, 128, 2, 3, 4, 4);  B(float, 2048, 16, 16, 8, 1, 128, 1, 6, 4, 4);
        P(float, 2048, 8, 16, 16, 1, 128, 2, 3, 4, 4);
        P(float, 16384, 32, 32, 16, 1, 512, 1, 1, 31, 4);   // staging does not fit next to 128 KiB: expected FAILED
        P(float, 3125, 25, 25, 5, 1, 125, 2, 3, 31, 4); B(float, 3125, 25, 25, 5, 1, 125, 1, 5, 31, 4);
        P(float, 3125, 25, 25, 5, 1, 125, 2, 2, 31, 4);
        P(double, 64, 8, 8, 1, 1, 8, 32, 2, 3, 4);      B(double, 64, 8, 8, 1, 1, 8, 32, 2, 3, 4);
        P(double, 128, 8, 4, 4, 1, 16, 16, 2, 3, 4);    B(double, 128, 8, 4, 4, 1, 16, 16, 2, 3, 4);
        P(double, 128, 16, 8, 1, 1, 8, 16, 3, 3, 4);
        P(double, 96, 8, 12, 1, 1, 12, 16, 2, 3, 4);    B(double, 96, 8, 12, 1, 1, 12, 16, 2, 3, 4);
        P(double, 192, 8, 8, 3, 1, 24, 8, 2, 3, 4);     B(double, 192, 8, 8, 3, 1, 24, 8, 2, 3, 4);
        P(double, 384, 8, 8, 6, 1, 48, 4, 2, 3, 4);     B(double, 384, 8, 8, 6, 1, 48, 4, 2, 3, 4);
        P(double, 144, 8, 18, 1, 1, 18, 8, 2, 3, 4);    B(double, 144, 8, 18, 1, 1, 18, 8, 2, 3, 4);
        P(double, 288, 8, 4, 9, 1, 36, 4, 2, 3, 4);     B(double, 288, 8, 4, 9, 1, 36, 4, 2, 3, 4);
        P(double, 576, 8, 8, 9, 1, 72, 2, 2, 3, 4);     B(double, 576, 8, 8, 9, 1, 72, 2, 2, 3, 4);
        P(double, 4096, 16, 16, 16, 1, 256, 1, 1, 4, 4);
        for (int mode = 1; mode <= 2; ++mode) {  // real flavours of the new complex entries
            g_mode = mode;
            P(float, 128, 16, 8, 1, 1, 8, 32, 2, 4, 4);   B(float, 128, 16, 8, 1, 1, 8, 16, 4, 4, 4);
            P(float, 256, 16, 16, 1, 1, 16, 8, 4, 4, 4);  B(float, 256, 16, 16, 1, 1, 16, 8, 4, 4, 4);
            P(float, 64, 8, 8, 1, 1, 8, 32, 2, 4, 4);     B(float, 64, 8, 8, 1, 1, 8, 32, 2, 4, 4);
            P(float, 1024, 32, 32, 1, 1, 32, 4, 2, 5, 4); P(float, 1024, 32, 32, 1, 1, 32, 4, 3, 5, 4);
        }
        g_mode = 0;
    }
    if (w == "tune5") {
        P(float, 768, 16, 16, 3, 1, 48, 4, 3, 4, 4);
        B(float, 768, 16, 16, 3, 1, 48, 4, 3, 4, 4);      // registered
        P(float, 384, 16, 24, 1, 1, 24, 8, 2, 4, 4);
        B(float, 384, 16, 24, 1, 1, 24, 8, 2, 4, 4);      // registered
        P(float, 576, 16, 4, 9, 1, 36, 4, 3, 4, 4);
        B(float, 576, 16, 4, 9, 1, 36, 4, 3, 4, 4);       // registered
        P(float, 1152, 16, 8, 9, 1, 72, 2, 3, 4, 4);
        B(float, 1152, 16, 8, 9, 1, 72, 2, 3, 4, 4);      // registered
        P(float, 9216, 32, 16, 18, 1, 288, 1, 1, 5, 4);
        B(float, 9216, 32, 16, 18, 1, 288, 1, 1, 5, 4);   // registered
        P(float, 256, 16, 16, 1, 1, 16, 8, 4, 4, 4);
        P(float, 128, 16, 8, 1, 1, 8, 32, 2, 4, 4);
        B(float, 128, 16, 8, 1, 1, 8, 32, 2, 4, 4);       // registered
        P(float, 64, 8, 8, 1, 1, 8, 32, 2, 4, 4);
        B(float, 64, 8, 8, 1, 1, 8, 32, 2, 4, 4);         // registered
        P(double, 2048, 8, 8, 8, 4, 256, 1, 2, 3, 4);     // registered
        P(double, 2048, 16, 16, 8, 1, 128, 1, 2, 3, 4);
        P(double, 2048, 8, 16, 16, 1, 128, 1, 2, 3, 4);
        P(double, 2048, 16, 16, 8, 1, 128, 1, 2, 4, 4);
        B(double, 512, 8, 8, 8, 1, 64, 4, 3, 3, 4);       // registered
        P(double, 512, 8, 8, 8, 1, 64, 4, 3, 3, 4);
        P(double, 512, 8, 8, 8, 1, 64, 2, 4, 3, 4);
        P(double, 512, 16, 8, 4, 1, 32, 4, 3, 3, 4);
        P(double, 512, 8, 4, 16, 1, 32, 4, 3, 3, 4);
        B(double, 256, 8, 8, 4, 1, 32, 8, 2, 3, 4);       // registered
        P(double, 256, 8, 8, 4, 1, 32, 8, 2, 3, 4);
        P(double, 256, 16, 16, 1, 1, 16, 8, 3, 3, 4);
        P(double, 256, 8, 8, 4, 1, 32, 4, 4, 3, 4);
        B(double, 1536, 8, 8, 8, 3, 192, 1, 2, 3, 4);     // registered
        P(double, 1536, 8, 8, 8, 3, 192, 1, 2, 3, 4);
        P(double, 1536, 16, 16, 6, 1, 96, 2, 2, 3, 4);
        B(double, 2304, 8, 8, 4, 9, 288, 1, 2, 3, 4);     // registered
        P(double, 2304, 8, 8, 4, 9, 288, 1, 2, 3, 4);
        P(double, 2304, 16, 16, 9, 1, 144, 1, 2, 3, 4);
        B(double, 3072, 8, 8, 8, 6, 384, 1, 1, 3, 4);     // registered
        P(double, 3072, 8, 8, 8, 6, 384, 1, 1, 3, 4);
        P(double, 3072, 16, 16, 12, 1, 192, 1, 1, 3, 4);
        P(double, 768, 8, 8, 12, 1, 96, 2, 2, 3, 4);
        B(double, 768, 8, 8, 12, 1, 96, 2, 2, 3, 4);      // registered
        P(double, 1152, 8, 8, 18, 1, 144, 1, 2, 3, 4);
        B(double, 1152, 8, 8, 18, 1, 144, 1, 2, 3, 4);    // registered
        P(double, 8192, 16, 8, 8, 8, 512, 1, 1, 31, 4);   // does not fit with staging? (2 x 128 KiB) -> expected FAILED
    }
    return 0;
}
