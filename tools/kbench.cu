// tools/kbench.cu -- times fused-kernel template variants in ONE GPU call (development tool, not shipped).
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr \
//        -I fft_b200/csrc -o gpurun_out/kbench tools/kbench.cu && gpurun -- ./gpurun_out/kbench
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
    return 0;
}
