// tools/l2_ceiling.cu -- what does the B200 memory system allow a TWO-PASS (four-step) transform?  (measurement tool)
//
// A four-step FFT whose intermediate stays in L2 moves every point through the L2 slices four times (HBM -> SM, SM -> L2
// scratch, L2 scratch -> SM, SM -> HBM) where a single-pass kernel moves it twice.  This program measures that traffic
// pattern with NO arithmetic, no transposition, no dependencies between CTAs and fully contiguous 16-byte accesses:
//
//   one  : out[i] = in[i]                                       (the single-pass pattern; should reach the copy peak)
//   two  : scratch[slot] = in[tile]; out[tile] = scratch[slot]   (the two-pass pattern, scratch of a few tens of MB in L2)
//   three: one more round trip through a second scratch          (a three-factor plan)
//
// Every number is reported the way bench.py reports a transform: ALGORITHMIC bytes (in + out, 16 B per fp32 complex point)
// per second and as a fraction of the measured HBM copy peak -- i.e. the roofline fraction an ideal FFT of that shape could
// reach.  Usage: l2_ceiling [peak_GBps]   (peak defaults to 6439.5, MEASURED_PEAKS.json)
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e_ = (x);                                                      \
        if (e_ != cudaSuccess) {                                                   \
            fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
            exit(1);                                                               \
        }                                                                          \
    } while (0)

__device__ __forceinline__ float4 ld_stream(const float4 *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld_cg(const float4 *p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_stream(float4 *p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_keep(float4 *p, float4 v) {
    asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// TILE float4 per CTA iteration, U per thread.  PASSES = number of passes over the data (1, 2 or 3).
template <int THREADS, int U, int PASSES>
__global__ void __launch_bounds__(THREADS) passes_kernel(const float4 *in, float4 *out, float4 *scratch, long long tiles) {
    constexpr int TILE = THREADS * U;
    float4 *mine = scratch + (long long)blockIdx.x * TILE * 2;  // two private scratch tiles per CTA
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ld_stream(in + t * TILE + u * THREADS + threadIdx.x);
#pragma unroll
        for (int p = 1; p < PASSES; ++p) {
            float4 *s = mine + (p - 1) * TILE;
            // written by one thread, read by another (like a transposition): rotate the tile by a quarter
#pragma unroll
            for (int u = 0; u < U; ++u) st_keep(s + u * THREADS + threadIdx.x, v[u]);
            __syncthreads();
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = ld_cg(s + ((u * THREADS + threadIdx.x + TILE / 4) % TILE));
            __syncthreads();
        }
#pragma unroll
        for (int u = 0; u < U; ++u) st_stream(out + t * TILE + u * THREADS + threadIdx.x, v[u]);
    }
}

template <int THREADS, int U, int PASSES>
static double run(const float4 *in, float4 *out, float4 *scratch, long long n4, int ctas_per_sm, int sms) {
    constexpr int TILE = THREADS * U;
    const long long tiles = n4 / TILE;
    const int grid = ctas_per_sm * sms;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) passes_kernel<THREADS, U, PASSES><<<grid, THREADS>>>(in, out, scratch, tiles);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 10; ++rep) {
        CK(cudaEventRecord(a));
        passes_kernel<THREADS, U, PASSES><<<grid, THREADS>>>(in, out, scratch, tiles);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return 2.0 * (double)tiles * TILE * 16 / (best * 1e-3) / 1e9;  // algorithmic GB/s
}

// The four-step access pattern: a tile is ROWS runs of RUNV float4 (RUNV * 16 bytes) at a pitch of PITCH4 float4 -- the
// column tile of a transform of ROWS x (2 PITCH4) complex points on the way in (SIN), the transposed store of a row tile
// on the way out (SOUT); the scratch round trip in between is contiguous as in the kernels.  Neighbouring CTAs take
// neighbouring tiles of the same transform, as the ticket queue hands them out.
template <int THREADS, int ROWS, int RUNV, bool SIN, bool SOUT, int PASSES>
__global__ void __launch_bounds__(THREADS) fourstep_like_kernel(const float4 *in, float4 *out, float4 *scratch, long long tiles, int pitch4) {
    constexpr int TILE = ROWS * RUNV, U = TILE / THREADS;
    static_assert(TILE % THREADS == 0, "whole tile per pass of the CTA");
    float4 *mine = scratch + (long long)blockIdx.x * TILE;
    const int tpt = pitch4 / RUNV;  // tiles per transform
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const long long b = t / tpt;
        const int j = (int)(t - b * tpt);
        const long long base = b * (long long)ROWS * pitch4 + (long long)j * RUNV;
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = u * THREADS + threadIdx.x, r = e / RUNV, i = e % RUNV;
            v[u] = ld_stream(in + (SIN ? base + (long long)r * pitch4 + i : t * TILE + e));
        }
        if (PASSES > 1) {
#pragma unroll
            for (int u = 0; u < U; ++u) st_keep(mine + u * THREADS + threadIdx.x, v[u]);
            __syncthreads();
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = ld_cg(mine + ((u * THREADS + threadIdx.x + TILE / 4) % TILE));
            __syncthreads();
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = u * THREADS + threadIdx.x, r = e / RUNV, i = e % RUNV;
            st_stream(out + (SOUT ? base + (long long)r * pitch4 + i : t * TILE + e), v[u]);
        }
    }
}

template <int THREADS, int ROWS, int RUNV, bool SIN, bool SOUT, int PASSES>
static double run_like(const float4 *in, float4 *out, float4 *scratch, long long n4, int ctas_per_sm, int sms, int pitch4) {
    constexpr int TILE = ROWS * RUNV;
    const long long tiles = n4 / TILE;
    const int grid = ctas_per_sm * sms;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    auto k = fourstep_like_kernel<THREADS, ROWS, RUNV, SIN, SOUT, PASSES>;
    for (int i = 0; i < 2; ++i) k<<<grid, THREADS>>>(in, out, scratch, tiles, pitch4);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        CK(cudaEventRecord(a));
        k<<<grid, THREADS>>>(in, out, scratch, tiles, pitch4);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return 2.0 * (double)tiles * TILE * 16 / (best * 1e-3) / 1e9;
}

int main(int argc, char **argv) {
    const double peak = argc > 1 ? atof(argv[1]) : 6439.5;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const long long n4 = (1LL << 30) / 16;  // 1 GiB in, 1 GiB out
    float4 *in, *out, *scratch;
    CK(cudaMalloc(&in, n4 * 16));
    CK(cudaMalloc(&out, n4 * 16));
    CK(cudaMalloc(&scratch, 256LL << 20));
    CK(cudaMemset(in, 1, n4 * 16));
    CK(cudaMemset(scratch, 0, 256LL << 20));
    printf("# l2_ceiling: %d SMs, 1 GiB in + 1 GiB out, contiguous 16-byte accesses, no arithmetic; peak %.1f GB/s\n", sms, peak);
    printf("# passes  threads x float4/thread  CTAs/SM  scratch in flight   algorithmic GB/s   fraction of the HBM roofline\n");
    struct Row { const char *what; double gbs; double scratch_mb; };
    auto report = [&](int passes, int threads, int u, int cps, double gbs) {
        const double mb = passes > 1 ? (double)cps * sms * threads * u * 16 * (passes - 1) / 1048576.0 : 0.0;
        printf("  %d       %4d x %d                %2d       %6.1f MiB        %8.1f           %5.1f %%\n", passes, threads, u, cps, mb, gbs, 100.0 * gbs / peak);
    };
    for (int cps : {2, 4, 8}) report(1, 256, 8, cps, run<256, 8, 1>(in, out, scratch, n4, cps, sms));
    for (int cps : {2, 4, 8}) report(2, 256, 8, cps, run<256, 8, 2>(in, out, scratch, n4, cps, sms));
    for (int cps : {2, 4}) report(2, 512, 8, cps, run<512, 8, 2>(in, out, scratch, n4, cps, sms));
    for (int cps : {2, 4}) report(2, 1024, 4, cps, run<1024, 4, 2>(in, out, scratch, n4, cps, sms));
    for (int cps : {4, 8}) report(3, 256, 8, cps, run<256, 8, 3>(in, out, scratch, n4, cps, sms));
    printf("#\n# four-step access pattern (tile = ROWS runs of RUN bytes at the row pitch of the transform), two passes, 4 CTAs/SM unless noted\n");
    printf("# transform        run    strided side        algorithmic GB/s   fraction of the HBM roofline\n");
    auto rep2 = [&](const char *shape, int run, const char *side, double gbs) {
        printf("  %-14s  %4d B  %-18s  %8.1f           %5.1f %%\n", shape, run, side, gbs, 100.0 * gbs / peak);
    };
    // 2^16 = 256 x 256: pitch 2 KiB = 128 float4
    rep2("256 x 256", 128, "in", run_like<256, 256, 8, true, false, 2>(in, out, scratch, n4, 4, sms, 128));
    rep2("256 x 256", 128, "out", run_like<256, 256, 8, false, true, 2>(in, out, scratch, n4, 4, sms, 128));
    rep2("256 x 256", 128, "in + out", run_like<256, 256, 8, true, true, 2>(in, out, scratch, n4, 4, sms, 128));
    rep2("256 x 256", 128, "in + out, 1 pass", run_like<256, 256, 8, true, true, 1>(in, out, scratch, n4, 4, sms, 128));
    rep2("256 x 256", 256, "in + out", run_like<512, 256, 16, true, true, 2>(in, out, scratch, n4, 2, sms, 128));
    rep2("256 x 256", 512, "in + out", run_like<1024, 256, 32, true, true, 2>(in, out, scratch, n4, 2, sms, 128));
    rep2("256 x 256", 64, "in + out", run_like<256, 256, 4, true, true, 2>(in, out, scratch, n4, 8, sms, 128));
    // 2^18 = 512 x 512: pitch 4 KiB
    rep2("512 x 512", 64, "in + out", run_like<256, 512, 4, true, true, 2>(in, out, scratch, n4, 4, sms, 256));
    rep2("512 x 512", 128, "in + out", run_like<512, 512, 8, true, true, 2>(in, out, scratch, n4, 2, sms, 256));
    rep2("512 x 512", 256, "in + out", run_like<1024, 512, 16, true, true, 2>(in, out, scratch, n4, 2, sms, 256));
    // 2^20 = 1024 x 1024: pitch 8 KiB
    rep2("1024 x 1024", 32, "in + out", run_like<256, 1024, 2, true, true, 2>(in, out, scratch, n4, 4, sms, 512));
    rep2("1024 x 1024", 64, "in + out", run_like<512, 1024, 4, true, true, 2>(in, out, scratch, n4, 2, sms, 512));
    rep2("1024 x 1024", 128, "in + out", run_like<1024, 1024, 8, true, true, 2>(in, out, scratch, n4, 2, sms, 512));
    return 0;
}
