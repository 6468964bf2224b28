// tools/ubench_f32x2.cu -- does Blackwell's packed fp32 (add/mul/fma.f32x2 -> FADD2 / FMUL2 / FFMA2) double the
// fp32 work per issue slot?  Measures warp-instructions per cycle per SM for scalar vs packed chains.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/ubench_f32x2.cu -o tools/ubench_f32x2.bin
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ILP = 8, ITERS = 4096;

template <int MODE>
__global__ void __launch_bounds__(256) k(float2 *out, float2 seed, long long *cycles) {
    float2 a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = make_float2(seed.x + i + threadIdx.x, seed.y - i);
    const float2 b = make_float2(seed.y, seed.x), c = make_float2(0.5f * seed.x, 0.25f * seed.y);
    const long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) { a[i].x = a[i].x + b.x; a[i].y = a[i].y + b.y; }              // 2 FADD
            else if (MODE == 1) { a[i] = __fadd2_rn(a[i], b); }                              // 1 FADD2
            else if (MODE == 2) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }  // 2 FFMA
            else if (MODE == 3) { a[i] = __ffma2_rn(a[i], b, c); }                           // 1 FFMA2
            else if (MODE == 4) { a[i].x = a[i].x * b.x; a[i].y = a[i].y * b.y; }           // 2 FMUL
            else { a[i] = __fmul2_rn(a[i], b); }                                             // 1 FMUL2
        }
    }
    const long long t1 = clock64();
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < ILP; ++i) { s.x += a[i].x; s.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int ctas_per_sm) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * ctas_per_sm;
    float2 *out;
    long long *cyc, h[4096];
    cudaMalloc(&out, sizeof(float2) * grid * 256);
    cudaMalloc(&cyc, sizeof(long long) * grid);
    k<MODE><<<grid, 256>>>(out, make_float2(1.0001f, 0.9999f), cyc);
    cudaDeviceSynchronize();
    k<MODE><<<grid, 256>>>(out, make_float2(1.0001f, 0.9999f), cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof(long long) * (grid < 4096 ? grid : 4096), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < grid && i < 4096; ++i) avg += (double)h[i];
    avg /= (grid < 4096 ? grid : 4096);
    // fp32 element-ops per thread = ITERS * ILP * 2; warps per SM = ctas_per_sm * 8
    const double elem_ops_per_sm = (double)ITERS * ILP * 2 * 256 * ctas_per_sm;
    printf("%-8s ctas/SM %d: %.0f cycles  ->  %.1f fp32 element-ops / cycle / SM   (%s)\n", name, ctas_per_sm, avg,
           elem_ops_per_sm / avg, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    for (int c : {1, 2, 4}) {
        run<0>("FADD", c);  run<1>("FADD2", c);
        run<2>("FFMA", c);  run<3>("FFMA2", c);
        run<4>("FMUL", c);  run<5>("FMUL2", c);
    }
    return 0;
}
