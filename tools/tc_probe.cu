// tools/tc_probe.cu -- the tensor-core question of the north star, measured: ONE radix-16 pass of an fp32 FFT as a
// DFT-matrix product on the 5th-generation tensor cores (tcgen05.mma, kind::tf32, accumulator in TMEM) against the same pass
// as the register codelet the library uses.  (measurement tool; nothing in the product path includes this file)
//
// The pass, as the four-step kernels run it: every thread owns one butterfly -- 16 complex points -- does the 16-point DFT,
// multiplies by the inter-pass twiddles and hands the points to the next pass through shared memory.
//
//   fp32 : LDS.64 x16 -> Dft<16> in registers (144 add + 24 mul) -> 15 complex twiddle multiplies -> STS.64 x16
//   tc   : the 16-point DFT of 128 butterflies is D[128 x 32] = A[128 x 32] * B[32 x 32]: A = the butterflies' inputs as
//          (re, im) rows, B = the real 32 x 32 form of the DFT matrix.  tf32 keeps 11 significant bits, so both operands
//          are split hi + lo and three products are accumulated (A_hi B_hi + A_lo B_hi + A_hi B_lo, "3xTF32"):
//          12 tcgen05.mma (M 128, N 32, K 8) issued by ONE thread per 2048 points.  What the other threads still do:
//          split their 32 reals (cvt.rna.tf32 + subtract), store hi and lo as 16 x STS.128 in the canonical K-major UMMA
//          layout (which doubles as the exchange), wait, tcgen05.ld their row of D (32 registers), twiddle multiplies.
//
// Both kernels keep the data on chip and repeat the pass REPS times (the output of one pass is the input of the next), so
// the time is the issue / shared-memory cost of a pass, which is what bounds the four-step kernels.
// Reported: ns per 16-point butterfly per SM-resident CTA set, thread-instructions per point from the SASS of each
// kernel (run `cuobjdump -sass tools/tc_probe.bin`), and the error of one pass against a double-precision DFT.
//
// usage: tc_probe [reps]
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../fft_b200/csrc/codelets.cuh"
#include "../fft_b200/csrc/cplx.cuh"

using namespace ssfft;

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e_ = (x);                                                      \
        if (e_ != cudaSuccess) {                                                   \
            fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
            exit(1);                                                               \
        }                                                                          \
    } while (0)

constexpr int kThreads = 128;  // one butterfly per thread, 128 butterflies = the M of one MMA

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------------------------------------------------------
// fp32 reference pass: exchange through shared memory + register codelet + twiddles
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) fp32_pass_kernel(const cx<float> *in, cx<float> *out, const cx<float> *tw, int reps) {
    __shared__ cx<float> ex[16 * (kThreads + 1)];
    const int m = threadIdx.x;
    cx<float> v[16], w[16];
    const cx<float> *src = in + (size_t)blockIdx.x * 16 * kThreads;
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = src[j * kThreads + m];
#pragma unroll
    for (int j = 1; j < 16; ++j) w[j] = tw[j * kThreads + m];
    for (int r = 0; r < reps; ++r) {
        Dft<16>::run(v);
#pragma unroll
        for (int j = 1; j < 16; ++j) v[j] = cmul(v[j], w[j]);
        // hand the points to "the next pass": a transposing trip through shared memory, as between two passes of a tile
#pragma unroll
        for (int j = 0; j < 16; ++j) ex[j * (kThreads + 1) + m] = v[j];
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = ex[((j + r) & 15) * (kThreads + 1) + m];
        __syncthreads();
    }
    cx<float> *dst = out + (size_t)blockIdx.x * 16 * kThreads;
#pragma unroll
    for (int j = 0; j < 16; ++j) dst[j * kThreads + m] = v[j];
}

// ------------------------------------------------------------------------------------------------------------------
// tensor-core pass
// ------------------------------------------------------------------------------------------------------------------
// canonical K-major, no swizzle: element (row, k) of 32-bit values at (row % 8) * 16 + (row / 8) * SBO + (k / 4) * LBO + (k % 4) * 4
constexpr uint32_t kSboA = 128, kLboA = 128 / 8 * 128;  // A: 128 rows  -> a K-slab of 4 values is 2 KiB
constexpr uint32_t kSboB = 128, kLboB = 32 / 8 * 128;   // B: 32 rows   -> 512 B per slab
constexpr uint32_t kBytesA = 8 * kLboA, kBytesB = 8 * kLboB;

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version of sm_100
    return d;                // layout type 0: no swizzle
}
// kind::tf32, D fp32, A and B K-major, M = 128, N = 32
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ float to_tf32(float x) {  // round to nearest tf32 (11 significant bits): two integer instructions
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);  // (cvt.rna.tf32.f32 expands to ~5 on sm_100a)
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t *bar, unsigned parity) {
    unsigned done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    return done != 0;
}

// G groups of 128 threads per CTA, each with its own A buffers, 32 TMEM columns and mbarrier (a kernel that allocates TMEM
// runs one CTA per SM here, so the overlap between butterflies' latency chains has to come from inside the CTA)
template <int G>
__global__ void __launch_bounds__(kThreads * G) tc_pass_kernel(const cx<float> *in, cx<float> *out, const cx<float> *tw, const float *bmat,
                                                               int reps, int *fault) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr uint32_t kTmemCols = G <= 1 ? 32 : G <= 2 ? 64 : G <= 4 ? 128 : 256;
    const int tid = threadIdx.x, g = tid / kThreads, m = tid % kThreads, warp = m >> 5;
    float *b_hi = reinterpret_cast<float *>(smem);
    float *b_lo = reinterpret_cast<float *>(smem + kBytesB);
    float *a_hi = reinterpret_cast<float *>(smem + 2 * kBytesB + (size_t)g * 2 * kBytesA);
    float *a_lo = reinterpret_cast<float *>(smem + 2 * kBytesB + (size_t)g * 2 * kBytesA + kBytesA);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 2 * kBytesB + (size_t)G * 2 * kBytesA);
    uint64_t *bar = bars + g;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + G);

    // B (hi and lo halves, already in the canonical layout) -> shared memory; barriers; TMEM columns
    for (int i = tid; i < (int)(2 * kBytesB / 4); i += kThreads * G) b_hi[i] = bmat[i];
    if (m == 0) mbar_init(bar, 1);
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot)), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot + (uint32_t)g * 32;  // this group's accumulator columns

    cx<float> v[16], w[16];
    const size_t tile = (size_t)blockIdx.x * G + g;
    const cx<float> *src = in + tile * 16 * kThreads;
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = src[j * kThreads + m];
#pragma unroll
    for (int j = 1; j < 16; ++j) w[j] = tw[j * kThreads + m];
    w[0] = mk<float>(1.f, 0.f);

    const uint32_t row_off = (uint32_t)(m & 7) * 16 + (uint32_t)(m >> 3) * kSboA;  // bytes, this thread's row of A
    const uint64_t da_hi = umma_desc(smem_addr(a_hi), kLboA, kSboA), da_lo = umma_desc(smem_addr(a_lo), kLboA, kSboA);
    const uint64_t db_hi = umma_desc(smem_addr(b_hi), kLboB, kSboB), db_lo = umma_desc(smem_addr(b_lo), kLboB, kSboB);
    unsigned phase = 0;
    for (int r = 0; r < reps; ++r) {
        // split: hi = the value rounded to tf32, lo = the rest (the tensor core drops lo's low bits: 2^-21 relative)
        // k = 2 j (re), 2 j + 1 (im): the K-slab of four values k / 4 = j / 2 holds points j, j + 1
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const int j = 2 * s;
            float4 h, l;
            h.x = to_tf32(v[j].x); h.y = to_tf32(v[j].y); h.z = to_tf32(v[j + 1].x); h.w = to_tf32(v[j + 1].y);
            l.x = v[j].x - h.x; l.y = v[j].y - h.y; l.z = v[j + 1].x - h.z; l.w = v[j + 1].y - h.w;
            *reinterpret_cast<float4 *>(reinterpret_cast<unsigned char *>(a_hi) + row_off + (uint32_t)s * kLboA) = h;
            *reinterpret_cast<float4 *>(reinterpret_cast<unsigned char *>(a_lo) + row_off + (uint32_t)s * kLboA) = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> the tensor core's reads
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(kThreads) : "memory");  // the group's 128 threads
        if (m == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {  // K = 32 in steps of 8 = two slabs
                const uint64_t adv_a = (uint64_t)((2 * ks * kLboA) >> 4), adv_b = (uint64_t)((2 * ks * kLboB) >> 4);
                umma_tf32(tmem, da_hi + adv_a, db_hi + adv_b, ks > 0);
                umma_tf32(tmem, da_lo + adv_a, db_hi + adv_b, 1);
                umma_tf32(tmem, da_hi + adv_a, db_lo + adv_b, 1);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
        }
        // bounded wait: a descriptor mistake must end as a reported fault, not as a hung GPU
        {
            long long t0 = clock64();
            while (!mbar_try(bar, phase)) {
                if (clock64() - t0 > 2000000000LL) {
                    if (m == 0) atomicExch(fault, 1);
                    __trap();
                }
            }
        }
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t d[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, "
            "%26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]), "=r"(d[9]),
              "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15]), "=r"(d[16]), "=r"(d[17]), "=r"(d[18]),
              "=r"(d[19]), "=r"(d[20]), "=r"(d[21]), "=r"(d[22]), "=r"(d[23]), "=r"(d[24]), "=r"(d[25]), "=r"(d[26]), "=r"(d[27]),
              "=r"(d[28]), "=r"(d[29]), "=r"(d[30]), "=r"(d[31])
            : "r"(tmem + ((uint32_t)(warp * 32) << 16)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = cmul(mk<float>(__uint_as_float(d[2 * j]), __uint_as_float(d[2 * j + 1])), w[j]);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");  // D is read before the next pass overwrites it
    }
    cx<float> *dst = out + tile * 16 * kThreads;
#pragma unroll
    for (int j = 0; j < 16; ++j) dst[j * kThreads + m] = v[j];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "n"(kTmemCols) : "memory");
}

// ------------------------------------------------------------------------------------------------------------------
static void fill_bmat(std::vector<float> &h) {
    // B[n][k], n = output real index (2 q = Re X[q], 2 q + 1 = Im X[q]), k = input real index (2 j, 2 j + 1):
    // X[q] = sum_j x[j] W^(q j), W = exp(-2 pi i / 16).  hi = rounded to tf32 (11 significant bits), lo = the rest.
    h.assign(2 * kBytesB / 4, 0.f);
    const double pi2 = 2.0 * 3.14159265358979323846;
    for (int n = 0; n < 32; ++n)
        for (int k = 0; k < 32; ++k) {
            const int q = n / 2, j = k / 2;
            const double c = cos(pi2 * (q * j % 16) / 16.0), s = -sin(pi2 * (q * j % 16) / 16.0);  // W^(qj) = c + i s
            double val;
            if (n % 2 == 0) val = (k % 2 == 0) ? c : -s;  // Re X += c xr - s xi
            else val = (k % 2 == 0) ? s : c;              // Im X += s xr + c xi
            float f = (float)val;
            uint32_t bits;
            memcpy(&bits, &f, 4);
            bits = (bits + 0x1000u) & 0xffffe000u;  // round to nearest tf32
            float hi;
            memcpy(&hi, &bits, 4);
            const float lo = (float)(val - (double)hi);
            const size_t off = (size_t)(n % 8) * 16 + (size_t)(n / 8) * kSboB + (size_t)(k / 4) * kLboB + (size_t)(k % 4) * 4;
            h[off / 4] = hi;
            h[(kBytesB + off) / 4] = lo;
        }
}

template <int G>
static size_t tc_smem() { return 2 * kBytesB + (size_t)G * 2 * kBytesA + 8 * G + 64; }

struct Bufs { float *dx, *dy, *dtw, *db; int *dfault; std::vector<float> hx, htw; int sms; };

static double check_pass(const Bufs &B, const char *name, int tiles) {
    std::vector<float> hy((size_t)2 * 16 * kThreads * tiles);
    CK(cudaMemcpy(hy.data(), B.dy, hy.size() * 4, cudaMemcpyDeviceToHost));
    long double num = 0, den = 0;
    for (int t = 0; t < tiles; ++t)
        for (int m = 0; m < kThreads; ++m) {
            const size_t base = (size_t)t * 16 * kThreads;
            double xr[16], xi[16];
            for (int j = 0; j < 16; ++j) { xr[j] = B.hx[2 * (base + j * kThreads + m)]; xi[j] = B.hx[2 * (base + j * kThreads + m) + 1]; }
            for (int q = 0; q < 16; ++q) {
                double sr = 0, si = 0;
                for (int j = 0; j < 16; ++j) {
                    const double a = -2.0 * 3.14159265358979323846 * (q * j % 16) / 16.0;
                    sr += xr[j] * cos(a) - xi[j] * sin(a);
                    si += xr[j] * sin(a) + xi[j] * cos(a);
                }
                const double tr = B.htw[2 * (q * kThreads + m)], ti = B.htw[2 * (q * kThreads + m) + 1];
                const double wr = q ? sr * tr - si * ti : sr, wi = q ? sr * ti + si * tr : si;
                const double gr = hy[2 * (base + q * kThreads + m)], gi = hy[2 * (base + q * kThreads + m) + 1];
                num += (gr - wr) * (gr - wr) + (gi - wi) * (gi - wi);
                den += wr * wr + wi * wi;
            }
        }
    const double err = (double)sqrtl(num / den);
    printf("%-34s one pass, relative L2 error vs double: %.3e   (parity bar of a transform: 1e-6 * log2 N)\n", name, err);
    return err;
}

template <int G>
static void run_tc(const Bufs &B, int reps, double pts_per_tile_set) {
    const size_t sm = tc_smem<G>();
    auto kern = tc_pass_kernel<G>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads * G, sm));
    const int grid = B.sms * (occ > 0 ? occ : 1);
    char name[96];
    snprintf(name, sizeof(name), "tcgen05 3xTF32, %d x 128 threads/CTA", G);
    kern<<<grid, kThreads * G, sm>>>((cx<float> *)B.dx, (cx<float> *)B.dy, (cx<float> *)B.dtw, B.db, 1, B.dfault);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(2); }
    check_pass(B, name, G);
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int it = 0; it < 4; ++it) {
        CK(cudaEventRecord(a));
        kern<<<grid, kThreads * G, sm>>>((cx<float> *)B.dx, (cx<float> *)B.dy, (cx<float> *)B.dtw, B.db, reps, B.dfault);
        CK(cudaEventRecord(b));
        e = cudaEventSynchronize(b);
        if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); exit(2); }
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (it > 0 && ms < best) best = ms;
    }
    const double pts = (double)grid * G * 16 * kThreads * reps;
    printf("  %d CTA(s)/SM x %d groups = %2d warps/SM: %8.3f ms = %7.1f Gpoint/s per pass\n", occ, G, occ * G * 4, best, pts / best / 1e6);
    (void)pts_per_tile_set;
}

static void run_fp(const Bufs &B, int reps, int ctas_per_sm) {
    const int grid = B.sms * ctas_per_sm;
    fp32_pass_kernel<<<grid, kThreads>>>((cx<float> *)B.dx, (cx<float> *)B.dy, (cx<float> *)B.dtw, 1);
    CK(cudaDeviceSynchronize());
    if (ctas_per_sm == 1) check_pass(B, "fp32 codelet", 1);
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int it = 0; it < 4; ++it) {
        CK(cudaEventRecord(a));
        fp32_pass_kernel<<<grid, kThreads>>>((cx<float> *)B.dx, (cx<float> *)B.dy, (cx<float> *)B.dtw, reps);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (it > 0 && ms < best) best = ms;
    }
    const double pts = (double)grid * 16 * kThreads * reps;
    printf("fp32 codelet, %2d CTA(s)/SM of 128 threads = %2d warps/SM: %8.3f ms = %7.1f Gpoint/s per pass\n", ctas_per_sm, 4 * ctas_per_sm, best,
           pts / best / 1e6);
}

int main(int argc, char **argv) {
    const int reps = argc > 1 ? atoi(argv[1]) : 2000;
    Bufs B;
    CK(cudaDeviceGetAttribute(&B.sms, cudaDevAttrMultiProcessorCount, 0));
    const int max_tiles = B.sms * 16;
    const size_t pts = (size_t)max_tiles * 16 * kThreads;
    B.hx.resize(2 * pts);
    B.htw.resize(2 * 16 * kThreads);
    std::vector<float> hb;
    uint64_t st = 88172645463325252ULL;
    auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (float)((st >> 11) * (1.0 / 9007199254740992.0) - 0.5); };
    for (auto &f : B.hx) f = rnd();
    for (int j = 0; j < 16; ++j)
        for (int m = 0; m < kThreads; ++m) {  // unit-modulus twiddles, as W_256^(j m')
            const double a = -2.0 * 3.14159265358979323846 * (double)(j * (m % 16)) / 256.0;
            B.htw[2 * (j * kThreads + m)] = (float)cos(a);
            B.htw[2 * (j * kThreads + m) + 1] = (float)sin(a);
        }
    fill_bmat(hb);
    CK(cudaMalloc(&B.dx, B.hx.size() * 4));
    CK(cudaMalloc(&B.dy, B.hx.size() * 4));
    CK(cudaMalloc(&B.dtw, B.htw.size() * 4));
    CK(cudaMalloc(&B.db, hb.size() * 4));
    CK(cudaMalloc(&B.dfault, 4));
    CK(cudaMemset(B.dfault, 0, 4));
    CK(cudaMemcpy(B.dx, B.hx.data(), B.hx.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(B.dtw, B.htw.data(), B.htw.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(B.db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
    printf("# tc_probe: one radix-16 pass (16-point DFT + twiddles + hand-over through shared memory), data on chip, %d passes per launch\n", reps);
    for (int c : {1, 2, 4, 6, 8, 12}) run_fp(B, reps, c);
    run_tc<1>(B, reps, 0);
    run_tc<2>(B, reps, 0);
    run_tc<4>(B, reps, 0);
    run_tc<6>(B, reps, 0);
    printf("for scale: 100 %% of the HBM roofline is 402 Gpoint/s for a WHOLE fp32 transform (16 B per point at 6.44 TB/s); a 2^16\n"
           "transform is four such passes plus its loads, stores and four-step twiddles\n");
    return 0;
}
