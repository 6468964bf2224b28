// Host stand-in for <cuda_runtime.h> (TEST INFRASTRUCTURE): lets g++ compile the library's kernel sources
// (fft_b200/csrc/*.cuh) UNMODIFIED so that tests/host/*_emul.cpp can execute them on the CPU, one CUDA thread per
// fiber (simt_emul.h).  Put this directory first on the include path and define __CUDACC__ and SSFFT_EMUL.
// Nothing here is used by the product: libssfft.so is built by nvcc against the real CUDA headers.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
// Static shared variables become function-local statics (one copy per kernel instantiation, shared by the fibers of
// the CTA; kernels that keep several CTAs alive use dynamic shared memory only, which is per CTA: SSFFT_DYNAMIC_SMEM).
#define __shared__ static
#define __grid_constant__

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;

struct __attribute__((aligned(8))) float2 { float x, y; };
struct __attribute__((aligned(16))) double2 { double x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline double2 make_double2(double x, double y) { return double2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

typedef struct simt_stream_st *cudaStream_t;
typedef struct simt_event_st *cudaEvent_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };

// cache-hinted accesses are plain accesses on the host -- but a misaligned vector access, which x86 tolerates and
// the GPU does not, aborts
inline void simt_check_aligned(const void *p, size_t size) {
    if ((uintptr_t)p % size) { fprintf(stderr, "simt: misaligned %zu-byte access at %p\n", size, p); abort(); }
}
template <typename V> inline V __ldcs(const V *p) { simt_check_aligned(p, sizeof(V)); return *p; }
template <typename V> inline V __ldg(const V *p) { simt_check_aligned(p, sizeof(V)); return *p; }
template <typename V> inline V __ldcg(const V *p) { simt_check_aligned(p, sizeof(V)); return *p; }
inline unsigned atomicAdd(unsigned *p, unsigned v) { const unsigned old = *p; *p += v; return old; }  // fibers never preempt
template <typename V> inline void __stcs(V *p, V v) { simt_check_aligned(p, sizeof(V)); *p = v; }

void __syncthreads();  // yields to the block scheduler (simt_emul.h)
inline void __threadfence() {}
inline void __threadfence_block() {}
inline void __nanosleep(unsigned) {}
inline long long clock64() { return 0; }
[[noreturn]] inline void __trap() { abort(); }
inline size_t __cvta_generic_to_shared(const void *p) { return (size_t)(uintptr_t)p; }

// "device" memory is host memory here, so the library's table builders (build_fused_twiddles ...) run unmodified
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
inline cudaError_t cudaMalloc(void **p, size_t bytes) { *p = aligned_alloc(256, (bytes + 255) / 256 * 256); return *p ? 0 : 2; }
inline cudaError_t cudaFree(void *p) { free(p); return 0; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t bytes, cudaMemcpyKind) { memcpy(d, s, bytes); return 0; }

inline void sincospi(double x, double *s, double *c) { *s = std::sin(M_PI * x); *c = std::cos(M_PI * x); }
inline void sincospif(float x, float *s, float *c) { *s = std::sin((float)M_PI * x); *c = std::cos((float)M_PI * x); }
