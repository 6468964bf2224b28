/*
 * ssfft.h -- C ABI of libssfft.so: the B200 (sm_100a) implementation of the Signalsmith FFT hot path.
 *
 * The reference (/root/reference/signalsmith-fft.h) is a header-only C++ class template with no FFI
 * layer; its public surface is FFT<V> (:326-387) and RealFFT<V,flags> (:402-503).  This ABI is what a
 * binding of that surface binds: one entry point per reference method on the transform path, plain
 * pointers and sizes, int status returns (0 = OK), nothing thrown across the boundary.  The header-only
 * C++ front end include/signalsmith-fft.h forwards to these symbols; INTEGRATION.md shows other bindings.
 *
 * Conventions (all taken from the reference):
 *   - complex data is interleaved (re, im): float2 / double2 == std::complex<float/double>
 *   - forward transform X[k] = sum_n x[n] exp(-2 pi i n k / N); both directions UNNORMALISED
 *     (ifft(fft(x)) == N * x, tests/00-fft.cpp:113-128)
 *   - out-of-place; the input is never modified (tests/00-fft.cpp:44-46); in == out is also accepted
 *   - batches are contiguous: transform b of a C2C plan lives at in + b*N complex; a real plan reads
 *     N reals at in + b*N and writes N/2 complex at out + b*(N/2), with bin 0 packing
 *     (DC, Nyquist) into (.re, .im) (signalsmith-fft.h:459-462)
 *   - there is NO CPU fallback and no cuFFT: every exec call launches hand-written sm_100a kernels
 */
#ifndef SSFFT_H
#define SSFFT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SSFFT_API __attribute__((visibility("default")))
#else
#define SSFFT_API
#endif

typedef struct ssfft_plan ssfft_plan;

/* status codes */
enum {
    SSFFT_OK = 0,
    SSFFT_ERR_INVALID = 1,      /* bad argument (null pointer, wrong plan kind, ...) */
    SSFFT_ERR_CUDA = 2,         /* a CUDA runtime call failed; see ssfft_last_cuda_error() */
    SSFFT_ERR_UNSUPPORTED = 3,  /* size has a prime factor too large for the on-chip paths */
    SSFFT_ERR_NO_DEVICE = 4,    /* no CUDA device: the library never falls back to the CPU */
    SSFFT_ERR_ALLOC = 5
};

/* plan kinds */
enum {
    SSFFT_C2C = 0,           /* FFT<V>                         signalsmith-fft.h:69-387  */
    SSFFT_REAL = 1,          /* RealFFT<V, 0>                  signalsmith-fft.h:393-503 */
    SSFFT_REAL_MODIFIED = 2  /* RealFFT<V, halfFreqShift> == ModifiedRealFFT<V>  :389-391, :505-508 */
};
/* precisions */
enum { SSFFT_F32 = 0, SSFFT_F64 = 1 };
/* directions for ssfft_exec_c2c */
enum { SSFFT_FORWARD = 1, SSFFT_INVERSE = -1 };

/* ---- size helpers (pure integer; usable without a GPU) ---- */
/* FFT<V>::sizeMinimum / sizeMaximum        signalsmith-fft.h:327-348 */
SSFFT_API size_t ssfft_size_minimum(size_t size);
SSFFT_API size_t ssfft_size_maximum(size_t size);
/* RealFFT<V>::sizeMinimum / sizeMaximum    signalsmith-fft.h:403-408 (quirks reproduced) */
SSFFT_API size_t ssfft_real_size_minimum(size_t size);
SSFFT_API size_t ssfft_real_size_maximum(size_t size);

/* ---- plans: replace FFT::setSize -> setPlan (:139-185, :356-363) and RealFFT::setSize (:416-435) ----
 * n is the transform length (the REAL length for real plans; odd n truncates to 2*(n/2) like the
 * reference).  device < 0 means the current device.  Twiddle tables are built here and live in HBM. */
SSFFT_API int ssfft_plan_create(ssfft_plan **out, int kind, int precision, size_t n, int device);
SSFFT_API int ssfft_plan_destroy(ssfft_plan *plan);
/* length the plan transforms (real plans: 2*(n/2), as RealFFT::size() :442-444) */
SSFFT_API size_t ssfft_plan_size(const ssfft_plan *plan);
/* human-readable plan: factors, passes, kernel choice (tests pin the plan builder with this) */
SSFFT_API int ssfft_plan_describe(const ssfft_plan *plan, char *buf, size_t buflen);

/* ---- execution on DEVICE pointers, asynchronous on `stream` (a cudaStream_t, may be NULL) ----
 * Both buffers of a call must be aligned to a whole complex<V> (8 bytes in float, 16 in double) -- the real side too,
 * whose samples move as pairs; other pointers return SSFFT_ERR_INVALID.  Buffers aligned to 16 bytes take the fastest
 * path for the four-step lengths (their tiles are fetched by tensor copies); others run the plan's other path. */
/* FFT<V>::fft (:374-379) when direction == SSFFT_FORWARD, FFT<V>::ifft (:381-386) when SSFFT_INVERSE */
SSFFT_API int ssfft_exec_c2c(ssfft_plan *plan, const void *d_in, void *d_out, size_t batch, int direction,
                             void *stream);
/* RealFFT<V>::fft (:446-473): batch x N reals -> batch x N/2 complex */
SSFFT_API int ssfft_exec_r2c(ssfft_plan *plan, const void *d_real_in, void *d_cplx_out, size_t batch,
                             void *stream);
/* RealFFT<V>::ifft (:475-502): batch x N/2 complex -> batch x N reals (scaled by N) */
SSFFT_API int ssfft_exec_c2r(ssfft_plan *plan, const void *d_cplx_in, void *d_real_out, size_t batch,
                             void *stream);

/* ---- extended execution: the callers either side of the transform ----
 * The reference has no counterpart: its users write these loops on the host around fft() / ifft() -- cut a signal into
 * overlapping frames and multiply by a window before RealFFT::fft (STFT), multiply a spectrum by a filter before
 * ifft (fast convolution), walk the columns of a matrix (2-D transforms).  On a GPU each such loop is one more trip
 * through HBM, so the same layouts and multipliers are applied inside the transform's first load and last store.
 *
 * "Elements" are REALS on the real side of a real plan (input of r2c, output of c2r) and COMPLEX values everywhere
 * else; strides, distances and multiplier tables all count in elements of their side.  A zero field means "default".
 * Every table lives in device memory.  Sizes with a fused kernel run these calls in ONE launch; every other plan runs
 * a gather pass, the plain transform and a scatter pass through a workspace owned by the plan (grown on demand:
 * one extended call in flight per plan). */
enum { SSFFT_MUL_NONE = 0, SSFFT_MUL_REAL = 1, SSFFT_MUL_COMPLEX = 2 };
typedef struct ssfft_io {
    int64_t in_stride;   /* elements between consecutive samples of one transform (0 or 1: contiguous) */
    int64_t in_dist;     /* elements between the first samples of consecutive transforms (0: the transform's length).
                            May be SMALLER than the length: overlapping STFT frames with hop = in_dist */
    int64_t out_stride;  /* same on the output side; the outputs of different transforms must not overlap */
    int64_t out_dist;
    const void *pre;     /* multiplier applied to every input element as it is loaded, NULL = none */
    int32_t pre_kind;    /* SSFFT_MUL_REAL: one real per element (a window; on a complex side it scales re and im);
                            SSFFT_MUL_COMPLEX: one complex per element, complex sides only (a filter).  On the half
                            spectrum of a real plan, bin 0 packs (DC, Nyquist) and is multiplied component by component,
                            which is what the product of two such spectra means */
    int32_t post_kind;
    int64_t pre_dist;    /* elements between the multipliers of consecutive transforms, 0 = one table shared by all */
    const void *post;    /* multiplier applied to every output element as it is stored, NULL = none */
    int64_t post_dist;
} ssfft_io;
/* io == NULL behaves exactly like the plain call.  in == out is accepted when both sides have the same byte layout.
 * SSFFT_ERR_INVALID: negative field, overlapping outputs, a complex multiplier on a real side, unknown kind. */
SSFFT_API int ssfft_exec_c2c_ex(ssfft_plan *plan, const void *d_in, void *d_out, size_t batch, int direction,
                                const ssfft_io *io, void *stream);
SSFFT_API int ssfft_exec_r2c_ex(ssfft_plan *plan, const void *d_real_in, void *d_cplx_out, size_t batch,
                                const ssfft_io *io, void *stream);
SSFFT_API int ssfft_exec_c2r_ex(ssfft_plan *plan, const void *d_cplx_in, void *d_real_out, size_t batch,
                                const ssfft_io *io, void *stream);

/* ---- execution on HOST pointers: the call a reference user makes (fft(in, out) on host containers, reference
 * signalsmith-fft.h:374-386, :446-502).  Returns when `h_out` is complete.  op = 0 C2C forward, 1 C2C inverse, 2 R2C,
 * 3 C2R; in/out mean what they mean for the device calls.
 *   - calls of at most 256 KiB per side go through pinned, mapped staging buffers of the plan: the kernels read and
 *     write them over PCIe themselves (no cudaMemcpy, one launch, one synchronisation);
 *   - larger calls are cut into slices that move through a ring of three device buffers per side on three streams of
 *     the plan (upload / kernels / download, linked by events), so device memory is bounded by the ring, the copies of
 *     neighbouring slices overlap the kernels, and all kernels of a call run in order on one stream.
 * h_in / h_out may be pageable; pinned memory (ssfft_host_alloc, cudaHostAlloc) is what makes the copies overlap.
 * A plan is not re-entrant (ssfft_plan: scratch, staging and counters are per plan): calls on one plan are serialised. */
SSFFT_API int ssfft_exec_host(ssfft_plan *plan, int op, const void *h_in, void *h_out, size_t batch);
SSFFT_API int ssfft_host_alloc(void **h_ptr, size_t bytes); /* pinned host memory */
SSFFT_API int ssfft_host_free(void *h_ptr);

/* ---- device-memory helpers so the header-only front end needs no CUDA headers ---- */
SSFFT_API int ssfft_device_count(int *count);
SSFFT_API int ssfft_malloc(void **d_ptr, size_t bytes);
SSFFT_API int ssfft_free(void *d_ptr);
SSFFT_API int ssfft_memcpy_h2d(void *d_dst, const void *h_src, size_t bytes, void *stream);
SSFFT_API int ssfft_memcpy_d2h(void *h_dst, const void *d_src, size_t bytes, void *stream);
SSFFT_API int ssfft_stream_synchronize(void *stream);
/* synthetic inputs, i.i.d. uniform [-0.5, 0.5) from the counter-based generator shared with the
 * oracle: count scalars starting at scalar index first_idx */
SSFFT_API int ssfft_fill_uniform(void *d_dst, size_t count, int precision, uint64_t seed, uint64_t first_idx,
                                 void *stream);

/* ---- local building blocks of the distributed four-step (single 1-D transform sharded over GPUs,
 * SURVEY.md section 8e; orchestration in fft_b200/dist.py).  Complex elements, device pointers. ----
 * out[b][c][r] = in[b][r][c] * W_N^((row0 + r) * c)   (conjugated when inverse != 0; n_total == 0: no twiddle) */
SSFFT_API int ssfft_transpose_twiddle(const void *d_in, void *d_out, size_t batch, size_t rows, size_t cols,
                                      size_t row0, uint64_t n_total, int inverse, int precision, void *stream);
/* out[b][a][c] = in[a][b][c] : swap the two outer dimensions of an [A][B][run] array */
SSFFT_API int ssfft_permute102(const void *d_in, void *d_out, size_t A, size_t B, size_t run, int precision,
                               void *stream);

/* ---- fused exchange over NVLink peer memory (one process per GPU; buffers shared with CUDA IPC) ----
 * ssfft_ipc_export / import: make a ssfft_malloc'ed buffer visible to the other ranks (64-byte opaque handle).
 * ssfft_exchange_transpose: src[rows][cols] is cut into `world` column blocks; block j is written, transposed and
 * optionally multiplied by W_N^((row0 + r) * c), directly into rank j's buffer at
 * dst_j[cl * dst_pitch + dst_col0 + r]  -- transpose + all-to-all + placement in one kernel, no NCCL on the path. */
SSFFT_API int ssfft_ipc_export(void *d_ptr, void *handle64);
SSFFT_API int ssfft_ipc_import(const void *handle64, void **d_ptr);
SSFFT_API int ssfft_ipc_close(void *d_ptr);
SSFFT_API int ssfft_exchange_transpose(const void *d_src, void *const *d_dst_ptrs, int world, size_t rows, size_t cols,
                                       size_t dst_pitch, size_t dst_col0, size_t row0, uint64_t n_total, int inverse,
                                       int precision, void *stream);
SSFFT_API int ssfft_memcpy_d2d(void *d_dst, const void *d_src, size_t bytes, void *stream);

/* ---- ONE transform sharded over the GPUs of this process (BASELINE config 5: N = 2^30 on 2 / 4 / 8 GPUs) ----
 * The reference has no multi-device path; this is the four-step decomposition N = N1 * N2 (the GPU analogue of its
 * cache-blocking branch, signalsmith-fft.h:130-133) with the all-to-all transposes done by peer stores over NVLink.
 * devices[r] owns block r of the natural order: d_in_shards[r] / d_out_shards[r] point at N / ndev complex elements in
 * that device's memory.  exec is asynchronous on streams of the plan; synchronize (or wait, from the caller's own
 * stream) before reading the output or reusing the input.  A device may be listed several times (logical ranks).
 * flags: SSFFT_DIST_TRANSPOSED_OUTPUT skips the third exchange, output shard r = rows k1 in [r N1/P, (r+1) N1/P) of
 * X[k1 + N1 k2], laid out [N1/P][N2].  SSFFT_ERR_UNSUPPORTED: N has no split with both factors divisible by ndev, or
 * the devices cannot reach each other's memory. */
#define SSFFT_DIST_TRANSPOSED_OUTPUT 1
typedef struct ssfft_dist_plan ssfft_dist_plan;
SSFFT_API int ssfft_dist_plan_create(ssfft_dist_plan **out, int precision, size_t n, int ndev, const int *devices, int flags);
SSFFT_API int ssfft_dist_plan_destroy(ssfft_dist_plan *plan);
SSFFT_API int ssfft_dist_plan_describe(const ssfft_dist_plan *plan, char *buf, size_t buflen);
SSFFT_API size_t ssfft_dist_plan_factor(const ssfft_dist_plan *plan, int which); /* 0: N1, 1: N2 */
SSFFT_API int ssfft_dist_exec_c2c(ssfft_dist_plan *plan, void *const *d_in_shards, void *const *d_out_shards, int direction);
SSFFT_API int ssfft_dist_synchronize(ssfft_dist_plan *plan);
SSFFT_API int ssfft_dist_wait(ssfft_dist_plan *plan, int r, void *stream);

/* ---- sharing the device with kernels of other streams ----
 * The four-step kernels are persistent launches that fill every SM.  A caller that wants another kernel to run beside
 * them (the distributed plan below overlaps its NVLink exchanges with its transforms this way) caps the launch at
 * ctas_per_sm CTAs per SM; 0 restores the default.  Plans without a persistent kernel ignore the call. */
SSFFT_API int ssfft_plan_limit_ctas(ssfft_plan *plan, int ctas_per_sm);

/* ---- diagnostics ---- */
SSFFT_API const char *ssfft_error_string(int status);
SSFFT_API const char *ssfft_last_cuda_error(void);
/* number of kernels this library has launched since load (bench.py reports it as gpu_launches) */
SSFFT_API uint64_t ssfft_launch_count(void);
SSFFT_API const char *ssfft_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SSFFT_H */
