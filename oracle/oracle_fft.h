/*
 * oracle_fft.h -- CPU restatement of the Signalsmith FFT hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity oracle for the B200 kernels.  It restates, in plain C, the algorithm of
 * the reference header /root/reference/signalsmith-fft.h (FFT<V> :69-387, RealFFT<V,flags>
 * :393-503).  Every function cites the reference lines it follows.
 *
 * PARITY IS PINNED: tests/test_oracle.py checks this port (a) against golden outputs generated
 * from the reference header itself (tests/golden/, script tests/golden/make_golden.py), (b)
 * against the compiled reference (oracle/_ref/libssfft_ref.so, built by oracle/Makefile from
 * /root/reference/signalsmith-fft.h where it lies) whenever that library is present, and (c)
 * against the known-answer properties of the reference's own tests (tests/00-fft.cpp,
 * tests/01-real.cpp).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product path (fft_b200/, include/) never does: it has no CPU fallback.
 */
#ifndef SSFFT_ORACLE_FFT_H
#define SSFFT_ORACLE_FFT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- size helpers: signalsmith-fft.h:317-348 (FFT), :403-408 (RealFFT) ---- */
size_t oracle_fft_size_minimum(size_t size);
size_t oracle_fft_size_maximum(size_t size);
size_t oracle_realfft_size_minimum(size_t size);
size_t oracle_realfft_size_maximum(size_t size);

/* ---- complex FFT plans (FFT<float> / FFT<double>), interleaved (re,im) buffers ---- */
typedef struct oracle_plan_f32 oracle_plan_f32;
typedef struct oracle_plan_f64 oracle_plan_f64;

oracle_plan_f32 *oracle_plan_create_f32(size_t n);
oracle_plan_f64 *oracle_plan_create_f64(size_t n);
void oracle_plan_destroy_f32(oracle_plan_f32 *p);
void oracle_plan_destroy_f64(oracle_plan_f64 *p);
/* out-of-place; inverse != 0 -> ifft (unnormalised both ways) */
void oracle_fft_f32(oracle_plan_f32 *p, const float *in, float *out, int inverse);
void oracle_fft_f64(oracle_plan_f64 *p, const double *in, double *out, int inverse);

/* plan introspection (used by tests to pin rows F/A/P of SURVEY.md section 8a) */
size_t oracle_plan_num_factors_f32(const oracle_plan_f32 *p);
size_t oracle_plan_num_factors_f64(const oracle_plan_f64 *p);
size_t oracle_plan_factor_f32(const oracle_plan_f32 *p, size_t i);
size_t oracle_plan_factor_f64(const oracle_plan_f64 *p, size_t i);
size_t oracle_plan_num_steps_f32(const oracle_plan_f32 *p);
size_t oracle_plan_num_steps_f64(const oracle_plan_f64 *p);
/* fills 5 size_t: type(0 generic,2,3,4), factor, startIndex, innerRepeats, outerRepeats */
void oracle_plan_step_f32(const oracle_plan_f32 *p, size_t i, size_t *out5);
void oracle_plan_step_f64(const oracle_plan_f64 *p, size_t i, size_t *out5);
size_t oracle_plan_num_twiddles_f32(const oracle_plan_f32 *p);
size_t oracle_plan_num_twiddles_f64(const oracle_plan_f64 *p);
/* permutation as dest-index-per-source: perm[to] = from (signalsmith-fft.h:288-293) */
void oracle_plan_permutation_f32(const oracle_plan_f32 *p, size_t *perm_n);
void oracle_plan_permutation_f64(const oracle_plan_f64 *p, size_t *perm_n);

/* ---- real FFT plans (RealFFT<V, flags>); modified != 0 -> FFTOptions::halfFreqShift ---- */
typedef struct oracle_rplan_f32 oracle_rplan_f32;
typedef struct oracle_rplan_f64 oracle_rplan_f64;

oracle_rplan_f32 *oracle_rplan_create_f32(size_t n, int modified);
oracle_rplan_f64 *oracle_rplan_create_f64(size_t n, int modified);
void oracle_rplan_destroy_f32(oracle_rplan_f32 *p);
void oracle_rplan_destroy_f64(oracle_rplan_f64 *p);
/* real size actually used (2*(n/2), signalsmith-fft.h:442-444) */
size_t oracle_rplan_size_f32(const oracle_rplan_f32 *p);
size_t oracle_rplan_size_f64(const oracle_rplan_f64 *p);
/* fft: N reals -> N/2 complex (bin 0 packs DC,Nyquist); ifft: N/2 complex -> N reals (scaled by N) */
void oracle_rfft_f32(oracle_rplan_f32 *p, const float *in, float *out);
void oracle_rfft_f64(oracle_rplan_f64 *p, const double *in, double *out);
void oracle_irfft_f32(oracle_rplan_f32 *p, const float *in, float *out);
void oracle_irfft_f64(oracle_rplan_f64 *p, const double *in, double *out);

/* ---- batched drivers: one plan per thread (reference objects are not re-entrant) ----
 * kind: 0 = C2C forward, 1 = C2C inverse, 2 = R2C, 3 = C2R, 4 = modified R2C, 5 = modified C2R
 * prec: 0 = f32, 1 = f64.  n is the transform length (real length for kinds 2..5).
 * Transform b reads in + b*in_stride scalars ... layouts are contiguous:
 *   C2C: in/out b*2n scalars;  R2C: in b*n scalars, out b*n scalars (n/2 complex);  C2R mirror.
 * Returns wall seconds of the slowest thread's compute loop (plan build excluded), <0 on error.
 */
double oracle_batch(int kind, int prec, size_t n, size_t batch, const void *in, void *out,
                    int threads);

/* ---- shared synthetic-input generator ----
 * value(seed, idx) = top bits of splitmix64(seed * 2^40 + idx) mapped to [-0.5, 0.5).
 * f32 uses the top 24 bits, f64 the top 53 bits.  idx counts scalars (re, im interleaved).
 */
void oracle_fill_uniform_f32(float *dst, size_t count, uint64_t seed, uint64_t first_idx);
void oracle_fill_uniform_f64(double *dst, size_t count, uint64_t seed, uint64_t first_idx);

#ifdef __cplusplus
}
#endif
#endif
