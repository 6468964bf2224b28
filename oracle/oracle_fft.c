/*
 * oracle_fft.c -- CPU restatement of the Signalsmith FFT hot path (TEST INFRASTRUCTURE ONLY).
 * See oracle_fft.h for the rules on who may load this.  Follows /root/reference/signalsmith-fft.h.
 */
#define _GNU_SOURCE
#include "oracle_fft.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846264338327950288
#endif

/* ---- size helpers ---- */

/* validSize filter  signalsmith-fft.h:317-325 : 0,1,2,3,4,6,8,9,12,16,18,24 */
static int valid_size(size_t s) {
    static const unsigned char ok[32] = {1, 1, 1, 1, 1, 0, 1, 0, 1, 1, 0, 0, 1, 0, 0, 0,
                                         1, 0, 1, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0};
    return ok[s];
}
/* FFT::sizeMinimum  :327-337 */
size_t oracle_fft_size_minimum(size_t size) {
    size_t power2 = 1;
    while (size >= 32) {
        size = (size - 1) / 2 + 1;
        power2 *= 2;
    }
    while (size < 32 && !valid_size(size)) ++size;
    return power2 * size;
}
/* FFT::sizeMaximum  :338-348 */
size_t oracle_fft_size_maximum(size_t size) {
    size_t power2 = 1;
    while (size >= 32) {
        size /= 2;
        power2 *= 2;
    }
    while (size > 1 && !valid_size(size)) --size;
    return power2 * size;
}
/* RealFFT::sizeMinimum :403-405, sizeMaximum :406-408 (quirks reproduced) */
size_t oracle_realfft_size_minimum(size_t size) { return (oracle_fft_size_minimum((size - 1) / 2) + 1) * 2; }
size_t oracle_realfft_size_maximum(size_t size) { return oracle_fft_size_minimum(size / 2) * 2; }

/* ---- precision-generic bodies ---- */
#define REAL float
#define NAME(x) x##_f32
#include "oracle_impl.inc"
#undef REAL
#undef NAME

#define REAL double
#define NAME(x) x##_f64
#include "oracle_impl.inc"
#undef REAL
#undef NAME

/* ---- synthetic inputs ---- */
static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
void oracle_fill_uniform_f32(float *dst, size_t count, uint64_t seed, uint64_t first_idx) {
    for (size_t i = 0; i < count; ++i) {
        uint64_t h = splitmix64((seed << 40) + first_idx + i);
        dst[i] = (float)(h >> 40) * (1.0f / 16777216.0f) - 0.5f;
    }
}
void oracle_fill_uniform_f64(double *dst, size_t count, uint64_t seed, uint64_t first_idx) {
    for (size_t i = 0; i < count; ++i) {
        uint64_t h = splitmix64((seed << 40) + first_idx + i);
        dst[i] = (double)(h >> 11) * (1.0 / 9007199254740992.0) - 0.5;
    }
}

/* ---- batched multi-thread driver ---- */
typedef struct {
    int kind, prec;
    size_t n, b0, b1;
    const char *in;
    char *out;
    double seconds;
    pthread_barrier_t *bar;
} job_t;

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void *worker(void *arg) {
    job_t *j = (job_t *)arg;
    const size_t n = j->n;
    const size_t sz = j->prec ? sizeof(double) : sizeof(float);
    const int modified = (j->kind >= 4);
    void *plan = NULL;
    /* one plan object per thread: reference objects hold mutable scratch (:73, :398) */
    if (j->kind <= 1) plan = j->prec ? (void *)oracle_plan_create_f64(n) : (void *)oracle_plan_create_f32(n);
    else plan = j->prec ? (void *)oracle_rplan_create_f64(n, modified) : (void *)oracle_rplan_create_f32(n, modified);
    pthread_barrier_wait(j->bar);
    double t0 = now_s();
    for (size_t b = j->b0; b < j->b1; ++b) {
        if (j->kind <= 1) {
            const char *src = j->in + b * 2 * n * sz;
            char *dst = j->out + b * 2 * n * sz;
            if (j->prec) oracle_fft_f64((oracle_plan_f64 *)plan, (const double *)src, (double *)dst, j->kind);
            else oracle_fft_f32((oracle_plan_f32 *)plan, (const float *)src, (float *)dst, j->kind);
        } else {
            const size_t nr = (n / 2) * 2;
            const char *src = j->in + b * nr * sz;
            char *dst = j->out + b * nr * sz;
            if (j->kind == 2 || j->kind == 4) {
                if (j->prec) oracle_rfft_f64((oracle_rplan_f64 *)plan, (const double *)src, (double *)dst);
                else oracle_rfft_f32((oracle_rplan_f32 *)plan, (const float *)src, (float *)dst);
            } else {
                if (j->prec) oracle_irfft_f64((oracle_rplan_f64 *)plan, (const double *)src, (double *)dst);
                else oracle_irfft_f32((oracle_rplan_f32 *)plan, (const float *)src, (float *)dst);
            }
        }
    }
    j->seconds = now_s() - t0;
    if (j->kind <= 1) { if (j->prec) oracle_plan_destroy_f64((oracle_plan_f64 *)plan); else oracle_plan_destroy_f32((oracle_plan_f32 *)plan); }
    else { if (j->prec) oracle_rplan_destroy_f64((oracle_rplan_f64 *)plan); else oracle_rplan_destroy_f32((oracle_rplan_f32 *)plan); }
    return NULL;
}

double oracle_batch(int kind, int prec, size_t n, size_t batch, const void *in, void *out, int threads) {
    if (kind < 0 || kind > 5 || prec < 0 || prec > 1) return -1.0;
    if (threads < 1) threads = 1;
    if ((size_t)threads > batch) threads = batch ? (int)batch : 1;
    pthread_t *tid = (pthread_t *)malloc((size_t)threads * sizeof(pthread_t));
    job_t *jobs = (job_t *)malloc((size_t)threads * sizeof(job_t));
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, NULL, (unsigned)threads);
    for (int t = 0; t < threads; ++t) {
        jobs[t].kind = kind; jobs[t].prec = prec; jobs[t].n = n;
        jobs[t].b0 = batch * (size_t)t / (size_t)threads;
        jobs[t].b1 = batch * (size_t)(t + 1) / (size_t)threads;
        jobs[t].in = (const char *)in; jobs[t].out = (char *)out;
        jobs[t].seconds = 0; jobs[t].bar = &bar;
        pthread_create(&tid[t], NULL, worker, &jobs[t]);
    }
    double worst = 0;
    for (int t = 0; t < threads; ++t) {
        pthread_join(tid[t], NULL);
        if (jobs[t].seconds > worst) worst = jobs[t].seconds;
    }
    pthread_barrier_destroy(&bar);
    free(tid); free(jobs);
    return worst;
}
