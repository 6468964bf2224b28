// dist.cu -- ONE 1-D complex transform sharded over several GPUs of one process (ssfft_dist_* of include/ssfft.h).
//
// BASELINE config 5 (N = 2^30) behind the C ABI: the four-step decomposition N = N1 * N2 with the input and the output
// block-distributed in natural order over P devices.  What a call enqueues, per device r:
//
//   exchange 1   x_r [a][N2]  --transpose, peer stores-->  A_q [b][N1]  of every device q          (a = N1/P, b = N2/P)
//   per chunk c  FFT_N1 of rows chunk c of A_r -> w_r          (stream "fft")
//                exchange 2 of that chunk: w_r --transpose * W_N^(n2 k1), peer stores--> B_q [a][N2]   (stream "xchg")
//   per chunk c  FFT_N2 of rows chunk c of B_r -> A_r          (stream "fft")
//                exchange 3 of that chunk: A_r --transpose, peer stores--> OUT_q [b][N1] = natural order  (stream "xchg")
//
// Every exchange is ONE kernel per chunk (exchange_transpose_kernel, real_kernels.cuh): transpose, twiddle and the
// all-to-all are peer stores straight into the remote HBM over NVLink -- no NCCL, no pack / unpack pass, and the last
// exchange lands in the caller's output shards (no copy-out).  The two streams per device let the exchange of chunk c
// run under the FFT of chunk c + 1; the phases that need every peer's data are ordered by events recorded on the
// peers' streams (cudaStreamWaitEvent across devices), not by host synchronisation.  Peer access is enabled between all
// devices at plan time.  The same device may be listed several times ("logical ranks": single-GPU tests).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/ssfft.h"

namespace {

struct DistDev {
    int device = 0;
    ssfft_plan *p1 = nullptr, *p2 = nullptr;  // batched length-n1 / length-n2 transforms
    void *A = nullptr, *B = nullptr, *w = nullptr;
    cudaStream_t s_fft = nullptr, s_x = nullptr;
    cudaEvent_t e_x1 = nullptr, e_x2 = nullptr, e_done = nullptr;
    std::vector<cudaEvent_t> e_f;  // per chunk: FFT of the chunk finished
};

int fail(const char *) { return SSFFT_ERR_CUDA; }
#define DCU(call)                                   \
    do {                                            \
        if ((call) != cudaSuccess) { cudaGetLastError(); return fail(#call); } \
    } while (0)

}  // namespace

struct ssfft_dist_plan {
    int prec = 0, ndev = 0, flags = 0, chunks = 1;
    size_t n = 0, n1 = 0, n2 = 0, a = 0, b = 0, elem = 0;
    bool first = true;
    std::vector<DistDev> dev;
};

extern "C" {

int ssfft_dist_plan_destroy(ssfft_dist_plan *p) {
    if (!p) return SSFFT_OK;
    for (DistDev &d : p->dev) {
        cudaSetDevice(d.device);
        if (d.s_fft) cudaStreamSynchronize(d.s_fft);
        if (d.s_x) cudaStreamSynchronize(d.s_x);
        if (d.p1) ssfft_plan_destroy(d.p1);
        if (d.p2) ssfft_plan_destroy(d.p2);
        for (void *q : {d.A, d.B, d.w})
            if (q) cudaFree(q);
        for (cudaEvent_t e : d.e_f)
            if (e) cudaEventDestroy(e);
        for (cudaEvent_t e : {d.e_x1, d.e_x2, d.e_done})
            if (e) cudaEventDestroy(e);
        if (d.s_fft) cudaStreamDestroy(d.s_fft);
        if (d.s_x) cudaStreamDestroy(d.s_x);
    }
    delete p;
    return SSFFT_OK;
}

int ssfft_dist_plan_create(ssfft_dist_plan **out, int precision, size_t n, int ndev, const int *devices, int flags) {
    if (!out) return SSFFT_ERR_INVALID;
    *out = nullptr;
    if ((precision != SSFFT_F32 && precision != SSFFT_F64) || ndev < 1 || ndev > 16 || !devices || n == 0) return SSFFT_ERR_INVALID;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { cudaGetLastError(); return SSFFT_ERR_NO_DEVICE; }
    for (int i = 0; i < ndev; ++i)
        if (devices[i] < 0 || devices[i] >= count) return SSFFT_ERR_INVALID;
    // N = N1 * N2, both divisible by P, as balanced as possible (N1 <= N2), each leg at most 2^20
    size_t best1 = 0;
    double best_score = 1e300;
    for (size_t d = 1; d * d <= n; ++d) {
        if (n % d) continue;
        const size_t c1 = d, c2 = n / d;
        if (c1 % (size_t)ndev || c2 % (size_t)ndev || c2 > ((size_t)1 << 20)) continue;
        const double score = (double)c2 / (double)c1;
        if (score < best_score) { best_score = score; best1 = c1; }
    }
    if (!best1) return SSFFT_ERR_UNSUPPORTED;
    int prev = 0;
    cudaGetDevice(&prev);
    ssfft_dist_plan *p = new ssfft_dist_plan();
    p->prec = precision; p->ndev = ndev; p->flags = flags; p->n = n; p->n1 = best1; p->n2 = n / best1;
    p->a = p->n1 / ndev; p->b = p->n2 / ndev;
    p->elem = precision == SSFFT_F32 ? 8 : 16;
    // chunks per phase.  MEASURED (profiles/bench_dist_local_8gpu_r02q.txt, bench_dist_share_r02r.txt): the exchange of
    // chunk c does not run beside the transform of chunk c + 1 to any effect -- 2 GPUs 13.9 / 14.1 / 14.0 / 14.1 ms with
    // 1 / 2 / 4 / 8 chunks, 8 GPUs 5.46 vs 5.67 ms with 1 vs 4, the same with the transforms capped at 2 of 3 CTAs per SM
    // and a higher stream priority: both kernels want the whole SM (the transforms are bound by instruction issue, the
    // exchange needs its full occupancy to keep NVLink busy), so sharing only time-slices them.  Default: one chunk.
    int chunks = 1;
    if (const char *e = getenv("SSFFT_DIST_CHUNKS")) chunks = atoi(e);
    while (chunks > 1 && (p->a % chunks || p->b % chunks)) --chunks;
    p->chunks = chunks < 1 ? 1 : chunks;
    p->dev.resize(ndev);
    int rc = SSFFT_OK;
    const size_t per = n / ndev * p->elem;
    for (int r = 0; r < ndev && rc == SSFFT_OK; ++r) {
        DistDev &d = p->dev[r];
        d.device = devices[r];
        if (cudaSetDevice(d.device) != cudaSuccess) { rc = SSFFT_ERR_CUDA; break; }
        for (int q = 0; q < ndev; ++q) {
            if (devices[q] == d.device) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, d.device, devices[q]);
            if (!can) { rc = SSFFT_ERR_UNSUPPORTED; break; }
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[q], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { rc = SSFFT_ERR_CUDA; break; }
            cudaGetLastError();
        }
        if (rc) break;
        if ((rc = ssfft_plan_create(&d.p1, SSFFT_C2C, precision, p->n1, d.device))) break;
        if ((rc = ssfft_plan_create(&d.p2, SSFFT_C2C, precision, p->n2, d.device))) break;
        if (cudaMalloc(&d.A, per) != cudaSuccess || cudaMalloc(&d.B, per) != cudaSuccess || cudaMalloc(&d.w, per) != cudaSuccess) {
            cudaGetLastError();
            rc = SSFFT_ERR_ALLOC;
            break;
        }
        // the exchange of a chunk runs BESIDE the transform of the next one only if both fit an SM together: the
        // transforms leave a third of every SM free (2 of 3 CTAs; measured no slower), and their stream has the higher
        // priority, so the short-lived exchange CTAs of the stream below never queue in front of them
        if (p->chunks > 1 && !getenv("SSFFT_DIST_NO_SHARE")) {
            ssfft_plan_limit_ctas(d.p1, 2);
            ssfft_plan_limit_ctas(d.p2, 2);
        }
        int prio_least = 0, prio_greatest = 0;
        cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
        bool ok = cudaStreamCreateWithPriority(&d.s_fft, cudaStreamNonBlocking, prio_greatest) == cudaSuccess &&
                  cudaStreamCreateWithPriority(&d.s_x, cudaStreamNonBlocking, prio_least) == cudaSuccess &&
                  cudaEventCreateWithFlags(&d.e_x1, cudaEventDisableTiming) == cudaSuccess &&
                  cudaEventCreateWithFlags(&d.e_x2, cudaEventDisableTiming) == cudaSuccess &&
                  cudaEventCreateWithFlags(&d.e_done, cudaEventDisableTiming) == cudaSuccess;
        d.e_f.assign((size_t)2 * p->chunks, nullptr);
        for (cudaEvent_t &e : d.e_f) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
        if (!ok) { cudaGetLastError(); rc = SSFFT_ERR_CUDA; }
    }
    cudaSetDevice(prev);
    if (rc) { ssfft_dist_plan_destroy(p); return rc; }
    *out = p;
    return SSFFT_OK;
}

int ssfft_dist_plan_describe(const ssfft_dist_plan *p, char *buf, size_t buflen) {
    if (!p || !buf || !buflen) return SSFFT_ERR_INVALID;
    snprintf(buf, buflen, "%s distributed N=%zu = %zu x %zu over %d device(s): 3 exchanges by peer stores (transposed output: %d), "
             "%d chunk(s) per phase, exchange of chunk c under the FFT of chunk c+1", p->prec == SSFFT_F32 ? "f32" : "f64", p->n, p->n1,
             p->n2, p->ndev, (p->flags & SSFFT_DIST_TRANSPOSED_OUTPUT) ? 1 : 0, p->chunks);
    return SSFFT_OK;
}

// d_in_shards[r] / d_out_shards[r]: N / P complex elements on device r (block r of the natural order; with
// SSFFT_DIST_TRANSPOSED_OUTPUT the output shard r holds rows k1 in [r a, (r+1) a) of X[k1 + N1 k2] as [a][N2]).
// Asynchronous: returns when the work is enqueued on the plan's streams; ssfft_dist_synchronize waits for it.
int ssfft_dist_exec_c2c(ssfft_dist_plan *p, void *const *d_in_shards, void *const *d_out_shards, int direction) {
    if (!p || !d_in_shards || !d_out_shards) return SSFFT_ERR_INVALID;
    if (direction != SSFFT_FORWARD && direction != SSFFT_INVERSE) return SSFFT_ERR_INVALID;
    const int P = p->ndev, K = p->chunks, inv = direction == SSFFT_INVERSE;
    const bool transposed = (p->flags & SSFFT_DIST_TRANSPOSED_OUTPUT) != 0;
    const size_t a = p->a, b = p->b, n1 = p->n1, n2 = p->n2, es = p->elem;
    int prev = 0;
    cudaGetDevice(&prev);
    std::vector<void *> tabA(P), tabB(P), tabOut(P);
    for (int q = 0; q < P; ++q) { tabA[q] = p->dev[q].A; tabB[q] = p->dev[q].B; tabOut[q] = d_out_shards[q]; }
    int rc = SSFFT_OK;
    auto wait_all = [&](cudaStream_t s, cudaEvent_t DistDev::*ev) {
        for (int q = 0; q < P; ++q) cudaStreamWaitEvent(s, p->dev[q].*ev, 0);
    };
    // ---- exchange 1 (after every device has finished the previous call: its buffers are free)
    for (int r = 0; r < P && !rc; ++r) {
        DistDev &d = p->dev[r];
        cudaSetDevice(d.device);
        if (!p->first) wait_all(d.s_x, &DistDev::e_done);
        rc = ssfft_exchange_transpose(d_in_shards[r], tabA.data(), P, a, n2, n1, (size_t)r * a, 0, 0, 0, p->prec, d.s_x);
        if (!rc && cudaEventRecord(d.e_x1, d.s_x) != cudaSuccess) rc = SSFFT_ERR_CUDA;
    }
    // ---- columns: FFT_N1 over my b rows of A, chunk by chunk; exchange 2 (+ twiddle) of a chunk under the next FFT
    for (int r = 0; r < P && !rc; ++r) {
        DistDev &d = p->dev[r];
        cudaSetDevice(d.device);
        wait_all(d.s_fft, &DistDev::e_x1);
        if (!p->first) wait_all(d.s_fft, &DistDev::e_done);  // w / A reuse against the previous call's last phase
        const size_t bc = b / K;
        for (int c = 0; c < K && !rc; ++c) {
            char *src = (char *)d.A + (size_t)c * bc * n1 * es, *dst = (char *)d.w + (size_t)c * bc * n1 * es;
            rc = ssfft_exec_c2c(d.p1, src, dst, bc, direction, d.s_fft);
            if (rc) break;
            cudaEventRecord(d.e_f[c], d.s_fft);
            cudaStreamWaitEvent(d.s_x, d.e_f[c], 0);
            rc = ssfft_exchange_transpose(dst, tabB.data(), P, bc, n1, n2, (size_t)r * b + (size_t)c * bc, (size_t)r * b + (size_t)c * bc,
                                          p->n, inv, p->prec, d.s_x);
        }
        if (!rc && cudaEventRecord(d.e_x2, d.s_x) != cudaSuccess) rc = SSFFT_ERR_CUDA;
    }
    // ---- rows: FFT_N2 over my a rows of B; exchange 3 of a chunk (into the output shards) under the next FFT
    for (int r = 0; r < P && !rc; ++r) {
        DistDev &d = p->dev[r];
        cudaSetDevice(d.device);
        wait_all(d.s_fft, &DistDev::e_x2);
        const size_t ac = a / K;
        for (int c = 0; c < K && !rc; ++c) {
            char *src = (char *)d.B + (size_t)c * ac * n2 * es;
            if (transposed) {
                rc = ssfft_exec_c2c(d.p2, src, (char *)d_out_shards[r] + (size_t)c * ac * n2 * es, ac, direction, d.s_fft);
                continue;
            }
            char *dst = (char *)d.A + (size_t)c * ac * n2 * es;  // A is free: its columns were transformed in the last phase
            rc = ssfft_exec_c2c(d.p2, src, dst, ac, direction, d.s_fft);
            if (rc) break;
            cudaEventRecord(d.e_f[K + c], d.s_fft);
            cudaStreamWaitEvent(d.s_x, d.e_f[K + c], 0);
            rc = ssfft_exchange_transpose(dst, tabOut.data(), P, ac, n2, n1, (size_t)r * a + (size_t)c * ac, 0, 0, 0, p->prec, d.s_x);
        }
        if (!rc) {
            if (transposed) { cudaEventRecord(d.e_done, d.s_fft); }
            else { cudaStreamWaitEvent(d.s_x, d.e_f[K + K - 1], 0); cudaEventRecord(d.e_done, d.s_x); }
        }
    }
    p->first = false;
    cudaSetDevice(prev);
    if (rc) return rc;
    if (cudaGetLastError() != cudaSuccess) return SSFFT_ERR_CUDA;
    return SSFFT_OK;
}

// Waits until every device has finished the last call: its own work AND the peers' stores into its output shard.
int ssfft_dist_synchronize(ssfft_dist_plan *p) {
    if (!p) return SSFFT_ERR_INVALID;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaError_t e = cudaSuccess;
    for (DistDev &d : p->dev) {
        cudaSetDevice(d.device);
        const cudaError_t e1 = cudaStreamSynchronize(d.s_fft), e2 = cudaStreamSynchronize(d.s_x);
        if (e == cudaSuccess) e = e1 != cudaSuccess ? e1 : e2;
    }
    cudaSetDevice(prev);
    return e == cudaSuccess ? SSFFT_OK : SSFFT_ERR_CUDA;
}

// The caller's own streams: make `stream` (on device index r of the plan) wait for the completion of the last call.
int ssfft_dist_wait(ssfft_dist_plan *p, int r, void *stream) {
    if (!p || r < 0 || r >= p->ndev) return SSFFT_ERR_INVALID;
    for (DistDev &d : p->dev)
        if (cudaStreamWaitEvent((cudaStream_t)stream, d.e_done, 0) != cudaSuccess) { cudaGetLastError(); return SSFFT_ERR_CUDA; }
    return SSFFT_OK;
}

size_t ssfft_dist_plan_factor(const ssfft_dist_plan *p, int which) { return !p ? 0 : which == 0 ? p->n1 : p->n2; }

}  // extern "C"
