// ssfft.cu -- plan construction, kernel dispatch and the C ABI of libssfft.so (include/ssfft.h).
//
// Host side of the boundary: replaces FFT<V>::setSize/setPlan (signalsmith-fft.h:139-185, :356-363),
// RealFFT<V>::setSize (:416-435) and the run<inverse> dispatcher (:295-315).  There is no CPU path:
// every exec entry point launches sm_100a kernels or returns an error.
#include "../../include/ssfft.h"

#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "bluestein.cuh"
#include "composite.cuh"
#include "exchange_tma.cuh"
#include "tiny.cuh"
#include "cluster.cuh"
#include "ex_request.h"
#include "flat.cuh"
#include "fused.cuh"
#include "generic.cuh"
#include "generic_plan.h"
#include "plan.h"
#include "planner.h"
#include "real_kernels.cuh"
#include "tiled.cuh"

using namespace ssfft;

namespace {

std::atomic<uint64_t> g_launches{0};
thread_local char g_cuda_err[256] = "";

int cuda_fail(cudaError_t e, const char *what) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", what, cudaGetErrorString(e));
    return SSFFT_ERR_CUDA;
}
#define CU(call)                                         \
    do {                                                 \
        cudaError_t e_ = (call);                         \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (dev != prev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Orders a call behind the previous call on the same plan (see plan.h): host-side mutex for the duration of the
// enqueue, device-side event when the stream changes.  Streams under capture are left alone (an outside event would
// break the capture); capturing code uses one plan per captured stream.
struct PlanExec {
    ssfft_plan *pl;
    cudaStream_t s;
    bool capturing = false;
    PlanExec(ssfft_plan *p, cudaStream_t st) : pl(p), s(st) {
        pl->exec_mu.lock();
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(s, &cs) != cudaSuccess) cudaGetLastError();
        capturing = cs != cudaStreamCaptureStatusNone;
        if (!capturing && pl->exec_any && pl->exec_last != s && pl->exec_done) cudaStreamWaitEvent(s, pl->exec_done, 0);
    }
    ~PlanExec() {
        if (!capturing) {
            if (!pl->exec_done && cudaEventCreateWithFlags(&pl->exec_done, cudaEventDisableTiming) != cudaSuccess) pl->exec_done = nullptr;
            if (pl->exec_done && cudaEventRecord(pl->exec_done, s) == cudaSuccess) { pl->exec_last = s; pl->exec_any = true; }
        }
        pl->exec_mu.unlock();
    }
};
bool plan_is_stateful(const ssfft_plan *pl) {  // owns device state that a concurrent call would trample
    return pl->d_scratch || pl->d_flat_scratch || pl->d_ex_in || pl->d_ex_out || pl->bs_inner || pl->comp_inner;
}

int max_optin_smem(int device) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess) return 48 * 1024;
    return v;
}

template <typename T>
int upload_roots(void **d_out, size_t n, size_t count, size_t step) {
    std::vector<T> h(2 * (count ? count : 1));
    fill_roots<T>(h.data(), n, count, step);
    CU(cudaMalloc(d_out, h.size() * sizeof(T)));
    CU(cudaMemcpy(*d_out, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return SSFFT_OK;
}

template <typename T>
int build_generic_stage(GenericStage &st, size_t n, int smem_max) {
    if (!plan_generic_stage(st, n, sizeof(cx<T>), smem_max)) return SSFFT_ERR_UNSUPPORTED;
    int rc = upload_roots<T>(&st.d_roots, n, n, 1);
    if (rc) return rc;
    if (st.smem_bytes > 48 * 1024) {
        CU(cudaFuncSetAttribute(generic_fft_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
    }
    return SSFFT_OK;
}

using Layout = GenericLayout;

template <typename T>
int launch_generic(const ssfft_plan *pl, const GenericStage &st, const void *in, void *out, long long batch,
                   const Layout &L, int inverse, bool epilogue, int ep_cols, cudaStream_t s) {
    if (batch <= 0) return SSFFT_OK;
    const GenericParams<T> p = make_generic_params<T>(st, st.d_roots, batch, L, inverse, epilogue ? pl->d_ep_lo : nullptr,
                                                      epilogue ? pl->d_ep_hi : nullptr, pl->ep_shift, ep_cols);
    const long long blocks = (batch + st.fpb - 1) / st.fpb;
    if (blocks > 0x7fffffffLL) return SSFFT_ERR_INVALID;
    dim3 grid((unsigned)blocks), block((unsigned)st.tx, (unsigned)st.fpb);
    generic_fft_kernel<T><<<grid, block, st.smem_bytes, s>>>((const cx<T> *)in, (cx<T> *)out, p);
    ++g_launches;
    CU(cudaGetLastError());
    return SSFFT_OK;
}

int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

template <typename T>
int launch_tile_flavor(int id, int flavor, const cx<T> *in, cx<T> *out, const void *tw, const void *tw4, int n1, int n2,
                       long long batch, long long in_stride, long long out_stride, int inverse, int ctb_log2,
                       cudaStream_t s) {
    TileParams<T> p;
    p.ctb_log2 = ctb_log2;
    p.in = in; p.out = out;
    p.tw = (const cx<T> *)tw; p.tw4 = (const cx<T> *)tw4;
    p.n1 = n1; p.n2 = n2; p.batch = batch;
    p.in_stride = in_stride; p.out_stride = out_stride; p.inverse = inverse;
    int rc = tile_registry()[id].launch[flavor](&p, s);
    ++g_launches;
    return rc ? SSFFT_ERR_CUDA : SSFFT_OK;
}

// Try to set up the tile-kernel four-step for total length `total` (complex n, or the REAL length for real
// plans).  Returns true when both factors have a tile kernel.
template <typename T>
int setup_tiled(ssfft_plan *pl, size_t total, bool real, bool *ok) {
    *ok = false;
    if (total == 0 || (total & (total - 1))) return SSFFT_OK;  // power-of-two only
    int lg = 0;
    while (((size_t)1 << lg) < total) ++lg;
    // below these sizes a fused single-pass kernel exists and is faster (fp32: up to 16384, fp64: up to 8192)
    const bool f64 = sizeof(T) == 8;
    const int min_lg = real ? env_int("SSFFT_TILE_MIN_LOG2_REAL", f64 ? 15 : 16) : env_int("SSFFT_TILE_MIN_LOG2", f64 ? 14 : 15);
    if (lg < min_lg) return SSFFT_OK;
    size_t n1 = (size_t)1 << (lg / 2), n2 = total / n1;
    int ia = find_tile<T>(n1), ib = find_tile<T>(n2);
    if (ia < 0 || ib < 0) return SSFFT_OK;
    pl->tile_a = ia; pl->tile_b = ib; pl->n1 = n1; pl->n2 = n2;
    int rc = build_tile_twiddles<T>(ia, &pl->d_tile_tw_a);
    if (rc) return rc;
    rc = build_tile_twiddles<T>(ib, &pl->d_tile_tw_b);
    if (rc) return rc;
    // scratch and twiddles are tile-major in blocks of CTB rows, CTB = lanes of the row-stage kernel (tiled.cuh)
    const size_t ctb = (size_t)tile_registry()[ib].ct;
    int ctb_log2 = 0;
    while (((size_t)1 << ctb_log2) < ctb) ++ctb_log2;
    pl->ctb_log2 = ctb_log2;
    std::vector<T> h;
    const size_t rows = fill_fourstep_twiddles<T>(h, total, n1, n2, ctb, real);
    CU(cudaMalloc(&pl->d_tw4, h.size() * sizeof(T)));
    CU(cudaMemcpy(pl->d_tw4, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    pl->scratch_per = rows * n2;
    const size_t per = pl->scratch_per * sizeof(cx<T>);
    // preferred: one persistent cluster launch per call, two scratch slots per co-resident cluster
    pl->fs_id = find_fourstep<T>(n1, n2);
    if (pl->fs_id >= 0) {
        const FourStepEntry &fe = fourstep_registry()[pl->fs_id];
        // CTAs per cluster.  Real 65536 has 8 + 9 tiles per transform: with 4 CTAs one of them runs 3 row tiles while
        // the others run 2 and everyone waits at the next cluster barrier; pairs split 4+4 / 5+4 (measured R2C
        // 35.3 -> 39.7 %, C2R 32.7 -> 36.4 % of roofline).  SSFFT_CLUSTER overrides.
        int cs = fourstep_cluster_size();
        if (!getenv("SSFFT_CLUSTER") && real && total == 65536) cs = 2;
        pl->fs_csize = cs;
        int clusters = 1 << 30;
        for (int kind = real ? 1 : 0; kind <= (real ? 2 : 0); ++kind) {
            int m = fe.max_clusters[kind](cs);
            if (m < clusters) clusters = m;
        }
        if (clusters >= 1) {
            pl->fs_clusters = clusters;
            // Transforms in flight = groups; each owns two scratch slots.  Grow the groups (clusters per transform)
            // until all slots fit the L2 budget or a group already has a CTA per tile.
            // MEASURED (profiles/sweep_groups_r01_float32.json): the kernels are latency-bound, not DRAM-bound, so keeping
            // the scratch in L2 only pays where it spills completely -- real transforms of 2^19 and more (+15 %);
            // complex 2^16..2^18 lost 3-7 points to the extra barrier.  Default: groups for those sizes only.
            const int dflt_mb = (real && total >= ((size_t)1 << 19)) ? 56 : (1 << 20);
            const size_t budget = (size_t)env_int("SSFFT_FS_L2_MB", dflt_mb) << 20;
            int max_groups = 1;
            for (int kind = real ? 1 : 0; kind <= (real ? 2 : 0); ++kind) {
                const int tiles = fe.tiles[kind][0] > fe.tiles[kind][1] ? fe.tiles[kind][0] : fe.tiles[kind][1];
                int g = 1;
                while ((size_t)2 * (size_t)(clusters / g) * per > budget && g * cs < tiles && 2 * g <= clusters) g *= 2;
                if (env_int("SSFFT_FS_GROUP", 0) > 0) g = env_int("SSFFT_FS_GROUP", 0);
                if (g > clusters) g = clusters;
                pl->fs_group[kind] = g;
                if (clusters / g > max_groups) max_groups = clusters / g;
            }
            CU(cudaMalloc(&pl->d_scratch, (size_t)2 * max_groups * per));
            pl->chunk = (size_t)2 * max_groups;  // transforms the scratch holds (two-launch fallback path)
            CU(cudaMalloc(&pl->d_fs_ctr, (size_t)max_groups * sizeof(unsigned)));
        } else {
            pl->fs_id = -1;  // clusters of this size cannot be scheduled here
        }
    }
    if (pl->fs_id < 0) {
        pl->chunk = ((size_t)env_int("SSFFT_SCRATCH_MB", 32) << 20) / per;
        if (pl->chunk < 1) pl->chunk = 1;
        CU(cudaMalloc(&pl->d_scratch, pl->chunk * per));
    }
    pl->tiled = true;
    *ok = true;
    return SSFFT_OK;
}

// kind: 0 = C2C (dir by `inverse`), 1 = R2C, 2 = C2R.  in/out are user buffers of `batch` transforms.
template <typename T>
int exec_tiled(ssfft_plan *pl, int kind, const void *in, void *out, long long batch, int inverse, cudaStream_t s) {
    const int n1 = (int)pl->n1, n2 = (int)pl->n2;
    const long long total = (long long)n1 * n2;
    const long long user_stride = kind == 0 ? total : total / 2;  // cx elements per transform on the user side
    const long long sp = (long long)pl->scratch_per;
    cx<T> *scratch = (cx<T> *)pl->d_scratch;
    if (pl->fs_id >= 0) {
        FourStepParams<T> q;
        q.in = (const cx<T> *)in; q.out = (cx<T> *)out; q.scratch = scratch;
        q.tw_a = (const cx<T> *)pl->d_tile_tw_a; q.tw_b = (const cx<T> *)pl->d_tile_tw_b; q.tw4 = (const cx<T> *)pl->d_tw4;
        q.n1 = n1; q.n2 = n2; q.batch = batch; q.user_stride = user_stride; q.scratch_per = sp; q.inverse = inverse;
        q.ctb_log2 = pl->ctb_log2;
        q.discard = env_int("SSFFT_DISCARD", 1);
        q.cluster_size = pl->fs_csize;
        q.group_clusters = pl->fs_group[kind];
        q.group_ctr = (unsigned *)pl->d_fs_ctr;
        const int groups = pl->fs_clusters / q.group_clusters;
        if (q.group_clusters > 1) CU(cudaMemsetAsync(pl->d_fs_ctr, 0, (size_t)groups * sizeof(unsigned), s));
        int rc = fourstep_registry()[pl->fs_id].launch[kind](&q, groups * q.group_clusters, s);
        if (rc != 3) {
            ++g_launches;
            if (rc) return cuda_fail(cudaGetLastError(), "fourstep_cluster_kernel launch");
            return SSFFT_OK;
        }
        // rc == 3: no tensor map for this input (pointer not 16-byte aligned): two launches per chunk instead
    }
    if (pl->chunk < 1) return SSFFT_ERR_INVALID;
    for (long long b0 = 0; b0 < batch; b0 += (long long)pl->chunk) {
        const long long nb = (batch - b0 < (long long)pl->chunk) ? batch - b0 : (long long)pl->chunk;
        const cx<T> *cin = (const cx<T> *)in + b0 * user_stride;
        cx<T> *cout = (cx<T> *)out + b0 * user_stride;
        int rc;
        if (kind == 0) {
            rc = launch_tile_flavor<T>(pl->tile_a, TILE_A_C2C, cin, scratch, pl->d_tile_tw_a, pl->d_tw4, n1, n2, nb,
                                       user_stride, sp, inverse, pl->ctb_log2, s);
            if (rc) return rc;
            rc = launch_tile_flavor<T>(pl->tile_b, TILE_B_C2C, scratch, cout, pl->d_tile_tw_b, pl->d_tw4, n1, n2, nb, sp,
                                       user_stride, inverse, pl->ctb_log2, s);
        } else if (kind == 1) {
            rc = launch_tile_flavor<T>(pl->tile_a, TILE_A_R2C, cin, scratch, pl->d_tile_tw_a, pl->d_tw4, n1, n2, nb,
                                       user_stride, sp, 0, pl->ctb_log2, s);
            if (rc) return rc;
            rc = launch_tile_flavor<T>(pl->tile_b, TILE_B_R2C, scratch, cout, pl->d_tile_tw_b, pl->d_tw4, n1, n2, nb, sp,
                                       user_stride, 0, pl->ctb_log2, s);
        } else {
            rc = launch_tile_flavor<T>(pl->tile_b, TILE_B_C2R, cin, scratch, pl->d_tile_tw_b, pl->d_tw4, n1, n2, nb,
                                       user_stride, sp, 1, pl->ctb_log2, s);
            if (rc) return rc;
            rc = launch_tile_flavor<T>(pl->tile_a, TILE_A_C2R, scratch, cout, pl->d_tile_tw_a, pl->d_tw4, n1, n2, nb, sp,
                                       user_stride, 1, pl->ctb_log2, s);
        }
        if (rc) return rc;
    }
    CU(cudaGetLastError());
    return SSFFT_OK;
}

// Ticket-queue four-step (flat.cuh), complex transforms: tables, scratch slots and dependency counters.  *ok = false
// leaves the plan on its other paths (which stay set up as the fallback for inputs a tensor map cannot describe).
// Schedule of a launch: `span` = phases (transforms) whose tickets are held by the CTAs at any time, `delay` = phases
// between the column tiles of a transform and its row tiles, `slots` = scratch slots; delay + 1 <= slots is what
// correctness needs, the head-room is what keeps the dependency waits free.
struct FlatSchedule { int span, delay, slots; };
FlatSchedule flat_schedule(int ctas, int ring, int tickets_per_phase) {
    const long long window = (long long)ctas * (ring == 1 ? 3 : ring);  // executing + loading (+ in hand)
    FlatSchedule f;
    f.span = (int)((window + tickets_per_phase - 1) / tickets_per_phase);
    if (f.span < 1) f.span = 1;
    f.delay = (3 * f.span + 1) / 2;
    const int pct = env_int("SSFFT_FLAT_DELAY_PCT", 100);  // A/B measurements of the schedule
    if (pct > 0 && pct != 100) f.delay = (int)((long long)f.delay * pct / 100);
    f.slots = f.delay + 2 * f.span + 2;  // the previous user of a slot finishes about a span after its tickets went out
    return f;
}
template <typename T>
int setup_flat(ssfft_plan *pl, bool *ok) {
    *ok = false;
    const size_t n = pl->n;  // complex length (real plans: N / 2)
    bool real = pl->kind == SSFFT_REAL;
    if ((pl->kind != SSFFT_C2C && !real) || n == 0 || env_int("SSFFT_DISABLE_FLAT", 0)) return SSFFT_OK;
    if (real && env_int("SSFFT_DISABLE_FLAT_REAL", 0)) return SSFFT_OK;
    size_t n1 = 0, n2 = 0;
    if ((n & (n - 1)) == 0) {
        int lg = 0;
        while (((size_t)1 << lg) < n) ++lg;
        // below: a single-pass kernel exists and is as fast (fp32 2^14: 60.5 vs 58.5 %); fp64 has none at 2^14
        if (lg < env_int("SSFFT_FLAT_MIN_LOG2", sizeof(T) == 8 ? 14 : 15)) return SSFFT_OK;
        n1 = (size_t)1 << (lg / 2);
        n2 = n / n1;
        // another registered split of the same length, by (part of) its name -- or the measured default of the length
        const char *want = getenv("SSFFT_FLAT_NAME");
        for (const FlatEntry &e : flat_registry())
            if (e.prec == (sizeof(T) == 4 ? 0 : 1) && (size_t)e.n1 * (size_t)e.n2 == n && (!real || (e.launch_real[0] && e.launch_real[1])) &&
                ((want && *want && strstr(e.name, want)) || (!want && strstr(e.name, "_dflt")))) { n1 = (size_t)e.n1; n2 = (size_t)e.n2; break; }
    } else {
        // 3 * 2^k: the registered (power of two) x (3 * 2^j) pair of this length, if there is one
        // (by name with SSFFT_FLAT_NAME, else the split marked "_dflt", else the first one registered)
        const char *want = getenv("SSFFT_FLAT_NAME");
        int rank_best = 0;
        for (const FlatEntry &e : flat_registry()) {
            if (e.prec != (sizeof(T) == 4 ? 0 : 1) || (size_t)e.n1 * (size_t)e.n2 != n) continue;
            const bool real_fit = !real || (e.launch_real[0] && e.launch_real[1]);
            const int rank = (want && *want && strstr(e.name, want)) ? 3 : (!want && strstr(e.name, "_dflt") && real_fit) ? 2 : 1;
            if (rank > rank_best) { rank_best = rank; n1 = (size_t)e.n1; n2 = (size_t)e.n2; }
        }
        if (!n1) return SSFFT_OK;
    }
    // a real plan whose length has no RealFFT kernels registered (fp64 3 * 2^k, 9 * 2^k) still runs its complex core
    // on these kernels, between the stand-alone twiddle passes
    if (real && (find_flat<T>(n1, n2, 0) < 0 || find_flat<T>(n1, n2, 1) < 0)) real = false;
    const int id = find_flat<T>(n1, n2, real ? 0 : -1);
    if (id < 0) return SSFFT_OK;
    const FlatEntry &e = flat_registry()[id];
    // real plans: forward and inverse may use different entries of the same tiles (ring depth, CTAs per SM)
    int id_inv = real ? find_flat<T>(n1, n2, 1) : id;
    if (id_inv < 0) return SSFFT_OK;
    {
        const FlatEntry &ei = flat_registry()[id_inv];
        bool same = ei.cta == e.cta && ei.ctb == e.ctb && ei.na_passes == e.na_passes && ei.nb_passes == e.nb_passes;
        for (int i = 0; i < 3; ++i) same = same && ei.ra[i] == e.ra[i] && ei.rb[i] == e.rb[i];
        if (!same) id_inv = e.launch_real[1] ? id : -1;  // the tables below belong to one set of tiles
        if (id_inv < 0) return SSFFT_OK;
    }
    const FlatEntry &einv = flat_registry()[id_inv];
    int ctas = real ? e.max_ctas_real[0]() : e.max_ctas[0]();
    int ctas_inv = real ? einv.max_ctas_real[1]() : e.max_ctas[1]();
    if (!real) { if (ctas_inv < ctas) ctas = ctas_inv; ctas_inv = ctas; }
    if (ctas < 1 || ctas_inv < 1) return SSFFT_OK;
    std::vector<T> ga[2], gb[2], s4, twb;
    fill_flat_tables<T>(ga, gb, s4, n1, n2, e.ra, e.na_passes);
    fill_flat_row_twiddles<T>(twb, e.n2, e.rb, e.nb_passes, e.tile_b_tw);
    auto up = [&](void **d, const std::vector<T> &h) -> int {
        CU(cudaMalloc(d, h.size() * sizeof(T)));
        CU(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
        return SSFFT_OK;
    };
    int rc;
    for (int p = 0; p + 1 < e.na_passes; ++p)
        if ((rc = up(&pl->d_flat_ga[p], ga[p])) || (rc = up(&pl->d_flat_gb[p], gb[p]))) return rc;
    if ((rc = up(&pl->d_flat_s4, s4)) || (rc = up(&pl->d_flat_twb, twb))) return rc;
    if (real)
        for (int inv = 0; inv < 2; ++inv) {
            std::vector<T> ra, rb;
            fill_flat_real_tables<T>(ra, rb, n1, n2, inv != 0);
            if ((rc = up(&pl->d_flat_ra[inv], ra)) || (rc = up(&pl->d_flat_rb[inv], rb))) return rc;
        }
    pl->flat_real = real;
    const int pt = (int)(n2 / e.cta + n1 / e.ctb);
    int slots = flat_schedule(ctas, e.nstage, pt).slots;
    if (flat_schedule(ctas_inv, einv.nstage, pt).slots > slots) slots = flat_schedule(ctas_inv, einv.nstage, pt).slots;
    if (env_int("SSFFT_FLAT_SLOTS", 0) > slots) slots = env_int("SSFFT_FLAT_SLOTS", 0);
    CU(cudaMalloc(&pl->d_flat_scratch, (size_t)slots * n * sizeof(cx<T>)));
    pl->flat_cap = 1 << 16;
    CU(cudaMalloc(&pl->d_flat_ctrl, (size_t)(32 + 2 * pl->flat_cap) * sizeof(unsigned)));
    pl->flat_id = id; pl->flat_ctas = ctas; pl->flat_slots = slots;
    pl->flat_id_inv = id_inv; pl->flat_ctas_inv = ctas_inv;
    pl->flat_ctas_max = ctas; pl->flat_ctas_inv_max = ctas_inv;
    if (!pl->n1 && !real) { pl->n1 = n1; pl->n2 = n2; }
    *ok = true;
    return SSFFT_OK;
}

// returns SSFFT_OK, an error, or -1: this input has no tensor map (pointer not 16-byte aligned) -- use the other path
// kind: 0 complex (direction = inverse), 1 RealFFT forward, 2 RealFFT inverse; n = complex length in every case
template <typename T>
int exec_flat(ssfft_plan *pl, const void *in, void *out, long long batch, int inverse, cudaStream_t s, int kind = 0) {
    const FlatEntry &e = flat_registry()[kind == 2 ? pl->flat_id_inv : pl->flat_id];
    const int max_ctas = kind == 2 ? pl->flat_ctas_inv : pl->flat_ctas;
    const long long n = (long long)pl->n;
    if ((reinterpret_cast<uintptr_t>(in) & 15u) != 0) return -1;
    const int tiles1 = e.n2 / e.cta, tiles2 = e.n1 / e.ctb, pt = tiles1 + tiles2;
    for (long long b0 = 0; b0 < batch; b0 += pl->flat_cap) {
        const long long nb = batch - b0 < pl->flat_cap ? batch - b0 : pl->flat_cap;
        long long ctas = nb * (tiles1 > tiles2 ? tiles1 : tiles2);
        if (ctas > max_ctas) ctas = max_ctas;
        const FlatSchedule fs = flat_schedule((int)ctas, e.nstage, pt);
        long long delay = env_int("SSFFT_FLAT_DELAY", -1) >= 0 ? env_int("SSFFT_FLAT_DELAY", -1) : fs.delay;
        if (delay > nb) delay = nb;
        if (delay > pl->flat_slots - 1) delay = pl->flat_slots - 1;
        long long slots = env_int("SSFFT_FLAT_SLOTS", 0) > 0 ? env_int("SSFFT_FLAT_SLOTS", 0) : delay + 2 * fs.span + 2;
        if (slots > pl->flat_slots) slots = pl->flat_slots;
        if (slots < delay + 1) slots = delay + 1;
        FlatParams<T> q;
        q.in = (const cx<T> *)in + b0 * n; q.out = (cx<T> *)out + b0 * n; q.scratch = (cx<T> *)pl->d_flat_scratch;
        q.tw_b = (const cx<T> *)pl->d_flat_twb;
        for (int p = 0; p < 2; ++p) { q.ga[p] = (const cx<T> *)pl->d_flat_ga[p]; q.gb[p] = (const cx<T> *)pl->d_flat_gb[p]; }
        q.s4 = (const cx<T> *)pl->d_flat_s4; q.ctrl = (unsigned *)pl->d_flat_ctrl;
        q.ra = (const cx<T> *)pl->d_flat_ra[kind == 2 ? 1 : 0]; q.rb = (const cx<T> *)pl->d_flat_rb[kind == 2 ? 1 : 0];
        q.batch = nb; q.user_stride = n; q.scratch_per = n; q.cap = nb;
        q.nslots = (int)slots; q.delay = (int)delay; q.discard = env_int("SSFFT_DISCARD", 1);
        q.stats = nullptr;
#if SSFFT_FLAT_STATS
        static unsigned long long *d_stats = nullptr;  // measurement build: one buffer, calls are serialised below
        if (!d_stats) CU(cudaMalloc(&d_stats, (size_t)4096 * kFlatStats * sizeof(unsigned long long)));
        CU(cudaMemsetAsync(d_stats, 0, (size_t)4096 * kFlatStats * sizeof(unsigned long long), s));
        q.stats = d_stats;
#endif
        CU(cudaMemsetAsync(pl->d_flat_ctrl, 0, (size_t)(32 + 2 * nb) * sizeof(unsigned), s));
        const int rc = kind ? e.launch_real[kind - 1](&q, (int)ctas, s) : e.launch[inverse ? 1 : 0](&q, (int)ctas, s);
        if (rc == 3) return -1;
        ++g_launches;
        if (rc) return cuda_fail(cudaGetLastError(), "fourstep_flat_kernel launch");
#if SSFFT_FLAT_STATS
        if (env_int("SSFFT_FLAT_STATS_PRINT", 0)) {
            CU(cudaStreamSynchronize(s));
            std::vector<unsigned long long> h((size_t)ctas * kFlatStats);
            CU(cudaMemcpy(h.data(), d_stats, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            double sum[kFlatStats] = {0};
            for (long long c = 0; c < ctas; ++c)
                for (int i = 0; i < kFlatStats; ++i) sum[i] += (double)h[(size_t)c * kFlatStats + i];
            const double items = sum[2] > 0 ? sum[2] : 1, tk = sum[4] > 0 ? sum[4] : 1;
            fprintf(stderr,
                    "flat stats (%s, %lld CTAs, delay %lld, slots %lld): cycles/CTA %.0f | consumers wait for data %.1f %% | per item: "
                    "cycles %.0f, ticket atomic %.0f, signal %.0f, dep wait col %.0f row %.0f (open polls %.2f), slot wait %.0f | "
                    "copy latency when waited: col %.0f (%.0f %% of items) row %.0f (%.0f %%)\n",
                    e.name, ctas, delay, slots, sum[0] / ctas, 100.0 * sum[1] / (sum[0] > 0 ? sum[0] : 1), sum[0] / items,
                    sum[3] / tk, sum[9] / items, sum[5] / items * 2, sum[6] / items * 2, sum[7] / items, sum[8] / items,
                    sum[11] > 0 ? sum[10] / sum[11] : 0.0, 200.0 * sum[11] / items, sum[13] > 0 ? sum[12] / sum[13] : 0.0,
                    200.0 * sum[13] / items);
        }
#endif
    }
    return SSFFT_OK;
}

// Cluster-resident four-step (cluster.cuh): the transform lives in the shared memory of a thread-block cluster.
// Complex length pl->n; real plans need both the R2C and the C2R kernel.  *ok = false leaves the plan untouched.
template <typename T>
int setup_clustered(ssfft_plan *pl, bool *ok) {
    *ok = false;
    if (env_int("SSFFT_DISABLE_DSMEM", 0)) return SSFFT_OK;
    const bool real = pl->kind != SSFFT_C2C;
    const int k0 = real ? 1 : 0, k1 = real ? 2 : 0;
    int ids[3] = {-1, -1, -1}, clusters[3] = {0, 0, 0};
    for (int k = k0; k <= k1; ++k) {
        ids[k] = find_cluster<T>(pl->n, k);
        if (ids[k] < 0) return SSFFT_OK;
        clusters[k] = cluster_registry()[ids[k]].max_clusters[k]();
        if (clusters[k] < 1) return SSFFT_OK;  // clusters of this size cannot be scheduled on this device
    }
    for (int k = k0; k <= k1; ++k) {
        pl->cl_id[k] = ids[k];
        pl->cl_clusters[k] = clusters[k];
        if (k > k0 && ids[k] == ids[k0]) {  // same kernel geometry: share the tables
            pl->d_cl_twa[k] = pl->d_cl_twa[k0]; pl->d_cl_twb[k] = pl->d_cl_twb[k0]; pl->d_cl_tw4[k] = pl->d_cl_tw4[k0];
            continue;
        }
        const ClusterEntry &e = cluster_registry()[ids[k]];
        std::vector<T> h;
        fill_cluster_pass_twiddles<T>(h, e.n1, e.ra0);
        CU(cudaMalloc(&pl->d_cl_twa[k], h.size() * sizeof(T)));
        CU(cudaMemcpy(pl->d_cl_twa[k], h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
        fill_cluster_pass_twiddles<T>(h, e.n2, e.rb0);
        CU(cudaMalloc(&pl->d_cl_twb[k], h.size() * sizeof(T)));
        CU(cudaMemcpy(pl->d_cl_twb[k], h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
        fill_cluster_tw4<T>(h, e.n1, e.n2);
        CU(cudaMalloc(&pl->d_cl_tw4[k], h.size() * sizeof(T)));
        CU(cudaMemcpy(pl->d_cl_tw4[k], h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    }
    pl->clustered = true;
    *ok = true;
    return SSFFT_OK;
}

template <typename T>
int exec_clustered(ssfft_plan *pl, int kind, const void *in, void *out, long long batch, int inverse, cudaStream_t s) {
    ClusterParams<T> q;
    q.in = (const cx<T> *)in; q.out = (cx<T> *)out;
    q.tw_a = (const cx<T> *)pl->d_cl_twa[kind]; q.tw_b = (const cx<T> *)pl->d_cl_twb[kind];
    q.tw4 = (const cx<T> *)pl->d_cl_tw4[kind]; q.rtw = (const cx<T> *)pl->d_rtw;
    q.batch = batch; q.inverse = inverse;
    q.use_tma = env_int("SSFFT_CLUSTER_TMA", 1);
    int rc = cluster_registry()[pl->cl_id[kind]].launch[kind](&q, pl->cl_clusters[kind], s);
    ++g_launches;
    if (rc) return cuda_fail(cudaGetLastError(), "cluster_fft_kernel launch");
    return SSFFT_OK;
}

// Composite plan (composite.cuh): is there a small radix R such that M = N / R has one of the fast plans?
// fast = a registered single-pass kernel, or a power of two the four-step kernels cover.  Larger R first: the inner
// transform is the slower part, and shorter inner transforms are faster.
template <typename T>
bool composite_inner_is_fast(size_t m) {
    if (find_fused<T>(m, 0) >= 0) return true;
    if (m >= ((size_t)1 << 14) && m <= ((size_t)1 << 20) && (m & (m - 1)) == 0) return true;
    for (const FlatEntry &e : flat_registry())  // 3 * 2^k, 9 * 2^k on the ticket-queue kernels
        if (e.prec == (sizeof(T) == 4 ? 0 : 1) && (size_t)e.n1 * (size_t)e.n2 == m) return true;
    return false;
}
template <typename T>
int choose_composite_radix(size_t n) {
    if (env_int("SSFFT_DISABLE_COMPOSITE", 0)) return 0;
    const int forced = env_int("SSFFT_COMPOSITE_RADIX", 0);
    static const int radices[] = {16, 9, 8, 4, 3, 2};
    if (forced > 0) {
        for (int r : radices)
            if (r == forced && n % (size_t)r == 0) return r;
        return 0;
    }
    for (int r : radices)
        if (n % (size_t)r == 0 && composite_inner_is_fast<T>(n / (size_t)r)) return r;
    // two levels (3 * 2^21 = 3 x (16 x 2^17), ...): the inner plan is itself a composite
    for (int r : radices)
        if (n % (size_t)r == 0)
            for (int r2 : radices)
                if ((n / (size_t)r) % (size_t)r2 == 0 && composite_inner_is_fast<T>(n / (size_t)r / (size_t)r2)) return r;
    return 0;
}
template <typename T>
int setup_composite(ssfft_plan *pl, int r) {
    const size_t n = pl->n, m = n / (size_t)r;
    int rc = ssfft_plan_create(&pl->comp_inner, SSFFT_C2C, pl->prec, m, pl->device);
    if (rc) return rc;
    pl->comp_r = r;
    std::vector<T> h;
    fill_composite_twiddles<T>(h, (size_t)r, m);
    CU(cudaMalloc(&pl->d_comp_tw, h.size() * sizeof(T)));
    CU(cudaMemcpy(pl->d_comp_tw, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    // transforms per trip through the work buffer.  MEASURED (profiles/sweep_r02n_mb_f32.txt): trips small enough to keep
    // the buffer in L2 (16 / 40 / 96 MiB) are slower than streaming everything through HBM (49152: 16.7 / 21.2 / 23.1 vs
    // 24.5 % of the roofline; 2^21: 11.7 / 15.4 / 17.9 vs 19.6 %) -- three short launches per trip cost more than the L2 hits
    // save -- so the buffer is only bounded to keep its footprint reasonable
    pl->comp_chunk = ((size_t)env_int("SSFFT_COMPOSITE_MB", 1024) << 20) / (n * sizeof(cx<T>));
    if (pl->comp_chunk < 1) pl->comp_chunk = 1;
    CU(cudaMalloc(&pl->d_comp_work, pl->comp_chunk * n * sizeof(cx<T>)));
    return SSFFT_OK;
}
template <typename T, int R>
int launch_composite_passes(bool pre, const cx<T> *in, cx<T> *out, const cx<T> *tw, long long m, long long nb, int inverse, cudaStream_t s) {
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (pre) {
        const long long per_thread = sizeof(T) == 4 ? 2 : 1;  // columns per thread (16-byte accesses)
        const bool vec = m % per_thread == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(tw)) & 15u) == 0;
        long long blocks = ((vec ? m / per_thread : m) * nb + 255) / 256;
        if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
        if (vec) radix_pass_kernel<T, R><<<(unsigned)blocks, 256, 0, s>>>(in, out, tw, m, nb, inverse);
        else radix_pass_scalar_kernel<T, R><<<(unsigned)blocks, 256, 0, s>>>(in, out, tw, m, nb, inverse);
    } else {
        constexpr int kW = sizeof(T) == 8 ? 16 : 32;  // k2 per warp tile
        long long blocks = (((m + kW - 1) / kW) * nb + 7) / 8;
        if (blocks > (long long)sms * 6) blocks = (long long)sms * 6;
        const int vec_ok = (reinterpret_cast<uintptr_t>(out) & 15u) == 0 && (m * R) % 2 == 0;
        interleave_kernel<T, R><<<(unsigned)blocks, 256, 0, s>>>(in, out, m, nb, inverse, vec_ok);
    }
    ++g_launches;
    CU(cudaGetLastError());
    return SSFFT_OK;
}
template <typename T>
int composite_pass(int r, bool pre, const cx<T> *in, cx<T> *out, const cx<T> *tw, long long m, long long nb, int inverse, cudaStream_t s) {
    switch (r) {
        case 2: return launch_composite_passes<T, 2>(pre, in, out, tw, m, nb, inverse, s);
        case 3: return launch_composite_passes<T, 3>(pre, in, out, tw, m, nb, inverse, s);
        case 4: return launch_composite_passes<T, 4>(pre, in, out, tw, m, nb, inverse, s);
        case 8: return launch_composite_passes<T, 8>(pre, in, out, tw, m, nb, inverse, s);
        case 9: return launch_composite_passes<T, 9>(pre, in, out, tw, m, nb, inverse, s);
        case 16: return launch_composite_passes<T, 16>(pre, in, out, tw, m, nb, inverse, s);
    }
    return SSFFT_ERR_INVALID;
}
template <typename T>
int exec_composite(ssfft_plan *pl, const void *in, void *out, long long batch, int inverse, cudaStream_t s) {
    const long long n = (long long)pl->n, r = pl->comp_r, m = n / r;
    cx<T> *work = (cx<T> *)pl->d_comp_work;
    for (long long b0 = 0; b0 < batch; b0 += (long long)pl->comp_chunk) {
        const long long nb = batch - b0 < (long long)pl->comp_chunk ? batch - b0 : (long long)pl->comp_chunk;
        int rc = composite_pass<T>((int)r, true, (const cx<T> *)in + b0 * n, work, (const cx<T> *)pl->d_comp_tw, m, nb, inverse, s);
        if (rc) return rc;
        // the re / im swap of an inverse transform is done by the two passes around it: the inner transform runs forward
        rc = ssfft_exec_c2c(pl->comp_inner, work, work, (size_t)(nb * r), SSFFT_FORWARD, s);
        if (rc) return rc;
        rc = composite_pass<T>((int)r, false, work, (cx<T> *)out + b0 * n, nullptr, m, nb, inverse, s);
        if (rc) return rc;
    }
    return SSFFT_OK;
}

// Bluestein: inner plan of length M = 2^k >= 2n - 1, chirp table, spectrum of the wrapped conjugate chirp, work buffers.
template <typename T>
int setup_bluestein(ssfft_plan *pl) {
    const size_t n = pl->n, m = bluestein_length(n);
    int rc = ssfft_plan_create(&pl->bs_inner, SSFFT_C2C, pl->prec, m, pl->device);
    if (rc) return rc;
    pl->bs_m = m;
    std::vector<T> chirp, wrapped;
    fill_bluestein_tables<T>(chirp, wrapped, n, m);
    CU(cudaMalloc(&pl->d_bs_chirp, chirp.size() * sizeof(T)));
    CU(cudaMemcpy(pl->d_bs_chirp, chirp.data(), chirp.size() * sizeof(T), cudaMemcpyHostToDevice));
    void *d_wrapped = nullptr;
    CU(cudaMalloc(&d_wrapped, wrapped.size() * sizeof(T)));
    CU(cudaMemcpy(d_wrapped, wrapped.data(), wrapped.size() * sizeof(T), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&pl->d_bs_filter, wrapped.size() * sizeof(T)));
    rc = ssfft_exec_c2c(pl->bs_inner, d_wrapped, pl->d_bs_filter, 1, SSFFT_FORWARD, nullptr);
    const cudaError_t e = cudaDeviceSynchronize();
    cudaFree(d_wrapped);
    if (rc) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "bluestein filter");
    // transforms per pass: two work buffers of at most ~128 MiB each
    pl->bs_chunk = ((size_t)env_int("SSFFT_BLUESTEIN_MB", 128) << 20) / (m * sizeof(cx<T>));
    if (pl->bs_chunk < 1) pl->bs_chunk = 1;
    for (int i = 0; i < 2; ++i) CU(cudaMalloc(&pl->d_bs_work[i], pl->bs_chunk * m * sizeof(cx<T>)));
    return SSFFT_OK;
}

template <typename T>
int exec_bluestein(ssfft_plan *pl, const void *in, void *out, long long batch, int inverse, cudaStream_t s) {
    const long long n = (long long)pl->n, m = (long long)pl->bs_m;
    cx<T> *w0 = (cx<T> *)pl->d_bs_work[0], *w1 = (cx<T> *)pl->d_bs_work[1];
    const cx<T> *chirp = (const cx<T> *)pl->d_bs_chirp, *filt = (const cx<T> *)pl->d_bs_filter;
    for (long long b0 = 0; b0 < batch; b0 += (long long)pl->bs_chunk) {
        const long long nb = batch - b0 < (long long)pl->bs_chunk ? batch - b0 : (long long)pl->bs_chunk;
        const dim3 grid((unsigned)((m + 255) / 256 > 1024 ? 1024 : (m + 255) / 256), (unsigned)(nb > 65535 ? 65535 : nb));
        bluestein_pre_kernel<T><<<grid, 256, 0, s>>>((const cx<T> *)in + b0 * n, w0, chirp, n, m, nb, inverse);
        ++g_launches;
        int rc = ssfft_exec_c2c(pl->bs_inner, w0, w1, (size_t)nb, SSFFT_FORWARD, s);
        if (rc) return rc;
        bluestein_mul_kernel<T><<<grid, 256, 0, s>>>(w1, filt, m, nb);
        ++g_launches;
        rc = ssfft_exec_c2c(pl->bs_inner, w1, w0, (size_t)nb, SSFFT_INVERSE, s);
        if (rc) return rc;
        bluestein_post_kernel<T><<<grid, 256, 0, s>>>(w0, (cx<T> *)out + b0 * n, chirp, n, m, nb, (T)(1.0 / (double)m), inverse);
        ++g_launches;
        CU(cudaGetLastError());
    }
    return SSFFT_OK;
}

// Complex core: batch contiguous transforms of length pl->n, in -> out (in == out allowed).
template <typename T>
int exec_complex(ssfft_plan *pl, const void *in, void *out, long long batch, int inverse, cudaStream_t s) {
    const long long n = (long long)pl->n;
    if (batch <= 0 || n == 0) return SSFFT_OK;
    if (pl->tiny) {
        const int rc = launch_tiny<T>((size_t)n, in, out, batch, inverse, s);
        if (rc == 0) { ++g_launches; return SSFFT_OK; }
        if (rc == 2) return cuda_fail(cudaGetLastError(), "tiny_fft_kernel launch");
    }
    if (pl->flat_id >= 0 && !pl->flat_real) {  // first: the other paths of such a plan are its fallback
        const int rc = exec_flat<T>(pl, in, out, batch, inverse, s);
        if (rc >= 0) return rc;
    }
    if (pl->bs_inner) return exec_bluestein<T>(pl, in, out, batch, inverse, s);
    if (pl->comp_inner) return exec_composite<T>(pl, in, out, batch, inverse, s);
    if (pl->clustered && pl->kind == SSFFT_C2C) return exec_clustered<T>(pl, 0, in, out, batch, inverse, s);
    if (pl->tiled && pl->kind == SSFFT_C2C) return exec_tiled<T>(pl, 0, in, out, batch, inverse, s);
    if (!pl->four_step) {
        if (pl->fused.id >= 0)
            return launch_fused<T>(pl->fused.id, pl->fused.d_twiddles, in, out, batch, inverse, FUSED_C2C, nullptr, s,
                                   &g_launches);
        Layout L{n, 0, 1, 1, n, 0, 1, 1};
        return launch_generic<T>(pl, pl->direct, in, out, batch, L, inverse, false, 1, s);
    }
    // four-step: x[n1][n2] row-major.  (1) length-n1 FFTs down the n2 columns + twiddle W_n^(c*k1) into the
    // L2-resident scratch, (2) length-n2 FFTs along rows, stored transposed: X[k1 + n1*k2].
    const long long n1 = (long long)pl->n1, n2 = (long long)pl->n2;
    const char *src = (const char *)in;
    char *dst = (char *)out;
    for (long long b0 = 0; b0 < batch; b0 += (long long)pl->chunk) {
        const long long nb = (batch - b0 < (long long)pl->chunk) ? batch - b0 : (long long)pl->chunk;
        const void *cin = src + (size_t)b0 * (size_t)n * pl->elem;
        void *cout = dst + (size_t)b0 * (size_t)n * pl->elem;
        const Layout L1 = fourstep_col_layout((size_t)n, (size_t)n2);
        int rc = launch_generic<T>(pl, pl->col, cin, pl->d_scratch, nb * n2, L1, inverse, true, (int)n2, s);
        if (rc) return rc;
        const Layout L2 = fourstep_row_layout((size_t)n, (size_t)n1, (size_t)n2);
        rc = launch_generic<T>(pl, pl->row, pl->d_scratch, cout, nb * n1, L2, inverse, false, 1, s);
        if (rc) return rc;
    }
    return SSFFT_OK;
}

template <typename T>
int build_plan_typed(ssfft_plan *pl) {
    const int smem_max = max_optin_smem(pl->device);
    pl->elem = sizeof(cx<T>);
    const size_t n = pl->n;
    char buf[512];
    if (n == 0) { pl->desc = "empty"; return SSFFT_OK; }
    const size_t limit = generic_limit(sizeof(cx<T>), smem_max);
    bool tiled_ok = false, clustered_ok = false;
    if (pl->kind == SSFFT_C2C || pl->kind == SSFFT_REAL) {
        int rc = setup_clustered<T>(pl, &clustered_ok);
        if (rc) return rc;
    }
    if (clustered_ok) {
        // nothing else to build: one kernel does the whole transform
    } else if (pl->kind == SSFFT_C2C) {
        int rc = setup_tiled<T>(pl, n, false, &tiled_ok);
        if (rc) return rc;
    } else if (pl->kind == SSFFT_REAL) {
        int rc = setup_tiled<T>(pl, pl->n_real, true, &tiled_ok);
        if (rc) return rc;
    }
    // the ticket-queue kernels need a tensor map of the input (16-byte aligned pointer); every plan that uses them also
    // carries another path for the inputs that have none: the round-1 kernels where they exist, else the generic one
    bool flat_ok = false;
    const bool have_fallback = clustered_ok || tiled_ok;
    if (pl->kind == SSFFT_C2C || pl->kind == SSFFT_REAL) {
        int rc = setup_flat<T>(pl, &flat_ok);
        if (rc) return rc;
    }
    std::string flat_desc;
    if (flat_ok && pl->flat_real) {
        const FlatEntry &e = flat_registry()[pl->flat_id];
        snprintf(buf, sizeof(buf), "real N=%zu as complex N/2 = %d x %d ticket-queue four-step (forward %s, %d CTAs; inverse %s, %d CTAs): "
                 "pairs of samples, post- / pre-twiddle fused into the row / column tiles, one persistent launch, %d scratch slots = %.1f MiB in L2",
                 pl->n_real, e.n1, e.n2, e.name, pl->flat_ctas, flat_registry()[pl->flat_id_inv].name, pl->flat_ctas_inv, pl->flat_slots,
                 pl->flat_slots * (double)n * sizeof(cx<T>) / 1048576.0);
        flat_desc = buf;
    } else if (flat_ok) {
        const FlatEntry &e = flat_registry()[pl->flat_id];
        snprintf(buf, sizeof(buf), "complex N=%zu ticket-queue four-step n1=%d x n2=%d (%s): one persistent launch of %d CTAs "
                 "(%d consumer threads + TMA producer and signaller warps each, ring of %d%s), %d scratch slots = %.1f MiB in L2", n, e.n1, e.n2,
                 e.name, pl->flat_ctas, e.threads - kFlatHelpers, e.nstage, e.inplace ? " in place" : "", pl->flat_slots, pl->flat_slots * (double)n * sizeof(cx<T>) / 1048576.0);
        flat_desc = buf;
    }
    if (flat_ok && have_fallback) {
        pl->desc = flat_desc;
    } else if (clustered_ok) {
        const int k = pl->kind == SSFFT_C2C ? 0 : 1;
        const ClusterEntry &e = cluster_registry()[pl->cl_id[k]];
        if (pl->kind == SSFFT_C2C)
            snprintf(buf, sizeof(buf), "complex N=%zu cluster-resident four-step n1=%d x n2=%d (%s): %d-CTA clusters hold a "
                     "transform in distributed shared memory, one launch, %d co-resident clusters", n, e.n1, e.n2, e.name,
                     e.csize, pl->cl_clusters[k]);
        else
            snprintf(buf, sizeof(buf), "complex N=%zu cluster-resident four-step (%s forward, %s inverse): %d-CTA clusters, "
                     "RealFFT twiddles fused, one launch", n, e.name, cluster_registry()[pl->cl_id[2]].name, e.csize);
        pl->desc = buf;
    } else if (tiled_ok) {
        if (pl->fs_id >= 0)
            snprintf(buf, sizeof(buf), "%s N=%zu four-step n1=%zu x n2=%zu, one persistent launch %s, clusters of %d CTAs x %d, "
                     "%d cluster(s) per transform, L2-resident scratch %.1f MiB", pl->kind == SSFFT_C2C ? "complex" : "real",
                     pl->kind == SSFFT_C2C ? n : pl->n_real, pl->n1, pl->n2, fourstep_registry()[pl->fs_id].name,
                     pl->fs_csize, pl->fs_clusters, pl->fs_group[pl->kind == SSFFT_C2C ? 0 : 1],
                     2.0 * (pl->fs_clusters / pl->fs_group[pl->kind == SSFFT_C2C ? 0 : 1]) * pl->scratch_per * sizeof(cx<T>) / 1048576.0);
        else
            snprintf(buf, sizeof(buf), "%s N=%zu four-step tiles n1=%zu (%s) x n2=%zu (%s) chunk=%zu L2-resident scratch",
                     pl->kind == SSFFT_C2C ? "complex" : "real", pl->kind == SSFFT_C2C ? n : pl->n_real, pl->n1,
                     tile_registry()[pl->tile_a].name, pl->n2, tile_registry()[pl->tile_b].name, pl->chunk);
        pl->desc = buf;
    } else if (n <= limit || find_fused<T>(n, pl->kind != SSFFT_C2C ? 1 : 0) >= 0) {
        pl->four_step = false;
        int rc = SSFFT_OK;
        pl->fused.id = find_fused<T>(n, pl->kind != SSFFT_C2C ? 1 : 0);
        if (pl->fused.id >= 0) {
            rc = build_fused_twiddles<T>(pl->fused.id, &pl->fused.d_twiddles);
            if (rc) return rc;
            pl->direct.radix = choose_radices(n);  // description only: the fused kernel has its own radices
        } else {
            rc = build_generic_stage<T>(pl->direct, n, smem_max);
            if (rc) return rc;
        }
        std::string rs;
        for (int r : pl->direct.radix) { rs += (rs.empty() ? "" : "x"); rs += std::to_string(r); }
        // a handful of points: one thread per transform (the interpreter stays built as the path of the real wrappers)
        pl->tiny = pl->kind == SSFFT_C2C && pl->fused.id < 0 && tiny_supported(n) && !env_int("SSFFT_DISABLE_TINY", 0);
        if (pl->tiny)
            snprintf(buf, sizeof(buf), "n=%zu one thread per transform (register codelet, 256 transforms per CTA through shared memory)", n);
        else if (pl->fused.id >= 0)
            snprintf(buf, sizeof(buf), "n=%zu single-pass fused kernel %s", n, fused_name(pl->fused.id));
        else
            snprintf(buf, sizeof(buf), "n=%zu single-pass generic radices=%s tx=%d fpb=%d smem=%zu", n, rs.c_str(),
                     pl->direct.tx, pl->direct.fpb, pl->direct.smem_bytes);
        pl->desc = buf;
    } else if (const int comp_r = choose_composite_radix<T>(n)) {
        int rc = setup_composite<T>(pl, comp_r);
        if (rc) return rc;
        char inner[400] = "";
        ssfft_plan_describe(pl->comp_inner, inner, sizeof(inner));
        snprintf(buf, sizeof(buf), "n=%zu composite: radix-%d pass + %d x (n=%zu) + interleave, %zu transform(s) per trip through a %.0f MiB "
                 "work buffer; inner plan: %.300s", n, comp_r, comp_r, n / (size_t)comp_r, pl->comp_chunk,
                 pl->comp_chunk * (double)n * sizeof(cx<T>) / 1048576.0, inner);
        pl->desc = buf;
    } else {
        GenericFourStep fs;
        bool split = plan_generic_fourstep(fs, n, limit);
        if (split) {  // both legs must fit the pass interpreter (a prime factor above the limit fits neither leg)
            GenericStage probe_col, probe_row;
            split = plan_generic_stage(probe_col, fs.n1, sizeof(cx<T>), smem_max) && plan_generic_stage(probe_row, fs.n2, sizeof(cx<T>), smem_max);
        }
        if (!split || env_int("SSFFT_FORCE_BLUESTEIN", 0)) {
            int rc = setup_bluestein<T>(pl);
            if (rc) return rc;
            char inner[400] = "";
            ssfft_plan_describe(pl->bs_inner, inner, sizeof(inner));
            snprintf(buf, sizeof(buf), "n=%zu Bluestein (largest prime factor fits no on-chip path): convolution of length %zu, "
                     "%zu transform(s) per pass; inner plan: %.300s", n, pl->bs_m, pl->bs_chunk, inner);
            pl->desc = buf;
            goto real_wrappers;
        }
        const size_t n1 = fs.n1, n2 = fs.n2;
        pl->four_step = true;
        pl->n1 = n1; pl->n2 = n2;
        int rc = build_generic_stage<T>(pl->col, n1, smem_max);
        if (rc) return rc;
        rc = build_generic_stage<T>(pl->row, n2, smem_max);
        if (rc) return rc;
        // epilogue twiddle W_n^q, q < n, split q = hi * 2^shift + lo
        pl->ep_shift = fs.ep_shift;
        rc = upload_roots<T>(&pl->d_ep_lo, n, fs.lo_count, 1);
        if (rc) return rc;
        rc = upload_roots<T>(&pl->d_ep_hi, n, fs.hi_count, fs.lo_count);
        if (rc) return rc;
        // keep the intermediate of one chunk inside L2 (126 MB): 32 MiB of scratch
        size_t per = n * sizeof(cx<T>);
        pl->chunk = (32u << 20) / per;
        if (pl->chunk < 1) pl->chunk = 1;
        CU(cudaMalloc(&pl->d_scratch, pl->chunk * per));
        snprintf(buf, sizeof(buf), "n=%zu four-step n1=%zu n2=%zu chunk=%zu (generic x generic)", n, n1, n2, pl->chunk);
        pl->desc = buf;
    }
real_wrappers:
    if (pl->kind != SSFFT_C2C && !pl->tiled) {
        const bool modified = pl->kind == SSFFT_REAL_MODIFIED;
        // twiddlesMinusI, followed (modified plans) by modifiedRotations so the fused kernel finds both behind one
        // pointer: rot = rtw + (n/2 + 1), n = complex length
        const size_t ntw = pl->n_real / 4 + 1, nrot = modified ? pl->n_real / 2 : 0;
        std::vector<T> h(2 * (ntw + nrot));
        fill_real_twiddles<T>(h.data(), pl->n_real, modified);
        if (modified) fill_modified_rotations<T>(h.data() + 2 * ntw, pl->n_real);
        CU(cudaMalloc(&pl->d_rtw, h.size() * sizeof(T)));
        CU(cudaMemcpy(pl->d_rtw, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
        if (modified) {
            std::vector<T> r(2 * (pl->n_real / 2 ? pl->n_real / 2 : 1));
            fill_modified_rotations<T>(r.data(), pl->n_real);
            CU(cudaMalloc(&pl->d_rot, r.size() * sizeof(T)));
            CU(cudaMemcpy(pl->d_rot, r.data(), r.size() * sizeof(T), cudaMemcpyHostToDevice));
        }
        pl->desc = std::string(modified ? "modified-real " : "real ") + "N=" + std::to_string(pl->n_real) + " core: " + pl->desc;
    }
    if (flat_ok && !have_fallback) pl->desc = flat_desc + "; inputs without a tensor map: " + pl->desc;
    return SSFFT_OK;
}

inline unsigned blocks_for(long long items, int threads) { return (unsigned)((items + threads - 1) / threads); }

template <typename T>
int exec_r2c_typed(ssfft_plan *pl, const void *in, void *out, long long batch, cudaStream_t s) {
    const long long h = (long long)pl->n;
    if (h == 0 || batch <= 0) return SSFFT_OK;
    const bool modified = pl->kind == SSFFT_REAL_MODIFIED;
    if (pl->flat_id >= 0 && pl->flat_real) {
        const int rc = exec_flat<T>(pl, in, out, batch, 0, s, 1);
        if (rc >= 0) return rc;
    }
    if (pl->clustered) return exec_clustered<T>(pl, 1, in, out, batch, 0, s);
    if (pl->tiled) return exec_tiled<T>(pl, 1, in, out, batch, 0, s);
    if (!pl->four_step && pl->fused.id >= 0)
        return launch_fused<T>(pl->fused.id, pl->fused.d_twiddles, in, out, batch, 0, modified ? FUSED_R2C_MOD : FUSED_R2C,
                               pl->d_rtw, s, &g_launches);
    const void *src = in;  // N reals == h complex pairs (:449-455)
    if (modified) {
        rotate_kernel<T><<<blocks_for(h * batch, 256), 256, 0, s>>>((cx<T> *)out, (const cx<T> *)in,
                                                                     (const cx<T> *)pl->d_rot, h, h * batch, 0);
        ++g_launches;
        CU(cudaGetLastError());
        src = out;
    }
    int rc = exec_complex<T>(pl, src, out, batch, 0, s);
    if (rc) return rc;
    r2c_post_kernel<T><<<blocks_for((h / 2 + 1) * batch, 256), 256, 0, s>>>((cx<T> *)out, (const cx<T> *)pl->d_rtw, h,
                                                                             batch, modified ? 1 : 0);
    ++g_launches;
    CU(cudaGetLastError());
    return SSFFT_OK;
}

template <typename T>
int exec_c2r_typed(ssfft_plan *pl, const void *in, void *out, long long batch, cudaStream_t s) {
    const long long h = (long long)pl->n;
    if (h == 0 || batch <= 0) return SSFFT_OK;
    const bool modified = pl->kind == SSFFT_REAL_MODIFIED;
    if (pl->flat_id >= 0 && pl->flat_real) {
        const int rc = exec_flat<T>(pl, in, out, batch, 1, s, 2);
        if (rc >= 0) return rc;
    }
    if (pl->clustered) return exec_clustered<T>(pl, 2, in, out, batch, 1, s);
    if (pl->tiled) return exec_tiled<T>(pl, 2, in, out, batch, 1, s);
    if (!pl->four_step && pl->fused.id >= 0)
        return launch_fused<T>(pl->fused.id, pl->fused.d_twiddles, in, out, batch, 1, modified ? FUSED_C2R_MOD : FUSED_C2R,
                               pl->d_rtw, s, &g_launches);
    c2r_pre_kernel<T><<<blocks_for((h / 2 + 1) * batch, 256), 256, 0, s>>>(
        (const cx<T> *)in, (cx<T> *)out, (const cx<T> *)pl->d_rtw, h, batch, modified ? 1 : 0);
    ++g_launches;
    CU(cudaGetLastError());
    int rc = exec_complex<T>(pl, out, out, batch, 1, s);
    if (rc) return rc;
    if (modified) {
        rotate_kernel<T><<<blocks_for(h * batch, 256), 256, 0, s>>>((cx<T> *)out, (const cx<T> *)out,
                                                                     (const cx<T> *)pl->d_rot, h, h * batch, 1);
        ++g_launches;
        CU(cudaGetLastError());
    }
    return SSFFT_OK;
}

// =============================================================================================
// extended execution (ssfft_exec_*_ex): strided / overlapping layouts and fused multipliers
// =============================================================================================
int ex_workspace(void **ptr, size_t *have, size_t need) {
    if (*have >= need) return SSFFT_OK;
    if (*ptr) CU(cudaFree(*ptr));  // cudaFree waits for the device; calls on one plan are ordered (PlanExec), captures must pre-size
    *ptr = nullptr; *have = 0;
    CU(cudaMalloc(ptr, need));
    *have = need;
    return SSFFT_OK;
}

template <typename T>
int launch_ex_copy(bool real_side, const void *src, void *dst, long long len, long long batch, long long src_dist,
                   long long src_stride, long long dst_dist, long long dst_stride, const void *mul, int mul_kind,
                   long long mul_dist, bool packed, cudaStream_t s) {
    const long long total = len * batch;
    if (total <= 0) return SSFFT_OK;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (real_side)
        ex_copy_kernel<T, true><<<(unsigned)blocks, 256, 0, s>>>(src, dst, len, batch, src_dist, src_stride, dst_dist, dst_stride,
                                                                  (const T *)mul, mul_kind, mul_dist, 0);
    else
        ex_copy_kernel<T, false><<<(unsigned)blocks, 256, 0, s>>>(src, dst, len, batch, src_dist, src_stride, dst_dist, dst_stride,
                                                                   (const T *)mul, mul_kind, mul_dist, packed ? 1 : 0);
    ++g_launches;
    CU(cudaGetLastError());
    return SSFFT_OK;
}

template <typename T>
int exec_plain(ssfft_plan *pl, int op, const void *in, void *out, long long batch, int inverse, cudaStream_t s) {
    if (op == EX_C2C) return exec_complex<T>(pl, in, out, batch, inverse, s);
    if (op == EX_R2C) return exec_r2c_typed<T>(pl, in, out, batch, s);
    return exec_c2r_typed<T>(pl, in, out, batch, s);
}

template <typename T>
int exec_ex_typed(ssfft_plan *pl, int op, const void *in, void *out, long long batch, int inverse, const ExRequest &x,
                  cudaStream_t s) {
    if (x.in_plain && x.out_plain) return exec_plain<T>(pl, op, in, out, batch, inverse, s);
    const bool fused = !pl->four_step && !pl->tiled && !pl->clustered && pl->fused.id >= 0 && pl->kind != SSFFT_REAL_MODIFIED &&
                       fused_registry()[pl->fused.id].launch_ex && !env_int("SSFFT_EX_UNFUSED", 0);
    if (fused) {
        // ONE launch: the layouts and multipliers ride on the first load and the last store of the fused kernel
        FusedIo<T> f;
        f.in_dist = x.id; f.in_stride = x.is; f.out_dist = x.od; f.out_stride = x.os;
        f.pre = (const T *)x.pre; f.post = (const T *)x.post;
        f.pre_dist = x.pre_dist; f.post_dist = x.post_dist;
        f.pre_kind = x.pre_kind; f.post_kind = x.post_kind;
        const int mode = op == EX_C2C ? FUSED_C2C : op == EX_R2C ? FUSED_R2C : FUSED_C2R;
        fused_io_finalize(f, mode, in, out);  // vector accesses on the real side, sides that need no staging
        const FusedEntry &fe = fused_registry()[pl->fused.id];
        // opt-in experiment: column layouts through the size's column configuration (4 or 2 transforms per CTA)
        const bool cols = fe.launch_ex_cols && (x.id < x.is || x.od < x.os) && env_int("SSFFT_EX_COLCFG", 0);
        int rc = (cols ? fe.launch_ex_cols : fe.launch_ex)(pl->fused.d_twiddles, in, out, batch, op == EX_C2R ? 1 : inverse, mode,
                                                           pl->d_rtw, &f, s);
        ++g_launches;
        if (rc) return cuda_fail(cudaGetLastError(), "fused_fft_kernel<EX> launch");
        return SSFFT_OK;
    }
    // no fused kernel for this plan: gather pass -> plain transform -> scatter pass through the plan's workspaces
    const size_t bytes = (size_t)batch * pl->n * pl->elem;  // one contiguous batch (N reals == N/2 complex values)
    const void *src = in;
    void *dst = out;
    int rc;
    // the columns of a row-major [N][batch] matrix (stride = batch, dist = 1, no multiplier) move with the tiled
    // transpose (32 x 32 tiles through shared memory, both sides coalesced) instead of the element-wise copy, whose
    // strided side touches one element per 32-byte sector
    const int prec = sizeof(T) == 4 ? SSFFT_F32 : SSFFT_F64;
    const long long max_rows = 65535LL * 32 * kRowTiles;  // grid.y limit of the transpose launch
    const bool col_ok = batch > 1 && batch <= max_rows && (long long)pl->n <= max_rows;
    const bool col_in = col_ok && !x.in_real && !x.pre && x.id == 1 && x.is == batch;
    const bool col_out = col_ok && !x.out_real && !x.post && x.od == 1 && x.os == batch;
    if (!x.in_plain) {
        if ((rc = ex_workspace(&pl->d_ex_in, &pl->ex_in_bytes, bytes))) return rc;
        if (col_in)  // ws[c][r] = in[r][c]
            rc = ssfft_transpose_twiddle(in, pl->d_ex_in, 1, (size_t)x.in_len, (size_t)batch, 0, 0, 0, prec, s);
        else
            rc = launch_ex_copy<T>(x.in_real, in, pl->d_ex_in, x.in_len, batch, x.id, x.is, x.in_len, 1, x.pre, x.pre_kind,
                                   x.pre_dist, x.packed && !x.in_real, s);
        if (rc) return rc;
        src = pl->d_ex_in;
    }
    if (!x.out_plain) {
        if ((rc = ex_workspace(&pl->d_ex_out, &pl->ex_out_bytes, bytes))) return rc;
        dst = pl->d_ex_out;
    }
    if ((rc = exec_plain<T>(pl, op, src, dst, batch, inverse, s))) return rc;
    if (!x.out_plain) {
        if (col_out)  // out[k][c] = ws[c][k]
            rc = ssfft_transpose_twiddle(dst, out, 1, (size_t)batch, (size_t)x.out_len, 0, 0, 0, prec, s);
        else
            rc = launch_ex_copy<T>(x.out_real, dst, out, x.out_len, batch, x.out_len, 1, x.od, x.os, x.post, x.post_kind,
                                   x.post_dist, x.packed && !x.out_real, s);
    }
    return rc;
}

int exec_ex(ssfft_plan *pl, int op, const void *d_in, void *d_out, size_t batch, int inverse, const ssfft_io *io, void *stream) {
    ExRequest x;
    int rc = ex_validate(pl->kind, pl->n, pl->n_real, pl->elem, op, io, (long long)batch, d_in, d_out, x);
    if (rc) return rc;
    DeviceGuard guard(pl->device);
    PlanExec order(pl, (cudaStream_t)stream);
    return pl->prec == SSFFT_F32 ? exec_ex_typed<float>(pl, op, d_in, d_out, (long long)batch, inverse, x, (cudaStream_t)stream)
                                 : exec_ex_typed<double>(pl, op, d_in, d_out, (long long)batch, inverse, x, (cudaStream_t)stream);
}

void free_stage(GenericStage &st) {
    if (st.d_roots) cudaFree(st.d_roots);
    st.d_roots = nullptr;
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

size_t ssfft_size_minimum(size_t size) { return size_minimum(size); }
size_t ssfft_size_maximum(size_t size) { return size_maximum(size); }
size_t ssfft_real_size_minimum(size_t size) { return real_size_minimum(size); }
size_t ssfft_real_size_maximum(size_t size) { return real_size_maximum(size); }

int ssfft_device_count(int *count) {
    if (!count) return SSFFT_ERR_INVALID;
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) { *count = 0; cuda_fail(e, "cudaGetDeviceCount"); return SSFFT_ERR_NO_DEVICE; }
    *count = c;
    return c > 0 ? SSFFT_OK : SSFFT_ERR_NO_DEVICE;
}

int ssfft_plan_create(ssfft_plan **out, int kind, int precision, size_t n, int device) {
    if (!out) return SSFFT_ERR_INVALID;
    *out = nullptr;
    if (kind < SSFFT_C2C || kind > SSFFT_REAL_MODIFIED) return SSFFT_ERR_INVALID;
    if (precision != SSFFT_F32 && precision != SSFFT_F64) return SSFFT_ERR_INVALID;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return SSFFT_ERR_NO_DEVICE;  // no CPU fallback, by design
    }
    if (device < 0) CU(cudaGetDevice(&device));
    if (device >= count) return SSFFT_ERR_INVALID;
    DeviceGuard guard(device);
    if (!guard.ok) return SSFFT_ERR_CUDA;
    ssfft_plan *pl = new ssfft_plan();
    pl->kind = kind; pl->prec = precision; pl->device = device;
    pl->n_user = n;
    if (kind == SSFFT_C2C) { pl->n = n; pl->n_real = 0; }
    else { pl->n_real = (n / 2) * 2; pl->n = n / 2; }  // RealFFT::setSize :416-435 (odd sizes truncate)
    int rc = precision == SSFFT_F32 ? build_plan_typed<float>(pl) : build_plan_typed<double>(pl);
    if (rc) { ssfft_plan_destroy(pl); return rc; }
    *out = pl;
    return SSFFT_OK;
}

int ssfft_plan_destroy(ssfft_plan *pl) {
    if (!pl) return SSFFT_OK;
    DeviceGuard guard(pl->device);
    free_stage(pl->direct); free_stage(pl->col); free_stage(pl->row);
    void *ptrs[] = {pl->fused.d_twiddles, pl->fused_col.d_twiddles, pl->fused_row.d_twiddles, pl->d_ep_lo, pl->d_ep_hi,
                    pl->d_scratch, pl->d_rtw, pl->d_rot, pl->d_tile_tw_a, pl->d_tile_tw_b,
                    pl->d_tw4, pl->d_fs_ctr, pl->d_ex_in, pl->d_ex_out, pl->d_flat_ga[0], pl->d_flat_ga[1], pl->d_flat_gb[0], pl->d_flat_gb[1], pl->d_flat_s4,
                    pl->d_flat_twb, pl->d_flat_scratch, pl->d_flat_ctrl, pl->d_flat_ra[0], pl->d_flat_ra[1], pl->d_flat_rb[0],
                    pl->d_flat_rb[1]};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    for (int k = 0; k < 3; ++k) {
        const bool shared = k > 0 && pl->d_cl_tw4[k] && (pl->d_cl_tw4[k] == pl->d_cl_tw4[k - 1] || (k == 2 && pl->d_cl_tw4[2] == pl->d_cl_tw4[0]));
        if (shared) continue;
        if (pl->d_cl_twa[k]) cudaFree(pl->d_cl_twa[k]);
        if (pl->d_cl_twb[k]) cudaFree(pl->d_cl_twb[k]);
        if (pl->d_cl_tw4[k]) cudaFree(pl->d_cl_tw4[k]);
    }
    if (pl->bs_inner) ssfft_plan_destroy(pl->bs_inner);
    if (pl->comp_inner) ssfft_plan_destroy(pl->comp_inner);
    for (void *p : {pl->d_comp_tw, pl->d_comp_work})
        if (p) cudaFree(p);
    for (void *p : {pl->d_bs_chirp, pl->d_bs_filter, pl->d_bs_work[0], pl->d_bs_work[1]})
        if (p) cudaFree(p);
    for (int i = 0; i < ssfft_plan::kRing; ++i) {
        if (pl->d_ring_in[i]) cudaFree(pl->d_ring_in[i]);
        if (pl->d_ring_out[i]) cudaFree(pl->d_ring_out[i]);
        if (pl->ev_h2d[i]) cudaEventDestroy(pl->ev_h2d[i]);
        if (pl->ev_comp[i]) cudaEventDestroy(pl->ev_comp[i]);
        if (pl->ev_d2h[i]) cudaEventDestroy(pl->ev_d2h[i]);
    }
    if (pl->h_zc_in) cudaFreeHost(pl->h_zc_in);
    if (pl->h_zc_out) cudaFreeHost(pl->h_zc_out);
    if (pl->st_h2d) cudaStreamDestroy(pl->st_h2d);
    if (pl->st_comp) cudaStreamDestroy(pl->st_comp);
    if (pl->st_d2h) cudaStreamDestroy(pl->st_d2h);
    if (pl->exec_done) cudaEventDestroy(pl->exec_done);
    delete pl;
    return SSFFT_OK;
}

size_t ssfft_plan_size(const ssfft_plan *pl) {
    if (!pl) return 0;
    return pl->kind == SSFFT_C2C ? pl->n : pl->n_real;
}

int ssfft_plan_limit_ctas(ssfft_plan *pl, int ctas_per_sm) {
    if (!pl || ctas_per_sm < 0) return SSFFT_ERR_INVALID;
    if (pl->flat_id < 0) return SSFFT_OK;
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pl->device) != cudaSuccess || sms < 1) { cudaGetLastError(); return SSFFT_ERR_CUDA; }
    auto cap = [&](int most) { return ctas_per_sm == 0 || ctas_per_sm * sms > most ? most : ctas_per_sm * sms; };
    pl->flat_ctas = cap(pl->flat_ctas_max);
    pl->flat_ctas_inv = cap(pl->flat_ctas_inv_max);
    return SSFFT_OK;
}

int ssfft_plan_describe(const ssfft_plan *pl, char *buf, size_t buflen) {
    if (!pl || !buf || !buflen) return SSFFT_ERR_INVALID;
    snprintf(buf, buflen, "%s %s", pl->prec == SSFFT_F32 ? "f32" : "f64", pl->desc.c_str());
    return SSFFT_OK;
}

int ssfft_exec_c2c(ssfft_plan *pl, const void *d_in, void *d_out, size_t batch, int direction, void *stream) {
    if (!pl || pl->kind != SSFFT_C2C) return SSFFT_ERR_INVALID;
    if (direction != SSFFT_FORWARD && direction != SSFFT_INVERSE) return SSFFT_ERR_INVALID;
    if (batch == 0 || pl->n == 0) return SSFFT_OK;
    if (!d_in || !d_out) return SSFFT_ERR_INVALID;
    // the kernels move whole complex values (real data as pairs): a pointer that is not aligned to one would fault
    if ((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_out)) % pl->elem) return SSFFT_ERR_INVALID;
    DeviceGuard guard(pl->device);
    const int inv = direction == SSFFT_INVERSE;
    if (plan_is_stateful(pl)) {
        PlanExec order(pl, (cudaStream_t)stream);
        return pl->prec == SSFFT_F32 ? exec_complex<float>(pl, d_in, d_out, (long long)batch, inv, (cudaStream_t)stream)
                                     : exec_complex<double>(pl, d_in, d_out, (long long)batch, inv, (cudaStream_t)stream);
    }
    return pl->prec == SSFFT_F32 ? exec_complex<float>(pl, d_in, d_out, (long long)batch, inv, (cudaStream_t)stream)
                                 : exec_complex<double>(pl, d_in, d_out, (long long)batch, inv, (cudaStream_t)stream);
}

int ssfft_exec_r2c(ssfft_plan *pl, const void *d_in, void *d_out, size_t batch, void *stream) {
    if (!pl || pl->kind == SSFFT_C2C) return SSFFT_ERR_INVALID;
    if (batch == 0 || pl->n == 0) return SSFFT_OK;
    if (!d_in || !d_out) return SSFFT_ERR_INVALID;
    // the kernels move whole complex values (real data as pairs): a pointer that is not aligned to one would fault
    if ((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_out)) % pl->elem) return SSFFT_ERR_INVALID;
    DeviceGuard guard(pl->device);
    PlanExec order(pl, (cudaStream_t)stream);
    return pl->prec == SSFFT_F32 ? exec_r2c_typed<float>(pl, d_in, d_out, (long long)batch, (cudaStream_t)stream)
                                 : exec_r2c_typed<double>(pl, d_in, d_out, (long long)batch, (cudaStream_t)stream);
}

int ssfft_exec_c2r(ssfft_plan *pl, const void *d_in, void *d_out, size_t batch, void *stream) {
    if (!pl || pl->kind == SSFFT_C2C) return SSFFT_ERR_INVALID;
    if (batch == 0 || pl->n == 0) return SSFFT_OK;
    if (!d_in || !d_out) return SSFFT_ERR_INVALID;
    // the kernels move whole complex values (real data as pairs): a pointer that is not aligned to one would fault
    if ((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_out)) % pl->elem) return SSFFT_ERR_INVALID;
    DeviceGuard guard(pl->device);
    PlanExec order(pl, (cudaStream_t)stream);
    return pl->prec == SSFFT_F32 ? exec_c2r_typed<float>(pl, d_in, d_out, (long long)batch, (cudaStream_t)stream)
                                 : exec_c2r_typed<double>(pl, d_in, d_out, (long long)batch, (cudaStream_t)stream);
}

int ssfft_exec_c2c_ex(ssfft_plan *pl, const void *d_in, void *d_out, size_t batch, int direction, const ssfft_io *io,
                      void *stream) {
    if (!io) return ssfft_exec_c2c(pl, d_in, d_out, batch, direction, stream);
    if (!pl || pl->kind != SSFFT_C2C) return SSFFT_ERR_INVALID;
    if (direction != SSFFT_FORWARD && direction != SSFFT_INVERSE) return SSFFT_ERR_INVALID;
    if (batch == 0 || pl->n == 0) return SSFFT_OK;
    if (!d_in || !d_out) return SSFFT_ERR_INVALID;
    return exec_ex(pl, EX_C2C, d_in, d_out, batch, direction == SSFFT_INVERSE, io, stream);
}

int ssfft_exec_r2c_ex(ssfft_plan *pl, const void *d_in, void *d_out, size_t batch, const ssfft_io *io, void *stream) {
    if (!io) return ssfft_exec_r2c(pl, d_in, d_out, batch, stream);
    if (!pl || pl->kind == SSFFT_C2C) return SSFFT_ERR_INVALID;
    if (batch == 0 || pl->n == 0) return SSFFT_OK;
    if (!d_in || !d_out) return SSFFT_ERR_INVALID;
    return exec_ex(pl, EX_R2C, d_in, d_out, batch, 0, io, stream);
}

int ssfft_exec_c2r_ex(ssfft_plan *pl, const void *d_in, void *d_out, size_t batch, const ssfft_io *io, void *stream) {
    if (!io) return ssfft_exec_c2r(pl, d_in, d_out, batch, stream);
    if (!pl || pl->kind == SSFFT_C2C) return SSFFT_ERR_INVALID;
    if (batch == 0 || pl->n == 0) return SSFFT_OK;
    if (!d_in || !d_out) return SSFFT_ERR_INVALID;
    return exec_ex(pl, EX_C2R, d_in, d_out, batch, 1, io, stream);
}

// One slice of a host call on the plan's own path (no PlanExec: the caller holds the order)
static int exec_host_kernels(ssfft_plan *pl, int op, const void *di, void *dout, size_t nb, cudaStream_t cs) {
    const long long b = (long long)nb;
    if (pl->prec == SSFFT_F32) {
        if (op == 0) return exec_complex<float>(pl, di, dout, b, 0, cs);
        if (op == 1) return exec_complex<float>(pl, di, dout, b, 1, cs);
        if (op == 2) return exec_r2c_typed<float>(pl, di, dout, b, cs);
        return exec_c2r_typed<float>(pl, di, dout, b, cs);
    }
    if (op == 0) return exec_complex<double>(pl, di, dout, b, 0, cs);
    if (op == 1) return exec_complex<double>(pl, di, dout, b, 1, cs);
    if (op == 2) return exec_r2c_typed<double>(pl, di, dout, b, cs);
    return exec_c2r_typed<double>(pl, di, dout, b, cs);
}

static int host_path_setup(ssfft_plan *pl) {
    if (pl->st_comp) return SSFFT_OK;
    CU(cudaStreamCreateWithFlags(&pl->st_h2d, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&pl->st_comp, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&pl->st_d2h, cudaStreamNonBlocking));
    for (int i = 0; i < ssfft_plan::kRing; ++i) {
        CU(cudaEventCreateWithFlags(&pl->ev_h2d[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&pl->ev_comp[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&pl->ev_d2h[i], cudaEventDisableTiming));
    }
    return SSFFT_OK;
}

// Host buffers in, host buffers out (the call a reference user makes: FFT<V>::fft(container, container)).
//  * tiny calls (both sides <= 256 KiB): the data goes through pinned, mapped staging buffers owned by the plan and the
//    kernels read / write them over PCIe themselves -- no cudaMemcpy, one launch, one stream synchronisation;
//  * everything else: slices of <= 64 MiB through a ring of three device buffers per side; H2D copies, kernels and D2H
//    copies run on three streams of the plan linked by events, so slice i+1 uploads and slice i-1 downloads while
//    slice i transforms (PCIe is full duplex) and ALL kernels of the call run on ONE stream in slice order -- the plan's
//    scratch is never used by two slices at once.  Device memory per call is bounded by the ring (6 slices), not by
//    the batch.  Pageable host memory works but serialises the copies; pinned memory (cudaHostAlloc / ssfft_host_alloc)
//    gives the overlap.
int ssfft_exec_host(ssfft_plan *pl, int op, const void *h_in, void *h_out, size_t batch) {
    if (!pl || op < 0 || op > 3) return SSFFT_ERR_INVALID;
    if ((op <= 1) != (pl->kind == SSFFT_C2C)) return SSFFT_ERR_INVALID;
    if (batch == 0 || pl->n == 0) return SSFFT_OK;
    if (!h_in || !h_out) return SSFFT_ERR_INVALID;
    DeviceGuard guard(pl->device);
    std::lock_guard<std::mutex> lock(pl->exec_mu);
    int rc = host_path_setup(pl);
    if (rc) return rc;
    // bytes per transform on each side: C2C n cx both; real: N reals == n cx on both sides as well
    const size_t per = pl->n * pl->elem, bytes = batch * per;
    if (pl->exec_any && pl->exec_done) CU(cudaStreamWaitEvent(pl->st_comp, pl->exec_done, 0));  // behind earlier device calls

    const size_t zc_limit = (size_t)env_int("SSFFT_HOST_ZEROCOPY_KB", 256) << 10;
    if (bytes <= zc_limit) {
        if (pl->zc_bytes < bytes) {
            if (pl->h_zc_in) cudaFreeHost(pl->h_zc_in);
            if (pl->h_zc_out) cudaFreeHost(pl->h_zc_out);
            pl->h_zc_in = pl->h_zc_out = nullptr; pl->zc_bytes = 0;
            const size_t cap = bytes < 65536 ? 65536 : bytes;
            CU(cudaHostAlloc(&pl->h_zc_in, cap, cudaHostAllocMapped));
            CU(cudaHostAlloc(&pl->h_zc_out, cap, cudaHostAllocMapped));
            pl->zc_bytes = cap;
        }
        memcpy(pl->h_zc_in, h_in, bytes);
        void *dz_in = nullptr, *dz_out = nullptr;
        CU(cudaHostGetDevicePointer(&dz_in, pl->h_zc_in, 0));
        CU(cudaHostGetDevicePointer(&dz_out, pl->h_zc_out, 0));
        rc = exec_host_kernels(pl, op, dz_in, dz_out, batch, pl->st_comp);
        if (rc) return rc;
        CU(cudaStreamSynchronize(pl->st_comp));
        memcpy(h_out, pl->h_zc_out, bytes);
        return SSFFT_OK;
    }

    // slices: large enough for PCIe to run at full rate, small enough that fill / drain of the pipeline stay short
    size_t slice_tr = ((size_t)env_int("SSFFT_HOST_SLICE_MB", 64) << 20) / per;
    if (slice_tr < 1) slice_tr = 1;
    const size_t min_slices = 8;  // a call of a few hundred MiB still overlaps its copies
    if (batch / slice_tr < min_slices && batch >= min_slices) slice_tr = (batch + min_slices - 1) / min_slices;
    if (slice_tr > batch) slice_tr = batch;
    const size_t slice_bytes = slice_tr * per;
    if (pl->ring_bytes < slice_bytes) {
        for (int i = 0; i < ssfft_plan::kRing; ++i) {
            if (pl->d_ring_in[i]) cudaFree(pl->d_ring_in[i]);
            if (pl->d_ring_out[i]) cudaFree(pl->d_ring_out[i]);
            pl->d_ring_in[i] = pl->d_ring_out[i] = nullptr;
        }
        pl->ring_bytes = 0;
        for (int i = 0; i < ssfft_plan::kRing; ++i) {
            CU(cudaMalloc(&pl->d_ring_in[i], slice_bytes));
            CU(cudaMalloc(&pl->d_ring_out[i], slice_bytes));
        }
        pl->ring_bytes = slice_bytes;
    }
    const size_t slices = (batch + slice_tr - 1) / slice_tr;
    cudaError_t e = cudaSuccess;
    for (size_t i = 0; i < slices && rc == SSFFT_OK && e == cudaSuccess; ++i) {
        const int k = (int)(i % ssfft_plan::kRing);
        const size_t b0 = i * slice_tr, nb = batch - b0 < slice_tr ? batch - b0 : slice_tr;
        const char *hi = (const char *)h_in + b0 * per;
        char *ho = (char *)h_out + b0 * per;
        // upload: the kernels of slice i - kRing have read this input buffer
        if (i >= (size_t)ssfft_plan::kRing) e = cudaStreamWaitEvent(pl->st_h2d, pl->ev_comp[k], 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(pl->d_ring_in[k], hi, nb * per, cudaMemcpyHostToDevice, pl->st_h2d);
        if (e == cudaSuccess) e = cudaEventRecord(pl->ev_h2d[k], pl->st_h2d);
        // transform: input uploaded, output buffer downloaded (slice i - kRing)
        if (e == cudaSuccess) e = cudaStreamWaitEvent(pl->st_comp, pl->ev_h2d[k], 0);
        if (e == cudaSuccess && i >= (size_t)ssfft_plan::kRing) e = cudaStreamWaitEvent(pl->st_comp, pl->ev_d2h[k], 0);
        if (e != cudaSuccess) break;
        rc = exec_host_kernels(pl, op, pl->d_ring_in[k], pl->d_ring_out[k], nb, pl->st_comp);
        if (rc) break;
        e = cudaEventRecord(pl->ev_comp[k], pl->st_comp);
        // download
        if (e == cudaSuccess) e = cudaStreamWaitEvent(pl->st_d2h, pl->ev_comp[k], 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(ho, pl->d_ring_out[k], nb * per, cudaMemcpyDeviceToHost, pl->st_d2h);
        if (e == cudaSuccess) e = cudaEventRecord(pl->ev_d2h[k], pl->st_d2h);
    }
    const cudaError_t e1 = cudaStreamSynchronize(pl->st_h2d), e2 = cudaStreamSynchronize(pl->st_comp),
                      e3 = cudaStreamSynchronize(pl->st_d2h);
    if (rc) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "host path enqueue");
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)
        return cuda_fail(e1 != cudaSuccess ? e1 : e2 != cudaSuccess ? e2 : e3, "cudaStreamSynchronize");
    return SSFFT_OK;
}

// Pinned (page-locked) host memory for the host path: copies from / to it overlap with the kernels.
int ssfft_host_alloc(void **h_ptr, size_t bytes) {
    if (!h_ptr) return SSFFT_ERR_INVALID;
    *h_ptr = nullptr;
    if (cudaHostAlloc(h_ptr, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return SSFFT_ERR_ALLOC; }
    return SSFFT_OK;
}
int ssfft_host_free(void *h_ptr) {
    if (h_ptr) CU(cudaFreeHost(h_ptr));
    return SSFFT_OK;
}

int ssfft_malloc(void **d_ptr, size_t bytes) {
    if (!d_ptr) return SSFFT_ERR_INVALID;
    *d_ptr = nullptr;
    if (cudaMalloc(d_ptr, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return SSFFT_ERR_ALLOC; }
    return SSFFT_OK;
}
int ssfft_free(void *d_ptr) {
    if (d_ptr) CU(cudaFree(d_ptr));
    return SSFFT_OK;
}
int ssfft_memcpy_h2d(void *d_dst, const void *h_src, size_t bytes, void *stream) {
    if (!bytes) return SSFFT_OK;
    CU(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return SSFFT_OK;
}
int ssfft_memcpy_d2h(void *h_dst, const void *d_src, size_t bytes, void *stream) {
    if (!bytes) return SSFFT_OK;
    CU(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return SSFFT_OK;
}
int ssfft_stream_synchronize(void *stream) {
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    return SSFFT_OK;
}
int ssfft_fill_uniform(void *d_dst, size_t count, int precision, uint64_t seed, uint64_t first_idx, void *stream) {
    if (!count) return SSFFT_OK;
    if (!d_dst) return SSFFT_ERR_INVALID;
    unsigned blocks = (unsigned)((count + 255) / 256 > 148 * 32 ? 148 * 32 : (count + 255) / 256);
    if (precision == SSFFT_F32)
        fill_uniform_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((float *)d_dst, count, seed, first_idx);
    else if (precision == SSFFT_F64)
        fill_uniform_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>((double *)d_dst, count, seed, first_idx);
    else
        return SSFFT_ERR_INVALID;
    ++g_launches;
    CU(cudaGetLastError());
    return SSFFT_OK;
}

int ssfft_transpose_twiddle(const void *d_in, void *d_out, size_t batch, size_t rows, size_t cols, size_t row0,
                            uint64_t n_total, int inverse, int precision, void *stream) {
    if (!batch || !rows || !cols) return SSFFT_OK;
    if (!d_in || !d_out || d_in == d_out) return SSFFT_ERR_INVALID;
    const size_t row_blocks = (rows + 32 * kRowTiles - 1) / (32 * kRowTiles);
    if (batch > 65535 || row_blocks > 65535) return SSFFT_ERR_INVALID;
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)row_blocks, (unsigned)batch), block(32, 8);
    if (precision == SSFFT_F32)
        transpose_twiddle_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>(
            (const cx<float> *)d_in, (cx<float> *)d_out, (long long)rows, (long long)cols, (long long)row0, n_total, inverse);
    else if (precision == SSFFT_F64)
        transpose_twiddle_kernel<double><<<grid, block, 0, (cudaStream_t)stream>>>(
            (const cx<double> *)d_in, (cx<double> *)d_out, (long long)rows, (long long)cols, (long long)row0, n_total, inverse);
    else
        return SSFFT_ERR_INVALID;
    ++g_launches;
    CU(cudaGetLastError());
    return SSFFT_OK;
}

int ssfft_permute102(const void *d_in, void *d_out, size_t A, size_t B, size_t run, int precision, void *stream) {
    if (!A || !B || !run) return SSFFT_OK;
    if (!d_in || !d_out || d_in == d_out) return SSFFT_ERR_INVALID;
    // grid.x covers one run (256 threads x 16 B), grid.y strides over the A*B runs
    const size_t per_run_blocks = (run / 2 + 255) / 256 ? (run / 2 + 255) / 256 : 1;
    unsigned gx = (unsigned)(per_run_blocks > 64 ? 64 : per_run_blocks);
    size_t gy = A * B;
    if (gy > 65535) gy = 65535;
    dim3 grid(gx, (unsigned)gy);
    if (precision == SSFFT_F32)
        permute102_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const cx<float> *)d_in, (cx<float> *)d_out,
                                                                          (long long)A, (long long)B, (long long)run);
    else if (precision == SSFFT_F64)
        permute102_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>((const cx<double> *)d_in, (cx<double> *)d_out,
                                                                           (long long)A, (long long)B, (long long)run);
    else
        return SSFFT_ERR_INVALID;
    ++g_launches;
    CU(cudaGetLastError());
    return SSFFT_OK;
}

int ssfft_ipc_export(void *d_ptr, void *handle64) {
    if (!d_ptr || !handle64) return SSFFT_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CU(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t *>(handle64), d_ptr));
    return SSFFT_OK;
}
int ssfft_ipc_import(const void *handle64, void **d_ptr) {
    if (!d_ptr || !handle64) return SSFFT_ERR_INVALID;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    CU(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return SSFFT_OK;
}
int ssfft_ipc_close(void *d_ptr) {
    if (d_ptr) CU(cudaIpcCloseMemHandle(d_ptr));
    return SSFFT_OK;
}
int ssfft_memcpy_d2d(void *d_dst, const void *d_src, size_t bytes, void *stream) {
    if (!bytes) return SSFFT_OK;
    CU(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return SSFFT_OK;
}

int ssfft_exchange_transpose(const void *d_src, void *const *d_dst_ptrs, int world, size_t rows, size_t cols,
                             size_t dst_pitch, size_t dst_col0, size_t row0, uint64_t n_total, int inverse, int precision,
                             void *stream) {
    if (!rows || !cols) return SSFFT_OK;
    if (!d_src || !d_dst_ptrs || world < 1 || world > kMaxPeers || cols % (size_t)world) return SSFFT_ERR_INVALID;
    // experiment, off by default: the TMA-driven kernel (exchange_tma.cuh), a few CTAs per SM instead of 64 warps per SM.
    // MEASURED SLOWER (profiles/bench_dist_tma_r02y.txt, 2^30 over 2 GPUs): 14.9 / 16.2 / 14.8 ms with 2 / 1 / 4 CTAs per SM
    // against 13.9 ms, and chunked phases still gain nothing beside it (14.5 ms at best)
    if (env_int("SSFFT_EXCHANGE_TMA", 0) && (precision == SSFFT_F32 || precision == SSFFT_F64)) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int max_ctas = sms * env_int("SSFFT_EXCHANGE_CTAS_PER_SM", 2);
        const int rc = precision == SSFFT_F32
                           ? launch_exchange_tma<float>(d_src, d_dst_ptrs, world, rows, cols, dst_pitch, dst_col0, row0, n_total, inverse, max_ctas, (cudaStream_t)stream)
                           : launch_exchange_tma<double>(d_src, d_dst_ptrs, world, rows, cols, dst_pitch, dst_col0, row0, n_total, inverse, max_ctas, (cudaStream_t)stream);
        if (rc == 0) { ++g_launches; return SSFFT_OK; }
        if (rc == 2) return cuda_fail(cudaGetLastError(), "exchange_tma_kernel launch");
    }
    const size_t row_blocks = (rows + 32 * kRowTiles - 1) / (32 * kRowTiles);
    if (row_blocks > 65535) return SSFFT_ERR_INVALID;
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)row_blocks), block(32, 8);
    const long long blk = (long long)(cols / (size_t)world);
    if (precision == SSFFT_F32) {
        PeerPtrs<float> pp;
        for (int i = 0; i < kMaxPeers; ++i) pp.p[i] = i < world ? (cx<float> *)d_dst_ptrs[i] : nullptr;
        exchange_transpose_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>(
            (const cx<float> *)d_src, pp, (long long)rows, (long long)cols, blk, (long long)dst_pitch, (long long)dst_col0,
            (long long)row0, n_total, inverse);
    } else if (precision == SSFFT_F64) {
        PeerPtrs<double> pp;
        for (int i = 0; i < kMaxPeers; ++i) pp.p[i] = i < world ? (cx<double> *)d_dst_ptrs[i] : nullptr;
        exchange_transpose_kernel<double><<<grid, block, 0, (cudaStream_t)stream>>>(
            (const cx<double> *)d_src, pp, (long long)rows, (long long)cols, blk, (long long)dst_pitch, (long long)dst_col0,
            (long long)row0, n_total, inverse);
    } else {
        return SSFFT_ERR_INVALID;
    }
    ++g_launches;
    CU(cudaGetLastError());
    return SSFFT_OK;
}

const char *ssfft_error_string(int status) {
    switch (status) {
        case SSFFT_OK: return "ok";
        case SSFFT_ERR_INVALID: return "invalid argument";
        case SSFFT_ERR_CUDA: return "CUDA runtime error";
        case SSFFT_ERR_UNSUPPORTED: return "unsupported size (prime factor too large for the on-chip paths)";
        case SSFFT_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
        case SSFFT_ERR_ALLOC: return "device allocation failed";
        default: return "unknown status";
    }
}
const char *ssfft_last_cuda_error(void) { return g_cuda_err; }
uint64_t ssfft_launch_count(void) { return g_launches.load(); }
/* 0.2: extended execution (ssfft_exec_*_ex) */
const char *ssfft_version(void) { return "ssfft-b200 0.2 (sm_100a)"; }

}  // extern "C"
