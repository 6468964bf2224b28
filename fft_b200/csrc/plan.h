// plan.h -- internal plan object behind the opaque ssfft_plan handle of include/ssfft.h.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <mutex>
#include <string>
#include <vector>

namespace ssfft {

// One single-launch transform of length n through the generic pass interpreter (generic.cuh).
struct GenericStage {
    int n = 0;
    std::vector<int> radix, prod;
    void *d_roots = nullptr;  // W_n^k, k < n  (cx<T>)
    int smem_stride = 0;
    int tx = 0, fpb = 0;
    int stage_input = 0;
    size_t smem_bytes = 0;
};

// Which specialised kernel (fused.cuh) serves a length, if any.
struct FusedChoice {
    int id = -1;              // index into the fused-kernel registry, -1 = none
    void *d_twiddles = nullptr;
};

}  // namespace ssfft

struct ssfft_plan {
    int kind = 0, prec = 0, device = 0;
    size_t n_user = 0;  // as given
    size_t n_real = 0;  // real plans: 2*(n_user/2)
    size_t n = 0;       // complex transform length (n_user, or n_real/2)
    size_t elem = 0;    // sizeof(cx<T>)

    // complex core
    bool four_step = false;
    size_t n1 = 0, n2 = 0;
    ssfft::GenericStage direct, col, row;  // direct: !four_step;  col (n1, strided) / row (n2) otherwise
    ssfft::FusedChoice fused;              // contiguous single-pass kernel for n, when registered
    ssfft::FusedChoice fused_col, fused_row;
    // fast four-step through the tile kernels (tiled.cuh): ids into tile_registry(), -1 = not used
    bool tiled = false;
    int tile_a = -1, tile_b = -1;          // length-n1 (column) and length-n2 (row) kernels
    void *d_tile_tw_a = nullptr, *d_tile_tw_b = nullptr;
    void *d_tw4 = nullptr;                 // W_N^(n2*k1) laid out [k1][n2]
    size_t scratch_per = 0;                // scratch elements (cx) per transform
    int ctb_log2 = 0;                      // tile-major block height (lanes of the row-stage kernel)
    int fs_id = -1;                        // cluster kernel (both stages in one launch), -1 = two launches per chunk
    int fs_clusters = 0;                   // co-resident clusters of the persistent launch
    int fs_csize = 4;                      // CTAs per cluster
    int fs_group[3] = {1, 1, 1};           // clusters that share one transform, per kind (keeps the scratch in L2)
    void *d_fs_ctr = nullptr;              // group barrier counters
    void *d_ep_lo = nullptr, *d_ep_hi = nullptr;
    int ep_shift = 0;
    void *d_scratch = nullptr;  // four-step intermediate, chunk transforms
    size_t chunk = 0;

    // ticket-queue four-step (flat.cuh): registry id, -1 = not used.  Complex transforms only for now.
    int flat_id = -1;                      // registry entry (complex plans; real plans: the forward transform)
    int flat_id_inv = -1;                  // real plans: entry of the inverse transform (same tiles, maybe another ring)
    int flat_ctas = 0;                     // co-resident CTAs of the persistent launch (real plans: forward)
    int flat_ctas_inv = 0;
    int flat_ctas_max = 0, flat_ctas_inv_max = 0;  // what the device can hold (ssfft_plan_limit_ctas lowers flat_ctas below it)
    int flat_slots = 0;                    // scratch slots (transforms) allocated
    void *d_flat_ga[2] = {nullptr, nullptr}, *d_flat_gb[2] = {nullptr, nullptr}, *d_flat_s4 = nullptr, *d_flat_twb = nullptr;
    void *d_flat_scratch = nullptr, *d_flat_ctrl = nullptr;
    long long flat_cap = 0;                // transforms per launch the dependency counters cover
    bool flat_real = false;                // real plan: the entry's RealFFT kernels (forward / inverse) are used
    void *d_flat_ra[2] = {nullptr, nullptr}, *d_flat_rb[2] = {nullptr, nullptr};  // post- / pre-twiddle factors [fwd, inv]

    // Bluestein (bluestein.cuh): lengths whose largest prime factor fits no on-chip path run as a convolution through
    bool tiny = false;                      // N <= 24 complex: one thread per transform (tiny.cuh)
    // composite plan (composite.cuh): N = comp_r * M, one radix pass + the plan of length M + an interleave pass
    ssfft_plan *comp_inner = nullptr;
    int comp_r = 0;
    size_t comp_chunk = 0;                  // transforms per trip through the work buffer
    void *d_comp_tw = nullptr, *d_comp_work = nullptr;
    // an inner power-of-two plan of length bs_m
    ssfft_plan *bs_inner = nullptr;
    size_t bs_m = 0, bs_chunk = 0;          // convolution length, transforms per pass through the work buffers
    void *d_bs_chirp = nullptr, *d_bs_filter = nullptr, *d_bs_work[2] = {nullptr, nullptr};

    // cluster-resident four-step (cluster.cuh): registry id per kind (C2C / R2C / C2R), -1 = not used
    bool clustered = false;
    int cl_id[3] = {-1, -1, -1};
    int cl_clusters[3] = {0, 0, 0};
    void *d_cl_twa[3] = {nullptr, nullptr, nullptr}, *d_cl_twb[3] = {nullptr, nullptr, nullptr},
         *d_cl_tw4[3] = {nullptr, nullptr, nullptr};  // tables per kind (kinds that share an entry share the pointers)

    // real wrappers
    void *d_rtw = nullptr;  // twiddlesMinusI
    void *d_rot = nullptr;  // modifiedRotations

    // Execution order.  A plan is NOT re-entrant: scratch, dependency counters, workspaces and the host-path staging
    // are per plan.  Calls on one plan are serialised on the host by exec_mu, and a call issued on another stream than
    // the previous one first waits (on the device) for that one's event, so concurrent use from several streams or
    // threads is slow but correct.  Independent work wants one plan per stream (copies of the C++ objects get one).
    std::mutex exec_mu;
    cudaEvent_t exec_done = nullptr;
    cudaStream_t exec_last = nullptr;
    bool exec_any = false;

    // host-pointer path (ssfft_exec_host): three streams (H2D, kernels, D2H) linked by events over a ring of
    // slice-sized device buffers, all owned by the plan; tiny calls go through pinned mapped host buffers instead
    static constexpr int kRing = 3;
    cudaStream_t st_h2d = nullptr, st_comp = nullptr, st_d2h = nullptr;
    cudaEvent_t ev_h2d[kRing] = {nullptr, nullptr, nullptr}, ev_comp[kRing] = {nullptr, nullptr, nullptr},
                ev_d2h[kRing] = {nullptr, nullptr, nullptr};
    void *d_ring_in[kRing] = {nullptr, nullptr, nullptr}, *d_ring_out[kRing] = {nullptr, nullptr, nullptr};
    size_t ring_bytes = 0;
    void *h_zc_in = nullptr, *h_zc_out = nullptr;  // pinned, mapped (zero-copy) staging for tiny transfers
    size_t zc_bytes = 0;

    // extended execution (ssfft_exec_*_ex) without a fused kernel: gather / scatter workspaces (grow-only)
    void *d_ex_in = nullptr, *d_ex_out = nullptr;
    size_t ex_in_bytes = 0, ex_out_bytes = 0;

    std::string desc;
};
