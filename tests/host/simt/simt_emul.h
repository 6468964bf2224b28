// simt_emul.h -- run a CUDA kernel's source on the CPU, one fiber per CUDA thread (TEST INFRASTRUCTURE).
//
//   simt::launch(grid, block, [&] { kernel<Cfg>(args...); });                        // CTAs one after the other
//   simt::launch(grid, block, body, simt::Launch{cluster_size, resident_ctas});      // clusters / persistent grids
//
// Every thread is a ucontext fiber executing the kernel body.  `resident_ctas` CTAs are alive at the same time (one
// cluster by default; the whole grid for persistent kernels whose CTAs wait for each other).  __syncthreads() parks a
// fiber until every fiber of its CTA that has not returned is parked too; the cluster barrier does the same over the
// CTAs of a cluster; spin-waits on global memory yield (simt::spin_yield).  A state in which nobody can run any more is
// reported as a deadlock instead of hanging.  The resume order can be reversed (State::reverse_order): the tests run
// every case under both orders, so a result that depends on which thread happens to run first fails.  Each live CTA has its own dynamic shared memory (SSFFT_DYNAMIC_SMEM in
// the kernels resolves to simt::dynamic_smem()), poisoned with NaN before the CTA starts.
//
// TMA bulk copy + mbarrier (fused.cuh, PF = 1 / 2; one mbarrier per CTA is all the kernels use) are emulated at the
// two extremes the hardware allows:
//   late_copy = false: the copy is performed AT ISSUE TIME (the earliest it could land), so a kernel that lets the
//                      copy overwrite shared memory some thread still has to read fails;
//   late_copy = true : the copy is performed when the first thread WAITS for it (the latest it could land), so a
//                      kernel that writes to the destination between issue and wait gets its data overwritten.
#pragma once
#include <ucontext.h>

#include <cstdio>
#include <functional>
#include <memory>
#include <vector>

#include "cuda_runtime.h"

uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

namespace simt {

constexpr size_t kSmemBytes = 232448;  // 227 KiB: the opt-in maximum of one sm_100 CTA

struct Launch {
    unsigned cluster = 1;    // CTAs per cluster (cluster_ctarank / cluster barrier)
    unsigned resident = 0;   // CTAs alive at the same time; 0 = one cluster.  Must be a multiple of `cluster`.
    size_t stack_bytes = 256 * 1024;
};

struct Fiber {
    ucontext_t ctx;
    std::unique_ptr<unsigned char[]> stack;
    size_t stack_bytes = 0;
    uint3 tid, bid;
    unsigned cta = 0;  // index among the resident CTAs
    enum { RUNNABLE, AT_BARRIER, AT_CLUSTER_BARRIER, DONE } st = RUNNABLE;
};

struct Mbar {  // emulated mbarrier of one CTA
    unsigned phase = 0;
    long long pending_tx = 0;
    bool armed = false;
    struct Pending { void *dst; const void *src; unsigned bytes; };
    std::vector<Pending> pending;
};

struct State {
    ucontext_t sched;
    std::vector<Fiber> fibers;
    int current = -1;
    const std::function<void()> *body = nullptr;
    std::vector<unsigned char *> smem;  // per resident CTA, 128-byte aligned
    std::vector<std::unique_ptr<unsigned char[]>> smem_store;
    std::vector<Mbar> mbar;             // per resident CTA
    unsigned cluster = 1;
    unsigned long long barriers = 0, cluster_barriers = 0, bulk_copies = 0, bulk_bytes = 0;
    bool progress = false;
    bool late_copy = false;
    // Fibers are resumed in index order, or in reverse when set: code that is only correct because "thread 0 runs
    // first" (a missing barrier, a flag read before it is written) gives different results under the two orders.
    bool reverse_order = false;
};
inline State &state() { static State s; return s; }
inline Fiber &self() { State &s = state(); return s.fibers[s.current]; }

inline unsigned char *dynamic_smem() { State &s = state(); return s.smem[self().cta]; }
inline unsigned cluster_ctarank() { return self().cta % state().cluster; }
inline unsigned cluster_nctarank() { return state().cluster; }

inline void fiber_entry() {
    State &s = state();
    (*s.body)();
    s.fibers[s.current].st = Fiber::DONE;
    s.progress = true;
    swapcontext(&s.fibers[s.current].ctx, &s.sched);
}

inline void yield_to_scheduler() {
    State &s = state();
    const int me = s.current;
    swapcontext(&s.fibers[me].ctx, &s.sched);
    threadIdx = s.fibers[me].tid;
    blockIdx = s.fibers[me].bid;
}
inline void spin_yield() { yield_to_scheduler(); }  // a polling loop lets everybody else run before it looks again
inline void cluster_barrier() {
    State &s = state();
    ++s.cluster_barriers;
    s.progress = true;
    self().st = Fiber::AT_CLUSTER_BARRIER;
    yield_to_scheduler();
}

// Runs `body` for every thread of every CTA.  Returns false on deadlock.
inline bool launch(dim3 grid, dim3 block, const std::function<void()> &body, Launch cfg = Launch()) {
    State &s = state();
    gridDim = grid;
    blockDim = block;
    const unsigned nthreads = block.x * block.y * block.z, nctas = grid.x * grid.y * grid.z;
    const unsigned cluster = cfg.cluster ? cfg.cluster : 1;
    unsigned resident = cfg.resident ? cfg.resident : cluster;
    if (resident > nctas) resident = nctas;
    if (nctas % cluster || resident % cluster) { fprintf(stderr, "simt: grid / resident CTAs must be multiples of the cluster size\n"); return false; }
    s.cluster = cluster;
    s.body = &body;
    s.fibers.resize((size_t)resident * nthreads);
    for (auto &f : s.fibers)
        if (f.stack_bytes != cfg.stack_bytes) { f.stack.reset(new unsigned char[cfg.stack_bytes]); f.stack_bytes = cfg.stack_bytes; }
    while (s.smem.size() < resident) {
        s.smem_store.emplace_back(new unsigned char[kSmemBytes + 128]);
        unsigned char *p = s.smem_store.back().get();
        s.smem.push_back(p + (128 - reinterpret_cast<uintptr_t>(p) % 128) % 128);
    }
    s.mbar.assign(resident, Mbar());
    for (unsigned first = 0; first < nctas; first += resident) {
        const unsigned live_ctas = nctas - first < resident ? nctas - first : resident;
        for (unsigned c = 0; c < live_ctas; ++c) {
            const unsigned b = first + c;
            memset(s.smem[c], 0xff, kSmemBytes);  // NaN pattern: reads of unwritten shared memory show up
            s.mbar[c] = Mbar();
            for (unsigned i = 0; i < nthreads; ++i) {
                Fiber &f = s.fibers[(size_t)c * nthreads + i];
                f.st = Fiber::RUNNABLE;
                f.cta = c;
                f.tid = uint3{i % block.x, (i / block.x) % block.y, i / (block.x * block.y)};
                f.bid = uint3{b % grid.x, (b / grid.x) % grid.y, b / (grid.x * grid.y)};
                getcontext(&f.ctx);
                f.ctx.uc_stack.ss_sp = f.stack.get();
                f.ctx.uc_stack.ss_size = f.stack_bytes;
                f.ctx.uc_link = nullptr;
                makecontext(&f.ctx, (void (*)())fiber_entry, 0);
            }
        }
        const size_t nf = (size_t)live_ctas * nthreads;
        for (;;) {
            s.progress = false;
            for (size_t k = 0; k < nf; ++k) {
                const size_t i = s.reverse_order ? nf - 1 - k : k;
                Fiber &f = s.fibers[i];
                if (f.st != Fiber::RUNNABLE) continue;
                s.current = (int)i;
                threadIdx = f.tid;
                blockIdx = f.bid;
                swapcontext(&s.sched, &f.ctx);
            }
            // release the barriers every participant has reached
            size_t runnable = 0, parked = 0;
            for (unsigned c = 0; c < live_ctas; ++c) {
                size_t at = 0, run = 0;
                for (unsigned i = 0; i < nthreads; ++i) {
                    const Fiber &f = s.fibers[(size_t)c * nthreads + i];
                    at += f.st == Fiber::AT_BARRIER;
                    run += f.st == Fiber::RUNNABLE || f.st == Fiber::AT_CLUSTER_BARRIER;
                }
                if (at && !run) {
                    for (unsigned i = 0; i < nthreads; ++i) {
                        Fiber &f = s.fibers[(size_t)c * nthreads + i];
                        if (f.st == Fiber::AT_BARRIER) f.st = Fiber::RUNNABLE;
                    }
                    s.progress = true;
                }
            }
            for (unsigned c0 = 0; c0 < live_ctas; c0 += cluster) {
                size_t at = 0, other = 0;
                for (size_t i = (size_t)c0 * nthreads; i < (size_t)(c0 + cluster) * nthreads; ++i) {
                    at += s.fibers[i].st == Fiber::AT_CLUSTER_BARRIER;
                    other += s.fibers[i].st == Fiber::RUNNABLE || s.fibers[i].st == Fiber::AT_BARRIER;
                }
                if (at && !other) {
                    for (size_t i = (size_t)c0 * nthreads; i < (size_t)(c0 + cluster) * nthreads; ++i)
                        if (s.fibers[i].st == Fiber::AT_CLUSTER_BARRIER) s.fibers[i].st = Fiber::RUNNABLE;
                    s.progress = true;
                }
            }
            for (size_t i = 0; i < nf; ++i) {
                runnable += s.fibers[i].st == Fiber::RUNNABLE;
                parked += s.fibers[i].st == Fiber::AT_BARRIER || s.fibers[i].st == Fiber::AT_CLUSTER_BARRIER;
            }
            if (!runnable && !parked) break;  // every thread returned
            if (!s.progress) {
                fprintf(stderr, "simt: deadlock (CTAs %u..%u): %zu thread(s) spinning, %zu parked at a barrier nobody else reaches\n",
                        first, first + live_ctas - 1, runnable, parked);
                return false;
            }
        }
    }
    return true;
}

// ---- hooks called by the kernels' PTX wrappers when SSFFT_EMUL is defined
inline Mbar &my_mbar() { return state().mbar[self().cta]; }
inline void mbar_complete_if_ready(Mbar &m) {
    if (m.armed && m.pending_tx == 0) { m.armed = false; ++m.phase; state().progress = true; }
}
inline void mbar_init() { my_mbar() = Mbar(); }
inline void mbar_expect_tx(unsigned bytes) {  // arrive (count 1) + expect-tx
    Mbar &m = my_mbar();
    m.pending_tx += bytes;
    m.armed = true;
    mbar_complete_if_ready(m);
}
inline void bulk_g2s(void *dst, const void *src, unsigned bytes) {
    State &s = state();
    Mbar &m = my_mbar();
    if (bytes % 16 || ((uintptr_t)dst & 15) || ((uintptr_t)src & 15)) {
        fprintf(stderr, "simt: cp.async.bulk needs 16-byte aligned addresses and sizes (dst %p src %p bytes %u)\n", dst, src, bytes);
        abort();
    }
    ++s.bulk_copies; s.bulk_bytes += bytes;
    if (s.late_copy) { m.pending.push_back({dst, src, bytes}); return; }
    memcpy(dst, src, bytes);
    m.pending_tx -= bytes;
    mbar_complete_if_ready(m);
}
inline void mbar_wait(unsigned parity) {
    State &s = state();
    Mbar &m = my_mbar();
    for (auto &c : m.pending) { memcpy(c.dst, c.src, c.bytes); m.pending_tx -= c.bytes; }
    m.pending.clear();
    mbar_complete_if_ready(m);
    while ((m.phase & 1u) == parity) yield_to_scheduler();  // the phase with this parity has not completed yet
    s.progress = true;
}

}  // namespace simt

inline void __syncthreads() {
    simt::State &s = simt::state();
    ++s.barriers;
    s.progress = true;
    simt::self().st = simt::Fiber::AT_BARRIER;
    simt::yield_to_scheduler();
}
