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
#include <map>
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
    enum { RUNNABLE, AT_BARRIER, DONE } st = RUNNABLE;
    unsigned cl_waits = 0;  // cluster-barrier phases this thread has waited for
    unsigned bar_id = 0, bar_count = 0;  // AT_BARRIER: 0 = __syncthreads (every live thread), else bar.sync id, count
};

struct Mbar {  // emulated mbarrier (arrival count + transaction bytes); a CTA's mbarriers are told apart by address
    unsigned phase = 0;
    unsigned expected = 1, arrived = 0;  // arrivals a phase needs / has seen
    long long pending_tx = 0;  // may go negative: transactions can complete before the expect-tx of their phase
    bool armed = false;        // every expected arrival of the current phase is in
    struct Pending { void *dst; const void *src; unsigned bytes; unsigned char value[16]; };  // src == nullptr: `value`
    std::vector<Pending> pending;
};
struct ClusterBar {  // split-phase hardware cluster barrier: arrive ... wait
    size_t arrived = 0;
    unsigned phase = 0;
};

struct State {
    ucontext_t sched;
    std::vector<Fiber> fibers;
    int current = -1;
    const std::function<void()> *body = nullptr;
    std::vector<unsigned char *> smem;  // per resident CTA, 128-byte aligned
    std::vector<std::unique_ptr<unsigned char[]>> smem_store;
    std::vector<std::map<const void *, Mbar>> mbar;  // per resident CTA
    std::vector<ClusterBar> cbar;       // per resident cluster
    unsigned cluster = 1, nthreads = 0;
    unsigned long long barriers = 0, cluster_barriers = 0, bulk_copies = 0, bulk_bytes = 0;
    bool progress = false;
    bool late_copy = false;
    // Fibers are resumed in index order, or in reverse when set: code that is only correct because "thread 0 runs
    // first" (a missing barrier, a flag read before it is written) gives different results under the two orders.
    bool reverse_order = false;
    // Skew between CTAs.  0: every runnable fiber of every resident CTA runs once per round (CTAs advance together).
    // 1 / 2: the lowest / highest-numbered CTA that can make progress runs alone until it blocks on another CTA, so
    // one CTA races as far ahead of its peers as the synchronisation allows -- a missing cross-CTA wait (a peer's
    // store landing in a buffer its owner is still reading) shows up.
    int cta_priority = 0;
};
inline State &state() { static State s; return s; }
inline Fiber &self() { State &s = state(); return s.fibers[s.current]; }

inline unsigned char *dynamic_smem() { State &s = state(); return s.smem[self().cta]; }
inline unsigned cluster_ctarank() { return self().cta % state().cluster; }
inline unsigned cluster_nctarank() { return state().cluster; }
inline unsigned cta_of_rank(unsigned rank) { State &s = state(); return self().cta / s.cluster * s.cluster + rank; }
// the dynamic shared memory of CTA `rank` of my cluster (what mapa.shared::cluster addresses)
inline unsigned char *peer_smem(unsigned rank) { return state().smem[cta_of_rank(rank)]; }

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
// barrier.cluster.arrive / barrier.cluster.wait: the phase completes when every thread of the cluster that has not
// returned has arrived; a thread may do other work between its arrive and its wait.
inline void cluster_check(unsigned cl) {
    State &s = state();
    size_t live = 0;
    for (size_t i = (size_t)cl * s.cluster * s.nthreads; i < (size_t)(cl + 1) * s.cluster * s.nthreads; ++i) live += s.fibers[i].st != Fiber::DONE;
    ClusterBar &b = s.cbar[cl];
    if (b.arrived && b.arrived >= live) { b.arrived = 0; ++b.phase; s.progress = true; }
}
inline void cluster_arrive() {
    State &s = state();
    const unsigned cl = self().cta / s.cluster;
    ++s.cbar[cl].arrived;
    s.progress = true;
    cluster_check(cl);
}
inline void cluster_wait() {
    State &s = state();
    const unsigned cl = self().cta / s.cluster;
    ++s.cluster_barriers;
    while (s.cbar[cl].phase <= self().cl_waits) yield_to_scheduler();
    ++self().cl_waits;
    s.progress = true;
}
inline void cluster_barrier() { cluster_arrive(); cluster_wait(); }

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
    s.nthreads = nthreads;
    s.body = &body;
    s.fibers.resize((size_t)resident * nthreads);
    for (auto &f : s.fibers)
        if (f.stack_bytes != cfg.stack_bytes) { f.stack.reset(new unsigned char[cfg.stack_bytes]); f.stack_bytes = cfg.stack_bytes; }
    while (s.smem.size() < resident) {
        s.smem_store.emplace_back(new unsigned char[kSmemBytes + 128]);
        unsigned char *p = s.smem_store.back().get();
        s.smem.push_back(p + (128 - reinterpret_cast<uintptr_t>(p) % 128) % 128);
    }
    s.mbar.assign(resident, std::map<const void *, Mbar>());
    s.cbar.assign(resident / cluster, ClusterBar());
    for (unsigned first = 0; first < nctas; first += resident) {
        const unsigned live_ctas = nctas - first < resident ? nctas - first : resident;
        for (unsigned c = 0; c < live_ctas; ++c) {
            const unsigned b = first + c;
            memset(s.smem[c], 0xff, kSmemBytes);  // NaN pattern: reads of unwritten shared memory show up
            s.mbar[c].clear();
            s.cbar[c / cluster] = ClusterBar();
            for (unsigned i = 0; i < nthreads; ++i) {
                Fiber &f = s.fibers[(size_t)c * nthreads + i];
                f.st = Fiber::RUNNABLE;
                f.cl_waits = 0;
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
        auto run_cta = [&](unsigned c) {  // resume every runnable fiber of CTA c once
            for (unsigned k = 0; k < nthreads; ++k) {
                const size_t i = (size_t)c * nthreads + (s.reverse_order ? nthreads - 1 - k : k);
                Fiber &f = s.fibers[i];
                if (f.st != Fiber::RUNNABLE) continue;
                s.current = (int)i;
                threadIdx = f.tid;
                blockIdx = f.bid;
                swapcontext(&s.sched, &f.ctx);
            }
        };
        auto release_barriers = [&]() {  // release the block barriers every participant has reached; cluster phases
            for (unsigned c = 0; c < live_ctas; ++c) {
                size_t at = 0, run = 0, named[16] = {0}, want[16] = {0};
                for (unsigned i = 0; i < nthreads; ++i) {
                    const Fiber &f = s.fibers[(size_t)c * nthreads + i];
                    if (f.st == Fiber::AT_BARRIER) {
                        if (f.bar_id == 0) ++at;
                        else { ++named[f.bar_id & 15]; want[f.bar_id & 15] = f.bar_count; }
                    }
                    run += f.st == Fiber::RUNNABLE;
                }
                bool any_named = false;
                for (unsigned id = 1; id < 16; ++id) {
                    if (!named[id]) continue;
                    any_named = true;
                    if (named[id] < want[id]) continue;
                    for (unsigned i = 0; i < nthreads; ++i) {
                        Fiber &f = s.fibers[(size_t)c * nthreads + i];
                        if (f.st == Fiber::AT_BARRIER && f.bar_id == id) f.st = Fiber::RUNNABLE;
                    }
                    s.progress = true;
                }
                if (at && !run && !any_named) {
                    for (unsigned i = 0; i < nthreads; ++i) {
                        Fiber &f = s.fibers[(size_t)c * nthreads + i];
                        if (f.st == Fiber::AT_BARRIER) f.st = Fiber::RUNNABLE;
                    }
                    s.progress = true;
                }
            }
            for (unsigned cl = 0; cl < live_ctas / cluster; ++cl) cluster_check(cl);  // threads that returned no longer count
        };
        int idle_rounds = 0;
        for (;;) {
            s.progress = false;
            if (s.cta_priority == 0) {
                for (unsigned k = 0; k < live_ctas; ++k) run_cta(s.reverse_order ? live_ctas - 1 - k : k);
            } else {
                for (unsigned k = 0; k < live_ctas; ++k) {
                    const unsigned c = s.cta_priority == 1 ? k : live_ctas - 1 - k;
                    s.progress = false;
                    release_barriers();     // so that a favoured CTA parked at its own block barrier goes on at once
                    run_cta(c);
                    if (s.progress) break;  // this CTA got somewhere: keep favouring it
                }
            }
            release_barriers();
            size_t runnable = 0, parked = 0;
            for (size_t i = 0; i < nf; ++i) {
                runnable += s.fibers[i].st == Fiber::RUNNABLE;
                parked += s.fibers[i].st == Fiber::AT_BARRIER;
            }
            if (!runnable && !parked) break;  // every thread returned
            // a poll that yields BEFORE it looks (ld_acquire_gpu) sees a change one round late: only several rounds in
            // a row in which nothing at all happened are a deadlock
            idle_rounds = s.progress ? 0 : idle_rounds + 1;
            if (idle_rounds >= 4) {
                fprintf(stderr, "simt: deadlock (CTAs %u..%u): %zu thread(s) spinning, %zu parked at a barrier nobody else reaches\n",
                        first, first + live_ctas - 1, runnable, parked);
                return false;
            }
        }
    }
    return true;
}

// ---- hooks called by the kernels' PTX wrappers when SSFFT_EMUL is defined
inline Mbar &mbar_of(unsigned cta, const void *bar) { return state().mbar[cta][bar]; }
inline void mbar_complete_if_ready(Mbar &m) {
    if (m.armed && m.pending_tx == 0) { m.armed = false; ++m.phase; state().progress = true; }
}
inline void mbar_init(const void *bar, unsigned count = 1) { Mbar m; m.expected = count ? count : 1; mbar_of(self().cta, bar) = m; }
inline void mbar_arrive_n(Mbar &m) {
    if (++m.arrived >= m.expected) { m.arrived = 0; m.armed = true; }
    state().progress = true;
    mbar_complete_if_ready(m);
}
inline void mbar_arrive(const void *bar) { mbar_arrive_n(mbar_of(self().cta, bar)); }  // plain arrival
inline void mbar_expect_tx(const void *bar, unsigned bytes) {  // one arrival + expect-tx
    Mbar &m = mbar_of(self().cta, bar);
    m.pending_tx += bytes;
    mbar_arrive_n(m);
}
inline void land(Mbar &m, const Mbar::Pending &c) {
    memcpy(c.dst, c.src ? c.src : (const void *)c.value, c.bytes);
    m.pending_tx -= c.bytes;
}
inline void bulk_g2s(void *dst, const void *src, unsigned bytes, const void *bar) {
    State &s = state();
    Mbar &m = mbar_of(self().cta, bar);
    if (bytes % 16 || ((uintptr_t)dst & 15) || ((uintptr_t)src & 15)) {
        fprintf(stderr, "simt: cp.async.bulk needs 16-byte aligned addresses and sizes (dst %p src %p bytes %u)\n", dst, src, bytes);
        abort();
    }
    ++s.bulk_copies; s.bulk_bytes += bytes;
    Mbar::Pending c{dst, src, bytes, {0}};
    if (s.late_copy) { m.pending.push_back(c); return; }
    land(m, c);
    mbar_complete_if_ready(m);
}
// one piece of a TMA tensor copy (any alignment), or its zero fill outside the tensor
inline void tma_copy(void *dst, const void *src, unsigned bytes, const void *bar) {
    State &s = state();
    Mbar &m = mbar_of(self().cta, bar);
    s.bulk_bytes += bytes;
    Mbar::Pending c{dst, src, bytes, {0}};
    if (s.late_copy) { m.pending.push_back(c); return; }
    land(m, c);
    mbar_complete_if_ready(m);
}
inline void tma_zero(void *dst, unsigned bytes, const void *bar) {
    State &s = state();
    Mbar &m = mbar_of(self().cta, bar);
    Mbar::Pending c{dst, nullptr, bytes, {0}};
    if (s.late_copy) { m.pending.push_back(c); return; }
    land(m, c);
    mbar_complete_if_ready(m);
}
// st.async to CTA `rank` of my cluster: the value lands in the peer's shared memory and completes `bytes` of the PEER's
// mbarrier at the same shared-memory offset as `bar` (now, or -- late_copy -- when the peer waits for it)
inline void remote_store_tx(unsigned rank, void *peer_dst, const void *value, unsigned bytes, const void *bar) {
    State &s = state();
    Mbar &m = mbar_of(cta_of_rank(rank), bar);
    Mbar::Pending c{peer_dst, nullptr, bytes, {0}};
    memcpy(c.value, value, bytes);
    if (s.late_copy) { m.pending.push_back(c); return; }
    land(m, c);
    mbar_complete_if_ready(m);
}
// mbarrier.test_wait: has the phase with this parity completed?  Never blocks.
inline bool mbar_test(const void *bar, unsigned parity) {
    Mbar &m = mbar_of(self().cta, bar);
    for (auto &c : m.pending) land(m, c);
    m.pending.clear();
    mbar_complete_if_ready(m);
    return (m.phase & 1u) != parity;
}
inline void mbar_wait(const void *bar, unsigned parity) {
    State &s = state();
    Mbar &m = mbar_of(self().cta, bar);
    for (;;) {
        for (auto &c : m.pending) land(m, c);  // late mode: whatever has been sent so far lands when somebody waits
        m.pending.clear();
        mbar_complete_if_ready(m);
        if ((m.phase & 1u) != parity) break;   // the phase with this parity has completed
        yield_to_scheduler();
    }
    s.progress = true;
}

}  // namespace simt

inline void __syncthreads() {
    simt::State &s = simt::state();
    ++s.barriers;
    s.progress = true;
    simt::self().bar_id = 0;
    simt::self().st = simt::Fiber::AT_BARRIER;
    simt::yield_to_scheduler();
}
namespace simt {
// bar.sync id, count (id >= 1): released when `count` threads of the CTA are parked at barrier `id`; threads that do
// not take part (a producer warp) keep running
inline void named_barrier(unsigned id, unsigned count) {
    State &s = state();
    ++s.barriers;
    s.progress = true;
    self().bar_id = id;
    self().bar_count = count;
    self().st = Fiber::AT_BARRIER;
    yield_to_scheduler();
}
}  // namespace simt
