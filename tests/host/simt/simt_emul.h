// simt_emul.h -- run a CUDA kernel's source on the CPU, one fiber per CUDA thread (TEST INFRASTRUCTURE).
//
//   simt::launch(grid, block, [&] { kernel<Cfg>(args...); });
//
// Blocks run one after the other; inside a block every thread is a ucontext fiber executing the kernel body.
// __syncthreads() switches back to the scheduler, which resumes the next fiber; a barrier completes when every
// fiber that has not returned has arrived.  Since only one block is live at a time, `extern __shared__` arrays
// bind to one host array (ssfft::ssfft_smem below) that plays the block's shared memory.
//
// TMA bulk copy + mbarrier (fused.cuh, PF = 1 / 2) are emulated at the two extremes the hardware allows:
//   late_copy = false: the copy is performed AT ISSUE TIME (the earliest it could land), so a kernel that lets the
//                      copy overwrite shared memory some thread still has to read fails;
//   late_copy = true : the copy is performed when the first thread WAITS for it (the latest it could land), so a
//                      kernel that writes to the destination between issue and wait gets its data overwritten.
// mbar_wait() yields until the phase it waits for has completed; a wait that can never complete is reported as a
// deadlock instead of hanging.
#pragma once
#include <ucontext.h>

#include <cstdio>
#include <functional>
#include <vector>

#include "cuda_runtime.h"

uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

namespace ssfft {
alignas(128) unsigned char ssfft_smem[232448];  // 227 KiB: the opt-in maximum of one sm_100 CTA
}

namespace simt {

struct Fiber {
    ucontext_t ctx;
    std::vector<unsigned char> stack;
    uint3 tid;
    enum { RUNNABLE, AT_BARRIER, DONE } st = RUNNABLE;
};

struct State {
    ucontext_t sched;
    std::vector<Fiber> fibers;
    int current = -1;
    const std::function<void()> *body = nullptr;
    // emulated mbarrier (one per CTA is all the kernels use) and statistics
    unsigned mbar_phase = 0;
    long long mbar_pending_tx = 0;
    bool mbar_armed = false;
    unsigned long long barriers = 0, bulk_copies = 0, bulk_bytes = 0;
    bool progress = false;
    bool late_copy = false;
    struct Pending { void *dst; const void *src; unsigned bytes; };
    std::vector<Pending> pending;
};
inline State &state() { static State s; return s; }

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
}

// Runs `body` for every thread of every block.  Returns false on deadlock (threads waiting for something that no
// thread can provide any more).
inline bool launch(dim3 grid, dim3 block, const std::function<void()> &body, size_t stack_bytes = 512 * 1024) {
    State &s = state();
    gridDim = grid;
    blockDim = block;
    const unsigned nthreads = block.x * block.y * block.z;
    s.fibers.resize(nthreads);
    for (auto &f : s.fibers) if (f.stack.size() != stack_bytes) f.stack.assign(stack_bytes, 0);
    s.body = &body;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                blockIdx = uint3{bx, by, bz};
                s.mbar_phase = 0; s.mbar_pending_tx = 0; s.mbar_armed = false; s.pending.clear();
                memset(ssfft::ssfft_smem, 0xff, sizeof(ssfft::ssfft_smem));  // NaN pattern: reads of unwritten shared memory show up
                for (unsigned i = 0; i < nthreads; ++i) {
                    Fiber &f = s.fibers[i];
                    f.st = Fiber::RUNNABLE;
                    f.tid = uint3{i % block.x, (i / block.x) % block.y, i / (block.x * block.y)};
                    getcontext(&f.ctx);
                    f.ctx.uc_stack.ss_sp = f.stack.data();
                    f.ctx.uc_stack.ss_size = f.stack.size();
                    f.ctx.uc_link = nullptr;
                    makecontext(&f.ctx, (void (*)())fiber_entry, 0);
                }
                // Barrier semantics: a fiber that reaches __syncthreads() stays parked until every fiber that has not
                // returned is parked too; fibers spinning in mbar_wait() stay runnable.
                for (;;) {
                    unsigned runnable = 0, parked = 0;
                    s.progress = false;
                    for (unsigned i = 0; i < nthreads; ++i) {
                        Fiber &f = s.fibers[i];
                        if (f.st != Fiber::RUNNABLE) continue;
                        s.current = (int)i;
                        threadIdx = f.tid;
                        swapcontext(&s.sched, &f.ctx);
                    }
                    for (auto &f : s.fibers) { runnable += f.st == Fiber::RUNNABLE; parked += f.st == Fiber::AT_BARRIER; }
                    if (!runnable) {
                        if (!parked) break;  // every thread returned
                        for (auto &f : s.fibers) if (f.st == Fiber::AT_BARRIER) f.st = Fiber::RUNNABLE;
                        continue;
                    }
                    if (!s.progress) {
                        fprintf(stderr, "simt: deadlock in block (%u,%u,%u): %u thread(s) wait on an mbarrier nobody completes, %u at a barrier\n",
                                bx, by, bz, runnable, parked);
                        return false;
                    }
                }
            }
    return true;
}

// ---- hooks called by the kernels' PTX wrappers when SSFFT_EMUL is defined
inline void mbar_init() { State &s = state(); s.mbar_phase = 0; s.mbar_pending_tx = 0; s.mbar_armed = false; }
inline void mbar_complete_if_ready() {
    State &s = state();
    if (s.mbar_armed && s.mbar_pending_tx == 0) { s.mbar_armed = false; ++s.mbar_phase; s.progress = true; }
}
inline void mbar_expect_tx(unsigned bytes) {  // arrive (count 1) + expect-tx
    State &s = state();
    s.mbar_pending_tx += bytes;
    s.mbar_armed = true;
    mbar_complete_if_ready();
}
inline void bulk_g2s(void *dst, const void *src, unsigned bytes) {
    State &s = state();
    if (bytes % 16 || ((uintptr_t)dst & 15) || ((uintptr_t)src & 15)) {
        fprintf(stderr, "simt: cp.async.bulk needs 16-byte aligned addresses and sizes (dst %p src %p bytes %u)\n", dst, src, bytes);
        abort();
    }
    ++s.bulk_copies; s.bulk_bytes += bytes;
    if (s.late_copy) { s.pending.push_back({dst, src, bytes}); return; }
    memcpy(dst, src, bytes);
    s.mbar_pending_tx -= bytes;
    mbar_complete_if_ready();
}
inline void mbar_wait(unsigned parity) {
    State &s = state();
    for (auto &c : s.pending) { memcpy(c.dst, c.src, c.bytes); s.mbar_pending_tx -= c.bytes; }
    s.pending.clear();
    mbar_complete_if_ready();
    while ((s.mbar_phase & 1u) == parity) yield_to_scheduler();  // the phase with this parity has not completed yet
    s.progress = true;
}

}  // namespace simt

inline void __syncthreads() {
    simt::State &s = simt::state();
    ++s.barriers;
    s.progress = true;
    s.fibers[s.current].st = simt::Fiber::AT_BARRIER;
    simt::yield_to_scheduler();
}
