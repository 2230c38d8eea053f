// Test infrastructure — a minimal SIMT emulator for running ONE hand-written CUDA kernel's *logic* on the CPU.
//
// Why: kernels written while no GPU was at hand (tests/test_simt_largek.py) can still be executed against the oracle:
// block-wide barriers, warp votes/shuffles/matches, shared-memory atomics and divergence are modelled; timing, the memory
// model and data races are NOT (threads of a block are cooperative fibers on one OS thread, switched only at barriers and
// warp collectives).  A collective that can never complete (divergent __syncthreads, a vote some lane of the mask never
// reaches) is reported as a deadlock instead of hanging, and threads of a block meeting at different __syncthreads() calls
// as a divergent barrier.  This is a checker for tests/, never a product path.
#pragma once

#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <functional>
#include <new>
#include <map>
#include <string>
#include <vector>

namespace simt {

struct Dim3 {
    unsigned x = 1, y = 1, z = 1;
};

struct Rendezvous {
    uint64_t vals[32];
    uint64_t snap[32];
    unsigned arrived = 0;
    unsigned gen = 0;
};

// Context switch.  glibc's swapcontext makes a system call (signal mask) per switch, which dominates kernels that
// synchronise often; on x86-64 a dozen instructions do (callee-saved registers + stack pointer).  Elsewhere: ucontext.
#if defined(__x86_64__)
extern "C" void simt_switch(void** save_sp, void* load_sp);
asm(".text\n"
    ".globl simt_switch\n"
    ".type simt_switch,@function\n"
    "simt_switch:\n"
    "    pushq %rbp\n    pushq %rbx\n    pushq %r12\n    pushq %r13\n    pushq %r14\n    pushq %r15\n"
    "    movq %rsp, (%rdi)\n"
    "    movq %rsi, %rsp\n"
    "    popq %r15\n    popq %r14\n    popq %r13\n    popq %r12\n    popq %rbx\n    popq %rbp\n"
    "    ret\n"
    ".size simt_switch, .-simt_switch\n");
struct Context {
    void* sp = nullptr;
};
inline void context_switch(Context* from, Context* to) { simt_switch(&from->sp, to->sp); }
inline void context_make(Context* c, char* stack, size_t bytes, void (*entry)()) {
    uintptr_t top = ((uintptr_t)stack + bytes) & ~(uintptr_t)15;
    void** sp = (void**)(top - 64);
    for (int i = 0; i < 6; ++i) sp[i] = nullptr;   // r15, r14, r13, r12, rbx, rbp
    sp[6] = (void*)entry;                          // popped by `ret`; the entry then sees rsp = top - 8 (ABI alignment)
    sp[7] = nullptr;
    c->sp = sp;
}
#else
struct Context {
    ucontext_t uc;
};
inline void context_switch(Context* from, Context* to) { swapcontext(&from->uc, &to->uc); }
inline void context_make(Context* c, char* stack, size_t bytes, void (*entry)()) {
    getcontext(&c->uc);
    c->uc.uc_stack.ss_sp = stack;
    c->uc.uc_stack.ss_size = bytes;
    c->uc.uc_link = nullptr;
    makecontext(&c->uc, entry, 0);
}
#endif

struct Fiber {
    Context ctx;
    bool done = false;
    unsigned tid = 0;
    char waiting_on[48] = "";   // what the thread is blocked at (diagnostics of a reported deadlock)
    unsigned wait_seq = 0;      // converged waits this thread has begun (see warp_converged_wait)
};
constexpr size_t kStackBytes = 128 * 1024;
inline std::vector<std::vector<char>> g_stacks;   // reused from block to block

struct BlockState {
    std::vector<Fiber> fibers;
    Context sched;
    int current = -1;
    unsigned bar_arrived = 0, bar_gen = 0, live = 0;
    unsigned warp_passed[32] = {};     // per warp: sequence number of the last converged wait known to have passed
    int bar_site = 0;                  // source line of the __syncthreads() the current generation's first thread arrived from
    bool divergent = false;
    unsigned long long progress = 0;   // bumped whenever any collective completes or a thread exits
    std::map<std::pair<int, unsigned>, Rendezvous> rv;   // (warp, mask) -> rendezvous
    std::function<void()> body;
    const char* deadlock = nullptr;
};

inline BlockState* g_block = nullptr;
// 0: threads are resumed in ascending order every pass.  Otherwise: a different pseudo-random order every pass, so that which
// thread arrives last at a barrier (and therefore runs ahead of everyone else to the next one) changes from barrier to
// barrier — code that reads shared state after a barrier which a thread running ahead may already have changed then shows
// up as wrong results or as a divergent barrier.
inline unsigned g_schedule_seed = 0;
inline char g_message[200];
inline Dim3 threadIdx, blockIdx, blockDim, gridDim;

inline void yield_to_scheduler() {
    BlockState* b = g_block;
    Fiber& f = b->fibers[b->current];
    context_switch(&f.ctx, &b->sched);
}

inline void fiber_entry() {
    BlockState* b = g_block;
    b->body();
    Fiber& f = b->fibers[b->current];
    f.done = true;
    b->live--;
    b->progress++;
    // a barrier the remaining threads are waiting at may now be complete (exited threads do not take part)
    if (b->live > 0 && b->bar_arrived == b->live) {
        b->bar_arrived = 0;
        b->bar_gen++;
    }
    context_switch(&f.ctx, &b->sched);   // never resumed
    abort();
}

// Runs `body` once per thread of every block, blocks one after the other.  Returns null, or a message on deadlock.
inline const char* launch(unsigned grid, unsigned block, const std::function<void()>& body) {
    gridDim.x = grid;
    blockDim.x = block;
    for (unsigned bx = 0; bx < grid; ++bx) {
        BlockState b;
        g_block = &b;
        b.body = body;
        b.fibers.resize(block);
        b.live = block;
        blockIdx.x = bx;
        if (g_stacks.size() < block) g_stacks.resize(block);
        for (unsigned t = 0; t < block; ++t) {
            Fiber& f = b.fibers[t];
            f.tid = t;
            if (g_stacks[t].empty()) g_stacks[t].resize(kStackBytes);
            context_make(&f.ctx, g_stacks[t].data(), g_stacks[t].size(), fiber_entry);
        }
        unsigned long long last_progress = ~0ull;
        std::vector<unsigned> order(block);
        for (unsigned t = 0; t < block; ++t) order[t] = t;
        unsigned long long passes = 0;
        while (b.live > 0) {
            if (++passes > 4000000ull) {   // watchdog: something keeps moving without ever finishing
                g_block = nullptr;
                return "livelock: the block is still running after 4M scheduler passes";
            }
            if (b.progress == last_progress) {
                if (getenv("SIMT_EMU_VERBOSE")) {
                    std::map<std::string, std::vector<unsigned>> who;
                    for (const Fiber& f : b.fibers)
                        if (!f.done) who[f.waiting_on].push_back(f.tid);
                    for (const auto& kv : who) {
                        fprintf(stderr, "simt deadlock, block %u: %zu thread(s) at [%s]:", bx, kv.second.size(), kv.first.c_str());
                        for (size_t i = 0; i < kv.second.size() && i < 12; ++i) fprintf(stderr, " %u", kv.second[i]);
                        fprintf(stderr, "\n");
                    }
                }
                g_block = nullptr;
                return "deadlock: no thread of the block can make progress (divergent barrier or incomplete warp collective)";
            }
            last_progress = b.progress;
            if (g_schedule_seed) {   // Fisher-Yates with a small LCG
                for (unsigned i = block - 1; i > 0; --i) {
                    g_schedule_seed = g_schedule_seed * 1664525u + 1013904223u;
                    std::swap(order[i], order[(g_schedule_seed >> 8) % (i + 1)]);
                }
            }
            for (unsigned i = 0; i < block; ++i) {
                const unsigned t = order[i];
                if (b.fibers[t].done) continue;
                b.current = (int)t;
                threadIdx.x = t;
                context_switch(&b.sched, &b.fibers[t].ctx);
            }
        }
        g_block = nullptr;
        if (b.divergent) return g_message;
    }
    return nullptr;
}

// `site` = source line of the call.  The hardware matches barrier arrivals by count only; two groups of threads that meet at
// different __syncthreads() calls are a bug (undefined behaviour) even when the counts happen to add up, so it is reported.
inline void syncthreads(int site) {
    BlockState* b = g_block;
    const unsigned my_gen = b->bar_gen;
    if (b->bar_arrived == 0) b->bar_site = site;
    else if (b->bar_site != site && !b->divergent) {
        b->divergent = true;
        snprintf(g_message, sizeof(g_message), "divergent barrier: threads of one block met at different __syncthreads() calls (generated lines %d and %d)",
                 b->bar_site, site);
    }
    b->bar_arrived++;
    if (b->bar_arrived == b->live) {
        b->bar_arrived = 0;
        b->bar_gen++;
        b->progress++;
        return;
    }
    snprintf(b->fibers[b->current].waiting_on, 48, "__syncthreads line %d", site);
    while (b->bar_gen == my_gen) {
        yield_to_scheduler();
        threadIdx.x = g_block->fibers[g_block->current].tid;
    }
    b->fibers[b->current].waiting_on[0] = 0;
    b->progress++;
}
// A spin-wait on some memory-resident condition (an mbarrier phase) that the hardware evaluates ONCE for a converged warp: all
// lanes of the warp execute the same sequence of such waits, so the n-th wait of a warp has passed for every lane as soon as
// any lane saw its condition hold — even if the condition has changed again by the time a late fiber looks.  Spinning is not
// progress: a wait whose condition never comes true ends as a reported deadlock.
template <class Pred>
inline void warp_converged_wait(Pred pred, const char* what, const void* p, unsigned v) {
    BlockState* b = g_block;
    Fiber& f = b->fibers[b->current];
    const unsigned warp = f.tid >> 5;
    const unsigned seq = ++f.wait_seq;
    while (!(b->warp_passed[warp] >= seq || pred())) {
        snprintf(f.waiting_on, 48, "%s %p %u", what, p, v);
        yield_to_scheduler();
        threadIdx.x = f.tid;
    }
    f.waiting_on[0] = 0;
    if (b->warp_passed[warp] < seq) {
        b->warp_passed[warp] = seq;
        b->progress++;
    }
}
inline void note_wait(const char* what, const void* p, unsigned v) {
    BlockState* b = g_block;
    snprintf(b->fibers[b->current].waiting_on, 48, "%s %p %u", what, p, v);
}

// every lane named in `mask` contributes v; returns the 32 contributed values (valid until this lane's next collective)
inline const uint64_t* warp_exchange(unsigned mask, uint64_t v) {
    BlockState* b = g_block;
    const unsigned tid = b->fibers[b->current].tid;
    const int warp = (int)(tid >> 5), lane = (int)(tid & 31);
    // lanes of a partial last warp that do not exist cannot arrive
    unsigned exist = 0xffffffffu;
    const unsigned n = (unsigned)b->fibers.size();
    if ((unsigned)(warp + 1) * 32 > n) exist = (1u << (n - warp * 32)) - 1u;
    mask &= exist;
    if (!(mask & (1u << lane))) {
        fprintf(stderr, "simt: lane %d called a collective whose mask %08x does not name it\n", lane, mask);
        abort();
    }
    Rendezvous& r = b->rv[std::make_pair(warp, mask)];
    r.vals[lane] = v;
    r.arrived |= 1u << lane;
    const unsigned my_gen = r.gen;
    if (r.arrived == mask) {
        memcpy(r.snap, r.vals, sizeof(r.snap));
        r.arrived = 0;
        r.gen++;
        b->progress++;
    } else {
        snprintf(b->fibers[b->current].waiting_on, 48, "warp collective mask %08x", mask);
        while (r.gen == my_gen) {
            yield_to_scheduler();
            threadIdx.x = g_block->fibers[g_block->current].tid;
        }
        b->fibers[b->current].waiting_on[0] = 0;
        b->progress++;   // (a thread leaving a wait is a state change too: a pass is only dead when nobody moved at all)
    }
    return r.snap;
}

template <typename T>
inline uint64_t to_bits(T v) {
    uint64_t u = 0;
    static_assert(sizeof(T) <= 8, "collective payloads are at most 64 bits");
    memcpy(&u, &v, sizeof(T));
    return u;
}
template <typename T>
inline T from_bits(uint64_t u) {
    T v;
    memcpy(&v, &u, sizeof(T));
    return v;
}
inline int lane_id() { return (int)(g_block->fibers[g_block->current].tid & 31); }

}  // namespace simt

// ---- the CUDA surface the kernels under test use ---------------------------------------------------------------------------
using simt::blockDim;
using simt::blockIdx;
using simt::gridDim;
using simt::threadIdx;
using std::max;
using std::min;

// (when the CUDA host headers were included first they have their own idea of these)
#undef __global__
#undef __device__
#undef __host__
#undef __forceinline__
#undef __launch_bounds__
#undef __align__
#undef __shared__
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#define __align__(n) alignas(n)

#define __syncthreads() simt::syncthreads(__LINE__)
inline void __syncwarp(unsigned mask = 0xffffffffu) { simt::warp_exchange(mask, 0); }
template <typename T>
inline T __shfl_sync(unsigned mask, T v, int src) {
    return simt::from_bits<T>(simt::warp_exchange(mask, simt::to_bits(v))[src & 31]);
}
template <typename T>
inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask) {
    return simt::from_bits<T>(simt::warp_exchange(mask, simt::to_bits(v))[(simt::lane_id() ^ lane_mask) & 31]);
}
template <typename T>
inline T __shfl_down_sync(unsigned mask, T v, unsigned delta) {
    const int src = simt::lane_id() + (int)delta;
    const uint64_t* s = simt::warp_exchange(mask, simt::to_bits(v));
    return src > 31 ? v : simt::from_bits<T>(s[src]);
}
template <typename T>
inline T __shfl_up_sync(unsigned mask, T v, unsigned delta) {
    const int src = simt::lane_id() - (int)delta;
    const uint64_t* s = simt::warp_exchange(mask, simt::to_bits(v));
    return src < 0 ? v : simt::from_bits<T>(s[src]);
}
inline unsigned __ballot_sync(unsigned mask, bool pred) {
    const uint64_t* s = simt::warp_exchange(mask, pred ? 1u : 0u);
    unsigned out = 0;
    for (int l = 0; l < 32; ++l)
        if ((mask >> l & 1u) && s[l]) out |= 1u << l;
    return out;
}
inline bool __any_sync(unsigned mask, bool pred) { return __ballot_sync(mask, pred) != 0; }
template <typename T>
inline unsigned __match_any_sync(unsigned mask, T v) {
    const uint64_t mine = simt::to_bits(v);
    const uint64_t* s = simt::warp_exchange(mask, mine);
    unsigned out = 0;
    for (int l = 0; l < 32; ++l)
        if ((mask >> l & 1u) && s[l] == mine) out |= 1u << l;
    return out;
}
template <typename T>
inline T atomicAdd(T* p, T v) {
    const T old = *p;
    *p = old + v;
    return old;
}
template <typename T>
inline T atomicMax(T* p, T v) {
    const T old = *p;
    if (v > old) *p = v;
    return old;
}
template <typename T>
inline T atomicOr(T* p, T v) {
    const T old = *p;
    *p = old | v;
    return old;
}
inline unsigned __float_as_uint(float f) { return simt::from_bits<unsigned>(simt::to_bits(f)); }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline void __threadfence_system() {}
template <typename T>
inline T __ldg(const T* p) { return *p; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
#if !defined(__VECTOR_TYPES_H__)
struct float4 {
    float x, y, z, w;
};
inline float4 make_float4(float x, float y, float z, float w) { float4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }
struct uint2 {
    unsigned x, y;
};
#endif
