// lockstep.h (host stand-in, warp-lockstep mode) -- TEST INFRASTRUCTURE, part of oracle/ref_shim.
//
// The sequential emulation of cuda_runtime.h cannot run kernels whose threads talk to each other.  The reference's zero-crossing
// extraction (FullScan6 / extract_kernel, kfusion/src/cuda/tsdf_volume.cu:511-710) is such a kernel: warp votes (__all, __ballot),
// a warp-synchronous prefix scan over `volatile` shared memory (scan_warp, :488-505) and a shared-memory staging of the points that
// relies on the 32 lanes of a warp executing every statement together.  With -DCUDAHOST_LOCKSTEP a launch runs ONE WARP AT A TIME
// as 32 fibers (ucontext) under a SIMT scheduler:
//   * every shared-memory access goes through a proxy (SArr / SPtr / Ref below; oracle/ref_shim/Makefile re-types the three
//     `__shared__` arrays, scan_warp's `volatile T*` parameter and the aliasing `volatile int*` with sed, on the fly) that parks the
//     lane with the program counter of the access (the access function's return address: the kernel and everything it calls are
//     force-inlined and compiled at -O0, so machine code order is source order);
//   * the scheduler always resumes the parked lane with the SMALLEST program counter -- the classic min-PC reconvergence rule --
//     so all lanes perform the loads of a statement before any lane performs its store, and statement k completes in every lane
//     before statement k+1 starts: the pre-Volta lock-step execution this code was written for;
//   * __ballot / __all / __any are warp barriers: lanes park until every live lane of the warp has voted;
//   * global atomics are naturally serialised (one lane runs at a time); warps and blocks run in launch order, so the global
//     append order of the points is deterministic (the tests compare the SORTED point set with the oracle's).
// Kernels with __syncthreads() are still not supported (extract_kernel has none).
#pragma once
#include <ucontext.h>
#include <vector>

namespace cudahost {

enum LaneState { LANE_READY, LANE_VOTE, LANE_DONE };

struct Lane {
    ucontext_t ctx;
    std::vector<char> stack;
    LaneState state;
    uintptr_t pc;          // program counter of the pending shared-memory access (0 = not parked at one)
    int pred;              // predicate handed to the pending vote
    unsigned vote_result, vote_active;
    uint3 tid;
};

struct WarpExec {
    ucontext_t sched;
    Lane lanes[32];
    int nlanes, current;
    void (*entry)(void *);
    void *arg;
};
inline WarpExec *g_warp = nullptr;

static inline void park()          // back to the scheduler
{
    WarpExec &w = *g_warp;
    swapcontext(&w.lanes[w.current].ctx, &w.sched);
}

__attribute__((noinline)) static void smem_access()
{
    if (!g_warp) return;
    Lane &l = g_warp->lanes[g_warp->current];
    l.pc = (uintptr_t)__builtin_return_address(0);
    l.state = LANE_READY;
    park();
}

static inline unsigned warp_vote(int pred, unsigned *active = nullptr)
{
    if (!g_warp) not_emulated("warp vote outside a lock-step launch");
    Lane &l = g_warp->lanes[g_warp->current];
    l.pred = pred;
    l.state = LANE_VOTE;
    park();
    if (active) *active = l.vote_active;
    return l.vote_result;
}

static inline unsigned ptx_special(const char *text)
{
    const int lane = g_warp ? g_warp->current : 0;
    if (std::strstr(text, "%laneid")) return (unsigned)lane;
    if (std::strstr(text, "%lanemask_lt")) return (1u << lane) - 1u;
    not_emulated(text);
}

static void lane_trampoline()
{
    WarpExec &w = *g_warp;
    w.entry(w.arg);
    w.lanes[w.current].state = LANE_DONE;
    park();
    std::abort();          // a finished lane is never resumed
}

// run one warp (lanes = consecutive flattened thread ids) to completion under the min-PC rule
static inline void run_warp(const uint3 *tids, int n, void (*entry)(void *), void *arg)
{
    static WarpExec w;
    w.nlanes = n; w.entry = entry; w.arg = arg;
    g_warp = &w;
    for (int i = 0; i < n; ++i) {
        Lane &l = w.lanes[i];
        if (l.stack.empty()) l.stack.resize(512 * 1024);
        getcontext(&l.ctx);
        l.ctx.uc_stack.ss_sp = l.stack.data();
        l.ctx.uc_stack.ss_size = l.stack.size();
        l.ctx.uc_link = nullptr;
        makecontext(&l.ctx, lane_trampoline, 0);
        l.state = LANE_READY; l.pc = 0; l.tid = tids[i];
    }
    for (;;) {
        int pick = -1, live = 0, voting = 0;
        for (int i = 0; i < n; ++i) {
            const Lane &l = w.lanes[i];
            if (l.state == LANE_DONE) continue;
            ++live;
            if (l.state == LANE_VOTE) { ++voting; continue; }
            if (pick < 0 || l.pc < w.lanes[pick].pc) pick = i;
        }
        if (!live) break;
        if (pick < 0) {                                   // every live lane has voted: exchange the predicates, release the warp
            unsigned mask = 0, active = 0;
            for (int i = 0; i < n; ++i)
                if (w.lanes[i].state == LANE_VOTE) { active |= 1u << i; if (w.lanes[i].pred) mask |= 1u << i; }
            for (int i = 0; i < n; ++i)
                if (w.lanes[i].state == LANE_VOTE) { w.lanes[i].vote_result = mask; w.lanes[i].vote_active = active; w.lanes[i].state = LANE_READY; w.lanes[i].pc = 0; }
            (void)voting;
            continue;
        }
        w.current = pick;
        threadIdx = w.lanes[pick].tid;
        swapcontext(&w.sched, &w.lanes[pick].ctx);
    }
    g_warp = nullptr;
}

// ---- shared-memory proxies --------------------------------------------------------------------------------------------------
#define CUDAHOST_AI inline __attribute__((always_inline))
template <class T> struct Ref {
    T *p;
    CUDAHOST_AI operator T() const { smem_access(); return *p; }
    CUDAHOST_AI T operator=(T v) const { smem_access(); *p = v; return v; }
    CUDAHOST_AI T operator=(const Ref &o) const { const T v = (T)o; smem_access(); *p = v; return v; }
};
template <class T> struct SPtr {
    T *base;
    CUDAHOST_AI Ref<T> operator[](long i) const { Ref<T> r = {base + i}; return r; }
    CUDAHOST_AI SPtr operator+(long i) const { SPtr s = {base + i}; return s; }
};
template <class T, int N> struct SArr {
    T data[N];
    CUDAHOST_AI Ref<T> operator[](long i) { Ref<T> r = {data + i}; return r; }
    CUDAHOST_AI SPtr<T> operator+(long i) { SPtr<T> s = {data + i}; return s; }
};
template <class U, class T> CUDAHOST_AI SPtr<U> sptr_cast(SPtr<T> p) { SPtr<U> s = {reinterpret_cast<U *>(p.base)}; return s; }

}  // namespace cudahost
