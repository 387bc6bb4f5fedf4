// TEST INFRASTRUCTURE.  Stand-in for <cuda_runtime.h> that lets g++ compile the WARP-level device code of maple_b200/csrc
// (search_fsm.cuh: warp_scan_job, fsm_warp_loop; scan2.cuh: warp_scan_job2) for the host and run it with its 32 lanes emulated:
// every lane is a coroutine (ucontext) on one OS thread, and a *_sync warp intrinsic blocks the calling lane until every lane
// of its mask has arrived at an intrinsic with the same mask -- the semantics of independent thread scheduling -- at which
// point the exchange is performed for all of them.  So shuffles, ballots, __syncwarp, __any/__all, __match_any behave as on the
// device for code that is correct on the device; code that would dead-lock there (a lane of the mask never arrives) aborts
// here with a message.  Shared memory is ordinary memory.  See ../shim/cuda_runtime.h for the one-lane stand-in used for the
// lane-level code.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <math.h>
#include <ucontext.h>
#include <vector>

#define MAPLE_HOST_WARP 1
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __restrict__

struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

namespace hostwarp {

enum Op { OP_SHFL, OP_SHFL_UP, OP_SHFL_XOR, OP_BALLOT, OP_SYNC, OP_MATCH };

struct Lane {
    ucontext_t ctx;
    char* stack = nullptr;
    bool done = true, waiting = false;
    unsigned mask = 0;
    int op = 0;
    uint64_t val = 0, result = 0;
    int aux = 0;
};

struct Warp {
    Lane lane[32];
};

struct Grid {
    std::vector<Warp> warps;
    ucontext_t sched;
    int curWarp = 0, cur = 0;
    std::function<void(int)> body;  // body(warp index), run once per lane
};

inline Grid*& current() {
    static thread_local Grid* g = nullptr;
    return g;
}

inline void complete(Warp& W, unsigned mask, int op) {
    for (int l = 0; l < 32; l++) {
        if (!((mask >> l) & 1u)) continue;
        Lane& L = W.lane[l];
        uint64_t r = 0;
        switch (op) {
            case OP_SHFL: { const int s = L.aux & 31; r = ((mask >> s) & 1u) ? W.lane[s].val : L.val; break; }
            case OP_SHFL_UP: { const int s = l - L.aux; r = (s >= 0 && ((mask >> s) & 1u)) ? W.lane[s].val : L.val; break; }
            case OP_SHFL_XOR: { const int s = (l ^ L.aux) & 31; r = ((mask >> s) & 1u) ? W.lane[s].val : L.val; break; }
            case OP_BALLOT:
                for (int q = 0; q < 32; q++)
                    if (((mask >> q) & 1u) && W.lane[q].val) r |= 1ull << q;
                break;
            case OP_MATCH:
                for (int q = 0; q < 32; q++)
                    if (((mask >> q) & 1u) && W.lane[q].val == L.val) r |= 1ull << q;
                break;
            default: break;
        }
        L.result = r;
    }
    for (int l = 0; l < 32; l++)
        if ((mask >> l) & 1u) W.lane[l].waiting = false;
}

inline uint64_t collective(int op, unsigned mask, uint64_t val, int aux) {
    Grid& G = *current();
    Warp& W = G.warps[G.curWarp];
    const int me = G.cur;
    Lane& L = W.lane[me];
    if (!((mask >> me) & 1u)) { fprintf(stderr, "hostwarp: lane %d calls a *_sync intrinsic with mask %08x that excludes it\n", me, mask); abort(); }
    L.waiting = true; L.mask = mask; L.op = op; L.val = val; L.aux = aux;
    for (;;) {
        if (!L.waiting) return L.result;  // a lane of my group arrived last and performed the exchange
        bool all = true;
        for (int l = 0; l < 32 && all; l++)
            if ((mask >> l) & 1u) {
                const Lane& o = W.lane[l];
                if (o.done) { fprintf(stderr, "hostwarp: lane %d waits for lane %d, which has exited\n", me, l); abort(); }
                if (!o.waiting || o.mask != mask) all = false;
                else if (o.op != op) { fprintf(stderr, "hostwarp: lanes %d and %d meet at different intrinsics (%d vs %d)\n", me, l, op, o.op); abort(); }
            }
        if (all) { complete(W, mask, op); continue; }
        swapcontext(&L.ctx, &G.sched);
    }
}

// a spinning lane (waiting for another warp through memory) lets everybody else run
inline void yield() {
    Grid& G = *current();
    swapcontext(&G.warps[G.curWarp].lane[G.cur].ctx, &G.sched);
}

inline void trampoline() {
    Grid& G = *current();
    G.body(G.curWarp);
    Lane& L = G.warps[G.curWarp].lane[G.cur];
    L.done = true;
    swapcontext(&L.ctx, &G.sched);
}

// runs body(w) once per lane of each of nWarps warps, all lanes interleaved at the warp intrinsics and at spin_pause()
inline void run_warps(int nWarps, const std::function<void(int)>& body) {
    constexpr size_t kStack = 1 << 20;
    Grid* G = new Grid;
    G->warps.resize(nWarps);
    for (auto& W : G->warps)
        for (auto& L : W.lane) L.stack = (char*)malloc(kStack);
    Grid* prev = current();
    current() = G;
    G->body = body;
    for (int w = 0; w < nWarps; w++)
        for (int l = 0; l < 32; l++) {
            Lane& L = G->warps[w].lane[l];
            getcontext(&L.ctx);
            L.ctx.uc_stack.ss_sp = L.stack;
            L.ctx.uc_stack.ss_size = kStack;
            L.ctx.uc_link = nullptr;
            L.done = false; L.waiting = false;
            makecontext(&L.ctx, (void (*)())trampoline, 0);
        }
    unsigned long long idlePasses = 0;
    for (;;) {
        bool anyLeft = false, progressed = false;
        for (int w = 0; w < nWarps; w++)
            for (int l = 0; l < 32; l++) {
                Lane& L = G->warps[w].lane[l];
                if (L.done) continue;
                anyLeft = true;
                if (L.waiting) {  // runnable only once its group is complete: let it re-check
                    bool all = true;
                    for (int q = 0; q < 32 && all; q++)
                        if ((L.mask >> q) & 1u) all = G->warps[w].lane[q].waiting && G->warps[w].lane[q].mask == L.mask;
                    if (!all) continue;
                }
                G->curWarp = w;
                G->cur = l;
                swapcontext(&G->sched, &L.ctx);
                progressed = true;
            }
        if (!anyLeft) break;
        if (!progressed) { fprintf(stderr, "hostwarp: dead-lock, every live lane waits for a lane that is not coming\n"); abort(); }
        if (++idlePasses > 400000000ull) { fprintf(stderr, "hostwarp: gave up after 4e8 scheduler passes (live-lock between warps?)\n"); abort(); }
    }
    for (auto& W : G->warps)
        for (auto& L : W.lane) free(L.stack);
    current() = prev;
    delete G;
}

inline void run_warp(const std::function<void()>& body) { run_warps(1, [&](int) { body(); }); }

struct Idx { unsigned x, y, z; };
inline Idx tidx() { return Idx{current() ? (unsigned)current()->cur : 0u, 0u, 0u}; }
inline int warp_index() { return current() ? current()->curWarp : 0; }

template <class T>
inline uint64_t bits(T v) { uint64_t u = 0; memcpy(&u, &v, sizeof v); return u; }
template <class T>
inline T unbits(uint64_t u) { T v; memcpy(&v, &u, sizeof v); return v; }

}  // namespace hostwarp

#define threadIdx (hostwarp::tidx())
#define blockIdx (hostwarp::Idx{0u, 0u, 0u})
#define blockDim (hostwarp::Idx{32u, 1u, 1u})

template <class T>
static inline T __ldg(const T* p) { return *p; }
static inline int __double2hiint(double x) { unsigned long long u; memcpy(&u, &x, 8); return (int)(u >> 32); }
static inline unsigned __activemask() { return 0xffffffffu; }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { hostwarp::collective(hostwarp::OP_SYNC, mask, 0, 0); }
static inline int __any_sync(unsigned mask, int p) { return hostwarp::collective(hostwarp::OP_BALLOT, mask, p != 0, 0) != 0; }
static inline int __all_sync(unsigned mask, int p) { return (unsigned)hostwarp::collective(hostwarp::OP_BALLOT, mask, p != 0, 0) == mask; }
static inline unsigned __ballot_sync(unsigned mask, int p) { return (unsigned)hostwarp::collective(hostwarp::OP_BALLOT, mask, p != 0, 0); }
template <class T> static inline T __shfl_sync(unsigned mask, T v, int src) { return hostwarp::unbits<T>(hostwarp::collective(hostwarp::OP_SHFL, mask, hostwarp::bits(v), src)); }
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned d) { return hostwarp::unbits<T>(hostwarp::collective(hostwarp::OP_SHFL_UP, mask, hostwarp::bits(v), (int)d)); }
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int x) { return hostwarp::unbits<T>(hostwarp::collective(hostwarp::OP_SHFL_XOR, mask, hostwarp::bits(v), x)); }
template <class T> static inline unsigned __match_any_sync(unsigned mask, T v) { return (unsigned)hostwarp::collective(hostwarp::OP_MATCH, mask, hostwarp::bits(v), 0); }
// position of the offset-th set bit of mask at or above base (offset >= 1), 0xffffffff if there is none
static inline unsigned __fns(unsigned mask, unsigned base, int offset) {
    if (offset == 0) return ((mask >> base) & 1u) ? base : 0xffffffffu;
    for (unsigned b = base; b < 32; b++)
        if ((mask >> b) & 1u) { if (--offset == 0) return b; }
    return 0xffffffffu;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline long long clock64() { return 0; }
static inline void host_yield() { hostwarp::yield(); }
static inline void __threadfence() {}
static inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long cmp, unsigned long long v) { const unsigned long long o = *p; if (o == cmp) *p = v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline double min(double a, double b) { return std::fmin(a, b); }
static inline double max(double a, double b) { return std::fmax(a, b); }
using std::fabs;
using std::fmax;
using std::fmin;
using std::isfinite;
using std::log;
