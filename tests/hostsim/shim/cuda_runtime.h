// TEST INFRASTRUCTURE.  Stand-in for <cuda_runtime.h> that lets g++ compile the lane-level device code of
// maple_b200/csrc (glist.cuh, likelihood.cuh, search.cuh, place.cuh) for the HOST, one "lane" at a time, so that the very
// source the kernels are built from can be checked against the reference's golden vectors in a container without a GPU
// (tests/hostsim/hostsim.cpp).  Warp-level code (search_fsm.cuh) is not covered: it needs the hardware.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <math.h>

#define MAPLE_HOST_LANES 1  // place_scan.cuh: the lanes of a phase run one after the other
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))

template <class T>
static inline T __ldg(const T* p) { return *p; }
static inline int __double2hiint(double x) { unsigned long long u; memcpy(&u, &x, 8); return (int)(u >> 32); }
static inline unsigned __activemask() { return 1u; }
static inline void __syncwarp(unsigned = 1u) {}
static inline int __any_sync(unsigned, int p) { return p != 0; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline double min(double a, double b) { return std::fmin(a, b); }
static inline double max(double a, double b) { return std::fmax(a, b); }
// Compile-only stand-ins for the warp intrinsics of search_fsm.cuh: its warp-cooperative scan (warp_scan_job) is never executed
// on the host (the state machine is driven with scanMinSize = 0, i.e. search variant 2); they only let the file compile so that
// its lane-level state machine (fsm_step, fsm_finish) can run.
struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
static inline int __all_sync(unsigned, int p) { return p != 0; }
#define __restrict__
static inline void host_yield() {}
static inline void __threadfence() {}
static inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long cmp, unsigned long long v) { const unsigned long long o = *p; if (o == cmp) *p = v; return o; }
struct HostIdx { unsigned x = 0, y = 0, z = 0; };
static const HostIdx threadIdx, blockIdx;
template <class T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned) { return v; }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
static inline unsigned __match_any_sync(unsigned, int) { return 1u; }
static inline unsigned __fns(unsigned, unsigned, int) { return 0u; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline long long clock64() { return 0; }
using std::fabs;
using std::fmax;
using std::isfinite;
using std::log;
