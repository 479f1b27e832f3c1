// cuda_shim.h -- just enough of the CUDA C++ dialect to compile the kernel headers of probnumdiffeq.jl_b200/csrc/kernels for the HOST
// with g++.  Two modes:
//   default       one "thread" per process: execution-space qualifiers vanish, threadIdx is 0, warp collectives are the identity on a
//                 warp of one, atomics are plain operations.  For the kernels that give a whole trajectory to ONE thread (pn_thread.cuh,
//                 the headline kernel; pn_kron.cuh, pn_blocks.cuh).
//   PN_SHIM_WARP  one CTA of one or more WARPS: every lane is a host thread (threadIdx is thread_local), the warp collectives (__shfl*_sync,
//                 __any / __all / __ballot_sync, __syncwarp) exchange through a per-warp slot array between two pthread barriers of that
//                 warp, __syncthreads* are barriers over the CTA, atomics are real atomics, "shared memory" is one array all lanes see.
//                 For the group kernels (pn_small.cuh, pn_smooth_col.cuh, pn_smooth_reg.cuh: G lanes per trajectory, one warp) and the
//                 CTA-per-trajectory kernel (pn_cta.cuh, 256 threads).  Because every cross-lane hand-over goes
//                 through a pthread barrier, ThreadSanitizer (-fsanitize=thread) sees exactly the ordering the device guarantees at
//                 __syncwarp / shuffle points and reports any shared-memory access pair that is not separated by one: a host-side
//                 racecheck of the exchange protocol (compute-sanitizer is closed on the GPU pool).
// Test infrastructure (tests/test_kernel_host_emulation.py); nothing in the product includes it.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#define PN_HOST_EMULATION 1
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __restrict__
#define __shared__
#define __align__(x) __attribute__((aligned(x)))

struct PnShimDim3 {
    unsigned x, y, z;
};
#ifdef PN_SHIM_WARP
#include <pthread.h>
extern thread_local PnShimDim3 threadIdx; /* set by the lane's host thread (pn_warp_host.cpp) */
extern PnShimDim3 blockDim;               /* {lanes, 1, 1} */
static const PnShimDim3 blockIdx{0, 0, 0}, gridDim{1, 1, 1};
struct PnShimWarp {
    pthread_barrier_t bar;
    pthread_barrier_t sub[4][16]; /* barriers of the aligned lane groups of 2, 4, 8, 16 lanes: collectives with a group mask (a subset of the warp
                                     that calls from divergent code, e.g. pn_ioup.cuh's butterflies in the retry phase) synchronise the group only */
    uint64_t slot[32];
    int lanes; /* host threads in the warp: a multiple of the group size, <= 32 (lanes that do not exist never take part) */
};
struct PnShimCta { /* a CTA of up to 32 warps: warp collectives stay inside a warp, __syncthreads* span the CTA */
    PnShimWarp warp[32];
    pthread_barrier_t bar;
    int vote[1024];
    int threads;
};
extern PnShimCta pn_shim_cta;
#define pn_shim_warp (pn_shim_cta.warp[threadIdx.x >> 5])
static inline void pn_shim_barrier(unsigned mask = 0xffffffffu) {
    PnShimWarp& w = pn_shim_warp;
    const unsigned all = w.lanes >= 32 ? 0xffffffffu : ((1u << w.lanes) - 1u);
    if ((mask & all) == all) {
        pthread_barrier_wait(&w.bar);
        return;
    }
    const int sz = __builtin_popcount(mask), first = __builtin_ctz(mask);
    const int lg = __builtin_ctz((unsigned)sz) - 1;
    if ((sz & (sz - 1)) || sz < 2 || sz > 16 || first % sz || mask != (((1u << sz) - 1u) << first) || !((mask >> (threadIdx.x & 31)) & 1u)) __builtin_trap();
    pthread_barrier_wait(&w.sub[lg][first / sz]);
}
static inline void pn_shim_cta_barrier() { pthread_barrier_wait(&pn_shim_cta.bar); }
#else
static const PnShimDim3 threadIdx{0, 0, 0}, blockIdx{0, 0, 0}, blockDim{1, 1, 1}, gridDim{1, 1, 1};
#endif

struct double2 {
    double x, y;
};
static inline double2 make_double2(double x, double y) { return double2{x, y}; }

#ifdef PN_SHIM_WARP
/* publish -> barrier -> read the source lane's slot -> barrier (the second one keeps a fast lane's next publish off a slow lane's read) */
template <class T>
static inline T pn_shim_exchange(unsigned mask, T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffles move at most 64 bits here");
    const int lane = threadIdx.x & 31;
    uint64_t b = 0;
    std::memcpy(&b, &v, sizeof(T));
    pn_shim_warp.slot[lane] = b;
    pn_shim_barrier(mask);
    const uint64_t r = pn_shim_warp.slot[(src_lane >= 0 && src_lane < pn_shim_warp.lanes) ? src_lane : lane];
    pn_shim_barrier(mask);
    T o;
    std::memcpy(&o, &r, sizeof(T));
    return o;
}
template <class T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    const int lane = threadIdx.x & 31;
    return pn_shim_exchange(mask, v, (lane & ~(width - 1)) | (src & (width - 1)));
}
template <class T>
static inline T __shfl_xor_sync(unsigned mask, T v, int m, int width = 32) {
    const int lane = threadIdx.x & 31, s = lane ^ m;
    return pn_shim_exchange(mask, v, ((s & ~(width - 1)) == (lane & ~(width - 1))) ? s : lane);
}
template <class T>
static inline T __shfl_down_sync(unsigned mask, T v, int dlt, int width = 32) {
    const int lane = threadIdx.x & 31, s = lane + dlt;
    return pn_shim_exchange(mask, v, ((s & ~(width - 1)) == (lane & ~(width - 1))) ? s : lane);
}
template <class T>
static inline T __shfl_up_sync(unsigned mask, T v, int dlt, int width = 32) {
    const int lane = threadIdx.x & 31, s = lane - dlt;
    return pn_shim_exchange(mask, v, (s >= 0 && (s & ~(width - 1)) == (lane & ~(width - 1))) ? s : lane);
}
#ifdef PN_SHIM_DROP_SYNCWARP /* negative control of the race check: without the kernels' __syncwarp()s ThreadSanitizer must report */
static inline void __syncwarp(unsigned = 0xffffffffu) {}
#else
static inline void __syncwarp(unsigned mask = 0xffffffffu) { pn_shim_barrier(mask); }
#endif
static inline unsigned __ballot_sync(unsigned, int p) {
    const int lane = threadIdx.x & 31;
    pn_shim_warp.slot[lane] = p ? 1u : 0u;
    pn_shim_barrier();
    unsigned m = 0;
    for (int l = 0; l < pn_shim_warp.lanes; l++) m |= (unsigned)pn_shim_warp.slot[l] << l;
    pn_shim_barrier();
    return m;
}
static inline int __any_sync(unsigned k, int p) { return __ballot_sync(k, p) != 0u; }
static inline int __all_sync(unsigned k, int p) { return __ballot_sync(k, !p) == 0u; }
static inline unsigned __activemask() { return pn_shim_warp.lanes >= 32 ? 0xffffffffu : ((1u << pn_shim_warp.lanes) - 1u); }
static inline void __syncthreads() { pn_shim_cta_barrier(); }
static inline int pn_shim_cta_count(int p) {
    pn_shim_cta.vote[threadIdx.x] = p ? 1 : 0;
    pn_shim_cta_barrier();
    int c = 0;
    for (int t = 0; t < pn_shim_cta.threads; t++) c += pn_shim_cta.vote[t];
    pn_shim_cta_barrier();
    return c;
}
static inline int __syncthreads_or(int p) { return pn_shim_cta_count(p) != 0; }
static inline int __syncthreads_and(int p) { return pn_shim_cta_count(p) == pn_shim_cta.threads; }
#else
template <class T>
static inline T __shfl_sync(unsigned, T v, int, int = 32) { return v; }
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int, int = 32) { return v; }
template <class T>
static inline T __shfl_down_sync(unsigned, T v, int, int = 32) { return v; }
template <class T>
static inline T __shfl_up_sync(unsigned, T v, int, int = 32) { return v; }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline void __syncthreads() {}
static inline int __syncthreads_or(int p) { return p; }
static inline int __syncthreads_and(int p) { return p; }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
static inline int __any_sync(unsigned, int p) { return p; }
static inline int __all_sync(unsigned, int p) { return p; }
static inline unsigned __activemask() { return 1u; }
#endif
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
template <class T>
static inline T __ldg(const T* p) { return *p; }

#ifdef PN_SHIM_WARP
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
#else
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
    const unsigned long long o = *p;
    *p = o + v;
    return o;
}
static inline int atomicAdd(int* p, int v) {
    const int o = *p;
    *p = o + v;
    return o;
}
static inline double atomicAdd(double* p, double v) {
    const double o = *p;
    *p = o + v;
    return o;
}
#endif

static inline int __double2hiint(double x) {
    int64_t b;
    std::memcpy(&b, &x, 8);
    return (int)(b >> 32);
}
static inline int __double2loint(double x) {
    int64_t b;
    std::memcpy(&b, &x, 8);
    return (int)(b & 0xffffffff);
}
static inline double __hiloint2double(int hi, int lo) {
    const int64_t b = ((int64_t)hi << 32) | (uint32_t)lo;
    double x;
    std::memcpy(&x, &b, 8);
    return x;
}
static inline double __longlong_as_double(long long b) {
    double x;
    std::memcpy(&x, &b, 8);
    return x;
}
static inline long long __double_as_longlong(double x) {
    long long b;
    std::memcpy(&b, &x, 8);
    return b;
}
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
using std::fma;
using std::fabs;
using std::fmax;
using std::fmin;
using std::sqrt;
using std::log;
using std::exp;
using std::pow;
using std::copysign;
using std::nextafter;
using std::isfinite;
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }

// seeds of the accuracy of MUFU.RCP64H / MUFU.RSQ64H (about 2^-23): the kernels refine them with two Newton steps
// (full exponent range like the hardware seed, mantissa cut to 23 bits)
static inline double pn_host_cut23(double v) {
    int64_t b;
    std::memcpy(&b, &v, 8);
    b &= ~((1LL << 29) - 1);
    std::memcpy(&v, &b, 8);
    return v;
}
static inline double pn_host_seed_rcp(double x) { return pn_host_cut23(1.0 / x); }
static inline double pn_host_seed_rsqrt(double x) { return pn_host_cut23(1.0 / std::sqrt(x)); }
