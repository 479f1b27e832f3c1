// cuda_shim.h -- just enough of the CUDA C++ dialect to compile the kernel headers of probnumdiffeq.jl_b200/csrc/kernels for the HOST
// with g++, one "thread" per process: execution-space qualifiers vanish, threadIdx is 0, warp collectives are the identity on a warp of
// one, atomics are plain operations.  Only kernels that give a whole trajectory to ONE thread (pn_thread.cuh, the headline kernel)
// are meaningful under this shim; the group / CTA kernels exchange data between lanes and are checked on the device only.
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
#define __align__(x) alignas(x)

struct PnShimDim3 {
    unsigned x, y, z;
};
static const PnShimDim3 threadIdx{0, 0, 0}, blockIdx{0, 0, 0}, blockDim{1, 1, 1}, gridDim{1, 1, 1};

struct double2 {
    double x, y;
};
static inline double2 make_double2(double x, double y) { return double2{x, y}; }

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
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
template <class T>
static inline T __ldg(const T* p) { return *p; }

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
