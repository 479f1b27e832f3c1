// pn_kernel_entry.cuh -- __global__ entry points shared by the nvcc (built-in problems) and NVRTC
// (user-traced problems) builds.  One persistent launch solves the whole ensemble: the grid is sized to
// a multiple of the SM count and groups pull trajectory indices from an atomic counter.
#pragma once


/* One 256-thread CTA per SM (8 warps, 2 per scheduler, ~255 registers each).  The warps of a CTA run free: the two warps
   that share a scheduler drift into different phases of the time loop (the scalar, latency-bound first phase of one overlaps
   the FMA-dense Gram / Cholesky phase of the other): ROBER 2.39e9 -> 2.63e9 steps/s, OREGO 2.05e9 -> 2.32e9
   (profiles/r2f_variants.jsonl).  PN_USE_CTA_SYNC re-synchronises the CTA once per iteration instead, which paid while the
   unrolled loop body (Householder sweep, ~100 KB) thrashed the instruction cache (round 1). */
#ifndef PN_THREADS
#define PN_THREADS 256
#endif
#ifndef PN_MIN_BLOCKS
#define PN_MIN_BLOCKS 1
#endif
#ifdef PN_USE_CTA_SYNC
#define PN_CTA_SYNC 1
#endif

#include "kernels/pn_small.cuh"

template <class Prob, int Q, int G, bool ADAPTIVE, bool DYNAMIC, int SAVE>
__global__ void __launch_bounds__(PN_THREADS, PN_MIN_BLOCKS) pn_small_kernel(const __grid_constant__ PnParams P) {
    pn_small_entry<Prob, Q, G, ADAPTIVE, DYNAMIC, SAVE>(P);
}

template <class Prob, int Q, bool ADAPTIVE>
__global__ void __launch_bounds__(128) pn_init_kernel(const __grid_constant__ PnParams P, double* init_mu, double* init_dt) {
    pn_init_entry<Prob, Q, ADAPTIVE>(P, init_mu, init_dt);
}

#include "kernels/pn_kron.cuh"

/* EK0 / IsometricKronecker: one thread per trajectory */
/* order <= 3: 4 CTAs of 128 threads per SM (<= 128 registers, 4 warps per scheduler): +9 % over the unconstrained 152
   registers / 3 warps on cfg2; higher orders need the registers */
#ifndef PN_KRON_MIN_BLOCKS
#define PN_KRON_MIN_BLOCKS(Q) ((Q) <= 3 ? 4 : 1)
#endif
template <class Prob, int Q, bool ADAPTIVE, bool DYNAMIC, int SAVE>
__global__ void __launch_bounds__(128, PN_KRON_MIN_BLOCKS(Q)) pn_kron_kernel(const __grid_constant__ PnParams P) {
    pn_kron_entry<Prob, Q, ADAPTIVE, DYNAMIC, SAVE>(P);
}

#include "kernels/pn_blocks.cuh"

/* BlocksOfDiagonals layout (EK0 + multivariate diffusion, DiagonalEK1): one thread per trajectory */
template <class Prob, int Q, bool DIAG1, bool MV, bool ADAPTIVE, bool DYNAMIC, int SAVE>
__global__ void __launch_bounds__(128) pn_blocks_kernel(const __grid_constant__ PnParams P) {
    pn_blocks_entry<Prob, Q, DIAG1, MV, ADAPTIVE, DYNAMIC, SAVE>(P);
}

#include "kernels/pn_cta.cuh"

/* large D (Pleiades, D = 112): one CTA per trajectory, R staged in shared memory */
template <class Prob, int Q, bool ADAPTIVE, bool DYNAMIC, int SAVE>
__global__ void __launch_bounds__(PN_CTA_THREADS, PN_CTA_MIN_BLOCKS) pn_cta_kernel(const __grid_constant__ PnParams P) {
    pn_cta_entry<Prob, Q, ADAPTIVE, DYNAMIC, SAVE>(P);
}

#include "kernels/pn_gen.cuh"

/* general dense transition pairs (IOUP with a vector / matrix / Jacobian-updated rate): one warp per trajectory */
template <class Prob, int Q, bool ADAPTIVE, bool DYNAMIC, int SAVE>
__global__ void __launch_bounds__(32 * PN_GEN_WARPS) pn_gen_kernel(const __grid_constant__ PnParams P) {
    pn_gen_entry<Prob, Q, ADAPTIVE, DYNAMIC, SAVE>(P);
}

#include "kernels/pn_thread.cuh"

/* small states, filter only: one thread per trajectory, Gram matrix in registers, factor packed in shared memory */
template <class Prob, int Q, bool ADAPTIVE, bool DYNAMIC>
__global__ void __launch_bounds__(PN_THREAD_NT, PN_THREAD_MINB) pn_thread_kernel(const __grid_constant__ PnParams P) {
    pn_thread_entry<Prob, Q, ADAPTIVE, DYNAMIC>(P);
}
