// pn_kernel_entry.cuh -- __global__ entry points shared by the nvcc (built-in problems) and NVRTC
// (user-traced problems) builds.  One persistent launch solves the whole ensemble: the grid is sized to
// a multiple of the SM count and groups pull trajectory indices from an atomic counter.
#pragma once
#include "kernels/pn_small.cuh"

#ifndef PN_THREADS
#define PN_THREADS 128
#endif
#ifndef PN_MIN_BLOCKS
#define PN_MIN_BLOCKS 2
#endif

template <class Prob, int Q, int G, bool ADAPTIVE, bool DYNAMIC, int SAVE>
__global__ void __launch_bounds__(PN_THREADS, PN_MIN_BLOCKS) pn_small_kernel(const __grid_constant__ PnParams P) {
    pn_small_entry<Prob, Q, G, ADAPTIVE, DYNAMIC, SAVE>(P);
}
