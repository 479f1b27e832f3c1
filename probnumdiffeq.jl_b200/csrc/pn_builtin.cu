// pn_builtin.cu -- ahead-of-time (nvcc, sm_100a) instantiations of the step kernels for the built-in
// problems BASELINE.json names.  Everything else is compiled at run time by NVRTC from the same headers
// (pnode_api.cu).  Keys must match pnode_api.cu::kernel_key().
#include "pn_kernel_entry.cuh"
#include "kernels/pn_builtin_problems.cuh"
#include "pn_builtin.h"
#include "kernels/pn_smooth.cuh"
#include "kernels/pn_smooth_reg.cuh"
#include "kernels/pn_smooth_col.cuh"
#include "kernels/pn_dense_gen.cuh"

template <int d, int Q, int G>
__global__ void __launch_bounds__(PN_SMR_THREADS, 1) pn_smooth_col_kernel(const __grid_constant__ PnSmoothParams P) {
    pn_smooth_col_entry<d, Q, G>(P);
}

template <int d, int Q, int G>
__global__ void __launch_bounds__(PN_SMR_THREADS, 1) pn_smooth_reg_kernel(const __grid_constant__ PnSmoothParams P) {
    pn_smooth_reg_entry<d, Q, G>(P);
}

#define PN_ENTRY(NAME, PROB, Q, G, A, DY, S) \
    { "builtin:" NAME ":small:q" #Q ":g" #G ":a" #A ":dyn" #DY ":s" #S, \
      (const void*)&pn_small_kernel<PROB, Q, G, (A != 0), (DY != 0), S>, PROB::d, PROB::n_p, \
      (const void*)&pn_init_kernel<PROB, Q, (A != 0)> }

#define PN_ENTRY_KRON(NAME, PROB, Q, A, DY, S) \
    { "builtin:" NAME ":kron:q" #Q ":g1:a" #A ":dyn" #DY ":s" #S, \
      (const void*)&pn_kron_kernel<PROB, Q, (A != 0), (DY != 0), S>, PROB::d, PROB::n_p, \
      (const void*)&pn_init_kernel<PROB, Q, (A != 0)> }

#define PN_ENTRY_CTA(NAME, PROB, Q, A, DY, S) \
    { "builtin:" NAME ":cta:q" #Q ":g256:a" #A ":dyn" #DY ":s" #S, \
      (const void*)&pn_cta_kernel<PROB, Q, (A != 0), (DY != 0), S>, PROB::d, PROB::n_p, \
      (const void*)&pn_init_kernel<PROB, Q, (A != 0)> }

#define PN_ENTRY_THREAD(NAME, PROB, Q, A, DY) \
    { "builtin:" NAME ":thread:q" #Q ":g1:a" #A ":dyn" #DY ":s0", \
      (const void*)&pn_thread_kernel<PROB, Q, (A != 0), (DY != 0)>, PROB::d, PROB::n_p, \
      (const void*)&pn_init_kernel<PROB, Q, (A != 0)> }

static const PnBuiltinEntry g_table[] = {
    /* filter-only solves of small states: one thread per trajectory (pn_thread.cuh) -- cfg5 (bench.py headline), cfg1 / cfg3 filter halves */
    PN_ENTRY_THREAD("rober", PnRober, 3, 1, 1),
    PN_ENTRY_THREAD("orego", PnOrego, 3, 1, 1),
    PN_ENTRY_THREAD("fitzhugh_nagumo", PnFitzHughNagumo, 3, 1, 1),
    PN_ENTRY_THREAD("vanderpol", PnVanDerPol, 5, 1, 1),
    PN_ENTRY_THREAD("lotka_volterra", PnLotkaVolterra, 3, 1, 1),
    /* cfg4: Pleiades EK1(order=3), D = 112, adaptive, CTA per trajectory */
    PN_ENTRY_CTA("pleiades", PnPleiades, 3, 1, 1, 0),
    /* cfg2: Lotka-Volterra EK0(order=3), adaptive, filter only, 1M trajectories */
    PN_ENTRY_KRON("lotka_volterra", PnLotkaVolterra, 3, 1, 1, 0),
    /* cfg5: ROBER / OREGO stiff EK1(order=3), adaptive, filter only (bench.py headline) */
    PN_ENTRY("rober", PnRober, 3, 2, 1, 1, 0),
    PN_ENTRY("orego", PnOrego, 3, 2, 1, 1, 0),
    /* cfg1: FitzHugh-Nagumo EK1(order=3) */
    PN_ENTRY("fitzhugh_nagumo", PnFitzHughNagumo, 3, 2, 1, 1, 0),
    /* cfg3: Van der Pol EK1(order=5) */
    PN_ENTRY("vanderpol", PnVanDerPol, 5, 2, 1, 1, 0),
    /* Lotka-Volterra through the dense EK1 path */
    PN_ENTRY("lotka_volterra", PnLotkaVolterra, 3, 2, 1, 1, 0),
};

/* rhs_name "builtin:<name>" -> struct in kernels/pn_builtin_problems.cuh (used when a configuration has no
   ahead-of-time instantiation and is compiled by NVRTC instead) */
extern "C" const char* pn_builtin_struct_name(const char* name) {
    static const struct { const char* n; const char* s; } map[] = {
        {"builtin:fitzhugh_nagumo", "PnFitzHughNagumo"}, {"builtin:lotka_volterra", "PnLotkaVolterra"},
        {"builtin:vanderpol", "PnVanDerPol"},           {"builtin:rober", "PnRober"},
        {"builtin:orego", "PnOrego"},                   {"builtin:pleiades", "PnPleiades"},
    };
    for (auto& m : map) {
        const char *a = m.n, *b = name;
        while (*a && *a == *b) { a++; b++; }
        if (!*a && !*b) return m.s;
    }
    return nullptr;
}

extern "C" const PnBuiltinEntry* pn_builtin_table(int* count) {
    *count = (int)(sizeof(g_table) / sizeof(g_table[0]));
    return g_table;
}

extern "C" const void* pn_smooth_kernel_ptr(void) { return (const void*)&pn_smooth_kernel; }
extern "C" const void* pn_dense_gen_kernel_ptr(void) { return (const void*)&pn_dense_gen_kernel; }

/* ahead-of-time instantiations of the register-resident smoother sweep (d, q, lanes): cfg1 FHN q=3, cfg3 VdP q=5 */
extern "C" const void* pn_smooth_reg_lookup(int d, int q, int g, int column_distributed) {
    static const struct { int d, q, g, col; const void* fn; } tab[] = {
        {2, 3, 4, 0, (const void*)&pn_smooth_reg_kernel<2, 3, 4>},
        {2, 5, 8, 0, (const void*)&pn_smooth_reg_kernel<2, 5, 8>},
        {2, 3, 4, 1, (const void*)&pn_smooth_col_kernel<2, 3, 4>},
        {2, 5, 8, 1, (const void*)&pn_smooth_col_kernel<2, 5, 8>},
    };
    for (auto& t : tab)
        if (t.d == d && t.q == q && t.g == g && t.col == column_distributed) return t.fn;
    return nullptr;
}
