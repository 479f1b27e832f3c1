#pragma once
struct PnBuiltinEntry {
    const char* key;
    const void* fn;
    int d, n_p;
    const void* init_fn; /* pn_init_kernel of the same (problem, order, adaptive) for the group-per-trajectory path, else null */
};
extern "C" const PnBuiltinEntry* pn_builtin_table(int* count);
extern "C" const char* pn_builtin_struct_name(const char* name);
extern "C" const void* pn_smooth_kernel_ptr(void);
extern "C" const void* pn_dense_gen_kernel_ptr(void); /* pn_dense_gen.cuh: dense-output predict, warp per (trajectory, saveat point) pair */
extern "C" const void* pn_smooth_reg_lookup(int d, int q, int g, int column_distributed);
