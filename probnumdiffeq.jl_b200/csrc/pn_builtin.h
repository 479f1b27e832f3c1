#pragma once
struct PnBuiltinEntry {
    const char* key;
    const void* fn;
    int d, n_p;
};
extern "C" const PnBuiltinEntry* pn_builtin_table(int* count);
