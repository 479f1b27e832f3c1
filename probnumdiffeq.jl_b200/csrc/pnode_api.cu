// pnode_api.cu -- C ABI of libpnode_cuda.so (include/pnode.h): configuration checks, NVRTC compilation of
// traced right-hand sides together with the step kernels, persistent-kernel launches, result handles.
//
// There is deliberately no CPU fallback: without a CUDA device every solve entry point fails with
// PNODE_ERR_NO_DEVICE.
#include <cuda.h>
#include <cuda_runtime.h>
#include <nvrtc.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pnode.h"
#include "kernels/pn_params.h"
#define PN_SMOOTH_DECL_ONLY
#include "kernels/pn_smooth.cuh"
#define PN_LTI_DECL_ONLY
#include "kernels/pn_lti.cuh" /* PN_GEN_WARPS, pn_gen_ws_doubles */
#define PN_THREAD_NT 256 /* pn_thread.cuh: threads per CTA of the ahead-of-time instantiations */
#define PN_SMR_THREADS 256 /* pn_smooth_reg.cuh */
#define PN_DENSE_DECL_ONLY
#include "kernels/pn_dense.cuh" /* PnDenseParams */
#include "kernels/pn_dense_gen.cuh" /* PN_DENSE_GEN_WS */
#include "kernels/pn_cta.cuh" /* PnCtaShared, PN_CTA_THREADS (templates are not instantiated here) */
#include "pn_builtin.h"
#include "pn_embedded_sources.h" /* generated: kernel headers as strings for NVRTC */

#define PN_THREADS 256

/* calibrate_solution! / set_diffusions! (src/integrator_utils.jl:46-88) for the per-step outputs of a
   FixedDiffusion solve: every saved covariance is scaled by the global MLE and sol.diffusions is overwritten. */
__global__ void pn_calibrate_saved(long long total, long long n_traj, const long long* __restrict__ off,
                                   const double* __restrict__ scale /* [n_traj] or null: sigma_hat^2 * initial_diffusion */, double scale_mul /* 1 / initial_diffusion */,
                                   const double* __restrict__ diffusion /* [n_traj] */, int dd,
                                   double* __restrict__ u_cov, double* __restrict__ s_diffusion, int d,
                                   const double* __restrict__ scale_mv /* [n_traj][d] or null */,
                                   const double* __restrict__ diffusion_mv /* [n_traj][d] or null */,
                                   double* __restrict__ s_diffusion_mv /* [total][d] or null */) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    long long lo = 0, hi = n_traj; /* largest tr with off[tr] <= i */
    while (hi - lo > 1) {
        long long mid = (lo + hi) >> 1;
        if (off[mid] <= i) lo = mid; else hi = mid;
    }
    if (scale_mv) { /* apply_diffusion! on blocks (apply_diffusion.jl:43-50): component a scaled by sigma_a^2 */
        /* sol.u has du = dd / du entries per point: d, or (du, u) = 2d for second-order problems (component = index mod d) */
        int du = 1;
        while (du * du < dd) du++;
        for (int a = 0; a < du; a++)
            for (int b = 0; b < du; b++) u_cov[i * dd + a * du + b] *= sqrt(scale_mv[lo * d + a % d] * scale_mv[lo * d + b % d]) * scale_mul;
    } else if (scale) {
        const double sc = scale[lo] * scale_mul;
        for (int k = 0; k < dd; k++) u_cov[i * dd + k] *= sc;
    }
    s_diffusion[i] = diffusion[lo];
    if (s_diffusion_mv)
        for (int a = 0; a < d; a++) s_diffusion_mv[i * d + a] = diffusion_mv[lo * d + a];
}

/* ... and the saved filter states of a filter-only solve (save_state): right factors scaled by sqrt(sigma_hat^2), column c
   belongs to component c mod d for the multivariate models (apply_diffusion.jl:35-104) */
__global__ void pn_calibrate_saved_state(long long total, long long n_traj, const long long* __restrict__ off,
                                         const double* __restrict__ scale, double scale_mul, const double* __restrict__ scale_mv, int D,
                                         int d, double* __restrict__ x_covfac) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total * D * D) return;
    const long long e = i / ((long long)D * D);
    const int c = (int)(i % D);
    long long lo = 0, hi = n_traj;
    while (hi - lo > 1) {
        const long long mid = (lo + hi) >> 1;
        if (off[mid] <= e) lo = mid; else hi = mid;
    }
    x_covfac[i] *= sqrt((scale_mv ? scale_mv[lo * d + c % d] : scale[lo]) * scale_mul);
}

/* Over-allocated (pitch) layout -> exact CSR: element e of trajectory tr lives at (tr * pitch + e) in the work buffers */
__global__ void pn_compact_saved(long long total, long long n_traj, const long long* __restrict__ off_exact, long long pitch,
                                 int width, const double* __restrict__ src, double* __restrict__ dst) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total * width) return;
    const long long e = i / width, w = i - e * width;
    long long lo = 0, hi = n_traj;
    while (hi - lo > 1) {
        const long long mid = (lo + hi) >> 1;
        if (off_exact[mid] <= e) lo = mid; else hi = mid;
    }
    dst[i] = src[(lo * pitch + (e - off_exact[lo])) * width + w];
}

/* Dense output, smoothed solutions: record 1 of every virtual trajectory is x_smooth[idx+1] of its parent trajectory
   (x_smooth[last] for a saveat point at or beyond the last saved time, with h2 = 0 so that the sweep copies it). */
__global__ void pn_dense_gather(long long V, long long n_saveat, int D, const long long* __restrict__ off,
                                const long long* __restrict__ counts, const long long* __restrict__ v_idx, const double* __restrict__ s_x_mean,
                                const double* __restrict__ s_x_covfac, double* __restrict__ v_mean, double* __restrict__ v_covfac) {
    const long long per = (long long)D * D + D;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= V * per) return;
    const long long v = i / per, e = i - v * per;
    const long long tr = v / n_saveat;
    const long long last = counts ? off[tr] + counts[tr] : off[tr + 1] - 1;
    long long rec = v_idx[v] + 1;
    if (rec > last) rec = last;
    if (e < (long long)D * D) v_covfac[(2 * v + 1) * (long long)D * D + e] = s_x_covfac[rec * (long long)D * D + e];
    else v_mean[(2 * v + 1) * (long long)D + (e - (long long)D * D)] = s_x_mean[rec * (long long)D + (e - (long long)D * D)];
}
/* ... and the marginals of the records 0 (of the records 1 where h2 == 0: the sweep's copy rule) are the dense output */
__global__ void pn_dense_collect(long long V, int d, const double* __restrict__ v_h, const double* __restrict__ vu_mean,
                                 const double* __restrict__ vu_cov, double* __restrict__ out_u, double* __restrict__ out_cov) {
    const long long per = (long long)d + (long long)d * d;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= V * per) return;
    const long long v = i / per, e = i - v * per;
    const long long src = (v_h[2 * v] == 0.0) ? 2 * v + 1 : 2 * v;
    if (e < d) out_u[v * d + e] = vu_mean[src * d + e];
    else out_cov[v * (long long)d * d + (e - d)] = vu_cov[src * (long long)d * d + (e - d)];
}

/* sol.stats totals of the ensemble: sum of naccept / nreject over the trajectories (two int64 instead of two arrays D2H) */
__global__ void pn_sum_counts(long long n, const long long* __restrict__ a, const long long* __restrict__ b,
                              unsigned long long* __restrict__ out) {
    unsigned long long sa = 0, sb = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        sa += (unsigned long long)a[i];
        sb += (unsigned long long)b[i];
    }
    for (int o = 16; o >= 1; o >>= 1) {
        sa += __shfl_xor_sync(0xffffffffu, sa, o);
        sb += __shfl_xor_sync(0xffffffffu, sb, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out, sa);
        atomicAdd(out + 1, sb);
    }
}

/* FP64 roofline probe: 8 independent DFMA chains per thread (the denominator bench.py reports against;
   MEASURED_PEAKS.json has no FP64 figure). */
__global__ void pn_dfma_probe(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * (long long)blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return fail(PNODE_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));             \
    } while (0)

/* ---- driver entry points (libcuda is never linked: the library must load on a CPU-only box) ---- */
typedef CUresult (*fn_cuModuleLoadData)(CUmodule*, const void*);
typedef CUresult (*fn_cuModuleUnload)(CUmodule);
typedef CUresult (*fn_cuModuleGetFunction)(CUfunction*, CUmodule, const char*);
typedef CUresult (*fn_cuLaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                                      CUstream, void**, void**);
typedef CUresult (*fn_cuOccupancy)(int*, CUfunction, int, size_t);
typedef CUresult (*fn_cuGetErrorString)(CUresult, const char**);
typedef CUresult (*fn_cuFuncSetAttribute)(CUfunction, CUfunction_attribute, int);
struct DriverApi {
    fn_cuModuleLoadData ModuleLoadData = nullptr;
    fn_cuModuleUnload ModuleUnload = nullptr;
    fn_cuModuleGetFunction ModuleGetFunction = nullptr;
    fn_cuLaunchKernel LaunchKernel = nullptr;
    fn_cuOccupancy Occupancy = nullptr;
    fn_cuGetErrorString GetErrorString = nullptr;
    fn_cuFuncSetAttribute FuncSetAttribute = nullptr;
    bool ok = false;
};
static DriverApi g_drv;
static int load_driver() {
    if (g_drv.ok) return 0;
    struct {
        const char* name;
        void** slot;
    } syms[] = {
        {"cuModuleLoadData", (void**)&g_drv.ModuleLoadData},
        {"cuModuleUnload", (void**)&g_drv.ModuleUnload},
        {"cuModuleGetFunction", (void**)&g_drv.ModuleGetFunction},
        {"cuLaunchKernel", (void**)&g_drv.LaunchKernel},
        {"cuOccupancyMaxActiveBlocksPerMultiprocessor", (void**)&g_drv.Occupancy},
        {"cuGetErrorString", (void**)&g_drv.GetErrorString},
        {"cuFuncSetAttribute", (void**)&g_drv.FuncSetAttribute},
    };
    for (auto& s : syms) {
        cudaDriverEntryPointQueryResult qr;
        cudaError_t e = cudaGetDriverEntryPoint(s.name, s.slot, cudaEnableDefault, &qr);
        if (e != cudaSuccess || qr != cudaDriverEntryPointSuccess || !*s.slot)
            return fail(PNODE_ERR_CUDA, std::string("cudaGetDriverEntryPoint(") + s.name + ") failed");
    }
    g_drv.ok = true;
    return 0;
}
static std::string cu_err(CUresult r) {
    const char* s = nullptr;
    if (g_drv.GetErrorString) g_drv.GetErrorString(r, &s);
    return s ? s : "unknown CUDA driver error";
}

/* ---------------------------------------------------------------------------------------------- */
static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

struct Kernel {
    int threads = PN_THREADS;
    const void* rt_fn = nullptr; /* nvcc-compiled (runtime API) */
    const void* rt_init = nullptr; /* pn_init_kernel (group-per-trajectory path) */
    CUfunction drv_init = nullptr;
    CUmodule mod = nullptr;      /* NVRTC-compiled */
    CUfunction drv_fn = nullptr;
    int blocks_per_sm = 1;
    bool valid() const { return rt_fn || drv_fn; }
};

struct pnode_solver {
    pnode_config cfg;
    std::string rhs_src, rhs_name;
    int D = 0, G = 4, n_sm = 0;
    bool builtin = false;
    bool kron = false; /* EK0: IsometricKronecker layout, one thread per trajectory */
    bool blocks = false; /* BlocksOfDiagonals layout (EK0 + multivariate diffusion, DiagonalEK1), thread per trajectory */
    bool mv = false;     /* multivariate (per-component) diffusion model */
    bool cta = false;  /* EK1, D > 32: one CTA per trajectory, R in shared memory */
    bool thread = false; /* small states, filter only: one thread per trajectory (pn_thread.cuh) */
    bool gen = false;  /* general-prior path (pn_gen.cuh): IOUP with a vector / matrix / Jacobian-updated rate, one warp per trajectory */
    double* d_rate = nullptr;   /* [d][d] rate matrix */
    double* d_gen_ws = nullptr; /* per-warp workspace when it does not fit in shared memory */
    size_t gen_ws_doubles = 0;  /* per warp */
    size_t smem = 0;   /* dynamic shared memory of the step kernel */
    Kernel k_final;  /* SAVE = 0 */
    Kernel k_save;   /* SAVE = 1 */
    Kernel k_dense;  /* dense-output predict kernel (pn_dense.cuh) */
    int Gd = 0;      /* lanes per (trajectory, saveat point) pair */
    double* d_saveat = nullptr;
    int64_t n_saveat = 0;
    Kernel k_smooth; /* register-resident smoother sweep (pn_smooth_reg.cuh); invalid => shared-memory sweep */
    int Gs = 0;      /* lanes per trajectory of the smoother sweep */
    size_t smem_smooth = 0;
    cudaStream_t stream = nullptr;
    unsigned long long* d_counter = nullptr;
    double* d_tstops = nullptr;
    std::vector<double> host_tstops;
    double* d_forced = nullptr;
    double* d_init_mu = nullptr; /* pn_init_kernel outputs, grown on demand */
    double* d_init_dt = nullptr;
    size_t init_cap = 0;
    double compile_s = 0.0;
    /* Fenrir data log-likelihood: device copies of the data set and the observation model */
    double *d_data_t = nullptr, *d_data_u = nullptr, *d_obs_H = nullptr, *d_obs_Rn = nullptr;
    bool data_forward = false; /* the data set drives forward updates in the step kernel (DataUpdateCallback) instead of the Fenrir pass */
    bool fenrir() const { return n_data > 0 && !data_forward; }
    int64_t n_data = 0;
    int n_obs = 0;
    std::string mass_src; /* generated preamble (PN_MASS, pn_mass, pn_mass_inv) when the problem has a mass matrix */
    PnParams base; /* constants filled at create */
};

struct Buf {
    void* d = nullptr;
    size_t bytes = 0;
};
struct pnode_result {
    pnode_solver* s = nullptr;
    int64_t n_traj = 0;
    Buf f[PNODE_F_COUNT];
    std::vector<int64_t> offsets;
    pnode_result_info info;
    bool owns_inputs = false;
    double* d_u0 = nullptr;
    double* d_p = nullptr;
};

/* ---------------------------------------------------------------------------------------------- */
extern "C" int pnode_abi_version(void) { return 1; }
extern "C" const char* pnode_last_error(void) { return g_err.c_str(); }
extern "C" int pnode_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

static double factorial_d(int k) {
    double f = 1.0;
    for (int i = 2; i <= k; i++) f *= (double)i;
    return f;
}

/* lanes per trajectory of the register-resident Householder sweeps (dense-output kernel; the step kernel when built with
   PN_PREDICT_CHOL=0): the 2D x D stack keeps 2 D ceil(D/G) doubles per lane */
static int auto_lanes_qr(int D) {
    const int cand[] = {2, 4, 8, 16, 32};
    for (int g : cand) {
        int rpl = (D + g - 1) / g;
        if (2 * rpl * D <= 80) return g;
    }
    return 32;
}
/* lanes per trajectory of the step kernel (Gram + Cholesky predict): a lane keeps its D ceil(D/G) entries of the factor and
   about half as many accumulators of the Gram matrix; the scalar part of a step (f, J, sigma^2, controller) is replicated
   in the group, so fewer lanes mean less replicated work: D = 12 runs 2 lanes (measured 2.06e9 against 1.45e9 steps/s
   with 4 on ROBER, profiles/r2c_variants.jsonl) */
static int auto_lanes(int D) {
    if (env_int("PNODE_QR_LANES", 0)) return auto_lanes_qr(D);
    const int cand[] = {2, 4, 8, 16, 32};
    for (int g : cand) {
        int rpl = (D + g - 1) / g;
        if (rpl * D <= 72) return g;
    }
    return 32;
}

/* lanes per trajectory of the register-resident smoother: the 2D x 2D sweep keeps 4 D ceil(D/G) doubles per lane */
static int smooth_lanes(int D) {
    const int cand[] = {2, 4, 8, 16, 32};
    for (int g : cand) {
        int rpl = (D + g - 1) / g;
        if (4 * rpl * D <= 100) return g;
    }
    return 0; /* too large: shared-memory sweep (pn_smooth.cuh) */
}
/* must match PnSmoothReg<d,Q,G>::SMEM_BYTES */
static size_t smooth_reg_smem_bytes(int d, int q, int G) {
    const int D = d * (q + 1), RPL = (D + G - 1) / G, ROWS = RPL * G, LD = D | 1;
    const int RAW = 2 * D * LD + ROWS * LD + 2 * D;
    const int PHASE = (G >= 16) ? 0 : G;
    const int REGION = ((RAW - PHASE + 15) / 16) * 16 + PHASE;
    return (size_t)REGION * (PN_SMR_THREADS / G) * 8;
}

/* group-per-trajectory kernel: threads per CTA (the ahead-of-time instantiations are built for PN_THREADS) and the dynamic
   shared memory of its per-group exchange regions (must match pn_xs_doubles<D>() in pn_small.cuh) */
static size_t small_smem_bytes(int D, int G, int threads) {
    const size_t xs = ((size_t)(D * D + 15) / 16) * 16 + 2;
    return (size_t)(threads / G) * xs * sizeof(double);
}

/* pn_thread.cuh::pn_thread_sm_doubles: packed factor + mean, doubles per thread */
static size_t thread_sm_doubles(int d, int n) {
    const int D = d * n;
    return (size_t)(2 * d * D + (D - 2 * d) * (D - 2 * d + 1) / 2 + D);
}
static bool is_dynamic(int diffusion) { return diffusion == PNODE_DIFF_DYNAMIC || diffusion == PNODE_DIFF_DYNAMIC_MV; }
static bool is_mv(int diffusion) { return diffusion == PNODE_DIFF_DYNAMIC_MV || diffusion == PNODE_DIFF_FIXED_MV; }

/* must match PnSmoothCol<d,Q,G>::SMEM_BYTES */
static size_t smooth_col_smem_bytes(int d, int q, int G) {
    const int D = d * (q + 1), LD = D | 1;
    const int recsz = D * D + D + ((D * D + D) & 1);
    const int off_rm = ((D % 2 == 0) ? 2 : 1) * recsz;
    const int RAW = off_rm + 2 * D * LD + 2 * D;
    const int PHASE = (G >= 16) ? 0 : G;
    const int REGION = ((RAW - PHASE + 15) / 16) * 16 + PHASE;
    return (size_t)REGION * (PN_SMR_THREADS / G) * 8;
}

static std::string kernel_key(const pnode_solver* s, int save) {
    char buf[256];
    const char* kind = s->blocks ? (s->cfg.alg == PNODE_DIAGONAL_EK1 ? (s->mv ? "blocks1mv" : "blocks1") : (s->mv ? "blocks0mv" : "blocks0"))
                                 : (s->kron ? "kron" : (s->cta ? "cta" : (s->thread ? "thread" : "small")));
    snprintf(buf, sizeof buf, "%s:%s:q%d:g%d:a%d:dyn%d:s%d", s->rhs_name.c_str(), kind, s->cfg.q, s->G, s->cfg.adaptive ? 1 : 0,
             is_dynamic(s->cfg.diffusion) ? 1 : 0, save);
    return buf;
}

/* Process-wide cache of NVRTC results, keyed by the complete input of a compilation (program text + every option; the embedded headers
   are fixed per build of this library).  Creating a second solver for the same right-hand side and configuration -- every call of the host
   mirrors' `solve` does -- then costs a lookup instead of seconds of compile time.  PNODE_NVRTC_CACHE=0 turns it off. */
#include <map>
#include <mutex>
static std::mutex g_cubin_mutex;
static std::map<std::string, std::vector<char>> g_cubin_cache;
static size_t g_cubin_cache_bytes = 0;
static int nvrtc_compile_uncached(const std::string& src, std::vector<std::string> defs, std::vector<char>* cubin);
static int nvrtc_compile(const std::string& src, std::vector<std::string> defs, std::vector<char>* cubin) {
    const char* off = getenv("PNODE_NVRTC_CACHE");
    if (off && atoi(off) == 0) return nvrtc_compile_uncached(src, defs, cubin);
    std::string key;
    for (auto& x : defs) {
        key += x;
        key.push_back('\n');
    }
    if (const char* e = getenv("PNODE_NVRTC_FLAGS")) key += e;
    key.push_back('\0');
    key += src;
    {
        std::lock_guard<std::mutex> lk(g_cubin_mutex);
        auto it = g_cubin_cache.find(key);
        if (it != g_cubin_cache.end()) {
            *cubin = it->second;
            return 0;
        }
    }
    int rc = nvrtc_compile_uncached(src, defs, cubin);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(g_cubin_mutex);
    if (g_cubin_cache_bytes + cubin->size() > ((size_t)512 << 20)) { /* bound the footprint: start over */
        g_cubin_cache.clear();
        g_cubin_cache_bytes = 0;
    }
    if (g_cubin_cache.emplace(key, *cubin).second) g_cubin_cache_bytes += cubin->size();
    return 0;
}
static int nvrtc_compile_uncached(const std::string& src, std::vector<std::string> defs, std::vector<char>* cubin) {
    nvrtcProgram prog;
    const char* hdr_src[PN_N_EMBEDDED];
    const char* hdr_name[PN_N_EMBEDDED];
    for (int i = 0; i < PN_N_EMBEDDED; i++) {
        hdr_src[i] = pn_embedded_sources[i].text;
        hdr_name[i] = pn_embedded_sources[i].name;
    }
    if (nvrtcCreateProgram(&prog, src.c_str(), "pnode_user.cu", PN_N_EMBEDDED, hdr_src, hdr_name) != NVRTC_SUCCESS)
        return fail(PNODE_ERR_COMPILE, "nvrtcCreateProgram failed");
    /* -default-device: the compile-time loops of pn_thread.cuh pass generic lambdas, which carry no execution-space annotation */
    std::vector<const char*> opts = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "-default-device"};
    if (const char* e = getenv("PNODE_NVRTC_FLAGS")) { /* developer knob: extra NVRTC options, space separated */
        std::string cur;
        for (const char* c = e;; c++) {
            if (*c == ' ' || *c == 0) {
                if (!cur.empty()) defs.push_back(cur);
                cur.clear();
                if (!*c) break;
            } else cur.push_back(*c);
        }
    }
    for (auto& x : defs) opts.push_back(x.c_str());
    nvrtcResult cr = nvrtcCompileProgram(prog, (int)opts.size(), opts.data());
    if (cr != NVRTC_SUCCESS) {
        size_t n = 0;
        nvrtcGetProgramLogSize(prog, &n);
        std::string log(n, '\0');
        if (n) nvrtcGetProgramLog(prog, &log[0]);
        nvrtcDestroyProgram(&prog);
        return fail(PNODE_ERR_COMPILE, "NVRTC compilation failed:\n" + log);
    }
    size_t nb = 0;
    nvrtcGetCUBINSize(prog, &nb);
    cubin->resize(nb);
    nvrtcGetCUBIN(prog, cubin->data());
    nvrtcDestroyProgram(&prog);
    return 0;
}

/* NVRTC: user RHS + step kernels -> sm_100a CUBIN.  Needs no device. */
static int nvrtc_cubin(const std::string& rhs_src, const std::string& struct_name, int q, int G, bool adaptive,
                       bool dynamic, int save, std::vector<char>* cubin, int threads = PN_THREADS, int min_blocks = 1,
                       bool kron = false, bool cta = false, int blocks = 0 /* 0 no, 1 EK0, 2 DiagonalEK1 */, bool mv = false,
                       bool ioup = false, bool ek0_dense = false, const std::string& preamble = std::string(), bool gen = false, bool thread = false) {
    std::string src = preamble;
    src += "#include \"pn_kernel_entry.cuh\"\n";
    src += rhs_src;
    if (thread)
        src += "\nextern \"C\" __global__ void __launch_bounds__(PN_THREAD_NT, PN_THREAD_MINB) pn_entry(const __grid_constant__ PnParams P) {\n"
               "  pn_thread_entry<" + struct_name + ", PN_Q, (PN_ADAPTIVE != 0), (PN_DYNAMIC != 0)>(P);\n}\n";
    else if (gen)
        src += "\nextern \"C\" __global__ void __launch_bounds__(PN_THREADS) pn_entry(const __grid_constant__ PnParams P) {\n"
               "  pn_gen_entry<" + struct_name + ", PN_Q, (PN_ADAPTIVE != 0), (PN_DYNAMIC != 0), PN_SAVE>(P);\n}\n";
    else if (blocks)
        src += std::string("\nextern \"C\" __global__ void __launch_bounds__(PN_THREADS) pn_entry(const __grid_constant__ PnParams P) {\n"
               "  pn_blocks_entry<") + struct_name + ", PN_Q, " + (blocks == 2 ? "true" : "false") + ", " + (mv ? "true" : "false") +
               ", (PN_ADAPTIVE != 0), (PN_DYNAMIC != 0), PN_SAVE>(P);\n}\n";
    else if (cta)
        src += "\nextern \"C\" __global__ void __launch_bounds__(PN_CTA_THREADS, PN_CTA_MIN_BLOCKS) pn_entry(const __grid_constant__ PnParams P) {\n"
               "  pn_cta_entry<" + struct_name + ", PN_Q, (PN_ADAPTIVE != 0), (PN_DYNAMIC != 0), PN_SAVE>(P);\n}\n";
    else if (kron)
        src += "\nextern \"C\" __global__ void __launch_bounds__(PN_THREADS, PN_KRON_MIN_BLOCKS(PN_Q)) pn_entry(const __grid_constant__ PnParams P) {\n"
               "  pn_kron_entry<" + struct_name + ", PN_Q, (PN_ADAPTIVE != 0), (PN_DYNAMIC != 0), PN_SAVE>(P);\n}\n";
    else
        src += "\nextern \"C\" __global__ void __launch_bounds__(PN_THREADS, PN_MIN_BLOCKS) pn_entry(const "
               "__grid_constant__ PnParams P) {\n  pn_small_entry<" +
               struct_name + ", PN_Q, PN_G, (PN_ADAPTIVE != 0), (PN_DYNAMIC != 0), PN_SAVE>(P);\n}\n";
    /* every step kernel takes its initial states from the pre-pass */
    src += "extern \"C\" __global__ void __launch_bounds__(128) pn_init(const __grid_constant__ PnParams P, double* init_mu, "
               "double* init_dt) {\n  pn_init_entry<" + struct_name + ", PN_Q, (PN_ADAPTIVE != 0)>(P, init_mu, init_dt);\n}\n";
    auto def = [](const char* name, int v) { return std::string("-D") + name + "=" + std::to_string(v); };
    return nvrtc_compile(src, {def(cta ? "PN_CTA_THREADS" : "PN_THREADS", threads), def(cta ? "PN_CTA_MIN_BLOCKS" : "PN_MIN_BLOCKS", min_blocks), def("PN_Q", q), def("PN_G", G),
                               def("PN_THREAD_NT", thread ? threads : 256), def("PN_THREAD_MINB", thread ? min_blocks : 1),
                               def("PN_ADAPTIVE", adaptive ? 1 : 0), def("PN_DYNAMIC", dynamic ? 1 : 0), def("PN_SAVE", save),
                               def("PN_IOUP", (ioup && !gen) ? 1 : 0), def("PN_EK0_DENSE", ek0_dense ? 1 : 0)},
                         cubin);
}

/* general-prior path: IOUP with a rate matrix (vector rates arrive as diagonal matrices) or the Jacobian-updated rate */
static bool is_gen(const pnode_config* cfg) {
    return cfg->prior == PNODE_PRIOR_IOUP && (cfg->rate_matrix != nullptr || cfg->update_rate_parameter != 0);
}
static size_t gen_ws_doubles(int d, int q) { return (size_t)pn_gen_ws_doubles(d, q + 1); }

/* Mass matrix M (d x d row-major) -> generated preamble: PN_MASS, constexpr pn_mass(a,i) = M[a][i] and pn_mass_inv(a,i) =
   MM^-1[a][i] with MM = M (+ 1e-20 I if M has a zero diagonal entry; src/initialization/autodiffinit.jl:20-31).  The entries
   are compile-time constants of the NVRTC build, so the zeros of M cost nothing.  Empty string for M = identity / null. */
static int mass_preamble(const double* M, int d, std::string* out) {
    out->clear();
    if (!M) return 0;
    bool identity = true;
    for (int a = 0; a < d; a++)
        for (int i = 0; i < d; i++) {
            if (!std::isfinite(M[a * d + i])) return fail(PNODE_ERR_ARGUMENT, "mass_matrix has a non-finite entry");
            if (M[a * d + i] != (a == i ? 1.0 : 0.0)) identity = false;
        }
    if (identity) return 0;
    std::vector<double> A(M, M + (size_t)d * d), Inv((size_t)d * d, 0.0);
    bool zero_diag = false;
    for (int i = 0; i < d; i++) zero_diag |= (M[i * d + i] == 0.0);
    for (int i = 0; i < d; i++) {
        if (zero_diag) A[i * d + i] += 1e-20;
        Inv[i * d + i] = 1.0;
    }
    for (int k = 0; k < d; k++) { /* Gauss-Jordan with partial pivoting */
        int piv = k;
        for (int r = k + 1; r < d; r++)
            if (std::fabs(A[r * d + k]) > std::fabs(A[piv * d + k])) piv = r;
        if (A[piv * d + k] == 0.0)
            return fail(PNODE_ERR_ARGUMENT, "mass_matrix: M (+ 1e-20 I) is singular -- the initial derivatives are undefined "
                                            "(SingularException in the reference's initial conditioning)");
        if (piv != k)
            for (int c = 0; c < d; c++) {
                std::swap(A[k * d + c], A[piv * d + c]);
                std::swap(Inv[k * d + c], Inv[piv * d + c]);
            }
        const double pv = A[k * d + k];
        for (int c = 0; c < d; c++) {
            A[k * d + c] /= pv;
            Inv[k * d + c] /= pv;
        }
        for (int r = 0; r < d; r++) {
            if (r == k) continue;
            const double f = A[r * d + k];
            if (f == 0.0) continue;
            for (int c = 0; c < d; c++) {
                A[r * d + c] -= f * A[k * d + c];
                Inv[r * d + c] -= f * Inv[k * d + c];
            }
        }
    }
    auto table = [&](const char* name, const double* T) {
        std::string t = std::string("__host__ __device__ constexpr double ") + name + "(int a, int i) {\n  return ";
        char buf[96];
        for (int a = 0; a < d; a++)
            for (int i = 0; i < d; i++) {
                if (T[a * d + i] == 0.0) continue;
                snprintf(buf, sizeof buf, "(a == %d && i == %d) ? %.17g : ", a, i, T[a * d + i]);
                t += buf;
            }
        t += "0.0;\n}\n";
        return t;
    };
    *out = "#define PN_MASS 1\n" + table("pn_mass", M) + table("pn_mass_inv", Inv.data());
    return 0;
}

/* NVRTC: register-resident smoother sweep for (d, q, G); independent of the right-hand side. */
static int nvrtc_smoother_cubin(int d, int q, int G, bool ioup, bool col, std::vector<char>* cubin, bool second = false,
                                bool fenrir = false) {
    std::string src = col ? "#include \"kernels/pn_smooth_col.cuh\"\n" : "#include \"kernels/pn_smooth_reg.cuh\"\n";
    src += "extern \"C\" __global__ void __launch_bounds__(PN_SMR_THREADS, 1) pn_smooth_entry(const __grid_constant__ "
           "PnSmoothParams P) {\n  ";
    src += col ? "pn_smooth_col_entry" : "pn_smooth_reg_entry";
    src += "<PN_SD, PN_SQ, PN_SG>(P);\n}\n";
    auto def = [](const char* name, int v) { return std::string("-D") + name + "=" + std::to_string(v); };
    return nvrtc_compile(src, {def("PN_SD", d), def("PN_SQ", q), def("PN_SG", G), def("PN_IOUP", ioup ? 1 : 0), def("PN_SECOND", second ? 1 : 0),
                               def("PN_FENRIR", fenrir ? 1 : 0)}, cubin);
}

/* NVRTC: dense-output predict kernel for (d, q, G) */
static int nvrtc_dense_cubin(int d, int q, int G, bool ioup, std::vector<char>* cubin, bool second = false) {
    std::string src =
        "#include \"kernels/pn_dense.cuh\"\n"
        "extern \"C\" __global__ void __launch_bounds__(256, 1) pn_dense_entry(const __grid_constant__ PnDenseParams P) {\n"
        "  pn_dense_predict_entry<PN_SD, PN_SQ, PN_SG>(P);\n}\n";
    auto def = [](const char* name, int v) { return std::string("-D") + name + "=" + std::to_string(v); };
    return nvrtc_compile(src, {def("PN_SD", d), def("PN_SQ", q), def("PN_SG", G), def("PN_IOUP", ioup ? 1 : 0), def("PN_SECOND", second ? 1 : 0)}, cubin);
}

/* Observation model of a data set: H = observation_matrix * SolProj (o x D, fenrir.jl:85 / dataupdate.jl:60-61; SolProj = E0 rows, or
   (E1; E0) rows for second-order problems), the noise covariance R (o x o) and its upper Cholesky factor R.R (cov2psdmatrix). */
static int build_obs_model(const pnode_config* cfg, int D, std::vector<double>* H, std::vector<double>* Cv, std::vector<double>* Rn) {
    const int o = cfg->n_obs, dul = cfg->second_order ? 2 * cfg->d : cfg->d;
    H->assign((size_t)o * D, 0.0);
    Cv->assign((size_t)o * o, 0.0);
    Rn->assign((size_t)o * o, 0.0);
    for (int a = 0; a < o; a++)
        for (int j = 0; j < dul; j++) {
            const double pv = cfg->observation_matrix ? cfg->observation_matrix[a * dul + j] : (a == j ? 1.0 : 0.0);
            const int sc = cfg->second_order ? (j < cfg->d ? cfg->d + j : j - cfg->d) : j;
            (*H)[(size_t)a * D + sc] = pv;
        }
    if (cfg->observation_noise_cov) { /* cov2psdmatrix(::AbstractMatrix): upper Cholesky factor (ProbNumDiffEq.jl:72-73) */
        Cv->assign(cfg->observation_noise_cov, cfg->observation_noise_cov + (size_t)o * o);
        for (int j = 0; j < o; j++) {
            double dsum = (*Cv)[j * o + j];
            for (int k = 0; k < j; k++) dsum -= (*Rn)[k * o + j] * (*Rn)[k * o + j];
            if (!(dsum > 0.0)) return fail(PNODE_ERR_ARGUMENT, "observation_noise_cov is not positive definite"); /* PosDefException */
            (*Rn)[j * o + j] = std::sqrt(dsum);
            for (int i = j + 1; i < o; i++) {
                double v = (*Cv)[j * o + i];
                for (int k = 0; k < j; k++) v -= (*Rn)[k * o + j] * (*Rn)[k * o + i];
                (*Rn)[j * o + i] = v / (*Rn)[j * o + j];
            }
        }
    } else {
        for (int a = 0; a < o; a++) {
            (*Cv)[a * o + a] = cfg->observation_noise_var;
            (*Rn)[a * o + a] = std::sqrt(cfg->observation_noise_var);
        }
    }
    return 0;
}
/* `__host__ __device__ constexpr double name(int a, int i)` returning T[a][i] of a rows x cols table (zeros are left out, so
   they fold away in the unrolled kernels) */
static std::string constexpr_table(const char* name, const double* T, int rows, int cols) {
    std::string t = std::string("__host__ __device__ constexpr double ") + name + "(int a, int i) {\n  return ";
    char buf[96];
    for (int a = 0; a < rows; a++)
        for (int i = 0; i < cols; i++) {
            if (T[(size_t)a * cols + i] == 0.0) continue;
            snprintf(buf, sizeof buf, "(a == %d && i == %d) ? %.17g : ", a, i, T[(size_t)a * cols + i]);
            t += buf;
        }
    t += "0.0;\n}\n";
    return t;
}
/* forward data updates (data_forward = 1) -> generated preamble: PN_DATA, PN_DATA_O, pn_data_H, pn_data_cov, pn_data_rt */
static int data_preamble(const pnode_config* cfg, std::string* out) {
    out->clear();
    if (!(cfg->n_data > 0 && cfg->data_forward)) return 0;
    const int D = cfg->d * (cfg->q + 1);
    std::vector<double> H, Cv, Rn;
    int rc = build_obs_model(cfg, D, &H, &Cv, &Rn);
    if (rc) return rc;
    *out = "#define PN_DATA 1\n#define PN_DATA_O " + std::to_string(cfg->n_obs) + "\n" + constexpr_table("pn_data_H", H.data(), cfg->n_obs, D) +
           constexpr_table("pn_data_cov", Cv.data(), cfg->n_obs, cfg->n_obs) + constexpr_table("pn_data_rt", Rn.data(), cfg->n_obs, cfg->n_obs);
    return 0;
}

/* pn_observation_noise (d x d row-major covariance, or var * I) -> generated preamble: PN_OBSN, constexpr pn_obsn_cov(a,b) = R[a][b]
   and pn_obsn_rt(a,b) = R.R[a][b], the upper Cholesky factor (cov2psdmatrix: a Number / UniformScaling x -> sqrt(x) I, a matrix ->
   cholesky(M).U; src/caches.jl:175-178).  Empty string without noise. */
static bool has_obsn(const pnode_config* cfg) { return cfg->pn_observation_noise != nullptr || cfg->pn_observation_noise_var > 0.0; }
static int obsn_preamble(const pnode_config* cfg, std::string* out) {
    out->clear();
    if (!has_obsn(cfg)) return 0;
    const int d = cfg->d;
    std::vector<double> Cv((size_t)d * d, 0.0), U((size_t)d * d, 0.0);
    for (int a = 0; a < d; a++)
        for (int b = 0; b < d; b++)
            Cv[a * d + b] = cfg->pn_observation_noise ? cfg->pn_observation_noise[a * d + b] : (a == b ? cfg->pn_observation_noise_var : 0.0);
    for (int a = 0; a < d; a++)
        for (int b = 0; b < d; b++) {
            if (!std::isfinite(Cv[a * d + b])) return fail(PNODE_ERR_ARGUMENT, "pn_observation_noise has a non-finite entry");
            if (std::fabs(Cv[a * d + b] - Cv[b * d + a]) > 1e-12 * (std::fabs(Cv[a * d + b]) + std::fabs(Cv[b * d + a])))
                return fail(PNODE_ERR_ARGUMENT, "pn_observation_noise must be symmetric");
        }
    for (int j = 0; j < d; j++) { /* upper Cholesky Cv = U'U */
        double sacc = Cv[j * d + j];
        for (int k = 0; k < j; k++) sacc -= U[k * d + j] * U[k * d + j];
        if (!(sacc > 0.0)) return fail(PNODE_ERR_ARGUMENT, "pn_observation_noise must be positive definite (PosDefException in cov2psdmatrix)");
        U[j * d + j] = std::sqrt(sacc);
        for (int i = j + 1; i < d; i++) {
            double v = Cv[j * d + i];
            for (int k = 0; k < j; k++) v -= U[k * d + j] * U[k * d + i];
            U[j * d + i] = v / U[j * d + j];
        }
    }
    auto table = [&](const char* name, const double* T) {
        std::string t = std::string("__host__ __device__ constexpr double ") + name + "(int a, int i) {\n  return ";
        char buf[96];
        for (int a = 0; a < d; a++)
            for (int i = 0; i < d; i++) {
                if (T[a * d + i] == 0.0) continue;
                snprintf(buf, sizeof buf, "(a == %d && i == %d) ? %.17g : ", a, i, T[a * d + i]);
                t += buf;
            }
        t += "0.0;\n}\n";
        return t;
    };
    *out = "#define PN_OBSN 1\n" + table("pn_obsn_cov", Cv.data()) + table("pn_obsn_rt", U.data());
    return 0;
}

static size_t cta_smem_bytes(int d, int q, int n_p, bool second_order = false, bool mass = false) {
    return (size_t)pn_cta_sm_doubles(d, q + 1, second_order ? 2 * d : d, second_order ? 3 : 2, n_p > 0 ? n_p : 1, mass,
                                     (int)sizeof(PnCtaShared)) * sizeof(double);
}
/* threads per CTA and resident CTAs per SM of the CTA-per-trajectory kernel.  A step is a chain of short phases separated by
   barriers (4-column Cholesky panels, single-thread f / J), so several small CTAs per SM beat one large one wherever shared
   memory allows: measured on Pleiades as 14 second-order equations (D = 56, profiles/r2i_cta_variants.jsonl) 4 x 128 threads
   6.6e6 steps/s, 2 x 256 5.6e6, 1 x 512 3.1e6; at D = 112 only one CTA fits and 256 = 512 threads.  The ahead-of-time
   instantiations use the compiled-in defaults (256 threads, one CTA). */
static void cta_shape(size_t smem, int* threads, int* min_blocks) {
    const size_t budget = 227 * 1024, per = smem + 1024; /* + the per-CTA reservation */
    int nb = (4 * per <= budget) ? 4 : (2 * per <= budget) ? 2 : 1;
    int th = (nb == 4) ? 128 : 256;
    *threads = env_int("PNODE_CTA_THREADS", th);
    *min_blocks = env_int("PNODE_CTA_MIN_BLOCKS", nb);
    if (*threads * *min_blocks > 1024) *min_blocks = 1024 / *threads;
}

static int compile_nvrtc(pnode_solver* s, int save, Kernel* out) {
    int rc = load_driver();
    if (rc) return rc;
    std::vector<char> cubin;
    std::string src = s->rhs_src, name = s->rhs_name;
    if (s->builtin) { /* built-in problem without an ahead-of-time instantiation for this configuration */
        const char* sn = pn_builtin_struct_name(s->rhs_name.c_str());
        if (!sn) return fail(PNODE_ERR_ARGUMENT, "unknown built-in problem '" + s->rhs_name + "'");
        src = "#include \"kernels/pn_builtin_problems.cuh\"\n";
        name = sn;
    }
    out->threads = (s->kron || s->blocks) ? 128 : env_int("PNODE_THREADS", PN_THREADS);
    int min_blocks = env_int("PNODE_MIN_BLOCKS", 1);
    if (s->cta) {
        cta_shape(s->smem, &out->threads, &min_blocks);
        s->G = out->threads; /* one CTA per trajectory */
    }
    if (s->thread) {
        out->threads = 256;
        while (out->threads > 32 && thread_sm_doubles(s->cfg.d, s->cfg.q + 1) * out->threads * 8 > 227 * 1024) out->threads /= 2;
        s->smem = thread_sm_doubles(s->cfg.d, s->cfg.q + 1) * out->threads * 8;
        min_blocks = 1;
    } else if (s->gen) {
        out->threads = 32 * PN_GEN_WARPS;
    } else if (!s->cta && !s->kron && !s->blocks) {
        /* one exchange region per group: large states with few lanes per trajectory run smaller CTAs */
        while (out->threads > 32 && small_smem_bytes(s->D, s->G, out->threads) > 200 * 1024) out->threads /= 2;
        s->smem = small_smem_bytes(s->D, s->G, out->threads);
    }
    rc = nvrtc_cubin(src, name, s->cfg.q, s->G, s->cfg.adaptive != 0, is_dynamic(s->cfg.diffusion), save, &cubin, out->threads,
                     min_blocks, s->kron, s->cta,
                     s->blocks ? (s->cfg.alg == PNODE_DIAGONAL_EK1 ? 2 : 1) : 0, s->mv, s->cfg.prior == PNODE_PRIOR_IOUP,
                     s->cfg.alg == PNODE_EK0 && !s->kron && !s->blocks, s->mass_src, s->gen, s->thread);
    if (rc) return rc;
    CUresult r = g_drv.ModuleLoadData(&out->mod, cubin.data());
    if (r != CUDA_SUCCESS) return fail(PNODE_ERR_COMPILE, "cuModuleLoadData: " + cu_err(r));
    r = g_drv.ModuleGetFunction(&out->drv_fn, out->mod, "pn_entry");
    if (r != CUDA_SUCCESS) return fail(PNODE_ERR_COMPILE, "cuModuleGetFunction: " + cu_err(r));
    r = g_drv.ModuleGetFunction(&out->drv_init, out->mod, "pn_init");
    if (r != CUDA_SUCCESS) return fail(PNODE_ERR_COMPILE, "cuModuleGetFunction(pn_init): " + cu_err(r));
    if (s->smem > 48 * 1024) {
        r = g_drv.FuncSetAttribute(out->drv_fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)s->smem);
        if (r != CUDA_SUCCESS) return fail(PNODE_ERR_CUDA, "cuFuncSetAttribute(max dynamic smem): " + cu_err(r));
    }
    int nblk = 0;
    r = g_drv.Occupancy(&nblk, out->drv_fn, out->threads, s->smem);
    if (r != CUDA_SUCCESS || nblk < 1) return fail(PNODE_ERR_CUDA, "occupancy query failed: " + cu_err(r));
    out->blocks_per_sm = nblk;
    return 0;
}

static int get_kernel(pnode_solver* s, int save, Kernel* out) {
    if (out->valid()) return 0;
    auto t0 = std::chrono::steady_clock::now();
    if (s->builtin && !env_int("PNODE_FORCE_NVRTC", 0) && s->cfg.prior == PNODE_PRIOR_IWP && s->mass_src.empty() &&
        !(s->cfg.alg == PNODE_EK0 && !s->kron && !s->blocks) &&
        !(s->thread && thread_sm_doubles(s->cfg.d, s->cfg.q + 1) * PN_THREAD_NT * 8 > 227 * 1024) /* CTA size of the ahead-of-time build */) {
        int n = 0;
        const PnBuiltinEntry* tab = pn_builtin_table(&n);
        std::string key = kernel_key(s, save);
        for (int i = 0; i < n; i++) {
            if (key == tab[i].key) {
                out->rt_fn = tab[i].fn;
                out->rt_init = tab[i].init_fn;
                out->threads = s->cta ? PN_CTA_THREADS : ((s->kron || s->blocks) ? 128 : PN_THREADS);
                if (s->cta) s->G = out->threads;
                if (s->thread) {
                    out->threads = PN_THREAD_NT;
                    s->smem = thread_sm_doubles(s->cfg.d, s->cfg.q + 1) * PN_THREAD_NT * 8;
                } else
                if (!s->cta && !s->kron && !s->blocks) s->smem = small_smem_bytes(s->D, s->G, out->threads);
                if (s->smem > 48 * 1024)
                    CUDA_TRY(cudaFuncSetAttribute(out->rt_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem));
                int nblk = 0;
                CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nblk, out->rt_fn, out->threads, s->smem));
                if (nblk < 1) return fail(PNODE_ERR_CUDA, "kernel does not fit on an SM");
                out->blocks_per_sm = nblk;
                return 0;
            }
        }
    }
    if (s->builtin && !pn_builtin_struct_name(s->rhs_name.c_str()))
        return fail(PNODE_ERR_ARGUMENT, "unknown built-in problem '" + s->rhs_name + "'");
    int rc = compile_nvrtc(s, save, out);
    s->compile_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return rc;
}

/* register-resident smoother: ahead-of-time instantiation if there is one, else NVRTC */
static int get_smoother(pnode_solver* s) {
    Kernel* out = &s->k_smooth;
    if (out->valid()) return 0;
    s->Gs = env_int("PNODE_SMOOTH_LANES", smooth_lanes(s->D));
    if (s->Gs <= 0 || s->cta || s->gen || env_int("PNODE_SMOOTH_GENERIC", 0)) return 0; /* shared-memory sweep */
    /* PNODE_SMOOTH_KIND: 1 (default) column-distributed sweep (pn_smooth_col.cuh), 0 row-distributed (pn_smooth_reg.cuh) */
    const bool col = env_int("PNODE_SMOOTH_KIND", 1) != 0;
    s->smem_smooth = col ? smooth_col_smem_bytes(s->cfg.d, s->cfg.q, s->Gs) : smooth_reg_smem_bytes(s->cfg.d, s->cfg.q, s->Gs);
    if (s->smem_smooth > 227 * 1024) return 0;
    auto t0 = std::chrono::steady_clock::now();
    out->threads = PN_SMR_THREADS;
    const bool ioup = (s->cfg.prior == PNODE_PRIOR_IOUP);
    const void* fn = (env_int("PNODE_FORCE_NVRTC", 0) || ioup || s->cfg.second_order || s->fenrir()) ? nullptr : pn_smooth_reg_lookup(s->cfg.d, s->cfg.q, s->Gs, col ? 1 : 0);
    if (fn) {
        out->rt_fn = fn;
        CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_smooth));
        int nblk = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nblk, fn, out->threads, s->smem_smooth));
        if (nblk < 1) return fail(PNODE_ERR_CUDA, "smoother kernel does not fit on an SM");
        out->blocks_per_sm = nblk;
        return 0;
    }
    int rc = load_driver();
    if (rc) return rc;
    std::vector<char> cubin;
    if ((rc = nvrtc_smoother_cubin(s->cfg.d, s->cfg.q, s->Gs, ioup, col, &cubin, s->cfg.second_order != 0, s->fenrir()))) return rc;
    CUresult r = g_drv.ModuleLoadData(&out->mod, cubin.data());
    if (r != CUDA_SUCCESS) return fail(PNODE_ERR_COMPILE, "cuModuleLoadData: " + cu_err(r));
    r = g_drv.ModuleGetFunction(&out->drv_fn, out->mod, "pn_smooth_entry");
    if (r != CUDA_SUCCESS) return fail(PNODE_ERR_COMPILE, "cuModuleGetFunction: " + cu_err(r));
    r = g_drv.FuncSetAttribute(out->drv_fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)s->smem_smooth);
    if (r != CUDA_SUCCESS) return fail(PNODE_ERR_CUDA, "cuFuncSetAttribute(max dynamic smem): " + cu_err(r));
    int nblk = 0;
    r = g_drv.Occupancy(&nblk, out->drv_fn, out->threads, s->smem_smooth);
    if (r != CUDA_SUCCESS || nblk < 1) return fail(PNODE_ERR_CUDA, "occupancy query failed: " + cu_err(r));
    out->blocks_per_sm = nblk;
    s->compile_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}

static int get_dense(pnode_solver* s) {
    Kernel* out = &s->k_dense;
    if (out->valid()) return 0;
    auto t0 = std::chrono::steady_clock::now();
    if (s->D > 32) { /* warp-per-pair predict kernel (ahead-of-time, size-generic) */
        s->Gd = 32;
        out->threads = 32 * PN_SM_WARPS;
        out->rt_fn = pn_dense_gen_kernel_ptr();
        out->blocks_per_sm = 2;
        return 0;
    }
    s->Gd = auto_lanes_qr(s->D);
    out->threads = 256;
    int rc = load_driver();
    if (rc) return rc;
    std::vector<char> cubin;
    if ((rc = nvrtc_dense_cubin(s->cfg.d, s->cfg.q, s->Gd, s->cfg.prior == PNODE_PRIOR_IOUP, &cubin, s->cfg.second_order != 0))) return rc;
    CUresult r = g_drv.ModuleLoadData(&out->mod, cubin.data());
    if (r != CUDA_SUCCESS) return fail(PNODE_ERR_COMPILE, "cuModuleLoadData: " + cu_err(r));
    r = g_drv.ModuleGetFunction(&out->drv_fn, out->mod, "pn_dense_entry");
    if (r != CUDA_SUCCESS) return fail(PNODE_ERR_COMPILE, "cuModuleGetFunction: " + cu_err(r));
    int nblk = 0;
    r = g_drv.Occupancy(&nblk, out->drv_fn, out->threads, 0);
    if (r != CUDA_SUCCESS || nblk < 1) return fail(PNODE_ERR_CUDA, "occupancy query failed: " + cu_err(r));
    out->blocks_per_sm = nblk;
    s->compile_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}

/* SAVE template level of the step kernel: 0 final values only, 1 per-step sol.t / sol.u / sol.pu, 2 additionally the filter
   state per step (what the smoother sweep, the dense output and save_state = sol.x_filt need) */
static int save_level(const pnode_config* cfg, bool cta) {
    if (!cfg->save_everystep) return 0;
    return (cfg->smooth || cfg->n_saveat > 0 || (cfg->save_state && !cta)) ? 2 : 1;
}

/* covariance_structure(alg, prior, diffusionmodel) (src/algorithms.jl:132-151) */
static int resolve_layout(const pnode_config* cfg) {
    if (cfg->alg == PNODE_EK0 && cfg->prior != PNODE_PRIOR_IWP) return PNODE_COV_DENSE; /* "other priors": dense */
    if (cfg->alg == PNODE_EK0) return is_mv(cfg->diffusion) ? PNODE_COV_BLOCKS : PNODE_COV_KRONECKER;
    if (cfg->alg == PNODE_DIAGONAL_EK1) return PNODE_COV_BLOCKS;
    return PNODE_COV_DENSE;
}

/* M == I (or no mass matrix at all): the ODE kernels apply */
static bool mass_is_identity(const pnode_config* cfg) {
    if (!cfg->mass_matrix) return true;
    for (int a = 0; a < cfg->d; a++)
        for (int i = 0; i < cfg->d; i++)
            if (cfg->mass_matrix[a * cfg->d + i] != (a == i ? 1.0 : 0.0)) return false;
    return true;
}
/* pn_thread.cuh: EK1, IWP, plain first-order ODE, filter only, small state */
static bool thread_eligible(const pnode_config* cfg) {
    const int D = cfg->d * (cfg->q + 1);
    return cfg->alg == PNODE_EK1 && cfg->prior == PNODE_PRIOR_IWP && !cfg->save_everystep && mass_is_identity(cfg) && !cfg->second_order &&
           cfg->n_data == 0 && !has_obsn(cfg) && cfg->q >= 1 && D <= 16 && (cfg->lanes_per_traj == 0 || cfg->lanes_per_traj == 1);
}

static bool use_thread(const pnode_config* cfg) {
    return thread_eligible(cfg) && (cfg->lanes_per_traj == 1 || (env_int("PNODE_THREAD_KERNEL", 1) != 0 && cfg->d * (cfg->q + 1) <= 12));
}

static int validate_config(const pnode_config* cfg) {
    if (cfg->struct_size != (int32_t)sizeof(pnode_config))
        return fail(PNODE_ERR_ARGUMENT, "pnode_config.struct_size mismatch (ABI)");
    /* ---- argument checks mirroring the reference's constructors ---- */
    if (cfg->d < 1) return fail(PNODE_ERR_ARGUMENT, "We currently don't support scalar-valued problems (d >= 1 vector required)"); /* caches.jl:96-98 */
    if (cfg->q < 1 || cfg->q + 1 > PN_MAX_N) return fail(PNODE_ERR_ARGUMENT, "order must satisfy 1 <= order <= 11");
    if (cfg->prior != PNODE_PRIOR_IWP && cfg->prior != PNODE_PRIOR_IOUP)
        return fail(PNODE_ERR_UNSUPPORTED, "only the IWP and IOUP priors are implemented");
    if (cfg->prior == PNODE_PRIOR_IOUP) {
        if (cfg->alg == PNODE_DIAGONAL_EK1 || is_mv(cfg->diffusion))
            return fail(PNODE_ERR_UNSUPPORTED, "the IOUP prior is implemented on the dense layout (EK0 / EK1 with a scalar diffusion)");
        if (cfg->update_rate_parameter && cfg->rate_matrix)
            return fail(PNODE_ERR_ARGUMENT, "Do not manually set the `rate_parameter` of the IOUP prior when using the "
                                            "`update_rate_parameter=true` option.Reset the prior and try again."); /* caches.jl:139-149 */
        if (cfg->update_rate_parameter && cfg->alg != PNODE_EK1)
            return fail(PNODE_ERR_UNSUPPORTED, "update_rate_parameter takes the rate from the traced Jacobian: EK1 only (RosenbrockExpEK)");
        if (is_gen(cfg)) {
            /* general-prior path: one warp per trajectory, dense transition pairs */
            if (cfg->rate_matrix)
                for (int i = 0; i < cfg->d * cfg->d; i++)
                    if (!std::isfinite(cfg->rate_matrix[i])) return fail(PNODE_ERR_ARGUMENT, "rate_parameter has a non-finite entry");
            if (cfg->d * (cfg->q + 1) > 64)
                return fail(PNODE_ERR_UNSUPPORTED, "IOUP priors with a vector / matrix / updated rate are implemented for D = d(q+1) <= 64");
            if (cfg->lanes_per_traj != 0 && cfg->lanes_per_traj != 32)
                return fail(PNODE_ERR_ARGUMENT, "IOUP priors with a vector / matrix / updated rate run one warp per trajectory (lanes_per_traj 0 or 32)");
            if (cfg->n_saveat > 0) return fail(PNODE_ERR_UNSUPPORTED, "dense output is not implemented for IOUP priors with a vector / matrix / updated rate");
            if (cfg->mass_matrix || cfg->second_order || cfg->n_data > 0 || has_obsn(cfg))
                return fail(PNODE_ERR_UNSUPPORTED, "IOUP priors with a vector / matrix / updated rate: plain first-order ODEs without mass matrix, data or observation noise");
        } else {
            if (cfg->d * (cfg->q + 1) > 24)
                return fail(PNODE_ERR_UNSUPPORTED, "the scalar-rate IOUP prior is implemented for D = d(q+1) <= 24 on the register-resident kernels and, for plain first-order problems without dense output, up to D = 64 on the general-prior path");
            if (cfg->lanes_per_traj == 256) return fail(PNODE_ERR_UNSUPPORTED, "the IOUP prior is not implemented in the CTA kernel");
        }
    } else if (cfg->rate_matrix || cfg->update_rate_parameter) {
        return fail(PNODE_ERR_ARGUMENT, "rate_matrix / update_rate_parameter belong to the IOUP prior");
    }
    if (cfg->alg != PNODE_EK0 && cfg->alg != PNODE_EK1 && cfg->alg != PNODE_DIAGONAL_EK1)
        return fail(PNODE_ERR_ARGUMENT, "Unknown algorithm"); /* derivative_utils.jl:31 */
    if (cfg->diffusion < PNODE_DIFF_DYNAMIC || cfg->diffusion > PNODE_DIFF_FIXED_MV)
        return fail(PNODE_ERR_ARGUMENT, "Unknown diffusion model");
    if (cfg->alg == PNODE_EK1) { /* ekargcheck, src/algorithms.jl:93-107 */
        if (cfg->diffusion == PNODE_DIFF_FIXED_MV && cfg->calibrate)
            return fail(PNODE_ERR_ARGUMENT,
                        "The `EK1` algorithm does not support automatic global calibration of multivariate diffusion models. "
                        "Either use a scalar diffusion model, or set `calibrate=false` and calibrate manually by optimizing "
                        "`sol.pnstats.log_likelihood`. Or use a different solve, like `EK0` or `DiagonalEK1`.");
        if (cfg->diffusion == PNODE_DIFF_DYNAMIC_MV)
            return fail(PNODE_ERR_ARGUMENT,
                        "The `EK1` algorithm does not support automatic calibration of local multivariate diffusion models. "
                        "Either use a scalar diffusion model, or use a different solve, like `EK0` or `DiagonalEK1`.");
        if (cfg->diffusion == PNODE_DIFF_FIXED_MV)
            return fail(PNODE_ERR_UNSUPPORTED, "EK1 with an uncalibrated multivariate diffusion is not implemented (dense layout)");
    }
    if (!cfg->adaptive && !(cfg->dt > 0.0))
        return fail(PNODE_ERR_ARGUMENT, "Fixed timestep methods require a choice of dt or choosing the tstops"); /* test/errors_thrown.jl:6-8 */
    if (cfg->smooth && !cfg->save_everystep)
        return fail(PNODE_ERR_ARGUMENT, "If you set `save_everystep=false` also set `smooth=false` in the alg!"); /* checks.jl:18-20 */
    if (!(cfg->t1 > cfg->t0)) return fail(PNODE_ERR_ARGUMENT, "tspan must satisfy t1 > t0");
    if (!cfg->rhs_name) return fail(PNODE_ERR_ARGUMENT, "rhs_name is required");
    const int D = cfg->d * (cfg->q + 1);
    const int layout = resolve_layout(cfg);
    if (cfg->cov == PNODE_COV_KRONECKER && cfg->alg != PNODE_EK0)
        return fail(PNODE_ERR_ARGUMENT, "IsometricKroneckerCovariance requires the EK0"); /* algorithms.jl:101-110 */
    if (cfg->cov != PNODE_COV_AUTO && cfg->cov != layout)
        return fail(PNODE_ERR_UNSUPPORTED,
                    "only the covariance factorization the reference selects for this algorithm / diffusion model is implemented "
                    "(covariance_structure, src/algorithms.jl:132-151)");
    /* pn_kron.cuh keeps the n x n factor and the n x d mean in one thread; beyond order 7 they no longer fit the register file and the
       compiler places them in thread-local memory (L1-resident): slower per step, same arithmetic (test/correctness.jl runs EK0(order=8)) */
    if (layout == PNODE_COV_BLOCKS) {
        /* pn_blocks.cuh keeps the d factors of n x n in one thread: registers up to about 160 doubles, thread-local memory beyond (e.g.
           Pleiades with the DiagonalEK1: 448 doubles); the bound keeps compile time and the local-memory footprint of a launch in check */
        if (cfg->d * (cfg->q + 1) * (cfg->q + 1) > 512)
            return fail(PNODE_ERR_UNSUPPORTED, "the blocks-of-diagonals thread-per-trajectory kernel is built for d (q+1)^2 <= 512 doubles of factors per trajectory");
        if (cfg->n_forced > 0) return fail(PNODE_ERR_UNSUPPORTED, "forced_diffusion is a scalar-diffusion test hook");
        if (cfg->smooth && smooth_lanes(D) == 0)
            return fail(PNODE_ERR_UNSUPPORTED, "smoothing on the blocks-of-diagonals layout needs the register-resident sweep (D <= 24)");
    }
    if (cfg->lanes_per_traj != 0 && cfg->lanes_per_traj != 1 && cfg->lanes_per_traj != 2 && cfg->lanes_per_traj != 4 && cfg->lanes_per_traj != 8 &&
        cfg->lanes_per_traj != 16 && cfg->lanes_per_traj != 32 && cfg->lanes_per_traj != 256)
        return fail(PNODE_ERR_ARGUMENT, "lanes_per_traj must be one of 0 (auto), 1 (thread per trajectory), 2, 4, 8, 16, 32, 256 (CTA per trajectory)");
    if (cfg->lanes_per_traj == 1 && !thread_eligible(cfg))
        return fail(PNODE_ERR_UNSUPPORTED, "lanes_per_traj = 1 (thread per trajectory) serves the EK1 with the IWP prior on first-order problems, "
                                           "filter only (save_everystep = 0), D = d(q+1) <= 16");
    if (!is_gen(cfg) && cfg->alg == PNODE_EK1 && (D > 32 || cfg->lanes_per_traj == 256)) {
        if (cta_smem_bytes(cfg->d, cfg->q, cfg->n_p, cfg->second_order != 0, cfg->mass_matrix != nullptr) > 227 * 1024)
            return fail(PNODE_ERR_UNSUPPORTED, "state dimension too large for the shared-memory CTA kernel");
        if (cfg->lanes_per_traj != 0 && cfg->lanes_per_traj != 256)
            return fail(PNODE_ERR_ARGUMENT, "D = d(q+1) > 32 runs one CTA per trajectory (lanes_per_traj 0 or 256)");
        if (cfg->smooth && cfg->prior != PNODE_PRIOR_IWP)
            return fail(PNODE_ERR_UNSUPPORTED, "smoothing with the CTA-per-trajectory kernel implements the IWP prior");
        if (cfg->smooth && cfg->second_order)
            return fail(PNODE_ERR_UNSUPPORTED, "smoothing of second-order problems needs the column-distributed register sweep (D <= 24)");
    }
    if (cfg->n_saveat > 0) { /* dense output sol(saveat) */
        if (!cfg->saveat) return fail(PNODE_ERR_ARGUMENT, "saveat is null but n_saveat > 0");
        if (!cfg->save_everystep) return fail(PNODE_ERR_ARGUMENT, "saveat (dense output) needs save_everystep = true");
        for (int64_t i = 0; i < cfg->n_saveat; i++) {
            if (cfg->saveat[i] < cfg->t0) return fail(PNODE_ERR_ARGUMENT, "Invalid t<t0"); /* src/solution.jl:340-342 */
            if (i > 0 && !(cfg->saveat[i] >= cfg->saveat[i - 1])) return fail(PNODE_ERR_ARGUMENT, "saveat must be ascending");
        }
        if (D > 32 || (cfg->smooth && smooth_lanes(D) == 0)) {
            /* beyond the register kernels: warp-per-pair predict (pn_dense_gen.cuh) + the shared-memory sweep */
            if (cfg->prior != PNODE_PRIOR_IWP || cfg->second_order || is_mv(cfg->diffusion) || layout != PNODE_COV_DENSE)
                return fail(PNODE_ERR_UNSUPPORTED, "dense output for D = d(q+1) > 32 (smoothed: > 24) is implemented for first-order problems, the IWP prior "
                                                   "and scalar diffusion models on the dense layout");
        }
    }
    if (cfg->n_data > 0) { /* fenrir_data_loglik (backward pass), or DataUpdateCallback (forward updates) */
        if (!cfg->data_forward && !cfg->smooth)
            return fail(PNODE_ERR_ARGUMENT, "fenrir only works with smoothing. Set `smooth=true`."); /* fenrir.jl:42-44 */
        if (!cfg->data_t || !cfg->data_u || cfg->n_obs < 1) return fail(PNODE_ERR_ARGUMENT, "data_t / data_u / n_obs are required with n_data > 0");
        const int dul = cfg->second_order ? 2 * cfg->d : cfg->d;
        if (!cfg->observation_matrix && cfg->n_obs != dul)
            return fail(PNODE_ERR_ARGUMENT, "without an observation_matrix every observation must have length(sol.u) entries"); /* DimensionMismatch */
        if (cfg->n_obs > D) return fail(PNODE_ERR_ARGUMENT, "n_obs must not exceed the state dimension");
        for (int64_t i = 0; i < cfg->n_data; i++) {
            if (!(cfg->data_t[i] >= cfg->t0 && cfg->data_t[i] <= cfg->t1)) return fail(PNODE_ERR_ARGUMENT, "data.t must lie in tspan");
            if (i > 0 && !(cfg->data_t[i] > cfg->data_t[i - 1])) return fail(PNODE_ERR_ARGUMENT, "data.t must be strictly ascending");
        }
        if (!cfg->observation_noise_cov && !(cfg->observation_noise_var > 0.0))
            return fail(PNODE_ERR_ARGUMENT, "observation_noise_cov must be positive (definite)");
        if (cfg->data_forward) {
            /* dataupdate.jl:69-71: "Partial observations only work with the EK1 right now"; here the forward updates are built into
               the dense-layout step kernel */
            if (layout != PNODE_COV_DENSE) {
                if (cfg->n_obs != dul) return fail(PNODE_ERR_ARGUMENT, "Partial observations only work with the EK1 right now");
                return fail(PNODE_ERR_UNSUPPORTED, "forward data updates (DataUpdateCallback) are implemented for the EK1 on the dense layout; "
                                                   "fenrir_data_loglik serves the EK0 and the DiagonalEK1");
            }
            if (D > 32 || cfg->lanes_per_traj == 256)
                return fail(PNODE_ERR_UNSUPPORTED, "forward data updates are implemented for D = d(q+1) <= 32");
            if (cfg->prior != PNODE_PRIOR_IWP) return fail(PNODE_ERR_UNSUPPORTED, "forward data updates are implemented for the IWP prior");
            std::vector<double> Hh, Cc, Rr;
            if (int orc = build_obs_model(cfg, D, &Hh, &Cc, &Rr)) return orc;
        } else if (cfg->prior != PNODE_PRIOR_IWP || is_mv(cfg->diffusion))
            return fail(PNODE_ERR_UNSUPPORTED, "the Fenrir backward pass is implemented for the IWP prior with scalar diffusion models");
        if (!cfg->data_forward && layout != PNODE_COV_DENSE) {
            /* the structured layouts carry the data update on their factors: full observations only
               (src/callbacks/dataupdate.jl:69-71) and a noise covariance of the layout's own shape
               (to_factorized_matrix, src/covariance_structure.jl:46-58: isotropic for Kronecker, diagonal for blocks) */
            bool ident = (cfg->n_obs == dul);
            if (ident && cfg->observation_matrix)
                for (int a = 0; a < dul; a++)
                    for (int j = 0; j < dul; j++) ident = ident && (cfg->observation_matrix[a * dul + j] == (a == j ? 1.0 : 0.0));
            if (!ident) return fail(PNODE_ERR_ARGUMENT, "Partial observations only work with the EK1 right now");
            if (cfg->observation_noise_cov) {
                const double* C = cfg->observation_noise_cov;
                bool diag = true, iso = true;
                for (int a = 0; a < dul; a++)
                    for (int j = 0; j < dul; j++) {
                        if (a != j && C[a * dul + j] != 0.0) diag = false;
                        if (a == j && C[a * dul + j] != C[0]) iso = false;
                    }
                if (!diag || (layout == PNODE_COV_KRONECKER && !iso))
                    return fail(PNODE_ERR_ARGUMENT,
                                layout == PNODE_COV_KRONECKER
                                    ? "observation_noise_cov must be a scalar or a multiple of the identity for the EK0 (Kronecker-factorized covariances); use the EK1 for a matrix-valued noise covariance"
                                    : "observation_noise_cov must be a scalar or a diagonal matrix for the block-diagonal covariance layout; use the EK1 for a full noise covariance");
            }
        }
        if (!cfg->data_forward && cfg->n_saveat > 0)
            return fail(PNODE_ERR_UNSUPPORTED, "saveat and a Fenrir data set cannot be combined (the records stay filtering estimates)");
        if (!cfg->data_forward && D > 32) return fail(PNODE_ERR_UNSUPPORTED, "the Fenrir backward pass is implemented for D = d(q+1) <= 32");
    }
    if (cfg->second_order) { /* SecondOrderODEProblem */
        if (cfg->q < 2) return fail(PNODE_ERR_ARGUMENT, "second-order problems need order >= 2 (the measurement reads the second derivative)");
        if (cfg->mass_matrix) return fail(PNODE_ERR_ARGUMENT, "second-order problems take no mass matrix (derivative_utils.jl:39-41 asserts mass_matrix === I)");
        if (cfg->prior != PNODE_PRIOR_IWP) return fail(PNODE_ERR_UNSUPPORTED, "second-order problems are implemented for the IWP prior");
        if (layout == PNODE_COV_DENSE && cfg->alg != PNODE_EK1 && (D > 32 || cfg->lanes_per_traj == 256))
            return fail(PNODE_ERR_UNSUPPORTED, "second-order problems with the EK0 on the dense layout are implemented for D = d(q+1) <= 32");
        if (cfg->smooth && (smooth_lanes(D) == 0 || env_int("PNODE_SMOOTH_KIND", 1) == 0 || env_int("PNODE_SMOOTH_GENERIC", 0)))
            return fail(PNODE_ERR_UNSUPPORTED, "smoothing of second-order problems needs the column-distributed register sweep (D <= 24)");
    }
    if (cfg->mass_matrix) {
        std::string pre;
        int mrc = mass_preamble(cfg->mass_matrix, cfg->d, &pre);
        if (mrc) return mrc;
        if (!pre.empty()) { /* M != I */
            const int d = cfg->d;
            const double* M = cfg->mass_matrix;
            bool diagonal = true, uniform = true, zero_on_diag = false;
            for (int a = 0; a < d; a++)
                for (int i = 0; i < d; i++) {
                    if (a != i && M[a * d + i] != 0.0) diagonal = false;
                    if (a == i && M[a * d + i] != M[0]) uniform = false;
                    if (a == i && M[a * d + i] == 0.0) zero_on_diag = true;
                }
            uniform = uniform && diagonal;
            if (layout == PNODE_COV_KRONECKER) {
                /* src/caches.jl:111-118: only a UniformScaling passes (H = lambda E1, derivative_utils.jl:45-47) */
                if (!uniform || zero_on_diag)
                    return fail(PNODE_ERR_ARGUMENT,
                                "The selected algorithm uses an efficient Kronecker-factorized implementation which is incompatible "
                                "with the provided mass matrix. Try using the `EK1` instead.");
            } else if (layout == PNODE_COV_BLOCKS) {
                if (!diagonal)
                    return fail(PNODE_ERR_UNSUPPORTED, "the blocks-of-diagonals layout (DiagonalEK1, EK0 with multivariate diffusion) takes diagonal mass matrices only");
                if (cfg->alg == PNODE_EK0 && zero_on_diag)
                    return fail(PNODE_ERR_ARGUMENT, "the EK0 cannot solve problems with a singular mass matrix (H_i = M_ii e_1' vanishes); use the EK1 or the DiagonalEK1");
            } else {
                if (cfg->alg != PNODE_EK1)
                    return fail(PNODE_ERR_UNSUPPORTED, "mass matrices on the dense layout are implemented for the EK1");
            }
        }
    }
    if (cfg->pn_observation_noise_var < 0.0) return fail(PNODE_ERR_ARGUMENT, "pn_observation_noise must be positive (semi)definite");
    if (has_obsn(cfg)) { /* ekargcheck (src/algorithms.jl:78-130) */
        if (!is_dynamic(cfg->diffusion) && cfg->calibrate)
            return fail(PNODE_ERR_ARGUMENT,
                        "Automatic calibration of global diffusion models is not possible when using observation noise. If you want to "
                        "calibrate a global diffusion parameter, do so setting `calibrate=false` and optimizing `sol.pnstats.log_likelihood` manually.");
        if (cfg->pn_observation_noise) {
            const int d = cfg->d;
            const double* Cn = cfg->pn_observation_noise;
            bool diag = true, iso = true;
            for (int a = 0; a < d; a++)
                for (int j = 0; j < d; j++) {
                    if (a != j && Cn[a * d + j] != 0.0) diag = false;
                    if (a == j && Cn[a * d + j] != Cn[0]) iso = false;
                }
            if (layout == PNODE_COV_KRONECKER && !(diag && iso))
                return fail(PNODE_ERR_ARGUMENT,
                            "The supplied `pn_observation_noise` is not compatible with the chosen `IsometricKroneckerCovariance` "
                            "factorization. Try one of `BlockDiagonalCovariance` or `DenseCovariance` instead!");
            if (layout == PNODE_COV_BLOCKS && !diag)
                return fail(PNODE_ERR_ARGUMENT,
                            "The supplied `pn_observation_noise` is not compatible with the chosen `BlockDiagonalCovariance` factorization. "
                            "Try `DenseCovariance` instead!");
        }
        std::string pre;
        int orc = obsn_preamble(cfg, &pre);
        if (orc) return orc;
        if (layout == PNODE_COV_DENSE && (D > 32 || cfg->lanes_per_traj == 256))
            return fail(PNODE_ERR_UNSUPPORTED, "pn_observation_noise is implemented for D = d(q+1) <= 32 on the dense layout");
        if (cfg->n_data > 0 && !cfg->data_forward)
            return fail(PNODE_ERR_UNSUPPORTED, "pn_observation_noise and a Fenrir data set cannot be combined yet");
    }
    const bool builtin = (std::string(cfg->rhs_name).rfind("builtin:", 0) == 0);
    if (!builtin && !cfg->rhs_src) return fail(PNODE_ERR_ARGUMENT, "rhs_src is required unless rhs_name is a builtin");
    return 0;
}

/* A scalar-rate IOUP beyond the register-resident kernels (24 < D <= 64): IOUP(q, r) is the SDE of IOUP(q, r I) (test/ioup.jl:48-55 runs
   the four spellings against the same threshold), so the configuration is handed to the general-prior path with the rate matrix r I.
   Only where that path serves the rest of the configuration; otherwise the original configuration goes on and validate_config names
   what is missing. */
struct RatePromotion {
    pnode_config cfg;
    std::vector<double> iso;
};
static const pnode_config* promote_scalar_rate(const pnode_config* cfg, RatePromotion* w) {
    if (!cfg || cfg->struct_size != (int32_t)sizeof(pnode_config)) return cfg;
    if (cfg->prior != PNODE_PRIOR_IOUP || cfg->rate_matrix || cfg->update_rate_parameter || cfg->d < 1 || cfg->q < 1) return cfg;
    const int D = cfg->d * (cfg->q + 1);
    if (D <= 24 || D > 64) return cfg;
    if (cfg->n_saveat > 0 || cfg->mass_matrix || cfg->second_order || cfg->n_data > 0 || has_obsn(cfg) ||
        (cfg->lanes_per_traj != 0 && cfg->lanes_per_traj != 32))
        return cfg;
    w->cfg = *cfg;
    w->iso.assign((size_t)cfg->d * cfg->d, 0.0);
    for (int i = 0; i < cfg->d; i++) w->iso[(size_t)i * cfg->d + i] = cfg->rate_parameter;
    w->cfg.rate_matrix = w->iso.data();
    return &w->cfg;
}

static int compile_check_impl(const pnode_config* cfg, int64_t* cubin_bytes);
/* Validate + NVRTC-compile only (no device needed).  Used by host-side tests and to pre-warm caches. */
extern "C" int pnode_compile_check(const pnode_config* cfg, int64_t* cubin_bytes) {
    RatePromotion promoted;
    return compile_check_impl(promote_scalar_rate(cfg, &promoted), cubin_bytes);
}
static int compile_check_impl(const pnode_config* cfg, int64_t* cubin_bytes) {
    if (!cfg) return fail(PNODE_ERR_ARGUMENT, "null argument");
    int rc = validate_config(cfg);
    if (rc) return rc;
    if (cubin_bytes) *cubin_bytes = 0;
    if (std::string(cfg->rhs_name).rfind("builtin:", 0) == 0) return 0;
    const int D = cfg->d * (cfg->q + 1);
    const int G = cfg->lanes_per_traj > 0 ? cfg->lanes_per_traj : auto_lanes(D);
    std::vector<char> cubin;
    const int layout = resolve_layout(cfg);
    std::string mass_src;
    if ((rc = mass_preamble(cfg->mass_matrix, cfg->d, &mass_src))) return rc;
    if (cfg->second_order) mass_src = "#define PN_SECOND 1\n" + mass_src;
    {
        std::string obs_src, data_src;
        if ((rc = obsn_preamble(cfg, &obs_src))) return rc;
        if ((rc = data_preamble(cfg, &data_src))) return rc;
        mass_src = obs_src + data_src + mass_src;
    }
    rc = nvrtc_cubin(cfg->rhs_src, cfg->rhs_name, cfg->q, G, cfg->adaptive != 0, is_dynamic(cfg->diffusion),
                     save_level(cfg, cfg->alg == PNODE_EK1 && (D > 32 || cfg->lanes_per_traj == 256)), &cubin, layout == PNODE_COV_DENSE ? PN_THREADS : 128, 1,
                     layout == PNODE_COV_KRONECKER, cfg->alg == PNODE_EK1 && (D > 32 || cfg->lanes_per_traj == 256),
                     layout == PNODE_COV_BLOCKS ? (cfg->alg == PNODE_DIAGONAL_EK1 ? 2 : 1) : 0, is_mv(cfg->diffusion),
                     cfg->prior == PNODE_PRIOR_IOUP, cfg->alg == PNODE_EK0 && layout == PNODE_COV_DENSE, mass_src, is_gen(cfg), use_thread(cfg));
    if (rc) return rc;
    int64_t total = (int64_t)cubin.size();
    if (is_gen(cfg)) { /* the smoother of this path is the ahead-of-time shared-memory sweep */
        if (cubin_bytes) *cubin_bytes = total;
        return 0;
    }
    /* the smoother sweep and the dense-output kernel of this configuration compile from the same embedded headers */
    if (cfg->smooth && smooth_lanes(D) > 0 && !(cfg->alg == PNODE_EK1 && D > 32)) {
        std::vector<char> c2;
        const bool col = env_int("PNODE_SMOOTH_KIND", 1) != 0;
        if ((rc = nvrtc_smoother_cubin(cfg->d, cfg->q, smooth_lanes(D), cfg->prior == PNODE_PRIOR_IOUP, col, &c2, cfg->second_order != 0, cfg->n_data > 0 && !cfg->data_forward))) return rc;
        total += (int64_t)c2.size();
    }
    if (cfg->n_saveat > 0) {
        std::vector<char> c3;
        if ((rc = nvrtc_dense_cubin(cfg->d, cfg->q, auto_lanes_qr(D), cfg->prior == PNODE_PRIOR_IOUP, &c3, cfg->second_order != 0))) return rc;
        total += (int64_t)c3.size();
    }
    if (cubin_bytes) *cubin_bytes = total;
    return 0;
}

static int create_impl(const pnode_config* cfg, pnode_solver** out);
extern "C" int pnode_create(const pnode_config* cfg, pnode_solver** out) {
    RatePromotion promoted; /* the rate matrix is copied to the device inside create_impl: nothing of `promoted` outlives the call */
    return create_impl(promote_scalar_rate(cfg, &promoted), out);
}
static int create_impl(const pnode_config* cfg, pnode_solver** out) {
    if (!cfg || !out) return fail(PNODE_ERR_ARGUMENT, "null argument");
    *out = nullptr;
    {
        int vrc = validate_config(cfg);
        if (vrc) return vrc;
    }
    const int D = cfg->d * (cfg->q + 1);

    pnode_solver* s = new pnode_solver();
    s->cfg = *cfg;
    s->D = D;
    s->rhs_name = cfg->rhs_name;
    s->builtin = (s->rhs_name.rfind("builtin:", 0) == 0);
    if (!s->builtin) s->rhs_src = cfg->rhs_src;
    s->cfg.rhs_src = nullptr;
    s->cfg.rhs_name = nullptr;
    mass_preamble(cfg->mass_matrix, cfg->d, &s->mass_src); /* validated above */
    if (cfg->second_order) s->mass_src = "#define PN_SECOND 1\n" + s->mass_src; /* generated preamble of the step kernels */
    {
        std::string obs_src, data_src;
        obsn_preamble(cfg, &obs_src); /* validated above */
        data_preamble(cfg, &data_src);
        s->mass_src = obs_src + data_src + s->mass_src;
        s->data_forward = (cfg->n_data > 0 && cfg->data_forward);
    }
    s->cfg.mass_matrix = nullptr;
    s->cfg.pn_observation_noise = nullptr;
    s->kron = (resolve_layout(cfg) == PNODE_COV_KRONECKER);
    s->blocks = (resolve_layout(cfg) == PNODE_COV_BLOCKS);
    s->mv = is_mv(cfg->diffusion);
    s->gen = is_gen(cfg);
    /* thread per trajectory: on request (lanes_per_traj = 1), or automatically where it is the faster mapping (PNODE_THREAD_KERNEL=0: off) */
    s->thread = use_thread(cfg);
    s->cta = !s->gen && !s->kron && !s->blocks && (D > 32 || cfg->lanes_per_traj == 256);
    s->G = (s->kron || s->blocks || s->thread) ? 1 : (s->gen ? 32 : (s->cta ? 256 : (cfg->lanes_per_traj > 0 ? cfg->lanes_per_traj : auto_lanes(D))));
    if (s->gen) {
        s->gen_ws_doubles = gen_ws_doubles(cfg->d, cfg->q);
        const size_t bytes = s->gen_ws_doubles * PN_GEN_WARPS * sizeof(double);
        s->smem = bytes <= 200 * 1024 ? bytes : 0; /* 0: global workspace (allocated below) */
    } else
    s->smem = s->cta ? cta_smem_bytes(cfg->d, cfg->q, cfg->n_p, cfg->second_order != 0, !s->mass_src.empty() && s->mass_src.find("PN_MASS") != std::string::npos) : 0; /* group-per-trajectory kernel: set with its thread count in get_kernel */
    /* ---- constants ---- */
    PnParams& B = s->base;
    memset(&B, 0, sizeof B);
    B.traj_stride = 1;
    B.t0 = cfg->t0;
    B.t1 = cfg->t1;
    B.dt0 = cfg->dt;
    B.abstol = cfg->abstol;
    B.reltol = cfg->reltol;
    B.dtmin = cfg->dtmin;
    B.dtmax = cfg->dtmax > 0.0 ? cfg->dtmax : (cfg->t1 - cfg->t0);
    B.beta2 = cfg->beta2 > 0.0 ? cfg->beta2 : 2.0 / (5.0 * cfg->q);
    B.beta1 = cfg->beta1 > 0.0 ? cfg->beta1 : 7.0 / (10.0 * cfg->q);
    B.qmin = cfg->qmin > 0.0 ? cfg->qmin : 0.2;
    B.qmax = cfg->qmax > 0.0 ? cfg->qmax : 10.0;
    B.gamma = cfg->gamma > 0.0 ? cfg->gamma : 0.9;
    B.qsteady_min = cfg->qsteady_min > 0.0 ? cfg->qsteady_min : 1.0;
    B.qsteady_max = cfg->qsteady_max > 0.0 ? cfg->qsteady_max : 1.0;
    B.qoldinit = cfg->qoldinit > 0.0 ? cfg->qoldinit : 1e-4;
    B.qoldinit_pb2 = std::pow(B.qoldinit, B.beta2);
    B.inv_qmax = 1.0 / B.qmax;
    B.inv_qmin = 1.0 / B.qmin;
    B.inv_gamma = 1.0 / B.gamma;
    B.initial_diffusion = cfg->initial_diffusion > 0.0 ? cfg->initial_diffusion : 1.0;
    B.maxiters = cfg->maxiters > 0 ? cfg->maxiters : 1000000;
    B.calibrate = cfg->calibrate;
    B.order = cfg->q;
    {
        /* left square root of the preconditioned IWP covariance (src/priors/iwp.jl:74-94) */
        const int q = cfg->q, n = q + 1;
        for (int m = 0; m <= q; m++)
            for (int c = 0; c <= m; c++) {
                double fq = factorial_d(q - c);
                B.L[m * n + c] = std::sqrt((double)(2 * q - 2 * m + 1)) * (fq * fq) / factorial_d(m - c) /
                                 factorial_d(2 * q - c - m + 1);
            }
    }
    {
        /* Qb = L'L accumulated in long double */
        const int n = cfg->q + 1;
        for (int a = 0; a < n; a++)
            for (int b = 0; b < n; b++) {
                long double acc = 0.0L;
                for (int m = 0; m < n; m++) acc += (long double)B.L[m * n + a] * (long double)B.L[m * n + b];
                B.Qb[a * n + b] = (double)acc;
            }
    }
    {
        /* U = qr(L).R in long double (U'U = L'L = Q_breve) */
        const int n = cfg->q + 1;
        long double A[PN_MAX_N][PN_MAX_N];
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) A[i][j] = (long double)B.L[i * n + j];
        for (int k = 0; k < n; k++) {
            long double tot = 0.0L;
            for (int r = k; r < n; r++) tot += A[r][k] * A[r][k];
            if (tot > 0.0L) {
                const long double alpha = A[k][k], nrm = sqrtl(tot);
                const long double beta = (alpha >= 0.0L) ? -nrm : nrm, vk = alpha - beta;
                const long double gam = 1.0L / (tot + fabsl(alpha) * nrm);
                for (int j = k + 1; j < n; j++) {
                    long double sdot = vk * A[k][j];
                    for (int r = k + 1; r < n; r++) sdot += A[r][k] * A[r][j];
                    sdot *= gam;
                    A[k][j] -= sdot * vk;
                    for (int r = k + 1; r < n; r++) A[r][j] -= sdot * A[r][k];
                }
                A[k][k] = beta;
                for (int r = k + 1; r < n; r++) A[r][k] = 0.0L;
            }
        }
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) B.U[i * n + j] = (j >= i) ? (double)A[i][j] : 0.0;
    }
    {
        /* 16-point Gauss-Legendre rule on [0, 1] (Newton on the Legendre recurrence) for the IOUP process noise */
        B.rate = cfg->rate_parameter;
        const int m = 16;
        for (int i = 0; i < m; i++) {
            double x = std::cos(M_PI * (i + 0.75) / (m + 0.5)), dp = 1.0;
            for (int it = 0; it < 100; it++) {
                double p0 = 1.0, p1 = x;
                for (int k = 2; k <= m; k++) {
                    const double p2 = ((2.0 * k - 1.0) * x * p1 - (k - 1.0) * p0) / k;
                    p0 = p1;
                    p1 = p2;
                }
                dp = m * (x * p1 - p0) / (x * x - 1.0);
                const double dx = p1 / dp;
                x -= dx;
                if (std::fabs(dx) < 1e-16) break;
            }
            B.glx[i] = 0.5 * (x + 1.0);
            B.glw[i] = 1.0 / ((1.0 - x * x) * dp * dp); /* = (2 / ((1-x^2) P'^2)) / 2 */
        }
    }
    /* ---- device ---- */
    int ndev = pnode_device_count();
    if (ndev <= 0) {
        delete s;
        return fail(PNODE_ERR_NO_DEVICE, "no CUDA device available: libpnode_cuda has no CPU fallback");
    }
    if (cfg->device < 0 || cfg->device >= ndev) {
        delete s;
        return fail(PNODE_ERR_ARGUMENT, "invalid device ordinal");
    }
    /* from here on a failure must release what the half-built solver already owns (stream, device buffers) */
#define CUDA_TRY_S(expr)                                                                                 \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            pnode_destroy(s);                                                                            \
            return fail(PNODE_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));             \
        }                                                                                                \
    } while (0)
    CUDA_TRY_S(cudaSetDevice(cfg->device));
    CUDA_TRY_S(cudaDeviceGetAttribute(&s->n_sm, cudaDevAttrMultiProcessorCount, cfg->device));
    CUDA_TRY_S(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    CUDA_TRY_S(cudaMalloc(&s->d_counter, sizeof(unsigned long long)));
    {
        /* tstops = union(data.t, tstops) (fenrir.jl:46) */
        std::vector<double> ts;
        if (cfg->n_tstops > 0 && cfg->tstops) ts.assign(cfg->tstops, cfg->tstops + cfg->n_tstops);
        if (cfg->n_data > 0) {
            ts.insert(ts.end(), cfg->data_t, cfg->data_t + cfg->n_data);
            std::sort(ts.begin(), ts.end());
            ts.erase(std::unique(ts.begin(), ts.end()), ts.end());
        }
        if (!ts.empty()) {
            CUDA_TRY_S(cudaMalloc(&s->d_tstops, sizeof(double) * ts.size()));
            CUDA_TRY_S(cudaMemcpy(s->d_tstops, ts.data(), sizeof(double) * ts.size(), cudaMemcpyHostToDevice));
            s->host_tstops = ts;
            B.tstops = s->d_tstops;
            B.n_tstops = (long long)ts.size();
        }
    }
    if (cfg->n_data > 0) {
        const int o = cfg->n_obs, dul = cfg->second_order ? 2 * cfg->d : cfg->d;
        s->n_data = cfg->n_data;
        s->n_obs = o;
        std::vector<double> H, Cv, Rn;
        if (int orc = build_obs_model(cfg, D, &H, &Cv, &Rn)) {
            pnode_destroy(s);
            return orc;
        }
        CUDA_TRY_S(cudaMalloc(&s->d_data_t, sizeof(double) * cfg->n_data));
        CUDA_TRY_S(cudaMalloc(&s->d_data_u, sizeof(double) * cfg->n_data * o));
        CUDA_TRY_S(cudaMalloc(&s->d_obs_H, sizeof(double) * H.size()));
        CUDA_TRY_S(cudaMalloc(&s->d_obs_Rn, sizeof(double) * Rn.size()));
        CUDA_TRY_S(cudaMemcpy(s->d_data_t, cfg->data_t, sizeof(double) * cfg->n_data, cudaMemcpyHostToDevice));
        CUDA_TRY_S(cudaMemcpy(s->d_data_u, cfg->data_u, sizeof(double) * cfg->n_data * o, cudaMemcpyHostToDevice));
        CUDA_TRY_S(cudaMemcpy(s->d_obs_H, H.data(), sizeof(double) * H.size(), cudaMemcpyHostToDevice));
        CUDA_TRY_S(cudaMemcpy(s->d_obs_Rn, Rn.data(), sizeof(double) * Rn.size(), cudaMemcpyHostToDevice));
    }
    if (cfg->n_forced > 0 && cfg->forced_diffusion) {
        CUDA_TRY_S(cudaMalloc(&s->d_forced, sizeof(double) * cfg->n_forced));
        CUDA_TRY_S(cudaMemcpy(s->d_forced, cfg->forced_diffusion, sizeof(double) * cfg->n_forced, cudaMemcpyHostToDevice));
        B.forced_diffusion = s->d_forced;
        B.n_forced = cfg->n_forced;
    }
    B.work_counter = s->d_counter;
    {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, cfg->device) == cudaSuccess) {
            unsigned long long thr = ~0ULL;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
    }
    if (cfg->n_saveat > 0) {
        CUDA_TRY_S(cudaMalloc(&s->d_saveat, sizeof(double) * cfg->n_saveat));
        CUDA_TRY_S(cudaMemcpy(s->d_saveat, cfg->saveat, sizeof(double) * cfg->n_saveat, cudaMemcpyHostToDevice));
        s->n_saveat = cfg->n_saveat;
    }
    if (s->gen) {
        B.update_rate = cfg->update_rate_parameter;
        if (cfg->rate_matrix) {
            CUDA_TRY_S(cudaMalloc(&s->d_rate, sizeof(double) * cfg->d * cfg->d));
            CUDA_TRY_S(cudaMemcpy(s->d_rate, cfg->rate_matrix, sizeof(double) * cfg->d * cfg->d, cudaMemcpyHostToDevice));
            B.rate_matrix = s->d_rate;
        }
    }
    s->cfg.rate_matrix = nullptr;
    s->cfg.saveat = nullptr;
    s->cfg.data_t = s->cfg.data_u = s->cfg.observation_matrix = s->cfg.observation_noise_cov = nullptr;
#undef CUDA_TRY_S
    int rc = get_kernel(s, save_level(cfg, s->cta), cfg->save_everystep ? &s->k_save : &s->k_final);
    if (!rc && s->gen && s->smem == 0) { /* per-warp workspace of the general-prior filter: one slot per resident warp of the persistent grid */
        Kernel& kk = cfg->save_everystep ? s->k_save : s->k_final;
        kk.blocks_per_sm = std::min(kk.blocks_per_sm, 2); /* keep the working set of the grid near the L2 capacity */
        const size_t slots = (size_t)s->n_sm * kk.blocks_per_sm * PN_GEN_WARPS;
        if (cudaMalloc(&s->d_gen_ws, slots * s->gen_ws_doubles * sizeof(double)) != cudaSuccess) rc = fail(PNODE_ERR_CUDA, "cudaMalloc(general-prior workspace)");
        s->base.gen_workspace = s->d_gen_ws;
    }
    if (!rc && cfg->smooth) rc = get_smoother(s);
    if (!rc && cfg->n_saveat > 0) rc = get_dense(s);
    if (rc) {
        std::string keep = g_err;
        pnode_destroy(s);
        g_err = keep;
        return rc;
    }
    *out = s;
    return 0;
}

extern "C" double pnode_compile_seconds(const pnode_solver* s) { return s ? s->compile_s : 0.0; }

extern "C" int pnode_measure_fp64_peak(int device, double* tflops) {
    if (!tflops) return fail(PNODE_ERR_ARGUMENT, "null argument");
    if (pnode_device_count() <= 0) return fail(PNODE_ERR_NO_DEVICE, "no CUDA device available");
    CUDA_TRY(cudaSetDevice(device));
    int n_sm = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
    const int threads = 256, blocks = n_sm * 8, iters = 1 << 15;
    double* d_out = nullptr;
    CUDA_TRY(cudaMalloc(&d_out, sizeof(double) * (size_t)threads * blocks));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        CUDA_TRY(cudaEventRecord(e0, 0));
        pn_dfma_probe<<<blocks, threads>>>(d_out, iters, 1.0000001, 1e-9);
        CUDA_TRY(cudaEventRecord(e1, 0));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double fl = 2.0 * 8.0 * (double)iters * threads * (double)blocks;
        if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    *tflops = best;
    return 0;
}

extern "C" void pnode_destroy(pnode_solver* s) {
    if (!s) return;
    cudaSetDevice(s->cfg.device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    for (Kernel* k : {&s->k_final, &s->k_save, &s->k_smooth, &s->k_dense})
        if (k->mod && g_drv.ModuleUnload) g_drv.ModuleUnload(k->mod);
    if (s->d_counter) cudaFree(s->d_counter);
    if (s->d_tstops) cudaFree(s->d_tstops);
    if (s->d_forced) cudaFree(s->d_forced);
    if (s->d_saveat) cudaFree(s->d_saveat);
    if (s->d_rate) cudaFree(s->d_rate);
    if (s->d_gen_ws) cudaFree(s->d_gen_ws);
    for (double* ptr : {s->d_data_t, s->d_data_u, s->d_obs_H, s->d_obs_Rn})
        if (ptr) cudaFree(ptr);
    if (s->d_init_mu) cudaFreeAsync(s->d_init_mu, s->stream);
    if (s->d_init_dt) cudaFreeAsync(s->d_init_dt, s->stream);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

/* pn_init_kernel: initial derivatives + initial step of every trajectory into the solver's init buffers */
static int launch_init(pnode_solver* s, Kernel& k, PnParams& P) {
    if (!k.rt_init && !k.drv_init) return 0;
    const size_t N = (size_t)P.n_traj;
    if (s->init_cap < N) {
        if (s->d_init_mu) CUDA_TRY(cudaFreeAsync(s->d_init_mu, s->stream));
        if (s->d_init_dt) CUDA_TRY(cudaFreeAsync(s->d_init_dt, s->stream));
        CUDA_TRY(cudaMallocAsync((void**)&s->d_init_mu, N * s->D * 8, s->stream));
        CUDA_TRY(cudaMallocAsync((void**)&s->d_init_dt, N * 8, s->stream));
        s->init_cap = N;
    }
    const unsigned grid = (unsigned)std::min<size_t>((N + 127) / 128, (size_t)s->n_sm * 16);
    void* args[] = {&P, &s->d_init_mu, &s->d_init_dt};
    if (k.rt_init) {
        CUDA_TRY(cudaLaunchKernel(k.rt_init, dim3(grid), dim3(128), args, 0, s->stream));
    } else {
        CUresult r = g_drv.LaunchKernel(k.drv_init, grid, 1, 1, 128, 1, 1, 0, (CUstream)s->stream, args, nullptr);
        if (r != CUDA_SUCCESS) return fail(PNODE_ERR_CUDA, "cuLaunchKernel(pn_init): " + cu_err(r));
    }
    P.init_mu = s->d_init_mu;
    P.init_dt = s->d_init_dt;
    return 0;
}

static int launch(pnode_solver* s, Kernel& k, PnParams& P, int64_t n_traj) {
    if (!P.init_mu && (k.rt_init || k.drv_init)) {
        int rci = launch_init(s, k, P);
        if (rci) return rci;
    }
    const int groups_per_block = std::max(1, k.threads / s->G);
    long long want = (n_traj + groups_per_block - 1) / groups_per_block;
    long long cap = (long long)s->n_sm * k.blocks_per_sm; /* persistent: a whole number of CTAs per SM */
    unsigned grid = (unsigned)std::max(1LL, std::min(want, cap));
    CUDA_TRY(cudaMemsetAsync(s->d_counter, 0, sizeof(unsigned long long), s->stream));
    if (k.rt_fn) {
        void* args[] = {&P};
        CUDA_TRY(cudaLaunchKernel(k.rt_fn, dim3(grid), dim3(k.threads), args, s->smem, s->stream));
    } else {
        void* args[] = {&P};
        CUresult r = g_drv.LaunchKernel(k.drv_fn, grid, 1, 1, k.threads, 1, 1, (unsigned)s->smem, (CUstream)s->stream, args, nullptr);
        if (r != CUDA_SUCCESS) return fail(PNODE_ERR_CUDA, "cuLaunchKernel: " + cu_err(r));
    }
    return 0;
}

static int alloc_field(pnode_result* r, int field, size_t bytes) {
    if (bytes == 0) return 0;
    void* p = nullptr;
    CUDA_TRY(cudaMallocAsync(&p, bytes, r->s->stream));
    r->f[field].d = p;
    r->f[field].bytes = bytes;
    return 0;
}

/* Temporaries of one solve_impl call (timing events, work buffers): released on every exit path, early returns included */
struct SolveScope {
    cudaStream_t st;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    std::vector<void*> bufs;
    explicit SolveScope(cudaStream_t stream) : st(stream) {}
    void track(void* ptr) {
        if (ptr) bufs.push_back(ptr);
    }
    void release(void* ptr) { /* free now */
        auto it = std::find(bufs.begin(), bufs.end(), ptr);
        if (it == bufs.end()) return;
        bufs.erase(it);
        cudaFreeAsync(ptr, st);
    }
    ~SolveScope() {
        for (void* ptr : bufs) cudaFreeAsync(ptr, st);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    }
};

static int solve_impl(pnode_solver* s, int64_t n_traj, const double* d_u0, const double* d_p, pnode_result* r,
                      bool force_count_pass, bool* need_count_pass) {
    const std::vector<double>& host_tstops = s->host_tstops;
    const int d = s->cfg.d, D = s->D;
    const int du = s->cfg.second_order ? 2 * d : d; /* length of sol.u: (du, u) for second-order problems */
    const size_t N = (size_t)n_traj;
    int rc;
    if ((rc = alloc_field(r, PNODE_F_T_END, N * 8))) return rc;
    if ((rc = alloc_field(r, PNODE_F_U_FINAL, N * du * 8))) return rc;
    if ((rc = alloc_field(r, PNODE_F_U_FINAL_COV, N * du * du * 8))) return rc;
    if (s->cfg.save_state) {
        if ((rc = alloc_field(r, PNODE_F_X_FINAL, N * D * 8))) return rc;
        if ((rc = alloc_field(r, PNODE_F_X_FINAL_COVFAC, N * D * D * 8))) return rc;
    }
    if ((rc = alloc_field(r, PNODE_F_LOGLIK, N * 8))) return rc;
    if (s->data_forward && (rc = alloc_field(r, PNODE_F_DATA_LOGLIK, N * 8))) return rc;
    if ((rc = alloc_field(r, PNODE_F_DIFFUSION, N * 8))) return rc;
    if (s->mv && (rc = alloc_field(r, PNODE_F_DIFFUSION_MV, N * d * 8))) return rc;
    if ((rc = alloc_field(r, PNODE_F_DT_LAST, N * 8))) return rc;
    if ((rc = alloc_field(r, PNODE_F_NACCEPT, N * 8))) return rc;
    if ((rc = alloc_field(r, PNODE_F_NREJECT, N * 8))) return rc;
    if ((rc = alloc_field(r, PNODE_F_NF, N * 8))) return rc;
    if ((rc = alloc_field(r, PNODE_F_RETCODE, N * 4))) return rc;

    PnParams P = s->base;
    P.n_traj = n_traj;
    P.u0 = d_u0;
    P.p = d_p;
    P.t_end = (double*)r->f[PNODE_F_T_END].d;
    P.u_mean = (double*)r->f[PNODE_F_U_FINAL].d;
    P.u_cov = (double*)r->f[PNODE_F_U_FINAL_COV].d;
    P.x_mean = (double*)r->f[PNODE_F_X_FINAL].d;
    P.x_covfac = (double*)r->f[PNODE_F_X_FINAL_COVFAC].d;
    P.loglik = (double*)r->f[PNODE_F_LOGLIK].d;
    if (s->data_forward) { /* DataUpdateCallback: the step kernel conditions on the observations as it passes their times */
        P.data_t = s->d_data_t;
        P.data_u = s->d_data_u;
        P.n_data = s->n_data;
        P.data_loglik = (double*)r->f[PNODE_F_DATA_LOGLIK].d;
    }
    P.diffusion = (double*)r->f[PNODE_F_DIFFUSION].d;
    P.diffusion_mv = (double*)r->f[PNODE_F_DIFFUSION_MV].d;
    P.dt_last = (double*)r->f[PNODE_F_DT_LAST].d;
    P.naccept = (long long*)r->f[PNODE_F_NACCEPT].d;
    P.nreject = (long long*)r->f[PNODE_F_NREJECT].d;
    P.nf = (long long*)r->f[PNODE_F_NF].d;
    P.retcode = (int*)r->f[PNODE_F_RETCODE].d;

    SolveScope scope(s->stream);
    CUDA_TRY(cudaEventCreate(&scope.e0));
    CUDA_TRY(cudaEventCreate(&scope.e1));
    cudaEvent_t e0 = scope.e0, e1 = scope.e1;
    CUDA_TRY(cudaEventRecord(e0, s->stream));
    int64_t launches = 0;
    {
        /* group-per-trajectory path: initial states of the whole ensemble first (inside the timed region) */
        Kernel& kmain = s->cfg.save_everystep ? s->k_save : s->k_final;
        if (kmain.rt_init || kmain.drv_init) {
            if ((rc = launch_init(s, kmain, P))) return rc;
            launches += 1;
        }
    }
    long long fixed_steps_total = -1; /* >= 0: the counting pass was skipped and every trajectory was allotted this many steps in total */
    if (!s->cfg.save_everystep) {
        if ((rc = launch(s, s->k_final, P, n_traj))) return rc;
        launches += 1;
    } else {
        r->offsets.resize(N + 1);
        r->offsets[0] = 0;
        long long fixed_steps = -1;
        const long long pilot_stride = 16;
        long long pitch = 0;
        bool pitch_mode = false;
        std::vector<void*> work_bufs; /* pitch mode: internal buffers in the over-allocated layout */
        const long long* d_counts = nullptr;
        if (!s->cfg.adaptive && !force_count_pass) {
            /* fixed steps: the number of steps is a function of (t0, t1, dt, tstops) alone -- replay the kernels' time
               bookkeeping (same IEEE operations) on the host and skip the counting pass.  A trajectory that stops early
               (retcode != Success) is detected from the totals afterwards and the solve is repeated with the counting pass. */
            double t = s->base.t0;
            long long nst = 0, it = 0, ti = 0;
            const double* ts = host_tstops.data();
            const long long nts = (long long)host_tstops.size();
            while (t < s->base.t1) {
                double tstop = s->base.t1;
                while (ti < nts && ts[ti] <= t) ti++;
                if (ti < nts && ts[ti] < s->base.t1) tstop = ts[ti];
                if (++it > s->base.maxiters) break;
                const double dt = std::fmin(s->base.dt0, tstop - t);
                if (dt != dt || !(t + dt > t)) break;
                const double ttmp = t + dt;
                const double ref = std::fmax(std::fabs(t), std::fabs(tstop));
                const double epsref = std::nextafter(ref, 1e300) - ref;
                t = (std::fabs(ttmp - tstop) < 100.0 * epsref) ? tstop : ttmp;
                nst++;
            }
            fixed_steps = nst;
            fixed_steps_total = nst * (long long)N;
            for (size_t i = 0; i < N; i++) r->offsets[i + 1] = r->offsets[i] + nst + 1;
        } else if (!force_count_pass && !s->cfg.save_state && !env_int("PNODE_TWO_PASS", 0) && (long long)N >= 64LL * pilot_stride) {
            /* adaptive steps, one full pass instead of two: a PILOT run of the same binary in counting mode on every
               pilot_stride-th trajectory gives the step-count scale of the ensemble; every trajectory is allotted a segment
               of pitch = margin * max_pilot (+ slack) saved points and the save pass writes into this over-allocated layout
               (writes beyond a segment are dropped, the count goes on).  The small user-facing fields are compacted to the
               exact CSR afterwards; the big state records stay internal.  If any trajectory overflows its segment the exact
               counts are known by then and the solve is repeated with the counting-pass layout. */
            PnParams Pp = P;
            Pp.traj_stride = pilot_stride;
            CUDA_TRY(cudaMemsetAsync(P.naccept, 0, N * 8, s->stream));
            if ((rc = launch(s, s->k_save, Pp, (n_traj + pilot_stride - 1) / pilot_stride))) return rc;
            launches += 1;
            std::vector<long long> nacc(N);
            CUDA_TRY(cudaMemcpyAsync(nacc.data(), P.naccept, N * 8, cudaMemcpyDeviceToHost, s->stream));
            CUDA_TRY(cudaStreamSynchronize(s->stream));
            long long mx = 0;
            for (size_t i = 0; i < N; i += (size_t)pilot_stride) mx = std::max(mx, nacc[i]);
            const double margin = env_int("PNODE_PILOT_MARGIN_PCT", 125) / 100.0;
            pitch = (long long)std::ceil((double)mx * margin) + 8 + 1;
            pitch_mode = true;
            for (size_t i = 0; i < N; i++) r->offsets[i + 1] = r->offsets[i] + pitch;
        } else {
            /* pass 1: count accepted steps with the *same* binary (null step_offsets => no stores), so both
               passes take bit-identical accept/reject decisions; pass 2 writes at exact CSR positions */
            if ((rc = launch(s, s->k_save, P, n_traj))) return rc;
            launches += 1;
            std::vector<long long> nacc(N);
            CUDA_TRY(cudaMemcpyAsync(nacc.data(), P.naccept, N * 8, cudaMemcpyDeviceToHost, s->stream));
            CUDA_TRY(cudaStreamSynchronize(s->stream));
            for (size_t i = 0; i < N; i++) r->offsets[i + 1] = r->offsets[i] + nacc[i] + 1;
        }
        const size_t T = (size_t)r->offsets[N]; /* points in the layout the kernels write (exact CSR, or N * pitch) */
        /* exact layout: the kernels write the result fields directly; pitch layout: internal work buffers */
        auto alloc_work = [&](int field, size_t bytes, void** out) -> int {
            *out = nullptr;
            if (bytes == 0) return 0;
            if (!pitch_mode) {
                int rc2 = alloc_field(r, field, bytes);
                *out = r->f[field].d;
                return rc2;
            }
            CUDA_TRY(cudaMallocAsync(out, bytes, s->stream));
            work_bufs.push_back(*out);
            scope.track(*out);
            return 0;
        };
        void *w_off = nullptr, *w_t = nullptr, *w_u = nullptr, *w_uc = nullptr, *w_df = nullptr, *w_dfmv = nullptr, *w_x = nullptr,
             *w_xc = nullptr;
        if ((rc = alloc_work(PNODE_F_STEP_OFFSETS, (N + 1) * 8, &w_off))) return rc;
        if ((rc = alloc_work(PNODE_F_T, T * 8, &w_t))) return rc;
        if ((rc = alloc_work(PNODE_F_U, T * du * 8, &w_u))) return rc;
        if ((rc = alloc_work(PNODE_F_U_COV, T * du * du * 8, &w_uc))) return rc;
        if ((rc = alloc_work(PNODE_F_DIFFUSIONS, T * 8, &w_df))) return rc;
        if (s->mv && (rc = alloc_work(PNODE_F_DIFFUSIONS_MV, T * d * 8, &w_dfmv))) return rc;
        /* slot k of a trajectory's diffusions belongs to the step k -> k+1: its last slot is never written, keep it defined */
        CUDA_TRY(cudaMemsetAsync(w_df, 0, T * 8, s->stream));
        if (w_dfmv) CUDA_TRY(cudaMemsetAsync(w_dfmv, 0, T * d * 8, s->stream));
        CUDA_TRY(cudaMemcpyAsync(w_off, r->offsets.data(), (N + 1) * 8, cudaMemcpyHostToDevice, s->stream));
        P.step_offsets = (const long long*)w_off;
        P.s_t = (double*)w_t;
        P.s_u_mean = (double*)w_u;
        P.s_u_cov = (double*)w_uc;
        P.s_diffusion = (double*)w_df;
        P.s_diffusion_mv = (double*)w_dfmv;
        double* d_h = nullptr;
        const bool dense = s->n_saveat > 0;
        const size_t V = N * (size_t)s->n_saveat;
        if (s->cfg.smooth || dense || (s->cfg.save_state && !s->cta)) { /* save_state without smoothing: sol.x_filt (group / thread kernels) */
            if ((rc = alloc_work(PNODE_F_X, T * D * 8, &w_x))) return rc;
            if ((rc = alloc_work(PNODE_F_X_COVFAC, T * D * D * 8, &w_xc))) return rc;
            CUDA_TRY(cudaMallocAsync((void**)&d_h, T * 8, s->stream));
            scope.track(d_h);
            CUDA_TRY(cudaMemsetAsync(d_h, 0, T * 8, s->stream));
            P.s_x_mean = (double*)w_x;
            P.s_x_covfac = (double*)w_xc;
            P.s_h = d_h;
            if (s->gen && s->base.update_rate && s->cfg.smooth) { /* the backward kernel of a step needs the rate it was taken with */
                double* d_sr = nullptr;
                CUDA_TRY(cudaMallocAsync((void**)&d_sr, T * d * d * 8, s->stream));
                scope.track(d_sr);
                P.s_rate = d_sr;
            }
        }
        if ((rc = launch(s, s->k_save, P, n_traj))) return rc;
        launches += 1;
        std::vector<long long> exact_off;
        if (pitch_mode) {
            /* the save pass counted every accepted step, stored or not: exact CSR offsets, overflow check */
            std::vector<long long> nacc(N);
            CUDA_TRY(cudaMemcpyAsync(nacc.data(), P.naccept, N * 8, cudaMemcpyDeviceToHost, s->stream));
            CUDA_TRY(cudaStreamSynchronize(s->stream));
            exact_off.resize(N + 1);
            exact_off[0] = 0;
            bool overflow = false;
            for (size_t i = 0; i < N; i++) {
                overflow = overflow || (nacc[i] + 1 > pitch);
                exact_off[i + 1] = exact_off[i] + nacc[i] + 1;
            }
            if (overflow) { /* repeat with the counting-pass layout (solve_common) */
                for (void* ptr : work_bufs) scope.release(ptr);
                if (d_h) scope.release(d_h);
                CUDA_TRY(cudaStreamSynchronize(s->stream));
                if (need_count_pass) *need_count_pass = true;
                return need_count_pass ? 0 : fail(PNODE_ERR_CUDA, "pilot segment overflow without a fallback");
            }
            d_counts = P.naccept;
        }
        /* the smoother sweep over a CSR of records (the real trajectories, or the two-point virtual ones of the dense output) */
        auto launch_sweep = [&](PnSmoothParams& Q, long long ntraj) -> int {
            Q.d = d;
            Q.q = s->cfg.q;
            memcpy(Q.L, s->base.L, sizeof Q.L);
            memcpy(Q.U, s->base.U, sizeof Q.U);
            Q.rate = s->base.rate;
            memcpy(Q.glx, s->base.glx, sizeof Q.glx);
            memcpy(Q.glw, s->base.glw, sizeof Q.glw);
            Q.n_traj = ntraj;
            Q.work_counter = s->d_counter;
            /* Fenrir mode: the column-distributed register sweep carries the data updates; the row-distributed variant and
               observation vectors longer than sol.u go through the shared-memory sweep (PNODE_FENRIR_GENERIC=1 forces it) */
            const int du_ = s->cfg.second_order ? 2 * d : d;
            const bool fen_reg = env_int("PNODE_SMOOTH_KIND", 1) != 0 && Q.n_obs <= du_ && !env_int("PNODE_FENRIR_GENERIC", 0);
            if (s->k_smooth.valid() && (!Q.fenrir_ll || fen_reg)) {
                /* register-resident sweep: Gs lanes per trajectory, persistent, one CTA per SM */
                Kernel& ks = s->k_smooth;
                const int groups = ks.threads / s->Gs;
                long long want = (ntraj + groups - 1) / groups;
                unsigned grid = (unsigned)std::max(1LL, std::min(want, (long long)s->n_sm * ks.blocks_per_sm));
                CUDA_TRY(cudaMemsetAsync(s->d_counter, 0, sizeof(unsigned long long), s->stream));
                void* args[] = {&Q};
                if (ks.rt_fn) {
                    CUDA_TRY(cudaLaunchKernel(ks.rt_fn, dim3(grid), dim3(ks.threads), args, s->smem_smooth, s->stream));
                } else {
                    CUresult cr = g_drv.LaunchKernel(ks.drv_fn, grid, 1, 1, ks.threads, 1, 1, (unsigned)s->smem_smooth,
                                                     (CUstream)s->stream, args, nullptr);
                    if (cr != CUDA_SUCCESS) return fail(PNODE_ERR_CUDA, "cuLaunchKernel(smoother): " + cu_err(cr));
                }
            } else {
                if (s->cfg.prior != PNODE_PRIOR_IWP && !s->gen)
                    return fail(PNODE_ERR_UNSUPPORTED, "the shared-memory smoother sweep implements the IWP prior and the general-rate IOUP only");
                const int n = s->cfg.q + 1;
                if (s->gen) { /* dense transition pairs recomputed per step (pn_lti.cuh) */
                    Q.gen_prior = 1;
                    Q.rate_matrix = s->d_rate;
                    Q.s_rate = P.s_rate;
                }
                const size_t per_warp =
                    (size_t)(7 * D * D + 4 * D + 2 * n * n + 2 * n + (Q.fenrir_ll ? PN_FENRIR_EXTRA(D, Q.n_obs) : 0) +
                             (s->gen ? PN_SMOOTH_GEN_EXTRA(D, d, n) : 0)) * sizeof(double);
                int warps = PN_SM_WARPS;
                const void* fn = pn_smooth_kernel_ptr();
                /* states whose work matrices (7 D^2 doubles per warp) do not fit in shared memory -- Pleiades, D = 112: 0.7 MB --
                   sweep out of a global workspace that stays L2-resident (filter + smoother) */
                const bool global_ws = per_warp > 200 * 1024;
                while (!global_ws && warps > 1 && per_warp * warps > 200 * 1024) warps--;
                const size_t smem = global_ws ? 0 : per_warp * warps;
                if (!global_ws) CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                int nblk = 0;
                CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nblk, fn, 32 * warps, smem));
                if (nblk < 1) return fail(PNODE_ERR_CUDA, "smoother kernel does not fit on an SM");
                if (global_ws) nblk = std::min(nblk, 2); /* 8 warps per SM x 0.7 MB: the working set of the whole grid stays below L2 */
                long long want = (ntraj + warps - 1) / warps;
                unsigned grid = (unsigned)std::max(1LL, std::min(want, (long long)s->n_sm * nblk));
                double* ws = nullptr;
                if (global_ws) {
                    if (warps != PN_SM_WARPS) return fail(PNODE_ERR_CUDA, "workspace indexing assumes PN_SM_WARPS warps per CTA");
                    CUDA_TRY(cudaMallocAsync((void**)&ws, (size_t)grid * warps * per_warp, s->stream));
                    scope.track(ws);
                }
                Q.workspace = ws;
                CUDA_TRY(cudaMemsetAsync(s->d_counter, 0, sizeof(unsigned long long), s->stream));
                void* args[] = {&Q};
                CUDA_TRY(cudaLaunchKernel(fn, dim3(grid), dim3(32 * warps), args, smem, s->stream));
                if (ws) scope.release(ws);
            }
            return 0;
        };
        /* dense output, first half: goal_pred = predict(x_filt[idx], h1) for every (trajectory, saveat point) pair; must run
           before the sweep overwrites the filtered records with the smoothed ones */
        PnDenseParams Dn;
        memset(&Dn, 0, sizeof Dn);
        double *v_mean = nullptr, *v_covfac = nullptr, *v_h = nullptr, *v_diff = nullptr, *vu_mean = nullptr, *vu_cov = nullptr, *v_diff_mv = nullptr;
        long long *v_idx = nullptr, *v_off = nullptr;
        if (dense) {
            if ((rc = alloc_field(r, PNODE_F_SAVEAT_U, V * du * 8))) return rc;
            if ((rc = alloc_field(r, PNODE_F_SAVEAT_U_COV, V * du * du * 8))) return rc;
            Dn.n_traj = n_traj;
            Dn.n_saveat = s->n_saveat;
            Dn.saveat = s->d_saveat;
            Dn.smooth = s->cfg.smooth;
            Dn.calibrate_scale = (!is_dynamic(s->cfg.diffusion) && s->cfg.calibrate) ? 1 : 0;
            Dn.step_offsets = P.step_offsets;
            Dn.counts = d_counts;
            Dn.s_t = P.s_t;
            Dn.s_h = d_h;
            Dn.s_diffusion = P.s_diffusion;
            Dn.s_x_mean = P.s_x_mean;
            Dn.s_x_covfac = P.s_x_covfac;
            Dn.gdiff = P.diffusion;
            Dn.s_diffusion_mv = s->mv ? P.s_diffusion_mv : nullptr;
            Dn.gdiff_mv = s->mv ? P.diffusion_mv : nullptr;
            Dn.initial_diffusion = s->base.initial_diffusion;
            memcpy(Dn.U, s->base.U, sizeof Dn.U);
            Dn.rate = s->base.rate;
            memcpy(Dn.glx, s->base.glx, sizeof Dn.glx);
            memcpy(Dn.glw, s->base.glw, sizeof Dn.glw);
            Dn.out_u = (double*)r->f[PNODE_F_SAVEAT_U].d;
            Dn.out_cov = (double*)r->f[PNODE_F_SAVEAT_U_COV].d;
            if (s->cfg.smooth) {
                CUDA_TRY(cudaMallocAsync((void**)&v_mean, 2 * V * D * 8, s->stream));
                scope.track(v_mean);
                CUDA_TRY(cudaMallocAsync((void**)&v_covfac, 2 * V * D * D * 8, s->stream));
                scope.track(v_covfac);
                CUDA_TRY(cudaMallocAsync((void**)&v_h, 2 * V * 8, s->stream));
                scope.track(v_h);
                CUDA_TRY(cudaMallocAsync((void**)&v_diff, 2 * V * 8, s->stream));
                scope.track(v_diff);
                CUDA_TRY(cudaMallocAsync((void**)&vu_mean, 2 * V * du * 8, s->stream));
                scope.track(vu_mean);
                CUDA_TRY(cudaMallocAsync((void**)&vu_cov, 2 * V * du * du * 8, s->stream));
                scope.track(vu_cov);
                CUDA_TRY(cudaMallocAsync((void**)&v_idx, V * 8, s->stream));
                scope.track(v_idx);
                CUDA_TRY(cudaMallocAsync((void**)&v_off, (V + 1) * 8, s->stream));
                scope.track(v_off);
                Dn.v_idx = v_idx;
                Dn.v_mean = v_mean;
                Dn.v_covfac = v_covfac;
                Dn.v_h = v_h;
                Dn.v_diffusion = v_diff;
                if (s->mv) {
                    CUDA_TRY(cudaMallocAsync((void**)&v_diff_mv, 2 * V * d * 8, s->stream));
                    scope.track(v_diff_mv);
                    Dn.v_diffusion_mv = v_diff_mv;
                }
            }
            Kernel& kd = s->k_dense;
            const int groups = kd.threads / s->Gd;
            long long want = ((long long)V + groups - 1) / groups;
            if (kd.rt_fn) { /* D > 32: warp per pair, workspace in shared memory or (Pleiades: 0.2 MB per warp) in a global buffer */
                const int n = s->cfg.q + 1;
                const size_t per_warp = (size_t)PN_DENSE_GEN_WS(D, n) * sizeof(double);
                const bool global_ws = per_warp * PN_SM_WARPS > 200 * 1024;
                const size_t smem = global_ws ? 0 : per_warp * PN_SM_WARPS;
                if (!global_ws && smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kd.rt_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                unsigned grid = (unsigned)std::max(1LL, std::min(want, (long long)s->n_sm * (global_ws ? 2 : 1)));
                double* ws = nullptr;
                if (global_ws) {
                    CUDA_TRY(cudaMallocAsync((void**)&ws, (size_t)grid * PN_SM_WARPS * per_warp, s->stream));
                    scope.track(ws);
                }
                int dd = d, qq = s->cfg.q;
                void* args[] = {&Dn, &dd, &qq, &ws};
                CUDA_TRY(cudaLaunchKernel(kd.rt_fn, dim3(grid), dim3(kd.threads), args, smem, s->stream));
                if (ws) scope.release(ws);
            } else {
            unsigned grid = (unsigned)std::max(1LL, std::min(want, (long long)s->n_sm * kd.blocks_per_sm * 4));
            void* args[] = {&Dn};
            CUresult cr = g_drv.LaunchKernel(kd.drv_fn, grid, 1, 1, kd.threads, 1, 1, 0, (CUstream)s->stream, args, nullptr);
            if (cr != CUDA_SUCCESS) return fail(PNODE_ERR_CUDA, "cuLaunchKernel(dense): " + cu_err(cr));
            }
            launches += 1;
        }
        if (s->cfg.smooth) {
            /* backward sweep over the real trajectories */
            PnSmoothParams Q;
            memset(&Q, 0, sizeof Q);
            Q.dynamic = is_dynamic(s->cfg.diffusion);
            Q.s_diffusion_mv = P.s_diffusion_mv;
            Q.gdiff_mv = P.diffusion_mv;
            Q.calibrate = s->cfg.calibrate;
            Q.initial_diffusion = s->base.initial_diffusion;
            Q.step_offsets = P.step_offsets;
            Q.counts = d_counts;
            Q.s_h = d_h;
            Q.s_diffusion = P.s_diffusion;
            Q.gdiff = P.diffusion;
            Q.s_x_mean = P.s_x_mean;
            Q.s_x_covfac = P.s_x_covfac;
            Q.s_u_mean = P.s_u_mean;
            Q.s_u_cov = P.s_u_cov;
            Q.write_x = (s->cfg.save_state || dense) ? 1 : 0; /* sol.x_smooth asked for, or the dense output reads the smoothed records */
            if (s->fenrir()) {
                /* Fenrir: the backward pass fits the PN posterior to the data instead of smoothing; the records and the
                   U fields stay the filtering estimates, uncalibrated (the reference stops with step!, no postamble) */
                if ((rc = alloc_field(r, PNODE_F_FENRIR_LOGLIK, N * 8))) return rc;
                Q.calibrate = 0;
                Q.s_t = P.s_t;
                Q.data_t = s->d_data_t;
                Q.data_u = s->d_data_u;
                Q.obs_H = s->d_obs_H;
                Q.obs_Rn = s->d_obs_Rn;
                Q.n_data = s->n_data;
                Q.n_obs = s->n_obs;
                Q.retcode = (const int*)r->f[PNODE_F_RETCODE].d;
                Q.fenrir_ll = (double*)r->f[PNODE_F_FENRIR_LOGLIK].d;
                Q.kron_logdet = s->kron ? 1 : 0; /* the reference's Kronecker log-density (update.jl:140-148, 182-194) */
            }
            if ((rc = launch_sweep(Q, n_traj))) return rc;
            launches += 1;
            if (dense) {
                /* second half: smooth(goal_pred, x_smooth[idx+1], h2) == one backward step of the sweep on two-point
                   virtual trajectories (records already carry the calibration scale; per-record diffusion) */
                const int thr = 256;
                const long long ng = (long long)V * ((long long)D * D + D);
                pn_dense_gather<<<(unsigned)((ng + thr - 1) / thr), thr, 0, s->stream>>>((long long)V, s->n_saveat, D, P.step_offsets, d_counts, v_idx,
                                                                                        P.s_x_mean, P.s_x_covfac, v_mean, v_covfac);
                std::vector<long long> voff(V + 1);
                for (size_t i = 0; i <= V; i++) voff[i] = 2 * (long long)i;
                CUDA_TRY(cudaMemcpyAsync(v_off, voff.data(), (V + 1) * 8, cudaMemcpyHostToDevice, s->stream));
                CUDA_TRY(cudaStreamSynchronize(s->stream)); /* voff is a stack-lifetime host buffer */
                PnSmoothParams Qv;
                memset(&Qv, 0, sizeof Qv);
                Qv.dynamic = 1;
                Qv.calibrate = 0;
                Qv.initial_diffusion = s->base.initial_diffusion;
                Qv.step_offsets = v_off;
                Qv.s_h = v_h;
                Qv.s_diffusion = v_diff;
                Qv.s_diffusion_mv = v_diff_mv;
                Qv.gdiff = P.diffusion;
                Qv.s_x_mean = v_mean;
                Qv.s_x_covfac = v_covfac;
                Qv.s_u_mean = vu_mean;
                Qv.s_u_cov = vu_cov;
                if ((rc = launch_sweep(Qv, (long long)V))) return rc;
                const long long nc = (long long)V * (du + du * du);
                pn_dense_collect<<<(unsigned)((nc + thr - 1) / thr), thr, 0, s->stream>>>((long long)V, du, v_h, vu_mean, vu_cov, Dn.out_u,
                                                                                         Dn.out_cov);
                launches += 3;
                for (void* ptr : {(void*)v_mean, (void*)v_covfac, (void*)v_h, (void*)v_diff, (void*)vu_mean, (void*)vu_cov, (void*)v_idx,
                                  (void*)v_off, (void*)v_diff_mv})
                    scope.release(ptr);
            }
        }
        if (d_h) scope.release(d_h);
        if (!is_dynamic(s->cfg.diffusion)) {
            /* DIFFUSION holds sigma_hat^2 * initial (calibrate) or the constant initial diffusion */
            const double* scale = nullptr;
            if (s->cfg.calibrate && !s->cfg.smooth) { /* the smoother sweep already applied the calibration */
                /* scale = sigma_hat^2 = diffusion / initial_diffusion; the kernel wrote diffusion = mle * initial */
                scale = P.diffusion;
            }
            const int thr = 256;
            const unsigned blocks = (unsigned)((T + thr - 1) / thr);
            pn_calibrate_saved<<<blocks, thr, 0, s->stream>>>((long long)T, (long long)N, P.step_offsets, scale, 1.0 / s->base.initial_diffusion, P.diffusion,
                                                             du * du, P.s_u_cov, P.s_diffusion, d,
                                                             (scale && s->mv) ? P.diffusion_mv : nullptr, P.diffusion_mv,
                                                             P.s_diffusion_mv);
            launches += 1;
            if (scale && P.s_x_covfac && !pitch_mode) { /* sol.x_filt of a filter-only solve: apply_diffusion! on every saved factor */
                const long long ne = (long long)T * D * D;
                pn_calibrate_saved_state<<<(unsigned)((ne + thr - 1) / thr), thr, 0, s->stream>>>((long long)T, (long long)N, P.step_offsets, scale,
                                                                                                1.0 / s->base.initial_diffusion,
                                                                                                s->mv ? P.diffusion_mv : nullptr, D, d, P.s_x_covfac);
                launches += 1;
            }
        }
        if (pitch_mode) {
            /* user-facing fields: over-allocated layout -> exact CSR */
            const size_t Te = (size_t)exact_off[N];
            if ((rc = alloc_field(r, PNODE_F_STEP_OFFSETS, (N + 1) * 8))) return rc;
            CUDA_TRY(cudaMemcpyAsync(r->f[PNODE_F_STEP_OFFSETS].d, exact_off.data(), (N + 1) * 8, cudaMemcpyHostToDevice, s->stream));
            const long long* d_eoff = (const long long*)r->f[PNODE_F_STEP_OFFSETS].d;
            struct { int field; int width; const void* src; } cp[] = {
                {PNODE_F_T, 1, w_t}, {PNODE_F_U, du, w_u}, {PNODE_F_U_COV, du * du, w_uc}, {PNODE_F_DIFFUSIONS, 1, w_df},
                {PNODE_F_DIFFUSIONS_MV, d, w_dfmv}};
            for (auto& c : cp) {
                if (!c.src) continue;
                if ((rc = alloc_field(r, c.field, Te * c.width * 8))) return rc;
                const long long ne = (long long)Te * c.width;
                pn_compact_saved<<<(unsigned)((ne + 255) / 256), 256, 0, s->stream>>>((long long)Te, (long long)N, d_eoff, pitch, c.width,
                                                                                     (const double*)c.src, (double*)r->f[c.field].d);
                launches += 1;
            }
            CUDA_TRY(cudaStreamSynchronize(s->stream)); /* exact_off is a host buffer of this scope */
            for (void* ptr : work_bufs) scope.release(ptr);
            r->offsets.assign(exact_off.begin(), exact_off.end());
        }
        r->info.total_saved = (int64_t)r->offsets[N];
    }
    CUDA_TRY(cudaEventRecord(e1, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    CUDA_TRY(cudaGetLastError());
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    r->info.kernel_ms = ms;
    r->info.n_launches = launches + 1; /* + pn_sum_counts below */
    r->info.n_traj = n_traj;
    r->info.d = d;
    r->info.q = s->cfg.q;
    r->info.D = D;
    r->info.lanes_per_traj = s->G;
    /* totals: reduced on the device, 16 bytes D2H */
    {
        unsigned long long* d_tot = nullptr;
        CUDA_TRY(cudaMallocAsync((void**)&d_tot, 16, s->stream));
        scope.track(d_tot);
        CUDA_TRY(cudaMemsetAsync(d_tot, 0, 16, s->stream));
        const unsigned blocks = (unsigned)std::min<long long>(((long long)N + 255) / 256, 1024);
        pn_sum_counts<<<blocks, 256, 0, s->stream>>>((long long)N, P.naccept, P.nreject, d_tot);
        unsigned long long tot[2] = {0, 0};
        CUDA_TRY(cudaMemcpyAsync(tot, d_tot, 16, cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        scope.release(d_tot);
        r->info.total_accepted = (int64_t)tot[0];
        r->info.total_rejected = (int64_t)tot[1];
        r->info.d2h_bytes += 16;
        if (need_count_pass) *need_count_pass = (fixed_steps_total >= 0 && (long long)tot[0] != fixed_steps_total);
    }
    return 0;
}

static int solve_common(pnode_solver* s, int64_t n_traj, const double* u0, const double* p, bool device_inputs,
                        pnode_result** out) {
    if (!s || !out) return fail(PNODE_ERR_ARGUMENT, "null argument");
    *out = nullptr;
    if (n_traj <= 0) return fail(PNODE_ERR_ARGUMENT, "trajectories must be positive");
    if (!u0) return fail(PNODE_ERR_ARGUMENT, "u0 is null");
    if (s->cfg.n_p > 0 && !p) return fail(PNODE_ERR_ARGUMENT, "p is null but n_p > 0");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    pnode_result* r = new pnode_result();
    r->s = s;
    r->n_traj = n_traj;
    memset(&r->info, 0, sizeof r->info);
    const double *d_u0 = u0, *d_p = p;
    if (!device_inputs) {
        const size_t bu = (size_t)n_traj * s->cfg.d * (s->cfg.second_order ? 2 : 1) * 8, bp = (size_t)n_traj * std::max(s->cfg.n_p, 0) * 8;
        auto t0 = std::chrono::steady_clock::now();
        cudaError_t e = cudaMallocAsync((void**)&r->d_u0, bu, s->stream);
        if (e == cudaSuccess && bp) e = cudaMallocAsync((void**)&r->d_p, bp, s->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(r->d_u0, u0, bu, cudaMemcpyHostToDevice, s->stream);
        if (e == cudaSuccess && bp) e = cudaMemcpyAsync(r->d_p, p, bp, cudaMemcpyHostToDevice, s->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
        if (e != cudaSuccess) {
            pnode_result_free(r);
            return fail(PNODE_ERR_CUDA, std::string("input staging: ") + cudaGetErrorString(e));
        }
        r->info.h2d_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        r->info.h2d_bytes = (int64_t)(bu + bp);
        r->owns_inputs = true;
        d_u0 = r->d_u0;
        d_p = r->d_p;
    }
    bool redo = false;
    int rc = solve_impl(s, n_traj, d_u0, d_p, r, false, &redo);
    if (!rc && redo) {
        /* a fixed-step trajectory stopped early: repeat with the counting pass so that the CSR offsets are exact */
        for (int i = 0; i < PNODE_F_COUNT; i++)
            if (r->f[i].d) {
                cudaFreeAsync(r->f[i].d, s->stream);
                r->f[i] = Buf();
            }
        r->offsets.clear();
        rc = solve_impl(s, n_traj, d_u0, d_p, r, true, nullptr);
    }
    if (rc) {
        std::string keep = g_err;
        pnode_result_free(r);
        g_err = keep;
        return rc;
    }
    *out = r;
    return 0;
}

extern "C" int pnode_solve_ensemble(pnode_solver* s, int64_t n_traj, const double* u0, const double* p,
                                    pnode_result** out) {
    return solve_common(s, n_traj, u0, p, false, out);
}
extern "C" int pnode_solve_ensemble_device(pnode_solver* s, int64_t n_traj, const double* d_u0, const double* d_p,
                                           pnode_result** out) {
    return solve_common(s, n_traj, d_u0, d_p, true, out);
}

extern "C" void pnode_shard_range(int64_t n_traj, int32_t shard, int32_t n_shards, int64_t* lo, int64_t* hi) {
    const int64_t base = n_traj / n_shards, rem = n_traj % n_shards;
    const int64_t l = shard * base + std::min<int64_t>(shard, rem);
    if (lo) *lo = l;
    if (hi) *hi = l + base + (shard < rem ? 1 : 0);
}

extern "C" int pnode_solve_ensemble_sharded(pnode_solver* const* solvers, int32_t n_solvers, int64_t n_traj, const double* u0,
                                            const double* p, pnode_result** results) {
    if (!solvers || !results || n_solvers < 1) return fail(PNODE_ERR_ARGUMENT, "null argument");
    if (n_traj < n_solvers) return fail(PNODE_ERR_ARGUMENT, "trajectories must be at least the number of devices");
    for (int i = 0; i < n_solvers; i++) {
        results[i] = nullptr;
        if (!solvers[i]) return fail(PNODE_ERR_ARGUMENT, "null solver");
        for (int j = 0; j < i; j++)
            if (solvers[j] == solvers[i]) return fail(PNODE_ERR_ARGUMENT, "a pnode_solver is not thread-safe: pass one solver per device");
        if (solvers[i]->cfg.d != solvers[0]->cfg.d || solvers[i]->cfg.n_p != solvers[0]->cfg.n_p ||
            solvers[i]->cfg.second_order != solvers[0]->cfg.second_order)
            return fail(PNODE_ERR_DIMENSION, "the solvers of one sharded solve must share the problem dimensions");
    }
    const int dim = solvers[0]->cfg.d * (solvers[0]->cfg.second_order ? 2 : 1), n_p = std::max(solvers[0]->cfg.n_p, 0);
    std::vector<int> rcs((size_t)n_solvers, 0);
    std::vector<std::string> errs((size_t)n_solvers);
    std::vector<std::thread> threads;
    for (int i = 0; i < n_solvers; i++)
        threads.emplace_back([&, i]() { /* one host thread (and the solver's own stream) per device */
            int64_t lo = 0, hi = 0;
            pnode_shard_range(n_traj, i, n_solvers, &lo, &hi);
            rcs[i] = solve_common(solvers[i], hi - lo, u0 + lo * dim, p ? p + lo * n_p : nullptr, false, &results[i]);
            if (rcs[i]) errs[i] = g_err; /* thread-local */
        });
    for (auto& t : threads) t.join();
    for (int i = 0; i < n_solvers; i++)
        if (rcs[i]) {
            for (int j = 0; j < n_solvers; j++)
                if (results[j]) {
                    pnode_result_free(results[j]);
                    results[j] = nullptr;
                }
            return fail(rcs[i], "shard " + std::to_string(i) + ": " + errs[i]);
        }
    return 0;
}

extern "C" int pnode_result_info_get(const pnode_result* r, pnode_result_info* info) {
    if (!r || !info) return fail(PNODE_ERR_ARGUMENT, "null argument");
    *info = r->info;
    return 0;
}
extern "C" int64_t pnode_result_field_bytes(const pnode_result* r, int field) {
    if (!r || field < 0 || field >= PNODE_F_COUNT) return 0;
    return (int64_t)r->f[field].bytes;
}
extern "C" const void* pnode_result_device_ptr(const pnode_result* r, int field) {
    if (!r || field < 0 || field >= PNODE_F_COUNT) return nullptr;
    return r->f[field].d;
}
extern "C" int pnode_result_copy(const pnode_result* r, int field, void* dst, size_t bytes) {
    if (!r || !dst) return fail(PNODE_ERR_ARGUMENT, "null argument");
    if (field < 0 || field >= PNODE_F_COUNT) return fail(PNODE_ERR_ARGUMENT, "unknown field");
    const Buf& b = r->f[field];
    if (!b.d) return fail(PNODE_ERR_ARGUMENT, "field was not produced by this solve");
    if (bytes != b.bytes) return fail(PNODE_ERR_DIMENSION, "destination size does not match the field size");
    CUDA_TRY(cudaSetDevice(r->s->cfg.device));
    CUDA_TRY(cudaMemcpyAsync(dst, b.d, bytes, cudaMemcpyDeviceToHost, r->s->stream));
    CUDA_TRY(cudaStreamSynchronize(r->s->stream));
    const_cast<pnode_result*>(r)->info.d2h_bytes += (int64_t)bytes;
    return 0;
}
extern "C" void pnode_result_free(pnode_result* r) {
    if (!r) return;
    cudaSetDevice(r->s->cfg.device);
    for (int i = 0; i < PNODE_F_COUNT; i++)
        if (r->f[i].d) cudaFreeAsync(r->f[i].d, r->s->stream);
    if (r->d_u0) cudaFreeAsync(r->d_u0, r->s->stream);
    if (r->d_p) cudaFreeAsync(r->d_p, r->s->stream);
    delete r;
}
