// pn_warp_host.cpp -- the group-per-trajectory kernels (csrc/kernels/pn_small.cuh: the step kernel of the saving / smoothing solves;
// pn_smooth_col.cuh and pn_smooth_reg.cuh: the register-resident RTS sweeps) compiled for the host under cuda_shim.h in PN_SHIM_WARP mode and
// run as ONE WARP of host threads: every lane is a pthread, shuffles / votes / __syncwarp go through pthread barriers, the kernels' shared
// memory is one array all lanes see.  32 / G trajectories are in flight at a time and the groups pull more from the kernels' own atomic work
// queue.  PN_EMU_PROBLEM / PN_EMU_Q / PN_EMU_G / PN_EMU_GS / PN_EMU_ADAPTIVE / PN_EMU_DYNAMIC select the instantiation at compile time, as
// the NVRTC build does.  Built with -fsanitize=thread the same program is a racecheck of the kernels' shared-memory protocol.
// The caller (tests/test_kernel_host_emulation.py) computes the host-side constants of the parameter blocks independently of pnode_api.cu and
// compares the outputs with the oracle.  Test infrastructure: nothing in the product includes this file.
#define PN_SHIM_WARP 1
#include "cuda_shim.h"
#include <cstdlib>
#include <vector>

thread_local PnShimDim3 threadIdx{0, 0, 0};
PnShimDim3 blockDim{32, 1, 1};
PnShimCta pn_shim_cta;

#ifndef PN_EMU_G
#define PN_EMU_G 4
#endif
#ifndef PN_EMU_GS
#define PN_EMU_GS 8
#endif
#define PN_THREADS 32     /* one warp is the CTA */
#define PN_SMR_THREADS 32
#ifndef PN_EMU_KERNEL
#define PN_EMU_KERNEL 0 /* 0: pn_small.cuh (G lanes per trajectory, one warp); 1: pn_cta.cuh (one CTA of PN_CTA_THREADS threads per trajectory) */
#endif
#include "kernels/pn_small.cuh"
#include "kernels/pn_smooth.cuh" /* the warp-per-trajectory sweep (shared memory up to D = 56, a global workspace beyond: Pleiades D = 112) */
#include "kernels/pn_smooth_col.cuh"
#include "kernels/pn_smooth_reg.cuh"
#include "kernels/pn_dense.cuh" /* dense output sol(t) on a saveat grid: register predict per (trajectory, grid point) pair */
#ifndef PN_EMU_GD
#define PN_EMU_GD 2 /* lanes per pair of the dense-output kernel (pnode_api.cu auto_lanes_qr(D)) */
#endif
#if PN_EMU_KERNEL == 1
#include "kernels/pn_cta.cuh"
alignas(16) double pn_cta_sm[65536];
#endif
#if PN_EMU_KERNEL == 2 /* pn_gen.cuh: general LTI prior (IOUP with a vector / matrix / Jacobian-updated rate), one warp per trajectory */
#include "kernels/pn_gen.cuh"
alignas(16) double pn_gen_shared[65536];
#endif
static const double* g_rate_matrix = nullptr; /* [d][d] row-major (PN_EMU_KERNEL 2) */
static int g_update_rate = 0;
extern "C" void pn_emu_set_rate(const double* rate, int update_rate) { g_rate_matrix = rate, g_update_rate = update_rate; }
/* Fenrir mode of the sweeps (built with PN_FENRIR=1 for the column sweep): observations, observation model, output [n_traj] */
static const double *g_data_t = nullptr, *g_data_u = nullptr, *g_obs_H = nullptr, *g_obs_Rn = nullptr;
static double* g_fenrir_ll = nullptr;
static long long g_n_data = 0;
static int g_n_obs = 0, g_kron_logdet = 0;
extern "C" void pn_emu_set_fenrir(const double* data_t, const double* data_u, long long n_data, int n_obs, const double* obs_H, const double* obs_Rn,
                                  double* fenrir_ll, int kron_logdet) {
    g_data_t = data_t, g_data_u = data_u, g_n_data = n_data, g_n_obs = n_obs, g_obs_H = obs_H, g_obs_Rn = obs_Rn, g_fenrir_ll = fenrir_ll;
    g_kron_logdet = kron_logdet;
}
/* forward data updates (DataUpdateCallback; pn_small.cuh built with a PN_DATA preamble): observations, which are tstops, and the output [n_traj] */
static const double *g_fwd_t = nullptr, *g_fwd_u = nullptr;
static long long g_fwd_n = 0;
static double* g_fwd_ll = nullptr;
extern "C" void pn_emu_set_data(const double* data_t, const double* data_u, long long n_data, double* data_loglik) {
    g_fwd_t = data_t, g_fwd_u = data_u, g_fwd_n = n_data, g_fwd_ll = data_loglik;
}
/* dense output of a filter-only solve (solution.jl:330-398 without the smoothing half): grid, outputs [n_traj][n_saveat][d] and [..][d][d] */
static const double* g_saveat = nullptr;
static long long g_n_saveat = 0;
static double *g_dense_u = nullptr, *g_dense_cov = nullptr;
extern "C" void pn_emu_set_saveat(const double* saveat, long long n, double* out_u, double* out_cov) {
    g_saveat = saveat, g_n_saveat = n, g_dense_u = out_u, g_dense_cov = out_cov;
}
static double g_rate = 0.0, g_glx[16], g_glw[16]; /* scalar-rate IOUP of the register kernels (built with PN_IOUP=1): rate, Gauss-Legendre rule on [0, 1] */
extern "C" void pn_emu_set_ioup(double rate, const double* glx, const double* glw) {
    g_rate = rate;
    for (int i = 0; i < 16; i++) g_glx[i] = glx[i], g_glw[i] = glw[i];
}
#include "kernels/pn_builtin_problems.cuh"
#ifdef PN_EMU_USER_SOURCE /* a traced right-hand side (tracer.cuda_source), as the NVRTC build appends it */
#include PN_EMU_USER_SOURCE
#endif

#ifndef PN_EMU_ADAPTIVE
#define PN_EMU_ADAPTIVE 0
#endif
#ifndef PN_EMU_DYNAMIC
#define PN_EMU_DYNAMIC 0
#endif
#ifndef PN_EMU_SAVE
#define PN_EMU_SAVE 2
#endif

alignas(16) double pn_small_shared[32768];
alignas(16) double pn_smc_shared[32768];
alignas(16) double pn_smr_shared[32768];
alignas(16) double pn_sm_shared[32768];

/* modelled dimension: Prob is the first-order form of dimension 2 d for a SecondOrderODEProblem (PN_SECOND), sol.u = (du, u) keeps length Prob::d */
static constexpr int PN_EMU_DM = PN_SECOND ? PN_EMU_PROBLEM::d / 2 : PN_EMU_PROBLEM::d;
extern "C" int pn_emu_d(void) { return PN_EMU_PROBLEM::d; }
extern "C" int pn_emu_n_p(void) { return PN_EMU_PROBLEM::n_p; }
extern "C" int pn_emu_q(void) { return PN_EMU_Q; }
extern "C" int pn_emu_g(void) { return PN_EMU_G; }
extern "C" int pn_emu_gs(void) { return PN_EMU_GS; }

template <class F>
static void run_warp(int threads, F body) { /* a CTA of `threads` host threads: ceil(threads / 32) warps */
    const int warps = (threads + 31) / 32;
    pn_shim_cta.threads = threads;
    blockDim.x = (unsigned)threads;
    pthread_barrier_init(&pn_shim_cta.bar, nullptr, (unsigned)threads);
    for (int w = 0; w < warps; w++) {
        pn_shim_cta.warp[w].lanes = (threads - 32 * w < 32) ? threads - 32 * w : 32;
        pthread_barrier_init(&pn_shim_cta.warp[w].bar, nullptr, (unsigned)pn_shim_cta.warp[w].lanes);
        for (int lg = 0; lg < 4; lg++)
            for (int g = 0; g < (32 >> (lg + 1)); g++) pthread_barrier_init(&pn_shim_cta.warp[w].sub[lg][g], nullptr, 2u << lg);
    }
    struct Arg {
        F* body;
        int tid;
    };
    std::vector<pthread_t> th(threads);
    std::vector<Arg> args(threads);
    for (int l = 0; l < threads; l++) {
        args[l] = Arg{&body, l};
        pthread_create(&th[l], nullptr, [](void* a) -> void* {
            Arg* x = (Arg*)a;
            threadIdx.x = (unsigned)x->tid;
            (*x->body)();
            return nullptr;
        }, &args[l]);
    }
    for (int l = 0; l < threads; l++) pthread_join(th[l], nullptr);
    pthread_barrier_destroy(&pn_shim_cta.bar);
    for (int w = 0; w < warps; w++) {
        pthread_barrier_destroy(&pn_shim_cta.warp[w].bar);
        for (int lg = 0; lg < 4; lg++)
            for (int g = 0; g < (32 >> (lg + 1)); g++) pthread_barrier_destroy(&pn_shim_cta.warp[w].sub[lg][g]);
    }
}

// scal_in : as pn_thread_host.cpp (16 doubles).  Per-trajectory inputs: init_mu [n_traj][D], init_dt [n_traj], p [n_traj][n_p].
// Per-trajectory outputs: u_mean [n_traj][d], u_cov [n_traj][d][d], scal_out [n_traj][4] (t_end, loglik, diffusion, dt_last), counts
// [n_traj][4] (naccept, nreject, nf, retcode).  With PN_EMU_SAVE >= 1 the CSR records (offsets [n_traj + 1]; s_t, s_u_mean, s_u_cov,
// s_diffusion; with SAVE = 2 also s_x_mean, s_x_covfac, s_h) are written; smooth = 1 / 2 then runs the column- / row-distributed sweep
// over them (in place, like the product: the records and s_u_mean / s_u_cov become the smoothed ones).
extern "C" void pn_emu_warp_solve(long long n_traj, int lanes, const double* scal_in, int calibrate, const double* L, const double* U,
                                  const double* Qb, const double* init_mu, const double* init_dt, const double* p, double* u_mean,
                                  double* u_cov, double* scal_out, long long* counts, const long long* offsets, double* s_t,
                                  double* s_u_mean, double* s_u_cov, double* s_diffusion, double* s_x_mean, double* s_x_covfac, double* s_h,
                                  int smooth, int write_x) {
    constexpr int n = PN_EMU_Q + 1;
    PnParams P;
    std::memset(&P, 0, sizeof P);
    unsigned long long counter = 0;
    std::vector<double> t_end(n_traj), loglik(n_traj), diffusion(n_traj), dt_last(n_traj);
    std::vector<long long> naccept(n_traj), nreject(n_traj), nf(n_traj);
    std::vector<int> retcode(n_traj, -1);
    P.n_traj = n_traj;
    P.p = p;
    P.t0 = scal_in[0], P.t1 = scal_in[1], P.dt0 = scal_in[2];
    P.abstol = scal_in[3], P.reltol = scal_in[4], P.dtmin = scal_in[5], P.dtmax = scal_in[6];
    P.beta1 = scal_in[7], P.beta2 = scal_in[8], P.qmin = scal_in[9], P.qmax = scal_in[10], P.gamma = scal_in[11];
    P.qsteady_min = scal_in[12], P.qsteady_max = scal_in[13], P.qoldinit = scal_in[14];
    P.qoldinit_pb2 = std::pow(P.qoldinit, P.beta2);
    P.inv_qmax = 1.0 / P.qmax, P.inv_qmin = 1.0 / P.qmin, P.inv_gamma = 1.0 / P.gamma;
    P.initial_diffusion = scal_in[15];
    P.maxiters = 1000000;
    P.calibrate = calibrate;
    P.order = PN_EMU_Q;
    P.traj_stride = 1;
    P.init_mu = init_mu;
    P.init_dt = init_dt;
    P.work_counter = &counter;
    for (int i = 0; i < n * n; i++) P.L[i] = L[i], P.U[i] = U[i], P.Qb[i] = Qb[i];
    if (g_fwd_ll) {
        P.data_t = g_fwd_t, P.data_u = g_fwd_u, P.n_data = g_fwd_n, P.data_loglik = g_fwd_ll;
        P.tstops = g_fwd_t, P.n_tstops = g_fwd_n; /* filtering.jl:24: the data times are tstops (those at t1 are harmless) */
    }
    P.rate = g_rate;
    for (int i = 0; i < 16; i++) P.glx[i] = g_glx[i], P.glw[i] = g_glw[i];
    P.t_end = t_end.data(), P.loglik = loglik.data(), P.diffusion = diffusion.data(), P.dt_last = dt_last.data();
    P.u_mean = u_mean, P.u_cov = u_cov;
    P.naccept = naccept.data(), P.nreject = nreject.data(), P.nf = nf.data(), P.retcode = retcode.data();
    if (PN_EMU_SAVE >= 1) {
        P.step_offsets = offsets;
        P.s_t = s_t, P.s_u_mean = s_u_mean, P.s_u_cov = s_u_cov, P.s_diffusion = s_diffusion;
        if (PN_EMU_SAVE >= 2) P.s_x_mean = s_x_mean, P.s_x_covfac = s_x_covfac, P.s_h = s_h;
    }
#if PN_EMU_KERNEL == 2
    (void)lanes;
    P.rate_matrix = g_rate_matrix, P.update_rate = g_update_rate;
    if (pn_gen_ws_doubles(PN_EMU_PROBLEM::d, n) * PN_GEN_WARPS > (int)(sizeof pn_gen_shared / 8)) std::abort();
    run_warp(32 * PN_GEN_WARPS, [&]() { pn_gen_entry<PN_EMU_PROBLEM, PN_EMU_Q, (PN_EMU_ADAPTIVE != 0), (PN_EMU_DYNAMIC != 0), PN_EMU_SAVE>(P); });
#elif PN_EMU_KERNEL == 1
    (void)lanes;
    run_warp(PN_CTA_THREADS, [&]() { pn_cta_entry<PN_EMU_PROBLEM, PN_EMU_Q, (PN_EMU_ADAPTIVE != 0), (PN_EMU_DYNAMIC != 0), PN_EMU_SAVE>(P); });
#else
    run_warp(lanes, [&]() { pn_small_entry<PN_EMU_PROBLEM, PN_EMU_Q, PN_EMU_G, (PN_EMU_ADAPTIVE != 0), (PN_EMU_DYNAMIC != 0), PN_EMU_SAVE>(P); });
#endif
    for (long long i = 0; i < n_traj; i++) {
        scal_out[4 * i + 0] = t_end[i], scal_out[4 * i + 1] = loglik[i], scal_out[4 * i + 2] = diffusion[i], scal_out[4 * i + 3] = dt_last[i];
        counts[4 * i + 0] = naccept[i], counts[4 * i + 1] = nreject[i], counts[4 * i + 2] = nf[i], counts[4 * i + 3] = retcode[i];
    }
#if PN_EMU_SAVE >= 2 && PN_EMU_KERNEL == 0
    if (g_saveat && !smooth) { /* pnode_api.cu solve_impl, dense && !smooth: the predict kernel writes the filtering marginals directly */
        PnDenseParams Dn;
        std::memset(&Dn, 0, sizeof Dn);
        Dn.n_traj = n_traj, Dn.n_saveat = g_n_saveat, Dn.saveat = g_saveat;
        Dn.smooth = 0;
        Dn.calibrate_scale = (calibrate && !(PN_EMU_DYNAMIC != 0)) ? 1 : 0;
        Dn.step_offsets = offsets;
        Dn.s_t = s_t, Dn.s_h = s_h, Dn.s_diffusion = s_diffusion, Dn.s_x_mean = s_x_mean, Dn.s_x_covfac = s_x_covfac;
        Dn.gdiff = diffusion.data();
        Dn.initial_diffusion = P.initial_diffusion;
        for (int i = 0; i < n * n; i++) Dn.U[i] = U[i];
        Dn.rate = g_rate;
        for (int i = 0; i < 16; i++) Dn.glx[i] = g_glx[i], Dn.glw[i] = g_glw[i];
        Dn.out_u = g_dense_u, Dn.out_cov = g_dense_cov;
        run_warp(lanes, [&]() { pn_dense_predict_entry<PN_EMU_DM, PN_EMU_Q, PN_EMU_GD>(Dn); });
    }
#endif
#if PN_EMU_SAVE >= 2
    if (smooth) {
        PnSmoothParams S;
        std::memset(&S, 0, sizeof S);
        counter = 0;
        S.n_traj = n_traj;
        S.d = PN_EMU_DM, S.q = PN_EMU_Q;
        S.dynamic = (PN_EMU_DYNAMIC != 0);
        S.calibrate = calibrate;
        S.initial_diffusion = P.initial_diffusion;
        S.step_offsets = offsets;
        S.s_h = s_h, S.s_diffusion = s_diffusion, S.gdiff = diffusion.data();
        S.s_x_mean = s_x_mean, S.s_x_covfac = s_x_covfac, S.s_u_mean = s_u_mean, S.s_u_cov = s_u_cov;
        for (int i = 0; i < n * n; i++) S.L[i] = L[i], S.U[i] = U[i];
        S.work_counter = &counter;
        S.write_x = write_x;
        S.rate = g_rate;
        for (int i = 0; i < 16; i++) S.glx[i] = g_glx[i], S.glw[i] = g_glw[i];
        if (g_fenrir_ll) { /* pnode_api.cu solve_impl: uncalibrated, the records stay the filtering estimates */
            S.calibrate = 0;
            S.s_t = s_t, S.data_t = g_data_t, S.data_u = g_data_u, S.obs_H = g_obs_H, S.obs_Rn = g_obs_Rn, S.n_data = g_n_data, S.n_obs = g_n_obs;
            S.retcode = retcode.data(), S.fenrir_ll = g_fenrir_ll, S.kron_logdet = g_kron_logdet;
        }
#if PN_EMU_KERNEL == 2
        S.gen_prior = 1, S.rate_matrix = g_rate_matrix;
#endif
        if (smooth == 3) { /* pn_smooth_kernel: PN_SM_WARPS warps per CTA, one trajectory per warp; work matrices in shared memory when the four
                              warps' share fits the device's 200 KB, else in a global workspace -- the product's rule (pnode_api.cu launch_sweep) */
            constexpr int D = PN_EMU_DM * n;
            const size_t per_warp = 7 * D * D + 4 * D + 2 * n * n + 2 * n + (S.gen_prior ? PN_SMOOTH_GEN_EXTRA(D, PN_EMU_DM, n) : 0) +
                                    (S.fenrir_ll ? PN_FENRIR_EXTRA(D, S.n_obs) : 0);
            std::vector<double> ws;
            if (per_warp * 8 > 200 * 1024) {
                ws.resize(per_warp * PN_SM_WARPS);
                S.workspace = ws.data();
            } else if (per_warp * PN_SM_WARPS > sizeof pn_sm_shared / 8) {
                std::abort();
            }
            run_warp(32 * PN_SM_WARPS, [&]() { pn_smooth_kernel(S); });
        }
#if PN_EMU_KERNEL == 0
        else if (smooth == 1)
            run_warp(lanes, [&]() { pn_smooth_col_entry<PN_EMU_DM, PN_EMU_Q, PN_EMU_GS>(S); });
        else
            run_warp(lanes, [&]() { pn_smooth_reg_entry<PN_EMU_DM, PN_EMU_Q, PN_EMU_GS>(S); });
#endif
    }
#endif
}

#ifdef PN_EMU_MAIN
// Stand-alone form for ThreadSanitizer (a sanitized library cannot be loaded into an unsanitized python): reads one ensemble from a file of
// doubles written by the test -- N, lanes, smooth, write_x, saved points per trajectory, scal_in[16], L, U, Qb, init_mu, init_dt, p -- runs the
// step kernel and the sweep, prints a checksum.  ThreadSanitizer's report (and its exit code 66) is the result.
#include <cstdio>
int main(int argc, char** argv) {
    if (argc < 2) return 2;
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    std::vector<double> in;
    double buf[1024];
    size_t got;
    while ((got = std::fread(buf, 8, 1024, f)) > 0) in.insert(in.end(), buf, buf + got);
    std::fclose(f);
    constexpr int n = PN_EMU_Q + 1, d = PN_EMU_PROBLEM::d, D = PN_EMU_DM * n, np_ = PN_EMU_PROBLEM::n_p;
    const double* c = in.data();
    const long long N = (long long)c[0];
    const int lanes = (int)c[1], smooth = (int)c[2], write_x = (int)c[3];
    const long long ns = (long long)c[4];
    c += 5;
    const double* scal = c;
    c += 16;
    const double *L = c, *U = c + n * n, *Qb = c + 2 * n * n;
    c += 3 * n * n;
    const double* init_mu = c;
    c += N * D;
    const double* init_dt = c;
    c += N;
    const double* p = c;
    c += N * (np_ > 0 ? np_ : 1);
    /* optional tagged sections: 1 rate matrix (update flag, d x d) | 2 scalar-rate IOUP (rate, glx[16], glw[16]) | 3 Fenrir (n_data, n_obs, data_t,
       data_u, H [n_obs][D], Rn [n_obs][n_obs]) | 4 forward data updates (n_data, n_obs, data_t, data_u) */
    std::vector<double> fll(N, 0.0);
    while (c - in.data() < (long long)in.size()) {
        const int tag = (int)*c++;
        if (tag == 1) {
            pn_emu_set_rate(c[1] < 0 ? nullptr : c + 2, (int)c[0]);
            c += 2 + d * d;
        } else if (tag == 2) {
            pn_emu_set_ioup(c[0], c + 1, c + 17);
            c += 33;
        } else if (tag == 3) {
            const long long nd = (long long)c[0];
            const int no = (int)c[1];
            pn_emu_set_fenrir(c + 2, c + 2 + nd, nd, no, c + 2 + nd + nd * no, c + 2 + nd + nd * no + (long long)no * D, fll.data(), 0);
            c += 2 + nd + nd * no + (long long)no * D + no * no;
        } else if (tag == 4) { /* forward data updates: n_data, n_obs, data_t, data_u */
            const long long nd = (long long)c[0];
            const int no = (int)c[1];
            pn_emu_set_data(c + 2, c + 2 + nd, nd, fll.data());
            c += 2 + nd + nd * no;
        } else {
            return 3;
        }
    }
    if (c - in.data() != (long long)in.size()) return 3;
    std::vector<long long> off(N + 1), counts(4 * N);
    for (long long i = 0; i <= N; i++) off[i] = i * ns;
    const long long T = off[N];
    std::vector<double> um(N * d), uc(N * d * d), so(4 * N), s_t(T), s_u(T * d), s_c(T * d * d), s_df(T), s_x(T * D), s_xr(T * D * D), s_h(T);
    pn_emu_warp_solve(N, lanes, scal, 0, L, U, Qb, init_mu, init_dt, p, um.data(), uc.data(), so.data(), counts.data(), off.data(), s_t.data(),
                      s_u.data(), s_c.data(), s_df.data(), s_x.data(), s_xr.data(), s_h.data(), smooth, write_x);
    double sum = 0.0;
    for (double v : um) sum += v;
    for (double v : fll) sum += v;
    for (double v : s_u) sum += v;
    long long acc = 0;
    for (long long i = 0; i < N; i++) acc += counts[4 * i];
    std::printf("accepted %lld checksum %.17g\n", acc, sum);
    return 0;
}
#endif
