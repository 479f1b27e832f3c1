// pn_thread_host.cpp -- the thread-per-trajectory step kernels (csrc/kernels/pn_thread.cuh: the headline kernel of bench.py; pn_kron.cuh and
// pn_blocks.cuh: the structured layouts) compiled for the host under cuda_shim.h and run on one trajectory.  PN_EMU_PROBLEM / PN_EMU_Q / PN_EMU_ADAPTIVE / PN_EMU_DYNAMIC select the instantiation at
// compile time, as the NVRTC build does.  The caller (tests/test_kernel_host_emulation.py) passes the host-side constants of PnParams computed
// independently of pnode_api.cu (IWP factor L, its upper twin U, Qb = L'L, controller constants) and the initial state, and compares the
// outputs with the oracle and with the extended-precision fixtures.  Test infrastructure: nothing in the product includes this file.
#include "cuda_shim.h"
#define PN_THREAD_NT 1
#include "kernels/pn_thread.cuh"
#include "kernels/pn_kron.cuh"   /* PN_EMU_KERNEL 1: EK0 on the IsometricKronecker layout, one thread per trajectory */
#include "kernels/pn_blocks.cuh" /* PN_EMU_KERNEL 2: BlocksOfDiagonals layout (PN_EMU_DIAG1: DiagonalEK1, else EK0; PN_EMU_MV: multivariate diffusion) */
#include "kernels/pn_builtin_problems.cuh"
#ifdef PN_EMU_USER_SOURCE
#include PN_EMU_USER_SOURCE
#endif

double pn_thread_shared[8192];

#ifndef PN_EMU_ADAPTIVE
#define PN_EMU_ADAPTIVE 0
#endif
#ifndef PN_EMU_DYNAMIC
#define PN_EMU_DYNAMIC 0
#endif
#ifndef PN_EMU_KERNEL
#define PN_EMU_KERNEL 0
#endif
#ifndef PN_EMU_DIAG1
#define PN_EMU_DIAG1 0
#endif
#ifndef PN_EMU_MV
#define PN_EMU_MV 0
#endif

extern "C" int pn_emu_d(void) { return PN_EMU_PROBLEM::d; }
extern "C" int pn_emu_n_p(void) { return PN_EMU_PROBLEM::n_p; }
extern "C" int pn_emu_q(void) { return PN_EMU_Q; }

// scal_in : t0, t1, dt0 (fixed step, or the first step of an adaptive solve), abstol, reltol, dtmin, dtmax, beta1, beta2, qmin, qmax, gamma,
//           qsteady_min, qsteady_max, qoldinit, initial_diffusion                                                     (16 doubles)
// scal_out: t_end, loglik, diffusion, dt_last                                                                          (4 doubles)
// counts  : naccept, nreject, nf, retcode                                                                              (4 x int64)
extern "C" void pn_emu_solve(const double* scal_in, int calibrate, const double* L, const double* U, const double* Qb, const double* init_mu,
                             const double* p, double* u_mean, double* u_cov, double* x_mean, double* x_covfac, double* scal_out,
                             long long* counts, const double* forced_diffusion, long long n_forced, double* diffusion_mv) {
    constexpr int n = PN_EMU_Q + 1;
    PnParams P;
    std::memset(&P, 0, sizeof P);
    unsigned long long counter = 0;
    double init_dt = scal_in[2];
    long long naccept = 0, nreject = 0, nf = 0;
    int retcode = -1;
    P.n_traj = 1;
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
    P.init_dt = &init_dt;
    P.work_counter = &counter;
    P.forced_diffusion = forced_diffusion; /* test hook of the product: sigma^2 of accepted step k (null: the local estimate) */
    P.n_forced = n_forced;
    for (int i = 0; i < n * n; i++) P.L[i] = L[i], P.U[i] = U[i], P.Qb[i] = Qb[i];
    P.t_end = &scal_out[0], P.loglik = &scal_out[1], P.diffusion = &scal_out[2], P.dt_last = &scal_out[3];
    P.u_mean = u_mean, P.u_cov = u_cov, P.x_mean = x_mean, P.x_covfac = x_covfac;
    P.naccept = &naccept, P.nreject = &nreject, P.nf = &nf, P.retcode = &retcode;
    P.diffusion_mv = diffusion_mv;
#if PN_EMU_KERNEL == 0
    pn_thread_entry<PN_EMU_PROBLEM, PN_EMU_Q, (PN_EMU_ADAPTIVE != 0), (PN_EMU_DYNAMIC != 0)>(P);
#elif PN_EMU_KERNEL == 1
    pn_kron_entry<PN_EMU_PROBLEM, PN_EMU_Q, (PN_EMU_ADAPTIVE != 0), (PN_EMU_DYNAMIC != 0), 0>(P);
#else
    pn_blocks_entry<PN_EMU_PROBLEM, PN_EMU_Q, (PN_EMU_DIAG1 != 0), (PN_EMU_MV != 0), (PN_EMU_ADAPTIVE != 0), (PN_EMU_DYNAMIC != 0), 0>(P);
#endif
    counts[0] = naccept, counts[1] = nreject, counts[2] = nf, counts[3] = retcode;
}
