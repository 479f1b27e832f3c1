// pn_kron.cuh -- EK0 + IWP + scalar diffusion: the IsometricKroneckerProduct layout (src/kronecker.jl:15-35,
// src/covariance_structure.jl) where every D x D matrix is B (x) I_d and only the (q+1) x (q+1) factor B is stored.
//
// One THREAD per trajectory: the factor R_B (n x n), the 2n x n QR stack and the n x d mean all live in the
// thread's registers, so a step needs no shuffles, no shared memory and no replicated work; the warp simply runs 32
// independent adaptive time loops and each lane pulls its next trajectory from the atomic work counter.
//
//   reference routine (file:line)                                       here (per thread, n x n factors)
//   predict_cov! Kronecker   src/filtering/predict.jl:97-117             Householder QR of [R_B T' ; sqrt(s2) L P^-1]
//   update! Kronecker        src/filtering/update.jl:151-195             S scalar, K = R^-'R^-[:,1]/S, mean as n x d matrix
//   invquad Kronecker        src/diffusions/calibration.jl:16-20         s2 = z'z / (W d), W = |Qh.R[:,1]|^2
//   diag! Kronecker          src/perform_step.jl:250-251                 e_i = h sqrt(s2 W) for all i
//   log-likelihood quirk     src/filtering/update.jl:147,182-194         logdet is NOT multiplied by d (replicated as is)
//
// NVRTC-compatible: no standard headers.
#ifndef PN_KRON_CUH
#define PN_KRON_CUH
#include "pn_params.h"
#include "pn_taylor.cuh"
#include "pn_small.cuh" /* pn_rcp, pn_rsqrt */

template <class Prob, int Q, bool ADAPTIVE, bool DYNAMIC, int SAVE>
struct PnKron {
    static constexpr int DP = Prob::d;                  /* dimension of the right-hand side */
    static constexpr int d = PN_SECOND ? DP / 2 : DP;   /* second-order problems (PN_SECOND, see pn_small.cuh): positions */
    static constexpr int DU = DP;                       /* length of sol.u = (du, u) */
    static constexpr int E1B = PN_SECOND ? 2 : 1;       /* H.B = e_2' (derivative_utils.jl:39-41) / e_1' */
    static constexpr int SB = PN_SECOND ? 1 : 0;        /* block the error estimate is scaled with (du / u) */
    static constexpr int n = Q + 1;
    static constexpr int D = d * n;
    static constexpr int NPA = Prob::n_p > 0 ? Prob::n_p : 1;

    // thread-local Householder QR of [Rt ; Bt] (2n x n); Rt returns upper triangular
    static PN_HD void qr_stack(double (&Rt)[n][n], double (&Bt)[n][n]) {
#pragma unroll
        for (int k = 0; k < n; k++) {
            double total = 0.0;
#pragma unroll
            for (int r = 0; r < n; r++) {
                if (r >= k) total += Rt[r][k] * Rt[r][k];
                total += Bt[r][k] * Bt[r][k];
            }
            const double alpha = Rt[k][k];
            const bool nz = total > 0.0;
            const double nrm = nz ? total * pn_rsqrt(total) : 0.0;
            const double beta = -copysign(nrm, alpha);
            const double vk = alpha - beta;
            const double gam = nz ? pn_rcp(total + fabs(alpha) * nrm) : 0.0;
#pragma unroll
            for (int j = 0; j < n; j++) {
                if (j <= k) continue;
                double s = vk * Rt[k][j];
#pragma unroll
                for (int r = 0; r < n; r++) {
                    if (r > k) s += Rt[r][k] * Rt[r][j];
                    s += Bt[r][k] * Bt[r][j];
                }
                s *= gam;
                Rt[k][j] -= s * vk;
#pragma unroll
                for (int r = 0; r < n; r++) {
                    if (r > k) Rt[r][j] -= s * Rt[r][k];
                    Bt[r][j] -= s * Bt[r][k];
                }
            }
            Rt[k][k] = beta;
#pragma unroll
            for (int r = 0; r < n; r++)
                if (r > k) Rt[r][k] = 0.0;
        }
    }

    /* sol.u = SolProj x and its covariance: E0 rows (c00 I), or (E1; E0) rows for second-order problems
       ([c11 I, c10 I; c10 I, c00 I]); cjk = sum_r RB[r][j] RB[r][k] */
    static __device__ __forceinline__ void write_pu(double* um, double* uc, const double (&M)[n][d], const double (&RB)[n][n], double sc) {
        double c00 = 0.0, c01 = 0.0, c11 = 0.0;
#pragma unroll
        for (int r = 0; r < n; r++) {
            c00 += RB[r][0] * RB[r][0];
            if (PN_SECOND) {
                c01 += RB[r][0] * RB[r][1];
                c11 += RB[r][1] * RB[r][1];
            }
        }
#pragma unroll
        for (int j = 0; j < DU; j++) um[j] = PN_SECOND ? (j < d ? M[1][j < d ? j : 0] : M[0][j >= d ? j - d : 0]) : M[0][j < d ? j : 0];
#pragma unroll
        for (int a = 0; a < DU; a++)
#pragma unroll
            for (int b = 0; b < DU; b++) {
                double v = 0.0;
                if (PN_SECOND) {
                    const int ia = a % d, ib = b % d;
                    if (ia == ib) v = (a < d && b < d) ? c11 : ((a >= d && b >= d) ? c00 : c01);
                } else {
                    v = (a == b) ? c00 : 0.0;
                }
                uc[a * DU + b] = v * sc;
            }
    }

    static __device__ void run(const PnParams& P) {
        double M[n][d];   /* mean as n x d matrix: M[j][i] = mu[j*d + i] */
        double RB[n][n];
        double pr[NPA];
        for (;;) {
            const unsigned long long idx = atomicAdd(P.work_counter, 1ULL);
            if ((long long)idx * P.traj_stride >= P.n_traj) break;
            const long long traj = (long long)idx * P.traj_stride;
            double t = P.t0, dt, loglik = 0.0, gdiff = 0.0, ldiff = 0.0, qold = P.qoldinit, qold_pb2 = P.qoldinit_pb2;
            long long naccept = 0, nreject = 0, nf = Q, iter = 0, tstop_idx = 0, save_pos = 0;
            int retcode = PN_RET_SUCCESS;
            double lgm = 1.0; /* lazily accumulated sum of log S (pn_logacc_mul, pn_small.cuh) */
            int lge = 0;
            double uprev[d];
            {
                /* initial state from pn_init_kernel (pre-pass over the whole ensemble: TaylorModeInit + initial step) */
#pragma unroll
                for (int i = 0; i < Prob::n_p; i++) pr[i] = P.p[traj * Prob::n_p + i];
#pragma unroll
                for (int j = 0; j < n; j++)
#pragma unroll
                    for (int i = 0; i < d; i++) M[j][i] = P.init_mu[traj * D + j * d + i];
#pragma unroll
                for (int i = 0; i < d; i++) uprev[i] = M[SB][i];
                dt = P.init_dt[traj];
                if (ADAPTIVE && !(P.dt0 > 0.0) && !PN_MASS) nf += 2;
            }
            const double dtcache = dt;
#pragma unroll
            for (int a = 0; a < n; a++)
#pragma unroll
                for (int b = 0; b < n; b++) RB[a][b] = 0.0;
            if (SAVE >= 1) {
                save_pos = P.step_offsets ? P.step_offsets[traj] : 0;
                if (P.step_offsets) {
                    P.s_t[save_pos] = t;
                    P.s_diffusion[save_pos] = 0.0;
                    write_pu(P.s_u_mean + save_pos * DU, P.s_u_cov + save_pos * DU * DU, M, RB, 1.0);
                    if (SAVE >= 2) save_state(P, save_pos, M, RB, 0.0, 0.0);
                }
                save_pos += 1;
            }

            while (t < P.t1 && retcode == PN_RET_SUCCESS) {
                double tstop = P.t1;
                while (tstop_idx < P.n_tstops && P.tstops[tstop_idx] <= t) tstop_idx++;
                if (tstop_idx < P.n_tstops && P.tstops[tstop_idx] < P.t1) tstop = P.tstops[tstop_idx];
                iter++;
                if (iter > P.maxiters) {
                    retcode = PN_RET_MAXITERS;
                    break;
                }
                if (ADAPTIVE) {
                    dt = fmin(dt, P.dtmax);
                    dt = fmax(dt, P.dtmin);
                    dt = fmin(dt, tstop - t);
                } else {
                    dt = fmin(dtcache, tstop - t);
                }
                const double h = dt;
                if (h != h || !(t + h > t) || (ADAPTIVE && h <= P.dtmin)) {
                    retcode = PN_RET_DTLESSTHANMIN;
                    break;
                }
                const double tnew = t + h;
                double hp[n], PIv[n];
                hp[0] = 1.0;
#pragma unroll
                for (int k = 1; k < n; k++) hp[k] = hp[k - 1] * h / (double)k;
                const double sqh = sqrt(h);
#pragma unroll
                for (int c = 0; c < n; c++) PIv[c] = hp[Q - c] * sqh;
                /* predicted mean (vec trick, src/kronecker.jl:193-223): M^- = T M */
                double Mp[n][d];
#pragma unroll
                for (int j = 0; j < n; j++)
#pragma unroll
                    for (int i = 0; i < d; i++) {
                        double s = M[j][i];
#pragma unroll
                        for (int k = 0; k < n; k++)
                            if (k > j) s += hp[k > j ? k - j : 0] * M[k][i];
                        Mp[j][i] = s;
                    }
#if PN_SECOND
                double fu[DP], z[d], x2[DP];
#pragma unroll
                for (int i = 0; i < d; i++) {
                    x2[i] = Mp[1][i]; /* (du, u) */
                    x2[d + i] = Mp[0][i];
                }
                Prob::template f<double>(fu, x2, pr, tnew);
#else
                double fu[d], z[d];
                Prob::template f<double>(fu, Mp[0], pr, tnew);
#endif
                nf += 1;
                double zz = 0.0;
#pragma unroll
                for (int i = 0; i < d; i++) {
#if PN_MASS
                    z[i] = pn_mass(0, 0) * Mp[1][i] - fu[i]; /* M = lambda I: H.B = lambda e_1' (derivative_utils.jl:45-47) */
#else
                    z[i] = Mp[E1B][i] - fu[i];
#endif
                    zz += z[i] * z[i];
                }
                /* W.B = |Qh.R.B[:,1]|^2 */
                double Wb = 0.0;
#pragma unroll
                for (int m = E1B; m < n; m++) {
#if PN_MASS
                    const double l1 = P.L[m * n + 1] * PIv[1] * pn_mass(0, 0);
#else
                    const double l1 = P.L[m * n + E1B] * PIv[E1B];
#endif
                    Wb += l1 * l1;
                }
                double EEst = 0.0, qq = 1.0, q11 = 0.0, lEE = 0.0;
                if (ADAPTIVE || DYNAMIC) {
                    ldiff = zz / Wb / (double)d;
                    if (P.forced_diffusion && naccept < P.n_forced) ldiff = P.forced_diffusion[naccept];
                }
                if (ADAPTIVE) {
                    const double e2c = ldiff * Wb * (h * h);
                    double e2 = 0.0;
#pragma unroll
                    for (int i = 0; i < d; i++) {
                        const double isc = pn_rcp(P.abstol + fmax(fabs(Mp[SB][i]), fabs(uprev[i])) * P.reltol);
                        e2 += isc * isc;
                    }
                    e2 = e2 * e2c * (1.0 / (double)d);
                    EEst = (e2 > 0.0) ? e2 * pn_rsqrt(e2) : e2;
                    if (EEst != EEst) {
                        retcode = PN_RET_UNSTABLE;
                        break;
                    }
                    if (EEst == 0.0) {
                        qq = 1.0 / P.qmax;
                    } else {
                        lEE = log(EEst);
                        q11 = exp(P.beta1 * lEE);
                        qq = q11 * pn_rcp(qold_pb2 * P.gamma);
                        qq = fmax(1.0 / P.qmax, fmin(1.0 / P.qmin, qq));
                    }
                    if (!(EEst < 1.0)) { /* reject */
                        nreject++;
                        dt = dt / fmin(1.0 / P.qmin, q11 / P.gamma);
                        continue;
                    }
                }
                double ediff = DYNAMIC ? ldiff : P.initial_diffusion;
                if (P.forced_diffusion && naccept < P.n_forced) ediff = P.forced_diffusion[naccept];
                /* predict_cov! on the factor */
                double Bt[n][n];
                {
                    const double sd = (ediff == 1.0) ? 1.0 : sqrt(ediff);
#pragma unroll
                    for (int r = 0; r < n; r++) {
#pragma unroll
                        for (int j = 0; j < n; j++) {
                            double s = RB[r][j];
#pragma unroll
                            for (int k = 0; k < n; k++)
                                if (k > j) s += hp[k > j ? k - j : 0] * RB[r][k];
                            RB[r][j] = s; /* ascending j: in place is safe */
                        }
#pragma unroll
                        for (int c = 0; c < n; c++) Bt[r][c] = (c <= r) ? P.L[r * n + c] * PIv[c] * sd : 0.0;
                    }
                }
                qr_stack(RB, Bt);
                /* update: H.B = e_1' */
                double c1[n];
                double S = 0.0;
#pragma unroll
                for (int r = 0; r < n; r++) {
#if PN_MASS
                    c1[r] = RB[r][1] * pn_mass(0, 0);
#else
                    c1[r] = RB[r][E1B];
#endif
                    S += c1[r] * c1[r];
                }
                if (S == 0.0) { /* Sigma^- == 0: update.jl:82-89 */
                    loglik += (zz == 0.0) ? PN_INF : -PN_INF;
#pragma unroll
                    for (int j = 0; j < n; j++)
#pragma unroll
                        for (int i = 0; i < d; i++) M[j][i] = Mp[j][i];
                } else {
#if PN_OBSN
                    S += pn_obsn_cov(0, 0); /* isotropic pn_observation_noise: the 1 x 1 left Kronecker factor of R (perform_step.jl:185-187) */
#endif
                    const double iS = pn_rcp(S);
                    loglik += -0.5 * zz * iS - 0.5 * (double)d * 1.8378770664093453;
                    if (!pn_logacc_mul(lgm, lge, S)) loglik -= 0.5 * log(S);
                    if (!DYNAMIC) {
                        const double inc = zz * iS * (1.0 / (double)d);
                        gdiff = (naccept == 0) ? inc : gdiff + (inc - gdiff) / (double)naccept;
                    }
                    double K[n];
#pragma unroll
                    for (int c = 0; c < n; c++) {
                        double s = 0.0;
#pragma unroll
                        for (int r = 0; r < n; r++) s += RB[r][c] * c1[r];
                        K[c] = s * iS;
                    }
#pragma unroll
                    for (int j = 0; j < n; j++)
#pragma unroll
                        for (int i = 0; i < d; i++) M[j][i] = Mp[j][i] - K[j] * z[i];
#pragma unroll
                    for (int r = 0; r < n; r++)
#pragma unroll
                        for (int c = 0; c < n; c++) RB[r][c] -= c1[r] * K[c];
#if PN_OBSN
                    { /* update.jl:125-136 on the factor: Sigma = R+'R+ + K r K', here as qr([R+ ; sqrt(r) K']) */
                        double Bn[n][n];
                        const double sr = sqrt(pn_obsn_cov(0, 0));
#pragma unroll
                        for (int r = 0; r < n; r++)
#pragma unroll
                            for (int c = 0; c < n; c++) Bn[r][c] = (r == 0) ? sr * K[c] : 0.0;
                        qr_stack(RB, Bn);
                    }
#endif
                }
                /* loopfooter! accept */
                double dtnew = dt;
                if (ADAPTIVE) {
                    if (P.qsteady_min <= qq && qq <= P.qsteady_max) qq = 1.0;
                    qold = fmax(EEst, P.qoldinit);
                    qold_pb2 = (EEst > P.qoldinit) ? exp(P.beta2 * lEE) : P.qoldinit_pb2;
                    dtnew = dt / qq;
                }
                const double ttmp = t + h;
                const double ref = fmax(fabs(t), fabs(tstop));
                const double epsref = nextafter(ref, 1e300) - ref;
                t = (fabs(ttmp - tstop) < 100.0 * epsref) ? tstop : ttmp;
                dt = dtnew;
                naccept++;
#pragma unroll
                for (int i = 0; i < d; i++) {
                    uprev[i] = M[SB][i];
                    if (M[0][i] != M[0][i]) retcode = PN_RET_UNSTABLE;
                }
                if (SAVE >= 1) {
                    if (P.step_offsets && save_pos < P.step_offsets[traj + 1]) {
                        P.s_t[save_pos] = t;
                        P.s_diffusion[save_pos - 1] = (SAVE >= 2) ? ediff : ((ADAPTIVE || DYNAMIC) ? ldiff : P.initial_diffusion);
                        write_pu(P.s_u_mean + save_pos * DU, P.s_u_cov + save_pos * DU * DU, M, RB, 1.0);
                        if (SAVE >= 2) save_state(P, save_pos, M, RB, h, 1.0);
                    }
                    save_pos += 1;
                }
            }
            /* postamble */
            double sc = 1.0, fd = ldiff;
            if (!DYNAMIC) {
                if (P.calibrate && naccept > 0) {
                    sc = gdiff;
                    fd = gdiff * P.initial_diffusion;
                } else {
                    fd = P.initial_diffusion;
                }
            }
            P.t_end[traj] = t;
            write_pu(P.u_mean + traj * DU, P.u_cov + traj * DU * DU, M, RB, sc);
            P.loglik[traj] = loglik - 0.5 * pn_logacc_value(lgm, lge);
            P.diffusion[traj] = fd;
            P.dt_last[traj] = dt;
            P.naccept[traj] = naccept;
            P.nreject[traj] = nreject;
            P.nf[traj] = nf;
            P.retcode[traj] = retcode;
            if (P.x_mean) {
#pragma unroll
                for (int j = 0; j < n; j++)
#pragma unroll
                    for (int i = 0; i < d; i++) P.x_mean[traj * D + j * d + i] = M[j][i];
            }
            if (P.x_covfac) { /* dense form: kron(R_B, I_d), row-major D x D */
                const double ss = sqrt(sc);
                for (int e = 0; e < D * D; e++) P.x_covfac[traj * D * D + e] = 0.0;
#pragma unroll
                for (int r = 0; r < n; r++)
#pragma unroll
                    for (int c = 0; c < n; c++)
#pragma unroll
                        for (int i = 0; i < d; i++) P.x_covfac[(traj * D + r * d + i) * D + c * d + i] = RB[r][c] * ss;
            }
        }
    }

    static PN_HD void save_state(const PnParams& P, long long pos, const double (&M)[n][d],
                                                   const double (&RB)[n][n], double h, double hvalid) {
        if (hvalid != 0.0) P.s_h[pos - 1] = h;
        for (int j = 0; j < n; j++)
            for (int i = 0; i < d; i++) P.s_x_mean[pos * D + j * d + i] = M[j][i];
        for (int e = 0; e < D * D; e++) P.s_x_covfac[pos * D * D + e] = 0.0;
        for (int r = 0; r < n; r++)
            for (int c = 0; c < n; c++)
                for (int i = 0; i < d; i++) P.s_x_covfac[(pos * D + r * d + i) * D + c * d + i] = RB[r][c];
    }
};

template <class Prob, int Q, bool ADAPTIVE, bool DYNAMIC, int SAVE>
__device__ __forceinline__ void pn_kron_entry(const PnParams& P) {
    PnKron<Prob, Q, ADAPTIVE, DYNAMIC, SAVE>::run(P);
}

#endif /* PN_KRON_CUH */
