// pn_blocks.cuh -- the BlocksOfDiagonals layout (src/blocksofdiagonals.jl:15-52): the D x D covariance factor is d
// independent (q+1) x (q+1) blocks, block i <-> state indices i:d:D.  The reference selects it for the EK0 with a
// multivariate diffusion and for the DiagonalEK1 (src/algorithms.jl:132-151, 343-393).
//
// One THREAD per trajectory, like pn_kron.cuh: the d factors, the 2n x n QR stack of the block being processed and
// the n x d mean live in the thread's registers; the warp runs 32 independent adaptive time loops.
//
//   reference routine (file:line)                                       here (per thread, per block i)
//   calc_H! DiagonalEK1      src/derivative_utils.jl:14-28               H_i = e_1' - J_ii e_0'   (EK0: H_i = e_1')
//   local_scalar_diffusion   src/diffusions/calibration.jl:21-29,123-133 s2 = sum_i z_i^2 / HQH_i / d
//   local_diagonal_diffusion src/diffusions/calibration.jl:151-180       s2_i = z_i^2 / HQH_i      (multivariate)
//   estimate_errors! Blocks  src/perform_step.jl:240-259                 e_i = h sqrt(s2_i HQH_i)
//   predict_cov! Blocks      src/filtering/predict.jl:120-141            QR of [R_i T' ; sqrt(s2_i) L P^-1]
//   update! Blocks           src/filtering/update.jl:197-239             scalar S_i, K_i = R_i^-' c_i / S_i, ll = sum_i ll_i
//   estimate_global_diffusion  src/diffusions/calibration.jl:49-66 (Fixed: invquad over blocks / d),
//                              :88-107 (FixedMV: z_j^2 / S[1,1] -- the first block's S for every j, as written there)
//   calibrate_solution!      src/integrator_utils.jl:46-65 + apply_diffusion! Blocks (apply_diffusion.jl:43-50)
//
// NVRTC-compatible: no standard headers.
#ifndef PN_BLOCKS_CUH
#define PN_BLOCKS_CUH
#include "pn_params.h"
#include "pn_taylor.cuh"
#include "pn_small.cuh" /* pn_rcp, pn_rsqrt, pn_initial_dt */

// thread-local Householder QR of [Rt ; Bt] (2n x n), Bt lower triangular; Rt returns upper triangular
template <int n>
PN_HD void pn_qr_local(double (&Rt)[n][n], double (&Bt)[n][n]) {
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

template <class Prob, int Q, bool DIAG1, bool MV, bool ADAPTIVE, bool DYNAMIC, int SAVE>
struct PnBlocks {
    static constexpr int DP = Prob::d;                  /* dimension of the right-hand side */
    static constexpr int d = PN_SECOND ? DP / 2 : DP;   /* second-order problems (PN_SECOND, pn_small.cuh): positions */
    static constexpr int DU = DP;                       /* length of sol.u = (du, u) */
    static constexpr int E1B = PN_SECOND ? 2 : 1;       /* H_i = e_2' - Jdu_ii e_1' - Ju_ii e_0' (derivative_utils.jl:14-28) / e_1' - J_ii e_0' */
    static constexpr int SB = PN_SECOND ? 1 : 0;        /* block the error estimate is scaled with (du / u) */
    static constexpr int n = Q + 1;
    static constexpr int D = d * n;
    static constexpr int NPA = Prob::n_p > 0 ? Prob::n_p : 1;

    /* sol.u = SolProj x and its covariance, component-wise scaled by sc[a]: E0 rows, or (E1; E0) rows for second-order
       problems (component a couples its du and u entries only) */
    static __device__ __forceinline__ void write_pu(double* um, double* uc, const double (&M)[n][d], const double (&RB)[d][n][n],
                                                    const double* sc) {
#pragma unroll
        for (int i = 0; i < DU * DU; i++) uc[i] = 0.0;
#pragma unroll
        for (int a = 0; a < d; a++) {
            double c00 = 0.0, c01 = 0.0, c11 = 0.0;
#pragma unroll
            for (int r = 0; r < n; r++) {
                c00 += RB[a][r][0] * RB[a][r][0];
                if (PN_SECOND) {
                    c01 += RB[a][r][0] * RB[a][r][1];
                    c11 += RB[a][r][1] * RB[a][r][1];
                }
            }
            const double s = sc ? sc[a] : 1.0;
            if (PN_SECOND) {
                um[a] = M[1][a];
                um[d + a] = M[0][a];
                uc[a * DU + a] = c11 * s;
                uc[a * DU + d + a] = c01 * s;
                uc[(d + a) * DU + a] = c01 * s;
                uc[(d + a) * DU + d + a] = c00 * s;
            } else {
                um[a] = M[0][a];
                uc[a * DU + a] = c00 * s;
            }
        }
    }

    static __device__ void run(const PnParams& P) {
        double M[n][d]; /* mean as n x d matrix: M[j][i] = mu[j*d + i] */
        double RB[d][n][n];
        double pr[NPA];
        for (;;) {
            const unsigned long long idx = atomicAdd(P.work_counter, 1ULL);
            if ((long long)idx * P.traj_stride >= P.n_traj) break;
            const long long traj = (long long)idx * P.traj_stride;
            double t = P.t0, dt, loglik = 0.0, qold = P.qoldinit, qold_pb2 = P.qoldinit_pb2;
            double gdiff[d], ldiff[d];
#pragma unroll
            for (int i = 0; i < d; i++) gdiff[i] = ldiff[i] = 0.0;
            long long naccept = 0, nreject = 0, nf = Q, iter = 0, tstop_idx = 0, save_pos = 0;
            int retcode = PN_RET_SUCCESS;
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
            for (int i = 0; i < d; i++)
#pragma unroll
                for (int a = 0; a < n; a++)
#pragma unroll
                    for (int b = 0; b < n; b++) RB[i][a][b] = 0.0;
            if (SAVE >= 1) {
                save_pos = P.step_offsets ? P.step_offsets[traj] : 0;
                if (P.step_offsets) {
                    P.s_t[save_pos] = t;
                    P.s_diffusion[save_pos] = 0.0;
#pragma unroll
                    for (int i = 0; i < d; i++)
                        if (P.s_diffusion_mv) P.s_diffusion_mv[save_pos * d + i] = 0.0;
                    write_pu(P.s_u_mean + save_pos * DU, P.s_u_cov + save_pos * DU * DU, M, RB, nullptr);
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
                double fu[DP], z[d], h0[d], h1[d], x2[DP];
#pragma unroll
                for (int i = 0; i < d; i++) {
                    x2[i] = Mp[1][i]; /* (du, u) */
                    x2[d + i] = Mp[0][i];
                    h1[i] = 0.0;
                }
                Prob::template f<double>(fu, x2, pr, tnew);
#else
                double fu[d], z[d], h0[d];
                Prob::template f<double>(fu, Mp[0], pr, tnew);
#endif
                nf += 1;
#pragma unroll
                for (int i = 0; i < d; i++) {
#if PN_MASS
                    z[i] = pn_mass(i, i) * Mp[1][i] - fu[i]; /* diagonal M: H_i = M_ii e_1' - J_ii e_0' */
#else
                    z[i] = Mp[E1B][i] - fu[i];
#endif
                    h0[i] = 0.0;
                }
                if (DIAG1) { /* H_i = e_1' - J_ii e_0' */
#if PN_SECOND
                    double J2[DP * DP]; /* ddu_diag = [Jdu_ii Ju_ii] times the (E1; E0) rows of block i (derivative_utils.jl:20-27) */
                    Prob::jac(J2, x2, pr, tnew);
#pragma unroll
                    for (int i = 0; i < d; i++) {
                        h1[i] = -J2[i + DP * i];
                        h0[i] = -J2[i + DP * (d + i)];
                    }
#else
                    double J[d * d];
                    Prob::jac(J, Mp[0], pr, tnew);
#pragma unroll
                    for (int i = 0; i < d; i++) h0[i] = -J[i + d * i];
#endif
                }
                /* HQH_i = |Qh.R H_i'|^2, (Qh.R H_i')[m] = L[m][0] PI[0] h0_i + L[m][1] PI[1] */
                double HQH[d];
#pragma unroll
                for (int i = 0; i < d; i++) {
                    double e = 0.0;
#pragma unroll
                    for (int m = 0; m < n; m++) {
#if PN_MASS
                        const double cq = P.L[m * n + 0] * PIv[0] * h0[i] + ((m >= 1) ? P.L[m * n + 1] * PIv[1] * pn_mass(i, i) : 0.0);
#elif PN_SECOND
                        const double cq = P.L[m * n + 0] * PIv[0] * h0[i] + ((m >= 1) ? P.L[m * n + 1] * PIv[1] * h1[i] : 0.0) +
                                          ((m >= 2) ? P.L[m * n + 2] * PIv[2] : 0.0);
#else
                        const double cq = P.L[m * n + 0] * PIv[0] * h0[i] + ((m >= 1) ? P.L[m * n + 1] * PIv[1] : 0.0);
#endif
                        e += cq * cq;
                    }
                    HQH[i] = e;
                }
                double EEst = 0.0, qq = 1.0, q11 = 0.0, lEE = 0.0;
                if (ADAPTIVE || DYNAMIC) {
                    double qd = 0.0;
#pragma unroll
                    for (int i = 0; i < d; i++) {
                        const double r = z[i] * z[i] / HQH[i];
                        ldiff[i] = r;
                        qd += r;
                    }
                    if (!MV) {
                        qd = qd / (double)d;
#pragma unroll
                        for (int i = 0; i < d; i++) ldiff[i] = qd;
                    }
                }
                if (ADAPTIVE) {
                    double e2 = 0.0;
#pragma unroll
                    for (int i = 0; i < d; i++) {
                        const double isc = pn_rcp(P.abstol + fmax(fabs(Mp[SB][i]), fabs(uprev[i])) * P.reltol);
                        e2 += (ldiff[i] * HQH[i]) * (isc * isc);
                    }
                    e2 = e2 * (h * h) * (1.0 / (double)d);
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
                double ediff[d], S[d];
                double llstep = 0.0;
#pragma unroll
                for (int i = 0; i < d; i++) {
                    ediff[i] = DYNAMIC ? ldiff[i] : P.initial_diffusion;
                    /* predict_cov! on block i */
                    double Bt[n][n];
                    const double sd = (ediff[i] == 1.0) ? 1.0 : sqrt(ediff[i]);
#pragma unroll
                    for (int r = 0; r < n; r++) {
#pragma unroll
                        for (int j = 0; j < n; j++) {
                            double s = RB[i][r][j];
#pragma unroll
                            for (int k = 0; k < n; k++)
                                if (k > j) s += hp[k > j ? k - j : 0] * RB[i][r][k];
                            RB[i][r][j] = s;
                        }
#pragma unroll
                        for (int c = 0; c < n; c++) Bt[r][c] = (c <= r) ? P.L[r * n + c] * PIv[c] * sd : 0.0;
                    }
                    pn_qr_local<n>(RB[i], Bt);
                    /* update! on block i: c = R^- H_i' */
                    double c1[n];
                    double Si = 0.0;
#pragma unroll
                    for (int r = 0; r < n; r++) {
#if PN_MASS
                        c1[r] = RB[i][r][1] * pn_mass(i, i) + h0[i] * RB[i][r][0];
#elif PN_SECOND
                        c1[r] = RB[i][r][2] + h1[i] * RB[i][r][1] + h0[i] * RB[i][r][0];
#else
                        c1[r] = RB[i][r][1] + h0[i] * RB[i][r][0];
#endif
                        Si += c1[r] * c1[r];
                    }
                    S[i] = Si;
                    if (Si == 0.0) { /* Sigma^- == 0: update.jl:82-89 */
                        llstep += (z[i] == 0.0) ? PN_INF : -PN_INF;
#pragma unroll
                        for (int j = 0; j < n; j++) M[j][i] = Mp[j][i];
                    } else {
#if PN_OBSN
                        Si += pn_obsn_cov(i, i); /* diagonal pn_observation_noise: block i's 1 x 1 factor (perform_step.jl:185-187) */
                        S[i] = Si;
#endif
                        const double iS = pn_rcp(Si);
                        llstep += -0.5 * z[i] * z[i] * iS - 0.5 * 1.8378770664093453 - 0.5 * log(Si);
                        double K[n];
#pragma unroll
                        for (int c = 0; c < n; c++) {
                            double s = 0.0;
#pragma unroll
                            for (int r = 0; r < n; r++) s += RB[i][r][c] * c1[r];
                            K[c] = s * iS;
                        }
#pragma unroll
                        for (int j = 0; j < n; j++) M[j][i] = Mp[j][i] - K[j] * z[i];
#pragma unroll
                        for (int r = 0; r < n; r++)
#pragma unroll
                            for (int c = 0; c < n; c++) RB[i][r][c] -= c1[r] * K[c];
#if PN_OBSN
                        { /* update.jl:125-136 on block i: qr([R+ ; sqrt(r_i) K']) */
                            double Bn[n][n];
                            const double sr = sqrt(pn_obsn_cov(i, i));
#pragma unroll
                            for (int r = 0; r < n; r++)
#pragma unroll
                                for (int c = 0; c < n; c++) Bn[r][c] = (r == 0) ? sr * K[c] : 0.0;
                            pn_qr_local<n>(RB[i], Bn);
                        }
#endif
                    }
                }
                loglik += llstep;
                if (!DYNAMIC) {
                    if (MV) { /* calibration.jl:88-107: z_j^2 / S[1,1] */
                        const double iS0 = 1.0 / S[0];
#pragma unroll
                        for (int j = 0; j < d; j++) {
                            const double inc = z[j] * z[j] * iS0;
                            gdiff[j] = (naccept == 0) ? inc : gdiff[j] + (inc - gdiff[j]) / (double)naccept;
                        }
                    } else {
                        double qd = 0.0;
#pragma unroll
                        for (int i = 0; i < d; i++) qd += z[i] * (z[i] / S[i]);
                        const double inc = qd / (double)d;
                        const double g = (naccept == 0) ? inc : gdiff[0] + (inc - gdiff[0]) / (double)naccept;
#pragma unroll
                        for (int j = 0; j < d; j++) gdiff[j] = g;
                    }
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
                        P.s_diffusion[save_pos - 1] = (SAVE >= 2) ? ediff[0] : ((ADAPTIVE || DYNAMIC) ? ldiff[0] : P.initial_diffusion);
#pragma unroll
                        for (int i = 0; i < d; i++) {
                            if (P.s_diffusion_mv)
                                P.s_diffusion_mv[(save_pos - 1) * d + i] =
                                    (SAVE >= 2) ? ediff[i] : ((ADAPTIVE || DYNAMIC) ? ldiff[i] : P.initial_diffusion);
                        }
                        write_pu(P.s_u_mean + save_pos * DU, P.s_u_cov + save_pos * DU * DU, M, RB, nullptr);
                        if (SAVE >= 2) save_state(P, save_pos, M, RB, h, 1.0);
                    }
                    save_pos += 1;
                }
            }
            /* postamble */
            P.t_end[traj] = t;
            double scs[d];
#pragma unroll
            for (int a = 0; a < d; a++) {
                double sc = 1.0, fd = ldiff[a];
                if (!DYNAMIC) {
                    if (P.calibrate && naccept > 0) {
                        sc = gdiff[a];
                        fd = gdiff[a] * P.initial_diffusion;
                    } else {
                        fd = P.initial_diffusion;
                    }
                }
                scs[a] = sc;
                if (a == 0) P.diffusion[traj] = fd;
                if (P.diffusion_mv) P.diffusion_mv[traj * d + a] = fd;
                if (P.x_covfac) {
                    const double ss = sqrt(sc);
                    if (a == 0)
                        for (int e = 0; e < D * D; e++) P.x_covfac[traj * D * D + e] = 0.0;
#pragma unroll
                    for (int r = 0; r < n; r++)
#pragma unroll
                        for (int c = 0; c < n; c++) P.x_covfac[(traj * D + r * d + a) * D + c * d + a] = RB[a][r][c] * ss;
                }
            }
            write_pu(P.u_mean + traj * DU, P.u_cov + traj * DU * DU, M, RB, scs);
            P.loglik[traj] = loglik;
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
        }
    }

    /* dense form of the Blocks state for the smoother sweep: block i occupies rows/columns i:d:D */
    static PN_HD void save_state(const PnParams& P, long long pos, const double (&M)[n][d], const double (&RB)[d][n][n],
                                 double h, double hvalid) {
        if (hvalid != 0.0) P.s_h[pos - 1] = h;
        for (int j = 0; j < n; j++)
            for (int i = 0; i < d; i++) P.s_x_mean[pos * D + j * d + i] = M[j][i];
        for (int e = 0; e < D * D; e++) P.s_x_covfac[pos * D * D + e] = 0.0;
        for (int i = 0; i < d; i++)
            for (int r = 0; r < n; r++)
                for (int c = 0; c < n; c++) P.s_x_covfac[(pos * D + r * d + i) * D + c * d + i] = RB[i][r][c];
    }
};

template <class Prob, int Q, bool DIAG1, bool MV, bool ADAPTIVE, bool DYNAMIC, int SAVE>
__device__ __forceinline__ void pn_blocks_entry(const PnParams& P) {
    PnBlocks<Prob, Q, DIAG1, MV, ADAPTIVE, DYNAMIC, SAVE>::run(P);
}

#endif /* PN_BLOCKS_CUH */
