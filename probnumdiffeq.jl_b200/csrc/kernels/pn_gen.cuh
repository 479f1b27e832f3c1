// pn_gen.cuh -- filter for priors whose transition is a general dense D x D pair (Ah, Qh): the integrated Ornstein-Uhlenbeck
// process with a VECTOR or MATRIX rate parameter, and with the Rosenbrock-style rate update (rate := Jacobian at the current
// state, every step).  One warp per trajectory; every matrix is dense, row-major, in a per-warp workspace (shared memory when it
// fits, else an L2-resident global buffer), and the arithmetic follows the reference formulas literally -- this is the
// functionality path for ExpEK(L = matrix) / RosenbrockExpEK, not the headline kernel (scalar rates and the IWP stay on the
// register-resident kernels of pn_small.cuh).
//
// Reference behaviour replaced (per trajectory):
//   perform_step!                          src/perform_step.jl:79-154 (rate update :88-96, make_new_transitions :37-50)
//   to_sde / update_sde_drift!             src/priors/ioup.jl:81-113   F = shift (x) I_d with F[q-block, q-block] = rate matrix
//   make_transition_matrices!(::IOUP)      src/priors/ioup.jl:115-131  (Ah, Qh.R) = exp_and_gram_chol(F, L, dt), preconditioned
//   predict_mean! / predict_cov!           src/filtering/predict.jl:53-94 (Gram + Cholesky, triangularize! on failure)
//   calc_H! / measurement                  src/derivative_utils.jl:1-33, src/measurement_models.jl:11-20
//   local diffusion / error estimate       src/diffusions/calibration.jl:123-133, src/perform_step.jl:199-247
//   update!                                src/filtering/update.jl:65-139
//   global diffusion MLE                   src/diffusions/calibration.jl:49-66
//
// Transition matrices: pn_lti.cuh (scaling and doubling in preconditioned coordinates).
//
// NVRTC-compatible: no standard headers.
#ifndef PN_GEN_CUH
#define PN_GEN_CUH
#include "pn_params.h"
#include "pn_small.cuh" /* pn_initial_dt is not needed here (pre-pass), but PN_EK0_DENSE, pn_ulp_above */

#include "pn_lti.cuh"

template <class Prob, int Q, bool ADAPTIVE, bool DYNAMIC, int SAVE>
__device__ void pn_gen_entry(const PnParams& P) {
    extern __shared__ double pn_gen_shared[];
    constexpr int d = Prob::d, n = Q + 1, D = d * n, DD = D * D, NPA = Prob::n_p > 0 ? Prob::n_p : 1;
    constexpr int per_warp = pn_gen_ws_doubles(d, n);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* W = P.gen_workspace ? P.gen_workspace + ((size_t)blockIdx.x * PN_GEN_WARPS + warp) * per_warp : pn_gen_shared + (size_t)warp * per_warp;
    double* Ah = W;             /* D x D */
    double* Qc = Ah + DD;       /* D x D  Qh.R (upper) */
    double* R = Qc + DD;        /* D x D  filtered right factor */
    double* Xs = R + DD;        /* 2D x D stack [R Ah'; sqrt(s2) Qh.R] */
    double* Mg = Xs + 2 * DD;   /* D x D  Gram matrix -> predicted factor */
    double* T1 = Mg + DD;       /* D x D scratch (x3: At, Qt, T2 below) */
    double* T2 = T1 + DD;
    double* At = T2 + DD;
    double* Qt = At + DD;
    double* mu = Qt + DD;       /* D */
    double* mup = mu + D;       /* D */
    double* Kz = mup + D;       /* D scratch */
    double* sp0 = Kz + D;       /* 3 D spare */
    double* Hm = sp0 + 3 * D;   /* d x D  H */
    double* C2 = Hm + d * D;    /* D x d  R^- H' (also Qh.R H') */
    double* Kt = C2 + D * d;    /* D x d  K */
    double* Sm = Kt + D * d;    /* d x d */
    double* Wm = Sm + d * d;    /* d x d */
    double* Jm = Wm + d * d;    /* d x d  column-major Jacobian */
    double* Rm = Jm + d * d;    /* d x d  row-major rate matrix of this step */
    double* Zm = Rm + d * d;    /* d x d */
    double* Sp = Zm + d * d;    /* d x d spare */
    double* z = Sp + d * d;     /* d */
    double* zt = z + d;         /* d */
    double* fu = zt + d;        /* d */
    double* uprev = fu + d;     /* d */
    double* wdiag = uprev + d;  /* d */
    double* u0x = wdiag + d;    /* 3 d spare */
    double* PIv = u0x + 3 * d;  /* n */
    (void)sp0; (void)Sp; (void)u0x;

    double pr[NPA];
    for (;;) {
        unsigned long long idx = 0;
        if (lane == 0) idx = atomicAdd(P.work_counter, 1ULL);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if ((long long)idx * P.traj_stride >= P.n_traj) break;
        const long long traj = (long long)idx * P.traj_stride;
        for (int c = lane; c < D; c += 32) mu[c] = P.init_mu[traj * D + c];
        for (int i = lane; i < DD; i += 32) R[i] = 0.0;
#pragma unroll
        for (int i = 0; i < NPA; i++) pr[i] = (i < Prob::n_p) ? P.p[traj * Prob::n_p + i] : 0.0;
        __syncwarp();
        for (int i = lane; i < d; i += 32) uprev[i] = mu[i];
        for (int i = lane; i < d * d; i += 32) Rm[i] = P.rate_matrix ? P.rate_matrix[i] : ((i / d == i % d) ? P.rate : 0.0);
        __syncwarp();
        double t = P.t0, dt = P.init_dt[traj];
        const double dtcache = dt;
        double qold_pb2 = P.qoldinit_pb2, loglik = 0.0, gdiff = 0.0, ldiff = 0.0;
        long long naccept = 0, nreject = 0, iter = 0, tstop_idx = 0;
        long long nf = Q + ((ADAPTIVE && !(P.dt0 > 0.0)) ? 2 : 0);
        int retcode = PN_RET_SUCCESS;
        long long save_pos = 0;
        if (SAVE >= 1) {
            save_pos = P.step_offsets ? P.step_offsets[traj] : 0;
            if (P.step_offsets) {
                if (lane == 0) {
                    P.s_t[save_pos] = t;
                    P.s_diffusion[save_pos] = 0.0;
                }
                for (int i = lane; i < d; i += 32) P.s_u_mean[save_pos * d + i] = mu[i];
                for (int i = lane; i < d * d; i += 32) P.s_u_cov[save_pos * d * d + i] = 0.0;
                if (SAVE >= 2) {
                    for (int c = lane; c < D; c += 32) P.s_x_mean[save_pos * D + c] = mu[c];
                    for (int i = lane; i < DD; i += 32) P.s_x_covfac[save_pos * DD + i] = 0.0;
                }
            }
            save_pos += 1;
        }
        bool rate_fresh = false; /* the Jacobian-updated rate is a function of (uprev, t): one evaluation per accepted step */

        while (t < P.t1) {
            double tstop = P.t1;
            while (tstop_idx < P.n_tstops && P.tstops[tstop_idx] <= t) tstop_idx++;
            if (tstop_idx < P.n_tstops && P.tstops[tstop_idx] < P.t1) tstop = P.tstops[tstop_idx];
            iter++;
            if (iter > P.maxiters) { retcode = PN_RET_MAXITERS; break; }
            if (ADAPTIVE) {
                dt = fmin(dt, P.dtmax);
                dt = fmax(dt, P.dtmin);
                dt = fmin(dt, tstop - t);
            } else {
                dt = fmin(dtcache, tstop - t);
            }
            const double h = dt;
            if (h != h || !(t + h > t) || (ADAPTIVE && h <= P.dtmin)) { retcode = PN_RET_DTLESSTHANMIN; break; }
            const double tnew = t + h;

            /* Rosenbrock-style update of the rate parameter (perform_step.jl:88-96): calc_J! at (uprev, t) */
            if (P.update_rate && !rate_fresh) {
                if (lane == 0) {
                    for (int i = 0; i < d * d; i++) Jm[i] = 0.0;
                    Prob::jac(Jm, uprev, pr, t);
                    for (int a = 0; a < d; a++)
                        for (int i = 0; i < d; i++) Rm[a * d + i] = Jm[a + d * i];
                }
                rate_fresh = true;
                __syncwarp();
            }
            if (!pn_gen_transition(Ah, Qc, Rm, h, d, n, At, Qt, T1, T2, Zm, PIv, lane)) { retcode = PN_RET_UNSTABLE; break; }

            /* predict_mean! */
            for (int r = lane; r < D; r += 32) {
                double s = 0.0;
                for (int c = 0; c < D; c++) s += Ah[r * D + c] * mu[c];
                mup[r] = s;
            }
            __syncwarp();
            /* measurement z = E1 mu^- - f(E0 mu^-), H = E1 - J E0 (EK1) or E1 (EK0) */
            if (lane == 0) {
                Prob::template f<double>(fu, mup, pr, tnew);
                for (int i = 0; i < d * d; i++) Jm[i] = 0.0;
#if !PN_EK0_DENSE
                Prob::jac(Jm, mup, pr, tnew);
#endif
            }
            nf += 1;
            __syncwarp();
            for (int i = lane; i < d; i += 32) z[i] = mup[d + i] - fu[i];
            for (int i = lane; i < d * D; i += 32) {
                const int a = i / D, c = i % D;
                double v = 0.0;
                if (c < d) v = -Jm[a + d * c];
                else if (c < 2 * d) v = (c - d == a) ? 1.0 : 0.0;
                Hm[i] = v;
            }
            __syncwarp();
            double EEst = 0.0, qq = 1.0, q11 = 0.0, lEE = 0.0;
            if (ADAPTIVE || DYNAMIC) {
                /* C = Qh.R H' (D x d), W = C'C, sigma^2 = z' W^-1 z / d */
                for (int i = lane; i < D * d; i += 32) {
                    const int r = i / d, a = i % d;
                    double s = 0.0;
                    for (int c = r; c < D; c++) s += Qc[r * D + c] * Hm[a * D + c];
                    C2[i] = s;
                }
                __syncwarp();
                for (int i = lane; i < d * d; i += 32) {
                    const int a = i / d, b = i % d;
                    double s = 0.0;
                    for (int r = 0; r < D; r++) s += C2[r * d + a] * C2[r * d + b];
                    Wm[i] = s;
                    if (a == b) wdiag[a] = s;
                }
                __syncwarp();
                const bool okw = pn_gen_chol(Wm, d, lane);
                if (!okw) { retcode = PN_RET_UNSTABLE; break; }
                if (lane == 0) { /* zt = W^-1 z */
                    for (int a = 0; a < d; a++) {
                        double s = z[a];
                        for (int k = 0; k < a; k++) s -= Wm[k * d + a] * zt[k];
                        zt[a] = s / Wm[a * d + a];
                    }
                    for (int a = d - 1; a >= 0; a--) {
                        double s = zt[a];
                        for (int k = a + 1; k < d; k++) s -= Wm[a * d + k] * zt[k];
                        zt[a] = s / Wm[a * d + a];
                    }
                }
                __syncwarp();
                double qd = 0.0;
                for (int a = 0; a < d; a++) qd += z[a] * zt[a];
                ldiff = qd / (double)d;
                if (P.forced_diffusion && naccept < P.n_forced) ldiff = P.forced_diffusion[naccept];
                if (ADAPTIVE) {
                    double e2 = 0.0;
                    for (int i = 0; i < d; i++) {
                        const double sc = P.abstol + fmax(fabs(mup[i]), fabs(uprev[i])) * P.reltol;
                        const double e = sqrt(ldiff * wdiag[i]) * h / sc;
                        e2 += e * e;
                    }
                    EEst = sqrt(e2 / (double)d);
                    if (EEst != EEst) { retcode = PN_RET_UNSTABLE; break; }
                }
            }
            const bool accept = !ADAPTIVE || EEst < 1.0;
            if (ADAPTIVE) { /* stepsize_controller!(::PIController) (OrdinaryDiffEqCore) */
                if (EEst == 0.0) {
                    qq = P.inv_qmax;
                } else {
                    lEE = log(EEst);
                    q11 = exp(P.beta1 * lEE);
                    qq = q11 / (qold_pb2 * P.gamma);
                    qq = fmax(P.inv_qmax, fmin(P.inv_qmin, qq));
                }
                if (!accept) {
                    nreject++;
                    dt = dt / fmin(P.inv_qmin, q11 / P.gamma);
                    continue;
                }
            }
            double ediff = DYNAMIC ? ldiff : P.initial_diffusion;
            if (P.forced_diffusion && naccept < P.n_forced) ediff = P.forced_diffusion[naccept];

            /* predict_cov! (predict.jl:62-94) */
            for (int i = lane; i < DD; i += 32) {
                const int r = i / D, c = i % D;
                double s = 0.0;
                for (int k = 0; k < D; k++) s += R[r * D + k] * Ah[c * D + k];
                Xs[i] = s;
            }
            __syncwarp();
            if (ediff == 0.0) { /* :71-74 fast_X_A_Xt!: Sigma^- = Ah Sigma Ah' (factor not triangular; nothing below needs it to be) */
                for (int i = lane; i < DD; i += 32) Mg[i] = Xs[i];
                __syncwarp();
            } else {
                const double sd = (ediff == 1.0) ? 1.0 : sqrt(ediff);
                for (int i = lane; i < DD; i += 32) Xs[DD + i] = (ediff == 1.0) ? Qc[i] : Qc[i] * sd;
                __syncwarp();
                for (int i = lane; i < DD; i += 32) { /* M = stack' stack (:84) */
                    const int a = i / D, b = i % D;
                    if (b < a) continue;
                    double s = 0.0;
                    for (int r = 0; r < 2 * D; r++) s += Xs[r * D + a] * Xs[r * D + b];
                    Mg[i] = s;
                }
                __syncwarp();
                if (!pn_gen_chol(Mg, D, lane)) { /* :90 triangularize!(stack) */
                    pn_gen_qr(Xs, 2 * D, D, lane);
                    for (int i = lane; i < DD; i += 32) Mg[i] = Xs[i];
                    __syncwarp();
                }
            }
            /* measurement covariance S = (R^- H')'(R^- H'), update! (update.jl:65-139) */
            double rr = 0.0;
            for (int i = lane; i < DD; i += 32) rr += fabs(Mg[i]);
            rr = pn_gen_wsum(rr);
            for (int i = lane; i < D * d; i += 32) {
                const int r = i / d, a = i % d;
                double s = 0.0;
                for (int c = 0; c < 2 * d && c < D; c++) s += Mg[r * D + c] * Hm[a * D + c];
                C2[i] = s;
            }
            __syncwarp();
            for (int i = lane; i < d * d; i += 32) {
                const int a = i / d, b = i % d;
                double s = 0.0;
                for (int r = 0; r < D; r++) s += C2[r * d + a] * C2[r * d + b];
                Sm[i] = s;
            }
            __syncwarp();
            if (rr == 0.0) { /* update.jl:82-89: zero predicted covariance -> copy, +-Inf */
                bool zz = true;
                for (int a = 0; a < d; a++) zz = zz && (z[a] == 0.0);
                loglik += zz ? PN_INF : -PN_INF;
                for (int c = lane; c < D; c += 32) mu[c] = mup[c];
                for (int i = lane; i < DD; i += 32) R[i] = Mg[i];
                __syncwarp();
            } else {
                const bool oks = pn_gen_chol(Sm, d, lane);
                if (!oks) retcode = PN_RET_UNSTABLE;
                /* K = Sigma^- H' S^-1 = R^-' C2 S^-1 */
                for (int i = lane; i < D * d; i += 32) {
                    const int c = i / d, a = i % d;
                    double s = 0.0;
                    for (int r = 0; r < D; r++) s += Mg[r * D + c] * C2[r * d + a];
                    Kt[i] = s;
                }
                __syncwarp();
                for (int c = lane; c <= D; c += 32) { /* rows of K (and z as row D) times (S.R' S.R)^-1 */
                    double* x = (c < D) ? Kt + c * d : zt;
                    if (c == D)
                        for (int a = 0; a < d; a++) zt[a] = z[a];
                    for (int a = 0; a < d; a++) {
                        double s = x[a];
                        for (int k = 0; k < a; k++) s -= x[k] * Sm[k * d + a];
                        x[a] = s / Sm[a * d + a];
                    }
                    for (int a = d - 1; a >= 0; a--) {
                        double s = x[a];
                        for (int k = a + 1; k < d; k++) s -= x[k] * Sm[a * d + k];
                        x[a] = s / Sm[a * d + a];
                    }
                }
                __syncwarp();
                double quad = 0.0, logdet = 0.0;
                for (int a = 0; a < d; a++) {
                    quad += z[a] * zt[a];
                    logdet += log(Sm[a * d + a]);
                }
                loglik += -0.5 * quad - 0.5 * (double)d * 1.8378770664093453 - logdet;
                if (!DYNAMIC) { /* estimate_global_diffusion(::FixedDiffusion), calibration.jl:49-66 */
                    const double inc = quad / (double)d;
                    gdiff = (naccept == 0) ? inc : gdiff + (inc - gdiff) / (double)naccept;
                }
                for (int c = lane; c < D; c += 32) {
                    double s = mup[c];
                    for (int a = 0; a < d; a++) s -= Kt[c * d + a] * z[a];
                    mu[c] = s;
                }
                /* R+ = R^- (I - K H)' = R^- - C2 K' */
                for (int i = lane; i < DD; i += 32) {
                    const int r = i / D, c = i % D;
                    double s = Mg[i];
                    for (int a = 0; a < d; a++) s -= C2[r * d + a] * Kt[c * d + a];
                    R[i] = s;
                }
                __syncwarp();
            }
            /* loopfooter! accept branch */
            double dtnew = dt;
            if (ADAPTIVE) {
                if (P.qsteady_min <= qq && qq <= P.qsteady_max) qq = 1.0;
                qold_pb2 = (EEst > P.qoldinit) ? exp(P.beta2 * lEE) : P.qoldinit_pb2;
                dtnew = dt / qq;
            }
            const double ttmp = t + h;
            const double ref = fmax(fabs(t), fabs(tstop));
            const double epsref = pn_ulp_above(ref);
            t = (fabs(ttmp - tstop) < 100.0 * epsref) ? tstop : ttmp;
            dt = dtnew;
            naccept++;
            rate_fresh = false;
            bool nanu = false;
            for (int i = 0; i < d; i++) nanu = nanu || (mu[i] != mu[i]);
            __syncwarp();
            for (int i = lane; i < d; i += 32) uprev[i] = mu[i];
            __syncwarp();
            if (nanu) retcode = PN_RET_UNSTABLE;
            if (SAVE >= 1) {
                const bool wr = P.step_offsets && save_pos < P.step_offsets[traj + 1];
                if (wr) {
                    if (lane == 0) {
                        P.s_t[save_pos] = t;
                        P.s_diffusion[save_pos - 1] = (SAVE >= 2) ? ediff : ((ADAPTIVE || DYNAMIC) ? ldiff : P.initial_diffusion);
                        if (SAVE >= 2) P.s_h[save_pos - 1] = h;
                    }
                    for (int i = lane; i < d; i += 32) P.s_u_mean[save_pos * d + i] = mu[i];
                    for (int i = lane; i < d * d; i += 32) {
                        const int a = i / d, b = i % d;
                        double s = 0.0;
                        for (int r = 0; r < D; r++) s += R[r * D + a] * R[r * D + b];
                        P.s_u_cov[save_pos * d * d + i] = s;
                    }
                    if (SAVE >= 2) {
                        for (int c = lane; c < D; c += 32) P.s_x_mean[save_pos * D + c] = mu[c];
                        for (int i = lane; i < DD; i += 32) P.s_x_covfac[save_pos * DD + i] = R[i];
                        if (P.s_rate) /* the backward kernel of this step needs the rate it was taken with */
                            for (int i = lane; i < d * d; i += 32) P.s_rate[(save_pos - 1) * d * d + i] = Rm[i];
                    }
                }
                save_pos += 1;
            }
            if (retcode != PN_RET_SUCCESS) break;
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
        __syncwarp();
        for (int i = lane; i < d * d; i += 32) {
            const int a = i / d, b = i % d;
            double s = 0.0;
            for (int r = 0; r < D; r++) s += R[r * D + a] * R[r * D + b];
            P.u_cov[traj * d * d + i] = s * sc;
        }
        for (int i = lane; i < d; i += 32) P.u_mean[traj * d + i] = mu[i];
        if (P.x_mean)
            for (int c = lane; c < D; c += 32) P.x_mean[traj * D + c] = mu[c];
        if (P.x_covfac) {
            const double ss = sqrt(sc);
            for (int i = lane; i < DD; i += 32) P.x_covfac[traj * DD + i] = R[i] * ss;
        }
        if (lane == 0) {
            P.t_end[traj] = t;
            P.loglik[traj] = loglik;
            P.diffusion[traj] = fd;
            P.dt_last[traj] = dt;
            P.naccept[traj] = naccept;
            P.nreject[traj] = nreject;
            P.nf[traj] = nf;
            P.retcode[traj] = retcode;
        }
        __syncwarp();
    }
}

#endif /* PN_GEN_CUH */
