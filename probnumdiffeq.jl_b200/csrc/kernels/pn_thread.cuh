// pn_thread.cuh -- filter for small states, ONE THREAD per trajectory (EK1, IWP prior, dense layout, filter only: the headline
// configuration BASELINE configs[4] and the filter halves of cfg1 / cfg3).
//
// Reference behaviour replaced: the same routines as pn_small.cuh -- perform_step! (src/perform_step.jl:79-154), predict_mean! /
// predict_cov! on the Gram + Cholesky route with triangularize! as the fallback (src/filtering/predict.jl:53-94), calc_H!
// (src/derivative_utils.jl:1-33), local diffusion (src/diffusions/calibration.jl:123-133), error estimate
// (src/perform_step.jl:199-247), update! (src/filtering/update.jl:65-139), global diffusion MLE (calibration.jl:49-66), the
// OrdinaryDiffEqCore controller / loop bookkeeping.
//
// Why a second mapping.  The group-per-trajectory kernel (pn_small.cuh, G = 2 lanes at D = 12) replicates the scalar phase of a
// step on both lanes, exchanges the factor rows through shared memory for the Gram matrix and through shuffles for the
// Cholesky pivots, and spills its row arrays (144 doubles per lane): 194 warp instructions per trajectory-step, half of them
// FP64 (profiles/r2p_*).  Here a warp carries 32 trajectories and no instruction is replicated or spent on exchange:
//   * the Gram matrix M = X'X + s2 Qh (upper triangle: D(D+1)/2 accumulators) lives in the thread's REGISTERS;
//   * the filtered factor lives in shared memory between steps, packed (the rows below 2d are the untouched upper-triangular
//     rows of the predicted factor: 2d D + (D-2d)(D-2d+1)/2 doubles instead of D^2), thread-interleaved (slot s of thread t at
//     s * NT + t: conflict-free); a step streams it once, row by row -- X = R Ah' is applied to the row in flight, 78 independent
//     multiply-adds per 12 loaded doubles at D = 12 -- and writes the updated factor back once;
//   * Cholesky, S, gain and the update run on the register-resident matrix with compile-time indices (zeros of the triangular /
//     H = [-J | I | 0] structure never issue);
//   * a rejected attempt costs its scalar phase only (no identity step: nothing is shared between the trajectories of a warp).
// Trajectories are pulled from the same atomic queue as in pn_small.cuh; the initial states come from the same pre-pass.
//
// NVRTC-compatible: no standard headers.
#ifndef PN_THREAD_CUH
#define PN_THREAD_CUH
#include "pn_params.h"
#include "pn_small.cuh"

#ifndef PN_THREAD_NT
#define PN_THREAD_NT 256 /* threads per CTA (compile-time: the shared-memory interleave stride) */
#endif
#ifndef PN_THREAD_MINB
#define PN_THREAD_MINB 1
#endif
#ifndef PN_THREAD_EARLY_GRAM
#define PN_THREAD_EARLY_GRAM 0 /* 1: X'X of the factor is accumulated during the attempt, next to the scalar chain mean -> f, J -> sigma^2 ->
                                  error estimate (independent work for the same warp); s2 Qh is added once the attempt is accepted -- the order
                                  of the reference's stacked product (predict.jl:75-85: factor rows first, noise rows after).  A rejected
                                  attempt then wastes its Gram phase. */
#endif

/* compile-time loop: f(PnInt<I>{}) for I = B ... E-1.  `#pragma unroll` is a hint that the compiler declines for the large bodies of this
   kernel (measured: the row loop of the Gram phase stayed rolled, with run-time structure tests and branches in the hot path) */
template <int V>
struct PnInt {
    static constexpr int value = V;
};
template <int B, int E, class F>
__device__ __forceinline__ void pn_static_for(F&& f) {
    if constexpr (B < E) {
        f(PnInt<B>{});
        pn_static_for<B + 1, E>(f);
    }
}

/* doubles of shared memory per thread (host and kernel must agree): packed factor + mean */
__host__ __device__ constexpr int pn_thread_sm_doubles(int d, int n) {
    return 2 * d * (d * n) + ((d * n) - 2 * d) * ((d * n) - 2 * d + 1) / 2 + d * n;
}

template <class Prob, int Q, bool ADAPTIVE, bool DYNAMIC>
struct PnThread {
    static constexpr int d = Prob::d, n = Q + 1, D = d * n, RU = 2 * d;
    static constexpr int NPA = Prob::n_p > 0 ? Prob::n_p : 1;
    static constexpr int NT = PN_THREAD_NT;
    static constexpr int NPK = RU * D + (D - RU) * (D - RU + 1) / 2; /* packed factor */
    static_assert(n >= 2 && RU <= D, "the measurement reads two derivative blocks");

    /* packed slot of factor entry (r, c); rows r >= RU hold columns c >= r only */
    static __device__ __forceinline__ constexpr int slot(int r, int c) {
        return r < RU ? r * D + c : RU * D + (r - RU) * (D - RU) - ((r - RU) * (r - RU - 1)) / 2 + (c - r);
    }
    static __device__ __forceinline__ constexpr bool stored(int r, int c) { return r < RU || c >= r; }
    /* X = R Ah' : entry (r, (j, i)) = sum_{k >= j} T[j][k] R[r][(k, i)] is structurally zero iff no k >= j has a stored entry */
    static __device__ __forceinline__ constexpr bool xnz(int r, int col) { return r < RU || (Q * d + (col % d)) >= r; }

    /* stream the factor: row r of X = R Ah' in flight, M += x x' (upper triangle) */
    static __device__ __forceinline__ void gram_rows(const double* my, const double (&hp)[n], double (&M)[D][D]) {
        pn_static_for<0, D>([&](auto rc) {
            constexpr int r = decltype(rc)::value;
            double x[D];
#pragma unroll
            for (int c = 0; c < D; c++) x[c] = stored(r, c) ? my[slot(r, stored(r, c) ? c : r) * NT] : 0.0;
#pragma unroll
            for (int i = 0; i < d; i++)
#pragma unroll
                for (int j = 0; j < n; j++) {
                    if (!xnz(r, j * d + i)) continue;
                    double s = stored(r, j * d + i) ? x[j * d + i] : 0.0;
#pragma unroll
                    for (int k = 0; k < n; k++)
                        if (k > j && stored(r, k * d + i)) s += hp[k > j ? k - j : 0] * x[k * d + i];
                    x[j * d + i] = s;
                }
#pragma unroll
            for (int a = 0; a < D; a++)
#pragma unroll
                for (int b = 0; b < D; b++)
                    if (b >= a && xnz(r, a) && xnz(r, b)) M[a][b] = fma(x[a], x[b], M[a][b]);
        });
    }

    static __device__ void run(const PnParams& P, double* smem) {
        double* const my = smem + threadIdx.x;                  /* slot s at my[s * NT] */
        double* const mus = my + (size_t)NPK * NT;              /* filtered mean, D slots */
        double pr[NPA], uprev[d];
        double t = 0.0, dt = 0.0, dtcache = 0.0, qold_pb2 = 1.0, loglik = 0.0, gdiff = 0.0, ldiff = 0.0;
        long long traj = -1, naccept = 0, nreject = 0, nf = 0, iter = 0, tstop_idx = 0;
        int retcode = PN_RET_SUCCESS;
        bool have = false;

        for (;;) {
            /* ============ (A) fetch ============ */
            if (!have) {
                const unsigned long long idx = atomicAdd(P.work_counter, 1ULL);
                if ((long long)idx * P.traj_stride >= P.n_traj) break;
                traj = (long long)idx * P.traj_stride;
#pragma unroll
                for (int c = 0; c < D; c++) mus[c * NT] = P.init_mu[traj * D + c];
#pragma unroll
                for (int i = 0; i < NPA; i++) pr[i] = (i < Prob::n_p) ? P.p[traj * Prob::n_p + i] : 0.0;
#pragma unroll
                for (int i = 0; i < d; i++) uprev[i] = P.init_mu[traj * D + i];
#pragma unroll
                for (int s = 0; s < NPK; s++) my[s * NT] = 0.0; /* Sigma_0.R = 0 */
                dt = P.init_dt[traj];
                nf = Q + ((ADAPTIVE && !(P.dt0 > 0.0)) ? 2 : 0);
                t = P.t0;
                naccept = nreject = iter = tstop_idx = 0;
                loglik = gdiff = ldiff = 0.0;
                qold_pb2 = P.qoldinit_pb2;
                retcode = PN_RET_SUCCESS;
                dtcache = dt;
                have = true;
            }

            /* ============ (B) one attempt: mean, f, J, sigma^2, error estimate, controller ============ */
            bool finished = !(t < P.t1);
            bool accept = false;
            double h = 0.0, tstop = P.t1, EEst = 0.0, qq = 1.0, ediff = 0.0, lEE = 0.0;
            double hp[n], PIv[n], z[d], J[d * d];
#if PN_THREAD_EARLY_GRAM
            double M[D][D]; /* upper triangle: X'X of this attempt -> (accepted) Gram matrix -> predicted factor U */
#pragma unroll
            for (int a = 0; a < D; a++)
#pragma unroll
                for (int b = 0; b < D; b++) M[a][b] = 0.0;
#endif
#pragma unroll
            for (int k = 0; k < n; k++) hp[k] = (k == 0) ? 1.0 : 0.0;
#pragma unroll
            for (int c = 0; c < n; c++) PIv[c] = 0.0;
#pragma unroll
            for (int i = 0; i < d; i++) z[i] = 0.0;
#pragma unroll
            for (int i = 0; i < d * d; i++) J[i] = 0.0;
            for (int attempt = 0; attempt < 1 && !finished; attempt++) {
                double q11 = 0.0;
                while (tstop_idx < P.n_tstops && P.tstops[tstop_idx] <= t) tstop_idx++;
                tstop = P.t1;
                if (tstop_idx < P.n_tstops && P.tstops[tstop_idx] < P.t1) tstop = P.tstops[tstop_idx];
                iter++;
                if (iter > P.maxiters) {
                    retcode = PN_RET_MAXITERS;
                    finished = true;
                    break;
                }
                if (ADAPTIVE) {
                    dt = fmin(dt, P.dtmax);
                    dt = fmax(dt, P.dtmin);
                    dt = fmin(dt, tstop - t);
                } else {
                    dt = fmin(dtcache, tstop - t);
                }
                h = dt;
                if (h != h || !(t + h > t) || (ADAPTIVE && h <= P.dtmin)) {
                    retcode = PN_RET_DTLESSTHANMIN;
                    finished = true;
                    break;
                }
                const double tnew = t + h;
                hp[0] = 1.0;
#pragma unroll
                for (int k = 1; k < n; k++) hp[k] = hp[k - 1] * h * (1.0 / (double)k);
                const double sqh = h * pn_rsqrt(h);
#pragma unroll
                for (int c = 0; c < n; c++) PIv[c] = hp[Q - c] * sqh;
#if PN_THREAD_EARLY_GRAM
                gram_rows(my, hp, M); /* independent of everything below up to the accept decision */
#endif
                /* predict_mean!: mu^- = (T (x) I) mu */
                double mup[D];
#pragma unroll
                for (int c = 0; c < D; c++) mup[c] = mus[c * NT];
#pragma unroll
                for (int j = 0; j < n; j++)
#pragma unroll
                    for (int i = 0; i < d; i++) {
                        double s = mup[j * d + i];
#pragma unroll
                        for (int k = 0; k < n; k++)
                            if (k > j) s += hp[k > j ? k - j : 0] * mup[k * d + i];
                        mup[j * d + i] = s; /* ascending j in place: entries k > j are still the old ones */
                    }
                double fu[d];
                Prob::template f<double>(fu, mup, pr, tnew);
                Prob::jac(J, mup, pr, tnew);
                nf += 1;
#pragma unroll
                for (int i = 0; i < d; i++) z[i] = mup[d + i] - fu[i];
                if (ADAPTIVE || DYNAMIC) {
                    /* W = H Qh H' = qh00 J J' - qh01 (J + J') + qh11 I  (pn_small.cuh, same operation order) */
                    double W[d][d];
                    const double w00 = P.Qb[0] * PIv[0] * PIv[0], w01 = P.Qb[1] * PIv[0] * PIv[1], w11 = P.Qb[n + 1] * PIv[1] * PIv[1];
#pragma unroll
                    for (int a = 0; a < d; a++)
#pragma unroll
                        for (int b = 0; b < d; b++) {
                            W[a][b] = 0.0;
                            if (b < a) continue;
                            double jj = 0.0;
#pragma unroll
                            for (int i = 0; i < d; i++) jj += J[a + d * i] * J[b + d * i];
                            W[a][b] = w00 * jj - w01 * (J[a + d * b] + J[b + d * a]) + ((a == b) ? w11 : 0.0);
                        }
                    double wdiag[d], zt[d], winv[d];
#pragma unroll
                    for (int a = 0; a < d; a++) {
                        wdiag[a] = W[a][a];
                        zt[a] = z[a];
                    }
                    const bool ok = pn_chol<d>(W, winv);
                    pn_chol_solve<d>(W, winv, zt);
                    double qd = 0.0;
#pragma unroll
                    for (int a = 0; a < d; a++) qd += z[a] * zt[a];
                    ldiff = qd * (1.0 / (double)d);
                    if (P.forced_diffusion && naccept < P.n_forced) ldiff = P.forced_diffusion[naccept];
                    if (!ok) {
                        retcode = PN_RET_UNSTABLE;
                        finished = true;
                        break;
                    }
                    if (ADAPTIVE) {
                        double e2 = 0.0;
#pragma unroll
                        for (int i = 0; i < d; i++) {
                            const double isc = pn_rcp(P.abstol + fmax(fabs(mup[i]), fabs(uprev[i])) * P.reltol);
                            e2 += (ldiff * wdiag[i]) * (isc * isc);
                        }
                        e2 = e2 * (h * h) * (1.0 / (double)d);
                        EEst = (e2 > 0.0) ? e2 * pn_rsqrt(e2) : e2;
                        if (EEst != EEst) {
                            retcode = PN_RET_UNSTABLE;
                            finished = true;
                            break;
                        }
                    }
                }
                accept = !ADAPTIVE || EEst < 1.0;
                if (ADAPTIVE) {
                    if (EEst == 0.0) {
                        qq = P.inv_qmax;
                    } else {
                        lEE = log(EEst);
                        q11 = exp(P.beta1 * lEE);
                        qq = q11 * pn_rcp(qold_pb2 * P.gamma);
                        qq = fmax(P.inv_qmax, fmin(P.inv_qmin, qq));
                    }
                    if (!accept) {
                        nreject++;
                        dt = dt * pn_rcp(fmin(P.inv_qmin, q11 * P.inv_gamma));
                    }
                }
                ediff = DYNAMIC ? ldiff : P.initial_diffusion;
                if (P.forced_diffusion && naccept < P.n_forced) ediff = P.forced_diffusion[naccept];
                if (accept) {
#pragma unroll
                    for (int c = 0; c < D; c++) mus[c * NT] = mup[c]; /* the update below subtracts K z */
                }
            }
            if (accept) {
                /* loopfooter! accept branch, scalar part */
                double dtnew = dt;
                if (ADAPTIVE) {
                    if (P.qsteady_min <= qq && qq <= P.qsteady_max) qq = 1.0;
                    qold_pb2 = (EEst > P.qoldinit) ? exp(P.beta2 * lEE) : P.qoldinit_pb2;
                    dtnew = dt * pn_rcp(qq);
                }
                const double ttmp = t + h;
                const double ref = fmax(fabs(t), fabs(tstop));
                const double epsref = pn_ulp_above(ref);
                t = (fabs(ttmp - tstop) < 100.0 * epsref) ? tstop : ttmp;
                dt = dtnew;

                /* ============ (C) covariance predict + update ============ */
#if !PN_THREAD_EARLY_GRAM
                double M[D][D]; /* upper triangle: Gram matrix -> predicted factor U */
#pragma unroll
                for (int a = 0; a < D; a++)
#pragma unroll
                    for (int b = 0; b < D; b++) M[a][b] = 0.0;
#endif
                /* s2 Qh = s2 (P^-1 Qb P^-1) (x) I_d */
#pragma unroll
                for (int m = 0; m < n; m++)
#pragma unroll
                    for (int c = 0; c < n; c++) {
                        if (c < m) continue;
                        const double v = P.Qb[m * n + c] * (PIv[m] * ediff) * PIv[c];
#pragma unroll
                        for (int i = 0; i < d; i++) M[m * d + i][c * d + i] = PN_THREAD_EARLY_GRAM ? M[m * d + i][c * d + i] + v : v;
                    }
#if !PN_THREAD_EARLY_GRAM
                gram_rows(my, hp, M);
#endif
                /* Cholesky in registers (right-looking, upper) */
                bool okc = true;
                pn_static_for<0, D>([&](auto kc) {
                    constexpr int k = decltype(kc)::value;
                    const double piv = M[k][k];
                    okc = okc && (piv > 0.0);
                    const double ri = pn_rsqrt(piv);
                    M[k][k] = piv * ri;
#pragma unroll
                    for (int j = 0; j < D; j++)
                        if (j > k) M[k][j] *= ri;
#pragma unroll
                    for (int i = 0; i < D; i++)
#pragma unroll
                        for (int j = 0; j < D; j++)
                            if (i > k && j >= i) M[i][j] = fma(-M[k][i], M[k][j], M[i][j]);
                });
                if (!okc || ediff == 0.0) { /* predict.jl:90 / :71-74: cold path, the stack in local memory */
                    double X[D * D];
                    fallback(P, my, h, ediff, X);
#pragma unroll
                    for (int a = 0; a < D; a++)
#pragma unroll
                        for (int b = 0; b < D; b++)
                            if (b >= a) M[a][b] = X[a * D + b];
                }

                /* C2 = U H' on the first 2d rows (zero below), S = C2'C2 */
                double C2[RU][d], S[d][d];
#pragma unroll
                for (int r = 0; r < RU; r++)
#pragma unroll
                    for (int a = 0; a < d; a++) {
                        double s = (d + a >= r) ? M[r][d + a] : 0.0;
#pragma unroll
                        for (int i = 0; i < d; i++)
                            if (i >= r) s -= M[r][i] * J[a + d * i];
                        C2[r][a] = s;
                    }
#pragma unroll
                for (int a = 0; a < d; a++)
#pragma unroll
                    for (int b = 0; b < d; b++) {
                        S[a][b] = 0.0;
                        if (b < a) continue;
                        double s = 0.0;
#pragma unroll
                        for (int r = 0; r < RU; r++) s += C2[r][a] * C2[r][b];
                        S[a][b] = s;
                    }
                double tr = 0.0;
#pragma unroll
                for (int a = 0; a < d; a++) tr += S[a][a];
                const bool passthrough = (tr == 0.0); /* Sigma^- H' == 0 (update.jl:82-89) */
                if (passthrough) {
#pragma unroll
                    for (int a = 0; a < d; a++) S[a][a] = 1.0;
                }
                double sinv[d], zt[d];
                const bool okS = pn_chol<d>(S, sinv);
#pragma unroll
                for (int a = 0; a < d; a++) zt[a] = z[a];
                pn_chol_solve<d>(S, sinv, zt);
                double quad = 0.0, dprod = 1.0;
#pragma unroll
                for (int a = 0; a < d; a++) {
                    quad += z[a] * zt[a];
                    dprod *= S[a][a];
                }
                if (!okS) retcode = PN_RET_UNSTABLE;
                if (passthrough) {
                    bool zz = true;
#pragma unroll
                    for (int a = 0; a < d; a++) zz = zz && (z[a] == 0.0);
                    loglik += zz ? PN_INF : -PN_INF;
                } else {
                    loglik += -0.5 * quad - 0.5 * (double)d * 1.8378770664093453 - log(dprod);
                    if (!DYNAMIC) {
                        const double inc = quad * (1.0 / (double)d);
                        gdiff = (naccept == 0) ? inc : gdiff + (inc - gdiff) / (double)naccept;
                    }
                }
                double V[RU][d];
#pragma unroll
                for (int r = 0; r < RU; r++) {
                    double v[d];
#pragma unroll
                    for (int a = 0; a < d; a++) v[a] = C2[r][a];
                    pn_chol_solve<d>(S, sinv, v);
#pragma unroll
                    for (int a = 0; a < d; a++) V[r][a] = v[a];
                }
                /* column by column: K[c] = U[:, c]' V ; mu[c] -= K[c] z ; R+[r][c] = U[r][c] - C2[r] K[c]' (r < 2d); rows >= 2d unchanged */
                bool nanu = false;
                pn_static_for<0, D>([&](auto cc) {
                    constexpr int c = decltype(cc)::value;
                    double Kc[d];
#pragma unroll
                    for (int a = 0; a < d; a++) {
                        double s = 0.0;
#pragma unroll
                        for (int r = 0; r < RU; r++)
                            if (r <= c) s += M[r][c] * V[r][a];
                        Kc[a] = s;
                    }
                    double m = mus[c * NT];
#pragma unroll
                    for (int a = 0; a < d; a++) m -= Kc[a] * z[a];
                    mus[c * NT] = m;
                    if (c < d) {
                        uprev[c < d ? c : 0] = m;
                        nanu = nanu || (m != m);
                    }
#pragma unroll
                    for (int r = 0; r < D; r++) {
                        if (r < RU) {
                            double v = (c >= r) ? M[r][c] : 0.0;
#pragma unroll
                            for (int a = 0; a < d; a++) v -= C2[r][a] * Kc[a];
                            my[slot(r, c) * NT] = v;
                        } else if (c >= r) {
                            my[slot(r, c) * NT] = M[r][c];
                        }
                    }
                });
                naccept++;
                if (nanu) retcode = PN_RET_UNSTABLE;
                if (retcode != PN_RET_SUCCESS) finished = true;
                if (!(t < P.t1)) finished = true;
            }

            /* ============ (D) postamble ============ */
            if (finished) {
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
#pragma unroll
                for (int a = 0; a < d; a++) {
                    P.u_mean[traj * d + a] = mus[a * NT];
#pragma unroll
                    for (int b = 0; b < d; b++) {
                        double s = 0.0; /* columns < d are touched by the rows < 2d only */
#pragma unroll
                        for (int r = 0; r < RU; r++) s += my[slot(r, a) * NT] * my[slot(r, b) * NT];
                        P.u_cov[traj * d * d + a * d + b] = s * sc;
                    }
                }
                P.loglik[traj] = loglik;
                P.diffusion[traj] = fd;
                P.dt_last[traj] = dt;
                P.naccept[traj] = naccept;
                P.nreject[traj] = nreject;
                P.nf[traj] = nf;
                P.retcode[traj] = retcode;
                if (P.x_mean) {
#pragma unroll
                    for (int c = 0; c < D; c++) P.x_mean[traj * D + c] = mus[c * NT];
                }
                if (P.x_covfac) {
                    const double ss = sqrt(sc);
#pragma unroll
                    for (int r = 0; r < D; r++)
#pragma unroll
                        for (int c = 0; c < D; c++)
                            P.x_covfac[(traj * D + r) * D + c] = stored(r, c) ? my[slot(r, stored(r, c) ? c : r) * NT] * ss : 0.0;
                }
                have = false;
            }
        }
    }

    /* predict_cov! when the Cholesky route does not apply: triangularize!([R Ah'; sqrt(s2) Qh.R]) (predict.jl:90), and the
       zero-diffusion case (predict.jl:71-74: Sigma^- = Ah Sigma Ah', brought to triangular form because the update above reads the
       structure).  One thread, the stack in local memory: a cold path (first steps of a trajectory, loss of definiteness). */
    static __device__ __noinline__ void fallback(const PnParams& P, const double* my, double h, double ediff, double* X) {
        double hp[n], PIv[n];
        hp[0] = 1.0;
        for (int k = 1; k < n; k++) hp[k] = hp[k - 1] * h * (1.0 / (double)k);
        const double sqh = h * pn_rsqrt(h);
        for (int c = 0; c < n; c++) PIv[c] = hp[Q - c] * sqh;
        for (int r = 0; r < D; r++) {
            double x[D];
            for (int c = 0; c < D; c++) x[c] = stored(r, c) ? my[slot(r, c) * NT] : 0.0;
            for (int i = 0; i < d; i++)
                for (int j = 0; j < n; j++) {
                    double s = x[j * d + i];
                    for (int k = j + 1; k < n; k++) s += hp[k - j] * x[k * d + i];
                    x[j * d + i] = s;
                }
            for (int c = 0; c < D; c++) X[r * D + c] = x[c];
        }
        double bf[n * n];
        const double sd = (ediff == 1.0) ? 1.0 : sqrt(ediff);
        for (int m = 0; m < n; m++)
            for (int c = 0; c < n; c++) bf[m * n + c] = (c >= m) ? P.U[m * n + c] * PIv[c] * sd : 0.0;
        pn_triangularize_serial<D, d, n>(X, bf);
    }
};

template <class Prob, int Q, bool ADAPTIVE, bool DYNAMIC>
__device__ __forceinline__ void pn_thread_entry(const PnParams& P) {
    extern __shared__ double pn_thread_shared[]; /* pn_thread_sm_doubles(d, n) * PN_THREAD_NT doubles */
    PnThread<Prob, Q, ADAPTIVE, DYNAMIC>::run(P, pn_thread_shared);
}

#endif /* PN_THREAD_CUH */
