// pn_smooth.cuh -- RTS smoother sweep (backward pass), one warp per trajectory, matrices staged in shared memory.
//
// Reference behaviour replaced (per trajectory, i = N-1 ... 1):
//   compute_backward_kernel!   src/filtering/markov_kernel.jl:190-235   G, b, Lambda.R (2D x D)
//   marginalize!               src/filtering/markov_kernel.jl:68-101    mu_s = G mu_s+ + b ; R_s = qr([R_s+ G' ; Lambda.R]).R
//   smooth_solution!           src/integrator_utils.jl:98-121           sweep + pu = SolProj * x_smooth
//   calibrate_solution!        src/integrator_utils.jl:46-65            sqrt(sigma_hat^2) scaling (FixedDiffusion)
//
// Data layout in HBM: the forward (filter) pass saved, per accepted step and trajectory-contiguous (CSR offsets),
// the filtered mean (D), the filtered right factor (D x D row-major), the step h and the diffusion used.  That is
// D^2 + D + 2 doubles per step (1.25 KB at D = 12) instead of the reference's 4D^2 + 2D (4.8 KB): the backward
// kernel of step i is *recomputed* here from (x_filt[i], h_i, sigma_i^2) -- it is a pure function of them -- rather
// than stored.  One warp streams its trajectory's records back to front; every record is a contiguous, fully
// coalesced read, and the smoothed state is written over the filtered one in place.
//
// This first version is size-generic at run time (d, q are arguments, loops are not unrolled) and keeps its
// working set in shared memory; the arithmetic follows the reference formulas literally.
//
// NVRTC-compatible: no standard headers.
#ifndef PN_SMOOTH_CUH
#define PN_SMOOTH_CUH
#include "pn_params.h"
#ifndef PN_SMOOTH_DECL_ONLY
#include "pn_lti.cuh"
#endif

struct PnSmoothParams {
    long long n_traj;
    int d, q;
    int dynamic;        /* DynamicDiffusion: per-step sigma^2 from s_diffusion; else initial_diffusion */
    int calibrate;      /* FixedDiffusion(calibrate=true): scale factors by sqrt(global MLE) */
    double initial_diffusion;
    const long long* step_offsets; /* [n_traj+1] */
    const long long* counts;       /* [n_traj] accepted steps per trajectory when the segments are over-allocated (pitch layout);
                                      null: segment i is exactly [off[i], off[i+1]) */
    const double* s_h;             /* [total] h of the step k -> k+1 stored at slot k */
    const double* s_diffusion;     /* [total] */
    const double* gdiff;           /* [n_traj] calibrated diffusion sigma_hat^2 * initial_diffusion (sol.diffusions); the covariance
                                      scale of calibrate_solution! is gdiff / initial_diffusion (integrator_utils.jl:46-57) */
    const double* s_diffusion_mv;  /* [total][d] per-component diffusions (multivariate models; null otherwise) */
    const double* gdiff_mv;        /* [n_traj][d] per-component global scale (FixedMVDiffusion, calibrate)       */
    double* s_x_mean;              /* [total][D]    in: filtered, out: smoothed */
    double* s_x_covfac;            /* [total][D][D] in: filtered, out: smoothed (row-major right factor) */
    double* s_u_mean;              /* [total][d] out */
    double* s_u_cov;               /* [total][d][d] out */
    double L[PN_MAX_N * PN_MAX_N]; /* IWP left sqrt factor, row-major n x n */
    double U[PN_MAX_N * PN_MAX_N]; /* upper-triangular right factor of the same covariance (pn_params.h) */
    double rate;              /* IOUP prior: scalar rate parameter and the Gauss-Legendre rule (pn_ioup.cuh) */
    double glx[16], glw[16];
    unsigned long long* work_counter;
    /* Fenrir data log-likelihood (src/data_likelihoods/fenrir.jl:68-137, fit_pnsolution_to_data!): a non-null fenrir_ll
       turns the shared-memory sweep into the backward pass that starts from the filtering estimates, marginalizes with the
       backward kernels and, at every data time, applies a Kalman update with observation matrix obs_H and noise factor
       obs_Rn, accumulating its log-likelihood.  Nothing is written back to the records in that mode. */
    const double* s_t;      /* [total] saved times */
    const double* data_t;   /* [n_data] ascending; every entry is a saved time of every trajectory (tstops) */
    const double* data_u;   /* [n_data][n_obs] */
    const double* obs_H;    /* [n_obs][D] row-major: proj * SolProj */
    const double* obs_Rn;   /* [n_obs][n_obs] row-major right factor of the observation-noise covariance */
    long long n_data;
    int n_obs;
    const int* retcode;     /* [n_traj] return codes of the solve: -Inf unless Success (fenrir.jl:55-59) */
    double* fenrir_ll;      /* [n_traj] out */
    double* workspace;      /* null: the warp-per-trajectory sweep keeps its work matrices in shared memory (D <= 56); else a global buffer
                               of gridDim.x * PN_SM_WARPS * per_warp doubles (large states, e.g. Pleiades D = 112: 7 D^2 doubles per
                               warp do not fit in shared memory; the buffer stays L2-resident) */
    int write_x;            /* 1: the smoothed state records (s_x_mean, s_x_covfac) are somebody's input or output (save_state, dense
                               output); 0: a sweep may leave the records alone and only emit sol.u / sol.pu -- a third of its HBM bytes */
    /* general LTI prior (IOUP with a vector / matrix rate, or the Jacobian-updated rate; pn_lti.cuh): the transition of a step is a
       dense pair (Ah, Qh.R) recomputed from (h, rate matrix); shared-memory sweep only */
    int gen_prior;
    const double* rate_matrix; /* [d][d] row-major (null: rate I_d) */
    const double* s_rate;      /* [total][d][d] per-step rate matrices saved by the filter (update_rate_parameter); null: rate_matrix */
    int kron_logdet;        /* 1: IsometricKronecker layout (EK0 + IWP + scalar diffusion): the reference's data update runs on the
                               n x n factors and takes logdet of the 1 x 1 factor of S without the factor d (update.jl:140-148,
                               182-194) -- log det S / n_obs instead of log det S */
};
/* extra doubles per warp with a general prior (host and kernel must agree) */
#define PN_SMOOTH_GEN_EXTRA(D, d, n) (2 * (D) * (D) + 2 * (d) * (d) + (n))
/* extra shared-memory doubles per warp in Fenrir mode (host and kernel must agree) */
#define PN_FENRIR_EXTRA(D, o) (3 * (D) * (o) + ((D) + (o)) * (o) + 4 * (o))

#define PN_SM_WARPS 4

#ifndef PN_SMOOTH_DECL_ONLY

__device__ __forceinline__ double pn_warp_sum(double x) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// Householder QR of the m x D row-major matrix A (leading dimension D) in shared memory, warp-cooperative.
// On return rows 0..D-1 hold the upper-triangular factor (below-diagonal entries zeroed).
__device__ void pn_sm_qr(double* A, int m, int D, int lane) {
    for (int k = 0; k < D; k++) {
        double part = 0.0;
        for (int r = k + lane; r < m; r += 32) part += A[r * D + k] * A[r * D + k];
        const double total = pn_warp_sum(part);
        const double alpha = A[k * D + k];
        __syncwarp();
        if (total > 0.0) {
            const double nrm = sqrt(total);
            const double beta = -copysign(nrm, alpha);
            const double vk = alpha - beta;
            const double gam = 1.0 / (total + fabs(alpha) * nrm);
            /* each lane owns columns j = k+1+lane, ... ; v_r = A[r][k] (r > k), v_k = vk */
            for (int j = k + 1 + lane; j < D; j += 32) {
                double s = vk * A[k * D + j];
                for (int r = k + 1; r < m; r++) s += A[r * D + k] * A[r * D + j];
                s *= gam;
                A[k * D + j] -= s * vk;
                for (int r = k + 1; r < m; r++) A[r * D + j] -= s * A[r * D + k];
            }
            __syncwarp();
            for (int r = k + lane; r < m; r += 32) A[r * D + k] = (r == k) ? beta : 0.0;
        }
        __syncwarp();
    }
}

// B (rows x D) <- B * (U'U)^{-1} with U upper triangular (D x D row-major); each lane owns whole rows of B.
__device__ void pn_sm_rdiv_chol(double* B, int rows, const double* U, int D, int lane) {
    for (int r = lane; r < rows; r += 32) {
        double* x = B + r * D;
        for (int c = 0; c < D; c++) { /* x U = b */
            double s = x[c];
            for (int k = 0; k < c; k++) s -= x[k] * U[k * D + c];
            x[c] = s / U[c * D + c];
        }
        for (int c = D - 1; c >= 0; c--) { /* y U' = x */
            double s = x[c];
            for (int k = c + 1; k < D; k++) s -= x[k] * U[c * D + k];
            x[c] = s / U[c * D + c];
        }
    }
    __syncwarp();
}

// measure_and_update! (fenrir.jl:127-137) + update!(...; R) (src/filtering/update.jl:65-139, QR route :132-136) on the
// posterior (mus, Rs) in shared memory; returns the log-likelihood of the observation u (same value on all lanes).
__device__ double pn_sm_fenrir_update(const PnSmoothParams& P, const double* u, double* mus, double* Rs, double* Xs, double* F, int D,
                                      int lane) {
    const int o = P.n_obs;
    double* Cb = F;                 /* D x o   C = R H'                  */
    double* Kb = Cb + D * o;        /* D x o   K                         */
    double* K1 = Kb + D * o;        /* D x o   K Rn'                     */
    double* Sb = K1 + D * o;        /* (D + o) x o stack -> S.R          */
    double* z = Sb + (D + o) * o;   /* o                                 */
    double* zt = z + o;             /* o                                 */
    double nrm = 0.0;
    for (int i = lane; i < D * D; i += 32) nrm += fabs(Rs[i]);
    nrm = pn_warp_sum(nrm);
    for (int a = lane; a < o; a += 32) {
        double sacc = 0.0;
        for (int c = 0; c < D; c++) sacc += P.obs_H[a * D + c] * mus[c];
        z[a] = sacc - u[a];
        zt[a] = z[a];
    }
    __syncwarp();
    if (nrm == 0.0) { /* update.jl:82-89: Sigma == 0 -> copy, +-Inf */
        double nz = 0.0;
        for (int a = 0; a < o; a++) nz += fabs(z[a]);
        return (nz == 0.0) ? PN_INF : -PN_INF;
    }
    for (int i = lane; i < D * o; i += 32) {
        const int r = i / o, a = i % o;
        double sacc = 0.0;
        for (int c = 0; c < D; c++) sacc += Rs[r * D + c] * P.obs_H[a * D + c];
        Cb[i] = sacc;
        Sb[i] = sacc;
    }
    for (int i = lane; i < o * o; i += 32) Sb[D * o + i] = P.obs_Rn[i];
    __syncwarp();
    pn_sm_qr(Sb, D + o, o, lane); /* make_obscov_sqrt: S.R = qr([R H'; Rn]).R */
    for (int i = lane; i < D * o; i += 32) {
        const int c = i / o, a = i % o;
        double sacc = 0.0;
        for (int r = 0; r < D; r++) sacc += Rs[r * D + c] * Cb[r * o + a];
        Kb[i] = sacc;
    }
    __syncwarp();
    pn_sm_rdiv_chol(Kb, D, Sb, o, lane); /* K = Sigma H' S^-1 */
    pn_sm_rdiv_chol(zt, 1, Sb, o, lane); /* S^-1 z            */
    double quad = 0.0, logdet = 0.0;
    for (int a = 0; a < o; a++) {
        quad += z[a] * zt[a];
        logdet += log(fabs(Sb[a * o + a]));
    }
    if (P.kron_logdet) logdet /= (double)o;
    const double ll = -0.5 * quad - 0.5 * (double)o * 1.8378770664093453 - logdet;
    for (int i = lane; i < D * o; i += 32) {
        const int c = i / o, a = i % o;
        double sacc = 0.0;
        for (int b = 0; b < o; b++) sacc += Kb[c * o + b] * P.obs_Rn[a * o + b];
        K1[i] = sacc; /* (K Rn')[c][a] */
    }
    /* stack [ R (I - K H)' ; (K Rn')' ] */
    for (int i = lane; i < D * D; i += 32) {
        const int r = i / D, c = i % D;
        double sacc = Rs[i];
        for (int a = 0; a < o; a++) sacc -= Cb[r * o + a] * Kb[c * o + a];
        Xs[i] = sacc;
    }
    __syncwarp();
    for (int i = lane; i < o * D; i += 32) {
        const int a = i / D, c = i % D;
        Xs[D * D + i] = K1[c * o + a];
    }
    for (int c = lane; c < D; c += 32) {
        double sacc = mus[c];
        for (int a = 0; a < o; a++) sacc -= Kb[c * o + a] * z[a];
        mus[c] = sacc;
    }
    __syncwarp();
    pn_sm_qr(Xs, D + o, D, lane);
    for (int i = lane; i < D * D; i += 32) Rs[i] = Xs[i];
    __syncwarp();
    return ll;
}

__global__ void __launch_bounds__(32 * PN_SM_WARPS) pn_smooth_kernel(const __grid_constant__ PnSmoothParams P) {
    extern __shared__ double pn_sm_shared[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int d = P.d, q = P.q, n = q + 1, D = d * n, DD = D * D;
    /* per-warp workspace */
    const bool fen = P.fenrir_ll != nullptr;
    const int fen_extra = fen ? PN_FENRIR_EXTRA(D, P.n_obs) : 0;
    const int per_warp = 7 * DD + 4 * D + 2 * n * n + 2 * n + fen_extra + (P.gen_prior ? PN_SMOOTH_GEN_EXTRA(D, d, n) : 0);
    double* W = P.workspace ? P.workspace + ((size_t)blockIdx.x * PN_SM_WARPS + warp) * per_warp : pn_sm_shared + (size_t)warp * per_warp;
    double* Rk = W;            /* D x D  filtered factor at k                 */
    double* Xs = Rk + DD;      /* 3D x D stack; first used as 2D x D predict  */
    double* Gm = Xs + 3 * DD;  /* D x D  smoothing gain                       */
    double* Rs = Gm + DD;      /* D x D  smoothed factor at k+1               */
    double* Tm = Rs + DD;      /* D x D  scratch                              */
    double* mus = Tm + DD;     /* D smoothed mean at k+1 */
    double* muk = mus + D;     /* D filtered mean at k   */
    double* mup = muk + D;     /* D predicted mean       */
    double* bv = mup + D;      /* D */
    double* T1 = bv + D;       /* n x n: Ah factor  T[i][j] = h^(j-i)/(j-i)! */
    double* LQ = T1 + n * n;   /* n x n: (L P^-1)                          */
    double* hp = LQ + n * n;   /* n */
    double* PIv = hp + n;      /* n */
    double* Fw = PIv + n;      /* Fenrir workspace (fen only) */
    double* Ahd = Fw + fen_extra; /* general prior: dense Ah, Qh.R, rate matrix, h * rate, diagonal of P^-1 */
    double* Qcd = Ahd + DD;
    double* Rmg = Qcd + DD;
    double* Zmg = Rmg + d * d;
    double* PIg = Zmg + d * d;
    const bool gen = P.gen_prior != 0;

    for (;;) {
        unsigned long long idx = 0;
        if (lane == 0) idx = atomicAdd(P.work_counter, 1ULL);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if ((long long)idx >= P.n_traj) break;
        const long long tr = (long long)idx;
        const long long a0 = P.step_offsets[tr], a1 = P.counts ? a0 + P.counts[tr] + 1 : P.step_offsets[tr + 1];
        const long long last = a1 - 1;
        const double sc = (!P.dynamic && P.calibrate) ? sqrt(P.gdiff[tr] / P.initial_diffusion) : 1.0;
        /* x_smooth[N] = x_filt[N] */
        for (int i = lane; i < DD; i += 32) Rs[i] = P.s_x_covfac[last * DD + i];
        for (int i = lane; i < D; i += 32) mus[i] = P.s_x_mean[last * D + i];
        __syncwarp();
        double LL = 0.0;
        long long data_idx = P.n_data - 1;
        if (fen && data_idx >= 0 && P.s_t[last] == P.data_t[data_idx]) /* fenrir.jl:87-97 */
            LL += pn_sm_fenrir_update(P, P.data_u + data_idx * P.n_obs, mus, Rs, Xs, Fw, D, lane);
        data_idx--; /* unconditional (fenrir.jl:99): a last observation that is not at sol.t[end] is skipped, as in the reference */
        /* outputs of the last point (calibrated) */
        if (!fen)
        for (int i = lane; i < d * d; i += 32) {
            const int a = i / d, b = i % d;
            double s = 0.0;
            for (int r = 0; r < D; r++) s += Rs[r * D + a] * Rs[r * D + b];
            P.s_u_cov[last * d * d + i] = s * sc * sc;
        }
        if (!fen)
        for (int i = lane; i < d; i += 32) P.s_u_mean[last * d + i] = mus[i];
        if (sc != 1.0 && !fen)
            for (int i = lane; i < DD; i += 32) P.s_x_covfac[last * DD + i] = Rs[i] * sc;
        __syncwarp();

        for (long long k = last - 1; k >= a0; k--) {
            const double h = P.s_h[k];
            const double ediff = P.dynamic ? P.s_diffusion[k] : P.initial_diffusion;
            for (int i = lane; i < DD; i += 32) Rk[i] = P.s_x_covfac[k * DD + i];
            for (int i = lane; i < D; i += 32) muk[i] = P.s_x_mean[k * D + i];
            if (lane == 0) {
                hp[0] = 1.0;
                for (int j = 1; j < n; j++) hp[j] = hp[j - 1] * h / (double)j;
                const double sqh = sqrt(h);
                for (int c = 0; c < n; c++) PIv[c] = hp[q - c] * sqh;
            }
            __syncwarp();
            for (int i = lane; i < n * n; i += 32) {
                const int r = i / n, c = i % n;
                T1[i] = (c >= r) ? hp[c - r] : 0.0;
                LQ[i] = (c <= r) ? P.L[r * n + c] * PIv[c] : 0.0;
            }
            __syncwarp();
            if (gen && h != 0.0) { /* dense transition of this step (Xs, Gm are free here) */
                for (int i = lane; i < d * d; i += 32)
                    Rmg[i] = P.s_rate ? P.s_rate[k * d * d + i] : (P.rate_matrix ? P.rate_matrix[i] : ((i / d == i % d) ? P.rate : 0.0));
                __syncwarp();
                pn_gen_transition(Ahd, Qcd, Rmg, h, d, n, Xs, Xs + DD, Xs + 2 * DD, Gm, Zmg, PIg, lane);
            }
            if (h == 0.0) { /* integrator_utils.jl:109-112: x_smooth[k] = x_smooth[k+1] */
                if (fen) continue; /* fenrir.jl:103-106 */
                for (int i = lane; i < DD; i += 32) P.s_x_covfac[k * DD + i] = Rs[i] * sc;
                for (int i = lane; i < D; i += 32) P.s_x_mean[k * D + i] = mus[i];
            } else {
                const double sd = (ediff == 1.0) ? 1.0 : sqrt(ediff);
                /* X = R_k Ah' (rows 0..D-1 of the stack), bottom = sd * (LQ (x) I)  (predict.jl:75-83) */
                for (int i = lane; i < DD; i += 32) {
                    const int r = i / D, c = i % D, j = c / d, ii = c % d;
                    double s = 0.0;
                    if (gen) {
                        for (int kk = 0; kk < D; kk++) s += Rk[r * D + kk] * Ahd[c * D + kk];
                        Xs[i] = s;
                        Tm[i] = s;
                        Xs[DD + i] = sd * Qcd[i];
                        continue;
                    }
                    for (int jj = j; jj < n; jj++) s += T1[j * n + jj] * Rk[r * D + jj * d + ii];
                    Xs[i] = s;
                    Tm[i] = s; /* keep X for G */
                    const int m = r / d, i2 = r % d;
                    Xs[DD + i] = (i2 == ii) ? sd * LQ[m * n + j] : 0.0;
                }
                /* mu^- = Ah mu_k */
                for (int i = lane; i < D; i += 32) {
                    const int j = i / d, ii = i % d;
                    double s = 0.0;
                    if (gen)
                        for (int kk = 0; kk < D; kk++) s += Ahd[i * D + kk] * muk[kk];
                    else
                    for (int jj = j; jj < n; jj++) s += T1[j * n + jj] * muk[jj * d + ii];
                    mup[i] = s;
                }
                __syncwarp();
                pn_sm_qr(Xs, 2 * D, D, lane); /* R^- in rows 0..D-1 of Xs (triangularize!) */
                /* G = R_k' X  (markov_kernel.jl:210-211) */
                double gpart = 0.0;
                for (int i = lane; i < DD; i += 32) {
                    const int a = i / D, b = i % D;
                    double s = 0.0;
                    for (int r = 0; r < D; r++) s += Rk[r * D + a] * Tm[r * D + b];
                    Gm[i] = s;
                    gpart += fabs(s);
                }
                gpart = pn_warp_sum(gpart);
                __syncwarp();
                if (gpart != 0.0) pn_sm_rdiv_chol(Gm, D, Xs, D, lane); /* :212-214 */
                /* b = mu_k - G mu^-  (:217-218) */
                for (int i = lane; i < D; i += 32) {
                    double s = 0.0;
                    for (int c = 0; c < D; c++) s += Gm[i * D + c] * mup[c];
                    bv[i] = muk[i] - s;
                }
                __syncwarp();
                /* Lambda.R = [ R_k (I - G Ah)' ; sd QhR G' ]  -> rows D..3D-1 of the smoother stack (:221-232) */
                /* first (I - G Ah)' into Tm: Tm[c][a] = delta - (G Ah)[a][c] ; (G Ah)[a][c=(j,i)] = sum_{jj<=j} G[a][(jj,i)] T[jj][j] */
                for (int i = lane; i < DD; i += 32) {
                    const int c = i / D, a = i % D, j = c / d, ii = c % d;
                    double s = 0.0;
                    if (gen)
                        for (int kk = 0; kk < D; kk++) s += Gm[a * D + kk] * Ahd[kk * D + c];
                    else
                    for (int jj = 0; jj <= j; jj++) s += Gm[a * D + jj * d + ii] * T1[jj * n + j];
                    Tm[i] = ((a == c) ? 1.0 : 0.0) - s;
                }
                __syncwarp();
                /* stack rows: [0,D): R_s+ G' ; [D,2D): R_k Tm ; [2D,3D): sd QhR G' */
                for (int i = lane; i < DD; i += 32) {
                    const int r = i / D, a = i % D;
                    double s0 = 0.0, s1 = 0.0;
                    for (int c = 0; c < D; c++) {
                        s0 += Rs[r * D + c] * Gm[a * D + c];
                        s1 += Rk[r * D + c] * Tm[c * D + a];
                    }
                    const int m = r / d, i2 = r % d;
                    double s2 = 0.0;
                    if (gen)
                        for (int kk = r; kk < D; kk++) s2 += Qcd[r * D + kk] * Gm[a * D + kk];
                    else
                    for (int j = 0; j <= m; j++) s2 += LQ[m * n + j] * Gm[a * D + j * d + i2];
                    Xs[i] = s0;
                    Xs[DD + i] = s1;
                    Xs[2 * DD + i] = sd * s2;
                }
                /* mu_s[k] = G mu_s[k+1] + b  (marginalize_mean!, :72-82) */
                for (int i = lane; i < D; i += 32) {
                    double s = bv[i];
                    for (int c = 0; c < D; c++) s += Gm[i * D + c] * mus[c];
                    mup[i] = s;
                }
                __syncwarp();
                pn_sm_qr(Xs, 3 * D, D, lane); /* marginalize_cov!: triangularize the 3D x D stack (:95-99) */
                for (int i = lane; i < DD; i += 32) {
                    const double v = Xs[i];
                    Rs[i] = v;
                    if (!fen) P.s_x_covfac[k * DD + i] = v * sc;
                }
                for (int i = lane; i < D; i += 32) {
                    mus[i] = mup[i];
                    if (!fen) P.s_x_mean[k * D + i] = mup[i];
                }
            }
            __syncwarp();
            if (fen) { /* fenrir.jl:110-123: data update at this time point */
                if (data_idx >= 0 && P.s_t[k] == P.data_t[data_idx]) {
                    const double ll = pn_sm_fenrir_update(P, P.data_u + data_idx * P.n_obs, mus, Rs, Xs, Fw, D, lane);
                    if (!(ll == PN_INF || ll == -PN_INF)) LL += ll; /* !isinf(ll): a NaN propagates (fenrir.jl:118-120) */
                    data_idx--;
                }
                continue;
            }
            /* pu[k] = SolProj x_smooth[k]  (integrator_utils.jl:117-118) */
            for (int i = lane; i < d * d; i += 32) {
                const int a = i / d, b = i % d;
                double s = 0.0;
                for (int r = 0; r < D; r++) s += Rs[r * D + a] * Rs[r * D + b];
                P.s_u_cov[k * d * d + i] = s * sc * sc;
            }
            for (int i = lane; i < d; i += 32) P.s_u_mean[k * d + i] = mus[i];
            __syncwarp();
        }
        if (fen && lane == 0) {
            const bool ok = (!P.retcode || P.retcode[tr] == PN_RET_SUCCESS) && data_idx < 0;
            P.fenrir_ll[tr] = ok ? LL : -PN_INF;
        }
    }
}

#endif /* PN_SMOOTH_DECL_ONLY */

#endif /* PN_SMOOTH_CUH */
