// pn_smooth_reg.cuh -- RTS smoother sweep for small filter states (D = d(q+1) <= 24), register-resident.
//
// A group of G lanes owns one trajectory and walks its saved records back to front.  Matrices are
// row-distributed over the group exactly as in pn_small.cuh (lane gl owns rows r == gl mod G), Householder
// reflections are reduced with G-lane butterflies, and the three D x D matrices every lane of the group must
// see in full (R^-, Y, the smoothed factor of the previous point) are exchanged through a small per-trajectory
// shared-memory region that is read with broadcast loads.
//
// Reference behaviour replaced (per trajectory, i = N-1 ... 1):
//   compute_backward_kernel!   src/filtering/markov_kernel.jl:190-235   G, b, Lambda.R (2D x D)
//   marginalize!               src/filtering/markov_kernel.jl:68-101    mu_s = G mu_s+ + b ; R_s = qr([R_s+ G' ; Lambda.R]).R
//   smooth_solution!           src/integrator_utils.jl:98-121           sweep, pu = SolProj x_smooth, dt == 0 copies
//   calibrate_solution!        src/integrator_utils.jl:46-65            sqrt(sigma_hat^2) scaling (FixedDiffusion)
//
// One orthogonal transformation produces everything the backward kernel needs.  With X = R Ah', B the
// upper-triangular right factor of sigma^2 Qh and R the filtered factor, the first D Householder reflections of
//
//        [ X  R ]        [ R^-  Y        ]       R^-'R^- = Sigma^-          (predict_cov!, predict.jl:62-94)
//        [ B  0 ]  --->  [ 0    Lambda_f ]       R^-'Y   = Ah Sigma   =>  G = Y' R^-^-T   (markov_kernel.jl:210-214)
//                                                Lambda_f'Lambda_f = Sigma - G Sigma^- G' = Lambda.R'Lambda.R  (:221-232)
//
// so   R_s+ G' = (R_s+ R^-^-1) Y   (row-local triangular solve, then a product against broadcast rows of Y),
//      mu_s    = mu + Y' R^-^-T (mu_s+ - Ah mu)                         (= G mu_s+ + b, :72-82 and :217-218),
//      R_s     = qr([R_s+ G' ; Lambda_f]).R                             (marginalize_cov!, :95-99, 2D x D instead of 3D x D).
// G itself is never formed.  The backward kernel is recomputed from the filtered record (x_filt[k], h_k, sigma_k^2)
// -- it is a pure function of it -- so the forward pass stores D^2 + D + 2 doubles per step instead of the
// reference's 4D^2 + 2D (integrator_utils.jl:153-167).
//
// NVRTC-compatible: no standard headers.
#ifndef PN_SMOOTH_REG_CUH
#define PN_SMOOTH_REG_CUH
#include "pn_params.h"
#include "pn_small.cuh" /* pn_gsum, pn_gbcast, pn_rcp, pn_rsqrt */
#ifndef PN_SMOOTH_DECL_ONLY
#define PN_SMOOTH_DECL_ONLY
#define PN_SMOOTH_REG_UNDEF_DECL
#endif
#include "pn_smooth.cuh" /* PnSmoothParams */
#ifdef PN_SMOOTH_REG_UNDEF_DECL
#undef PN_SMOOTH_DECL_ONLY
#undef PN_SMOOTH_REG_UNDEF_DECL
#endif

#ifndef PN_SMR_THREADS
#define PN_SMR_THREADS 256
#endif

// Row of D doubles at p (16-byte aligned when D is even: record and row offsets are multiples of D doubles).
template <int D_>
PN_HD void pn_ld_row(const double* __restrict__ p, double (&dst)[D_], bool pred) {
    if (D_ % 2 == 0) {
#pragma unroll
        for (int c = 0; c < D_; c += 2) {
            double2 v = make_double2(0.0, 0.0);
            if (pred) v = *reinterpret_cast<const double2*>(p + c);
            dst[c] = v.x;
            dst[c + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int c = 0; c < D_; c++) dst[c] = pred ? p[c] : 0.0;
    }
}
// column c = (j, i) is scaled by scv[i] (per-component calibration; a scalar model passes d equal entries)
template <int D_, int d_>
PN_HD void pn_st_row(double* __restrict__ p, const double (&src)[D_], const double (&scv)[d_]) {
    if (D_ % 2 == 0) {
#pragma unroll
        for (int c = 0; c < D_; c += 2)
            *reinterpret_cast<double2*>(p + c) = make_double2(src[c] * scv[c % d_], src[c + 1] * scv[(c + 1) % d_]);
    } else {
#pragma unroll
        for (int c = 0; c < D_; c++) p[c] = src[c] * scv[c % d_];
    }
}

template <int d, int Q, int G>
struct PnSmoothReg {
    static constexpr int n = Q + 1;
    static constexpr int D = d * n;
    static constexpr int RPL = (D + G - 1) / G;
    static constexpr int ROWS = RPL * G;
    static constexpr int W = 2 * D;
    static constexpr int LD = D | 1; /* odd row stride (doubles): a group's own-row accesses hit distinct banks */
    /* per-trajectory shared-memory region (doubles) */
    static constexpr int OFF_RM = 0;                  /* R^-       D x LD (upper triangle used)          */
    static constexpr int OFF_YM = OFF_RM + D * LD;    /* Y         D x LD                                */
    static constexpr int OFF_SM = OFF_YM + D * LD;    /* R_s+      ROWS x LD  (own rows parked here)     */
    static constexpr int OFF_DI = OFF_SM + ROWS * LD; /* 1/diag(R^-)  D                                  */
    static constexpr int OFF_MS = OFF_DI + D;         /* mu_s+        D                                  */
    static constexpr int RAW = OFF_MS + D;
    /* region stride: the 16/G groups of a half-warp must broadcast-read from distinct banks */
    static constexpr int PHASE = (G >= 16) ? 0 : G;
    static constexpr int REGION = ((RAW - PHASE + 15) / 16) * 16 + PHASE;
    static constexpr int GROUPS = PN_SMR_THREADS / G;
    static constexpr int SMEM_BYTES = REGION * GROUPS * 8;
    static constexpr int JC = 8;

    static __device__ void run(const PnSmoothParams& P, double* smem_all) {
        constexpr unsigned FULL = 0xffffffffu;
        const int lane = threadIdx.x & 31;
        const int gl = lane % G;
        const int leader = lane - gl;
        double* sm = smem_all + (size_t)(threadIdx.x / G) * REGION;
        double* Rm = sm + OFF_RM;
        double* Ym = sm + OFF_YM;
        double* Sm = sm + OFF_SM;
        double* Di = sm + OFF_DI;
        double* Ms = sm + OFF_MS;

        int mB[RPL], iB[RPL];
#pragma unroll
        for (int lr = 0; lr < RPL; lr++) {
            const int rb = lr * G + gl;
            mB[lr] = rb / d;
            iB[lr] = rb - mB[lr] * d;
        }

        long long k = -1, a0 = 0;
        double scv[d]; /* sqrt of the global calibration scale, per component */
#pragma unroll
        for (int i = 0; i < d; i++) scv[i] = 1.0;
        bool have = false, queue_empty = false;

        for (;;) {
            /* ===== (A) groups without work fetch a trajectory; x_smooth[N] = x_filt[N] ===== */
            const bool need = !have && !queue_empty;
            if (__any_sync(FULL, need)) {
                unsigned long long idx = 0;
                if (need && gl == 0) idx = atomicAdd(P.work_counter, 1ULL);
                idx = __shfl_sync(FULL, idx, leader);
                bool got = false;
                long long last = 0;
                if (need) {
                    if ((long long)idx >= P.n_traj) {
                        queue_empty = true;
                    } else {
                        got = true;
                        a0 = P.step_offsets[idx];
                        last = P.counts ? a0 + P.counts[idx] : P.step_offsets[idx + 1] - 1;
#pragma unroll
                        for (int i = 0; i < d; i++)
                            scv[i] = (!P.dynamic && P.calibrate) ? sqrt((P.gdiff_mv ? P.gdiff_mv[idx * d + i] : P.gdiff[idx]) / P.initial_diffusion) : 1.0;
                    }
                }
                double Rl[RPL][D];
#pragma unroll
                for (int lr = 0; lr < RPL; lr++) {
                    const int row = lr * G + gl;
                    const bool ld = got && row < D;
                    pn_ld_row<D>(P.s_x_covfac + (last * D + row) * D, Rl[lr], ld);
                }
                double cv[d * d];
#pragma unroll
                for (int a = 0; a < d; a++)
#pragma unroll
                    for (int b = 0; b < d; b++) {
                        if (b < a) continue;
                        double s = 0.0;
#pragma unroll
                        for (int lr = 0; lr < RPL; lr++) s += Rl[lr][a] * Rl[lr][b];
                        s = pn_gsum<G>(s) * (scv[a] * scv[b]);
                        cv[a * d + b] = s;
                        cv[b * d + a] = s;
                    }
                if (got) {
#pragma unroll
                    for (int lr = 0; lr < RPL; lr++) {
                        const int row = lr * G + gl;
#pragma unroll
                        for (int c = 0; c < D; c++) Sm[row * LD + c] = Rl[lr][c];
                        if (!P.dynamic && P.calibrate && row < D) pn_st_row<D, d>(P.s_x_covfac + (last * D + row) * D, Rl[lr], scv);
                    }
                    if (gl == 0) {
#pragma unroll
                        for (int c = 0; c < D; c++) Ms[c] = P.s_x_mean[last * D + c];
#pragma unroll
                        for (int i = 0; i < d; i++) P.s_u_mean[last * d + i] = P.s_x_mean[last * D + i];
#pragma unroll
                        for (int i = 0; i < d * d; i++) P.s_u_cov[last * d * d + i] = cv[i];
                    }
                    k = last - 1;
                    have = (k >= a0);
                }
                __syncwarp();
            }
            if (__syncthreads_or(((have || !queue_empty) ? 1 : 0)) == 0) break;

            /* ===== (B) one backward step for every group that has one ===== */
            const bool active = have;
            const double h = active ? P.s_h[k] : 1.0;
            const bool copy = active && (h == 0.0); /* integrator_utils.jl:109-112 */
            const bool doit = active && !copy;
            if (__any_sync(FULL, copy)) {
                if (copy) {
#pragma unroll
                    for (int lr = 0; lr < RPL; lr++) {
                        const int row = lr * G + gl;
                        if (row < D) {
#pragma unroll
                            for (int c = 0; c < D; c++) P.s_x_covfac[(k * D + row) * D + c] = Sm[row * LD + c] * scv[c % d];
                        }
                    }
                    if (gl == 0) {
#pragma unroll
                        for (int c = 0; c < D; c++) P.s_x_mean[k * D + c] = Ms[c];
                    }
                }
            }
            if (__any_sync(FULL, doit)) {
                const double hh = doit ? h : 1.0;
                const double ediff = doit ? (P.dynamic ? P.s_diffusion[k] : P.initial_diffusion) : 1.0;
                double hp[n], PIv[n];
                hp[0] = 1.0;
#pragma unroll
                for (int j = 1; j < n; j++) hp[j] = hp[j - 1] * hh / (double)j;
                {
                    const double sqh = sqrt(hh);
#pragma unroll
                    for (int c = 0; c < n; c++) PIv[c] = hp[Q - c] * sqh;
                }
#if PN_IOUP
                double tq[n], Uh[n][n];
                pn_ioup_step<Q, G>(P.rate, P.glx, P.glw, hh, gl, tq, Uh);
#endif
                double Lf[RPL][D]; /* Lambda_f rows (bottom-right block after the sweep) */
                {
                    double Tt[RPL][W], Bt[RPL][W];
                    /* top block: [ R Ah' | R ] from the filtered record (each lane streams its own rows) */
#pragma unroll
                    for (int lr = 0; lr < RPL; lr++) {
                        const int row = lr * G + gl;
                            const bool ld = doit && row < D;
                        {
                            double rr[D];
                            pn_ld_row<D>(P.s_x_covfac + (k * D + row) * D, rr, ld);
#pragma unroll
                            for (int c = 0; c < D; c++) Tt[lr][D + c] = rr[c];
                            /* next record towards L2 while this one is processed */
#ifndef PN_HOST_EMULATION
                            if (ld && k > a0) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.s_x_covfac + ((k - 1) * D + row) * D));
#endif
                        }
#pragma unroll
                        for (int i = 0; i < d; i++)
#pragma unroll
                            for (int j = 0; j < n; j++) {
#if PN_IOUP
                                double s = (j == Q) ? tq[Q] * Tt[lr][D + j * d + i] : Tt[lr][D + j * d + i];
#pragma unroll
                                for (int kk = 0; kk < n; kk++)
                                    if (kk > j) s += ((kk == Q) ? tq[j] : hp[kk > j ? kk - j : 0]) * Tt[lr][D + kk * d + i];
#else
                                double s = Tt[lr][D + j * d + i];
#pragma unroll
                                for (int kk = 0; kk < n; kk++)
                                    if (kk > j) s += hp[kk > j ? kk - j : 0] * Tt[lr][D + kk * d + i];
#endif
                                Tt[lr][j * d + i] = s;
                            }
                    }
                    /* bottom block: [ sd (U P^-1) (x) I_d | 0 ], upper triangular */
                    {
                        const double sd0 = (ediff == 1.0) ? 1.0 : sqrt(ediff);
#pragma unroll
                        for (int lr = 0; lr < RPL; lr++) {
                            /* multivariate diffusion: row (m, i) of the noise factor carries sigma_i */
                            const double sd = (P.s_diffusion_mv && P.dynamic && doit) ? sqrt(P.s_diffusion_mv[k * d + iB[lr]]) : sd0;
#pragma unroll
                            for (int c = 0; c < n; c++) {
                                double usel = 0.0;
#pragma unroll
                                for (int m = 0; m < n; m++)
#if PN_IOUP
                                    if (m <= c) usel = (mB[lr] == m) ? Uh[m][c] : usel;
#else
                                    if (m <= c) usel = (mB[lr] == m) ? P.U[m * n + c] : usel;
#endif
                                const double val = usel * PIv[c] * sd;
#pragma unroll
                                for (int i = 0; i < d; i++) Bt[lr][c * d + i] = (iB[lr] == i) ? val : 0.0;
                            }
#pragma unroll
                            for (int c = 0; c < D; c++) Bt[lr][D + c] = 0.0;
                        }
                    }
                    pn_qr_rows<G, RPL, W, D, true, JC>(Tt, Bt, gl, Di);
                    /* publish R^- and Y for the group */
#pragma unroll
                    for (int lr = 0; lr < RPL; lr++) {
                        const int row = lr * G + gl;
                        if (row < D) {
#pragma unroll
                            for (int c = 0; c < D; c++) {
                                if (c >= lr * G) Rm[row * LD + c] = Tt[lr][c];
                                Ym[row * LD + c] = Tt[lr][D + c];
                            }
                        }
#pragma unroll
                        for (int c = 0; c < D; c++) Lf[lr][c] = Bt[lr][D + c];
                    }
                }
                __syncwarp();
                /* Z = R_s+ R^-^-1 (own rows) and w = R^-^-T (mu_s+ - Ah mu) (replicated) */
                double Z[RPL][D], w[D], mu[D];
#pragma unroll
                for (int c = 0; c < D; c++) mu[c] = doit ? P.s_x_mean[k * D + c] : 0.0;
#pragma unroll
                for (int j = 0; j < n; j++)
#pragma unroll
                    for (int i = 0; i < d; i++) {
#if PN_IOUP
                        double s = (j == Q) ? tq[Q] * mu[j * d + i] : mu[j * d + i];
#pragma unroll
                        for (int kk = 0; kk < n; kk++)
                            if (kk > j) s += ((kk == Q) ? tq[j] : hp[kk > j ? kk - j : 0]) * mu[kk * d + i];
#else
                        double s = mu[j * d + i];
#pragma unroll
                        for (int kk = 0; kk < n; kk++)
                            if (kk > j) s += hp[kk > j ? kk - j : 0] * mu[kk * d + i];
#endif
                        w[j * d + i] = Ms[j * d + i] - s;
                    }
#pragma unroll
                for (int lr = 0; lr < RPL; lr++) {
                    const int row = lr * G + gl;
#pragma unroll
                    for (int c = 0; c < D; c++) Z[lr][c] = Sm[row * LD + c];
                }
#pragma unroll
                for (int j = 0; j < D; j++) {
                    const double dj = Di[j];
                    double rcol[D];
#pragma unroll
                    for (int kk = 0; kk < D; kk++)
                        if (kk < j) rcol[kk] = Rm[kk * LD + j];
#pragma unroll
                    for (int lr = 0; lr < RPL; lr++) {
                        double s = Z[lr][j];
#pragma unroll
                        for (int kk = 0; kk < D; kk++)
                            if (kk < j) s -= Z[lr][kk] * rcol[kk];
                        Z[lr][j] = s * dj;
                    }
                    double s = w[j];
#pragma unroll
                    for (int kk = 0; kk < D; kk++)
                        if (kk < j) s -= w[kk] * rcol[kk];
                    w[j] = s * dj;
                }
                /* top of the marginalisation stack: R_s+ G' = Z Y ; smoothed mean mu + Y' w */
                double St[RPL][D];
#pragma unroll
                for (int lr = 0; lr < RPL; lr++)
#pragma unroll
                    for (int c = 0; c < D; c++) St[lr][c] = 0.0;
#pragma unroll
                for (int kk = 0; kk < D; kk++) {
                    double yrow[D];
#pragma unroll
                    for (int c = 0; c < D; c++) yrow[c] = Ym[kk * LD + c];
#pragma unroll
                    for (int lr = 0; lr < RPL; lr++)
#pragma unroll
                        for (int c = 0; c < D; c++) St[lr][c] += Z[lr][kk] * yrow[c];
#pragma unroll
                    for (int c = 0; c < D; c++) mu[c] += w[kk] * yrow[c];
                }
                __syncwarp(); /* everyone is done reading Rm / Ym / Sm / Ms */
                pn_qr_rows<G, RPL, D, D, false, JC>(St, Lf, gl, nullptr);
                /* outputs: x_smooth[k] over x_filt[k], pu[k] = SolProj x_smooth[k] */
                double cv[d * d];
#pragma unroll
                for (int a = 0; a < d; a++)
#pragma unroll
                    for (int b = 0; b < d; b++) {
                        if (b < a) continue;
                        double s = 0.0;
#pragma unroll
                        for (int lr = 0; lr < RPL; lr++) s += St[lr][a] * St[lr][b];
                        s = pn_gsum<G>(s) * (scv[a] * scv[b]);
                        cv[a * d + b] = s;
                        cv[b * d + a] = s;
                    }
                if (doit) {
#pragma unroll
                    for (int lr = 0; lr < RPL; lr++) {
                        const int row = lr * G + gl;
#pragma unroll
                        for (int c = 0; c < D; c++) Sm[row * LD + c] = St[lr][c];
                        if (row < D) pn_st_row<D, d>(P.s_x_covfac + (k * D + row) * D, St[lr], scv);
                    }
                    if (gl == 0) {
#pragma unroll
                        for (int c = 0; c < D; c++) {
                            Ms[c] = mu[c];
                            P.s_x_mean[k * D + c] = mu[c];
                        }
#pragma unroll
                        for (int i = 0; i < d; i++) P.s_u_mean[k * d + i] = mu[i];
#pragma unroll
                        for (int i = 0; i < d * d; i++) P.s_u_cov[k * d * d + i] = cv[i];
                    }
                }
                __syncwarp();
            }
            if (active) {
                k -= 1;
                have = (k >= a0);
            }
        }
    }
};

template <int d, int Q, int G>
__device__ __forceinline__ void pn_smooth_reg_entry(const PnSmoothParams& P) {
    extern __shared__ double pn_smr_shared[];
    PnSmoothReg<d, Q, G>::run(P, pn_smr_shared);
}

#endif /* PN_SMOOTH_REG_CUH */
