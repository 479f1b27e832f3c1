// pn_dense.cuh -- dense output `sol(t)` on a common `saveat` grid, evaluated on the device from the saved step records.
//
// Reference behaviour replaced: `interpolate` (src/solution.jl:330-398), called by `sol(t)` / `sol(ts)`:
//   idx = last saved point with t[idx] <= tval;  h1 = tval - t[idx];  h2 = t[idx+1] - tval
//   goal_pred     = predict(x_filt[idx], Ah(h1), diffusion * Qh(h1))                       (:355-371)
//   goal_smoothed = smooth(goal_pred, x_smooth[idx+1], Ah(h2), diffusion * Qh(h2))         (:373-397, smoothed solutions)
// with diffusion = diffusions[min(idx, end)] (:351).  (The reference evaluates both in preconditioned coordinates; the
// formulas are the same.)
//
// Here: pn_dense_predict_kernel does the first half for every (trajectory, saveat point) pair -- G lanes per pair, rows
// in registers, the same upper-triangular-noise Householder sweep as the step kernel -- and writes the result as
// record 0 of a two-point *virtual trajectory*; after the smoother sweep of the real trajectories, pn_dense_gather
// appends x_smooth[idx+1] as record 1, and the SAME sweep kernel (pn_smooth_col / pn_smooth_reg) is launched over the
// virtual trajectories: one backward step of a sweep is exactly `smooth` (compute_backward_kernel! + marginalize!).
// pn_dense_collect then picks the solution marginals of the records 0.
//
// NVRTC-compatible: no standard headers.
#ifndef PN_DENSE_CUH
#define PN_DENSE_CUH
#include "pn_params.h"
#include "pn_small.cuh"

struct PnDenseParams {
    long long n_traj, n_saveat;
    const double* saveat;          /* [n_saveat] ascending, inside [t0, t1] */
    int smooth;                    /* 1: fill the virtual trajectories; 0: write the filtering marginals directly */
    int calibrate_scale;           /* FixedDiffusion(calibrate=true): marginals / records are scaled by gdiff */
    const long long* step_offsets; /* [n_traj+1] */
    const long long* counts;       /* [n_traj] or null (see PnSmoothParams) */
    const double* s_t;             /* [total] */
    const double* s_h;             /* [total] step k -> k+1 at slot k */
    const double* s_diffusion;     /* [total] diffusion used for step k -> k+1 at slot k */
    const double* s_x_mean;        /* [total][D] filtered */
    const double* s_x_covfac;      /* [total][D][D] filtered */
    const double* gdiff;           /* [n_traj] sigma_hat^2 * initial_diffusion */
    double initial_diffusion;      /* the covariance scale of calibrate_solution! is gdiff / initial_diffusion */
    double U[PN_MAX_N * PN_MAX_N];
    double rate;
    double glx[16], glw[16];
    /* virtual trajectories (two records per pair) */
    long long* v_idx;    /* [V] global record index of x_filt[idx] */
    double* v_mean;      /* [2V][D]    */
    double* v_covfac;    /* [2V][D][D] */
    double* v_h;         /* [2V] slot 2v: h2 */
    double* v_diffusion; /* [2V] slot 2v */
    /* multivariate diffusion models (DynamicMVDiffusion / FixedMVDiffusion on the blocks layout): per-component diffusions of the saved
       steps, per-component calibration scale, per-component diffusion of the virtual steps; null for the scalar models */
    const double* s_diffusion_mv; /* [total][d] */
    const double* gdiff_mv;       /* [n_traj][d] */
    double* v_diffusion_mv;       /* [2V][d] slot 2v */
    /* direct outputs (smooth == 0) */
    double* out_u;       /* [V][d]    */
    double* out_cov;     /* [V][d][d] */
};

#ifndef PN_DENSE_DECL_ONLY

template <int d, int Q, int G>
struct PnDensePredict {
    static constexpr int n = Q + 1;
    static constexpr int D = d * n;
    static constexpr int RPL = (D + G - 1) / G;

    static __device__ void run(const PnDenseParams& P) {
        const int lane = threadIdx.x & 31;
        const int gl = lane % G;
        const long long V = P.n_traj * P.n_saveat;
        const long long groups_total = (long long)gridDim.x * (blockDim.x / G);
        const long long g0 = (long long)blockIdx.x * (blockDim.x / G) + threadIdx.x / G;
        int mB[RPL], iB[RPL];
#pragma unroll
        for (int lr = 0; lr < RPL; lr++) {
            const int rb = lr * G + gl;
            mB[lr] = rb / d;
            iB[lr] = rb - mB[lr] * d;
        }
        /* warp-convergent grid-stride loop: every lane of the warp runs the same number of iterations */
        const long long iters = (V + groups_total - 1) / groups_total;
        for (long long it = 0; it < iters; it++) {
            const long long v = g0 + it * groups_total;
            const bool act = v < V;
            const long long tr = act ? v / P.n_saveat : 0;
            const long long j = act ? v - tr * P.n_saveat : 0;
            const double tval = P.saveat[j];
            const long long a0 = P.step_offsets[tr], last = P.counts ? a0 + P.counts[tr] : P.step_offsets[tr + 1] - 1;
            /* idx: last saved point with t <= tval (binary search; t[a0] = t0 <= tval) */
            long long lo = a0, hi = last + 1;
            while (hi - lo > 1) {
                const long long mid = (lo + hi) >> 1;
                if (P.s_t[mid] <= tval) lo = mid; else hi = mid;
            }
            const long long rec = lo;
            const double h1 = act ? tval - P.s_t[rec] : 0.0;
            const long long dslot = (rec < last) ? rec : last - 1;
            const double ediff = (last > a0) ? P.s_diffusion[dslot] : 0.0;
            double edc[d], gc[d]; /* per-component diffusion and calibration scale (all equal for the scalar models) */
#pragma unroll
            for (int i = 0; i < d; i++) {
                edc[i] = (P.s_diffusion_mv && last > a0) ? P.s_diffusion_mv[dslot * d + i] : ediff;
                gc[i] = (P.calibrate_scale && act) ? (P.gdiff_mv ? P.gdiff_mv[tr * d + i] : P.gdiff[tr]) / P.initial_diffusion : 1.0;
            }
            const double h2 = (rec < last) ? P.s_t[rec + 1] - tval : 0.0;
            double hp[n], PIv[n];
            hp[0] = 1.0;
#pragma unroll
            for (int k = 1; k < n; k++) hp[k] = hp[k - 1] * h1 / (double)k;
            {
                const double sqh = sqrt(h1);
#pragma unroll
                for (int c = 0; c < n; c++) PIv[c] = hp[Q - c] * sqh;
            }
#if PN_IOUP
            double tq[n], Uh[n][n];
            if (h1 > 0.0) {
                pn_ioup_step<Q, G>(P.rate, P.glx, P.glw, h1, gl, tq, Uh);
            } else {
#pragma unroll
                for (int k = 0; k < n; k++) {
                    tq[k] = (k == Q) ? 1.0 : 0.0;
#pragma unroll
                    for (int c = 0; c < n; c++) Uh[k][c] = 0.0;
                }
            }
#endif
            double mu[D], mup[D];
#pragma unroll
            for (int c = 0; c < D; c++) mu[c] = act ? P.s_x_mean[rec * D + c] : 0.0;
#pragma unroll
            for (int jj = 0; jj < n; jj++)
#pragma unroll
                for (int i = 0; i < d; i++) {
#if PN_IOUP
                    double s = (jj == Q) ? tq[Q] * mu[jj * d + i] : mu[jj * d + i];
#pragma unroll
                    for (int k = 0; k < n; k++)
                        if (k > jj) s += ((k == Q) ? tq[jj] : hp[k > jj ? k - jj : 0]) * mu[k * d + i];
#else
                    double s = mu[jj * d + i];
#pragma unroll
                    for (int k = 0; k < n; k++)
                        if (k > jj) s += hp[k > jj ? k - jj : 0] * mu[k * d + i];
#endif
                    mup[jj * d + i] = s;
                }
            double R[RPL][D], B[RPL][D];
#pragma unroll
            for (int lr = 0; lr < RPL; lr++) {
                const int row = lr * G + gl;
                const bool ld = act && row < D;
#pragma unroll
                for (int c = 0; c < D; c++) R[lr][c] = ld ? P.s_x_covfac[(rec * D + row) * D + c] : 0.0;
#pragma unroll
                for (int i = 0; i < d; i++)
#pragma unroll
                    for (int jj = 0; jj < n; jj++) {
#if PN_IOUP
                        double s = (jj == Q) ? tq[Q] * R[lr][jj * d + i] : R[lr][jj * d + i];
#pragma unroll
                        for (int k = 0; k < n; k++)
                            if (k > jj) s += ((k == Q) ? tq[jj] : hp[k > jj ? k - jj : 0]) * R[lr][k * d + i];
#else
                        double s = R[lr][jj * d + i];
#pragma unroll
                        for (int k = 0; k < n; k++)
                            if (k > jj) s += hp[k > jj ? k - jj : 0] * R[lr][k * d + i];
#endif
                        R[lr][jj * d + i] = s;
                    }
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
                    double sdr = 0.0; /* sqrt of the diffusion of the row's component */
#pragma unroll
                    for (int i = 0; i < d; i++) sdr = (iB[lr] == i) ? ((edc[i] == 1.0) ? 1.0 : sqrt(edc[i])) : sdr;
                    const double val = usel * PIv[c] * sdr;
#pragma unroll
                    for (int i = 0; i < d; i++) B[lr][c * d + i] = (iB[lr] == i) ? val : 0.0;
                }
            }
            pn_qr_rows<G, RPL, D, D, true, D>(R, B, gl, nullptr);
            if (P.smooth) {
                if (act) {
                    if (gl == 0) {
                        P.v_idx[v] = rec;
                        P.v_h[2 * v] = h2;
                        P.v_h[2 * v + 1] = 0.0;
                        P.v_diffusion[2 * v] = ediff * gc[0];
                        P.v_diffusion[2 * v + 1] = 0.0;
                        if (P.v_diffusion_mv) {
#pragma unroll
                            for (int i = 0; i < d; i++) {
                                P.v_diffusion_mv[2 * v * d + i] = edc[i] * gc[i];
                                P.v_diffusion_mv[(2 * v + 1) * d + i] = 0.0;
                            }
                        }
#pragma unroll
                        for (int c = 0; c < D; c++) P.v_mean[2 * v * D + c] = mup[c];
                    }
                    double sg[d];
#pragma unroll
                    for (int i = 0; i < d; i++) sg[i] = sqrt(gc[i]);
#pragma unroll
                    for (int lr = 0; lr < RPL; lr++) {
                        const int row = lr * G + gl;
                        if (row < D) {
#pragma unroll
                            for (int c = 0; c < D; c++) P.v_covfac[(2 * v * D + row) * D + c] = R[lr][c] * sg[c % d];
                        }
                    }
                }
            } else {
                constexpr int DU = PN_SECOND ? 2 * d : d; /* sol(t) = SolProj x: (E1; E0) rows for second-order problems */
                double cv[DU * DU];
#pragma unroll
                for (int a = 0; a < DU; a++)
#pragma unroll
                    for (int b = 0; b < DU; b++) {
                        if (b < a) continue;
                        double s = 0.0;
#pragma unroll
                        for (int lr = 0; lr < RPL; lr++) s += R[lr][pn_solcol<d>(a)] * R[lr][pn_solcol<d>(b)];
                        s = pn_gsum<G>(s) * sqrt(gc[pn_solcol<d>(a) % d] * gc[pn_solcol<d>(b) % d]);
                        cv[a * DU + b] = s;
                        cv[b * DU + a] = s;
                    }
                if (act && gl == 0) {
#pragma unroll
                    for (int i = 0; i < DU; i++) P.out_u[v * DU + i] = mup[pn_solcol<d>(i)];
#pragma unroll
                    for (int i = 0; i < DU * DU; i++) P.out_cov[v * DU * DU + i] = cv[i];
                }
            }
        }
    }
};

template <int d, int Q, int G>
__device__ __forceinline__ void pn_dense_predict_entry(const PnDenseParams& P) {
    PnDensePredict<d, Q, G>::run(P);
}

#endif /* PN_DENSE_DECL_ONLY */
#endif /* PN_DENSE_CUH */
