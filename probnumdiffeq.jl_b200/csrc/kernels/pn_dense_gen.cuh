// pn_dense_gen.cuh -- dense output `sol(t)` for states beyond the register kernels of pn_dense.cuh (D = d(q+1) > 32: the CTA-per-
// trajectory path, e.g. Pleiades D = 112 / 56): the same first half of `interpolate` (src/solution.jl:330-371),
//   goal_pred = predict(x_filt[idx], Ah(h1), diffusion * Qh(h1)),
// one WARP per (trajectory, saveat point) pair, matrices in a per-warp workspace (shared memory, or an L2-resident global buffer when
// 2 D^2 doubles per warp do not fit), size-generic at run time like the shared-memory smoother sweep whose Householder routine it
// uses.  The second half (smooth over h2) is one backward step of that sweep on the two-point virtual trajectories, as in pn_dense.cuh.
// IWP prior, scalar diffusion models, first-order problems.
#ifndef PN_DENSE_GEN_CUH
#define PN_DENSE_GEN_CUH
#include "pn_dense.cuh"
#include "pn_smooth.cuh" /* pn_sm_qr, PN_SM_WARPS */

/* per-warp workspace, doubles (host and kernel must agree) */
#define PN_DENSE_GEN_WS(D, n) (2 * (D) * (D) + 2 * (D) + 2 * (n) * (n) + 2 * (n))

#ifndef PN_SMOOTH_DECL_ONLY
__global__ void __launch_bounds__(32 * PN_SM_WARPS) pn_dense_gen_kernel(const __grid_constant__ PnDenseParams P, int d, int q, double* workspace) {
    extern __shared__ double pn_dg_shared[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = q + 1, D = d * n, DD = D * D;
    const int per_warp = PN_DENSE_GEN_WS(D, n);
    double* W = workspace ? workspace + ((size_t)blockIdx.x * PN_SM_WARPS + warp) * per_warp : pn_dg_shared + (size_t)warp * per_warp;
    double* Xs = W;            /* 2D x D stack [R Ah'; sqrt(s2) (U P^-1) (x) I] */
    double* mu = Xs + 2 * DD;  /* D */
    double* mup = mu + D;      /* D */
    double* T1 = mup + D;      /* n x n  T[i][j] = h^(j-i)/(j-i)! */
    double* UQ = T1 + n * n;   /* n x n  U P^-1 (upper) */
    double* hp = UQ + n * n;   /* n */
    double* PIv = hp + n;      /* n */
    const long long V = P.n_traj * P.n_saveat;
    const long long warps_total = (long long)gridDim.x * PN_SM_WARPS;
    for (long long v = (long long)blockIdx.x * PN_SM_WARPS + warp; v < V; v += warps_total) {
        const long long tr = v / P.n_saveat, j = v - tr * P.n_saveat;
        const double tval = P.saveat[j];
        const long long a0 = P.step_offsets[tr], last = P.counts ? a0 + P.counts[tr] : P.step_offsets[tr + 1] - 1;
        long long lo = a0, hi = last + 1;
        while (hi - lo > 1) {
            const long long mid = (lo + hi) >> 1;
            if (P.s_t[mid] <= tval) lo = mid; else hi = mid;
        }
        const long long rec = lo;
        const double h1 = tval - P.s_t[rec];
        const double ediff = (last > a0) ? P.s_diffusion[(rec < last) ? rec : last - 1] : 0.0;
        const double h2 = (rec < last) ? P.s_t[rec + 1] - tval : 0.0;
        const double g = P.calibrate_scale ? P.gdiff[tr] / P.initial_diffusion : 1.0;
        __syncwarp();
        if (lane == 0) {
            hp[0] = 1.0;
            for (int k = 1; k < n; k++) hp[k] = hp[k - 1] * h1 / (double)k;
            const double sqh = sqrt(h1);
            for (int c = 0; c < n; c++) PIv[c] = hp[q - c] * sqh;
        }
        for (int c = lane; c < D; c += 32) mu[c] = P.s_x_mean[rec * D + c];
        __syncwarp();
        for (int i = lane; i < n * n; i += 32) {
            const int r = i / n, c = i % n;
            T1[i] = (c >= r) ? hp[c - r] : 0.0;
            UQ[i] = (c >= r) ? P.U[r * n + c] * PIv[c] : 0.0;
        }
        __syncwarp();
        const double sd = (ediff == 1.0) ? 1.0 : sqrt(ediff);
        const double* Rk = P.s_x_covfac + rec * (long long)DD;
        for (int i = lane; i < DD; i += 32) {
            const int r = i / D, c = i % D, jj = c / d, ii = c % d;
            double s = 0.0;
            for (int k = jj; k < n; k++) s += T1[jj * n + k] * Rk[r * D + k * d + ii];
            Xs[i] = s;
            const int m = r / d, i2 = r % d;
            Xs[DD + i] = (i2 == ii) ? sd * UQ[m * n + jj] : 0.0;
        }
        for (int i = lane; i < D; i += 32) {
            const int jj = i / d, ii = i % d;
            double s = 0.0;
            for (int k = jj; k < n; k++) s += T1[jj * n + k] * mu[k * d + ii];
            mup[i] = s;
        }
        __syncwarp();
        pn_sm_qr(Xs, 2 * D, D, lane);
        if (P.smooth) {
            if (lane == 0) {
                P.v_idx[v] = rec;
                P.v_h[2 * v] = h2;
                P.v_h[2 * v + 1] = 0.0;
                P.v_diffusion[2 * v] = ediff * g;
                P.v_diffusion[2 * v + 1] = 0.0;
            }
            const double sg = sqrt(g);
            for (int c = lane; c < D; c += 32) P.v_mean[2 * v * D + c] = mup[c];
            for (int i = lane; i < DD; i += 32) P.v_covfac[2 * v * (long long)DD + i] = Xs[i] * sg;
        } else {
            for (int i = lane; i < d * d; i += 32) {
                const int a = i / d, b = i % d;
                double s = 0.0;
                for (int r = 0; r <= (a < b ? a : b); r++) s += Xs[r * D + a] * Xs[r * D + b]; /* upper triangular factor */
                P.out_cov[v * d * d + i] = s * g;
            }
            for (int i = lane; i < d; i += 32) P.out_u[v * d + i] = mup[i];
        }
        __syncwarp();
    }
}
#endif

#endif /* PN_DENSE_GEN_CUH */
