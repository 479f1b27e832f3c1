// pn_smooth_col.cuh -- RTS smoother sweep for small filter states (D = d(q+1) <= 24), register-resident,
// COLUMN-distributed: a group of G lanes owns one trajectory and lane gl holds columns c == gl (mod G) of the
// 2D x 2D sweep matrix.  A Householder reflector is then *broadcast* (one shuffle per entry of v) and every dot
// product v'a_j is lane-local, instead of the row-distributed scheme of pn_small.cuh / pn_smooth_reg.cuh where every
// one of the ~D^2 dot products per sweep needs a log2(G)-level butterfly.  At G = 8 that is 4x fewer shuffles and
// 40 % fewer FP64 instructions per step.
//
// Reference behaviour replaced (per trajectory, i = N-1 ... 1):
//   compute_backward_kernel!   src/filtering/markov_kernel.jl:190-235   G, b, Lambda.R (2D x D)
//   marginalize!               src/filtering/markov_kernel.jl:68-101    mu_s = G mu_s+ + b ; R_s = qr([R_s+ G' ; Lambda.R]).R
//   smooth_solution!           src/integrator_utils.jl:98-121           sweep, pu = SolProj x_smooth, dt == 0 copies
//   calibrate_solution!        src/integrator_utils.jl:46-65            sqrt(sigma_hat^2) scaling (FixedDiffusion)
//
// The algebra is the one of pn_smooth_reg.cuh: the first D reflections of [X R; B 0] (X = R Ah', B the upper-
// triangular noise factor) give [R^- Y; 0 Lambda_f]; G' = R^-^-1 Y column by column (lane-local back substitution
// against broadcast reads of R^-), R_s+ G' likewise against broadcast reads of R_s+, mu_s = mu + G (mu_s+ - Ah mu), and
// R_s = qr([R_s+ G'; Lambda_f]).  Per step the filtered record is staged global -> shared with coalesced 16-byte
// copies and the smoothed record leaves the same way.
//
// NVRTC-compatible: no standard headers.
#ifndef PN_SMOOTH_COL_CUH
#define PN_SMOOTH_COL_CUH
#include "pn_params.h"
#include "pn_small.cuh" /* pn_rcp, pn_rsqrt, PN_IOUP */
#ifndef PN_SMOOTH_DECL_ONLY
#define PN_SMOOTH_DECL_ONLY
#define PN_SMOOTH_COL_UNDEF_DECL
#endif
#include "pn_smooth.cuh" /* PnSmoothParams */
#ifdef PN_SMOOTH_COL_UNDEF_DECL
#undef PN_SMOOTH_DECL_ONLY
#undef PN_SMOOTH_COL_UNDEF_DECL
#endif

#ifndef PN_SMR_THREADS
#define PN_SMR_THREADS 256
#endif

// Householder sweep on a column-distributed matrix A[NROW][NS] (slot s of lane gl holds column s*G + gl).  The top
// block has TOPR rows, the bottom block NROW - TOPR.  Columns C0 .. C0+NE-1 are eliminated; the pivot of column C0+kk is
// top row kk.  BOTUP: bottom row r joins at reflector r (upper-triangular bottom block).  dinv (shared memory,
// optional): lane 0 of the group stores 1/diag.  The pivot entry of the eliminated column is replaced by the diagonal.
template <int G, int NROW, int NS, int TOPR, int C0, int NE, bool BOTUP>
PN_HD void pn_qr_cols(double (&A)[NROW][NS], int gl, double* dinv) {
#pragma unroll
    for (int kk = 0; kk < NE; kk++) {
        const int col = C0 + kk;
        const int ks = col / G, kl = col % G;
        double v[NROW];
        double total = 0.0;
#pragma unroll
        for (int r = 0; r < NROW; r++) {
            const bool act = (r < TOPR) ? (r >= kk) : (!BOTUP || (r - TOPR) <= kk);
            if (act) {
                v[r] = __shfl_sync(0xffffffffu, A[r][ks], kl, G);
                total += v[r] * v[r];
            } else {
                v[r] = 0.0;
            }
        }
        const double alpha = v[kk];
        const bool nz = total > 0.0;
        const double rsq = pn_rsqrt(total);
        const double nrm = nz ? total * rsq : 0.0;
        const double beta = -copysign(nrm, alpha);
        const double gam = nz ? pn_rcp(total + fabs(alpha) * nrm) : 0.0;
        v[kk] = alpha - beta;
        if (dinv != nullptr && gl == 0) dinv[kk] = nz ? -copysign(rsq, alpha) : 0.0;
#pragma unroll
        for (int s = 0; s < NS; s++) {
            if (s < ks) continue;
            const bool pred = (s > ks) || (gl > kl);
            double dot = 0.0;
#pragma unroll
            for (int r = 0; r < NROW; r++) {
                const bool act = (r < TOPR) ? (r >= kk) : (!BOTUP || (r - TOPR) <= kk);
                if (act) dot += v[r] * A[r][s];
            }
            const double f = pred ? dot * gam : 0.0;
#pragma unroll
            for (int r = 0; r < NROW; r++) {
                const bool act = (r < TOPR) ? (r >= kk) : (!BOTUP || (r - TOPR) <= kk);
                if (act) A[r][s] -= f * v[r];
            }
        }
        if (gl == kl) A[kk][ks] = beta;
    }
}

#ifndef PN_FENRIR
#define PN_FENRIR 0 /* 1: build the sweep with the Fenrir data updates (NVRTC builds of solvers that carry a data set) */
#endif
// Fenrir data update on the posterior (Ms, Sm) of one trajectory in shared memory: executed by ONE lane of the group (data
// points are rare compared with steps), work arrays in a local-memory frame that only this cold path touches.
#include "pn_fenrir_update.cuh"

template <int d, int Q, int G>
struct PnSmoothCol {
    static constexpr int n = Q + 1;
    static constexpr int D = d * n;
    static constexpr int W = 2 * D;
    static constexpr int NS = (W + G - 1) / G; /* column slots per lane */
    static constexpr int SR0 = D / G;          /* first slot that can hold a right-block column */
    static constexpr int LD = D | 1;
    /* per-trajectory shared-memory region (doubles) */
    static constexpr int OFF_REC = 0;                    /* staged record: R (D x D row-major), mean (D)   */
    static constexpr int RECSZ = D * D + D + ((D * D + D) & 1);
    static constexpr bool ASYNC = (D % 2 == 0);          /* 16-byte cp.async staging needs even D */
    static constexpr int OFF_RM = OFF_REC + (ASYNC ? 2 : 1) * RECSZ; /* R^- (strict upper part), D x LD; staging is double-buffered */
    static constexpr int OFF_SM = OFF_RM + D * LD;       /* R_s+       D x LD                              */
    static constexpr int OFF_DI = OFF_SM + D * LD;       /* 1/diag(R^-) D                                  */
    static constexpr int OFF_MS = OFF_DI + D;            /* mu_s+       D                                  */
    static constexpr int RAW = OFF_MS + D;
    static constexpr int PHASE = (G >= 16) ? 0 : G;
    static constexpr int REGION = ((RAW - PHASE + 15) / 16) * 16 + PHASE;
    static constexpr int GROUPS = PN_SMR_THREADS / G;
    static constexpr int SMEM_BYTES = REGION * GROUPS * 8;

    /* u-block of Sigma = S'S from the shared-memory factor (all D rows: the factor of the last point is dense) */
    static PN_HD void write_pu(const PnSmoothParams& P, long long k, const double* Sm, const double* Ms, const double (&scv)[d]) {
        /* sol.pu = SolProj x: E0 rows, or (E1; E0) rows for second-order problems (PN_SECOND, pn_small.cuh) */
        constexpr int DU = PN_SECOND ? 2 * d : d;
#pragma unroll
        for (int a = 0; a < DU; a++) {
            const int ca = pn_solcol<d>(a);
            P.s_u_mean[k * DU + a] = Ms[ca];
#pragma unroll
            for (int b = 0; b < DU; b++) {
                if (b < a) continue;
                const int cb = pn_solcol<d>(b);
                double s = 0.0;
#pragma unroll
                for (int r = 0; r < D; r++) s += Sm[r * LD + ca] * Sm[r * LD + cb];
                s *= scv[ca % d] * scv[cb % d];
                P.s_u_cov[k * DU * DU + a * DU + b] = s;
                P.s_u_cov[k * DU * DU + b * DU + a] = s;
            }
        }
    }
    /* smoothed record (factor scaled per component, mean) shared -> global, G lanes cooperating, coalesced */
    static PN_HD void write_record(const PnSmoothParams& P, long long k, const double* Sm, const double* Ms, const double (&scv)[d],
                                   int gl) {
        for (int e = gl; e < D * D; e += G) {
            const int r = e / D, c = e - r * D;
            double sc = scv[0];
#pragma unroll
            for (int i = 1; i < d; i++) sc = (c % d == i) ? scv[i] : sc;
            P.s_x_covfac[k * D * D + e] = Sm[r * LD + c] * sc;
        }
        for (int c = gl; c < D; c += G) P.s_x_mean[k * D + c] = Ms[c];
    }

    /* global -> shared copy of record k (factor + mean), 16 bytes per cp.async, G lanes cooperating */
    static PN_HD void stage_async(const PnSmoothParams& P, long long k, double* buf, int gl) {
        const double* src = P.s_x_covfac + k * D * D;
        const double* srm = P.s_x_mean + k * D;
#ifdef PN_HOST_EMULATION /* tests/emulation: the copy lands at once -- the earliest moment the hardware allows */
        for (int e = 2 * gl; e < D * D; e += 2 * G) buf[e] = src[e], buf[e + 1] = src[e + 1];
        for (int c = 2 * gl; c < D; c += 2 * G) buf[D * D + c] = srm[c], buf[D * D + c + 1] = srm[c + 1];
#else
        for (int e = 2 * gl; e < D * D; e += 2 * G) {
            const unsigned sa = (unsigned)__cvta_generic_to_shared(buf + e);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src + e) : "memory");
        }
        for (int c = 2 * gl; c < D; c += 2 * G) {
            const unsigned sa = (unsigned)__cvta_generic_to_shared(buf + D * D + c);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(srm + c) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
#endif
    }

    static __device__ void run(const PnSmoothParams& P, double* smem_all) {
        constexpr unsigned FULL = 0xffffffffu;
        const int lane = threadIdx.x & 31;
        const int gl = lane % G;
        const int leader = lane - gl;
        double* sm = smem_all + (size_t)(threadIdx.x / G) * REGION;
        double* Rec = sm + OFF_REC;
        int cur = 0; /* staging buffer that holds (or is receiving) the record of the next step */
        double* Rm = sm + OFF_RM;
        double* Sm = sm + OFF_SM;
        double* Di = sm + OFF_DI;
        double* Ms = sm + OFF_MS;

        long long k = -1, a0 = 0;
        /* Fenrir mode (PnSmoothParams::fenrir_ll): per-trajectory data cursor and log-likelihood, held by lane 0 of the group */
        const bool fen = PN_FENRIR && P.fenrir_ll != nullptr; /* compile-time off in the plain smoother builds */
        constexpr int OMAX = PN_SECOND ? 2 * d : d;
        long long tr_cur = 0, data_idx = -1;
        double LL = 0.0;
        double scv[d];
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
                        for (int e = gl; e < D * D; e += G) {
                            const int r = e / D, c = e - r * D;
                            Sm[r * LD + c] = P.s_x_covfac[last * D * D + e];
                        }
                        for (int c = gl; c < D; c += G) Ms[c] = P.s_x_mean[last * D + c];
                    }
                }
                __syncwarp();
                if (got && fen) {
                    tr_cur = (long long)idx;
                    data_idx = P.n_data - 1;
                    LL = 0.0;
                    if (P.s_t[last] == P.data_t[data_idx]) { /* fenrir.jl:87-97 */
                        if (gl == 0)
                            LL += pn_fenrir_update_serial<D, LD, OMAX>(P.obs_H, P.obs_Rn, P.data_u + data_idx * P.n_obs, P.n_obs, Sm, Ms,
                                                                       P.kron_logdet != 0);
                    }
                    data_idx--; /* unconditional, as fenrir.jl:99 (`data_idx = length(data.u) - 1`): a last observation that is
                                   not at sol.t[end] is skipped by the reference, and so it is here */
                }
                if (got) {
                    if (!P.dynamic && P.calibrate && !fen && P.write_x) write_record(P, last, Sm, Ms, scv, gl);
                    if (gl == 0 && !fen) write_pu(P, last, Sm, Ms, scv);
                    k = last - 1;
                    have = (k >= a0);
                    if (fen && !have && gl == 0) /* a single saved point: nothing to sweep */
                        P.fenrir_ll[tr_cur] = ((!P.retcode || P.retcode[tr_cur] == PN_RET_SUCCESS) && data_idx < 0) ? LL : -PN_INF;
                    if (ASYNC && have && P.s_h[k] != 0.0) stage_async(P, k, sm + OFF_REC + cur * RECSZ, gl);
                }
                __syncwarp();
            }
            /* (Fenrir build: ncu, profiles/r1t, shows 49 % of the samples at this barrier -- the other warps wait for a warp's
               single-lane data update.  Dropping the lockstep for a warp-level exit test was measured 45 % SLOWER (the
               warps drift apart in the unrolled body, instruction fetch); the remedy was a cheaper update -- the pre-array
               form of pn_fenrir_update.cuh, 2.3x on the whole call -- and, next, a group-cooperative one.) */
            if (__syncthreads_or(((have || !queue_empty) ? 1 : 0)) == 0) break;

            /* ===== (B) one backward step for every group that has one ===== */
            const bool active = have;
            const double h = active ? P.s_h[k] : 1.0;
            const bool copy = active && (h == 0.0); /* integrator_utils.jl:109-112 */
            const bool doit = active && !copy;
            if (copy) {
                if (!fen && P.write_x) write_record(P, k, Sm, Ms, scv, gl);
                /* a copy step consumes no staged record: request the one of the next real step from here */
                if (ASYNC && k > a0 && P.s_h[k - 1] != 0.0) stage_async(P, k - 1, sm + OFF_REC + cur * RECSZ, gl);
            }
            __syncwarp();
            if (__any_sync(FULL, doit)) {
                const double hh = doit ? h : 1.0;
                const double ediff = doit ? (P.dynamic ? P.s_diffusion[k] : P.initial_diffusion) : 1.0;
                /* the filtered record of this step was requested one step ago (cp.async, double-buffered); request the
                   next one now so that its HBM / L2 latency hides under this step's arithmetic */
                if (ASYNC) {
                    Rec = sm + OFF_REC + cur * RECSZ;
#ifndef PN_HOST_EMULATION
                    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
                    if (!doit)
                        for (int e = gl; e < D * D + D; e += G) Rec[e] = 0.0;
                    if (doit && k > a0 && P.s_h[k - 1] != 0.0) stage_async(P, k - 1, sm + OFF_REC + (cur ^ 1) * RECSZ, gl);
                    if (doit) cur ^= 1;
                } else if (doit) {
                    const double* src = P.s_x_covfac + k * D * D;
                    for (int e = gl; e < D * D; e += G) Rec[e] = src[e];
                    for (int c = gl; c < D; c += G) Rec[D * D + c] = P.s_x_mean[k * D + c];
                } else {
                    for (int e = gl; e < D * D + D; e += G) Rec[e] = 0.0;
                }
                __syncwarp();
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
                double sdv[d]; /* sqrt of the diffusion of this step, per component */
                {
                    const double sd0 = (ediff == 1.0) ? 1.0 : sqrt(ediff);
#pragma unroll
                    for (int i = 0; i < d; i++) sdv[i] = (P.s_diffusion_mv && P.dynamic && doit) ? sqrt(P.s_diffusion_mv[k * d + i]) : sd0;
                }
                double A[W][NS];
                bool isR[NS];
                int cR[NS]; /* right-block column index (0..D-1) held in slot s, or 0 */
#pragma unroll
                for (int s = 0; s < NS; s++) {
                    const int c = s * G + gl;
                    const bool left = c < D;
                    isR[s] = (c >= D) && (c < W);
                    cR[s] = isR[s] ? c - D : 0;
                    const int cl = left ? c : 0;
                    const int j = cl / d, i = cl - j * d;
                    /* noise column: U[m][j] PI[j] (upper triangular: zero for m > j) */
                    double upj[n];
#pragma unroll
                    for (int m = 0; m < n; m++) {
                        double u = 0.0;
#pragma unroll
                        for (int jj = 0; jj < n; jj++) {
                            if (jj < m) continue;
#if PN_IOUP
                            u = (jj == j) ? Uh[m][jj] * PIv[jj] : u;
#else
                            u = (jj == j) ? P.U[m * n + jj] * PIv[jj] : u;
#endif
                        }
                        upj[m] = left ? u : 0.0;
                    }
#if PN_IOUP
                    double tqj = 0.0;
#pragma unroll
                    for (int jj = 0; jj < n; jj++) tqj = (jj == j) ? tq[jj] : tqj;
#endif
#pragma unroll
                    for (int r = 0; r < D; r++) {
                        /* left: X[r][c] = sum_m T[j][j+m] R[r][c + m d] ; right: R[r][c - D] */
                        double x = 0.0;
#pragma unroll
                        for (int m = 0; m < n; m++) {
                            const bool in = left && (c + m * d < D);
#if PN_IOUP
                            const double coef = (j + m == Q) ? tqj : hp[m];
#else
                            const double coef = hp[m];
#endif
                            const double rv = in ? Rec[r * D + c + m * d] : 0.0;
                            x += coef * rv;
                        }
                        const double rr = isR[s] ? Rec[r * D + cR[s]] : 0.0;
                        A[r][s] = left ? x : rr;
                        /* bottom row r = (mr, ir) */
                        const int mr = r / d, ir = r - mr * d;
                        A[D + r][s] = (left && ir == i) ? upj[mr] * sdv[ir] : 0.0;
                    }
                }
                pn_qr_cols<G, W, NS, D, 0, D, true>(A, gl, Di);
                /* mean bookkeeping (replicated): delta = mu_s+ - Ah mu */
                double delta[D];
#pragma unroll
                for (int j = 0; j < n; j++)
#pragma unroll
                    for (int i = 0; i < d; i++) {
                        double s = 0.0;
#pragma unroll
                        for (int m = 0; m < n; m++) {
                            if (j + m >= n) continue;
#if PN_IOUP
                            const double coef = (j + m == Q) ? tq[j] : hp[m];
#else
                            const double coef = hp[m];
#endif
                            s += coef * Rec[D * D + (j + m) * d + i];
                        }
                        delta[j * d + i] = Ms[j * d + i] - s;
                    }
                /* publish the strict upper part of R^- */
#pragma unroll
                for (int s = 0; s < NS; s++) {
                    const int c = s * G + gl;
                    if (s * G >= D) continue;
#pragma unroll
                    for (int r = 0; r < D; r++)
                        if (r < c && c < D) Rm[r * LD + c] = A[r][s];
                }
                __syncwarp();
                /* G' = R^-^-1 Y for the right-block columns of this lane (back substitution, broadcast reads of R^-) */
                double gc[NS][D];
#pragma unroll
                for (int r = D - 1; r >= 0; r--) {
                    const double dr = Di[r];
                    double rrow[D];
#pragma unroll
                    for (int kk = 0; kk < D; kk++)
                        if (kk > r) rrow[kk] = Rm[r * LD + kk];
#pragma unroll
                    for (int s = 0; s < NS; s++) {
                        if (s < SR0) continue;
                        double acc = A[r][s];
#pragma unroll
                        for (int kk = 0; kk < D; kk++)
                            if (kk > r) acc -= rrow[kk] * gc[s][kk];
                        gc[s][r] = acc * dr;
                    }
                }
                /* smoothed mean of the own columns and the top block R_s+ G' */
                double mnew[NS];
#pragma unroll
                for (int s = 0; s < NS; s++) {
                    if (s < SR0) continue;
                    double acc = Rec[D * D + cR[s]];
#pragma unroll
                    for (int r = 0; r < D; r++) acc += gc[s][r] * delta[r];
                    mnew[s] = acc;
                }
#pragma unroll
                for (int r = 0; r < D; r++) {
                    double srow[D];
#pragma unroll
                    for (int kk = 0; kk < D; kk++) srow[kk] = Sm[r * LD + kk];
#pragma unroll
                    for (int s = 0; s < NS; s++) {
                        if (s < SR0) continue;
                        double acc = 0.0;
#pragma unroll
                        for (int kk = 0; kk < D; kk++) acc += srow[kk] * gc[s][kk];
                        A[r][s] = isR[s] ? acc : 0.0;
                    }
                }
                __syncwarp(); /* everyone is done reading Rm / Sm / Ms / Rec */
                pn_qr_cols<G, W, NS, D, D, D, false>(A, gl, nullptr);
                if (doit) {
#pragma unroll
                    for (int s = 0; s < NS; s++) {
                        if (s < SR0) continue;
                        if (isR[s]) {
#pragma unroll
                            for (int r = 0; r < D; r++) Sm[r * LD + cR[s]] = (r <= cR[s]) ? A[r][s] : 0.0;
                            Ms[cR[s]] = mnew[s];
                        }
                    }
                }
                __syncwarp();
                if (doit && !fen) {
                    if (P.write_x) write_record(P, k, Sm, Ms, scv, gl);
                    if (gl == 0) write_pu(P, k, Sm, Ms, scv);
                }
                if (doit && fen && data_idx >= 0 && P.s_t[k] == P.data_t[data_idx]) { /* fenrir.jl:110-123 */
                    if (gl == 0) {
                        const double ll = pn_fenrir_update_serial<D, LD, OMAX>(P.obs_H, P.obs_Rn, P.data_u + data_idx * P.n_obs, P.n_obs, Sm, Ms,
                                                                               P.kron_logdet != 0);
                        if (!(ll == PN_INF || ll == -PN_INF)) LL += ll; /* !isinf(ll): a NaN propagates, as in fenrir.jl:118-120 */
                    }
                    data_idx--;
                }
                __syncwarp();
            }
            if (active) {
                k -= 1;
                have = (k >= a0);
                if (fen && !have && gl == 0) {
                    const bool ok = (!P.retcode || P.retcode[tr_cur] == PN_RET_SUCCESS) && data_idx < 0;
                    P.fenrir_ll[tr_cur] = ok ? LL : -PN_INF;
                }
            }
        }
    }
};

template <int d, int Q, int G>
__device__ __forceinline__ void pn_smooth_col_entry(const PnSmoothParams& P) {
    extern __shared__ double pn_smc_shared[];
    PnSmoothCol<d, Q, G>::run(P, pn_smc_shared);
}

#endif /* PN_SMOOTH_COL_CUH */
