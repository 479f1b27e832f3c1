// pn_small.cuh -- group-per-trajectory ODE-filter kernel for small filter states (D = d(q+1) <= 32).
//
// A "group" of G lanes (G in {2,4,8,16,32}, 32/G trajectories per warp) owns one trajectory for the
// whole time loop.  The square-root covariance factor R (D x D, Sigma = R'R) lives in registers,
// row-distributed: lane gl owns rows r == gl (mod G).  Row ownership makes every step of the reference
// algorithm either lane-local or a G-lane butterfly reduction:
//
//   reference routine (file:line)                         here
//   -------------------------------------------------     ---------------------------------------------
//   make_transition_matrices!   priors/iwp.jl:182-190      h-powers in registers; Ah = T (x) I_d and
//   make_preconditioner[_inv]!  preconditioning.jl:30-58   Qh.R = (L P^-1) (x) I_d are never materialised
//   predict_mean!               filtering/predict.jl:53    replicated, n(n+1)/2 * d FMAs
//   measurement z, H            measurement_models.jl:11   user f / jac inlined (NVRTC or builtin)
//                               derivative_utils.jl:1-33
//   local_scalar_diffusion      diffusions/calibration.jl:123-133  W = C'C with C = Qh.R H' (structured)
//   compute_scaled_error_estimate!  perform_step.jl:199-247  from diag(W)
//   PI controller, accept/reject    OrdinaryDiffEqCore, per trajectory, in-kernel
//   predict_cov! + triangularize!   filtering/predict.jl:62-94, fast_linalg.jl:70-79
//                                   Householder QR of the 2D x D stack [R Ah' ; sqrt(s2) Qh.R], rows in
//                                   registers, column norms / v'a reduced with warp shuffles
//   compute_measurement_covariance! perform_step.jl:182-188   row-local C2 = R^- H', S all-reduced
//   update!                         filtering/update.jl:65-139 K all-reduced, Joseph form R+ = R^- - C2 K'
//   estimate_global_diffusion       diffusions/calibration.jl:49-66
//
// Groups pull trajectories from an atomic work counter (ragged adaptive step counts).
// All arithmetic is FP64; tensor cores are not used (per-trajectory matrices are <= 32 x 32).
//
// NVRTC-compatible: no standard headers.
#ifndef PN_SMALL_CUH
#define PN_SMALL_CUH
#include "pn_params.h"
#include "pn_taylor.cuh"

// Butterfly all-reduce over the G lanes of a group.  Called warp-convergently (full mask), so it compiles
// to bare SHFL.BFLY pairs; every lane of the group ends with the bitwise-identical sum.
template <int G>
PN_HD double pn_gsum(double x) {
#pragma unroll
    for (int o = G >> 1; o >= 1; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o, 32);
    return x;
}
template <int G>
PN_HD double pn_gbcast(double x, int src) {
    return __shfl_sync(0xffffffffu, x, src, G);
}

// 1/x and 1/sqrt(x) from the MUFU seed (rcp.approx / rsqrt.approx: ~2^-23, full exponent range) plus two Newton
// steps in FP64: <= 2 ulp, no slow-path branches.  The IEEE-rounded `1.0/x` and `sqrt(x)` each cost a
// ~25-instruction sequence with a call on the slow path; the QR needs one of each per column.
PN_HD double pn_rcp(double x) {
    double r;
#ifdef PN_HOST_EMULATION /* tests/emulation: the kernel headers compiled for the host; a seed of the same accuracy */
    r = pn_host_seed_rcp(x);
#else
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#endif
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}
PN_HD double pn_rsqrt(double x) {
    double y;
#ifdef PN_HOST_EMULATION
    y = pn_host_seed_rsqrt(x);
#else
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#endif
    double t = x * y;
    double e = fma(-t, y, 1.0);
    y = fma(0.5 * y, e, y);
    t = x * y;
    e = fma(-t, y, 1.0);
    y = fma(0.5 * y, e, y);
    return y;
}

#ifndef PN_FAST_SCALAR
#define PN_FAST_SCALAR 1 /* 1: the scalar bookkeeping of a step (powers of h, controller factors, eps(t)) multiplies by reciprocals
                            (pn_rcp / host-computed constants) instead of IEEE divisions and takes eps from the exponent field:
                            <= 2 ulp per quantity, ~250 fewer instructions and several ~100-cycle dependent chains per step;
                            0: the literal divisions / nextafter of round 1 */
#endif
/* nextafter(x, +inf) - x for finite x > 0: the spacing of the doubles in x's binade, 2^(e - 52), from the exponent field
   (exact; subnormal / huge arguments take the library route) */
__device__ __forceinline__ double pn_ulp_above(double x) {
    const int ex = (__double2hiint(x) >> 20) & 0x7ff;
    if (ex > 52 && ex < 0x7ff) return __hiloint2double((ex - 52) << 20, 0);
    return nextafter(x, 1e300) - x;
}

// Upper Cholesky of a symmetric M (row-major, only j >= i read) in place: M = U'U, inv[j] = 1/U[j][j].
// Returns false if a pivot is not positive.
template <int M_>
PN_HD bool pn_chol(double (&A)[M_][M_], double (&invd)[M_]) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < M_; j++) {
        double s = A[j][j];
#pragma unroll
        for (int k = 0; k < M_; k++)
            if (k < j) s -= A[k][j] * A[k][j];
        if (!(s > 0.0)) ok = false;
        const double inv = pn_rsqrt(s);
        const double ujj = s * inv;
        A[j][j] = ujj;
        invd[j] = inv;
#pragma unroll
        for (int i = 0; i < M_; i++) {
            if (i <= j) continue;
            double v = A[j][i];
#pragma unroll
            for (int k = 0; k < M_; k++)
                if (k < j) v -= A[k][j] * A[k][i];
            A[j][i] = v * inv;
        }
    }
    return ok;
}
// x <- (U'U)^{-1} x
template <int M_>
PN_HD void pn_chol_solve(const double (&U)[M_][M_], const double (&invd)[M_], double (&x)[M_]) {
#pragma unroll
    for (int a = 0; a < M_; a++) {
#pragma unroll
        for (int k = 0; k < M_; k++)
            if (k < a) x[a] -= U[k][a] * x[k];
        x[a] *= invd[a];
    }
#pragma unroll
    for (int a = M_ - 1; a >= 0; a--) {
#pragma unroll
        for (int k = 0; k < M_; k++)
            if (k > a) x[a] -= U[a][k] * x[k];
        x[a] *= invd[a];
    }
}

// Householder QR of the stack [Tt ; Bt], rows in registers: slot lr of lane gl holds row lr*G + gl of the block.
// Columns 0..NE-1 are eliminated; columns NE..NC-1 only receive the reflections.  BOTUP: the first NE columns of
// the bottom block are upper triangular (row r zero in columns < r), so bottom slot lr joins at column lr*G and
// entries left of its first row are never touched.  On return Tt[:, 0:NE] is the upper-triangular factor (zeros
// below the diagonal); if dinv != nullptr lane (k % G) stores 1/diag[k] (0 for a zero column) to dinv[k].
// The j loop is chunked (JC columns at a time) to bound the number of live partial sums.
// live = false (per group): every reflector is the identity and the matrix is left untouched bit for bit -- the shuffles
// still run for the whole warp; used for groups with nothing to commit, so that a trajectory's arithmetic never depends on
// which other trajectories happen to share its warp.
// Householder scalars of one pivot column
struct PnHouse {
    double beta, vk, gam, binv;
};
// norm, pivot broadcast and reflector scalars of column k of the stack (rows >= k of the top block, active bottom slots)
template <int G, int RPL, int NC, bool BOTUP>
PN_HD PnHouse pn_qr_house(const double (&Tt)[RPL][NC], const double (&Bt)[RPL][NC], int gl, int k, bool live) {
    const int ks = k / G, kl = k % G;
    double part = 0.0, xp = 0.0;
#pragma unroll
    for (int lr = 0; lr < RPL; lr++) {
        const int base = lr * G;
        if (base + G - 1 < k) continue;
        const double xv = (base >= k) ? Tt[lr][k] : ((base + gl >= k) ? Tt[lr][k] : 0.0);
        if (lr == ks) xp = xv;
        part += xv * xv;
    }
#pragma unroll
    for (int lr = 0; lr < RPL; lr++)
        if (!BOTUP || lr * G <= k) part += Bt[lr][k] * Bt[lr][k];
    const double total = pn_gsum<G>(part);
    const double alpha = pn_gbcast<G>(xp, kl);
    const bool nz = total > 0.0;
    const double rsq = pn_rsqrt(total);
    const double nrm = nz ? total * rsq : 0.0;
    PnHouse h;
    h.beta = -copysign(nrm, alpha);
    h.vk = alpha - h.beta;
    h.gam = (nz && live) ? pn_rcp(total + fabs(alpha) * nrm) : 0.0; /* 2 / v'v; 0: the reflector is the identity */
    h.binv = nz ? -copysign(rsq, alpha) : 0.0;
    return h;
}

template <int G, int RPL, int NC, int NE, bool BOTUP, int JC>
PN_HD void pn_qr_rows(double (&Tt)[RPL][NC], double (&Bt)[RPL][NC], int gl, double* dinv, bool live = true) {
    /* software-pipelined: the scalars of pivot k+1 (butterfly, rsqrt, rcp: a ~200-cycle dependent chain) are started
       right after reflector k has updated column k+1, so the chain runs under the updates of columns k+2... */
    PnHouse hs = pn_qr_house<G, RPL, NC, BOTUP>(Tt, Bt, gl, 0, live);
#pragma unroll
    for (int k = 0; k < NE; k++) {
        const int ks = k / G, kl = k % G;
        const PnHouse hk = hs;
        double x[RPL];
#pragma unroll
        for (int lr = 0; lr < RPL; lr++) {
            const int base = lr * G;
            if (base + G - 1 < k) {
                x[lr] = 0.0;
            } else if (base >= k) {
                x[lr] = Tt[lr][k];
            } else {
                x[lr] = (base + gl >= k) ? Tt[lr][k] : 0.0;
            }
        }
        if (dinv != nullptr && gl == kl) dinv[k] = hk.binv;
        if (gl == kl) x[ks] = hk.vk;
        /* column k+1 first, then the next pivot's scalars */
        if (k + 1 < NC) {
            const int j = k + 1;
            double s = 0.0;
#pragma unroll
            for (int lr = 0; lr < RPL; lr++)
                if (lr * G + G - 1 >= k) s += x[lr] * Tt[lr][j];
#pragma unroll
            for (int lr = 0; lr < RPL; lr++)
                if (!BOTUP || lr * G <= k) s += Bt[lr][k] * Bt[lr][j];
            s = pn_gsum<G>(s) * hk.gam;
#pragma unroll
            for (int lr = 0; lr < RPL; lr++)
                if (lr * G + G - 1 >= k) Tt[lr][j] -= s * x[lr];
#pragma unroll
            for (int lr = 0; lr < RPL; lr++)
                if (!BOTUP || lr * G <= k) Bt[lr][j] -= s * Bt[lr][k];
        }
        if (k + 1 < NE) hs = pn_qr_house<G, RPL, NC, BOTUP>(Tt, Bt, gl, k + 1, live);
#pragma unroll
        for (int j0 = 0; j0 < NC; j0 += JC) {
            if (j0 + JC - 1 <= k + 1) continue;
            double sj[JC];
#pragma unroll
            for (int jj = 0; jj < JC; jj++) {
                const int j = j0 + jj;
                sj[jj] = 0.0;
                if (j <= k + 1 || j >= NC) continue;
                double s = 0.0;
#pragma unroll
                for (int lr = 0; lr < RPL; lr++)
                    if (lr * G + G - 1 >= k) s += x[lr] * Tt[lr][j];
#pragma unroll
                for (int lr = 0; lr < RPL; lr++)
                    if (!BOTUP || lr * G <= k) s += Bt[lr][k] * Bt[lr][j];
                sj[jj] = s;
            }
#pragma unroll
            for (int jj = 0; jj < JC; jj++) {
                const int j = j0 + jj;
                if (j <= k + 1 || j >= NC) continue;
                sj[jj] = pn_gsum<G>(sj[jj]) * hk.gam;
            }
#pragma unroll
            for (int jj = 0; jj < JC; jj++) {
                const int j = j0 + jj;
                if (j <= k + 1 || j >= NC) continue;
#pragma unroll
                for (int lr = 0; lr < RPL; lr++)
                    if (lr * G + G - 1 >= k) Tt[lr][j] -= sj[jj] * x[lr];
#pragma unroll
                for (int lr = 0; lr < RPL; lr++)
                    if (!BOTUP || lr * G <= k) Bt[lr][j] -= sj[jj] * Bt[lr][k];
            }
        }
#pragma unroll
        for (int lr = 0; lr < RPL; lr++) {
            const int base = lr * G;
            if (base + G - 1 < k) continue;
            const int row = base + gl;
            { /* a group that is not live keeps its matrix bit for bit (gam = 0 above); selects, not branches: the
                 predicated form made ptxas demote the row arrays to local memory */
                const double keep = Tt[lr][k];
                const double fin = (row == k) ? hk.beta : ((row > k) ? 0.0 : keep);
                Tt[lr][k] = live ? fin : keep;
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// predict_cov! on the reference's own route (filtering/predict.jl:84-85): M = X'X + s2 Qh, Sigma^-.R = cholesky(M).U, with
// triangularize! only when the factorisation fails (:90).  X = R Ah' is row-distributed in registers; the Gram matrix
// needs every row on every lane, so the rows are published once in the group's shared-memory exchange region and each
// lane streams all of them (16-byte broadcast loads) into the accumulators of ITS rows of M -- no cross-lane reduction,
// and X leaves the registers for the duration of the factorisation.  Arithmetic per lane at D = 12, G = 4: 288 DFMA
// (Gram) + ~250 (Cholesky) against ~1030 for the Householder sweep of the 2D x D stack.
// ---------------------------------------------------------------------------------------------------------------------
#ifndef PN_PREDICT_CHOL
#define PN_PREDICT_CHOL 1 /* 0: always triangularize! (Householder QR of the stack, the round-1 kernel) */
#endif
/* doubles per group of the exchange region: D*D rounded so that the 32/G groups of a warp start 16 bytes apart modulo the
   128 bytes of banks (their broadcast loads do not collide) */
template <int D>
PN_HD constexpr int pn_xs_doubles() {
    return ((D * D + 15) / 16) * 16 + 2;
}

#ifndef PN_R_IN_SMEM
#define PN_R_IN_SMEM 0 /* 1 (needs PN_PREDICT_CHOL): the filtered factor lives in the group's exchange region BETWEEN iterations and is
                          loaded at the start of phase 2, so the registers of R are free during phase 1 (f, J, sigma^2, controller) */
#endif
/* own rows of a row-cyclic D x D matrix <-> the group's exchange region (row-major) */
template <int G, int RPL, int D>
PN_HD void pn_rows_store(double* sx, const double (&R)[RPL][D], int gl) {
#pragma unroll
    for (int lr = 0; lr < RPL; lr++) {
        const int row = lr * G + gl;
        if (row < D) {
            double* dst = sx + row * D;
            if constexpr (D % 2 == 0) {
#pragma unroll
                for (int c = 0; c < D; c += 2) *reinterpret_cast<double2*>(dst + c) = make_double2(R[lr][c], R[lr][c + 1]);
            } else {
#pragma unroll
                for (int c = 0; c < D; c++) dst[c] = R[lr][c];
            }
        }
    }
}
template <int G, int RPL, int D>
PN_HD void pn_rows_load(const double* sx, double (&R)[RPL][D], int gl) {
#pragma unroll
    for (int lr = 0; lr < RPL; lr++) {
        const int row = lr * G + gl;
        const double* src = sx + (row < D ? row : 0) * D;
        if constexpr (D % 2 == 0) {
#pragma unroll
            for (int c = 0; c < D; c += 2) {
                const double2 v = *reinterpret_cast<const double2*>(src + c);
                R[lr][c] = (row < D) ? v.x : 0.0;
                R[lr][c + 1] = (row < D) ? v.y : 0.0;
            }
        } else {
#pragma unroll
            for (int c = 0; c < D; c++) R[lr][c] = (row < D) ? src[c] : 0.0;
        }
    }
}

// Own rows of X'X added to the accumulators Mw (entries j >= lr*G of slot lr): every row of X is streamed from the group's
// exchange region (16-byte broadcast loads), no cross-lane reduction.
#ifndef PN_GRAM_UNROLL
#define PN_GRAM_UNROLL 2
#endif
#define PN_STR2(x) #x
#define PN_PRAGMA_UNROLL(k) _Pragma(PN_STR2(unroll k))
template <int G, int RPL, int D>
PN_HD void pn_gram_rows(const double* sx, double (&Mw)[RPL][D], int gl) {
    PN_PRAGMA_UNROLL(PN_GRAM_UNROLL)
    for (int r = 0; r < D; r++) {
        double xr[D], xo[RPL];
        const double* src = sx + r * D;
        if constexpr (D % 2 == 0) {
#pragma unroll
            for (int c = 0; c < D; c += 2) {
                const double2 v = *reinterpret_cast<const double2*>(src + c);
                xr[c] = v.x;
                xr[c + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int c = 0; c < D; c++) xr[c] = src[c];
        }
#pragma unroll
        for (int lr = 0; lr < RPL; lr++) xo[lr] = (lr * G + gl < D) ? src[lr * G + gl] : 0.0;
#pragma unroll
        for (int lr = 0; lr < RPL; lr++)
#pragma unroll
            for (int j = 0; j < D; j++)
                if (j >= lr * G) Mw[lr][j] = fma(xo[lr], xr[j], Mw[lr][j]);
    }
}

// Right-looking Cholesky of the row-cyclic symmetric M (slot lr of lane gl holds row lr*G + gl; entries j >= row valid, the
// rest is never read).  On return R holds the upper-triangular factor in the same distribution (zeros left of the
// diagonal) and the function tells whether every pivot was positive.  The pivot row is broadcast with shuffles; a lane's
// element of the pivot column is, by symmetry, one of the broadcast values (selected by its lane index).
template <int G, int RPL, int D>
PN_HD bool pn_chol_rows(double (&Mw)[RPL][D], double (&R)[RPL][D], int gl) {
    bool ok = true;
#pragma unroll
    for (int k = 0; k < D; k++) {
        const int ks = k / G, kl = k % G;
        double rk[D];
#pragma unroll
        for (int j = 0; j < D; j++) rk[j] = (j >= k) ? pn_gbcast<G>(Mw[ks][j], kl) : 0.0;
        ok = ok && (rk[k] > 0.0);
        const double pinv = pn_rcp(rk[k]);
#pragma unroll
        for (int lr = 0; lr < RPL; lr++) {
            const int base = lr * G;
            if (base + G - 1 <= k) continue; /* every row of this slot is finished */
            double mik = 0.0;                /* M[row][k] = M[k][row] */
#pragma unroll
            for (int g = 0; g < G; g++)
                if (base + g > k && base + g < D) mik = (gl == g) ? rk[base + g] : mik;
            const double f = mik * pinv;
#pragma unroll
            for (int j = 0; j < D; j++)
                if (j >= base && j > k) Mw[lr][j] = fma(-f, rk[j], Mw[lr][j]);
        }
    }
    /* U[row][j] = M[row][j] / sqrt(pivot of the row), pivot = M[row][row] as left by the eliminations before it */
#pragma unroll
    for (int lr = 0; lr < RPL; lr++) {
        const int base = lr * G;
        double piv = 1.0;
#pragma unroll
        for (int g = 0; g < G; g++)
            if (base + g < D) piv = (gl == g) ? Mw[lr][base + g] : piv;
        const double ri = pn_rsqrt(piv);
#pragma unroll
        for (int j = 0; j < D; j++) R[lr][j] = (j >= base + gl && base + gl < D) ? Mw[lr][j] * ri : 0.0;
    }
    return ok;
}

// triangularize! (fast_linalg.jl:70-79) of the stack [X ; B] for ONE trajectory, executed by a single lane on the group's
// exchange region: the cold path of predict_cov! (predict.jl:90), taken when the Cholesky factorisation of the Gram matrix
// fails or the diffusion is exactly zero.  X (D x D row-major in shared memory) is reduced in place by Householder
// reflections, then the rows of the upper-triangular noise block B (brow(m, c) = entry of row (m, i) at column (c, i)) are
// folded in one at a time with Givens rotations.  Not inlined: its local-memory frame belongs to this path only.
template <int D>
__device__ __noinline__ void pn_house_serial(double* X) {
    for (int k = 0; k < D; k++) {
        double tot = 0.0;
        for (int r = k; r < D; r++) tot += X[r * D + k] * X[r * D + k];
        if (tot > 0.0) {
            const double alpha = X[k * D + k], nrm = sqrt(tot), beta = -copysign(nrm, alpha), vk = alpha - beta;
            const double gam = 1.0 / (tot + fabs(alpha) * nrm);
            for (int j = k + 1; j < D; j++) {
                double sj = vk * X[k * D + j];
                for (int r = k + 1; r < D; r++) sj += X[r * D + k] * X[r * D + j];
                sj *= gam;
                X[k * D + j] -= sj * vk;
                for (int r = k + 1; r < D; r++) X[r * D + j] -= sj * X[r * D + k];
            }
            X[k * D + k] = beta;
        }
        for (int r = k + 1; r < D; r++) X[r * D + k] = 0.0;
    }
}
/* fold the dense row b (entries left of j0 zero) into the upper-triangular X with Givens rotations; b is destroyed */
template <int D>
__device__ __forceinline__ void pn_givens_fold(double* X, double* b, int j0) {
    for (int j = j0; j < D; j++) { /* rotate (X[j][j:], b[j:]) so that b[j] vanishes */
        const double a = X[j * D + j], bb = b[j];
        if (bb == 0.0) continue;
        const double rr = sqrt(a * a + bb * bb), cs = a / rr, sn = bb / rr;
        for (int c = j; c < D; c++) {
            const double xv = X[j * D + c], bv = b[c];
            X[j * D + c] = cs * xv + sn * bv;
            b[c] = cs * bv - sn * xv;
        }
    }
}
template <int D, int d, int n>
__device__ __noinline__ void pn_triangularize_serial(double* X, const double* bfac /* n x n row-major, upper */) {
    pn_house_serial<D>(X);
    double b[D];
    for (int m = 0; m < n; m++)
        for (int i = 0; i < d; i++) {
            for (int c = 0; c < D; c++) b[c] = 0.0;
            bool any = false;
            for (int c = m; c < n; c++) {
                b[c * d + i] = bfac[m * n + c];
                any = any || (b[c * d + i] != 0.0);
            }
            if (!any) continue;
            pn_givens_fold<D>(X, b, m * d + i);
        }
}
/* triangularize!([X ; E]) with NE dense extra rows E (row-major NE x D, destroyed): the fallback of the observation-noise
   update (update.jl:133-135, [x_out.Sigma.R ; (K R.R')']) */
template <int D, int NE>
__device__ __noinline__ void pn_triangularize_serial_dense(double* X, double* E) {
    pn_house_serial<D>(X);
    for (int a = 0; a < NE; a++) pn_givens_fold<D>(X, E + a * D, 0);
}

// ode_determine_initdt (OrdinaryDiffEqCore, Hairer-Wanner; restated from SURVEY.md A.6).  Two f evaluations.
template <class Prob>
PN_HD double pn_initial_dt(const PnParams& P, const double* u0, const double* pr) {
    constexpr int d = Prob::d;
    double dt;
    double sk[d], f0[d], f1[d], u1[d];
    const double epst = fabs(P.t0) > 0.0 ? (nextafter(fabs(P.t0), 1e300) - fabs(P.t0)) : 4.9406564584124654e-324;
    const double dtmin = nextafter(fmax(P.dtmin, epst), 1e300);
    const double smalldt = fmax(dtmin, 1e-6);
    double d0 = 0.0, d1 = 0.0, d2 = 0.0;
    Prob::template f<double>(f0, u0, pr, P.t0);
#pragma unroll
    for (int i = 0; i < d; i++) {
        sk[i] = P.abstol + fabs(u0[i]) * P.reltol;
        const double a = u0[i] / sk[i], b = f0[i] / sk[i];
        d0 += a * a;
        d1 += b * b;
    }
    d0 = sqrt(d0 / (double)d);
    d1 = sqrt(d1 / (double)d);
    double dta = (d0 < 1e-5 || d1 < 1e-5) ? smalldt : (d0 / d1) / 100.0;
    dta = fmin(dta, P.dtmax);
    if (dta < 10.0 * 2.220446049250313e-16) {
        dt = fmax(smalldt, dtmin);
    } else {
#pragma unroll
        for (int i = 0; i < d; i++) u1[i] = u0[i] + dta * f0[i];
        Prob::template f<double>(f1, u1, pr, P.t0 + dta);
#pragma unroll
        for (int i = 0; i < d; i++) {
            const double c = (f1[i] - f0[i]) / sk[i];
            d2 += c * c;
        }
        d2 = sqrt(d2 / (double)d) / dta;
        const double mx = fmax(d1, d2);
        const double dtb = (mx <= 1e-15) ? fmax(smalldt, dta * 1e-3) : pow(10.0, -(2.0 + log10(mx)) / (double)P.order);
        dt = fmax(dtmin, fmin(fmin(100.0 * dta, dtb), P.dtmax));
    }
    return dt;
}

#ifndef PN_IOUP
#define PN_IOUP 0 /* 1: integrated Ornstein-Uhlenbeck prior with scalar rate P.rate (pn_ioup.cuh); 0: IWP */
#endif
#ifndef PN_EK0_DENSE
#define PN_EK0_DENSE 0 /* 1: EK0 measurement (H = E1, no Jacobian) on the dense layout -- EK0 with a non-IWP prior */
#endif
/* Lazily accumulated sum of logarithms: sum_i log(x_i) = log(m) + e ln 2, with m the running product of the mantissas kept
   in [1, 2) and e the sum of the exponents.  The log-likelihood needs log det S of every accepted step; taking ONE log per
   trajectory instead of one per step removes a slow-path call from the step loop (used by pn_kron: +5 % on cfg2).  Returns false (and leaves the
   accumulator alone) for factors outside the normal range -- zero, subnormal, huge, Inf, NaN: the caller takes the
   logarithm directly so that +-Inf / NaN propagate as before. */
__device__ __forceinline__ bool pn_logacc_mul(double& m, int& e, double x) {
    if (!(x >= 1e-290 && x <= 1e290)) return false;
    m *= x;
    const int hi = __double2hiint(m);
    const int ex = ((hi >> 20) & 0x7ff) - 1023;
    e += ex;
    m = __hiloint2double(hi - (ex << 20), __double2loint(m));
    return true;
}
__device__ __forceinline__ double pn_logacc_value(double m, int e) { return log(m) + (double)e * 0.6931471805599453; }
#ifndef PN_SECOND
#define PN_SECOND 0 /* 1: SecondOrderODEProblem.  Prob is the first-order form F([du; u]) = [f1(du, u); du] of dimension
                       2d (what the reference's DynamicalODEFunction evaluates on ArrayPartition(du, u)); the state models
                       the d positions u and their q derivatives, z = E2 x - f1(E1 x, E0 x) (measurement_models.jl:29-49),
                       H = E2 - J_du E1 - J_u E0 (derivative_utils.jl:1-53), sol.u = SolProj x = (E1 x, E0 x)
                       (caches.jl:126); error scaling uses the du partition (perform_step.jl:204-216). */
#endif
/* SolProj row j -> state entry: E0 rows, or [E1; E0] for second-order problems */
template <int d>
PN_HD constexpr int pn_solcol(int j) {
    return PN_SECOND ? (j < d ? d + j : j - d) : j;
}
#ifndef PN_OBSN
#define PN_OBSN 0 /* 1: EK0/EK1(pn_observation_noise = R): the generated source defines constexpr pn_obsn_cov(a,b) = (R.R'R.R)[a][b] and
                     pn_obsn_rt(a,b) = R.R[a][b] (cov2psdmatrix, caches.jl:175-178).  S += R (perform_step.jl:185-187), the filtered
                     covariance gains K R K' and is re-factorised (update.jl:125-136). */
#endif
#ifndef PN_DATA
#define PN_DATA 0 /* 1: forward data updates at the observation times (DataUpdateCallback, src/callbacks/dataupdate.jl:39-84, as used by
                     filtering_data_loglik / dalton_data_loglik): the generated source defines PN_DATA_O = length of an observation and
                     constexpr pn_data_H(a,c) = (observation_matrix * SolProj)[a][c], pn_data_cov(a,b) = R[a][b], pn_data_rt(a,b) = R.R[a][b] */
#endif
#ifndef PN_MASS
#define PN_MASS 0 /* 1: mass matrix M u' = f; the generated source defines constexpr pn_mass(a,i) = M[a][i] and
                     pn_mass_inv(a,i) = (M + 1e-20 I if M has a zero diagonal entry)^-1 (autodiffinit.jl:20-31), so zero
                     entries fold away at compile time.  z = M E1 x - f, H = M E1 - J E0. */
#endif
#if PN_IOUP
#include "pn_ioup.cuh"
#endif

template <class Prob, int Q, int G, bool ADAPTIVE, bool DYNAMIC, int SAVE>
struct PnSmall {
    static constexpr int DP = Prob::d;                       /* dimension of the right-hand side */
    static constexpr int d = PN_SECOND ? DP / 2 : DP;        /* modelled components (positions for second-order problems) */
    static constexpr int DU = DP;                            /* length of sol.u */
    static constexpr int E1B = PN_SECOND ? 2 : 1;            /* derivative block the measurement reads (E2 / E1) */
    static constexpr int n = Q + 1;
    static constexpr int D = d * n;
    static constexpr int RPL = (D + G - 1) / G; /* rows per lane, top and bottom block each */
    static constexpr int NPA = Prob::n_p > 0 ? Prob::n_p : 1;
    /* H = [-J | I | 0 ...] touches the first 2d columns only and R^- is upper triangular, so R^- H' is zero in
       rows >= 2d: row slots lr >= RU never enter the measurement update */
    static constexpr int RU = ((E1B + 1) * d + G - 1) / G < RPL ? ((E1B + 1) * d + G - 1) / G : RPL;

    static constexpr int XS = pn_xs_doubles<D>(); /* exchange region per group, doubles */

    // update.jl:125-136: Sigma = R'R + K1 K1' (K1 = K R.R', D x O) re-factorised as cholesky(...).U, triangularize!([R ; K1']) if that
    // fails.  Same machinery as predict_cov!: own rows published, Gram matrix streamed, row-cyclic Cholesky.  Groups with
    // live = false keep their factor bit for bit.  Warp-convergent (full-mask shuffles).
    template <int O>
    static PN_HD void noise_refactor(double* sx, double (&R)[RPL][D], const double (&K1all)[D][O], bool live, int gl) {
        constexpr unsigned FULL = 0xffffffffu;
        pn_rows_store<G, RPL, D>(sx, R, gl);
        __syncwarp();
        double Mw[RPL][D];
#pragma unroll
        for (int lr = 0; lr < RPL; lr++) {
            double k1o[O]; /* own row of K1 */
#pragma unroll
            for (int a = 0; a < O; a++) {
                double v = 0.0;
#pragma unroll
                for (int g = 0; g < G; g++)
                    if (lr * G + g < D) v = (gl == g) ? K1all[lr * G + g < D ? lr * G + g : 0][a] : v;
                k1o[a] = v;
            }
#pragma unroll
            for (int j = 0; j < D; j++) {
                if (j < lr * G) continue;
                double v = 0.0;
#pragma unroll
                for (int a = 0; a < O; a++) v += k1o[a] * K1all[j][a];
                Mw[lr][j] = v;
            }
        }
        pn_gram_rows<G, RPL, D>(sx, Mw, gl);
        const bool okn = pn_chol_rows<G, RPL, D>(Mw, R, gl);
        const bool redo = live && !okn;
        if (__any_sync(FULL, redo)) {
            if (redo && gl == 0) {
                double E[O * D];
#pragma unroll
                for (int a = 0; a < O; a++)
#pragma unroll
                    for (int c = 0; c < D; c++) E[a * D + c] = K1all[c][a];
                pn_triangularize_serial_dense<D, O>(sx, E);
            }
            __syncwarp();
        }
        if (redo || !live) pn_rows_load<G, RPL, D>(sx, R, gl);
        __syncwarp();
    }

#if PN_DATA
    // One Kalman update of (mu, R) on the observation y = H x + eps, eps ~ N(0, R_obs), for every group with live = true
    // (DataUpdateCallback's affect!, dataupdate.jl:47-80: obs_cov = (R H')'(R H') + R_obs, update!(...; R = R_obs)); returns the
    // update's log-likelihood (pn_logpdf!, update.jl:140-148; +-Inf for a zero predicted covariance, :82-89).  H, R_obs and its
    // factor are compile-time tables of the NVRTC build (zeros fold away).  Warp-convergent.
    static PN_HD double data_update(double* sx, double (&mu)[D], double (&R)[RPL][D], const double* y, bool live, int gl) {
        constexpr int O = PN_DATA_O;
        double zd[O];
#pragma unroll
        for (int a = 0; a < O; a++) {
            double s = 0.0;
#pragma unroll
            for (int c = 0; c < D; c++)
                if (pn_data_H(a, c) != 0.0) s += pn_data_H(a, c) * mu[c];
            zd[a] = s - (live ? y[a] : 0.0);
        }
        double C2[RPL][O], S[O][O];
#pragma unroll
        for (int lr = 0; lr < RPL; lr++)
#pragma unroll
            for (int a = 0; a < O; a++) {
                double s = 0.0;
#pragma unroll
                for (int c = 0; c < D; c++)
                    if (pn_data_H(a, c) != 0.0) s += R[lr][c] * pn_data_H(a, c);
                C2[lr][a] = live ? s : 0.0;
            }
#pragma unroll
        for (int a = 0; a < O; a++)
#pragma unroll
            for (int b = 0; b < O; b++) {
                if (b < a) continue;
                double s = 0.0;
#pragma unroll
                for (int lr = 0; lr < RPL; lr++) s += C2[lr][a] * C2[lr][b];
                S[a][b] = pn_gsum<G>(s);
            }
        /* iszero(Sigma.R) (update.jl:82): all rows of R, not only R H' */
        double rr = 0.0;
#pragma unroll
        for (int lr = 0; lr < RPL; lr++)
#pragma unroll
            for (int c = 0; c < D; c++) rr += R[lr][c] * R[lr][c];
        rr = pn_gsum<G>(rr);
        const bool zero_cov = (rr == 0.0);
        const bool act = live && !zero_cov;
#pragma unroll
        for (int a = 0; a < O; a++)
#pragma unroll
            for (int b = 0; b < O; b++)
                if (b >= a) S[a][b] = act ? S[a][b] + pn_data_cov(a, b) : ((a == b) ? 1.0 : 0.0);
        double sinv[O];
        pn_chol<O>(S, sinv);
        double zt[O];
#pragma unroll
        for (int a = 0; a < O; a++) zt[a] = zd[a];
        pn_chol_solve<O>(S, sinv, zt);
        double quad = 0.0, logdet = 0.0;
#pragma unroll
        for (int a = 0; a < O; a++) {
            quad += zd[a] * zt[a];
            logdet += log(S[a][a]);
        }
        double ll = 0.0;
        if (live) {
            bool zz = true;
#pragma unroll
            for (int a = 0; a < O; a++) zz = zz && (zd[a] == 0.0);
            ll = zero_cov ? (zz ? PN_INF : -PN_INF) : (-0.5 * quad - 0.5 * (double)O * 1.8378770664093453 - logdet);
        }
        double V[RPL][O];
#pragma unroll
        for (int lr = 0; lr < RPL; lr++) {
            double v[O];
#pragma unroll
            for (int a = 0; a < O; a++) v[a] = act ? C2[lr][a] : 0.0;
            pn_chol_solve<O>(S, sinv, v);
#pragma unroll
            for (int a = 0; a < O; a++) V[lr][a] = v[a];
        }
        double K1all[D][O];
#pragma unroll
        for (int c = 0; c < D; c++) {
            double Kc[O];
#pragma unroll
            for (int a = 0; a < O; a++) {
                double s = 0.0;
#pragma unroll
                for (int lr = 0; lr < RPL; lr++) s += R[lr][c] * V[lr][a];
                Kc[a] = pn_gsum<G>(s);
            }
            double s = mu[c];
#pragma unroll
            for (int a = 0; a < O; a++) s -= Kc[a] * zd[a];
            mu[c] = act ? s : mu[c];
#pragma unroll
            for (int lr = 0; lr < RPL; lr++) {
                double r = R[lr][c];
#pragma unroll
                for (int a = 0; a < O; a++) r -= C2[lr][a] * Kc[a];
                R[lr][c] = act ? r : R[lr][c];
            }
#pragma unroll
            for (int a = 0; a < O; a++) {
                double k1 = 0.0;
#pragma unroll
                for (int b = 0; b < O; b++)
                    if (pn_data_rt(a, b) != 0.0) k1 += Kc[b] * pn_data_rt(a, b);
                K1all[c][a] = k1;
            }
        }
        noise_refactor<O>(sx, R, K1all, act, gl);
        return ll;
    }
#endif

    static __device__ void run(const PnParams& P, double* smem) {
        constexpr unsigned FULL = 0xffffffffu;
        const int lane = threadIdx.x & 31;
        const int gl = lane % G;
        const int leader = lane - gl;
        double* const sx = smem + (size_t)(threadIdx.x / G) * XS; /* this group's exchange region (PN_PREDICT_CHOL) */

        /* bottom-block row bookkeeping: slot lr holds stack row D + rb, rb = lr*G + gl = m*d + i */
        int mB[RPL], iB[RPL];
#pragma unroll
        for (int lr = 0; lr < RPL; lr++) {
            const int rb = lr * G + gl;
            mB[lr] = rb / d;
            iB[lr] = rb - mB[lr] * d;
        }

        double mu[D];          /* filter mean, derivative-major, replicated in the group */
        double R[RPL][D];      /* own rows of the right factor */
        double pr[NPA];
        double uprev[d];
        double t = 0.0, dt = 0.0, dtcache = 0.0, qold_pb2 = 1.0, loglik = 0.0, gdiff = 0.0, ldiff = 0.0;
        long long traj = -1, naccept = 0, nreject = 0, nf = 0, iter = 0, tstop_idx = 0, save_pos = 0;
        int retcode = PN_RET_SUCCESS;
        bool have = false, queue_empty = false;
#if PN_DATA
        long long data_idx = 0; /* next observation (the data times are tstops, so every one of them is hit exactly) */
        double data_ll = 0.0;   /* DataUpdateLogLikelihood.ll */
#endif
#pragma unroll
        for (int lr = 0; lr < RPL; lr++)
#pragma unroll
            for (int c = 0; c < D; c++) R[lr][c] = 0.0;
#pragma unroll
        for (int c = 0; c < D; c++) mu[c] = 0.0;

        /* The warp runs this loop convergently (full-mask shuffles); groups differ only in predicates. */
        for (;;) {
            /* ============ (A) groups without work fetch the next trajectory ============ */
            const bool need = !have && !queue_empty;
            if (__any_sync(FULL, need)) {
                unsigned long long idx = 0;
                if (need && gl == 0) idx = atomicAdd(P.work_counter, 1ULL);
                idx = __shfl_sync(FULL, idx, leader);
                if (need) {
                    if ((long long)idx * P.traj_stride >= P.n_traj) {
                        queue_empty = true;
                    } else {
                        traj = (long long)idx * P.traj_stride;
                        {
                            /* initial state: computed for the whole ensemble by pn_init_kernel before this kernel started
                               (TaylorModeInit + initial step; one thread per trajectory, off the persistent loop's
                               critical path -- in line it cost ~3 % of the ROBER run and a local-memory frame) */
#pragma unroll
                            for (int c = 0; c < D; c++) mu[c] = P.init_mu[traj * D + c];
#pragma unroll
                            for (int i = 0; i < NPA; i++) pr[i] = (i < Prob::n_p) ? P.p[traj * Prob::n_p + i] : 0.0;
#pragma unroll
                            for (int i = 0; i < d; i++) uprev[i] = mu[pn_solcol<d>(i)];
                            dt = P.init_dt[traj];
                            nf = Q + ((ADAPTIVE && !(P.dt0 > 0.0) && !PN_MASS) ? 2 : 0);
                            t = P.t0;
                        }
#pragma unroll
                        for (int lr = 0; lr < RPL; lr++)
#pragma unroll
                            for (int c = 0; c < D; c++) R[lr][c] = 0.0;
#if PN_R_IN_SMEM && PN_PREDICT_CHOL
                        pn_rows_store<G, RPL, D>(sx, R, gl); /* Sigma_0.R = 0 */
#endif
                        naccept = nreject = iter = tstop_idx = 0;
#if PN_DATA
                        data_idx = 0;
                        data_ll = 0.0;
                        while (data_idx < P.n_data && P.data_t[data_idx] <= P.t0) data_idx++; /* PresetTimeCallback: t0 is not a step end */
#endif
                        loglik = 0.0;
                        gdiff = 0.0;
                        ldiff = 0.0;
                        qold_pb2 = P.qoldinit_pb2;
                        retcode = PN_RET_SUCCESS;
                        dtcache = dt;
                        have = true;
                        if (SAVE >= 1) {
                            save_pos = P.step_offsets ? P.step_offsets[traj] : 0;
                            if (P.step_offsets && gl == 0) {
                                P.s_t[save_pos] = t;
                                P.s_diffusion[save_pos] = 0.0;
#pragma unroll
                                for (int i = 0; i < DU; i++) P.s_u_mean[save_pos * DU + i] = mu[pn_solcol<d>(i)];
#pragma unroll
                                for (int i = 0; i < DU * DU; i++) P.s_u_cov[save_pos * DU * DU + i] = 0.0;
                                if (SAVE >= 2) {
#pragma unroll
                                    for (int c = 0; c < D; c++) P.s_x_mean[save_pos * D + c] = mu[c];
                                }
                            }
                            if (SAVE >= 2 && P.step_offsets) {
#pragma unroll
                                for (int lr = 0; lr < RPL; lr++) {
                                    const int row = lr * G + gl;
                                    if (row < D) {
#pragma unroll
                                        for (int c = 0; c < D; c++) P.s_x_covfac[(save_pos * D + row) * D + c] = 0.0;
                                    }
                                }
                            }
                            save_pos += 1;
                        }
                    }
                }
            }
#ifdef PN_CTA_SYNC
            /* keep the warps of a CTA in step: the two warps that share a scheduler then fetch the same
               instruction lines (the unrolled loop body is ~100 KB, far beyond the L0 instruction cache) */
            if (__syncthreads_or(have ? 1 : 0) == 0) break;
#else
            if (__all_sync(FULL, !have)) break;
#endif

            /* ============ (B) phase 1: attempt steps until one is accepted (no cross-lane traffic) ============ */
            bool finished = !have || !(t < P.t1);
            bool accept = false;
            double h = 0.0, tstop = P.t1, EEst = 0.0, qq = 1.0, ediff = 0.0, lEE = 0.0;
            double hp[n], PIv[n], mup[D], z[d], J[d * d];
#if PN_SECOND
            double Jv[d * d]; /* d f1 / d du; J holds d f1 / d u */
#pragma unroll
            for (int i = 0; i < d * d; i++) Jv[i] = 0.0;
#endif
#if PN_IOUP
            double tq[n], Uh[n][n]; /* last column of the 1-d transition matrix; upper factor of the 1-d process noise */
#pragma unroll
            for (int k = 0; k < n; k++) {
                tq[k] = (k == Q) ? 1.0 : 0.0;
#pragma unroll
                for (int c = 0; c < n; c++) Uh[k][c] = 0.0;
            }
#endif
#pragma unroll
            for (int k = 0; k < n; k++) hp[k] = (k == 0) ? 1.0 : 0.0;
#pragma unroll
            for (int c = 0; c < n; c++) PIv[c] = 0.0;
#pragma unroll
            for (int c = 0; c < D; c++) mup[c] = mu[c];
#pragma unroll
            for (int i = 0; i < d; i++) z[i] = 0.0;
#pragma unroll
            for (int i = 0; i < d * d; i++) J[i] = 0.0;

            /* ONE attempt per iteration of the warp's loop: a group whose step is rejected sits out phase 2 (identity
               step) and retries in the next iteration.  (Retrying inside this phase kept the other groups of the warp
               waiting and let the warps of the CTA drift apart in the ~100 KB unrolled body: measured slower, git history.) */
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
                /* hp[k] = h^k/k! ; PIv[c] = h^(q-c+1/2)/(q-c)!  (preconditioning.jl:45-54) */
                hp[0] = 1.0;
#pragma unroll
#if PN_FAST_SCALAR
                for (int k = 1; k < n; k++) hp[k] = hp[k - 1] * h * (1.0 / (double)k);
                const double sqh = h * pn_rsqrt(h);
#else
                for (int k = 1; k < n; k++) hp[k] = hp[k - 1] * h / (double)k;
                const double sqh = sqrt(h);
#endif
#pragma unroll
                for (int c = 0; c < n; c++) PIv[c] = hp[Q - c] * sqh;
#if PN_IOUP
                if (!pn_ioup_step<Q, G>(P.rate, P.glx, P.glw, h, gl, tq, Uh)) {
                    retcode = PN_RET_UNSTABLE;
                    finished = true;
                    break;
                }
#endif
                /* predict_mean!: mu^- = (T (x) I) mu, T[j][k] = h^(k-j)/(k-j)!  (IOUP: last column tq) */
#pragma unroll
                for (int j = 0; j < n; j++)
#pragma unroll
                    for (int i = 0; i < d; i++) {
#if PN_IOUP
                        double s = (j == Q) ? tq[Q] * mu[j * d + i] : mu[j * d + i];
#pragma unroll
                        for (int k = 0; k < n; k++)
                            if (k > j) s += ((k == Q) ? tq[j] : hp[k > j ? k - j : 0]) * mu[k * d + i];
#else
                        double s = mu[j * d + i];
#pragma unroll
                        for (int k = 0; k < n; k++)
                            if (k > j) s += hp[k > j ? k - j : 0] * mu[k * d + i];
#endif
                        mup[j * d + i] = s;
                    }
                /* measurement: z = E1 mu^- - f(E0 mu^-), J = df/du */
#if PN_SECOND
                double fu[DP], x2[DP];
#pragma unroll
                for (int i = 0; i < d; i++) {
                    x2[i] = mup[d + i]; /* (du, u) = (E1 mu^-, E0 mu^-) */
                    x2[d + i] = mup[i];
                }
                Prob::template f<double>(fu, x2, pr, tnew);
#if !PN_EK0_DENSE
                {
                    double J2[DP * DP];
                    Prob::jac(J2, x2, pr, tnew);
#pragma unroll
                    for (int a = 0; a < d; a++)
#pragma unroll
                        for (int i = 0; i < d; i++) {
                            Jv[a + d * i] = J2[a + DP * i];
                            J[a + d * i] = J2[a + DP * (d + i)];
                        }
                }
#endif
#else
                double fu[d];
                Prob::template f<double>(fu, mup, pr, tnew);
#if !PN_EK0_DENSE
                Prob::jac(J, mup, pr, tnew);
#endif
#endif
                nf += 1;
#if PN_MASS
#pragma unroll
                for (int a = 0; a < d; a++) { /* z = M E1 mu^- - f (measurement_models.jl:18) */
                    double sm = 0.0;
#pragma unroll
                    for (int i = 0; i < d; i++)
                        if (pn_mass(a, i) != 0.0) sm += pn_mass(a, i) * mup[d + i];
                    z[a] = sm - fu[a];
                }
#else
#pragma unroll
                for (int i = 0; i < d; i++) z[i] = mup[E1B * d + i] - fu[i];
#endif

                if (ADAPTIVE || DYNAMIC) {
                    /* C[(m,i),a] = (Qh.R H')[(m,i),a] = -LQ[m][0] J[a][i] + LQ[m][1] (i==a) ; W = C'C */
                    double W[d][d];
#pragma unroll
                    for (int a = 0; a < d; a++)
#pragma unroll
                        for (int b = 0; b < d; b++) W[a][b] = 0.0;
#if PN_FAST_SCALAR && !PN_IOUP && !PN_MASS && !PN_SECOND
                    /* closed form of the same W = H Qh H' with Qh = (P^-1 Qb P^-1) (x) I_d, H = [-J | I | 0]:
                       W = qh00 J J' - qh01 (J + J') + qh11 I  (3x fewer multiply-adds than summing C'C over the n derivative rows) */
                    {
                        const double w00 = P.Qb[0] * PIv[0] * PIv[0], w01 = P.Qb[1] * PIv[0] * PIv[1], w11 = P.Qb[n + 1] * PIv[1] * PIv[1];
#pragma unroll
                        for (int a = 0; a < d; a++)
#pragma unroll
                            for (int b = 0; b < d; b++) {
                                if (b < a) continue;
                                double jj = 0.0;
#pragma unroll
                                for (int i = 0; i < d; i++) jj += J[a + d * i] * J[b + d * i];
                                W[a][b] = w00 * jj - w01 * (J[a + d * b] + J[b + d * a]) + ((a == b) ? w11 : 0.0);
                            }
                    }
#else
#pragma unroll
                    for (int m = 0; m < n; m++) {
#if PN_IOUP
                        /* Qh.R = Uh P^-1 (upper): column 0 lives in row 0, column 1 in rows 0 and 1 */
                        if (m > 1) continue;
                        const double l0 = (m == 0) ? Uh[0][0] * PIv[0] : 0.0;
                        const double l1 = Uh[m][1] * PIv[1];
#else
                        const double l0 = P.L[m * n + 0] * PIv[0];
                        const double l1 = (m >= 1) ? P.L[m * n + 1] * PIv[1] : 0.0;
#if PN_SECOND
                        const double l2 = (m >= 2) ? P.L[m * n + 2] * PIv[2] : 0.0;
#endif
#endif
#pragma unroll
                        for (int i = 0; i < d; i++) {
                            double c[d];
#pragma unroll
#if PN_MASS
                            for (int a = 0; a < d; a++) c[a] = -l0 * J[a + d * i] + ((pn_mass(a, i) != 0.0) ? l1 * pn_mass(a, i) : 0.0);
#elif PN_SECOND
                            for (int a = 0; a < d; a++) c[a] = -l0 * J[a + d * i] - l1 * Jv[a + d * i] + ((i == a) ? l2 : 0.0);
#else
                            for (int a = 0; a < d; a++) c[a] = -l0 * J[a + d * i] + ((i == a) ? l1 : 0.0);
#endif
#pragma unroll
                            for (int a = 0; a < d; a++)
#pragma unroll
                                for (int b = 0; b < d; b++)
                                    if (b >= a) W[a][b] += c[a] * c[b];
                        }
                    }
#endif
                    double wdiag[d];
#pragma unroll
                    for (int a = 0; a < d; a++) wdiag[a] = W[a][a];
                    double zt[d];
#pragma unroll
                    for (int a = 0; a < d; a++) zt[a] = z[a];
                    double winv[d];
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
                        /* EEst^2 = mean_i (h sqrt(s2 W_ii) / sc_i)^2 : one rsqrt instead of d sqrt + d div */
                        double e2 = 0.0;
#pragma unroll
                        for (int i = 0; i < d; i++) {
                            const double isc = pn_rcp(P.abstol + fmax(fabs(mup[pn_solcol<d>(i)]), fabs(uprev[i])) * P.reltol);
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
                    /* stepsize_controller!(::PIController) (OrdinaryDiffEqCore) */
                    if (EEst == 0.0) {
                        qq = P.inv_qmax;
                    } else {
                        /* EEst^beta1 / qold^beta2 as exp(beta * log x); log(EEst) is reused for qold^beta2 of the
                           next step (qold = max(EEst, qoldinit)), so a step costs one log and two exp */
                        lEE = log(EEst);
                        q11 = exp(P.beta1 * lEE);
                        qq = q11 * pn_rcp(qold_pb2 * P.gamma);
                        qq = fmax(P.inv_qmax, fmin(P.inv_qmin, qq));
                    }
                    if (!accept) { /* step_reject_controller! */
                        nreject++;
#if PN_FAST_SCALAR
                        dt = dt * pn_rcp(fmin(P.inv_qmin, q11 * P.inv_gamma));
#else
                        dt = dt / fmin(P.inv_qmin, q11 / P.gamma);
#endif
                    }
                }
                ediff = DYNAMIC ? ldiff : P.initial_diffusion;
                if (P.forced_diffusion && naccept < P.n_forced) ediff = P.forced_diffusion[naccept];
            }
            if (!accept) {
                /* identity step for groups that have nothing to commit: h = 0, no process noise, K = 0 */
#pragma unroll
                for (int k = 1; k < n; k++) hp[k] = 0.0;
#if PN_IOUP
#pragma unroll
                for (int k = 0; k < n; k++) tq[k] = (k == Q) ? 1.0 : 0.0;
#endif
                ediff = 0.0;
            }
            /* from here on one mean vector is live: the predicted mean of an accepted attempt, else the filtered mean (the
               update below subtracts K z with K = 0 for groups that commit nothing) */
#pragma unroll
            for (int c = 0; c < D; c++) mu[c] = accept ? mup[c] : mu[c];
            double t_next = t;
            if (accept) {
                /* loopfooter! accept branch (OrdinaryDiffEqCore), scalar part: everything that does not depend on
                   the covariance update is settled here, so that the controller's temporaries are dead during phase 2 */
                double dtnew = dt;
                if (ADAPTIVE) {
                    if (P.qsteady_min <= qq && qq <= P.qsteady_max) qq = 1.0;
                    qold_pb2 = (EEst > P.qoldinit) ? exp(P.beta2 * lEE) : P.qoldinit_pb2; /* qold = max(EEst, qoldinit), kept as qold^beta2 */
#if PN_FAST_SCALAR
                    dtnew = dt * pn_rcp(qq);
#else
                    dtnew = dt / qq;
#endif
                }
                const double ttmp = t + h;
                const double ref = fmax(fabs(t), fabs(tstop));
#if PN_FAST_SCALAR
                const double epsref = pn_ulp_above(ref);
#else
                const double epsref = nextafter(ref, 1e300) - ref;
#endif
                t_next = (fabs(ttmp - tstop) < 100.0 * epsref) ? tstop : ttmp;
                dt = dtnew;
            }
            t = t_next;

            /* ============ (C) phase 2: covariance predict + update (warp-convergent) ============ */
            /* Phase 2 runs in EVERY iteration, also for a warp in which no group accepted: a group without an accepted
               step takes the identity step (R <- qr(R), same R'R), and it does so after each of ITS OWN rejections --
               never depending on what the other trajectories of the warp did.  That keeps every trajectory's arithmetic
               a function of its own history only (bit-reproducible runs; the counting and the save pass agree). */
            {
                double B[RPL][D];
#if PN_R_IN_SMEM && PN_PREDICT_CHOL
                __syncwarp();
                pn_rows_load<G, RPL, D>(sx, R, gl); /* the filtered factor of the previous iteration */
#endif
                /* top block: X = R Ah'  (row-local; ascending j is safe in place) */
#pragma unroll
                for (int lr = 0; lr < RPL; lr++)
#pragma unroll
                    for (int i = 0; i < d; i++)
#pragma unroll
                        for (int j = 0; j < n; j++) {
#if PN_IOUP
                            double s = (j == Q) ? tq[Q] * R[lr][j * d + i] : R[lr][j * d + i];
#pragma unroll
                            for (int k = 0; k < n; k++)
                                if (k > j) s += ((k == Q) ? tq[j] : hp[k > j ? k - j : 0]) * R[lr][k * d + i];
#else
                            double s = R[lr][j * d + i];
#pragma unroll
                            for (int k = 0; k < n; k++)
                                if (k > j) s += hp[k > j ? k - j : 0] * R[lr][k * d + i];
#endif
                            R[lr][j * d + i] = s;
                        }
#if PN_PREDICT_CHOL
                /* predict_cov! (predict.jl:75-92): M = X'X + s2 Qh, R^- = cholesky(M).U; triangularize! if that fails.
                   Groups without an accepted step keep their factor bit for bit (h = 0 made X = R above). */
                {
                    (void)B;
                    /* (1) publish the own rows of X */
                    pn_rows_store<G, RPL, D>(sx, R, gl);
                    __syncwarp();
                    /* (2) accumulators of the own rows start from s2 Qh = s2 (P^-1 Qb P^-1) (x) I_d, Qb = L'L (iwp.jl:74-108):
                       row (m,i) has s2 PI[m] Qb[m][c] PI[c] at column (c,i) */
                    double Mw[RPL][D];
                    if constexpr (G <= 4) {
                        /* few lanes: the (derivative, component) pair of an own row is one of G compile-time alternatives, so the
                           row of s2 Qh is picked with G selects per entry from the n(n+1)/2 distinct values qh[m][c] */
                        double qh[n][n];
#pragma unroll
                        for (int m = 0; m < n; m++)
#pragma unroll
                            for (int c = 0; c < n; c++) {
                                if (c < m) continue;
#if PN_IOUP
                                double qmc = 0.0; /* Qb = Uh'Uh */
#pragma unroll
                                for (int r = 0; r < n; r++)
                                    if (r <= m) qmc += Uh[r][m] * Uh[r][c];
#else
                                const double qmc = P.Qb[m * n + c];
#endif
                                qh[m][c] = qmc * (PIv[m] * ediff) * PIv[c];
                            }
#pragma unroll
                        for (int lr = 0; lr < RPL; lr++)
#pragma unroll
                            for (int j = 0; j < D; j++) {
                                if (j < lr * G) continue; /* never read (pn_chol_rows) */
                                double v = 0.0;
#pragma unroll
                                for (int g = 0; g < G; g++) {
                                    const int rb = lr * G + g;
                                    if (rb < D && (rb % d) == (j % d)) {
                                        const int m = rb / d, c = j / d;
                                        v = (gl == g) ? qh[m < c ? m : c][m < c ? c : m] : v;
                                    }
                                }
                                Mw[lr][j] = v;
                            }
                    } else {
#pragma unroll
                    for (int lr = 0; lr < RPL; lr++) {
                        double pim = 0.0; /* s2 PI[m] of the own row (m, i) */
#pragma unroll
                        for (int m = 0; m < n; m++) pim = (mB[lr] == m) ? PIv[m] * ediff : pim;
#pragma unroll
                        for (int c = 0; c < n; c++) {
#if PN_IOUP
                            double qsel = 0.0; /* Qb = Uh'Uh */
#pragma unroll
                            for (int m = 0; m < n; m++) {
                                double qmc = 0.0;
#pragma unroll
                                for (int r = 0; r < n; r++)
                                    if (r <= m && r <= c) qmc += Uh[r][m] * Uh[r][c];
                                qsel = (mB[lr] == m) ? qmc : qsel;
                            }
#else
                            const double qsel = (mB[lr] < n) ? P.Qb[mB[lr] * n + c] : 0.0; /* constant-bank load, run-time row */
#endif
                            const double val = qsel * pim * PIv[c];
#pragma unroll
                            for (int i = 0; i < d; i++) Mw[lr][c * d + i] = (iB[lr] == i) ? val : 0.0;
                        }
                    }
                    }
                    /* (3) Gram matrix: stream every row of X from the exchange region */
                    pn_gram_rows<G, RPL, D>(sx, Mw, gl);
                    /* (4) factorise; idle groups and zero diffusion (predict.jl:71-74 keeps X, which the update below needs
                       in triangular form) do not count as failures of the Cholesky route */
                    bool ok = pn_chol_rows<G, RPL, D>(Mw, R, gl);
#ifdef PN_FORCE_CHOL_FAIL
                    ok = ok && ((naccept % PN_FORCE_CHOL_FAIL) != 0); /* test hook: exercise the fallback on every N-th step */
#endif
                    const bool redo = accept && (!ok || ediff == 0.0);
                    if (__any_sync(FULL, redo)) { /* predict.jl:90: triangularize!(stack) -- cold path, one lane per trajectory */
                        if (redo && gl == 0) {
                            double bf[n * n]; /* sqrt(s2) (U P^-1): the noise block's n x n factor */
                            const double sd = (ediff == 1.0) ? 1.0 : sqrt(ediff);
#pragma unroll
                            for (int m = 0; m < n; m++)
#pragma unroll
                                for (int c = 0; c < n; c++) /* (unrolled: a run-time index would demote PIv to local memory) */
#if PN_IOUP
                                    bf[m * n + c] = (c >= m) ? Uh[m][c] * PIv[c] * sd : 0.0;
#else
                                    bf[m * n + c] = (c >= m) ? P.U[m * n + c] * PIv[c] * sd : 0.0;
#endif
                            pn_triangularize_serial<D, d, n>(sx, bf);
                        }
                        __syncwarp();
                    }
                    if (redo || !accept) /* take the factor from the exchange region: the fallback's result / the untouched X = R */
                        pn_rows_load<G, RPL, D>(sx, R, gl);
                    __syncwarp(); /* the region is rewritten in the next iteration */
                }
#else
                /* bottom block: sqrt(s2) (U P^-1) (x) I_d with U'U = L'L upper triangular (pn_params.h): row (m,i) has
                   U[m][c] PI[c] at column (c,i), c >= m, so the block is upper triangular and its rows join the
                   Householder sweep one slot at a time */
                {
                    const double sd = (ediff == 1.0) ? 1.0 : sqrt(ediff);
#pragma unroll
                    for (int lr = 0; lr < RPL; lr++)
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
                            for (int i = 0; i < d; i++) B[lr][c * d + i] = (iB[lr] == i) ? val : 0.0;
                        }
                }
#ifndef PN_QR_JC
#define PN_QR_JC D
#endif
                pn_qr_rows<G, RPL, D, D, true, PN_QR_JC>(R, B, gl, nullptr);

#endif /* PN_PREDICT_CHOL */
                /* C2 = R^- H' (row-local; rows >= 2d are exactly zero, see RU), S = C2'C2 */
                double C2[RU][d];
                double S[d][d];
#pragma unroll
                for (int lr = 0; lr < RU; lr++)
#pragma unroll
                    for (int a = 0; a < d; a++) {
#if PN_MASS
                        double s = 0.0; /* H = [-J | M | 0] */
#pragma unroll
                        for (int i = 0; i < d; i++)
                            if (pn_mass(a, i) != 0.0) s += R[lr][d + i] * pn_mass(a, i);
#else
                        double s = R[lr][E1B * d + a];
#endif
#pragma unroll
                        for (int i = 0; i < d; i++) s -= R[lr][i] * J[a + d * i];
#if PN_SECOND
#pragma unroll
                        for (int i = 0; i < d; i++) s -= R[lr][d + i] * Jv[a + d * i];
#endif
                        C2[lr][a] = accept ? s : 0.0;
                    }
#pragma unroll
                for (int a = 0; a < d; a++)
#pragma unroll
                    for (int b = 0; b < d; b++) {
                        if (b < a) continue;
                        double s = 0.0;
#pragma unroll
                        for (int lr = 0; lr < RU; lr++) s += C2[lr][a] * C2[lr][b];
                        S[a][b] = pn_gsum<G>(s);
                    }
                double tr = 0.0;
#pragma unroll
                for (int a = 0; a < d; a++) tr += S[a][a];
                const bool passthrough = (tr == 0.0); /* Sigma^- H' == 0 (update.jl:82-89) or idle group */
                if (passthrough) {
#pragma unroll
                    for (int a = 0; a < d; a++) S[a][a] = 1.0;
                }
#if PN_OBSN
#pragma unroll
                for (int a = 0; a < d; a++) /* compute_measurement_covariance!: + R.R'R.R (perform_step.jl:185-187) */
#pragma unroll
                    for (int b = 0; b < d; b++)
                        if (b >= a && pn_obsn_cov(a, b) != 0.0) S[a][b] += passthrough ? 0.0 : pn_obsn_cov(a, b);
                double K1all[D][d]; /* K R.R' (update.jl:128) */
#endif
                double sinv[d];
                const bool okS = pn_chol<d>(S, sinv);
                double zt[d];
#pragma unroll
                for (int a = 0; a < d; a++) zt[a] = z[a];
                pn_chol_solve<d>(S, sinv, zt);
                double quad = 0.0, logdet = 0.0, dprod = 1.0;
#pragma unroll
                for (int a = 0; a < d; a++) {
                    quad += z[a] * zt[a];
                    dprod *= S[a][a];
                    if ((a & 3) == 3 || a == d - 1) { /* one log per 4 Cholesky pivots (the lazy accumulator pn_logacc_mul
                                                         costs three more registers here: measured -3 % on ROBER) */
                        logdet += log(dprod);
                        dprod = 1.0;
                    }
                }
                if (accept) {
                    if (!okS) retcode = PN_RET_UNSTABLE;
                    if (passthrough) {
                        bool zz = true;
#pragma unroll
                        for (int a = 0; a < d; a++) zz = zz && (z[a] == 0.0);
                        loglik += zz ? PN_INF : -PN_INF;
                    } else {
                        loglik += -0.5 * quad - 0.5 * (double)d * 1.8378770664093453 - logdet;
                        if (!DYNAMIC) { /* estimate_global_diffusion(::FixedDiffusion), calibration.jl:49-66 */
                            const double inc = quad * (1.0 / (double)d);
                            gdiff = (naccept == 0) ? inc : gdiff + (inc - gdiff) / (double)naccept;
                        }
                    }
                }
                /* V = C2 S^-1 (row-local) ; K = R^-' V all-reduced ; mu = mu^- - K z ; R+ = R^- - C2 K' */
                double V[RU][d];
#pragma unroll
                for (int lr = 0; lr < RU; lr++) {
                    double v[d];
#pragma unroll
                    for (int a = 0; a < d; a++) v[a] = C2[lr][a];
                    pn_chol_solve<d>(S, sinv, v);
#pragma unroll
                    for (int a = 0; a < d; a++) V[lr][a] = v[a];
                }
#pragma unroll
                for (int c = 0; c < D; c++) {
                    double Kc[d];
#pragma unroll
                    for (int a = 0; a < d; a++) {
                        double s = 0.0;
#pragma unroll
                        for (int lr = 0; lr < RU; lr++) s += R[lr][c] * V[lr][a];
                        Kc[a] = pn_gsum<G>(s);
                    }
                    double s = mu[c];
#pragma unroll
                    for (int a = 0; a < d; a++) s -= Kc[a] * z[a];
                    mu[c] = s;
#if PN_OBSN
#pragma unroll
                    for (int a = 0; a < d; a++) {
                        double k1 = 0.0;
#pragma unroll
                        for (int b = 0; b < d; b++)
                            if (pn_obsn_rt(a, b) != 0.0) k1 += Kc[b] * pn_obsn_rt(a, b);
                        K1all[c][a] = k1;
                    }
#endif
#pragma unroll
                    for (int lr = 0; lr < RU; lr++) {
                        double r = R[lr][c];
#pragma unroll
                        for (int a = 0; a < d; a++) r -= C2[lr][a] * Kc[a];
                        R[lr][c] = r;
                    }
                }
#if PN_OBSN
                /* update.jl:125-136: + K R K'.  Groups that commit nothing (or passed the prediction through, update.jl:82-89) keep their
                   factor bit for bit. */
                noise_refactor<d>(sx, R, K1all, accept && !passthrough, gl);
#endif
#if PN_R_IN_SMEM && PN_PREDICT_CHOL
                if (accept) pn_rows_store<G, RPL, D>(sx, R, gl); /* x_filt.Sigma.R for the next iteration (idle groups: the region still holds it) */
#endif
                if (accept) {
                    /* loopfooter! accept branch, the part that needs the updated mean (the scalars were settled before phase 2) */
                    naccept++;
                    bool nanu = false;
#pragma unroll
                    for (int i = 0; i < d; i++) {
                        uprev[i] = mu[pn_solcol<d>(i)];
                        nanu = nanu || (mu[i] != mu[i]);
                    }
                    if (nanu) retcode = PN_RET_UNSTABLE;
                    if (retcode != PN_RET_SUCCESS) finished = true;
                    if (!(t < P.t1)) finished = true;
                }
#if PN_DATA
                {
                    /* PresetTimeCallback at the data times (dataupdate.jl:83): condition the filtered state on the observation before it
                       is saved (callbacks run before savevalues!); uprev keeps the value from before the update (update_uprev! copies
                       integ.u, which the callback does not touch) */
                    const bool du = accept && data_idx < P.n_data && t == P.data_t[data_idx];
                    if (__any_sync(FULL, du)) data_ll += data_update(sx, mu, R, P.data_u + (du ? data_idx : 0) * PN_DATA_O, du, gl);
                    if (du) data_idx++;
                }
#endif
                if (SAVE >= 1) {
                    /* never write past this trajectory's CSR segment (the host-computed offsets are exact; belt and braces) */
                    const bool wr = accept && P.step_offsets && save_pos < P.step_offsets[traj + 1];
                    double cv[DU * DU];
#pragma unroll
                    for (int a = 0; a < DU; a++)
#pragma unroll
                        for (int b = 0; b < DU; b++) {
                        if (b < a) continue;
                            double s = 0.0;
#pragma unroll
                            for (int lr = 0; lr < RPL; lr++) s += R[lr][pn_solcol<d>(a)] * R[lr][pn_solcol<d>(b)];
                            s = pn_gsum<G>(s);
                            cv[a * DU + b] = s;
                            cv[b * DU + a] = s;
                        }
                    if (wr && gl == 0) {
                        P.s_t[save_pos] = t;
                        P.s_diffusion[save_pos - 1] = (SAVE >= 2) ? ediff : ((ADAPTIVE || DYNAMIC) ? ldiff : P.initial_diffusion);
#pragma unroll
                        for (int i = 0; i < DU; i++) P.s_u_mean[save_pos * DU + i] = mu[pn_solcol<d>(i)];
#pragma unroll
                        for (int i = 0; i < DU * DU; i++) P.s_u_cov[save_pos * DU * DU + i] = cv[i];
                        if (SAVE >= 2) { /* what the smoother sweep needs: x_filt[k+1] and the step that produced it */
                            P.s_h[save_pos - 1] = h;
#pragma unroll
                            for (int c = 0; c < D; c++) P.s_x_mean[save_pos * D + c] = mu[c];
                        }
                    }
                    if (SAVE >= 2 && wr) {
#pragma unroll
                        for (int lr = 0; lr < RPL; lr++) {
                            const int row = lr * G + gl;
                            if (row < D) {
                                double* dst = P.s_x_covfac + (save_pos * D + row) * D;
                                if (D % 2 == 0) { /* 16-byte stores: the row start is a multiple of D doubles */
#pragma unroll
                                    for (int c = 0; c < D; c += 2)
                                        *reinterpret_cast<double2*>(dst + c) = make_double2(R[lr][c], R[lr][(c + 1 < D) ? c + 1 : c]);
                                } else {
#pragma unroll
                                    for (int c = 0; c < D; c++) dst[c] = R[lr][c];
                                }
                            }
                        }
                    }
                    if (accept) save_pos += 1;
                }
            }

            /* ============ (D) postamble of finished trajectories ============ */
            const bool fin = have && finished;
            if (__any_sync(FULL, fin)) {
                double sc = 1.0, fd = ldiff;
                if (!DYNAMIC) {
                    if (P.calibrate && naccept > 0) {
                        sc = gdiff;
                        fd = gdiff * P.initial_diffusion;
                    } else {
                        fd = P.initial_diffusion;
                    }
                }
                double cv[DU * DU];
#pragma unroll
                for (int a = 0; a < DU; a++)
#pragma unroll
                    for (int b = 0; b < DU; b++) {
                        if (b < a) continue;
                        double s = 0.0;
#pragma unroll
                        for (int lr = 0; lr < RPL; lr++) s += R[lr][pn_solcol<d>(a)] * R[lr][pn_solcol<d>(b)];
                        s = pn_gsum<G>(s) * sc;
                        cv[a * DU + b] = s;
                        cv[b * DU + a] = s;
                    }
                if (fin) {
                    if (gl == 0) {
                        P.t_end[traj] = t;
#pragma unroll
                        for (int i = 0; i < DU; i++) P.u_mean[traj * DU + i] = mu[pn_solcol<d>(i)];
#pragma unroll
                        for (int i = 0; i < DU * DU; i++) P.u_cov[traj * DU * DU + i] = cv[i];
                        P.loglik[traj] = loglik;
#if PN_DATA
                        if (P.data_loglik) P.data_loglik[traj] = data_ll;
#endif
                        P.diffusion[traj] = fd;
                        P.dt_last[traj] = dt;
                        P.naccept[traj] = naccept;
                        P.nreject[traj] = nreject;
                        P.nf[traj] = nf;
                        P.retcode[traj] = retcode;
                        if (P.x_mean) {
#pragma unroll
                            for (int c = 0; c < D; c++) P.x_mean[traj * D + c] = mu[c];
                        }
                    }
                    if (P.x_covfac) {
                        const double ss = sqrt(sc);
#pragma unroll
                        for (int lr = 0; lr < RPL; lr++) {
                            const int row = lr * G + gl;
                            if (row < D) {
#pragma unroll
                                for (int c = 0; c < D; c++) P.x_covfac[(traj * D + row) * D + c] = R[lr][c] * ss;
                            }
                        }
                    }
                    have = false;
                }
            }
        }
    }

};

// Initial state + initial step of every trajectory, one thread each (launched before the persistent step kernel):
// exact initial derivatives (TaylorModeInit, autodiffinit.jl:52-67; Sigma0.R = 0) and
// ode_determine_initdt (when adaptive and no dt is given).
template <class Prob, int Q, bool ADAPTIVE>
__device__ __forceinline__ void pn_init_entry(const PnParams& P, double* init_mu, double* init_dt) {
    constexpr int DP = Prob::d, d = PN_SECOND ? DP / 2 : DP, D = d * (Q + 1), NPA = Prob::n_p > 0 ? Prob::n_p : 1;
    for (long long traj = blockIdx.x * (long long)blockDim.x + threadIdx.x; traj < P.n_traj;
         traj += (long long)gridDim.x * blockDim.x) {
        double u0[DP], mu[D], pr[NPA];
#pragma unroll
        for (int i = 0; i < NPA; i++) pr[i] = 0.0;
#pragma unroll
        for (int i = 0; i < DP; i++) u0[i] = P.u0[traj * DP + i];
#pragma unroll
        for (int i = 0; i < Prob::n_p; i++) pr[i] = P.p[traj * Prob::n_p + i];
#if PN_SECOND
        { /* derivatives of the first-order form (du, u); the state takes the u-part (autodiffinit.jl:34-37) */
            double mu2[DP * (Q + 1)];
            pn_taylor_init<Prob, Q>(mu2, u0, pr, P.t0);
#pragma unroll
            for (int o = 0; o <= Q; o++)
#pragma unroll
                for (int i = 0; i < d; i++) mu[o * d + i] = mu2[o * DP + d + i];
        }
#else
        pn_taylor_init<Prob, Q>(mu, u0, pr, P.t0);
#endif
#if PN_MASS
        /* conditioning on MM E_o x = d^o/dt^o of u' = f(u), o >= 1 (autodiffinit.jl:33-47): x_o = MM^-1 df_o */
#pragma unroll
        for (int o = 1; o <= Q; o++) {
            double xo[d];
#pragma unroll
            for (int a = 0; a < d; a++) {
                double sm = 0.0;
#pragma unroll
                for (int i = 0; i < d; i++)
                    if (pn_mass_inv(a, i) != 0.0) sm += pn_mass_inv(a, i) * mu[o * d + i];
                xo[a] = sm;
            }
#pragma unroll
            for (int a = 0; a < d; a++) mu[o * d + a] = xo[a];
        }
        /* initial step with a mass matrix: max(smalldt, dtmin) (see oracle initial_dt; ode_determine_initdt's isdae /
           failed-linsolve branches) */
        double dt = P.dt0;
        if (ADAPTIVE && !(P.dt0 > 0.0)) {
            const double epst = fabs(P.t0) > 0.0 ? (nextafter(fabs(P.t0), 1e300) - fabs(P.t0)) : 4.9406564584124654e-324;
            const double dtmin = nextafter(fmax(P.dtmin, epst), 1e300);
            dt = fmax(fmax(dtmin, 1e-6), dtmin);
        }
#else
        const double dt = (ADAPTIVE && !(P.dt0 > 0.0)) ? pn_initial_dt<Prob>(P, u0, pr) : P.dt0;
#endif
#pragma unroll
        for (int c = 0; c < D; c++) init_mu[traj * D + c] = mu[c];
        init_dt[traj] = dt;
    }
}

template <class Prob, int Q, int G, bool ADAPTIVE, bool DYNAMIC, int SAVE>
__device__ __forceinline__ void pn_small_entry(const PnParams& P) {
    extern __shared__ double pn_small_shared[]; /* (PN_THREADS / G) exchange regions of pn_xs_doubles<D>() doubles */
    PnSmall<Prob, Q, G, ADAPTIVE, DYNAMIC, SAVE>::run(P, pn_small_shared);
}

#endif /* PN_SMALL_CUH */
