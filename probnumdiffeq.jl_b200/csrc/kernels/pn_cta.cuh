// pn_cta.cuh -- CTA-per-trajectory ODE-filter kernel for large filter states (D = d(q+1) > 32, e.g. Pleiades
// d = 28, q = 3, D = 112).  One 256-thread CTA owns one trajectory for its whole adaptive time loop; the right
// factor R (D x D) is staged in shared memory (98 KB at D = 112) and never leaves the SM.
//
// predict_cov! follows the reference literally (src/filtering/predict.jl:84-91): the Gram matrix
// M = X'X + s2 Qh (x) I_d of the stack is formed (register-tiled SYRK, upper triangle in packed storage) and
// factorised by a right-looking Cholesky with ONE barrier per column (rows stay unscaled until the end, so a column
// step only reads row j and writes rows > j); only if a pivot is not positive does it fall back to Householder QR
// (`triangularize!`), exactly as the reference does.  The fallback avoids the dense 2D x D stack:
//
// Shared-memory budget: the dense 2D x D stack of predict_cov! (196 KB) does not fit next to the
// update workspace, so the stack is never materialised:
//   1. X = R Ah' in place (row-local, Ah = T (x) I_d),
//   2. Householder QR of X in place  ->  upper-triangular R1,
//   3. the process-noise block sqrt(s2) (L P^-1) (x) I_d is folded into R1 one derivative level at a time
//      (a d x D batch, 25 KB): QR of [R1 ; batch] touches only the pivot row of R1 and the d batch rows per column.
// The result has the same R'R as qr([R Ah' ; sqrt(s2) Qh.R]) of src/filtering/predict.jl:62-94.
// The measurement update uses H = [-J | I | 0...]: R^- H' is zero in rows >= 2d, so K, S and the Joseph update only
// touch the first 2d rows (src/filtering/update.jl:65-139).
//
// Everything is FP64 on the vector pipe; thread j owns column j in the QR updates (conflict-free shared-memory
// access, the Householder vector is a broadcast read).
//
// NVRTC-compatible: no standard headers.
#ifndef PN_CTA_CUH
#define PN_CTA_CUH
#include "pn_params.h"
#include "pn_taylor.cuh"
#include "pn_small.cuh" /* pn_initial_dt */

#ifndef PN_CTA_THREADS
#define PN_CTA_THREADS 256 /* threads per CTA (one trajectory); the host may build with 512 */
#endif
#ifndef PN_CTA_MIN_BLOCKS
#define PN_CTA_MIN_BLOCKS 1 /* resident CTAs per SM the register allocation must allow (mid-sized states, e.g. D = 56, fit several) */
#endif

/* doubles of dynamic shared memory of the CTA kernel (host and device use this one formula):
   d modelled components, n = q + 1, dp = dimension of the right-hand side (2d for second-order problems), hb = number of
   d-column blocks the measurement matrix H touches (2; 3 for second-order problems), mass: M is staged in shared memory */
__host__ __device__ constexpr long long pn_cta_ev(long long x) { return (x + 1) & ~1LL; } /* keep every buffer 16-byte aligned */
/* doubles of the tile-packed upper factor of an M x M matrix (4 x 4 tiles, PN_CTA_TS = 18 doubles apart) */
__host__ __device__ constexpr long long pn_cta_tiles(long long M) { return ((M + 3) / 4) * ((M + 3) / 4 + 1) / 2 * 18; }
__host__ __device__ constexpr long long pn_cta_uni(int d, int n) {
    const long long D = (long long)d * n, tp = pn_cta_tiles(D), bk = pn_cta_ev(d * D) + d * D; /* tile-packed factor | fold batch + gain */
    return pn_cta_ev(tp > bk ? tp : bk);
}
__host__ __device__ constexpr long long pn_cta_sm_doubles(int d, int n, int dp, int hb, int npa, bool mass, int shared_struct_bytes) {
    const long long D = (long long)d * n, NR = (long long)hb * d;
    return pn_cta_ev(D * D) + pn_cta_uni(d, n) + pn_cta_ev(NR * d) /*C2*/ + pn_cta_ev(pn_cta_tiles(d)) /*tile-packed factor of S, W*/ +
           pn_cta_ev((long long)d * d) /*S*/ + pn_cta_ev(hb == 3 ? (long long)dp * dp : (long long)d * d) /*J (second order: of the first-order form)*/ +
           pn_cta_ev(mass ? (long long)d * d : 0) /*M*/ + pn_cta_ev(3 * D) /*mu,mup,tmp*/ + pn_cta_ev(3 * d) /*z,zt,uprev*/ +
           pn_cta_ev(2 * dp) /*fu,x2*/ + pn_cta_ev(npa) + (shared_struct_bytes + 7) / 8;
}

/* number of independent slices the problem struct offers for f / J (tracer.py, jac_parts); 1 if it does not say */
template <class P, class = void>
struct PnJacParts {
    static constexpr int value = 1;
};
template <class P>
struct PnJacParts<P, decltype((void)P::jac_parts)> {
    static constexpr int value = P::jac_parts;
};

struct PnCtaShared {
    double red[32]; /* one slot per warp; sized for 1024 threads so that host and NVRTC builds agree on sizeof */
    double hp[PN_MAX_N], PIv[PN_MAX_N];
    double t, dt, h, tnew, tstop, EEst, qq, q11, lEE, qold, qold_pb2, ediff, ldiff, loglik, gdiff, dtcache;
    double total, total2, quad;
    double hv[2][4]; /* Householder scalars {beta, vk, gam} of the current / next pivot */
    double Qn[PN_MAX_N * PN_MAX_N]; /* (P^-1 L'L P^-1)[c][c']: the 1-d process-noise covariance of this step */
    long long traj, naccept, nreject, nf, iter, tstop_idx, save_pos;
    int retcode, accept, finished, okflag, chol_fail;
    double cinv[4]; /* reciprocal pivots of the current panel (gram_chol_tiles) */
};
#define PN_CTA_TS 18 /* doubles between consecutive 4 x 4 tiles of the tile-packed factor: 144 bytes, so that the eight lanes of a
                        quarter warp reading consecutive tiles with 16-byte loads hit distinct bank groups */

template <class Prob, int Q, bool ADAPTIVE, bool DYNAMIC, int SAVE>
struct PnCta {
    static constexpr int DP = Prob::d;                 /* dimension of the right-hand side */
    static constexpr int d = PN_SECOND ? DP / 2 : DP;  /* modelled components (positions for second-order problems) */
    static constexpr int DU = DP;                      /* length of sol.u: (du, u) for second-order problems */
    static constexpr int E1B = PN_SECOND ? 2 : 1;      /* derivative block the measurement reads (E2 / E1) */
    static constexpr int HB = E1B + 1;                 /* d-column blocks H touches: H = [-J | M | 0..] or [-J_u | -J_du | I | 0..] */
    static constexpr int NR = HB * d;                  /* R^- is upper triangular, so R^- H' vanishes in rows >= NR */
    static constexpr int n = Q + 1;
    static constexpr int D = d * n;
    static constexpr int NPA = Prob::n_p > 0 ? Prob::n_p : 1;
    static constexpr int NT = PN_CTA_THREADS;
    /* threads that share one column in the Householder sweeps (adjacent lanes, power of two) */
    static constexpr int TPC = (NT / D >= 8) ? 8 : (NT / D >= 4) ? 4 : (NT / D >= 2) ? 2 : 1;
    static constexpr int NB = (d + TPC - 1) / TPC; /* batch rows per thread in the fold sweeps */
    /* doubles of dynamic shared memory */
    /* packed upper triangle of the Gram matrix; the region is shared with the fold batch Bt (d x D, QR fallback only)
       and with the gain K (D x d, computed after the factor has been copied into R) */
    static constexpr int UNI = (int)pn_cta_uni(d, n);
    static constexpr int JS = PN_SECOND ? DP * DP : d * d; /* Jacobian buffer (second order: of the first-order form) */
    static constexpr int SM_DOUBLES = (int)pn_cta_sm_doubles(d, n, DP, HB, NPA, PN_MASS != 0, (int)sizeof(PnCtaShared));

    static __device__ __forceinline__ double block_sum(double x, PnCtaShared* sh) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) sh->red[threadIdx.x >> 5] = x;
        __syncthreads();
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NT / 32; w++) s += sh->red[w];
        return s;
    }

    // beta, v_k = alpha - beta, gam = 2 / v'v of the reflector that maps (alpha, x) with |(alpha,x)|^2 = total onto beta e_1
    static __device__ __forceinline__ void house_scalars(double* hv, double total, double alpha, bool nz) {
        const double nrm = nz ? sqrt(total) : 0.0;
        const double beta = nz ? -copysign(nrm, alpha) : alpha;
        hv[0] = beta;
        hv[1] = alpha - beta;
        hv[2] = nz ? 1.0 / (total + fabs(alpha) * nrm) : 0.0;
    }

    // ----------------------------------------------------------------------------------------------------------------
    // Gram matrix + Cholesky on 4 x 4 register tiles (predict.jl:84-85 for the D x D factor; the d x d matrices S and W use
    // the same routine).  Tile (I, J), I <= J, of the symmetric M x M matrix belongs to thread e % NT (e = row-major index over
    // the upper triangle of tiles) for the whole routine and lives in its REGISTERS: the Gram phase accumulates it from the
    // rows of A (two 32-byte operands per 16 multiply-adds), and the right-looking factorisation updates it in place; a tile
    // is written to shared memory once, when its tile row becomes the panel.  Per panel: the owner of the diagonal tile
    // factorises it (4 pivots) and publishes it with the reciprocal pivots; barrier; the owners of the panel row solve their
    // own tiles against it and publish them; barrier; every owner below subtracts U_PI' U_PJ (two tiles read, 64 FMAs).  Two
    // barriers per 4 columns, no read-modify-write traffic on shared memory.
    //   A      rows x M row-major (leading dimension M) or nullptr (no Gram phase: tiles start from init alone)
    //   init   (gi, gj) -> value added to entry (gi, gj), gi <= gj
    //   Tp     tile-packed result: tile (I, J) at Tp + tix(I, J) * PN_CTA_TS, row-major 4 x 4 (upper triangle of U)
    //   invd   optional: 1 / U[j][j]
    // Returns false (CTA-uniform) as soon as a pivot is not positive (the reference's PosDefException, predict.jl:85-91);
    // Tp is then incomplete.
    // ----------------------------------------------------------------------------------------------------------------
    template <int M>
    static __device__ __forceinline__ int tix(int I, int J) {
        constexpr int TI = (M + 3) / 4;
        return I * TI - (I * (I - 1)) / 2 + (J - I);
    }
    template <int M, class Init>
    static __device__ bool gram_chol_tiles(const double* A, int rows, double* Tp, double* invd, PnCtaShared* sh, Init init) {
        constexpr int TI = (M + 3) / 4, NTILES = TI * (TI + 1) / 2, TPT = (NTILES + NT - 1) / NT, TS = PN_CTA_TS;
        constexpr bool ALIGNED = (M % 4 == 0);
        const int tid = threadIdx.x;
        int tI[TPT], tJ[TPT];
        double acc[TPT][4][4];
#pragma unroll
        for (int k = 0; k < TPT; k++) {
            int e = tid + k * NT, I = 0;
            if (e >= NTILES) e = -1;
            int rem = e;
            while (e >= 0 && rem >= TI - I) {
                rem -= TI - I;
                I++;
            }
            tI[k] = (e >= 0) ? I : -1;
            tJ[k] = (e >= 0) ? I + rem : -1;
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) acc[k][a][b] = 0.0;
        }
        if (A != nullptr) {
#pragma unroll 2
            for (int r = 0; r < rows; r++) {
                const double* row = A + r * M;
#pragma unroll
                for (int k = 0; k < TPT; k++) {
                    if (tI[k] < 0) continue;
                    double av[4], bv[4];
                    if constexpr (ALIGNED) {
                        const double2 a01 = *reinterpret_cast<const double2*>(row + 4 * tI[k]);
                        const double2 a23 = *reinterpret_cast<const double2*>(row + 4 * tI[k] + 2);
                        const double2 b01 = *reinterpret_cast<const double2*>(row + 4 * tJ[k]);
                        const double2 b23 = *reinterpret_cast<const double2*>(row + 4 * tJ[k] + 2);
                        av[0] = a01.x, av[1] = a01.y, av[2] = a23.x, av[3] = a23.y;
                        bv[0] = b01.x, bv[1] = b01.y, bv[2] = b23.x, bv[3] = b23.y;
                    } else {
#pragma unroll
                        for (int a = 0; a < 4; a++) {
                            av[a] = (4 * tI[k] + a < M) ? row[4 * tI[k] + a] : 0.0;
                            bv[a] = (4 * tJ[k] + a < M) ? row[4 * tJ[k] + a] : 0.0;
                        }
                    }
#pragma unroll
                    for (int a = 0; a < 4; a++)
#pragma unroll
                        for (int b = 0; b < 4; b++) acc[k][a][b] = fma(av[a], bv[b], acc[k][a][b]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < TPT; k++) {
            if (tI[k] < 0) continue;
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const int gi = 4 * tI[k] + a, gj = 4 * tJ[k] + b;
                    if (gi < M && gj < M) {
                        if (gi <= gj) acc[k][a][b] += init(gi, gj);
                    } else {
                        acc[k][a][b] = (gi == gj) ? 1.0 : 0.0; /* padding of the last tile row / column: identity */
                    }
                }
        }
        /* diagonal tile of a panel: factor in registers, publish U_PP (upper, zeros below) and the reciprocal pivots */
        auto factor_diag = [&](int k, int P) {
            double inv[4];
            const bool ok = pn_chol<4>(acc[k], inv);
            double* dst = Tp + tix<M>(P, P) * TS;
#pragma unroll
            for (int a = 0; a < 4; a++) {
                *reinterpret_cast<double2*>(dst + 4 * a) = make_double2(a <= 0 ? acc[k][a][0] : 0.0, a <= 1 ? acc[k][a][1] : 0.0);
                *reinterpret_cast<double2*>(dst + 4 * a + 2) = make_double2(a <= 2 ? acc[k][a][2] : 0.0, acc[k][a][3]);
                sh->cinv[a] = inv[a];
                if (invd != nullptr && 4 * P + a < M) invd[4 * P + a] = inv[a];
            }
            if (!ok) sh->chol_fail = 1;
        };
        if (tid == 0) sh->chol_fail = 0;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TPT; k++)
            if (tI[k] == 0 && tJ[k] == 0) factor_diag(k, 0);
        bool ok = true;
        for (int P = 0; P < TI; P++) {
            __syncthreads(); /* U_PP, cinv, chol_fail of panel P are visible */
            if (sh->chol_fail) {
                ok = false;
                break;
            }
            /* panel row: U_PJ = U_PP^-T M_PJ, solved in the owner's registers */
#pragma unroll
            for (int k = 0; k < TPT; k++) {
                if (tI[k] != P || tJ[k] == P) continue;
                const double* dg = Tp + tix<M>(P, P) * TS;
                const double u01 = dg[1], u02 = dg[2], u03 = dg[3], u12 = dg[6], u13 = dg[7], u23 = dg[11];
                const double i0 = sh->cinv[0], i1 = sh->cinv[1], i2 = sh->cinv[2], i3 = sh->cinv[3];
                double* dst = Tp + tix<M>(P, tJ[k]) * TS;
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const double x0 = acc[k][0][b] * i0;
                    const double x1 = (acc[k][1][b] - u01 * x0) * i1;
                    const double x2 = (acc[k][2][b] - u02 * x0 - u12 * x1) * i2;
                    const double x3 = (acc[k][3][b] - u03 * x0 - u13 * x1 - u23 * x2) * i3;
                    acc[k][0][b] = x0, acc[k][1][b] = x1, acc[k][2][b] = x2, acc[k][3][b] = x3;
                }
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    *reinterpret_cast<double2*>(dst + 4 * a) = make_double2(acc[k][a][0], acc[k][a][1]);
                    *reinterpret_cast<double2*>(dst + 4 * a + 2) = make_double2(acc[k][a][2], acc[k][a][3]);
                }
            }
            if (P + 1 == TI) break;
            __syncthreads(); /* the panel row is visible */
#pragma unroll
            for (int k = 0; k < TPT; k++) {
                if (tI[k] <= P) continue; /* also skips unused slots (-1) */
                const double* pi = Tp + tix<M>(P, tI[k]) * TS;
                const double* pj = Tp + tix<M>(P, tJ[k]) * TS;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const double2 a01 = *reinterpret_cast<const double2*>(pi + 4 * q), a23 = *reinterpret_cast<const double2*>(pi + 4 * q + 2);
                    const double2 b01 = *reinterpret_cast<const double2*>(pj + 4 * q), b23 = *reinterpret_cast<const double2*>(pj + 4 * q + 2);
                    const double av[4] = {a01.x, a01.y, a23.x, a23.y};
                    const double bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
                    for (int a = 0; a < 4; a++)
#pragma unroll
                        for (int b = 0; b < 4; b++) acc[k][a][b] = fma(-av[a], bv[b], acc[k][a][b]);
                }
                if (tI[k] == P + 1 && tJ[k] == P + 1) factor_diag(k, P + 1);
            }
        }
        __syncthreads();
        return ok;
    }
    /* tile-packed upper factor -> dense row-major M x M (zeros below the diagonal) */
    template <int M>
    static __device__ void expand_tiles(const double* Tp, double* out) {
        for (int e = threadIdx.x; e < M * M; e += NT) {
            const int i = e / M, k = e - i * M;
            out[e] = (k >= i) ? Tp[tix<M>(i >> 2, k >> 2) * PN_CTA_TS + (i & 3) * 4 + (k & 3)] : 0.0;
        }
    }

    // x <- (U'U)^-1 x for m <= 32 by one warp (call with all 32 lanes of the warp): lane l holds x[l]; each solved
    // entry is broadcast and eliminated from the remaining lanes (column-oriented substitution).
    static __device__ void chol_solve_warp(const double* U, const double* invd, double* x, int m) {
        const int lane = threadIdx.x & 31;
        double xv = (lane < m) ? x[lane] : 0.0;
        for (int a = 0; a < m; a++) { /* U' y = x */
            const double ya = __shfl_sync(0xffffffffu, xv, a) * invd[a];
            if (lane == a) xv = ya;
            else if (lane > a && lane < m) xv -= U[a * m + lane] * ya;
        }
        for (int a = m - 1; a >= 0; a--) { /* U x = y */
            const double xa = __shfl_sync(0xffffffffu, xv, a) * invd[a];
            if (lane == a) xv = xa;
            else if (lane < a) xv -= U[lane * m + a] * xa;
        }
        if (lane < m) x[lane] = xv;
    }
    // serial variant (one thread) for m > 32
    static __device__ void chol_solve_serial(const double* U, const double* invd, double* x, int m) {
        for (int a = 0; a < m; a++) {
            double s = x[a];
            for (int k = 0; k < a; k++) s -= U[k * m + a] * x[k];
            x[a] = s * invd[a];
        }
        for (int a = m - 1; a >= 0; a--) {
            double s = x[a];
            for (int k = a + 1; k < m; k++) s -= U[a * m + k] * x[k];
            x[a] = s * invd[a];
        }
    }

    // K' <- S^-1 K' column by column (one thread per column of the d x D matrix at K), mu = mu^- - K z.  S holds the upper
    // Cholesky factor, invd its reciprocal diagonal.  Column-oriented substitution in registers: per pivot one dependent
    // multiply and one level of independent FMAs; the factor is read as shared-memory broadcasts.  Not inlined: the routine
    // wants d + d^2/2-ish registers of its own, which the surrounding kernel cannot spare.
    static __device__ __forceinline__ void gain_columns(double* K, const volatile double* S, const double* invd, const double* mup, double* mu,
                                                     const double* z) {
        for (int c = threadIdx.x; c < D; c += NT) {
            double x[d];
#pragma unroll
            for (int a = 0; a < d; a++) x[a] = K[a * D + c];
#pragma unroll
            for (int a = 0; a < d; a++) { /* U' y = b */
                x[a] *= invd[a];
#pragma unroll
                for (int j = 0; j < d; j++)
                    if (j > a) x[j] = fma(-S[a * d + j], x[a], x[j]);
            }
#pragma unroll
            for (int a = d - 1; a >= 0; a--) { /* U k = y */
                double acc = x[a];
#pragma unroll
                for (int j = 0; j < d; j++)
                    if (j > a) acc = fma(-S[a * d + j], x[j], acc);
                x[a] = acc * invd[a];
            }
            double m = mup[c];
#pragma unroll
            for (int a = 0; a < d; a++) {
                K[a * D + c] = x[a];
                m = fma(-x[a], z[a], m);
            }
            mu[c] = m; /* mu = mu^- - K z (update.jl:115) */
        }
    }

    static __device__ void run(const PnParams& P) {
        extern __shared__ __align__(16) double pn_cta_sm[];
        const int tid = threadIdx.x;
        double* R = pn_cta_sm;
        double* Mp = R + pn_cta_ev(D * D); /* tile-packed factor of the Gram matrix | Bt, K */
        double* Bt = Mp;
        double* K = Mp + pn_cta_ev(d * D);
        double* C2 = Mp + UNI;
        double* Ts = C2 + pn_cta_ev(NR * d); /* tile-packed factors of the d x d matrices W and S */
        double* S = Ts + pn_cta_ev(pn_cta_tiles(d));
        double* J = S + pn_cta_ev(d * d);
        double* Mm = J + pn_cta_ev(JS); /* mass matrix (PN_MASS) */
        double* mu = Mm + pn_cta_ev(PN_MASS ? d * d : 0);
        double* mup = mu + D;
        double* sinv = mup + D; /* D doubles of scratch: first d used as inverse Cholesky diagonal */
        double* z = sinv + D + (pn_cta_ev(3 * D) - 3 * D);
        double* zt = z + d;
        double* uprev = zt + d;
        double* fu = uprev + d + (pn_cta_ev(3 * d) - 3 * d);
        double* x2 = fu + DP;
        double* pr = x2 + DP;
        PnCtaShared* sh = (PnCtaShared*)(pr + pn_cta_ev(NPA));
        /* entries of H's column blocks (as used in W = H Qh H' and C2 = R^- H'):
           first order  H = [-J | M | 0 ...]   (M = I without a mass matrix; measurement_models.jl:11-20, derivative_utils.jl:1-33)
           second order H = [-J_u | -J_du | I | 0 ...]   (measurement_models.jl:29-49, derivative_utils.jl:35-53) */
        auto Ju = [&](int a, int i) -> double { return PN_SECOND ? J[a + DP * (d + i)] : J[a + d * i]; }; /* d f / d u */
        auto Jd = [&](int a, int i) -> double { return PN_SECOND ? J[a + DP * i] : 0.0; };                /* d f1 / d du */
        auto Hb = [&](int b, int a, int i) -> double {
            if (b == 0) return -Ju(a, i);
            if (PN_SECOND) return (b == 1) ? -Jd(a, i) : ((a == i) ? 1.0 : 0.0);
            return PN_MASS ? Mm[a * d + i] : ((a == i) ? 1.0 : 0.0);
        };
#if PN_MASS
        for (int e = tid; e < d * d; e += NT) Mm[e] = pn_mass(e / d, e % d);
#endif

        for (;;) {
            /* ---- fetch ---- */
            __syncthreads();
            if (tid == 0) sh->traj = (long long)atomicAdd(P.work_counter, 1ULL) * P.traj_stride;
            __syncthreads();
            const long long traj = sh->traj;
            if (traj >= P.n_traj) break;
            /* initial state and first step: computed for the whole ensemble by pn_init_kernel (TaylorModeInit incl. the
               second-order / mass-matrix conditioning, ode_determine_initdt) before this kernel started */
            for (int c = tid; c < D; c += NT) mu[c] = P.init_mu[traj * D + c];
            for (int i = tid; i < Prob::n_p; i += NT) pr[i] = P.p[traj * Prob::n_p + i];
            __syncthreads();
            for (int i = tid; i < d; i += NT) uprev[i] = mu[pn_solcol<d>(i)];
            if (tid == 0) {
                sh->t = P.t0;
                sh->dt = P.init_dt[traj];
                sh->nf = Q + ((ADAPTIVE && !(P.dt0 > 0.0) && !PN_MASS) ? 2 : 0);
                sh->dtcache = sh->dt;
                sh->naccept = sh->nreject = sh->iter = sh->tstop_idx = 0;
                sh->loglik = sh->gdiff = sh->ldiff = 0.0;
                sh->qold = P.qoldinit;
                sh->qold_pb2 = P.qoldinit_pb2;
                sh->retcode = PN_RET_SUCCESS;
                sh->finished = 0;
                sh->save_pos = 0;
                if (SAVE >= 1) {
                    long long sp = P.step_offsets ? P.step_offsets[traj] : 0;
                    if (P.step_offsets) {
                        P.s_t[sp] = P.t0;
                        P.s_diffusion[sp] = 0.0;
                        for (int i = 0; i < DU; i++) P.s_u_mean[sp * DU + i] = mu[pn_solcol<d>(i)];
                        for (int i = 0; i < DU * DU; i++) P.s_u_cov[sp * DU * DU + i] = 0.0;
                        if (SAVE >= 2)
                            for (int c = 0; c < D; c++) P.s_x_mean[sp * D + c] = mu[c];
                    }
                    sh->save_pos = sp + 1;
                }
            }
            for (int e = tid; e < D * D; e += NT) R[e] = 0.0;
            if (SAVE >= 2 && P.step_offsets) { /* Sigma_0.R = 0 */
                double* dst = P.s_x_covfac + P.step_offsets[traj] * (long long)D * D;
                for (int e = tid; e < D * D; e += NT) dst[e] = 0.0;
            }
            __syncthreads();

            /* ---- adaptive time loop ---- */
            while (!sh->finished) {
                /* phase 1a: scalars (thread 0) */
                if (tid == 0) {
                    double t = sh->t, dt = sh->dt, tstop = P.t1;
                    long long ti = sh->tstop_idx;
                    while (ti < P.n_tstops && P.tstops[ti] <= t) ti++;
                    if (ti < P.n_tstops && P.tstops[ti] < P.t1) tstop = P.tstops[ti];
                    sh->tstop_idx = ti;
                    sh->iter++;
                    if (!(t < P.t1)) {
                        sh->finished = 1;
                    } else if (sh->iter > P.maxiters) {
                        sh->retcode = PN_RET_MAXITERS;
                        sh->finished = 1;
                    } else {
                        if (ADAPTIVE) {
                            dt = fmin(dt, P.dtmax);
                            dt = fmax(dt, P.dtmin);
                            dt = fmin(dt, tstop - t);
                        } else {
                            dt = fmin(sh->dtcache, tstop - t);
                        }
                        if (dt != dt || !(t + dt > t) || (ADAPTIVE && dt <= P.dtmin)) {
                            sh->retcode = PN_RET_DTLESSTHANMIN;
                            sh->finished = 1;
                        }
                        sh->dt = dt;
                        sh->h = dt;
                        sh->tnew = t + dt;
                        sh->tstop = tstop;
                        sh->hp[0] = 1.0;
                        for (int k = 1; k < n; k++) sh->hp[k] = sh->hp[k - 1] * dt / (double)k;
                        const double sqh = sqrt(dt);
                        for (int c = 0; c < n; c++) sh->PIv[c] = sh->hp[Q - c] * sqh;
                    }
                }
                __syncthreads();
                if (sh->finished) break;
                const double h = sh->h;
                /* predict_mean! */
                for (int c = tid; c < D; c += NT) {
                    const int j = c / d, i = c - j * d;
                    double s = mu[c];
                    for (int k = j + 1; k < n; k++) s += sh->hp[k - j] * mu[k * d + i];
                    mup[c] = s;
                }
                __syncthreads();
                /* measurement: f on one thread, J on a thread of another warp (or in Prob::jac_parts independent slices,
                   slice k on the first lane of warp k) */
                const double* xin = mup;
#if PN_SECOND
                for (int i = tid; i < d; i += NT) { /* (du, u) = (E1 mu^-, E0 mu^-) */
                    x2[i] = mup[d + i];
                    x2[d + i] = mup[i];
                }
                __syncthreads();
                xin = x2;
#endif
                {
                    constexpr int JP = PnJacParts<Prob>::value;
                    static_assert(JP >= 1 && JP <= NT / 32, "jac_parts must not exceed the number of warps of the CTA");
                    if (JP == 1) {
                        if (tid == 0) Prob::template f<double>(fu, xin, pr, sh->tnew);
                        if (tid == 32) Prob::jac(J, xin, pr, sh->tnew);
                    }
#define PN_CTA_RHS_SLICE(K)                                                              \
    if (JP > 1 && JP > K && tid == 32 * K) {                                                  \
        Prob::template f_part<(JP > K ? K : 0)>(fu, xin, pr, sh->tnew);                  \
        Prob::template jac_part<(JP > K ? K : 0)>(J, xin, pr, sh->tnew);                 \
    }
                    if constexpr (JP > 1) {
                    PN_CTA_RHS_SLICE(0)
                    PN_CTA_RHS_SLICE(1)
                    PN_CTA_RHS_SLICE(2)
                    PN_CTA_RHS_SLICE(3)
                    PN_CTA_RHS_SLICE(4)
                    PN_CTA_RHS_SLICE(5)
                    PN_CTA_RHS_SLICE(6)
                    PN_CTA_RHS_SLICE(7)
                    }
#undef PN_CTA_RHS_SLICE
                }
                __syncthreads();
                for (int a = tid; a < d; a += NT) {
#if PN_MASS
                    double sm = 0.0; /* z = M E1 mu^- - f (measurement_models.jl:18) */
                    for (int i = 0; i < d; i++) sm += Mm[a * d + i] * mup[d + i];
                    z[a] = sm - fu[a];
#else
                    z[a] = mup[E1B * d + a] - fu[a];
#endif
                }
                if (tid == 0) sh->nf += 1;
                if (ADAPTIVE || DYNAMIC) {
                    /* W = C'C, C[(m,i),a] = -LQ[m][0] J[a][i] + (i==a) LQ[m][1]   (calibration.jl:123-133); summed over m in
                       closed form: W = alpha J J' - beta (J + J') + gamma I with alpha = sum LQ[m][0]^2, beta = sum LQ[m][0] LQ[m][1],
                       gamma = sum LQ[m][1]^2 */
                    if (!PN_SECOND && !PN_MASS) {
                        double al = 0.0, be = 0.0, ga = 0.0;
                        for (int m = 0; m < n; m++) {
                            const double l0 = P.L[m * n + 0] * sh->PIv[0];
                            const double l1 = (m >= 1) ? P.L[m * n + 1] * sh->PIv[1] : 0.0;
                            al += l0 * l0;
                            be += l0 * l1;
                            ga += l1 * l1;
                        }
                        for (int e = tid; e < d * d; e += NT) {
                            const int a = e / d, b = e - a * d;
                            if (b < a) continue;
                            double jj = 0.0;
                            for (int i = 0; i < d; i++) jj += J[a + d * i] * J[b + d * i];
                            S[a * d + b] = al * jj - be * (J[a + d * b] + J[b + d * a]) + ((a == b) ? ga : 0.0);
                        }
                    } else {
                        /* general form: W = sum_{b1,b2} qh[b1][b2] H_b1 H_b2', qh = P^-1 (L'L) P^-1 restricted to the HB blocks of H */
                        double qh[HB][HB];
                        for (int b1 = 0; b1 < HB; b1++)
                            for (int b2 = 0; b2 < HB; b2++) {
                                double acc = 0.0;
                                for (int m = 0; m < n; m++) acc += P.L[m * n + b1] * P.L[m * n + b2];
                                qh[b1][b2] = acc * sh->PIv[b1] * sh->PIv[b2];
                            }
                        for (int e = tid; e < d * d; e += NT) {
                            const int a = e / d, b = e - a * d;
                            if (b < a) continue;
                            double w = 0.0;
                            for (int b1 = 0; b1 < HB; b1++)
                                for (int b2 = 0; b2 < HB; b2++) {
                                    double acc = 0.0;
                                    for (int i = 0; i < d; i++) acc += Hb(b1, a, i) * Hb(b2, b, i);
                                    w += qh[b1][b2] * acc;
                                }
                            S[a * d + b] = w;
                        }
                    }
                    __syncthreads();
                    for (int i = tid; i < d; i += NT) {
                        sinv[d + i] = S[i * d + i]; /* W_ii for the error estimate */
                        zt[i] = z[i];
                    }
                    __syncthreads();
                    const bool okW = gram_chol_tiles<d>(nullptr, 0, Ts, sinv, sh, [&](int gi, int gj) -> double { return S[gi * d + gj]; });
                    expand_tiles<d>(Ts, S);
                    __syncthreads();
                    if (d <= 32) {
                        if (tid < 32) chol_solve_warp(S, sinv, zt, d);
                    } else if (tid == 0) {
                        chol_solve_serial(S, sinv, zt, d);
                    }
                    __syncthreads();
                    if (tid == 0) {
                        sh->okflag = okW ? 1 : 0;
                        double qd = 0.0;
                        for (int a = 0; a < d; a++) qd += z[a] * zt[a];
                        double ldiff = qd * (1.0 / (double)d);
                        if (P.forced_diffusion && sh->naccept < P.n_forced) ldiff = P.forced_diffusion[sh->naccept];
                        sh->ldiff = ldiff;
                        if (!sh->okflag) {
                            sh->retcode = PN_RET_UNSTABLE;
                            sh->finished = 1;
                        }
                    }
                    __syncthreads();
                }
                /* error estimate + controller (thread 0) */
                if (tid == 0 && !sh->finished) {
                    double EEst = 0.0, qq = 1.0, q11 = 0.0, lEE = 0.0;
                    int accept = 1;
                    if (ADAPTIVE) {
                        double e2 = 0.0;
                        for (int i = 0; i < d; i++) {
                            const double isc = 1.0 / (P.abstol + fmax(fabs(mup[pn_solcol<d>(i)]), fabs(uprev[i])) * P.reltol);
                            e2 += (sh->ldiff * sinv[d + i]) * (isc * isc);
                        }
                        e2 = e2 * (h * h) * (1.0 / (double)d);
                        EEst = (e2 > 0.0) ? sqrt(e2) : e2;
                        if (EEst != EEst) {
                            sh->retcode = PN_RET_UNSTABLE;
                            sh->finished = 1;
                        }
                        if (EEst == 0.0) {
                            qq = 1.0 / P.qmax;
                        } else {
                            lEE = log(EEst);
                            q11 = exp(P.beta1 * lEE);
                            qq = q11 / (sh->qold_pb2 * P.gamma);
                            qq = fmax(1.0 / P.qmax, fmin(1.0 / P.qmin, qq));
                        }
                        accept = (EEst < 1.0) && !sh->finished;
                        if (!accept && !sh->finished) {
                            sh->nreject++;
                            sh->dt = sh->dt / fmin(1.0 / P.qmin, q11 / P.gamma);
                        }
                    }
                    sh->EEst = EEst;
                    sh->qq = qq;
                    sh->q11 = q11;
                    sh->lEE = lEE;
                    sh->accept = accept;
                    double ediff = DYNAMIC ? sh->ldiff : P.initial_diffusion;
                    if (P.forced_diffusion && sh->naccept < P.n_forced) ediff = P.forced_diffusion[sh->naccept];
                    sh->ediff = ediff;
                }
                __syncthreads();
                if (sh->finished) break;
                if (!sh->accept) continue;

                /* ================= phase 2 ================= */
                /* 1. X = R Ah' in place (ascending j is safe) */
                for (int e = tid; e < D * d; e += NT) { /* one (row, component) pair per work item: conflict-free */
                    const int r = e / d, i = e - r * d;
                    double* row = R + r * D + i;
                    double v[n];
#pragma unroll
                    for (int j = 0; j < n; j++) v[j] = row[j * d];
#pragma unroll
                    for (int j = 0; j < n; j++) {
                        double s = v[j];
#pragma unroll
                        for (int k = 0; k < n; k++)
                            if (k > j) s += sh->hp[k > j ? k - j : 0] * v[k];
                        row[j * d] = s;
                    }
                }
                __syncthreads();
                /* 2. Gram matrix + Cholesky (predict.jl:84-85) */
                bool need_qr = true;
#ifndef PN_CTA_QR_ONLY
                {
                    if (tid < n * n) { /* Qn = P^-1 (L'L) P^-1 */
                        const int c = tid / n, c2 = tid - c * n;
                        double acc = 0.0;
                        for (int m = 0; m < n; m++) acc += P.L[m * n + c] * P.L[m * n + c2];
                        sh->Qn[tid] = acc * sh->PIv[c] * sh->PIv[c2];
                    }
                    __syncthreads();
                    const double ediff = sh->ediff;
                    /* M = X'X + s2 Qh (x) I_d (predict.jl:79-85), factorised tile by tile in registers */
                    const bool ok = gram_chol_tiles<D>(R, D, Mp, nullptr, sh, [&](int gi, int gj) -> double {
                        const int ji = gi / d, jj = gj / d;
                        return (gi - ji * d == gj - jj * d) ? ediff * sh->Qn[ji * n + jj] : 0.0;
                    });
                    if (ok) { /* R^- = U: expand the tile-packed factor into the dense one */
                        expand_tiles<D>(Mp, R);
                        need_qr = false;
                    }
                    __syncthreads();
                }
#endif
                /* 2'.+3'. fallback (predict.jl:90, triangularize!): Householder sweeps with ONE barrier per pivot.  TPC
                   adjacent lanes share a column (rows interleaved); the owners of column k+1 compute its norm right after
                   updating it, and the owners of column k-1 (idle by then) write its diagonal and clear its sub-diagonal. */
                if (need_qr) {
                    const int col = tid / TPC, sub = tid % TPC;
                    /* Householder scalars of pivot k live in sh->hv[k & 1] = {beta, vk, gam}: they are computed once, by
                       the owner of the pivot column right after it has updated that column (double-buffered, so a
                       pivot needs exactly one barrier). */
                    /* ---- sweep A: QR of X in place ---- */
                    {
                        double part = 0.0;
                        if (col == 0)
                            for (int r = sub; r < D; r += TPC) part += R[r * D] * R[r * D];
#pragma unroll
                        for (int o = TPC >> 1; o >= 1; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                        if (col == 0 && sub == 0) house_scalars(sh->hv[0], part, R[0], part > 0.0);
                    }
                    double beta_prev = 0.0;
                    for (int k = 0; k < D; k++) {
                        __syncthreads();
                        const double beta = sh->hv[k & 1][0], vk = sh->hv[k & 1][1], gam = sh->hv[k & 1][2];
                        double s = 0.0;
                        const bool mine = (col > k && col < D);
                        if (mine) {
                            /* loads are batched four rows at a time: the compiler cannot prove that column k and
                               column col do not alias, so a plain loop serialises on the shared-memory latency */
                            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                            int r = k + 1 + sub;
                            for (; r + 3 * TPC < D; r += 4 * TPC) {
                                const double v0 = R[r * D + k], v1 = R[(r + TPC) * D + k], v2 = R[(r + 2 * TPC) * D + k],
                                             v3 = R[(r + 3 * TPC) * D + k];
                                const double a0 = R[r * D + col], a1 = R[(r + TPC) * D + col], a2 = R[(r + 2 * TPC) * D + col],
                                             a3 = R[(r + 3 * TPC) * D + col];
                                s0 += v0 * a0;
                                s1 += v1 * a1;
                                s2 += v2 * a2;
                                s3 += v3 * a3;
                            }
                            for (; r < D; r += TPC) s0 += R[r * D + k] * R[r * D + col];
                            s = (s0 + s1) + (s2 + s3);
                            if (sub == 0) s += vk * R[k * D + col];
                        }
#pragma unroll
                        for (int o = TPC >> 1; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                        s *= gam;
                        double np = 0.0, anext = 0.0;
                        if (mine) {
                            const bool next = (col == k + 1);
                            int r = k + 1 + sub;
                            for (; r + 3 * TPC < D; r += 4 * TPC) {
                                const double v0 = R[r * D + k], v1 = R[(r + TPC) * D + k], v2 = R[(r + 2 * TPC) * D + k],
                                             v3 = R[(r + 3 * TPC) * D + k];
                                double a0 = R[r * D + col], a1 = R[(r + TPC) * D + col], a2 = R[(r + 2 * TPC) * D + col],
                                       a3 = R[(r + 3 * TPC) * D + col];
                                a0 -= s * v0;
                                a1 -= s * v1;
                                a2 -= s * v2;
                                a3 -= s * v3;
                                R[r * D + col] = a0;
                                R[(r + TPC) * D + col] = a1;
                                R[(r + 2 * TPC) * D + col] = a2;
                                R[(r + 3 * TPC) * D + col] = a3;
                                if (next) {
                                    np += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
                                    if (r == k + 1) anext = a0;
                                }
                            }
                            for (; r < D; r += TPC) {
                                const double v = R[r * D + col] - s * R[r * D + k];
                                R[r * D + col] = v;
                                if (next) {
                                    np += v * v;
                                    if (r == k + 1) anext = v;
                                }
                            }
                            if (sub == 0) R[k * D + col] -= s * vk;
                        }
#pragma unroll
                        for (int o = TPC >> 1; o >= 1; o >>= 1) np += __shfl_xor_sync(0xffffffffu, np, o);
                        if (col == k + 1 && sub == 0 && k + 1 < D) house_scalars(sh->hv[(k + 1) & 1], np, anext, np > 0.0);
                        /* column k-1 is dead for everyone now: finish it */
                        if (col == k - 1 && k >= 1) {
                            for (int r = k + sub; r < D; r += TPC) R[r * D + col] = 0.0;
                            if (sub == 0) R[(k - 1) * D + col] = beta_prev;
                        }
                        beta_prev = beta;
                    }
                    __syncthreads();
                    if (tid == 0) R[(D - 1) * D + (D - 1)] = beta_prev;
                    __syncthreads();

                    /* ---- sweeps B: fold sqrt(s2) (U P^-1) (x) I_d into R, one derivative level (d rows) at a time.
                       U is the *upper* triangular factor with U'U = L'L (host, pnode_api.cu): level m is zero in the
                       columns of derivatives < m, so its sweep starts at pivot m*d. ---- */
                    const double sd = (sh->ediff == 1.0) ? 1.0 : sqrt(sh->ediff);
                    for (int m = 0; m < n; m++) {
                        if (sh->ediff == 0.0) break; /* predict.jl:71-74: no process noise */
                        const int k0 = m * d;
                        for (int e = tid; e < d * D; e += NT) {
                            const int i2 = e / D, c = e - i2 * D, j = c / d, ii = c - j * d;
                            Bt[e] = (ii == i2 && j >= m) ? sd * P.U[m * n + j] * sh->PIv[j] : 0.0;
                        }
                        __syncthreads();
                        {
                            double part = 0.0;
                            if (col == k0)
                                for (int i = sub; i < d; i += TPC) part += Bt[i * D + k0] * Bt[i * D + k0];
#pragma unroll
                            for (int o = TPC >> 1; o >= 1; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                            if (col == k0 && sub == 0) {
                                const double a0 = R[k0 * D + k0];
                                house_scalars(sh->hv[k0 & 1], part + a0 * a0, a0, part > 0.0);
                            }
                        }
                        double bprev = 0.0;
                        bool have_prev = false;
                        for (int k = k0; k < D; k++) {
                            __syncthreads();
                            const double beta = sh->hv[k & 1][0], vk = sh->hv[k & 1][1], gam = sh->hv[k & 1][2];
                            double s = 0.0;
                            const bool mine = (col > k && col < D);
                            double bv[NB], ba[NB]; /* this thread's rows of batch columns k and col, kept in registers */
                            if (mine) {
#pragma unroll
                                for (int ib = 0; ib < NB; ib++) {
                                    const int i = sub + ib * TPC;
                                    bv[ib] = (i < d) ? Bt[i * D + k] : 0.0;
                                    ba[ib] = (i < d) ? Bt[i * D + col] : 0.0;
                                }
                                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                                for (int ib = 0; ib < NB; ib++) {
                                    if (ib & 1) s1 += bv[ib] * ba[ib];
                                    else s0 += bv[ib] * ba[ib];
                                }
                                s = s0 + s1;
                                if (sub == 0) s += vk * R[k * D + col];
                            }
#pragma unroll
                            for (int o = TPC >> 1; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                            s *= gam;
                            double np = 0.0;
                            if (mine) {
                                const bool next = (col == k + 1);
#pragma unroll
                                for (int ib = 0; ib < NB; ib++) {
                                    const int i = sub + ib * TPC;
                                    const double v = ba[ib] - s * bv[ib];
                                    if (i < d) Bt[i * D + col] = v;
                                    if (next) np += v * v;
                                }
                                if (sub == 0) R[k * D + col] -= s * vk;
                            }
#pragma unroll
                            for (int o = TPC >> 1; o >= 1; o >>= 1) np += __shfl_xor_sync(0xffffffffu, np, o);
                            if (col == k + 1 && sub == 0 && k + 1 < D) {
                                const double a1 = R[(k + 1) * D + (k + 1)];
                                house_scalars(sh->hv[(k + 1) & 1], np + a1 * a1, a1, np > 0.0);
                            }
                            if (have_prev && col == k - 1 && sub == 0) R[(k - 1) * D + col] = bprev;
                            bprev = beta;
                            have_prev = true;
                        }
                        __syncthreads();
                        if (tid == 0) R[(D - 1) * D + (D - 1)] = bprev;
                        __syncthreads();
                    }
                }
                /* 4. measurement update on the first NR rows: C2 = R^- H' */
                for (int e = tid; e < NR * d; e += NT) {
                    const int r = e / d, a = e - r * d;
                    double s = 0.0;
                    if (!PN_SECOND && !PN_MASS) {
                        s = R[r * D + d + a];
                        for (int i = 0; i < d; i++) s -= R[r * D + i] * J[a + d * i];
                    } else {
                        for (int b = 0; b < HB; b++)
                            for (int i = 0; i < d; i++) s += R[r * D + b * d + i] * Hb(b, a, i);
                    }
                    C2[e] = s;
                }
                __syncthreads();
                /* S = C2'C2 = U_S'U_S (perform_step.jl:182-188, update.jl:100): Gram + Cholesky on register tiles; the
                   tile-packed factor goes to Ts */
                const bool okS = gram_chol_tiles<d>(C2, NR, Ts, sinv, sh, [](int, int) -> double { return 0.0; });
                /* a singular S is either Sigma^- H' == 0 (pass-through, update.jl:82-89) or a breakdown */
                double trS = 0.0;
                for (int e = tid; e < NR * d; e += NT) trS += C2[e] * C2[e];
                trS = block_sum(trS, sh);
                if (trS == 0.0) { /* Sigma^- H' == 0 : pass-through (update.jl:82-89) */
                    for (int c = tid; c < D; c += NT) mu[c] = mup[c];
                    if (tid == 0) {
                        bool zz = true;
                        for (int a = 0; a < d; a++) zz = zz && (z[a] == 0.0);
                        sh->loglik += zz ? PN_INF : -PN_INF;
                    }
                    __syncthreads();
                } else {
                    expand_tiles<d>(Ts, S);
                    for (int i = tid; i < d; i += NT) zt[i] = z[i];
                    __syncthreads();
                    /* K' = S^-1 (C2' R^-)  (update.jl:100-109: K = Sigma^-.R' (Sigma^-.R H') / chol(S)).  First Bm = C2' R^- (d x D,
                       stored where the gain goes), TA rows per thread so that an entry of R^- is reused; then one thread per
                       column solves U_S' U_S k = b in registers (column-oriented substitution: one dependent multiply and one
                       level of independent FMAs per pivot), the factor read as shared-memory broadcasts.  Warp 0 meanwhile
                       solves for z (log-likelihood) on its own. */
                    {
                        constexpr int TA = (d % 7 == 0) ? 7 : 4;
                        constexpr int NAB = (d + TA - 1) / TA;
                        for (int e = tid; e < D * NAB; e += NT) {
                            const int ab = e / D, c = e - ab * D;
                            const int rmax = (c + 1 < NR) ? c + 1 : NR;
                            double acc[TA];
#pragma unroll
                            for (int t = 0; t < TA; t++) acc[t] = 0.0;
                            for (int r = 0; r < rmax; r++) {
                                const double rv = R[r * D + c];
#pragma unroll
                                for (int t = 0; t < TA; t++)
                                    if (ab * TA + t < d) acc[t] += rv * C2[r * d + ab * TA + t];
                            }
#pragma unroll
                            for (int t = 0; t < TA; t++)
                                if (ab * TA + t < d) K[(ab * TA + t) * D + c] = acc[t];
                        }
                    }
                    __syncthreads();
                    if (tid >= NT - 32) { /* last warp: z solve + log-likelihood bookkeeping */
                        if (d <= 32) {
                            chol_solve_warp(S, sinv, zt, d);
                        } else if (tid == NT - 32) {
                            chol_solve_serial(S, sinv, zt, d);
                        }
                        __syncwarp();
                        if (tid == NT - 32) {
                            sh->okflag = okS ? 1 : 0;
                            double quad = 0.0, logdet = 0.0;
                            for (int a = 0; a < d; a++) {
                                quad += z[a] * zt[a];
                                logdet += log(S[a * d + a]);
                            }
                            sh->loglik += -0.5 * quad - 0.5 * (double)d * 1.8378770664093453 - logdet;
                            if (!DYNAMIC) {
                                const double inc = quad * (1.0 / (double)d);
                                sh->gdiff = (sh->naccept == 0) ? inc : sh->gdiff + (inc - sh->gdiff) / (double)sh->naccept;
                            }
                            if (!sh->okflag) sh->retcode = PN_RET_UNSTABLE;
                        }
                    }
                    gain_columns(K, S, sinv, mup, mu, z);
                    __syncthreads();
                    /* R+ = R^- - C2 K' on the first NR rows, TR rows per thread */
                    {
                        constexpr int TR = 4;
                        constexpr int NRB = (NR + TR - 1) / TR;
                        for (int e = tid; e < NRB * D; e += NT) {
                            const int rb = e / D, c = e - rb * D;
                            double acc[TR];
#pragma unroll
                            for (int t = 0; t < TR; t++) acc[t] = 0.0;
                            for (int a = 0; a < d; a++) {
                                const double kv = K[a * D + c];
#pragma unroll
                                for (int t = 0; t < TR; t++)
                                    if (rb * TR + t < NR) acc[t] += C2[(rb * TR + t) * d + a] * kv;
                            }
#pragma unroll
                            for (int t = 0; t < TR; t++)
                                if (rb * TR + t < NR) R[(rb * TR + t) * D + c] -= acc[t];
                        }
                    }
                    __syncthreads();
                }
                /* loopfooter! accept (thread 0) */
                if (tid == 0) {
                    double dtnew = sh->dt, qq = sh->qq;
                    if (ADAPTIVE) {
                        if (P.qsteady_min <= qq && qq <= P.qsteady_max) qq = 1.0;
                        sh->qold = fmax(sh->EEst, P.qoldinit);
                        sh->qold_pb2 = (sh->EEst > P.qoldinit) ? exp(P.beta2 * sh->lEE) : P.qoldinit_pb2;
                        dtnew = sh->dt / qq;
                    }
                    const double ttmp = sh->t + h;
                    const double ref = fmax(fabs(sh->t), fabs(sh->tstop));
                    const double epsref = nextafter(ref, 1e300) - ref;
                    sh->t = (fabs(ttmp - sh->tstop) < 100.0 * epsref) ? sh->tstop : ttmp;
                    sh->dt = dtnew;
                    sh->naccept++;
                    bool nanu = false;
                    for (int i = 0; i < d; i++) {
                        uprev[i] = mu[pn_solcol<d>(i)];
                        nanu = nanu || (mu[i] != mu[i]);
                    }
                    if (nanu) sh->retcode = PN_RET_UNSTABLE;
                    if (sh->retcode != PN_RET_SUCCESS || !(sh->t < P.t1)) sh->finished = 1;
                }
                __syncthreads();
                if (SAVE >= 1) {
                    if (P.step_offsets && sh->save_pos < P.step_offsets[traj + 1]) {
                        const long long sp = sh->save_pos;
                        for (int e = tid; e < DU * DU; e += NT) {
                            const int a = pn_solcol<d>(e / DU), b = pn_solcol<d>(e % DU);
                            double s = 0.0;
                            for (int r = 0; r < D; r++) s += R[r * D + a] * R[r * D + b];
                            P.s_u_cov[sp * DU * DU + e] = s;
                        }
                        for (int i = tid; i < DU; i += NT) P.s_u_mean[sp * DU + i] = mu[pn_solcol<d>(i)];
                        if (tid == 0) {
                            P.s_t[sp] = sh->t;
                            P.s_diffusion[sp - 1] = (SAVE >= 2) ? sh->ediff : ((ADAPTIVE || DYNAMIC) ? sh->ldiff : P.initial_diffusion);
                            if (SAVE >= 2) P.s_h[sp - 1] = h;
                        }
                        if (SAVE >= 2) { /* what the smoother sweep needs: x_filt[k+1] (D^2 + D doubles, coalesced) */
                            double* dst = P.s_x_covfac + sp * (long long)D * D;
                            for (int e = tid; e < D * D; e += NT) dst[e] = R[e];
                            for (int c = tid; c < D; c += NT) P.s_x_mean[sp * D + c] = mu[c];
                        }
                    }
                    __syncthreads();
                    if (tid == 0) sh->save_pos += 1;
                    __syncthreads();
                }
            }
            /* ---- postamble ---- */
            __syncthreads();
            double sc = 1.0, fd = sh->ldiff;
            if (!DYNAMIC) {
                if (P.calibrate && sh->naccept > 0) {
                    sc = sh->gdiff;
                    fd = sh->gdiff * P.initial_diffusion;
                } else {
                    fd = P.initial_diffusion;
                }
            }
            for (int e = tid; e < DU * DU; e += NT) {
                const int a = pn_solcol<d>(e / DU), b = pn_solcol<d>(e % DU);
                double s = 0.0;
                for (int r = 0; r < D; r++) s += R[r * D + a] * R[r * D + b];
                P.u_cov[traj * DU * DU + e] = s * sc;
            }
            for (int i = tid; i < DU; i += NT) P.u_mean[traj * DU + i] = mu[pn_solcol<d>(i)];
            if (P.x_mean)
                for (int c = tid; c < D; c += NT) P.x_mean[traj * D + c] = mu[c];
            if (P.x_covfac) {
                const double ss = sqrt(sc);
                for (int e = tid; e < D * D; e += NT) P.x_covfac[traj * D * D + e] = R[e] * ss;
            }
            if (tid == 0) {
                P.t_end[traj] = sh->t;
                P.loglik[traj] = sh->loglik;
                P.diffusion[traj] = fd;
                P.dt_last[traj] = sh->dt;
                P.naccept[traj] = sh->naccept;
                P.nreject[traj] = sh->nreject;
                P.nf[traj] = sh->nf;
                P.retcode[traj] = sh->retcode;
            }
        }
    }
};

template <class Prob, int Q, bool ADAPTIVE, bool DYNAMIC, int SAVE>
__device__ __forceinline__ void pn_cta_entry(const PnParams& P) {
    PnCta<Prob, Q, ADAPTIVE, DYNAMIC, SAVE>::run(P);
}
#endif /* PN_CTA_CUH */
