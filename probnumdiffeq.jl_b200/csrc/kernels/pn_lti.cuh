// pn_lti.cuh -- discretisation of a linear time-invariant SDE prior with a general drift in the last derivative block: the
// integrated Ornstein-Uhlenbeck process with a vector or matrix rate parameter (src/priors/ioup.jl:81-131).  Warp-cooperative dense
// routines on row-major matrices in a per-warp workspace; shared by the general-prior filter (pn_gen.cuh) and the shared-memory
// smoother sweep (pn_smooth.cuh).
//
// Reference behaviour replaced: make_transition_matrices!(cache, ::IOUP, dt) (src/priors/ioup.jl:115-131), which calls
// FiniteHorizonGramians.exp_and_gram_chol! (third party, not vendored: scaling and doubling of (exp(F t), Gramian(t))) on
// F = shift (x) I_d with F[q-block, q-block] = rate matrix (to_sde, ioup.jl:81-101) and then forms A = P Ah PI, Q = P Qh P.
// The same scheme runs here in the PRECONDITIONED coordinates x~ = P x (src/preconditioning.jl:30-58), where the SDE over
// tau = t/h in [0, 1] has the O(1) drift
//     F~ = superdiag(q, q-1, ..., 1) (x) I_d  +  e_q e_q' (x) (h * rate),      dispersion  e_q (x) I_d:
//   (1) tau0 = 2^-s with ||F~|| tau0 <= 1/4;  A(tau0) by its Taylor series, Q(tau0) = sum_k tau0^(k+1)/(k+1)! M_k with
//       M_0 = L L', M_(k+1) = F~ M_k + M_k F~'  (F~ is applied through its block structure: d D^2 operations per product);
//   (2) s doublings  Q(2 tau) = Q(tau) + A(tau) Q(tau) A(tau)',  A(2 tau) = A(tau)^2;
//   (3) Ah = P^-1 A(1) P,  Qh.R = chol(Q(1)).U P^-1   (test/core/priors.jl:220-223 identities).
//
// NVRTC-compatible: no standard headers.
#ifndef PN_LTI_CUH
#define PN_LTI_CUH
#include "pn_params.h"

#ifndef PN_GEN_WARPS
#define PN_GEN_WARPS 4
#endif
#define PN_GEN_TAYLOR 18

/* doubles of per-warp workspace (host and kernel must agree) */
__host__ __device__ constexpr int pn_gen_ws_doubles(int d, int n) {
    return 10 * (d * n) * (d * n) + 6 * (d * n) + 3 * (d * n) * d + 6 * d * d + 8 * d + 4 * n + 16;
}

#ifndef PN_LTI_DECL_ONLY

__device__ __forceinline__ double pn_gen_wsum(double x) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
__device__ __forceinline__ double pn_gen_wmax(double x) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}

/* C (N x N) = A B   (all row-major N x N; warp-cooperative; C must not alias A or B) */
__device__ void pn_gen_mm(double* C, const double* A, const double* B, int N, int lane) {
    for (int i = lane; i < N * N; i += 32) {
        const int r = i / N, c = i % N;
        double s = 0.0;
        for (int k = 0; k < N; k++) s += A[r * N + k] * B[k * N + c];
        C[i] = s;
    }
    __syncwarp();
}
/* C (N x N) = A B'  */
__device__ void pn_gen_mmt(double* C, const double* A, const double* B, int N, int lane) {
    for (int i = lane; i < N * N; i += 32) {
        const int r = i / N, c = i % N;
        double s = 0.0;
        for (int k = 0; k < N; k++) s += A[r * N + k] * B[c * N + k];
        C[i] = s;
    }
    __syncwarp();
}
/* Y = F~ X with F~ = superdiag(q-a) (x) I_d + e_q e_q' (x) Z  (X, Y: D x D row-major, Z: d x d row-major) */
__device__ void pn_gen_apply_drift(double* Y, const double* X, const double* Z, int d, int n, int lane) {
    const int D = d * n, q = n - 1;
    for (int i = lane; i < D * D; i += 32) {
        const int r = i / D, c = i % D, a = r / d, ii = r % d;
        double s;
        if (a < q) {
            s = (double)(q - a) * X[(r + d) * D + c];
        } else {
            s = 0.0;
            for (int k = 0; k < d; k++) s += Z[ii * d + k] * X[(q * d + k) * D + c];
        }
        Y[i] = s;
    }
    __syncwarp();
}

/* In-place upper Cholesky of the symmetric D x D matrix M (row-major; reads the upper triangle): M <- U with U'U = M, strict lower
   part zeroed.  Returns false on a non-positive pivot (the same on all lanes). */
__device__ bool pn_gen_chol(double* M, int N, int lane) {
    bool ok = true;
    for (int k = 0; k < N; k++) {
        const double piv = M[k * N + k];
        __syncwarp();
        if (!(piv > 0.0)) { ok = false; break; }
        const double rk = sqrt(piv), ik = 1.0 / rk;
        for (int j = k + lane; j < N; j += 32) M[k * N + j] = (j == k) ? rk : M[k * N + j] * ik;
        __syncwarp();
        const int rem = N - k - 1;
        for (int i = lane; i < rem * rem; i += 32) {
            const int r = k + 1 + i / rem, c = k + 1 + i % rem;
            if (c >= r) M[r * N + c] -= M[k * N + r] * M[k * N + c];
        }
        __syncwarp();
    }
    if (ok)
        for (int i = lane; i < N * N; i += 32)
            if (i % N < i / N) M[i] = 0.0;
    __syncwarp();
    return ok;
}

/* Householder QR of the m x N row-major matrix A, warp-cooperative; rows 0..N-1 end up as the upper-triangular factor
   (triangularize!, src/fast_linalg.jl) */
__device__ void pn_gen_qr(double* A, int m, int N, int lane) {
    for (int k = 0; k < N; k++) {
        double part = 0.0;
        for (int r = k + lane; r < m; r += 32) part += A[r * N + k] * A[r * N + k];
        const double total = pn_gen_wsum(part);
        const double alpha = A[k * N + k];
        __syncwarp();
        if (total > 0.0) {
            const double nrm = sqrt(total);
            const double beta = -copysign(nrm, alpha);
            const double vk = alpha - beta;
            const double gam = 1.0 / (total + fabs(alpha) * nrm);
            for (int j = k + 1 + lane; j < N; j += 32) {
                double s = vk * A[k * N + j];
                for (int r = k + 1; r < m; r++) s += A[r * N + k] * A[r * N + j];
                s *= gam;
                A[k * N + j] -= s * vk;
                for (int r = k + 1; r < m; r++) A[r * N + j] -= s * A[r * N + k];
            }
            __syncwarp();
            for (int r = k + lane; r < m; r += 32) A[r * N + k] = (r == k) ? beta : 0.0;
        }
        __syncwarp();
    }
}

/* (Ah, Qc = Qh.R) of one step of length h for the rate matrix Rm (d x d row-major).  At, Qt, T1, T2 are D x D scratch; Zm d x d.
   Returns false if the process-noise covariance is not numerically positive definite. */
__device__ bool pn_gen_transition(double* Ah, double* Qc, const double* Rm, double h, int d, int n, double* At, double* Qt, double* T1,
                                  double* T2, double* Zm, double* PIv, int lane) {
    const int D = d * n, q = n - 1, DD = D * D;
    /* Z = h Rm and its infinity norm */
    double zn = 0.0;
    for (int r = lane; r < d; r += 32) {
        double s = 0.0;
        for (int c = 0; c < d; c++) s += fabs(Rm[r * d + c]);
        zn = fmax(zn, s * fabs(h));
    }
    zn = pn_gen_wmax(zn);
    for (int i = lane; i < d * d; i += 32) Zm[i] = h * Rm[i];
    if (lane == 0) { /* PIv[c] = h^(q-c+1/2)/(q-c)!  (preconditioning.jl:45-54: the diagonal of P^-1) */
        const double sqh = sqrt(h);
        double hp = 1.0;
        for (int k = 0; k < n; k++) {
            PIv[q - k] = hp * sqh;
            hp = hp * h / (double)(k + 1);
        }
    }
    const double nrm = fmax((double)q, zn); /* ||F~||_inf */
    int s = 0;
    double tau0 = 1.0;
    while (nrm * tau0 > 0.25 && s < 60) {
        tau0 *= 0.5;
        s++;
    }
    /* (1) Taylor series at tau0: A = sum (F~ tau0)^k / k!,  Q = sum tau0^(k+1)/(k+1)! M_k */
    for (int i = lane; i < DD; i += 32) {
        const int r = i / D, c = i % D;
        At[i] = (r == c) ? 1.0 : 0.0;
        T1[i] = (r == c) ? 1.0 : 0.0; /* running term of A */
    }
    __syncwarp();
    for (int k = 1; k <= PN_GEN_TAYLOR; k++) {
        pn_gen_apply_drift(T2, T1, Zm, d, n, lane);
        const double f = tau0 / (double)k;
        for (int i = lane; i < DD; i += 32) {
            const double v = T2[i] * f;
            T1[i] = v;
            At[i] += v;
        }
        __syncwarp();
    }
    for (int i = lane; i < DD; i += 32) {
        const int r = i / D, c = i % D;
        const double m0 = (r == c && r >= q * d) ? 1.0 : 0.0; /* L L' = e_q e_q' (x) I_d */
        T1[i] = m0;
        Qt[i] = tau0 * m0;
    }
    __syncwarp();
    double coef = tau0;
    for (int k = 0; k < PN_GEN_TAYLOR; k++) {
        pn_gen_apply_drift(T2, T1, Zm, d, n, lane); /* F~ M_k */
        coef = coef * tau0 / (double)(k + 2);
        for (int i = lane; i < DD; i += 32) {
            const int r = i / D, c = i % D;
            if (c >= r) {
                const double v = T2[r * D + c] + T2[c * D + r]; /* M_(k+1) = F~ M_k + (F~ M_k)' */
                Qc[r * D + c] = v; /* staged in Qc: both T2[r][c] and T2[c][r] are still being read by other lanes */
                Qc[c * D + r] = v;
            }
        }
        __syncwarp();
        for (int i = lane; i < DD; i += 32) {
            T1[i] = Qc[i];
            Qt[i] += coef * Qc[i];
        }
        __syncwarp();
    }
    /* (2) doublings */
    for (int it = 0; it < s; it++) {
        pn_gen_mm(T1, At, Qt, D, lane);  /* A Q    */
        pn_gen_mmt(T2, T1, At, D, lane); /* A Q A' */
        for (int i = lane; i < DD; i += 32) {
            const int r = i / D, c = i % D;
            if (c >= r) {
                const double v = Qt[r * D + c] + 0.5 * (T2[r * D + c] + T2[c * D + r]);
                Qc[r * D + c] = v;
                Qc[c * D + r] = v;
            }
        }
        pn_gen_mm(T1, At, At, D, lane);
        for (int i = lane; i < DD; i += 32) {
            At[i] = T1[i];
            Qt[i] = Qc[i];
        }
        __syncwarp();
    }
    /* (3) factorise and leave the preconditioned coordinates */
    const bool ok = pn_gen_chol(Qt, D, lane);
    for (int i = lane; i < DD; i += 32) {
        const int r = i / D, c = i % D;
        Ah[i] = PIv[r / d] * At[i] / PIv[c / d];
        Qc[i] = (c >= r) ? Qt[i] * PIv[c / d] : 0.0;
    }
    __syncwarp();
    return ok;
}

#endif /* PN_LTI_DECL_ONLY */

#endif /* PN_LTI_CUH */
