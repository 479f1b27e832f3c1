/*
 * pnode_oracle.c -- CPU ORACLE (test infrastructure; see pnode_oracle.h).
 *
 * Plain-C FP64 restatement of ProbNumDiffEq.jl v0.18.0's ODE-filter path.  Citations are
 * to files under the reference checkout (src/...).  Matrices are column-major.
 * Nothing here is shared with, or reachable from, the CUDA product.
 */
#include "pnode_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define AT(A, i, j, ld) ((A)[(size_t)(i) + (size_t)(j) * (size_t)(ld)])

/* ------------------------------------------------------------------------------------------ */
/* small dense helpers                                                                        */
/* ------------------------------------------------------------------------------------------ */

/* C (m x n) = alpha * op(A) * op(B) + beta * C ; op = transpose if t? != 0 ; inner dim k */
static void gemm(int m, int n, int k, double alpha, const double* A, int lda, int tA, const double* B, int ldb,
                 int tB, double beta, double* C, int ldc) {
    for (int j = 0; j < n; j++)
        for (int i = 0; i < m; i++) {
            double s = 0.0;
            for (int l = 0; l < k; l++) {
                double a = tA ? AT(A, l, i, lda) : AT(A, i, l, lda);
                double b = tB ? AT(B, j, l, ldb) : AT(B, l, j, ldb);
                s += a * b;
            }
            AT(C, i, j, ldc) = (beta == 0.0 ? 0.0 : beta * AT(C, i, j, ldc)) + alpha * s;
        }
}

static int all_zero(const double* x, size_t n) {
    for (size_t i = 0; i < n; i++)
        if (x[i] != 0.0) return 0;
    return 1;
}

static double factorial_d(int k) {
    double f = 1.0;
    for (int i = 2; i <= k; i++) f *= (double)i;
    return f;
}
static double binomial_d(int n, int k) {
    if (k < 0 || k > n) return 0.0;
    double r = 1.0;
    for (int i = 1; i <= k; i++) r = r * (double)(n - k + i) / (double)i;
    return floor(r + 0.5);
}

/* Upper Cholesky M = U'U in place (LAPACK potrf 'U' semantics); strictly-lower part zeroed.
   Returns 0 on success, j+1 if the j-th pivot is not positive (issuccess == false). */
int oracle_cholesky_upper(int n, double* M) {
    for (int j = 0; j < n; j++) {
        double s = AT(M, j, j, n);
        for (int k = 0; k < j; k++) s -= AT(M, k, j, n) * AT(M, k, j, n);
        if (!(s > 0.0) || isnan(s)) return j + 1;
        double ujj = sqrt(s);
        AT(M, j, j, n) = ujj;
        for (int i = j + 1; i < n; i++) {
            double v = AT(M, j, i, n);
            for (int k = 0; k < j; k++) v -= AT(M, k, j, n) * AT(M, k, i, n);
            AT(M, j, i, n) = v / ujj;
        }
    }
    for (int j = 0; j < n; j++)
        for (int i = j + 1; i < n; i++) AT(M, i, j, n) = 0.0;
    return 0;
}

/* triangularize!(A) = qr!(A).R  (src/fast_linalg.jl:70-79; LAPACK geqrt! == Householder QR with
   dlarfg conventions; only R is used so blocking (nb=36) does not change the result beyond rounding).
   A is m x n (m >= n), destroyed.  Rout = triu(A[0:n,0:n]) (getupperright!, fast_linalg.jl:54-57). */
void oracle_triangularize(int m, int n, double* A, double* Rout) {
    for (int k = 0; k < n; k++) {
        double alpha = AT(A, k, k, m);
        double xn2 = 0.0;
        for (int i = k + 1; i < m; i++) xn2 += AT(A, i, k, m) * AT(A, i, k, m);
        if (xn2 == 0.0) continue; /* tau = 0, H = I */
        double xnorm = sqrt(xn2);
        double beta = -copysign(hypot(alpha, xnorm), alpha);
        double tau = (beta - alpha) / beta;
        double scal = 1.0 / (alpha - beta);
        for (int i = k + 1; i < m; i++) AT(A, i, k, m) *= scal; /* v (v_k = 1 implicit) */
        AT(A, k, k, m) = beta;
        for (int j = k + 1; j < n; j++) {
            double w = AT(A, k, j, m);
            for (int i = k + 1; i < m; i++) w += AT(A, i, k, m) * AT(A, i, j, m);
            w *= tau;
            AT(A, k, j, m) -= w;
            for (int i = k + 1; i < m; i++) AT(A, i, j, m) -= w * AT(A, i, k, m);
        }
    }
    for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) AT(Rout, i, j, n) = (i <= j) ? AT(A, i, j, m) : 0.0;
}

/* ------------------------------------------------------------------------------------------ */
/* IWP prior + preconditioning                                                                */
/* ------------------------------------------------------------------------------------------ */

/* preconditioned_discretize_1d (src/priors/iwp.jl:96-108): A_breve = binomial.(q:-1:0, (q:-1:0)') ;
   QR_breve = make_preconditioned_iwp_transition_cov_lsqrt_1d (iwp.jl:74-94), lower triangular,
   Q_breve = L' L  (PSDMatrix(R) == R'R, src/filtering/predict.jl:15-16). */
void oracle_iwp_preconditioned(int q, double* A_breve, double* L_breve) {
    int n = q + 1;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            AT(A_breve, i, j, n) = binomial_d(q - i, q - j);
            AT(L_breve, i, j, n) = 0.0;
        }
    for (int m = 0; m <= q; m++)
        for (int c = 0; c <= m; c++) {
            double fqn = factorial_d(q - c);
            AT(L_breve, m, c, n) =
                sqrt((double)(2 * q - 2 * m + 1)) * (fqn * fqn) / factorial_d(m - c) / factorial_d(2 * q - c - m + 1);
        }
}

/* make_preconditioner! / make_preconditioner_inv! (src/preconditioning.jl:30-58), 1-d factors. */
void oracle_preconditioner(int q, double h, double* P, double* PI) {
    double val = factorial_d(q) / pow(h, (double)q + 0.5);
    for (int j = 0; j <= q; j++) {
        P[j] = val;
        val /= (double)(q - j) / h;
    }
    val = pow(h, (double)q + 0.5) / factorial_d(q);
    for (int j = 0; j <= q; j++) {
        PI[j] = val;
        val *= (double)(q - j) / h;
    }
}

/* make_transition_matrices!(cache, ::IWP, dt) (src/priors/iwp.jl:182-190):
   Ah = PI * (A * P) ; Qh.R = Q.R * PI' = L * diag(PI)  (fast_X_A_Xt!, src/fast_linalg.jl:88-91). */
void oracle_iwp_transition(int q, double h, double* Ah, double* QhR) {
    int n = q + 1;
    double A[32 * 32], L[32 * 32], P[32], PI[32];
    oracle_iwp_preconditioned(q, A, L);
    oracle_preconditioner(q, h, P, PI);
    for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) {
            AT(Ah, i, j, n) = PI[i] * (AT(A, i, j, n) * P[j]);
            AT(QhR, i, j, n) = AT(L, i, j, n) * PI[j];
        }
}


/* ------------------------------------------------------------------------------------------ */
/* IOUP prior (src/priors/ioup.jl), scalar rate parameter                                      */
/* ------------------------------------------------------------------------------------------ */

/* Solve A X = B (A n x n, B n x m, column-major) by Gaussian elimination with partial pivoting; B <- X. */
static void lu_solve(int n, int m, double* A, double* B) {
    for (int k = 0; k < n; k++) {
        int piv = k;
        for (int i = k + 1; i < n; i++)
            if (fabs(AT(A, i, k, n)) > fabs(AT(A, piv, k, n))) piv = i;
        if (piv != k) {
            for (int j = 0; j < n; j++) { double t = AT(A, k, j, n); AT(A, k, j, n) = AT(A, piv, j, n); AT(A, piv, j, n) = t; }
            for (int j = 0; j < m; j++) { double t = AT(B, k, j, n); AT(B, k, j, n) = AT(B, piv, j, n); AT(B, piv, j, n) = t; }
        }
        for (int i = k + 1; i < n; i++) {
            double f = AT(A, i, k, n) / AT(A, k, k, n);
            if (f == 0.0) continue;
            for (int j = k + 1; j < n; j++) AT(A, i, j, n) -= f * AT(A, k, j, n);
            for (int j = 0; j < m; j++) AT(B, i, j, n) -= f * AT(B, k, j, n);
        }
    }
    for (int j = 0; j < m; j++)
        for (int i = n - 1; i >= 0; i--) {
            double sacc = AT(B, i, j, n);
            for (int k = i + 1; k < n; k++) sacc -= AT(A, i, k, n) * AT(B, k, j, n);
            AT(B, i, j, n) = sacc / AT(A, i, i, n);
        }
}

/* exp(M), M n x n column-major: scaling and squaring with the degree-13 Pade approximant (Higham 2005; what Julia's
   `exp(::Matrix)` -- used by matrix_fraction_decomposition, src/priors/ltisde.jl:52-63 -- implements). */
void oracle_expm(int n, const double* M, double* E) {
    static const double b[14] = {64764752532480000.0, 32382376266240000.0, 7771770303897600.0, 1187353796428800.0,
                                 129060195264000.0,   10559470521600.0,    670442572800.0,    33522128640.0,
                                 1323241920.0,        40840800.0,          960960.0,          16380.0, 182.0, 1.0};
    size_t nn = (size_t)n * n;
    double nrm = 0.0; /* 1-norm */
    for (int j = 0; j < n; j++) {
        double c = 0.0;
        for (int i = 0; i < n; i++) c += fabs(AT(M, i, j, n));
        if (c > nrm) nrm = c;
    }
    int sq = 0;
    if (nrm > 5.371920351148152) sq = (int)ceil(log2(nrm / 5.371920351148152));
    if (sq < 0) sq = 0;
    double sc = ldexp(1.0, -sq);
    double* A = (double*)malloc(sizeof(double) * nn * 8);
    double *A2 = A + nn, *A4 = A2 + nn, *A6 = A4 + nn, *U = A6 + nn, *V = U + nn, *T1 = V + nn, *T2 = T1 + nn;
    for (size_t i = 0; i < nn; i++) A[i] = M[i] * sc;
    gemm(n, n, n, 1.0, A, n, 0, A, n, 0, 0.0, A2, n);
    gemm(n, n, n, 1.0, A2, n, 0, A2, n, 0, 0.0, A4, n);
    gemm(n, n, n, 1.0, A4, n, 0, A2, n, 0, 0.0, A6, n);
    /* U = A [A6 (b13 A6 + b11 A4 + b9 A2) + b7 A6 + b5 A4 + b3 A2 + b1 I] */
    for (size_t i = 0; i < nn; i++) T1[i] = b[13] * A6[i] + b[11] * A4[i] + b[9] * A2[i];
    gemm(n, n, n, 1.0, A6, n, 0, T1, n, 0, 0.0, T2, n);
    for (size_t i = 0; i < nn; i++) T2[i] += b[7] * A6[i] + b[5] * A4[i] + b[3] * A2[i];
    for (int i = 0; i < n; i++) AT(T2, i, i, n) += b[1];
    gemm(n, n, n, 1.0, A, n, 0, T2, n, 0, 0.0, U, n);
    /* V = A6 (b12 A6 + b10 A4 + b8 A2) + b6 A6 + b4 A4 + b2 A2 + b0 I */
    for (size_t i = 0; i < nn; i++) T1[i] = b[12] * A6[i] + b[10] * A4[i] + b[8] * A2[i];
    gemm(n, n, n, 1.0, A6, n, 0, T1, n, 0, 0.0, V, n);
    for (size_t i = 0; i < nn; i++) V[i] += b[6] * A6[i] + b[4] * A4[i] + b[2] * A2[i];
    for (int i = 0; i < n; i++) AT(V, i, i, n) += b[0];
    /* (V - U) E = V + U */
    for (size_t i = 0; i < nn; i++) { T1[i] = V[i] - U[i]; E[i] = V[i] + U[i]; }
    lu_solve(n, n, T1, E);
    for (int k = 0; k < sq; k++) {
        gemm(n, n, n, 1.0, E, n, 0, E, n, 0, 0.0, T1, n);
        memcpy(E, T1, sizeof(double) * nn);
    }
    free(A);
}

/* matrix_fraction_decomposition (src/priors/ltisde.jl:52-63): exp(dt [F LL'; 0 -F']) = [A *; 0 *], Q = Mexp[1:d, d+1:end] A'. */
void oracle_matrix_fraction_decomposition(int n, const double* F, const double* Lvec /* n, dispersion column */, double dt,
                                          double* A, double* Q) {
    int m = 2 * n;
    double* M = (double*)calloc((size_t)m * m, sizeof(double));
    double* E = (double*)malloc(sizeof(double) * (size_t)m * m);
    for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) {
            AT(M, i, j, m) = dt * AT(F, i, j, n);
            AT(M, i, n + j, m) = dt * Lvec[i] * Lvec[j];
            AT(M, n + i, n + j, m) = -dt * AT(F, j, i, n);
        }
    oracle_expm(m, M, E);
    for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) AT(A, i, j, n) = AT(E, i, j, m);
    for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) {
            double sacc = 0.0;
            for (int k = 0; k < n; k++) sacc += AT(E, i, n + k, m) * AT(A, j, k, n);
            AT(Q, i, j, n) = sacc;
        }
    free(M);
    free(E);
}

/* make_transition_matrices!(cache, ::IOUP, dt) (src/priors/ioup.jl:115-131), 1-d factors, scalar rate r:
   the reference calls FiniteHorizonGramians.exp_and_gram_chol! (v0.2.1, third party, not vendored) on
   F = shift + r e_q e_q', L = e_q over the horizon dt and then forms A = P Ah PI, Q = P Qh P.  The in-repo equivalent that
   the reference's own tests pin it against is matrix_fraction_decomposition (test/core/priors.jl:29-32,266-272).  Here
   the decomposition is evaluated in the *preconditioned* coordinates x~ = P x, where the SDE over tau = t/dt in [0, 1]
   has the O(1) drift F~ = superdiag(q, q-1, ..., 1) + (r dt) e_q e_q' and dispersion e_q (for r = 0 this is exactly the
   IWP's A_breve, Q_breve of iwp.jl:96-108), and mapped back: Ah = PI A~ P, Qh.R = chol(Q~).U PI (preconditioning
   identities of test/core/priors.jl:220-223). */
void oracle_ioup_transition(int q, double rate, double h, double* Ah, double* QhR) {
    int n = q + 1;
    double F[32 * 32], Lv[32], At[32 * 32], Qt[32 * 32], P[32], PI[32];
    memset(F, 0, sizeof(F));
    for (int j = 0; j < q; j++) AT(F, j, j + 1, n) = (double)(q - j);
    AT(F, q, q, n) = rate * h;
    for (int j = 0; j < n; j++) Lv[j] = (j == q) ? 1.0 : 0.0;
    oracle_matrix_fraction_decomposition(n, F, Lv, 1.0, At, Qt);
    for (int j = 0; j < n; j++) /* symmetrise before the factorisation */
        for (int i = 0; i < j; i++) {
            double v = 0.5 * (AT(Qt, i, j, n) + AT(Qt, j, i, n));
            AT(Qt, i, j, n) = v;
            AT(Qt, j, i, n) = v;
        }
    oracle_cholesky_upper(n, Qt);
    oracle_preconditioner(q, h, P, PI);
    for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) {
            AT(Ah, i, j, n) = PI[i] * (AT(At, i, j, n) * P[j]);
            AT(QhR, i, j, n) = (i <= j) ? AT(Qt, i, j, n) * PI[j] : 0.0;
        }
}

/* matrix_fraction_decomposition (src/priors/ltisde.jl:52-63) with a general dispersion: LLt = L L' (n x n). */
static void mfd_general(int n, const double* F, const double* LLt, double dt, double* A, double* Q) {
    int m = 2 * n;
    double* M = (double*)calloc((size_t)m * m, sizeof(double));
    double* E = (double*)malloc(sizeof(double) * (size_t)m * m);
    for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) {
            AT(M, i, j, m) = dt * AT(F, i, j, n);
            AT(M, i, n + j, m) = dt * AT(LLt, i, j, n);
            AT(M, n + i, n + j, m) = -dt * AT(F, j, i, n);
        }
    oracle_expm(m, M, E);
    for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) AT(A, i, j, n) = AT(E, i, j, m);
    for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) {
            double sacc = 0.0;
            for (int k = 0; k < n; k++) sacc += AT(E, i, n + k, m) * AT(A, j, k, n);
            AT(Q, i, j, n) = sacc;
        }
    free(M);
    free(E);
}

/* make_transition_matrices!(cache, ::IOUP, dt) (src/priors/ioup.jl:115-131) for a vector / matrix rate parameter: to_sde
   (ioup.jl:81-101) gives F = kron(shift, I_d) with F[q-block, q-block] = R and L = kron(e_q, I_d) in the derivative-major state
   ordering; the discretisation is the matrix fraction decomposition the reference's own tests pin exp_and_gram_chol! against
   (test/core/priors.jl:29-32, 229-272), evaluated as in oracle_ioup_transition in the preconditioned coordinates:
   F~ = kron(superdiag(q, ..., 1), I_d) + kron(e_q e_q', R dt), dispersion kron(e_q, I_d), horizon 1. */
int oracle_ioup_transition_matrix(int d, int q, const double* rate, double h, double* Ah, double* QhR) {
    int n = q + 1, D = d * n;
    size_t DD = (size_t)D * D;
    double* F = (double*)calloc(DD, sizeof(double));
    double* LLt = (double*)calloc(DD, sizeof(double));
    double* At = (double*)malloc(sizeof(double) * DD);
    double* Qt = (double*)malloc(sizeof(double) * DD);
    double P[32], PI[32];
    for (int a = 0; a < q; a++)
        for (int i = 0; i < d; i++) AT(F, a * d + i, (a + 1) * d + i, D) = (double)(q - a);
    for (int i = 0; i < d; i++) {
        for (int j = 0; j < d; j++) AT(F, q * d + i, q * d + j, D) = AT(rate, i, j, d) * h;
        AT(LLt, q * d + i, q * d + i, D) = 1.0;
    }
    mfd_general(D, F, LLt, 1.0, At, Qt);
    for (int j = 0; j < D; j++)
        for (int i = 0; i < j; i++) {
            double v = 0.5 * (AT(Qt, i, j, D) + AT(Qt, j, i, D));
            AT(Qt, i, j, D) = v;
            AT(Qt, j, i, D) = v;
        }
    int rc = oracle_cholesky_upper(D, Qt);
    oracle_preconditioner(q, h, P, PI);
    for (int j = 0; j < D; j++)
        for (int i = 0; i < D; i++) {
            AT(Ah, i, j, D) = PI[i / d] * (AT(At, i, j, D) * P[j / d]);
            AT(QhR, i, j, D) = (i <= j) ? AT(Qt, i, j, D) * PI[j / d] : 0.0;
        }
    free(F);
    free(LLt);
    free(At);
    free(Qt);
    return rc == 0 ? 0 : -1;
}

/* kron(B, I_d) (src/kronecker.jl:7): out[(r*d+i), (c*d+i)] = B[r,c]. */
void oracle_kron_identity(int d, int n_rows, int n_cols, const double* B, double* out) {
    int M = n_rows * d, N = n_cols * d;
    memset(out, 0, sizeof(double) * (size_t)M * (size_t)N);
    for (int r = 0; r < n_rows; r++)
        for (int c = 0; c < n_cols; c++)
            for (int i = 0; i < d; i++) AT(out, r * d + i, c * d + i, M) = AT(B, r, c, n_rows);
}

/* ------------------------------------------------------------------------------------------ */
/* filtering components (generic size N; EK1 dense: N = D, EK0 Kronecker: N = q+1 on factors)  */
/* ------------------------------------------------------------------------------------------ */

/* how often predict_cov!'s Cholesky (predict.jl:85) succeeded / fell back to triangularize! (:90); diagnostics only */
static long long g_cov_total = 0, g_cov_fallback = 0;
void oracle_cov_stats(long long* total, long long* fallback, int reset) {
    if (total) *total = g_cov_total;
    if (fallback) *fallback = g_cov_fallback;
    if (reset) g_cov_total = g_cov_fallback = 0;
}

/* predict_cov! (src/filtering/predict.jl:62-94). */
void oracle_predict_cov(int N, const double* R, const double* Ah, const double* QhR, double diffusion, int backend,
                        double* Rout, int* used_qr) {
    if (used_qr) *used_qr = 0;
    if (diffusion == 0.0) { /* :71-74 */
        gemm(N, N, N, 1.0, R, N, 0, Ah, N, 1, 0.0, Rout, N);
        return;
    }
    double* S = (double*)malloc(sizeof(double) * 2 * N * N); /* C_2DxD */
    double* M = (double*)malloc(sizeof(double) * N * N);     /* C_DxD  */
    double* top = (double*)malloc(sizeof(double) * N * N);
    gemm(N, N, N, 1.0, R, N, 0, Ah, N, 1, 0.0, top, N); /* :78 */
    double s = (diffusion == 1.0) ? 1.0 : sqrt(diffusion);
    for (int j = 0; j < N; j++)
        for (int i = 0; i < N; i++) {
            AT(S, i, j, 2 * N) = AT(top, i, j, N);
            AT(S, N + i, j, 2 * N) = (diffusion == 1.0) ? AT(QhR, i, j, N) : AT(QhR, i, j, N) * s; /* :79-83 */
        }
    int ok = 0;
    if (backend == ORACLE_COV_REF) {
        gemm(N, N, 2 * N, 1.0, S, 2 * N, 1, S, 2 * N, 0, 0.0, M, N); /* :84 */
        ok = (oracle_cholesky_upper(N, M) == 0);                     /* :85 */
        if (ok) memcpy(Rout, M, sizeof(double) * N * N);
#pragma omp atomic
        g_cov_total += 1;
#pragma omp atomic
        g_cov_fallback += ok ? 0 : 1;
    }
    if (!ok) { /* :90 */
        oracle_triangularize(2 * N, N, S, Rout);
        if (used_qr) *used_qr = 1;
    }
    free(S);
    free(M);
    free(top);
}

/* Solve X * (U'U) = B in place for X (rdiv!(B, Cholesky(U,'U'))), B is m x n, U upper n x n
   (only the upper triangle of U is read, as LAPACK trsm does). */
static void rdiv_chol_upper(int m, int n, double* B, const double* U) {
    /* X1 = B U^{-1} */
    for (int j = 0; j < n; j++) {
        for (int k = 0; k < j; k++) {
            double ukj = AT(U, k, j, n);
            if (ukj != 0.0)
                for (int i = 0; i < m; i++) AT(B, i, j, m) -= AT(B, i, k, m) * ukj;
        }
        double ujj = AT(U, j, j, n);
        for (int i = 0; i < m; i++) AT(B, i, j, m) /= ujj;
    }
    /* X = X1 U^{-T} */
    for (int j = n - 1; j >= 0; j--) {
        double ujj = AT(U, j, j, n);
        for (int i = 0; i < m; i++) AT(B, i, j, m) /= ujj;
        for (int k = 0; k < j; k++) {
            double ukj = AT(U, k, j, n);
            if (ukj != 0.0)
                for (int i = 0; i < m; i++) AT(B, i, k, m) -= AT(B, i, j, m) * ukj;
        }
    }
}

/* Generic update! (src/filtering/update.jl:65-139) with measurement covariance from
   compute_measurement_covariance! (src/perform_step.jl:182-188).
   State mean is an N x nrhs "matrix" addressed mu[r*rs + c*cs]; z is m x nrhs addressed z[a*zs_a + c*zs_c].
   Dense EK1: N=D, m=d, nrhs=1.  Kronecker EK0: N=q+1, m=1, nrhs=d, rs=d, cs=1 (update.jl:151-195).
   Returns 0 ok, 1 if S is not positive definite (Julia would throw PosDefException). */
static int update_generic(int N, int m, int nrhs, const double* mu_p, int rs, int cs, const double* R_p,
                          const double* z, int zs_a, int zs_c, const double* H /* m x N */, double* mu_out,
                          double* R_out, double* loglik, const double* Rn /* m x m upper factor of pn_observation_noise, or NULL */) {
    if (all_zero(R_p, (size_t)N * N)) { /* :82-89 */
        for (int r = 0; r < N; r++)
            for (int c = 0; c < nrhs; c++) mu_out[r * rs + c * cs] = mu_p[r * rs + c * cs];
        memcpy(R_out, R_p, sizeof(double) * N * N);
        int zz = 1;
        for (int a = 0; a < m; a++)
            for (int c = 0; c < nrhs; c++)
                if (z[a * zs_a + c * zs_c] != 0.0) zz = 0;
        *loglik = zz ? INFINITY : -INFINITY;
        return 0;
    }
    double* K1 = (double*)malloc(sizeof(double) * N * m);
    double* K = (double*)malloc(sizeof(double) * N * m);
    double* S = (double*)malloc(sizeof(double) * m * m);
    double* Mc = (double*)malloc(sizeof(double) * N * N);
    double* zt = (double*)malloc(sizeof(double) * m * nrhs);
    gemm(N, m, N, 1.0, R_p, N, 0, H, m, 1, 0.0, K1, N);  /* C_Dxd = R_p H'  (perform_step.jl:183, update.jl:101) */
    gemm(m, m, N, 1.0, K1, N, 1, K1, N, 0, 0.0, S, m);   /* S = C' C        (perform_step.jl:184) */
    gemm(N, m, N, 1.0, R_p, N, 1, K1, N, 0, 0.0, K, N);  /* K2 = R_p' K1    (update.jl:102) */
    if (Rn) /* + R.R'R.R (perform_step.jl:185-187) */
        gemm(m, m, m, 1.0, Rn, m, 1, Rn, m, 0, 1.0, S, m);
    double logdet;
    int fail = 0;
    if (m == 1) { /* S_chol = S[1] (update.jl:108) */
        double s = S[0];
        for (int r = 0; r < N; r++) K[r] /= s;
        for (int c = 0; c < nrhs; c++) zt[c] = z[c * zs_c] / s;
        logdet = log(s);
    } else {
        if (oracle_cholesky_upper(m, S) != 0) fail = 1;
        rdiv_chol_upper(N, m, K, S); /* rdiv!(K, S_chol) :109 */
        /* ldiv!(S_chol, z): solve (U'U) x = z for each rhs */
        for (int c = 0; c < nrhs; c++) {
            double* x = zt + (size_t)c * m;
            for (int a = 0; a < m; a++) x[a] = z[a * zs_a + c * zs_c];
            for (int a = 0; a < m; a++) { /* U' y = z */
                for (int k = 0; k < a; k++) x[a] -= AT(S, k, a, m) * x[k];
                x[a] /= AT(S, a, a, m);
            }
            for (int a = m - 1; a >= 0; a--) { /* U x = y */
                for (int k = a + 1; k < m; k++) x[a] -= AT(S, a, k, m) * x[k];
                x[a] /= AT(S, a, a, m);
            }
        }
        logdet = 0.0;
        for (int a = 0; a < m; a++) logdet += log(AT(S, a, a, m));
        logdet *= 2.0;
    }
    /* pn_logpdf! (update.jl:140-148): d = length(reshape(mu,:)) = m*nrhs; logdet NOT multiplied by nrhs */
    double quad = 0.0;
    for (int c = 0; c < nrhs; c++)
        for (int a = 0; a < m; a++) quad += z[a * zs_a + c * zs_c] * zt[(size_t)c * m + a];
    *loglik = -0.5 * quad - 0.5 * (double)(m * nrhs) * log(2.0 * M_PI) - 0.5 * logdet;
    /* mu = m_p - K z (:115) */
    for (int r = 0; r < N; r++)
        for (int c = 0; c < nrhs; c++) {
            double s = 0.0;
            for (int a = 0; a < m; a++) s += AT(K, r, a, N) * z[a * zs_a + c * zs_c];
            mu_out[r * rs + c * cs] = mu_p[r * rs + c * cs] - s;
        }
    /* M = I - K H (:118-121) ; R_out = R_p M' (:123, fast_X_A_Xt!) */
    gemm(N, N, m, -1.0, K, N, 0, H, m, 0, 0.0, Mc, N);
    for (int i = 0; i < N; i++) AT(Mc, i, i, N) += 1.0;
    gemm(N, N, N, 1.0, R_p, N, 0, Mc, N, 1, 0.0, R_out, N);
    if (Rn) { /* update.jl:125-136: Sigma = R_out'R_out + K R K'; cholesky, or triangularize!([R_out; (K R.R')']) if it fails */
        gemm(N, m, m, 1.0, K, N, 0, Rn, m, 1, 0.0, K1, N);          /* K1 = K R.R' */
        gemm(N, N, N, 1.0, R_out, N, 1, R_out, N, 0, 0.0, Mc, N);
        gemm(N, N, m, 1.0, K1, N, 0, K1, N, 1, 1.0, Mc, N);
        if (oracle_cholesky_upper(N, Mc) == 0) {
            memcpy(R_out, Mc, sizeof(double) * N * N);
        } else {
            double* St = (double*)malloc(sizeof(double) * (N + m) * N);
            for (int c = 0; c < N; c++) {
                for (int r = 0; r < N; r++) AT(St, r, c, N + m) = AT(R_out, r, c, N);
                for (int a = 0; a < m; a++) AT(St, N + a, c, N + m) = AT(K1, c, a, N);
            }
            oracle_triangularize(N + m, N, St, R_out);
            free(St);
        }
    }
    free(K1);
    free(K);
    free(S);
    free(Mc);
    free(zt);
    return fail;
}

int oracle_update(int D, int d, const double* mu_p, const double* R_p, const double* z, const double* H,
                  double* mu_out, double* R_out, double* loglik) {
    return update_generic(D, d, 1, mu_p, 1, 0, R_p, z, 1, 0, H, mu_out, R_out, loglik, NULL);
}

/* compute_backward_kernel! (src/filtering/markov_kernel.jl:190-235; Kronecker :237-272). */
static void backward_kernel_generic(int N, int nrhs, int rs, int cs, const double* mu, const double* R,
                                    const double* mu_pred, const double* R_pred, const double* Ah,
                                    const double* QhR, double diffusion, double* G, double* b, double* LR) {
    double* C = (double*)malloc(sizeof(double) * N * N);
    gemm(N, N, N, 1.0, R, N, 0, Ah, N, 1, 0.0, C, N); /* C_DxD = x.Sigma.R A' (:210) */
    gemm(N, N, N, 1.0, R, N, 1, C, N, 0, 0.0, G, N);  /* G = x.Sigma.R' C     (:211) */
    if (!all_zero(G, (size_t)N * N)) rdiv_chol_upper(N, N, G, R_pred); /* :212-214 */
    for (int r = 0; r < N; r++) /* b = mu - G mu_pred (:217-218) */
        for (int c = 0; c < nrhs; c++) {
            double s = 0.0;
            for (int k = 0; k < N; k++) s += AT(G, r, k, N) * mu_pred[k * rs + c * cs];
            b[r * rs + c * cs] = mu[r * rs + c * cs] - s;
        }
    gemm(N, N, N, -1.0, Ah, N, 1, G, N, 1, 0.0, C, N); /* C = -A' G' (:221) */
    for (int i = 0; i < N; i++) AT(C, i, i, N) += 1.0; /* :222-224 */
    double* top = (double*)malloc(sizeof(double) * N * N);
    double* bot = (double*)malloc(sizeof(double) * N * N);
    gemm(N, N, N, 1.0, R, N, 0, C, N, 0, 0.0, top, N); /* Lambda.R[1:D] = R (I - G A)' (:225) */
    if (diffusion == 1.0) {
        gemm(N, N, N, 1.0, QhR, N, 0, G, N, 1, 0.0, bot, N); /* :228 */
    } else {
        double s = sqrt(diffusion);
        for (int i = 0; i < N * N; i++) C[i] = QhR[i] * s; /* apply_diffusion! :230 */
        gemm(N, N, N, 1.0, C, N, 0, G, N, 1, 0.0, bot, N); /* :231 */
    }
    for (int j = 0; j < N; j++)
        for (int i = 0; i < N; i++) {
            AT(LR, i, j, 2 * N) = AT(top, i, j, N);
            AT(LR, N + i, j, 2 * N) = AT(bot, i, j, N);
        }
    free(C);
    free(top);
    free(bot);
}

void oracle_backward_kernel(int D, const double* mu, const double* R, const double* mu_pred,
                            const double* R_pred, const double* Ah, const double* QhR, double diffusion,
                            double* G, double* b, double* LR) {
    backward_kernel_generic(D, 1, 1, 0, mu, R, mu_pred, R_pred, Ah, QhR, diffusion, G, b, LR);
}

/* marginalize! (src/filtering/markov_kernel.jl:68-101; Kronecker :103-121). */
static void marginalize_generic(int N, int nrhs, int rs, int cs, const double* mu_next, const double* R_next,
                                const double* G, const double* b, const double* LR, double* mu_out,
                                double* R_out) {
    for (int r = 0; r < N; r++) /* marginalize_mean! :72-82 */
        for (int c = 0; c < nrhs; c++) {
            double s = 0.0;
            for (int k = 0; k < N; k++) s += AT(G, r, k, N) * mu_next[k * rs + c * cs];
            mu_out[r * rs + c * cs] = s + b[r * rs + c * cs];
        }
    double* S = (double*)malloc(sizeof(double) * 3 * N * N); /* C_3DxD */
    double* top = (double*)malloc(sizeof(double) * N * N);
    gemm(N, N, N, 1.0, R_next, N, 0, G, N, 1, 0.0, top, N); /* :95 */
    for (int j = 0; j < N; j++) {
        for (int i = 0; i < N; i++) AT(S, i, j, 3 * N) = AT(top, i, j, N);
        for (int i = 0; i < 2 * N; i++) AT(S, N + i, j, 3 * N) = AT(LR, i, j, 2 * N); /* :96 */
    }
    oracle_triangularize(3 * N, N, S, R_out); /* :98 */
    free(S);
    free(top);
}

void oracle_marginalize(int D, const double* mu_next, const double* R_next, const double* G, const double* b,
                        const double* LR, double* mu_out, double* R_out) {
    marginalize_generic(D, 1, 1, 0, mu_next, R_next, G, b, LR, mu_out, R_out);
}

/* ------------------------------------------------------------------------------------------ */
/* growable per-trajectory storage                                                             */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    double* p;
    size_t len, cap;
} dvec;
static void dvec_push(dvec* v, const double* src, size_t n) {
    if (v->len + n > v->cap) {
        size_t nc = v->cap ? v->cap * 2 : 1024;
        while (nc < v->len + n) nc *= 2;
        v->p = (double*)realloc(v->p, nc * sizeof(double));
        v->cap = nc;
    }
    memcpy(v->p + v->len, src, n * sizeof(double));
    v->len += n;
}

/* ------------------------------------------------------------------------------------------ */
/* the solver loop                                                                             */
/* ------------------------------------------------------------------------------------------ */

static double rms_norm(const double* x, int n) { /* ODE_DEFAULT_NORM (DiffEqBase, 3rd party) */
    double s = 0.0;
    for (int i = 0; i < n; i++) s += x[i] * x[i];
    return sqrt(s / (double)n);
}

/* ode_determine_initdt (OrdinaryDiffEqCore, 3rd party; restated from SURVEY.md A.6 -- parity unpinned). */
/* Mass matrix M (column-major d x d).  `singular`: Gaussian elimination with partial pivoting meets an exactly zero pivot
   (ArrayInterface.issingular = !issuccess(lu(M))).  `MMinv` = inverse of MM = M (+ 1e-20 I if any diagonal entry of M is zero,
   src/initialization/autodiffinit.jl:20-31).  Conditioning the standard-normal initial state on E_0 x = u0 and
   MM E_o x = df_o, o = 1..q (autodiffinit.jl:33-47, common.jl:122-139) gives the mean blocks MM^-1 df_o and a covariance
   that is zero up to rounding; the restatement takes it as exactly zero, as for M = I. */
static void mass_setup(int d, const double* M, double* MMinv, int* singular) {
    double A[64 * 64], B[64 * 64];
    memcpy(A, M, sizeof(double) * d * d);
    *singular = 0;
    for (int k = 0; k < d && !*singular; k++) {
        int piv = k;
        for (int i = k + 1; i < d; i++)
            if (fabs(AT(A, i, k, d)) > fabs(AT(A, piv, k, d))) piv = i;
        if (AT(A, piv, k, d) == 0.0) { *singular = 1; break; }
        if (piv != k)
            for (int j = 0; j < d; j++) { double t = AT(A, k, j, d); AT(A, k, j, d) = AT(A, piv, j, d); AT(A, piv, j, d) = t; }
        for (int i = k + 1; i < d; i++) {
            double f = AT(A, i, k, d) / AT(A, k, k, d);
            for (int j = k + 1; j < d; j++) AT(A, i, j, d) -= f * AT(A, k, j, d);
        }
    }
    memcpy(A, M, sizeof(double) * d * d);
    int zero_diag = 0;
    for (int i = 0; i < d; i++) zero_diag |= (AT(M, i, i, d) == 0.0);
    if (zero_diag)
        for (int i = 0; i < d; i++) AT(A, i, i, d) += 1e-20;
    memset(B, 0, sizeof(double) * d * d);
    for (int i = 0; i < d; i++) AT(B, i, i, d) = 1.0;
    lu_solve(d, d, A, B);
    memcpy(MMinv, B, sizeof(double) * d * d);
}

/* With a mass matrix M != I the initial step is max(smalldt, dtmin) = 1e-6 and no f evaluation is counted:
   a singular M makes the integrator a DAE integrator (`isdae`, early return); for a non-singular M upstream solves
   M \ f0 through `integrator.alg.linsolve` inside a try block, a field the EK algorithms do not have, so the catch
   branch returns the same value.  (3rd party, recalled; consistent with test/mass_matrix.jl:15-16 -- rtol 1e-10 with
   M = -100 I -- which the Hairer-Wanner step of 0.1 would miss by four orders of magnitude with this restatement.) */
static double initial_dt(const oracle_config* c, oracle_rhs_fn f, const double* u0, const double* p, int order, int mass) {
    int d = c->second_order ? 2 * c->d : c->d; /* the ArrayPartition (du, u) of a second-order problem */
    double sk[64], tmp[64], f0[64], f1[64], u1[64];
    double t = c->t0;
    double dtmax = c->dtmax;
    double dtmin = nextafter(fmax(c->dtmin, nextafter(fabs(t), INFINITY) - fabs(t)), INFINITY);
    double smalldt = fmax(dtmin, 1e-6);
    if (mass) return fmax(smalldt, dtmin);
    for (int i = 0; i < d; i++) sk[i] = c->abstol + fabs(u0[i]) * c->reltol;
    for (int i = 0; i < d; i++) tmp[i] = u0[i] / sk[i];
    double d0 = rms_norm(tmp, d);
    f(f0, u0, p, t);
    for (int i = 0; i < d; i++) tmp[i] = f0[i] / sk[i];
    double d1 = rms_norm(tmp, d);
    double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? smalldt : (d0 / d1) / 100.0;
    dt0 = fmin(dt0, dtmax);
    if (dt0 < 10.0 * 2.220446049250313e-16) return fmax(smalldt, dtmin);
    for (int i = 0; i < d; i++) u1[i] = u0[i] + dt0 * f0[i];
    f(f1, u1, p, t + dt0);
    for (int i = 0; i < d; i++) tmp[i] = (f1[i] - f0[i]) / sk[i];
    double d2 = rms_norm(tmp, d) / dt0;
    double mx = fmax(d1, d2);
    double dt1 = (mx <= 1e-15) ? fmax(smalldt, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(mx)) / (double)order);
    return fmax(dtmin, fmin(fmin(100.0 * dt0, dt1), dtmax));
}


/* ------------------------------------------------------------------------------------------ */
/* BlocksOfDiagonals layout (src/blocksofdiagonals.jl:15-52): d independent (q+1) x (q+1) factors, */
/* block i <-> state indices i:d:D.  Selected for EK0 + multivariate diffusion and for DiagonalEK1 */
/* (src/algorithms.jl:132-151).  The per-block routines are the generic ones above with N = q+1.   */
/* ------------------------------------------------------------------------------------------ */
static double data_update_factor(int n, int nc, int rs, const double* u, int us, double sn, double* mu, double* R);
static int solve_blocks(const oracle_config* c, oracle_rhs_fn f, oracle_jac_fn jac, oracle_taylor_fn taylor,
                        const double* u0, const double* p, oracle_traj* out) {
    const int d = c->d, q = c->q, n = q + 1, D = d * n;
    const int diag1 = (c->alg == ORACLE_DIAGONAL_EK1);
    const int so = c->second_order ? 1 : 0, du = so ? 2 * d : d, e1 = so ? 2 : 1; /* second-order problems: see oracle_solve */
    const int mv = (c->diffusion == ORACLE_DIFF_DYNAMIC_MV || c->diffusion == ORACLE_DIFF_FIXED_MV);
    const int dynamic = (c->diffusion == ORACLE_DIFF_DYNAMIC || c->diffusion == ORACLE_DIFF_DYNAMIC_MV);
    memset(out, 0, sizeof(*out));
    out->d = d; out->q = q; out->D = D; out->smooth = c->smooth; out->alg = c->alg;
    out->global_diffusion = NAN;
    if (du > 64 || n > 32) return -1;
    if (c->alg == ORACLE_EK1) return -2; /* EK1 + MV diffusion is rejected by the reference (algorithms.jl:93-107) */
    const size_t nn = (size_t)n * n, NB = nn * d; /* all blocks of one state */
    double* Ah1 = (double*)malloc(sizeof(double) * nn);
    double* Qh1 = (double*)malloc(sizeof(double) * nn);
    double* mu = (double*)calloc(D, sizeof(double));
    double* R = (double*)calloc(NB, sizeof(double));
    double* mu_pred = (double*)calloc(D, sizeof(double));
    double* R_pred = (double*)calloc(NB, sizeof(double));
    double* mu_filt = (double*)calloc(D, sizeof(double));
    double* R_filt = (double*)calloc(NB, sizeof(double));
    double* H = (double*)calloc((size_t)d * n, sizeof(double)); /* block i: 1 x n at H + i*n */
    double* J = (double*)calloc((size_t)du * du, sizeof(double));
    double* G = (double*)malloc(sizeof(double) * NB);
    double* bb = (double*)malloc(sizeof(double) * D);
    double* LR = (double*)malloc(sizeof(double) * 2 * NB);
    double* Qs = (double*)malloc(sizeof(double) * nn);
    double z[64], fu[64], uprev[64], upred[64], err[64], HQH[64], Cq[32];
    double local_diffusion[64], global_diffusion[64], default_diffusion[64], ediff[64];
    for (int i = 0; i < d; i++) {
        local_diffusion[i] = NAN;
        global_diffusion[i] = NAN;
        default_diffusion[i] = c->initial_diffusion;
    }
    dvec ts = {0}, diffs = {0}, xmu = {0}, xR = {0}, bG = {0}, bB = {0}, bL = {0};

    if (so) {
        double* t2 = (double*)malloc(sizeof(double) * (size_t)du * n);
        taylor(t2, u0, p, c->t0, q);
        for (int o = 0; o <= q; o++)
            for (int i = 0; i < d; i++) mu[o * d + i] = t2[o * du + d + i];
        free(t2);
    } else {
        taylor(mu, u0, p, c->t0, q);
    }
    out->nf += q;
    for (int i = 0; i < d; i++) uprev[i] = so ? mu[d + i] : mu[i];
    const double* Mm = c->mass_matrix; /* diagonal (checked by the caller) */
    if (Mm) {
        int zero_diag = 0;
        for (int i = 0; i < d; i++) zero_diag |= (AT(Mm, i, i, d) == 0.0);
        for (int o = 1; o <= q; o++) /* x_o = MM^-1 df_o, MM = M (+ 1e-20 I) (autodiffinit.jl:20-47) */
            for (int i = 0; i < d; i++) mu[o * d + i] = (1.0 / (AT(Mm, i, i, d) + (zero_diag ? 1e-20 : 0.0))) * mu[o * d + i];
    }
    double t = c->t0, dt;
    if (c->adaptive && !(c->dt > 0.0)) {
        dt = initial_dt(c, f, u0, p, q, Mm != NULL);
        if (!Mm) out->nf += 2;
    } else {
        dt = c->dt;
    }
    const double dtcache = dt;
    double qold = c->qoldinit;
    int64_t iter = 0, success_iter = 0, tstop_idx = 0;
    int retcode = ORACLE_RET_SUCCESS;
    double loglik_total = 0.0;
    dvec_push(&ts, &t, 1);
    dvec_push(&xmu, mu, D);
    dvec_push(&xR, R, NB);

    while (t < c->t1) {
        double tstop = c->t1;
        while (tstop_idx < c->n_tstops && c->tstops[tstop_idx] <= t) tstop_idx++;
        if (tstop_idx < c->n_tstops && c->tstops[tstop_idx] < c->t1) tstop = c->tstops[tstop_idx];
        iter++;
        if (iter > c->maxiters) { retcode = ORACLE_RET_MAXITERS; break; }
        if (c->adaptive) {
            dt = fmin(dt, c->dtmax);
            dt = fmax(dt, c->dtmin);
            dt = fmin(dt, tstop - t);
        } else {
            dt = fmin(dtcache, tstop - t);
        }
        if (isnan(dt) || !(t + dt > t) || (c->adaptive && dt <= c->dtmin)) { retcode = ORACLE_RET_DTLESSTHANMIN; break; }
        double tnew = t + dt;
        oracle_iwp_transition(q, dt, Ah1, Qh1); /* every block shares Ah, Qh (iwp.jl:165-171 on the Blocks layout) */
        for (int i = 0; i < d; i++) /* predict_mean! */
            for (int r = 0; r < n; r++) {
                double s = 0.0;
                for (int k = 0; k < n; k++) s += AT(Ah1, r, k, n) * mu[k * d + i];
                mu_pred[r * d + i] = s;
            }
        for (int i = 0; i < d; i++) upred[i] = mu_pred[i];
        if (so)
            for (int i = 0; i < d; i++) { upred[i] = mu_pred[d + i]; upred[d + i] = mu_pred[i]; }
        f(fu, upred, p, tnew);
        out->nf += 1;
        for (int i = 0; i < d; i++) z[i] = (Mm ? AT(Mm, i, i, d) : 1.0) * mu_pred[e1 * d + i] - fu[i];
        /* calc_H! (derivative_utils.jl:1-53): EK0 H = M E1; DiagonalEK1 H = M E1 - Diagonal(diag(J)) E0 */
        if (diag1) {
            jac(J, upred, p, tnew);
            out->njac += 1;
        }
        for (int i = 0; i < d; i++) {
            double* Hi = H + (size_t)i * n;
            for (int k = 0; k < n; k++) Hi[k] = 0.0;
            Hi[e1] = Mm ? AT(Mm, i, i, d) : 1.0;
            if (diag1 && !so) Hi[0] = -AT(J, i, i, d);
            if (diag1 && so) { /* ddu_diag = [Jdu_ii Ju_ii] on the (E1; E0) rows of block i (derivative_utils.jl:20-27) */
                Hi[1] = -AT(J, i, i, du);
                Hi[0] = -AT(J, i, d + i, du);
            }
        }
        if (c->adaptive || dynamic) {
            /* HQH block i = |Qh.R H_i'|^2 ; scalar: invquad(z, HQH)/d (calibration.jl:21-29,123-133);
               multivariate: z_i^2 / Q11_i (local_diagonal_diffusion, calibration.jl:151-180) */
            double qd = 0.0;
            for (int i = 0; i < d; i++) {
                const double* Hi = H + (size_t)i * n;
                double e = 0.0;
                for (int r = 0; r < n; r++) {
                    double cq = 0.0;
                    for (int k = 0; k < n; k++) cq += AT(Qh1, r, k, n) * Hi[k];
                    e += cq * cq;
                }
                HQH[i] = e;
                qd += z[i] * (z[i] / e);
            }
            for (int i = 0; i < d; i++) local_diffusion[i] = mv ? z[i] * z[i] / HQH[i] : qd / (double)d;
        }
        double EEst = 0.0;
        if (c->adaptive) { /* estimate_errors! (perform_step.jl:240-259): block i of (Qh.R sqrt(s_i)) H_i' */
            for (int i = 0; i < d; i++) {
                const double* Hi = H + (size_t)i * n;
                double s = sqrt(local_diffusion[i]);
                for (size_t k = 0; k < nn; k++) Qs[k] = Qh1[k] * s;
                double e = 0.0;
                for (int r = 0; r < n; r++) {
                    Cq[r] = 0.0;
                    for (int k = 0; k < n; k++) Cq[r] += AT(Qs, r, k, n) * Hi[k];
                    e += Cq[r] * Cq[r];
                }
                err[i] = sqrt(e) * dt;
                err[i] = err[i] / (c->abstol + fmax(fabs(upred[i]), fabs(uprev[i])) * c->reltol);
            }
            EEst = rms_norm(err, d);
            if (isnan(EEst)) { retcode = ORACLE_RET_UNSTABLE; break; }
        }
        int accept = (!c->adaptive) || (EEst < 1.0);
        double q11 = 0.0, qq = 1.0;
        if (c->adaptive) {
            if (EEst == 0.0) {
                qq = 1.0 / c->qmax;
            } else {
                q11 = pow(EEst, c->beta1);
                qq = q11 / pow(qold, c->beta2);
                qq = fmax(1.0 / c->qmax, fmin(1.0 / c->qmin, qq / c->gamma));
            }
        }
        if (!accept) {
            out->nreject++;
            dt = dt / fmin(1.0 / c->qmin, q11 / c->gamma);
            continue;
        }
        double ll_step = 0.0, S[64];
        int ufail = 0;
        for (int i = 0; i < d; i++) {
            ediff[i] = dynamic ? local_diffusion[i] : default_diffusion[i];
            const double* Hi = H + (size_t)i * n;
            oracle_predict_cov(n, R + i * nn, Ah1, Qh1, ediff[i], c->cov_backend, R_pred + i * nn, NULL); /* predict.jl:120-141 */
            if (c->smooth) /* markov_kernel.jl:274-323 */
                backward_kernel_generic(n, 1, d, 0, mu + i, R + i * nn, mu_pred + i, R_pred + i * nn, Ah1, Qh1, ediff[i],
                                        G + i * nn, bb + i, LR + i * 2 * nn);
            double e = 0.0; /* S_i = |R^-_i H_i'|^2 (perform_step.jl:182-188) */
            for (int r = 0; r < n; r++) {
                double cq = 0.0;
                for (int k = 0; k < n; k++) cq += AT(R_pred + i * nn, r, k, n) * Hi[k];
                e += cq * cq;
            }
            S[i] = e;
            double ll;
            /* blocks layout: the observation noise is Diagonal (to_factorized_matrix, covariance_structure.jl:46-58), block i takes
               the 1 x 1 factor R.R[i][i] */
            const double rni = c->pn_obs_Rn ? c->pn_obs_Rn[(size_t)i * d + i] : 0.0;
            if (c->pn_obs_Rn) S[i] += rni * rni;
            ufail |= update_generic(n, 1, 1, mu_pred + i, d, 0, R_pred + i * nn, z + i, 1, 0, Hi, mu_filt + i, R_filt + i * nn,
                                    &ll, c->pn_obs_Rn ? &rni : NULL); /* update.jl:197-239 */
            ll_step += ll;
        }
        if (ufail) { retcode = ORACLE_RET_UNSTABLE; break; }
        loglik_total += ll_step;
        if (!dynamic) {
            if (mv) { /* estimate_global_diffusion(::FixedMVDiffusion) (calibration.jl:88-107): z_j^2 / S[1,1] */
                for (int j = 0; j < d; j++) {
                    double inc = z[j] * z[j] / S[0];
                    global_diffusion[j] = (success_iter == 0) ? inc : global_diffusion[j] + (inc - global_diffusion[j]) / (double)success_iter;
                }
            } else { /* FixedDiffusion: invquad(z, S)/d with S BlocksOfDiagonals (calibration.jl:21-29,49-66) */
                double qd = 0.0;
                for (int i = 0; i < d; i++) qd += z[i] * (z[i] / S[i]);
                double inc = qd / (double)d;
                double g = (success_iter == 0) ? inc : global_diffusion[0] + (inc - global_diffusion[0]) / (double)success_iter;
                for (int j = 0; j < d; j++) global_diffusion[j] = g;
            }
        }
        memcpy(mu, mu_filt, sizeof(double) * D);
        memcpy(R, R_filt, sizeof(double) * NB);
        double dtnew = dt;
        if (c->adaptive) {
            if (c->qsteady_min <= qq && qq <= c->qsteady_max) qq = 1.0;
            qold = fmax(EEst, c->qoldinit);
            dtnew = dt / qq;
        }
        double ttmp = t + dt;
        double ref = fmax(fabs(t), fabs(tstop));
        double epsref = nextafter(ref, INFINITY) - ref;
        t = (fabs(ttmp - tstop) < 100.0 * epsref) ? tstop : ttmp;
        dt = dtnew;
        success_iter++;
        out->naccept++;
        int nanu = 0;
        for (int i = 0; i < d; i++) {
            uprev[i] = so ? mu[d + i] : mu[i];
            if (isnan(mu[i])) nanu = 1;
        }
        /* DataUpdateCallback on the blocks (update.jl:197-239): one scalar observation per block */
        for (int64_t j = 0; j < c->n_data; j++)
            if (c->data_t[j] == t)
                for (int i = 0; i < d; i++)
                    out->data_loglik += data_update_factor(n, 1, d, c->data_u + j * c->n_obs + i, 1, AT(c->obs_Rn, i, i, d), mu + i, R + i * nn);
        if (c->save_everystep) {
            dvec_push(&ts, &t, 1);
            dvec_push(&diffs, local_diffusion, d);
            dvec_push(&xmu, mu, D);
            dvec_push(&xR, R, NB);
            if (c->smooth) {
                dvec_push(&bG, G, NB);
                dvec_push(&bB, bb, D);
                dvec_push(&bL, LR, 2 * NB);
            }
        }
        if (nanu) { retcode = ORACLE_RET_UNSTABLE; break; }
    }
    if (!c->save_everystep) {
        dvec_push(&ts, &t, 1);
        dvec_push(&diffs, local_diffusion, d);
        dvec_push(&xmu, mu, D);
        dvec_push(&xR, R, NB);
    }
    int64_t ns = (int64_t)ts.len;
    out->n = ns;
    out->retcode = retcode;
    out->loglik = loglik_total;
    out->global_diffusion = global_diffusion[0];
    out->global_diffusion_mv = (double*)malloc(sizeof(double) * d);
    memcpy(out->global_diffusion_mv, global_diffusion, sizeof(double) * d);
    {
        double pad[64];
        for (int i = 0; i < d; i++) pad[i] = NAN;
        while ((int64_t)diffs.len < ns * d) dvec_push(&diffs, pad, d);
    }
    if (!dynamic) {
        if (c->calibrate && !isnan(global_diffusion[0])) { /* calibrate_solution!, apply_diffusion! on blocks */
            for (int64_t k = 0; k < ns; k++)
                for (int i = 0; i < d; i++) {
                    diffs.p[k * d + i] = global_diffusion[i] * default_diffusion[i];
                    double sm = sqrt(global_diffusion[i]);
                    double* r = xR.p + (size_t)k * NB + i * nn;
                    for (size_t e = 0; e < nn; e++) r[e] *= sm;
                    if (c->smooth && k + 1 < ns) {
                        double* l = bL.p + (size_t)k * 2 * NB + i * 2 * nn;
                        for (size_t e = 0; e < 2 * nn; e++) l[e] *= sm;
                    }
                }
        } else {
            for (int64_t k = 0; k < ns; k++)
                for (int i = 0; i < d; i++) diffs.p[k * d + i] = default_diffusion[i];
        }
    }
    out->t = ts.p;
    out->diffusions_mv = diffs.p;
    out->diffusions = (double*)malloc(sizeof(double) * (size_t)ns);
    for (int64_t k = 0; k < ns; k++) out->diffusions[k] = diffs.p[k * d];
    out->x_filt_mu = xmu.p;
    out->x_filt_R = (double*)calloc((size_t)ns * D * D, sizeof(double));
    out->pu_mu = (double*)malloc(sizeof(double) * (size_t)ns * du);
    out->pu_R = (double*)malloc(sizeof(double) * (size_t)ns * D * du);
    /* dense form of a Blocks matrix: block i occupies rows/columns i:d:D (blocksofdiagonals.jl:44-52) */
#define PN_BLOCKS_TO_DENSE(dst, src)                                                        \
    for (int i = 0; i < d; i++)                                                             \
        for (int cc = 0; cc < n; cc++)                                                      \
            for (int r = 0; r < n; r++) AT(dst, r * d + i, cc * d + i, D) = AT((src) + i * nn, r, cc, n);
    for (int64_t k = 0; k < ns; k++) {
        double* dst = out->x_filt_R + (size_t)k * D * D;
        PN_BLOCKS_TO_DENSE(dst, xR.p + (size_t)k * NB)
    }
    double* smu = NULL;
    double* sR = NULL;
    if (c->smooth && c->save_everystep) { /* smooth_solution! ; marginalize! Blocks (markov_kernel.jl:123-143) */
        smu = (double*)malloc(sizeof(double) * (size_t)ns * D);
        sR = (double*)malloc(sizeof(double) * (size_t)ns * NB);
        memcpy(smu, xmu.p, sizeof(double) * (size_t)ns * D);
        memcpy(sR, xR.p, sizeof(double) * (size_t)ns * NB);
        for (int64_t k = ns - 2; k >= 0; k--) {
            double dti = ts.p[k + 1] - ts.p[k];
            if (dti == 0.0) {
                memcpy(smu + (size_t)k * D, smu + (size_t)(k + 1) * D, sizeof(double) * D);
                memcpy(sR + (size_t)k * NB, sR + (size_t)(k + 1) * NB, sizeof(double) * NB);
                continue;
            }
            for (int i = 0; i < d; i++)
                marginalize_generic(n, 1, d, 0, smu + (size_t)(k + 1) * D + i, sR + (size_t)(k + 1) * NB + i * nn,
                                    bG.p + (size_t)k * NB + i * nn, bB.p + (size_t)k * D + i, bL.p + (size_t)k * 2 * NB + i * 2 * nn,
                                    smu + (size_t)k * D + i, sR + (size_t)k * NB + i * nn);
        }
        out->x_smooth_mu = smu;
        out->x_smooth_R = (double*)calloc((size_t)ns * D * D, sizeof(double));
        for (int64_t k = 0; k < ns; k++) {
            double* dst = out->x_smooth_R + (size_t)k * D * D;
            PN_BLOCKS_TO_DENSE(dst, sR + (size_t)k * NB)
        }
        /* dense-form backward kernels for inspection (block i <-> indices i:d:, also in the 2D x D factor of Lambda) */
        int64_t nk = ns - 1;
        out->bk_G = (double*)calloc((size_t)(nk > 0 ? nk : 1) * D * D, sizeof(double));
        out->bk_b = (double*)malloc(sizeof(double) * (size_t)(nk > 0 ? nk : 1) * D);
        out->bk_LR = (double*)calloc((size_t)(nk > 0 ? nk : 1) * 2 * D * D, sizeof(double));
        for (int64_t k = 0; k < nk; k++) {
            double* Gd = out->bk_G + (size_t)k * D * D;
            double* Ld = out->bk_LR + (size_t)k * 2 * D * D;
            PN_BLOCKS_TO_DENSE(Gd, bG.p + (size_t)k * NB)
            for (int i = 0; i < d; i++)
                for (int cc = 0; cc < n; cc++)
                    for (int r = 0; r < 2 * n; r++)
                        AT(Ld, r * d + i, cc * d + i, 2 * D) = AT(bL.p + (size_t)k * 2 * NB + (size_t)i * 2 * nn, r, cc, 2 * n);
            memcpy(out->bk_b + (size_t)k * D, bB.p + (size_t)k * D, sizeof(double) * D);
        }
    }
#undef PN_BLOCKS_TO_DENSE
    for (int64_t k = 0; k < ns; k++) {
        const double* xm = (smu ? smu : xmu.p) + (size_t)k * D;
        const double* xr = (smu ? out->x_smooth_R : out->x_filt_R) + (size_t)k * D * D;
        for (int j = 0; j < du; j++) { /* SolProj row j selects state entry sc: E0 rows, or (E1; E0) for second order */
            const int sc = so ? (j < d ? d + j : j - d) : j;
            out->pu_mu[(size_t)k * du + j] = xm[sc];
            for (int r = 0; r < D; r++) out->pu_R[(size_t)k * D * du + (size_t)j * D + r] = AT(xr, r, sc, D);
        }
    }
    free(sR);
    free(xR.p); free(bG.p); free(bB.p); free(bL.p);
    free(Ah1); free(Qh1); free(mu); free(R); free(mu_pred); free(R_pred); free(mu_filt); free(R_filt);
    free(H); free(J); free(G); free(bb); free(LR); free(Qs);
    return 0;
}

/* The same data update on ONE n x n factor with the mean handled as an n x nc matrix (column c at mu[r*rs + c]):
   the Kronecker layout (nc = d: update.jl:151-195, with the log-density of the 1 x 1 factor -- logdet is NOT multiplied by
   d, the quirk of update.jl:147) and one block of the BlocksOfDiagonals layout (nc = 1: update.jl:197-239).  Observation
   row e_0' (full observations of the solution), noise standard deviation sn.  Returns the log-likelihood. */
static double data_update_factor(int n, int nc, int rs, const double* u, int us, double sn, double* mu, double* R) {
    double f[32], K[32], z[64];
    double* St = (double*)malloc(sizeof(double) * (n + 1) * n);
    double s2 = sn * sn, quad = 0.0, ll;
    for (int r = 0; r < n; r++) { f[r] = AT(R, r, 0, n); s2 += f[r] * f[r]; } /* R H' with H = e_0' */
    for (int c = 0; c < nc; c++) z[c] = mu[c] - u[c * us];
    if (all_zero(R, (size_t)n * n)) { /* update.jl:82-89 */
        ll = all_zero(z, (size_t)nc) ? INFINITY : -INFINITY;
        free(St);
        return ll;
    }
    for (int k = 0; k < n; k++) {
        double sacc = 0.0;
        for (int r = 0; r < n; r++) sacc += AT(R, r, k, n) * f[r];
        K[k] = sacc / s2;
    }
    for (int c = 0; c < nc; c++) quad += z[c] * z[c] / s2;
    ll = -0.5 * quad - 0.5 * (double)nc * log(2.0 * M_PI) - 0.5 * log(s2);
    for (int r = 0; r < n; r++)
        for (int c = 0; c < nc; c++) mu[r * rs + c] -= K[r] * z[c];
    for (int k = 0; k < n; k++) {
        for (int r = 0; r < n; r++) AT(St, r, k, n + 1) = AT(R, r, k, n) - f[r] * K[k]; /* R (I - K H)' */
        AT(St, n, k, n + 1) = sn * K[k];                                                 /* (K Rn')'    */
    }
    oracle_triangularize(n + 1, n, St, R);
    free(St);
    return ll;
}

/* observation model check for the structured layouts: observation_matrix = I (obs_H = E0) and a diagonal noise factor */
static int data_model_is_componentwise(const oracle_config* c, int D) {
    if (c->n_obs != c->d) return 0;
    for (int a = 0; a < c->d; a++) {
        for (int k = 0; k < D; k++)
            if (AT(c->obs_H, a, k, c->d) != ((k == a) ? 1.0 : 0.0)) return 0;
        for (int b = 0; b < c->d; b++)
            if (a != b && AT(c->obs_Rn, a, b, c->d) != 0.0) return 0;
    }
    return 1;
}

/* DataUpdateCallback affect! (src/callbacks/dataupdate.jl:48-82): condition the filter state (mu, R) on the observation
   u = H x + eps, eps ~ N(0, Rn'Rn): obs_cov factor make_obscov_sqrt = qr([R H'; Rn]).R (:86-87), then
   update!(...; R) (src/filtering/update.jl:65-139) with the QR route of :132-136.  Returns the log-likelihood. */
static double data_update(int D, int o, const double* H /* o x D */, const double* Rn /* o x o */, const double* u, double* mu,
                          double* R) {
    double* z = (double*)malloc(sizeof(double) * o);
    double* zt = (double*)malloc(sizeof(double) * o);
    double* Cm = (double*)malloc(sizeof(double) * D * o);
    double* St = (double*)malloc(sizeof(double) * (D + o) * o);
    double* Sr = (double*)malloc(sizeof(double) * o * o);
    double* K = (double*)malloc(sizeof(double) * D * o);
    double* Mc = (double*)malloc(sizeof(double) * D * D);
    double* R2 = (double*)malloc(sizeof(double) * (D + o) * D);
    double* K1 = (double*)malloc(sizeof(double) * D * o);
    double ll;
    for (int a = 0; a < o; a++) {
        double sacc = 0.0;
        for (int c = 0; c < D; c++) sacc += AT(H, a, c, o) * mu[c];
        z[a] = sacc - u[a];
        zt[a] = z[a];
    }
    if (all_zero(R, (size_t)D * D)) { /* update.jl:82-89 */
        ll = all_zero(z, (size_t)o) ? INFINITY : -INFINITY;
    } else {
        gemm(D, o, D, 1.0, R, D, 0, H, o, 1, 0.0, Cm, D);           /* R H' */
        for (int a = 0; a < o; a++) {
            for (int r = 0; r < D; r++) AT(St, r, a, D + o) = AT(Cm, r, a, D);
            for (int b = 0; b < o; b++) AT(St, D + b, a, D + o) = AT(Rn, b, a, o);
        }
        oracle_triangularize(D + o, o, St, Sr);                     /* S = Sr'Sr */
        gemm(D, o, D, 1.0, R, D, 1, Cm, D, 0, 0.0, K, D);           /* Sigma H' */
        rdiv_chol_upper(D, o, K, Sr);                               /* K = Sigma H' S^-1 */
        rdiv_chol_upper(1, o, zt, Sr);                              /* z' S^-1 */
        double quad = 0.0, logdet = 0.0;
        for (int a = 0; a < o; a++) {
            quad += z[a] * zt[a];
            logdet += 2.0 * log(fabs(AT(Sr, a, a, o)));
        }
        ll = -0.5 * quad - 0.5 * (double)o * log(2.0 * M_PI) - 0.5 * logdet;
        for (int c = 0; c < D; c++)
            for (int a = 0; a < o; a++) mu[c] -= AT(K, c, a, D) * z[a];
        gemm(D, D, o, -1.0, K, D, 0, H, o, 0, 0.0, Mc, D);          /* I - K H */
        for (int i = 0; i < D; i++) AT(Mc, i, i, D) += 1.0;
        gemm(D, o, o, 1.0, K, D, 0, Rn, o, 1, 0.0, K1, D);          /* K Rn' */
        for (int c = 0; c < D; c++) {
            for (int r = 0; r < D; r++) {
                double sacc = 0.0;
                for (int l = 0; l < D; l++) sacc += AT(R, r, l, D) * AT(Mc, c, l, D); /* R (I - K H)' */
                AT(R2, r, c, D + o) = sacc;
            }
            for (int a = 0; a < o; a++) AT(R2, D + a, c, D + o) = AT(K1, c, a, D);
        }
        oracle_triangularize(D + o, D, R2, R);
    }
    free(z); free(zt); free(Cm); free(St); free(Sr); free(K); free(Mc); free(R2); free(K1);
    return ll;
}

int oracle_solve(const oracle_config* c, oracle_rhs_fn f, oracle_jac_fn jac, oracle_taylor_fn taylor,
                 const double* u0, const double* p, oracle_traj* out) {
    /* mass matrices: EK1 on the dense layout takes any M; the Kronecker layout (EK0 + IWP + scalar diffusion) only a
       UniformScaling (ArgumentError otherwise, caches.jl:111-118; H = lambda E1, derivative_utils.jl:45-47); the
       BlocksOfDiagonals layout is restated for diagonal M */
    if (c->second_order && (c->mass_matrix || c->q < 2)) return -2;
    if (c->n_data > 0 && (c->second_order || c->mass_matrix) && c->alg != ORACLE_EK1) return -2;
    if (c->n_data > 0 && (c->alg == ORACLE_DIAGONAL_EK1 || c->diffusion == ORACLE_DIFF_DYNAMIC_MV || c->diffusion == ORACLE_DIFF_FIXED_MV ||
                          (c->alg == ORACLE_EK0 && c->prior == ORACLE_PRIOR_IWP))) {
        /* structured layouts: full observations with component-wise noise only ("Partial observations only work with the
           EK1 right now", dataupdate.jl:69-71); the Kronecker layout additionally needs isotropic noise */
        if (!data_model_is_componentwise(c, c->d * (c->q + 1))) return -2;
        if (c->alg == ORACLE_EK0 && c->diffusion != ORACLE_DIFF_DYNAMIC_MV && c->diffusion != ORACLE_DIFF_FIXED_MV)
            for (int a = 1; a < c->d; a++)
                if (AT(c->obs_Rn, a, a, c->d) != c->obs_Rn[0]) return -2;
    }
    if (c->mass_matrix) {
        const double* M = c->mass_matrix;
        int diagonal = 1, uniform = 1, zero_diag = 0;
        for (int a = 0; a < c->d; a++)
            for (int i = 0; i < c->d; i++) {
                if (a != i && AT(M, a, i, c->d) != 0.0) diagonal = 0;
                if (a == i && AT(M, a, i, c->d) != M[0]) uniform = 0;
                if (a == i && AT(M, a, i, c->d) == 0.0) zero_diag = 1;
            }
        const int blocks_layout = (c->alg == ORACLE_DIAGONAL_EK1 || c->diffusion == ORACLE_DIFF_DYNAMIC_MV || c->diffusion == ORACLE_DIFF_FIXED_MV);
        if (blocks_layout && (!diagonal || (c->alg == ORACLE_EK0 && zero_diag))) return -2;
        if (!blocks_layout && c->alg == ORACLE_EK0 && c->prior == ORACLE_PRIOR_IWP && !(diagonal && uniform && !zero_diag)) return -2;
        if (!blocks_layout && c->alg == ORACLE_EK0 && c->prior != ORACLE_PRIOR_IWP) return -2;
    }
    if (c->alg == ORACLE_DIAGONAL_EK1 || c->diffusion == ORACLE_DIFF_DYNAMIC_MV || c->diffusion == ORACLE_DIFF_FIXED_MV)
        return solve_blocks(c, f, jac, taylor, u0, p, out);
    const int d = c->d, q = c->q, n = q + 1, D = d * n;
    /* covariance_structure (src/algorithms.jl:132-151): EK0 is Kronecker only with the IWP prior, else dense */
    const int kron = (c->alg == ORACLE_EK0 && c->prior == ORACLE_PRIOR_IWP);
    const int N = kron ? n : D;          /* size of the matrices the filter works on */
    const int nrhs = kron ? d : 1;       /* mean handled as N x nrhs (update.jl:166-174) */
    const int rs = kron ? d : 1, cs = kron ? 1 : 0;
    const int m = kron ? 1 : d;          /* rows of H on this layout */
    memset(out, 0, sizeof(*out));
    out->d = d; out->q = q; out->D = D; out->smooth = c->smooth; out->alg = c->alg;
    out->global_diffusion = NAN;
    const int so = c->second_order ? 1 : 0;
    const int du = so ? 2 * d : d;       /* length of sol.u: (du, u) for second-order problems */
    const int e1 = so ? 2 : 1;           /* the measured derivative: E2 x for second-order problems, E1 x otherwise */
    if (du > 64 || n > 32) return -1;

    /* --- work buffers (EKCache, src/caches.jl:198-214) --- */
    size_t NN = (size_t)N * N;
    double* Ah1 = (double*)malloc(sizeof(double) * n * n);
    double* Qh1 = (double*)malloc(sizeof(double) * n * n);
    double* Ah = (double*)malloc(sizeof(double) * NN);
    double* QhR = (double*)malloc(sizeof(double) * NN);
    double* mu = (double*)calloc(D, sizeof(double));
    double* R = (double*)calloc(NN, sizeof(double));
    double* mu_pred = (double*)calloc(D, sizeof(double));
    double* R_pred = (double*)calloc(NN, sizeof(double));
    double* mu_filt = (double*)calloc(D, sizeof(double));
    double* R_filt = (double*)calloc(NN, sizeof(double));
    double* H = (double*)calloc((size_t)m * N, sizeof(double));
    double* J = (double*)calloc((size_t)du * du, sizeof(double));
    double* Cq = (double*)malloc(sizeof(double) * (size_t)N * m);
    double* W = (double*)malloc(sizeof(double) * (size_t)m * m);
    double* G = (double*)malloc(sizeof(double) * NN);
    double* bb = (double*)malloc(sizeof(double) * D);
    double* LR = (double*)malloc(sizeof(double) * 2 * NN);
    double* Rdense = (double*)malloc(sizeof(double) * (size_t)D * D);
    double* tmpDD = (double*)malloc(sizeof(double) * (size_t)2 * D * D);
    double z[64], zt[64], fu[64], uprev[64], upred[64], err[64];
    double* rate_now = (double*)calloc((size_t)d * d + 1, sizeof(double));

    dvec ts = {0}, diffs = {0}, xmu = {0}, xR = {0}, bG = {0}, bB = {0}, bL = {0};

    /* --- initial state: exact Taylor-mode derivatives, Sigma0.R == 0 exactly
       (src/initialization/autodiffinit.jl:1-47 + common.jl:122-139; SURVEY H7;
        test/solution.jl:59-62 asserts iszero(pu[1].Sigma.R)) --- */
    if (so) { /* derivatives of the first-order form; the state takes the u-part (autodiffinit.jl:34-37: df.x[2]) */
        double* t2 = (double*)malloc(sizeof(double) * (size_t)du * n);
        taylor(t2, u0, p, c->t0, q);
        for (int o = 0; o <= q; o++)
            for (int i = 0; i < d; i++) mu[o * d + i] = t2[o * du + d + i];
        free(t2);
    } else {
        taylor(mu, u0, p, c->t0, q);
    }
    out->nf += q; /* autodiffinit.jl:17 */
    /* error scaling uses integ.u / integ.uprev, for second-order problems their first partition du = E1 x
       (perform_step.jl:204-216) */
    for (int i = 0; i < d; i++) uprev[i] = so ? mu[d + i] : mu[i];
    const double* Mm = c->mass_matrix;
    double* MMinv = NULL;
    int m_singular = 0;
    if (Mm) {
        MMinv = (double*)malloc(sizeof(double) * d * d);
        mass_setup(d, Mm, MMinv, &m_singular);
        (void)m_singular;
        for (int o = 1; o <= q; o++) { /* x_o = MM^-1 df_o */
            for (int a = 0; a < d; a++) {
                double sacc = 0.0;
                for (int i = 0; i < d; i++) sacc += AT(MMinv, a, i, d) * mu[o * d + i];
                z[a] = sacc;
            }
            for (int a = 0; a < d; a++) mu[o * d + a] = z[a];
        }
    }

    double t = c->t0;
    double dt;
    const int order = q; /* alg_order (src/alg_utils.jl:45) */
    if (c->adaptive && !(c->dt > 0.0)) {
        dt = initial_dt(c, f, u0, p, order, Mm != NULL);
        if (!Mm) out->nf += 2;
    } else {
        dt = c->dt;
    }
    const double dtcache = dt;
    double qold = c->qoldinit;
    int64_t iter = 0, success_iter = 0;
    double local_diffusion = NAN;                 /* caches.jl:257 (initdiff * NaN) */
    double global_diffusion = NAN;
    const double default_diffusion = c->initial_diffusion;
    const int dynamic = (c->diffusion == ORACLE_DIFF_DYNAMIC);
    int retcode = ORACLE_RET_SUCCESS;
    int64_t tstop_idx = 0;
    double loglik_total = 0.0;

    /* save initial (perform_step.jl:30-32) */
    dvec_push(&ts, &t, 1);
    dvec_push(&xmu, mu, D);
    dvec_push(&xR, R, NN);

    while (t < c->t1) {
        double tstop = c->t1;
        while (tstop_idx < c->n_tstops && c->tstops[tstop_idx] <= t) tstop_idx++;
        if (tstop_idx < c->n_tstops && c->tstops[tstop_idx] < c->t1) tstop = c->tstops[tstop_idx];

        /* loopheader! (OrdinaryDiffEqCore) */
        iter++;
        if (iter > c->maxiters) { retcode = ORACLE_RET_MAXITERS; break; }
        if (c->adaptive) {
            dt = fmin(dt, c->dtmax);
            dt = fmax(dt, c->dtmin);
            dt = fmin(dt, tstop - t);
        } else {
            dt = fmin(dtcache, tstop - t);
        }
        if (isnan(dt) || !(t + dt > t) || (c->adaptive && dt <= c->dtmin)) {
            retcode = ORACLE_RET_DTLESSTHANMIN;
            break;
        }

        /* ---- perform_step! (src/perform_step.jl:79-154) ---- */
        double tnew = t + dt;                                  /* :85 */
        const int gen_rate = (c->prior == ORACLE_PRIOR_IOUP) && (c->rate_matrix || c->update_rate_parameter);
        if (gen_rate) {
            if (c->update_rate_parameter) { /* :88-96: calc_J!(rate, integ, cache, false) -- the Jacobian at (uprev, t) */
                jac(J, uprev, p, t);
                out->njac += 1;
                memcpy(rate_now, J, sizeof(double) * d * d);
            } else {
                memcpy(rate_now, c->rate_matrix, sizeof(double) * d * d);
            }
            if (oracle_ioup_transition_matrix(d, q, rate_now, dt, Ah, QhR) != 0) { retcode = ORACLE_RET_UNSTABLE; break; }
        } else if (c->prior == ORACLE_PRIOR_IOUP) oracle_ioup_transition(q, c->rate_parameter, dt, Ah1, Qh1); /* ioup.jl:115-131 */
        else oracle_iwp_transition(q, dt, Ah1, Qh1);           /* :87-99, make_transition_matrices! */
        if (gen_rate) {
            /* dense pair already in place */
        } else if (kron) {
            memcpy(Ah, Ah1, sizeof(double) * NN);
            memcpy(QhR, Qh1, sizeof(double) * NN);
        } else {
            oracle_kron_identity(d, n, n, Ah1, Ah);            /* iwp.jl:165-171 (Matrix(A)) */
            oracle_kron_identity(d, n, n, Qh1, QhR);
        }
        /* predict_mean! (:102; predict.jl:53-60) */
        for (int r = 0; r < N; r++)
            for (int cc = 0; cc < nrhs; cc++) {
                double s = 0.0;
                for (int k = 0; k < N; k++) s += AT(Ah, r, k, N) * mu[k * rs + cc * cs];
                mu_pred[r * rs + cc * cs] = s;
            }
        for (int i = 0; i < d; i++) upred[i] = mu_pred[i];     /* integ.u = E0 mu_pred (:103-104) */
        if (so) /* integ.u = SolProj mu_pred = (E1 mu_pred, E0 mu_pred) */
            for (int i = 0; i < d; i++) { upred[i] = mu_pred[d + i]; upred[d + i] = mu_pred[i]; }
        /* evaluate_ode! (:107, :167-180): z = E1 mu_pred - f(E0 mu_pred) (measurement_models.jl:11-20) */
        f(fu, upred, p, tnew);
        out->nf += 1;
        for (int i = 0; i < d; i++) z[i] = mu_pred[e1 * d + i] - fu[i]; /* second order: E2 x - f1(E1 x, E0 x) (:38-48) */
        if (Mm) /* z = M E1 mu_pred - f (measurement_models.jl:18); on the Kronecker layout M = lambda I */
            for (int a = 0; a < d; a++) {
                double sacc = 0.0;
                for (int i = 0; i < d; i++) sacc += AT(Mm, a, i, d) * mu_pred[d + i];
                z[a] = sacc - fu[a];
            }
        /* calc_H! (derivative_utils.jl:1-33) */
        if (kron) {
            for (int k = 0; k < N; k++) H[k] = (k == e1) ? (Mm ? Mm[0] : 1.0) : 0.0; /* E1.B = e_1' (projection.jl:20-29), times lambda; second order: E2 (derivative_utils.jl:39-41) */
        } else {
            memset(H, 0, sizeof(double) * (size_t)m * N);
            for (int i = 0; i < d; i++) AT(H, i, e1 * d + i, d) = 1.0;    /* H = E1 (second order: E2) */
            if (Mm) /* H = M E1 (derivative_utils.jl:43-50) */
                for (int a = 0; a < d; a++)
                    for (int i = 0; i < d; i++) AT(H, a, d + i, d) = AT(Mm, a, i, d);
            if (c->alg == ORACLE_EK1) {
                jac(J, upred, p, tnew);
                out->njac += 1;
                if (so) { /* H -= ddu[1:d, :] * SolProj, SolProj = [E1; E0] (:12-13): J is the 2d x 2d Jacobian of the first-order form */
                    for (int a = 0; a < d; a++)
                        for (int i = 0; i < d; i++) {
                            AT(H, a, d + i, d) -= AT(J, a, i, du);     /* d f1_a / d du_i -> E1 columns */
                            AT(H, a, i, d) -= AT(J, a, d + i, du);     /* d f1_a / d u_i  -> E0 columns */
                        }
                } else
                for (int a = 0; a < d; a++)
                    for (int i = 0; i < d; i++) AT(H, a, i, d) -= AT(J, a, i, d); /* H -= J * E0 (:13) */
            }
        }
        /* local diffusion (:110-112; calibration.jl:123-133, invquad :9-29) */
        int have_local = 0;
        if (c->adaptive || dynamic) {
            gemm(N, m, N, 1.0, QhR, N, 0, H, m, 1, 0.0, Cq, N); /* C_Dxd = Qh.R H' */
            gemm(m, m, N, 1.0, Cq, N, 1, Cq, N, 0, 0.0, W, m);  /* HQH = C'C      */
            double sig;
            if (kron) {
                double zz = 0.0;
                for (int i = 0; i < d; i++) zz += z[i] * z[i];
                sig = zz / W[0] / (double)d; /* calibration.jl:16-20 */
            } else {
                if (oracle_cholesky_upper(m, W) != 0) { retcode = ORACLE_RET_UNSTABLE; break; }
                for (int a = 0; a < m; a++) zt[a] = z[a];
                for (int a = 0; a < m; a++) {
                    for (int k = 0; k < a; k++) zt[a] -= AT(W, k, a, m) * zt[k];
                    zt[a] /= AT(W, a, a, m);
                }
                for (int a = m - 1; a >= 0; a--) {
                    for (int k = a + 1; k < m; k++) zt[a] -= AT(W, a, k, m) * zt[k];
                    zt[a] /= AT(W, a, a, m);
                }
                double qd = 0.0;
                for (int a = 0; a < m; a++) qd += z[a] * zt[a];
                sig = qd / (double)d;
            }
            local_diffusion = sig;
            have_local = 1;
        }
        (void)have_local;
        if (c->forced_diffusion && success_iter < c->n_forced) local_diffusion = c->forced_diffusion[success_iter];
        /* error estimate (:113-119; :199-247) */
        double EEst = 0.0;
        if (c->adaptive) {
            double s = sqrt(local_diffusion);
            /* _Q.R = Qh.R * sqrt(sigma^2) ; C = _Q.R H' ; e_i = sqrt(sum_k C[k,i]^2) (:240-247) */
            double* Qs = tmpDD;
            for (size_t i = 0; i < NN; i++) Qs[i] = QhR[i] * s;
            gemm(N, m, N, 1.0, Qs, N, 0, H, m, 1, 0.0, Cq, N);
            if (kron) {
                double e = 0.0;
                for (int k = 0; k < N; k++) e += Cq[k] * Cq[k]; /* diag!(..IsometricKronecker) :251-252 */
                e = sqrt(e);
                for (int i = 0; i < d; i++) err[i] = e;
            } else {
                for (int i = 0; i < d; i++) {
                    double e = 0.0;
                    for (int k = 0; k < N; k++) e += AT(Cq, k, i, N) * AT(Cq, k, i, N);
                    err[i] = sqrt(e);
                }
            }
            for (int i = 0; i < d; i++) { /* calculate_residuals! (3rd party) */
                err[i] *= dt;
                err[i] = err[i] / (c->abstol + fmax(fabs(upred[i]), fabs(uprev[i])) * c->reltol);
            }
            EEst = rms_norm(err, d);
            if (isnan(EEst)) { retcode = ORACLE_RET_UNSTABLE; break; }
        }
        int accept = (!c->adaptive) || (EEst < 1.0); /* early return at EEst >= 1 (:116-118) */
        double q11 = 0.0, qq = 1.0;
        if (c->adaptive) { /* stepsize_controller!(::PIController) (3rd party; exact pow) */
            if (EEst == 0.0) {
                qq = 1.0 / c->qmax;
            } else {
                q11 = pow(EEst, c->beta1);
                qq = q11 / pow(qold, c->beta2);
                qq = fmax(1.0 / c->qmax, fmin(1.0 / c->qmin, qq / c->gamma));
            }
        }
        if (!accept) { /* step_reject_controller! */
            out->nreject++;
            dt = dt / fmin(1.0 / c->qmin, q11 / c->gamma);
            continue;
        }
        /* ---- accepted: covariance predict / backward kernel / update ---- */
        double ediff = dynamic ? local_diffusion : default_diffusion;               /* :121-122 */
        if (c->forced_diffusion && success_iter < c->n_forced) ediff = c->forced_diffusion[success_iter];
        oracle_predict_cov(N, R, Ah, QhR, ediff, c->cov_backend, R_pred, NULL);     /* :122-124 */
        if (c->smooth)                                                               /* :126-131 */
            backward_kernel_generic(N, nrhs, rs, cs, mu, R, mu_pred, R_pred, Ah, QhR, ediff, G, bb, LR);
        double ll;
        /* Kronecker layout: isotropic observation noise, the factor of the 1 x 1 left Kronecker factor is R.R[0][0] */
        int ufail = update_generic(N, m, nrhs, mu_pred, rs, cs, R_pred, z, 1, 1, H, mu_filt, R_filt, &ll, c->pn_obs_Rn); /* :134-138 */
        if (ufail) { retcode = ORACLE_RET_UNSTABLE; break; }
        loglik_total += ll;                                                          /* :142-143 */
        if (!dynamic) { /* estimate_global_diffusion(::FixedDiffusion) (calibration.jl:49-66) */
            /* invquad(z, S) / d with S = measurement covariance from the predicted state */
            gemm(N, m, N, 1.0, R_pred, N, 0, H, m, 1, 0.0, Cq, N);
            gemm(m, m, N, 1.0, Cq, N, 1, Cq, N, 0, 0.0, W, m);
            if (c->pn_obs_Rn) gemm(m, m, m, 1.0, c->pn_obs_Rn, m, 1, c->pn_obs_Rn, m, 0, 1.0, W, m); /* measurement.Sigma includes R */
            double inc;
            if (kron) {
                double zz = 0.0;
                for (int i = 0; i < d; i++) zz += z[i] * z[i];
                inc = zz / W[0] / (double)d;
            } else {
                if (oracle_cholesky_upper(m, W) != 0) { retcode = ORACLE_RET_UNSTABLE; break; }
                for (int a = 0; a < m; a++) zt[a] = z[a];
                for (int a = 0; a < m; a++) {
                    for (int k = 0; k < a; k++) zt[a] -= AT(W, k, a, m) * zt[k];
                    zt[a] /= AT(W, a, a, m);
                }
                for (int a = m - 1; a >= 0; a--) {
                    for (int k = a + 1; k < m; k++) zt[a] -= AT(W, a, k, m) * zt[k];
                    zt[a] /= AT(W, a, a, m);
                }
                double qd = 0.0;
                for (int a = 0; a < m; a++) qd += z[a] * zt[a];
                inc = qd / (double)d;
            }
            /* divisor is success_iter BEFORE this step is counted (:56-62) */
            global_diffusion = (success_iter == 0) ? inc : global_diffusion + (inc - global_diffusion) / (double)success_iter;
        }
        memcpy(mu, mu_filt, sizeof(double) * D); /* x <- x_filt (:151) */
        memcpy(R, R_filt, sizeof(double) * NN);

        /* loopfooter! accept branch (3rd party) */
        double dtnew = dt;
        if (c->adaptive) {
            if (c->qsteady_min <= qq && qq <= c->qsteady_max) qq = 1.0;
            qold = fmax(EEst, c->qoldinit);
            dtnew = dt / qq;
        }
        double ttmp = t + dt;
        double ref = fmax(fabs(t), fabs(tstop));
        double epsref = nextafter(ref, INFINITY) - ref;
        t = (fabs(ttmp - tstop) < 100.0 * epsref) ? tstop : ttmp;
        dt = dtnew;
        success_iter++;
        out->naccept++;
        int nanu = 0;
        for (int i = 0; i < d; i++) {
            uprev[i] = so ? mu[d + i] : mu[i]; /* update_uprev! (integrator_utils.jl:173-183) */
            if (isnan(mu[i])) nanu = 1;
        }
        /* PresetTimeCallback at the data times (dataupdate.jl:83): condition the state on the observation */
        for (int64_t j = 0; j < c->n_data; j++)
            if (c->data_t[j] == t) {
                if (kron) out->data_loglik += data_update_factor(n, d, d, c->data_u + j * c->n_obs, 1, c->obs_Rn[0], mu, R);
                else out->data_loglik += data_update(D, c->n_obs, c->obs_H, c->obs_Rn, c->data_u + j * c->n_obs, mu, R);
            }
        /* savevalues! (integrator_utils.jl:143-171) */
        if (c->save_everystep) {
            dvec_push(&ts, &t, 1);
            dvec_push(&diffs, &local_diffusion, 1);
            dvec_push(&xmu, mu, D);
            dvec_push(&xR, R, NN);
            if (c->smooth) {
                dvec_push(&bG, G, NN);
                dvec_push(&bB, bb, D);
                dvec_push(&bL, LR, 2 * NN);
            }
        }
        if (nanu) { retcode = ORACLE_RET_UNSTABLE; break; }
    }
    if (!c->save_everystep) { /* save_end (integrator_utils.jl:123-141) */
        dvec_push(&ts, &t, 1);
        dvec_push(&diffs, &local_diffusion, 1);
        dvec_push(&xmu, mu, D);
        dvec_push(&xR, R, NN);
    }

    /* ---- postamble! (integrator_utils.jl:9-36) ---- */
    int64_t ns = (int64_t)ts.len;
    out->n = ns;
    out->retcode = retcode;
    out->loglik = loglik_total;
    out->global_diffusion = global_diffusion;
    {
        double pad = NAN;
        while ((int64_t)diffs.len < ns) dvec_push(&diffs, &pad, 1);
    }
    if (!dynamic) {
        if (c->calibrate && !isnan(global_diffusion)) { /* calibrate_solution! (:46-65) */
            double sm = sqrt(global_diffusion);
            for (int64_t k = 0; k < ns; k++) diffs.p[k] = global_diffusion * default_diffusion;
            for (size_t i = 0; i < xR.len; i++) xR.p[i] *= sm;
            for (int64_t k = 0; k + 1 < ns && c->smooth; k++) {
                double* l = bL.p + (size_t)k * 2 * NN;
                for (size_t i = 0; i < 2 * NN; i++) l[i] *= sm;
            }
        } else {
            for (int64_t k = 0; k < ns; k++) diffs.p[k] = default_diffusion; /* set_diffusions! (:74-88) */
        }
    }

    /* ---- outputs in dense form ---- */
    out->t = ts.p;
    out->diffusions = diffs.p;
    out->x_filt_mu = xmu.p;
    out->x_filt_R = (double*)malloc(sizeof(double) * (size_t)ns * D * D);
    out->pu_mu = (double*)malloc(sizeof(double) * (size_t)ns * du);
    out->pu_R = (double*)malloc(sizeof(double) * (size_t)ns * D * du);
    for (int64_t k = 0; k < ns; k++) {
        double* dst = out->x_filt_R + (size_t)k * D * D;
        if (kron) oracle_kron_identity(d, n, n, xR.p + (size_t)k * NN, dst);
        else memcpy(dst, xR.p + (size_t)k * NN, sizeof(double) * NN);
    }
    double* smu = NULL;
    double* sR = NULL; /* smoothed, on the working layout */
    if (c->smooth && c->save_everystep) { /* smooth_solution! (:98-121) */
        smu = (double*)malloc(sizeof(double) * (size_t)ns * D);
        sR = (double*)malloc(sizeof(double) * (size_t)ns * NN);
        memcpy(smu, xmu.p, sizeof(double) * (size_t)ns * D);
        memcpy(sR, xR.p, sizeof(double) * (size_t)ns * NN);
        for (int64_t i = ns - 2; i >= 0; i--) {
            double dti = ts.p[i + 1] - ts.p[i];
            if (dti == 0.0) { /* :109-112 */
                memcpy(smu + (size_t)i * D, smu + (size_t)(i + 1) * D, sizeof(double) * D);
                memcpy(sR + (size_t)i * NN, sR + (size_t)(i + 1) * NN, sizeof(double) * NN);
                continue;
            }
            marginalize_generic(N, nrhs, rs, cs, smu + (size_t)(i + 1) * D, sR + (size_t)(i + 1) * NN,
                                bG.p + (size_t)i * NN, bB.p + (size_t)i * D, bL.p + (size_t)i * 2 * NN,
                                smu + (size_t)i * D, sR + (size_t)i * NN);
        }
        out->x_smooth_mu = smu;
        out->x_smooth_R = (double*)malloc(sizeof(double) * (size_t)ns * D * D);
        for (int64_t k = 0; k < ns; k++) {
            double* dst = out->x_smooth_R + (size_t)k * D * D;
            if (kron) oracle_kron_identity(d, n, n, sR + (size_t)k * NN, dst);
            else memcpy(dst, sR + (size_t)k * NN, sizeof(double) * NN);
        }
        /* dense-form backward kernels for inspection */
        int64_t nk = ns - 1;
        out->bk_G = (double*)malloc(sizeof(double) * (size_t)(nk > 0 ? nk : 1) * D * D);
        out->bk_b = (double*)malloc(sizeof(double) * (size_t)(nk > 0 ? nk : 1) * D);
        out->bk_LR = (double*)malloc(sizeof(double) * (size_t)(nk > 0 ? nk : 1) * 2 * D * D);
        for (int64_t k = 0; k < nk; k++) {
            if (kron) {
                oracle_kron_identity(d, n, n, bG.p + (size_t)k * NN, out->bk_G + (size_t)k * D * D);
                oracle_kron_identity(d, 2 * n, n, bL.p + (size_t)k * 2 * NN, out->bk_LR + (size_t)k * 2 * D * D);
            } else {
                memcpy(out->bk_G + (size_t)k * D * D, bG.p + (size_t)k * NN, sizeof(double) * NN);
                memcpy(out->bk_LR + (size_t)k * 2 * D * D, bL.p + (size_t)k * 2 * NN, sizeof(double) * 2 * NN);
            }
            memcpy(out->bk_b + (size_t)k * D, bB.p + (size_t)k * D, sizeof(double) * D);
        }
    }
    /* pu = SolProj * x  (_gaussian_mul!, src/gaussians.jl:135-139): mu_u = E0 mu, R_u = R E0' (D x d) */
    for (int64_t k = 0; k < ns; k++) {
        const double* xm = (smu ? smu : xmu.p) + (size_t)k * D;
        const double* xr = (smu ? out->x_smooth_R : out->x_filt_R) + (size_t)k * D * D;
        for (int j = 0; j < du; j++) { /* SolProj row j selects state entry sc: E0 rows, or (E1; E0) for second order */
            const int sc = so ? (j < d ? d + j : j - d) : j;
            out->pu_mu[(size_t)k * du + j] = xm[sc];
            for (int r = 0; r < D; r++) out->pu_R[(size_t)k * D * du + (size_t)j * D + r] = AT(xr, r, sc, D);
        }
    }
    free(sR);
    free(xR.p); free(bG.p); free(bB.p); free(bL.p);
    free(Ah1); free(Qh1); free(Ah); free(QhR); free(mu); free(R); free(mu_pred); free(R_pred);
    free(mu_filt); free(R_filt); free(H); free(J); free(Cq); free(W); free(G); free(bb); free(LR);
    free(Rdense); free(tmpDD); free(MMinv); free(rate_now);
    return 0;
}

void oracle_traj_free(oracle_traj* tr) {
    free(tr->t); free(tr->diffusions); free(tr->x_filt_mu); free(tr->x_filt_R);
    free(tr->x_smooth_mu); free(tr->x_smooth_R); free(tr->pu_mu); free(tr->pu_R);
    free(tr->bk_G); free(tr->bk_b); free(tr->bk_LR); free(tr->diffusions_mv); free(tr->global_diffusion_mv);
    memset(tr, 0, sizeof(*tr));
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* "EnsembleThreads" restatement: independent solves, one trajectory per OpenMP task. */
int oracle_solve_ensemble(const oracle_config* cfg, oracle_rhs_fn f, oracle_jac_fn jac, oracle_taylor_fn taylor,
                          int64_t n_traj, const double* u0, const double* p, int n_threads, double* u_final,
                          double* cov_final, int64_t* naccept, int64_t* nreject, int64_t* nf, int32_t* retcode,
                          double* loglik) {
    const int d = cfg->d, D = cfg->d * (cfg->q + 1);
    int err = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t i = 0; i < n_traj; i++) {
        oracle_traj tr;
        int rc = oracle_solve(cfg, f, jac, taylor, u0 + (size_t)i * d, p + (size_t)i * cfg->n_p, &tr);
        if (rc != 0) {
#pragma omp atomic write
            err = rc;
        } else {
            int64_t last = tr.n - 1;
            if (u_final)
                for (int k = 0; k < d; k++) u_final[(size_t)i * d + k] = tr.pu_mu[(size_t)last * d + k];
            if (cov_final) {
                const double* ru = tr.pu_R + (size_t)last * D * d; /* D x d */
                for (int a = 0; a < d; a++)
                    for (int b = 0; b < d; b++) {
                        double s = 0.0;
                        for (int r = 0; r < D; r++) s += ru[(size_t)a * D + r] * ru[(size_t)b * D + r];
                        cov_final[(size_t)i * d * d + (size_t)a * d + b] = s;
                    }
            }
            if (naccept) naccept[i] = tr.naccept;
            if (nreject) nreject[i] = tr.nreject;
            if (nf) nf[i] = tr.nf;
            if (retcode) retcode[i] = tr.retcode;
            if (loglik) loglik[i] = tr.loglik;
        }
        oracle_traj_free(&tr);
    }
    return err;
}
