/*
 * pnode_oracle.h -- CPU ORACLE for the ODE-filter hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (libpnode_cuda.so) never links, loads or calls anything in oracle/.
 *
 * It is a plain-C restatement (FP64) of ProbNumDiffEq.jl v0.18.0's filter/smoother
 * path, one function per reference routine, each citing the reference file:line it
 * follows.  Julia is not installable here, so the restatement is
 * pinned against the reference's own known-answer values and formula identities
 * (tests/test_oracle_*.py): IWP golden matrices test/core/priors.jl:49-192, the
 * dense-formula identities of test/core/filtering.jl, exact initialisation
 * test/state_init.jl:9-56 / test/solution.jl:59-62 and the end-to-end accuracy
 * thresholds of test/correctness.jl:12-87 and test/exponential_integrators.jl:6-41 (the
 * tightest pin the reference has on the adaptive path), plus the per-feature tests
 * listed in DESIGN.md section 5.  The adaptive step controller lives in
 * OrdinaryDiffEqCore (3rd party, not vendored): step-sequence parity against live
 * Julia is UNPINNED ("parity unpinned" items).
 *
 * All matrices are column-major (Julia layout).  State order is derivative-major:
 * x[j*d + i] = j-th derivative of component i.
 */
#ifndef PNODE_ORACLE_H
#define PNODE_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void (*oracle_rhs_fn)(double* du, const double* u, const double* p, double t);
/* J column-major d x d: J[a + d*i] = d f_a / d u_i */
typedef void (*oracle_jac_fn)(double* J, const double* u, const double* p, double t);
/* out[k*d + i] = d^k u_i / dt^k (t0), k = 0..q  (true derivatives, not Taylor coefficients) */
typedef void (*oracle_taylor_fn)(double* out, const double* u, const double* p, double t, int q);

enum { ORACLE_EK0 = 0, ORACLE_EK1 = 1, ORACLE_DIAGONAL_EK1 = 2 /* BlocksOfDiagonals layout, src/algorithms.jl:343-393 */ };
enum { ORACLE_DIFF_DYNAMIC = 0, ORACLE_DIFF_FIXED = 1, ORACLE_DIFF_DYNAMIC_MV = 2, ORACLE_DIFF_FIXED_MV = 3 /* src/diffusions/typedefs.jl */ };
enum { ORACLE_PRIOR_IWP = 0, ORACLE_PRIOR_IOUP = 1 /* scalar rate parameter, src/priors/ioup.jl */ };
enum { ORACLE_COV_REF = 0 /* Gram+Cholesky, QR fallback (predict.jl:84-91) */, ORACLE_COV_QR = 1 };
enum {
    ORACLE_RET_SUCCESS = 0,
    ORACLE_RET_MAXITERS = 1,
    ORACLE_RET_DTLESSTHANMIN = 2,
    ORACLE_RET_UNSTABLE = 3
};

typedef struct {
    int32_t d, q, n_p;
    int32_t alg;          /* ORACLE_EK0 (IsometricKronecker layout) / ORACLE_EK1 (dense) */
    int32_t diffusion;    /* ORACLE_DIFF_* */
    int32_t calibrate;    /* FixedDiffusion(calibrate=...) */
    int32_t smooth;
    int32_t adaptive;
    int32_t save_everystep;
    int32_t cov_backend;  /* ORACLE_COV_* */
    int64_t maxiters;
    double initial_diffusion;
    double t0, t1, dt;
    double abstol, reltol;
    double dtmin, dtmax;
    double beta1, beta2, qmin, qmax, gamma, qsteady_min, qsteady_max, qoldinit;
    /* optional extra stop times (sorted ascending, inside (t0,t1)); n_tstops may be 0 */
    const double* tstops;
    int64_t n_tstops;
    /* optional sigma^2 teacher forcing: value used in place of the local diffusion for
       accepted step k (k = 0..n_forced-1); NULL = free running */
    const double* forced_diffusion;
    int64_t n_forced;
    int32_t prior;          /* ORACLE_PRIOR_* */
    int32_t reserved;
    double rate_parameter;  /* IOUP: F[q,q] (src/priors/ioup.jl:66-79) */
    /* constant mass matrix M (d x d, column-major) of M u' = f(u,p,t); NULL = identity.  z = M E1 x - f
       (src/measurement_models.jl:11-20), H = M E1 - J E0 (src/derivative_utils.jl:1-53).  EK1 (dense) only. */
    const double* mass_matrix;
    /* SecondOrderODEProblem(f1, du0, u0, tspan, p): d = length(u) positions, the state models u and its q derivatives,
       z = E2 x - f1(E1 x, E0 x) (src/measurement_models.jl:29-49), H = E2 - J_du E1 - J_u E0 (EK1; EK0: E2;
       src/derivative_utils.jl:1-53), SolProj = [E1; E0] (src/caches.jl:126).  The rhs callbacks are those of the
       first-order form F([du; u]) = [f1(du, u); du] of dimension 2d (what the DynamicalODEFunction evaluates on the
       ArrayPartition); u0 holds (du0, u0); pu_* have 2d entries in SolProj order (du, u). */
    int32_t second_order;
    int32_t reserved1;
    /* DataUpdateCallback (src/callbacks/dataupdate.jl:39-84) as used by filtering_data_loglik
       (src/data_likelihoods/filtering.jl): whenever an accepted step ends on a data time the filter state is updated on
       the observation y = Hx + eps, eps ~ N(0, Rn'Rn), and the update's log-likelihood is accumulated in
       oracle_traj.data_loglik.  The caller puts the data times into tstops.  Dense layout only. */
    const double* data_t;   /* [n_data] */
    int64_t n_data;
    const double* data_u;   /* [n_data][n_obs] */
    int32_t n_obs;
    int32_t reserved2;
    const double* obs_H;    /* n_obs x D, column-major: observation_matrix * SolProj */
    const double* obs_Rn;   /* n_obs x n_obs, column-major upper factor of the observation-noise covariance */
    /* EK0/EK1(pn_observation_noise = R) (src/algorithms.jl:188-393, src/caches.jl:175-178): d x d column-major upper factor R.R of
       the noise covariance the ODE measurement z = 0 is observed with; NULL = noise-free.  S += R.R'R.R
       (src/perform_step.jl:182-188) and the filtered covariance gains K R K' (src/filtering/update.jl:125-136). */
    const double* pn_obs_Rn;
    /* IOUP(num_derivatives, rate_parameter::Union{AbstractVector,AbstractMatrix}) (src/priors/ioup.jl:81-113): d x d COLUMN-major
       rate matrix (a vector rate as Diagonal(r), as to_sde does); NULL = the scalar rate_parameter.  update_rate_parameter = 1:
       the rate is set to the Jacobian J(uprev, t) before every step (src/perform_step.jl:88-96, RosenbrockExpEK). */
    const double* rate_matrix;
    int32_t update_rate_parameter;
    int32_t reserved3;
} oracle_config;

/* One solved trajectory.  Arrays are owned by the struct; free with oracle_traj_free. */
typedef struct {
    int64_t n;            /* number of saved time points (naccept+1 if save_everystep, else 2) */
    int32_t d, q, D, smooth, alg;
    int32_t retcode;
    int64_t naccept, nreject, nf, njac;
    double loglik;
    double global_diffusion; /* FixedDiffusion MLE (calibration.jl:49-66); NaN otherwise */
    double* t;            /* [n] */
    double* diffusions;   /* [n]   (slot k belongs to the step k -> k+1; slot n-1 unused) */
    double* x_filt_mu;    /* [n][D] */
    double* x_filt_R;     /* [n][D*D] col-major right factor, Sigma = R'R (dense form, also for EK0) */
    double* x_smooth_mu;  /* [n][D]   (smooth only) */
    double* x_smooth_R;   /* [n][D*D] (smooth only) */
    double* pu_mu;        /* [n][d]   solution means (smoothed if smooth) */
    double* pu_R;         /* [n][D*d] col-major D x d right factor of the d x d solution covariance */
    double* bk_G;         /* [n-1][D*D] backward kernels (smooth only) */
    double* bk_b;         /* [n-1][D] */
    double* bk_LR;        /* [n-1][2D*D] col-major 2D x D */
    double* diffusions_mv;      /* [n][d] per-component diffusions (BlocksOfDiagonals solves only, else NULL) */
    double* global_diffusion_mv; /* [d] (BlocksOfDiagonals solves only, else NULL) */
    double data_loglik;   /* sum of the data updates' log-likelihoods (DataUpdateLogLikelihood); 0 without a data set */
} oracle_traj;

/* ---- component-level entry points (unit-tested against dense formulas) ---- */
void oracle_iwp_preconditioned(int q, double* A_breve /*n*n*/, double* L_breve /*n*n*/);
void oracle_preconditioner(int q, double h, double* P /*n*/, double* PI /*n*/);
void oracle_iwp_transition(int q, double h, double* Ah /*n*n*/, double* QhR /*n*n*/);
void oracle_expm(int n, const double* M, double* E);
void oracle_matrix_fraction_decomposition(int n, const double* F, const double* Lvec, double dt, double* A, double* Q);
void oracle_ioup_transition(int q, double rate, double h, double* Ah /*n*n*/, double* QhR /*n*n upper*/);
/* the same for a d x d column-major rate matrix: dense D x D column-major Ah and Qh.R; returns -1 if Qh is not positive definite */
int oracle_ioup_transition_matrix(int d, int q, const double* rate, double h, double* Ah, double* QhR);
void oracle_kron_identity(int d, int n_rows, int n_cols, const double* B, double* out /* (n_rows*d) x (n_cols*d) */);
int oracle_cholesky_upper(int n, double* M); /* in place; returns 0 on success */
void oracle_triangularize(int m, int n, double* A /* m x n, in place */, double* Rout /* n x n */);
void oracle_predict_cov(int D, const double* R, const double* Ah, const double* QhR, double diffusion,
                        int backend, double* Rout, int* used_qr);
int oracle_update(int D, int d, const double* mu_p, const double* R_p, const double* z, const double* H,
                  double* mu_out, double* R_out, double* loglik);
void oracle_backward_kernel(int D, const double* mu, const double* R, const double* mu_pred,
                            const double* R_pred, const double* Ah, const double* QhR, double diffusion,
                            double* G, double* b, double* LR /* 2D x D */);
void oracle_marginalize(int D, const double* mu_next, const double* R_next, const double* G, const double* b,
                        const double* LR /* 2D x D */, double* mu_out, double* R_out);

/* ---- solves ---- */
int oracle_solve(const oracle_config* cfg, oracle_rhs_fn f, oracle_jac_fn jac, oracle_taylor_fn taylor,
                 const double* u0, const double* p, oracle_traj* out);
void oracle_traj_free(oracle_traj* tr);

/* Ensemble: OpenMP over trajectories ("EnsembleThreads" restatement).  Returns per-trajectory
   final solution mean (d), final solution covariance (d*d, dense), stats.  Any pointer may be NULL. */
int oracle_solve_ensemble(const oracle_config* cfg, oracle_rhs_fn f, oracle_jac_fn jac, oracle_taylor_fn taylor,
                          int64_t n_traj, const double* u0 /*[n_traj][d]*/, const double* p /*[n_traj][n_p]*/,
                          int n_threads, double* u_final /*[n_traj][d]*/, double* cov_final /*[n_traj][d*d]*/,
                          int64_t* naccept, int64_t* nreject, int64_t* nf, int32_t* retcode, double* loglik);
int oracle_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
