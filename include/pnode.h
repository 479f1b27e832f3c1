/*
 * pnode.h -- C ABI of libpnode_cuda.so, the B200 (sm_100a) ODE-filter ensemble solver.
 *
 * This is the drop-in boundary for ProbNumDiffEq.jl's filter hot path.  A host package (Julia:
 * julia/ProbNumDiffEqB200, `EnsembleB200()`; here: the Python ctypes mirror in
 * probnumdiffeq.jl_b200/lib.py) gathers the per-trajectory u0/p, traces the user RHS to CUDA C++ and makes
 * ONE call per ensemble solve.  Each entry point names the reference interface it replaces
 * (paths relative to the reference checkout):
 *
 *   pnode_create            OrdinaryDiffEqCore.alg_cache(alg::AbstractEK, ...)     src/caches.jl:78-262
 *                           + covariance_structure / ekargcheck                     src/algorithms.jl:78-151
 *                           + make_measurement_model / calc_H! (RHS, Jacobian)      src/measurement_models.jl:11-20,
 *                                                                                    src/derivative_utils.jl:1-33
 *   pnode_solve_ensemble    SciMLBase.__solve(::EnsembleProblem, alg, ensemblealg; trajectories), i.e.
 *                           per trajectory: initialize! src/perform_step.jl:13-35, the solve! loop calling
 *                           perform_step! src/perform_step.jl:79-154, savevalues! src/integrator_utils.jl:143-171,
 *                           postamble! (calibrate_solution!, smooth_solution!) src/integrator_utils.jl:9-121
 *   pnode_result_*          ProbODESolution fields (t, u, pu, x_filt, x_smooth, diffusions, stats, pnstats, retcode)
 *                           src/solution.jl:33-56, build_solution src/solution.jl:89-151
 *   pnode_last_error        Julia exceptions thrown at construction (ArgumentError / DimensionMismatch /
 *                           ErrorException): src/algorithms.jl:87-128, src/caches.jl:96-147, src/checks.jl:1-21
 *
 * Conventions
 *   - All floating-point data is IEEE FP64.  State order is derivative-major: x[j*d + i] = j-th derivative of
 *     component i (src/preconditioning.jl:34).  D = d*(q+1).
 *   - Covariances are returned either dense (d x d, row-major, symmetric) or as *right* square-root factors
 *     R with Sigma = R'R (PSDMatrices.jl convention, src/filtering/predict.jl:15-16), stored row-major D x D.
 *     Factors are unique only up to an orthogonal transformation from the left: compare R'R, never R.
 *   - Inputs are caller-owned and only read during the call.  Results are library-owned handles; copy fields out
 *     with pnode_result_copy and release with pnode_result_free.  No pointer outlives its handle.
 *   - Every function returns 0 on success and a negative pnode_status otherwise; the message is available from
 *     pnode_last_error() (thread-local).  Nothing unwinds across the boundary.  Per-trajectory failures of the
 *     time loop are reported in the RETCODE field, like sol.retcode (src/solution.jl:55).
 *   - A pnode_solver is bound to one CUDA device and is not thread-safe; use one per thread / per GPU.
 */
#ifndef PNODE_H
#define PNODE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum pnode_status {
    PNODE_OK = 0,
    PNODE_ERR_ARGUMENT = -1,   /* Julia: ArgumentError     */
    PNODE_ERR_DIMENSION = -2,  /* Julia: DimensionMismatch */
    PNODE_ERR_COMPILE = -3,    /* NVRTC / module load      */
    PNODE_ERR_CUDA = -4,
    PNODE_ERR_UNSUPPORTED = -5,
    PNODE_ERR_NO_DEVICE = -6
} pnode_status;

typedef enum pnode_alg { PNODE_EK0 = 0, PNODE_EK1 = 1, PNODE_DIAGONAL_EK1 = 2 } pnode_alg; /* src/algorithms.jl:188-393 */
/* src/covariance_structure.jl:1-63; AUTO follows covariance_structure(alg, prior, diffusion) src/algorithms.jl:132-151 */
typedef enum pnode_cov { PNODE_COV_AUTO = 0, PNODE_COV_DENSE = 1, PNODE_COV_KRONECKER = 2, PNODE_COV_BLOCKS = 3 } pnode_cov;
/* src/priors/iwp.jl; src/priors/ioup.jl (scalar rate_parameter, or rate_matrix / update_rate_parameter below) */
typedef enum pnode_prior { PNODE_PRIOR_IWP = 0, PNODE_PRIOR_IOUP = 1 } pnode_prior;
/* src/diffusions/typedefs.jl: DynamicDiffusion, FixedDiffusion, DynamicMVDiffusion, FixedMVDiffusion */
typedef enum pnode_diffusion {
    PNODE_DIFF_DYNAMIC = 0,
    PNODE_DIFF_FIXED = 1,
    PNODE_DIFF_DYNAMIC_MV = 2,
    PNODE_DIFF_FIXED_MV = 3
} pnode_diffusion;
typedef enum pnode_retcode {
    PNODE_RET_SUCCESS = 0,
    PNODE_RET_MAXITERS = 1,
    PNODE_RET_DTLESSTHANMIN = 2,
    PNODE_RET_UNSTABLE = 3
} pnode_retcode;

/* Flattened EK0/EK1 constructor keywords (src/algorithms.jl:195-207, 258-295) + SciML solve keywords. */
typedef struct pnode_config {
    int32_t struct_size;       /* sizeof(pnode_config), for ABI checks */
    int32_t device;            /* CUDA device ordinal */
    int32_t d, q, n_p;         /* ODE dimension, prior order (num_derivatives), parameters per trajectory */
    int32_t alg;               /* pnode_alg */
    int32_t cov;               /* pnode_cov (AUTO: EK0 + scalar diffusion -> Kronecker, EK0 + multivariate diffusion and
                                  DiagonalEK1 -> blocks of diagonals, EK1 -> dense; src/algorithms.jl:132-151) */
    int32_t prior;             /* pnode_prior */
    int32_t diffusion;         /* pnode_diffusion */
    int32_t calibrate;         /* FixedDiffusion(calibrate=...) */
    int32_t smooth;            /* alg.smooth */
    int32_t adaptive;          /* solve(...; adaptive) */
    int32_t save_everystep;    /* solve(...; save_everystep) */
    int32_t save_state;        /* also return the full filter state (x_filt / x_smooth) not only pu */
    int32_t lanes_per_traj;    /* 0 = auto; 1 = one thread per trajectory (EK1, IWP, filter only, D <= 16); 2,4,8,16,32 lanes cooperating on one
                                  trajectory; 256 = one CTA per trajectory (the automatic choice for the EK1 with D > 32) */
    int32_t reserved0;
    int64_t maxiters;
    double initial_diffusion;
    double t0, t1, dt;         /* dt: fixed step, or initial step if adaptive (0 = automatic) */
    double abstol, reltol;
    double dtmin, dtmax;       /* dtmax <= 0: t1 - t0 */
    /* PI controller (OrdinaryDiffEqCore defaults when beta1/beta2 <= 0: 7/(10q), 2/(5q)) */
    double beta1, beta2, qmin, qmax, gamma, qsteady_min, qsteady_max, qoldinit;
    /* right-hand side: either the name of a built-in problem ("builtin:rober", ...) in rhs_name with rhs_src = NULL,
       or CUDA C++ source defining `struct <rhs_name>` with static d, n_p, f<T>, jac (see tracer.py) */
    const char* rhs_src;
    const char* rhs_name;
    /* optional extra stop times inside (t0, t1), ascending; host pointer, copied at create */
    const double* tstops;
    int64_t n_tstops;
    /* test hook: sigma^2 used for accepted step k instead of the local estimate (host pointer, copied) */
    const double* forced_diffusion;
    int64_t n_forced;
    /* IOUP prior: rate_parameter (F[q,q] of the SDE drift, src/priors/ioup.jl:66-79); ignored for the IWP.  Register-resident kernels for
       D <= 24; larger states (D <= 64) run as the rate matrix rate_parameter * I on the general-prior path (see rate_matrix) */
    double rate_parameter;
    /* dense output sol(saveat) (src/solution.jl:330-398): ascending times in [t0, t1]; host pointer, copied at create.
       Needs save_everystep; smoothed interpolation if smooth, else the filtering prediction */
    const double* saveat;
    int64_t n_saveat;
    /* constant mass matrix M of  M u' = f(u, p, t)  (`ODEFunction(f; mass_matrix=M)`): d x d, row-major, host pointer,
       copied at create; NULL = identity.  Measurement z = M E1 x - f(E0 x) (src/measurement_models.jl:11-20),
       H = M E1 - J E0 (src/derivative_utils.jl:1-53), initial derivatives as src/initialization/autodiffinit.jl:20-47.
       May be singular (index-1 DAEs).  EK1 on the dense layout (register kernels for D <= 32, the CTA-per-trajectory kernel beyond): any M; DiagonalEK1 / EK0 with a multivariate
       diffusion (blocks layout): diagonal M; EK0 on the Kronecker layout: M = lambda I only, otherwise
       PNODE_ERR_ARGUMENT with the reference's message (src/caches.jl:111-118). */
    const double* mass_matrix;
    /* SecondOrderODEProblem(f1, du0, u0, tspan, p)  (src/measurement_models.jl:29-49, src/caches.jl:100-126): 1 = `d` is the
       number of positions u; the state models u and its q >= 2 derivatives (D = d(q+1)); z = E2 x - f1(E1 x, E0 x);
       H = E2 (EK0) or E2 - J_du E1 - J_u E0 (EK1).  rhs_src / rhs_name describe the FIRST-ORDER FORM of dimension 2d,
       F([du; u]) = [f1(du, u); du] (what the reference's DynamicalODEFunction evaluates on ArrayPartition(du, u)); the
       kernels use its first d rows.  u0 is [n_traj][2d] = (du0, u0); every U field has 2d entries per point in the
       reference's sol.u order (du, u), every U_COV field (2d)^2.  EK0, DiagonalEK1 and EK1 (dense; D > 32 on the CTA-per-trajectory kernel:
       Pleiades as 14 second-order equations, D = 56; smoothing of second-order problems needs D <= 24) with
       every diffusion model the first-order path supports; IWP prior. */
    int32_t second_order;
    int32_t reserved1;
    /* Fenrir data log-likelihood, `fenrir_data_loglik(prob, alg; data=(t, u), observation_matrix, observation_noise_cov)`
       (src/data_likelihoods/fenrir.jl:30-137).  n_data > 0: the data times are added to tstops (:46), the solve keeps the
       filtering estimates (the reference stops with `step!`: no smoothing, no calibration), and a backward pass fits the
       PN posterior to the data (fit_pnsolution_to_data!, :68-125), one log-likelihood per trajectory in
       PNODE_F_FENRIR_LOGLIK (-Inf for a trajectory whose solve did not succeed, :55-59).  Needs smooth = 1
       (ArgumentError otherwise, :42-44).  One data set for the whole ensemble (parameter sweeps).  The value is the dense
       formula; with the EK0 the reference's Kronecker update omits the factor d of log det S (update.jl:147 quirk), see
       DESIGN.md section 5. */
    const double* data_t;             /* [n_data] ascending, in [t0, t1]; host pointer, copied at create            */
    int64_t n_data;
    const double* data_u;             /* [n_data][n_obs]                                                            */
    int32_t n_obs;                    /* length of one observation                                                  */
    int32_t data_forward;             /* 0: Fenrir (backward pass, above).  1: DataUpdateCallback (src/callbacks/dataupdate.jl:39-84)
                                         as used by filtering_data_loglik / dalton_data_loglik (src/data_likelihoods/filtering.jl,
                                         dalton.jl): whenever an accepted step ends on a data time the FILTER state is updated on
                                         y = H x + eps, eps ~ N(0, R), before it is saved, and the updates' log-likelihoods are
                                         summed per trajectory in PNODE_F_DATA_LOGLIK (DataUpdateLogLikelihood.ll).  smooth is free.
                                         Dense layout (EK1, D <= 32); partial observations and matrix noise allowed. */
    const double* observation_matrix; /* [n_obs][len(sol.u)] row-major; NULL = identity (n_obs == len(sol.u))        */
    const double* observation_noise_cov; /* [n_obs][n_obs] symmetric positive definite; NULL = observation_noise_var * I */
    double observation_noise_var;
    /* EK0 / EK1 / DiagonalEK1(pn_observation_noise = R) (src/algorithms.jl:195-393; cov2psdmatrix, src/caches.jl:175-178): the ODE
       residual z = 0 is observed with noise covariance R.  The measurement covariance gains R (src/perform_step.jl:182-188) and the
       filtered covariance K R K', re-factorised by Cholesky with triangularize! as the fallback (src/filtering/update.jl:125-136).
       pn_observation_noise: [d][d] row-major symmetric positive definite, host pointer, copied at create; NULL with
       pn_observation_noise_var > 0: R = var * I (a Number / UniformScaling); both unset: noise-free.  As in ekargcheck
       (src/algorithms.jl:78-130): not with FixedDiffusion(calibrate = true) / FixedMVDiffusion(calibrate = true); the Kronecker
       layout (EK0 + scalar diffusion) takes isotropic R only, the blocks layout (DiagonalEK1, EK0 + multivariate diffusion)
       diagonal R only -- PNODE_ERR_ARGUMENT with the reference's messages. */
    const double* pn_observation_noise;
    double pn_observation_noise_var;
    /* IOUP(num_derivatives, rate_parameter::Union{AbstractVector,AbstractMatrix}) (src/priors/ioup.jl:81-113): [d][d] row-major rate
       matrix of the SDE drift, F[q-block, q-block] = rate (a vector rate is passed as its diagonal matrix, as to_sde does); host
       pointer, copied at create; NULL = the scalar rate_parameter.  IOUP(num_derivatives; update_rate_parameter = true)
       (RosenbrockExpEK, src/algorithms.jl:438-459): update_rate_parameter = 1 sets the rate to the Jacobian J(uprev, t) before every
       step (src/perform_step.jl:88-96); rate_matrix must then be NULL ("Do not manually set the `rate_parameter` ...",
       src/caches.jl:139-149) and the EK1 is required (the Jacobian is the traced one).  Both run on the general-prior path
       (one warp per trajectory, dense transition pairs; filter and smoother; EK0 and EK1 with scalar diffusion models). */
    const double* rate_matrix;
    int32_t update_rate_parameter;
    int32_t reserved2;
} pnode_config;

typedef struct pnode_solver pnode_solver;
typedef struct pnode_result pnode_result;

typedef enum pnode_field {
    PNODE_F_T_END = 0,        /* double [n_traj]              */
    PNODE_F_U_FINAL = 1,      /* double [n_traj][d]           sol.u[end]                 */
    PNODE_F_U_FINAL_COV = 2,  /* double [n_traj][d][d]        Matrix(sol.pu[end].Sigma)  */
    PNODE_F_X_FINAL = 3,      /* double [n_traj][D]           sol.x_filt[end].mu   (save_state) */
    PNODE_F_X_FINAL_COVFAC = 4, /* double [n_traj][D][D]      sol.x_filt[end].Sigma.R    (save_state) */
    PNODE_F_LOGLIK = 5,       /* double [n_traj]              sol.pnstats.log_likelihood */
    PNODE_F_DIFFUSION = 6,    /* double [n_traj]              final / global diffusion   */
    PNODE_F_DT_LAST = 7,      /* double [n_traj]                                          */
    PNODE_F_NACCEPT = 8,      /* int64  [n_traj]              sol.stats.naccept          */
    PNODE_F_NREJECT = 9,      /* int64  [n_traj]              sol.stats.nreject          */
    PNODE_F_NF = 10,          /* int64  [n_traj]              sol.stats.nf               */
    PNODE_F_RETCODE = 11,     /* int32  [n_traj]              sol.retcode                */
    /* save_everystep (CSR over trajectories) */
    PNODE_F_STEP_OFFSETS = 12, /* int64 [n_traj+1]            saved points of trajectory i are [off[i], off[i+1]) */
    PNODE_F_T = 13,            /* double [total]              sol.t                      */
    PNODE_F_U = 14,            /* double [total][d]           sol.u  (smoothed if smooth) */
    PNODE_F_U_COV = 15,        /* double [total][d][d]        Matrix(sol.pu.Sigma)       */
    PNODE_F_DIFFUSIONS = 16,   /* double [total]              sol.diffusions (slot k: step k -> k+1) */
    PNODE_F_X = 17,            /* double [total][D]           sol.x_filt / sol.x_smooth mean (save_state; not produced by the
                                  CTA-per-trajectory kernel, D > 32, which returns the final state only) */
    PNODE_F_X_COVFAC = 18,     /* double [total][D][D]        ... right factor            (save_state) */
    /* multivariate diffusion models (Diagonal(sigma_1^2 .. sigma_d^2), src/diffusions/typedefs.jl) */
    PNODE_F_DIFFUSION_MV = 19,  /* double [n_traj][d]         final / global diffusion per component */
    PNODE_F_DIFFUSIONS_MV = 20, /* double [total][d]          sol.diffusions per component           */
    /* dense output on the saveat grid */
    PNODE_F_SAVEAT_U = 21,      /* double [n_traj][n_saveat][d]      mean(sol(saveat[j]))            */
    PNODE_F_SAVEAT_U_COV = 22,  /* double [n_traj][n_saveat][d][d]   Matrix(sol(saveat[j]).Sigma)    */
    PNODE_F_FENRIR_LOGLIK = 23, /* double [n_traj]  Fenrir data log-likelihood (config data_t / data_u) */
    PNODE_F_DATA_LOGLIK = 24,   /* double [n_traj]  sum of the forward data updates' log-likelihoods (data_forward = 1) */
    PNODE_F_COUNT
} pnode_field;

typedef struct pnode_result_info {
    int64_t n_traj;
    int64_t total_saved;       /* sum of saved points (0 if !save_everystep) */
    int64_t total_accepted;    /* sum over trajectories of naccept */
    int64_t total_rejected;
    int32_t d, q, D;
    int32_t lanes_per_traj;
    double kernel_ms;          /* device time of the solve kernels (CUDA events on the solver's stream) */
    double h2d_ms, d2h_ms;     /* host<->device staging inside pnode_solve_ensemble */
    int64_t n_launches;        /* kernels launched by this solve */
    int64_t h2d_bytes, d2h_bytes;
} pnode_result_info;

/* Number of CUDA devices visible (0 if none / no driver).  Replaces nothing in the reference. */
int pnode_device_count(void);
/* Thread-local message of the last failing call ("" if none). */
const char* pnode_last_error(void);
/* ABI version of this header. */
int pnode_abi_version(void);

/* Validate the configuration, compile (NVRTC, sm_100a) or look up the kernels, allocate the work queue.  NVRTC results are cached per
   process (key: generated program text + compile-time options), so a second solver of the same right-hand side and compile-time
   configuration -- another tspan, tolerance, ensemble -- does not compile again (PNODE_NVRTC_CACHE=0 turns the cache off). */
int pnode_create(const pnode_config* cfg, pnode_solver** out);
void pnode_destroy(pnode_solver* s);
/* Validate the configuration and run NVRTC only (no device needed); *cubin_bytes = size of the sm_100a CUBIN
   (0 for built-in problems).  Same argument errors as pnode_create (src/algorithms.jl:78-130). */
int pnode_compile_check(const pnode_config* cfg, int64_t* cubin_bytes);
/* Seconds spent in NVRTC + module load for this solver (reported separately from solve time). */
double pnode_compile_seconds(const pnode_solver* s);

/* Measured FP64 FMA throughput of the device in TFLOP/s (DFMA-saturating probe kernel); the denominator of the
   FP64 roofline fraction (MEASURED_PEAKS.json carries no FP64 figure). */
int pnode_measure_fp64_peak(int device, double* tflops);

/* Solve n_traj independent trajectories.  u0: [n_traj][d], p: [n_traj][n_p] (host pointers). */
int pnode_solve_ensemble(pnode_solver* s, int64_t n_traj, const double* u0, const double* p, pnode_result** out);
/* Same with inputs already resident on the solver's device (device pointers); outputs stay on the device until
   pnode_result_copy.  Used to time the kernels without staging. */
int pnode_solve_ensemble_device(pnode_solver* s, int64_t n_traj, const double* d_u0, const double* d_p,
                                pnode_result** out);

/* One ensemble over several devices of one node (replaces the reference's single-process
   EnsembleThreads loop, SciMLBase `solve(::EnsembleProblem, alg, ensemblealg; trajectories)`): solvers[i] -- created from the
   same configuration with `device` = the i-th GPU -- solves the contiguous trajectory range pnode_shard_range(n_traj, i,
   n_solvers) on its own host thread and stream; there is no exchange between the shards.  results[i] receives shard i
   (concatenate in shard order: saved-point offsets of shard i are relative to that shard).  On failure nothing is returned
   and the first failing shard's message is in pnode_last_error(). */
int pnode_solve_ensemble_sharded(pnode_solver* const* solvers, int32_t n_solvers, int64_t n_traj, const double* u0,
                                 const double* p, pnode_result** results /* [n_solvers] */);
/* [lo, hi) of shard `shard` out of `n_shards`: floor split, the remainder goes to the low shards. */
void pnode_shard_range(int64_t n_traj, int32_t shard, int32_t n_shards, int64_t* lo, int64_t* hi);

int pnode_result_info_get(const pnode_result* r, pnode_result_info* info);
/* Bytes needed for a field (0 if the field was not produced). */
int64_t pnode_result_field_bytes(const pnode_result* r, int field);
/* Copy a field into caller memory (host). bytes must equal pnode_result_field_bytes. */
int pnode_result_copy(const pnode_result* r, int field, void* dst, size_t bytes);
/* Device pointer of a field (NULL if absent); valid until pnode_result_free. */
const void* pnode_result_device_ptr(const pnode_result* r, int field);
void pnode_result_free(pnode_result* r);

#ifdef __cplusplus
}
#endif
#endif /* PNODE_H */
