// pn_params.h -- plain-old-data kernel parameter block shared by host (pnode_api.cu) and device code.
// NVRTC-compatible: no standard headers.
#ifndef PN_PARAMS_H
#define PN_PARAMS_H

#define PN_MAX_N 12 /* q + 1 <= 12 */
#define PN_INF (__longlong_as_double(0x7ff0000000000000LL))

enum PnRetcode { PN_RET_SUCCESS = 0, PN_RET_MAXITERS = 1, PN_RET_DTLESSTHANMIN = 2, PN_RET_UNSTABLE = 3 };

struct PnParams {
    long long n_traj;
    const double* u0; /* [n_traj][d]   */
    const double* p;  /* [n_traj][n_p] */
    double t0, t1, dt0;
    double abstol, reltol, dtmin, dtmax;
    double beta1, beta2, qmin, qmax, gamma, qsteady_min, qsteady_max, qoldinit;
    double qoldinit_pb2; /* qoldinit^beta2 (host-computed) */
    double inv_qmax, inv_qmin, inv_gamma; /* 1/qmax, 1/qmin, 1/gamma (host-computed, IEEE divisions) */
    double initial_diffusion;
    long long maxiters;
    int calibrate;
    int order; /* controller order = q (src/alg_utils.jl:45) */
    /* optional schedules (device pointers or null) */
    const double* tstops;
    long long n_tstops;
    const double* forced_diffusion;
    long long n_forced;
    /* work item idx of the queue is trajectory idx * traj_stride (1 for a normal solve; > 1 for the step-count pilot that
       sizes the save buffers from a strided subsample of the ensemble) */
    long long traj_stride;
    /* initial states computed by pn_init_kernel before the step kernel starts (group-per-trajectory path): exact initial
       derivatives mu0[n_traj][D] and the initial step dt0[n_traj]; null: the step kernel initialises in line */
    const double* init_mu;
    const double* init_dt;
    /* dynamic work queue */
    unsigned long long* work_counter;
    /* IWP constants (src/priors/iwp.jl:74-108), row-major n x n, host-computed */
    double L[PN_MAX_N * PN_MAX_N];
    /* upper-triangular U with U'U = L'L (Householder QR of L in extended precision on the host), row-major n x n:
       an equivalent right factor of the same preconditioned process-noise covariance, used where an upper-triangular
       noise block lets a sweep skip its leading zero columns */
    double U[PN_MAX_N * PN_MAX_N];
    /* Qb = L'L, the preconditioned IWP process-noise covariance itself (iwp.jl:84-108), row-major n x n: the Gram + Cholesky
       route of predict_cov! (predict.jl:84-85) adds s2 (P^-1 Qb P^-1) (x) I_d to X'X */
    double Qb[PN_MAX_N * PN_MAX_N];
    /* IOUP prior (src/priors/ioup.jl): scalar rate parameter, 16-point Gauss-Legendre rule on [0, 1] (host-computed) */
    double rate;
    double glx[16], glw[16];
    /* forward data updates (DataUpdateCallback, src/callbacks/dataupdate.jl:39-84; kernels built with PN_DATA): observation times
       (all of them are tstops), observations [n_data][PN_DATA_O], and the per-trajectory sum of the updates' log-likelihoods */
    const double* data_t;
    const double* data_u;
    long long n_data;
    double* data_loglik; /* [n_traj] */
    /* general-prior filter (pn_gen.cuh): IOUP with a vector / matrix rate parameter (d x d row-major device array; null: rate I_d),
       Rosenbrock-style rate update (rate := J(uprev, t), perform_step.jl:88-96), per-warp workspace (null: shared memory) */
    const double* rate_matrix;
    int update_rate;
    double* gen_workspace;
    double* s_rate; /* [total][d][d] rate matrix of the step k -> k+1 at slot k (smoothing with update_rate only; null otherwise) */
    /* ---- per-trajectory outputs ---- */
    double* t_end;      /* [n_traj]       */
    double* u_mean;     /* [n_traj][d]    */
    double* u_cov;      /* [n_traj][d][d] */
    double* x_mean;     /* [n_traj][D]    (may be null) */
    double* x_covfac;   /* [n_traj][D][D] row-major right factor, Sigma = R'R (may be null) */
    double* loglik;     /* [n_traj] */
    double* diffusion;  /* [n_traj] final diffusion (global MLE for FixedDiffusion, last local otherwise) */
    double* diffusion_mv; /* [n_traj][d] per-component diffusion (multivariate models; null otherwise) */
    double* dt_last;    /* [n_traj] */
    long long* naccept; /* [n_traj] */
    long long* nreject;
    long long* nf;
    int* retcode;
    /* ---- save_everystep outputs (CSR over trajectories; null in the counting pass) ---- */
    const long long* step_offsets; /* [n_traj+1] saved points per trajectory incl. the initial one */
    double* s_t;                   /* [total]         */
    double* s_u_mean;              /* [total][d]      */
    double* s_u_cov;               /* [total][d][d]   */
    double* s_diffusion;           /* [total]         */
    double* s_diffusion_mv;        /* [total][d]      (multivariate diffusion models; null otherwise) */
    double* s_x_mean;              /* [total][D]      (smoothing only) */
    double* s_x_covfac;            /* [total][D][D]   (smoothing only) */
    double* s_h;                   /* [total] step size of the step k -> k+1 at slot k (smoothing only) */
};

#endif /* PN_PARAMS_H */
