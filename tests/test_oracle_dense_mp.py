"""The oracle against an independent implementation that shares nothing with it: the textbook (covariance-form) Gaussian ODE filter and
RTS smoother in plain coordinates -- no square roots, no preconditioning, closed-form IWP transition matrices -- evaluated with 120
significant digits (mpmath), exact initial derivatives by symbolic differentiation.  In double precision that formulation loses the
solution after a few dozen steps (the process-noise covariance has entries h^(2q+1)); with enough digits (the lost digits grow with the number of steps: 120 carry the runs below with a wide margin) it is the exact answer the
square-root algorithm of the reference (src/filtering/predict.jl, update.jl, markov_kernel.jl) approximates.

Pins, over whole fixed-step trajectories: filtered and smoothed means and covariances, the local diffusions, the running global
diffusion MLE with the reference's divisor (src/diffusions/calibration.jl:56-62) and its a-posteriori scaling of the covariances
(src/integrator_utils.jl:46-65), the log-likelihood (src/filtering/update.jl:140-148) and, for the EK0, the reference's Kronecker
log-determinant as written (update.jl:147: log det of the 1 x 1 factor, not multiplied by d)."""
import mpmath as mp
import numpy as np
import pytest
import sympy as sp

from dense_reference import dense_filter_smoother

import probnumdiffeq_jl_b200  # noqa: F401
from probnumdiffeq_jl_b200 import problems, tracer


def _cov(R):
    return np.einsum("kri,krj->kij", R, R)


def _relc(a, b):
    return np.max(np.linalg.norm(a - b, axis=(1, 2)) / np.linalg.norm(b, axis=(1, 2)))


def _fhn(du, u, p, t):      # FitzHugh-Nagumo of README.md:35-43 with exact rational arithmetic for the symbolic side
    du[0] = p[2] * (u[0] - u[0] ** 3 / 3 + u[1])
    du[1] = -(u[0] - p[0] - p[1] * u[1]) / p[2]


CASES = [("lotka_volterra", problems.lotka_volterra, [1.0, 1.0], [1.5, 1.0, 3.0, 1.0], 3, "0.02", 50),
         ("fitzhugh_nagumo_q2", _fhn, [-1.0, 1.0], [0.2, 0.2, 3.0], 2, "0.01", 60),
         ("fitzhugh_nagumo_q5", _fhn, [-1.0, 1.0], [0.2, 0.2, 3.0], 5, "0.02", 40)]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("alg", ["EK1", "EK0"])
def test_fixed_diffusion_trajectory(oracle, case, alg):
    """FixedDiffusion, calibrated: filtered states (uncalibrated run: solution components 1e-11, all derivatives 1e-9), then the calibrated smoothed solution -- every mean and
    every covariance (scaled by the global MLE) -- the MLE itself and the log-likelihood.  Both covariance back-ends of the oracle."""
    mp.mp.dps = 120
    _, f, u0, p, q, h, N = case
    d = len(u0)
    ref = dense_filter_smoother(f, u0, p, q, h, N, alg == "EK1", False)
    rhs = oracle.compile_rhs(tracer.trace(f, d, len(p)), q)
    oalg = oracle.EK1 if alg == "EK1" else oracle.EK0
    for backend in (oracle.COV_REF, oracle.COV_QR):
        kw = dict(alg=oalg, diffusion=oracle.DIFF_FIXED, adaptive=False, dt=float(h), t0=0.0, t1=float(h) * N, save_everystep=True, cov_backend=backend)
        s = oracle.solve(oracle.make_config(d, q, len(p), calibrate=False, smooth=False, **kw), rhs, u0, p)
        assert s["n"] == N + 1 and s["retcode"] == "Success"
        scale = np.abs(ref["fm"]).max(axis=0)
        assert np.max(np.abs(s["x_filt_mu"] - ref["fm"])[:, :d] / scale[:d]) < 1e-11      # the solution components
        assert np.max(np.abs(s["x_filt_mu"] - ref["fm"]) / scale) < 1e-9                  # all derivatives (the q-th carries the rounding)
        assert _relc(_cov(s["x_filt_R"])[1:], ref["fP"][1:]) < 1e-9
        assert s["global_diffusion"] == pytest.approx(ref["gd"], rel=1e-9)
        assert s["loglik"] == pytest.approx(ref["ll"] if alg == "EK1" else ref["ll_kron"], rel=1e-10)
        s = oracle.solve(oracle.make_config(d, q, len(p), calibrate=True, smooth=True, **kw), rhs, u0, p)
        assert np.max(np.abs(s["pu_mu"] - ref["sm"][:, :d]) / scale[:d]) < 1e-10
        assert _relc(s["pu_cov"][1:], ref["gd"] * ref["sP"][1:, :d, :d]) < 1e-8


@pytest.mark.parametrize("case", CASES[:2], ids=[c[0] for c in CASES[:2]])
def test_dynamic_diffusion_trajectory(oracle, case):
    """DynamicDiffusion (EK1): sigma^2_k = z'(H Q H')^-1 z / d is a ratio of rounding-sensitive quantities (the residual z is pure cancellation), so
    the local diffusions are compared at 1e-6, the means -- which a wrong sigma^2 would move -- at 1e-9, filtered and smoothed covariances at 1e-6."""
    mp.mp.dps = 120
    _, f, u0, p, q, h, N = case
    d = len(u0)
    ref = dense_filter_smoother(f, u0, p, q, h, N, True, True)
    rhs = oracle.compile_rhs(tracer.trace(f, d, len(p)), q)
    kw = dict(alg=oracle.EK1, diffusion=oracle.DIFF_DYNAMIC, adaptive=False, dt=float(h), t0=0.0, t1=float(h) * N, save_everystep=True)
    s = oracle.solve(oracle.make_config(d, q, len(p), smooth=False, **kw), rhs, u0, p)
    scale = np.abs(ref["fm"]).max(axis=0)
    assert np.max(np.abs(s["diffusions"][:N] - ref["sig"]) / ref["sig"]) < 1e-6
    assert np.max(np.abs(s["x_filt_mu"] - ref["fm"]) / scale) < 1e-9
    assert _relc(_cov(s["x_filt_R"])[1:], ref["fP"][1:]) < 1e-6
    assert s["loglik"] == pytest.approx(ref["ll"], rel=1e-8)
    s = oracle.solve(oracle.make_config(d, q, len(p), smooth=True, **kw), rhs, u0, p)
    assert np.max(np.abs(s["pu_mu"] - ref["sm"][:, :d]) / scale[:d]) < 1e-9
    assert _relc(s["pu_cov"][1:], ref["sP"][1:, :d, :d]) < 1e-6


def _check(s, ref, d, cols, tol_mu=1e-10, tol_cov=1e-8, ll=None):
    scale = np.abs(ref["fm"]).max(axis=0)
    assert s["retcode"] == "Success" and s["n"] == len(ref["fm"])
    assert np.max(np.abs(s["x_filt_mu"] - ref["fm"])[:, :d] / scale[:d]) < tol_mu
    assert np.max(np.abs(s["x_filt_mu"] - ref["fm"]) / scale) < 100 * tol_mu
    assert _relc(_cov(s["x_filt_R"])[1:], ref["fP"][1:]) < tol_cov
    assert s["loglik"] == pytest.approx(ref["ll"] if ll is None else ll, rel=1e-9)


@pytest.mark.parametrize("rate", [-0.7, np.array([[-1.0, 0.3], [-0.2, -0.5]])], ids=["scalar", "matrix"])
def test_ioup_prior_trajectory(oracle, rate):
    """IOUP(3, rate) (src/priors/ioup.jl:66-131): the oracle's transition pair (Pade matrix exponential of the Van Loan matrix in
    preconditioned coordinates; scalar rates by the matrix-fraction decomposition) against the 60-digit matrix exponential in plain
    coordinates, through a whole filtered + smoothed trajectory."""
    mp.mp.dps = 120
    f, u0, p, q, h, N = problems.lotka_volterra, [1.0, 1.0], [1.5, 1.0, 3.0, 1.0], 3, "0.02", 40
    R = rate * np.eye(2) if np.isscalar(rate) else rate
    ref = dense_filter_smoother(f, u0, p, q, h, N, True, False, rate=R)
    rhs = oracle.compile_rhs(tracer.trace(f, 2, 4), q)
    kw = dict(alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, calibrate=False, adaptive=False, dt=float(h), t0=0.0, t1=float(h) * N, save_everystep=True,
              prior=oracle.PRIOR_IOUP, rate_parameter=rate)
    _check(oracle.solve(oracle.make_config(2, q, 4, smooth=False, **kw), rhs, u0, p), ref, 2, None)
    s = oracle.solve(oracle.make_config(2, q, 4, smooth=True, **kw), rhs, u0, p)
    assert np.max(np.abs(s["pu_mu"] - ref["sm"][:, :2])) < 1e-10
    assert _relc(s["pu_cov"][1:], ref["sP"][1:, :2, :2]) < 1e-8


def test_mass_matrix_trajectory(oracle):
    """M u' = f with a nonsingular, non-diagonal M (src/measurement_models.jl:11-20, src/derivative_utils.jl:43-50): z = M E1 x - f, H = M E1 - J E0,
    started from the reference's initial state (src/initialization/autodiffinit.jl:16-48: M x^(o) = Taylor coefficients of u' = f)."""
    mp.mp.dps = 120
    f, u0, p, q, h, N = problems.lotka_volterra, [1.0, 1.0], [1.5, 1.0, 3.0, 1.0], 3, "0.02", 40
    M = np.array([[2.0, 0.5], [0.25, 1.5]])
    ref = dense_filter_smoother(f, u0, p, q, h, N, True, False, mass=M)
    rhs = oracle.compile_rhs(tracer.trace(f, 2, 4), q)
    kw = dict(alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, calibrate=False, adaptive=False, dt=float(h), t0=0.0, t1=float(h) * N, save_everystep=True,
              mass_matrix=M)
    _check(oracle.solve(oracle.make_config(2, q, 4, smooth=False, **kw), rhs, u0, p), ref, 2, None)


def _twobody(dx, x, p, t):      # first-order form of twobody2 (test/secondorderode.jl:15-19): x = (v1, v2, r1, r2)
    r3 = (x[2] * x[2] + x[3] * x[3]) ** sp.Rational(3, 2) if isinstance(x[0], sp.Basic) else (x[2] * x[2] + x[3] * x[3]) ** 1.5
    dx[0] = -x[2] / r3
    dx[1] = -x[3] / r3
    dx[2] = x[0]
    dx[3] = x[1]


@pytest.mark.parametrize("alg", ["EK1", "EK0"])
def test_second_order_trajectory(oracle, alg):
    """SecondOrderODEProblem (src/measurement_models.jl:29-49, src/derivative_utils.jl:35-53, SolProj = [E1; E0] src/caches.jl:100-126): z = E2 x -
    f1(E1 x, E0 x), H = E2 - J_du E1 - J_u E0 (EK0: H = E2); sol.u = (du, u)."""
    mp.mp.dps = 120
    x0, q, h, N = [0.0, 2.0, 0.4, 0.0], 3, "0.002", 40
    ref = dense_filter_smoother(_twobody, x0, [], q, h, N, alg == "EK1", False, second_order=True)
    rhs = oracle.compile_rhs(tracer.trace(_twobody, 4, 0), q)
    kw = dict(alg=oracle.EK1 if alg == "EK1" else oracle.EK0, diffusion=oracle.DIFF_FIXED, calibrate=False, adaptive=False, dt=float(h), t0=0.0,
              t1=float(h) * N, save_everystep=True, second_order=True)
    s = oracle.solve(oracle.make_config(2, q, 0, smooth=False, **kw), rhs, x0, [])
    _check(s, ref, 2, None, ll=ref["ll"] if alg == "EK1" else ref["ll_kron"])
    assert np.max(np.abs(s["pu_mu"] - np.hstack([ref["fm"][:, 2:4], ref["fm"][:, 0:2]]))) < 1e-10      # (du, u)
    s = oracle.solve(oracle.make_config(2, q, 0, smooth=True, **kw), rhs, x0, [])
    assert np.max(np.abs(s["pu_mu"] - np.hstack([ref["sm"][:, 2:4], ref["sm"][:, 0:2]]))) < 1e-10


@pytest.mark.parametrize("noise", [0.05, np.array([[0.05, 0.01], [0.01, 0.08]])], ids=["scalar", "matrix"])
def test_pn_observation_noise_trajectory(oracle, noise):
    """pn_observation_noise R (src/perform_step.jl:182-188, src/filtering/update.jl:125-136): S = H P^- H' + R and the Joseph-form covariance, through
    a whole trajectory (DynamicDiffusion: the reference rejects global calibration with observation noise, src/algorithms.jl:87-91)."""
    mp.mp.dps = 120
    f, u0, p, q, h, N = problems.lotka_volterra, [1.0, 1.0], [1.5, 1.0, 3.0, 1.0], 3, "0.02", 40
    R = noise * np.eye(2) if np.isscalar(noise) else noise
    ref = dense_filter_smoother(f, u0, p, q, h, N, True, True, obs_noise=R)
    rhs = oracle.compile_rhs(tracer.trace(f, 2, 4), q)
    kw = dict(alg=oracle.EK1, diffusion=oracle.DIFF_DYNAMIC, adaptive=False, dt=float(h), t0=0.0, t1=float(h) * N, save_everystep=True,
              pn_observation_noise=noise)
    s = oracle.solve(oracle.make_config(2, q, 4, smooth=False, **kw), rhs, u0, p)
    assert np.max(np.abs(s["diffusions"][:N] - ref["sig"]) / ref["sig"]) < 1e-6
    _check(s, ref, 2, None, tol_mu=1e-9, tol_cov=1e-6)


def test_blocks_layout_trajectories(oracle):
    """The BlocksOfDiagonals configurations (src/algorithms.jl:132-151): DiagonalEK1 with FixedDiffusion, EK0 with DynamicMVDiffusion, EK0 with
    FixedMVDiffusion (calibrated: the smoothed covariance of component i is scaled by its own MLE)."""
    mp.mp.dps = 120
    f, u0, p, q, h, N = problems.lotka_volterra, [1.0, 1.0], [1.5, 1.0, 3.0, 1.0], 3, "0.02", 40
    rhs = oracle.compile_rhs(tracer.trace(f, 2, 4), q)
    base = dict(adaptive=False, dt=float(h), t0=0.0, t1=float(h) * N, save_everystep=True)
    # DiagonalEK1, FixedDiffusion
    ref = dense_filter_smoother(f, u0, p, q, h, N, True, False, diag_jac=True)
    kw = dict(alg=oracle.DIAGONAL_EK1, diffusion=oracle.DIFF_FIXED, calibrate=False, **base)
    _check(oracle.solve(oracle.make_config(2, q, 4, smooth=False, **kw), rhs, u0, p), ref, 2, None)
    s = oracle.solve(oracle.make_config(2, q, 4, smooth=True, **kw), rhs, u0, p)
    assert np.max(np.abs(s["pu_mu"] - ref["sm"][:, :2])) < 1e-10 and _relc(s["pu_cov"][1:], ref["sP"][1:, :2, :2]) < 1e-8
    # EK0, DynamicMVDiffusion
    ref = dense_filter_smoother(f, u0, p, q, h, N, False, True, mv=True)
    kw = dict(alg=oracle.EK0, diffusion=oracle.DIFF_DYNAMIC_MV, **base)
    s = oracle.solve(oracle.make_config(2, q, 4, smooth=False, **kw), rhs, u0, p)
    assert np.max(np.abs(s["diffusions_mv"][:N] - ref["sig"]) / ref["sig"]) < 1e-6
    _check(s, ref, 2, None, tol_mu=1e-9, tol_cov=1e-6)
    s = oracle.solve(oracle.make_config(2, q, 4, smooth=True, **kw), rhs, u0, p)
    assert np.max(np.abs(s["pu_mu"] - ref["sm"][:, :2])) < 1e-9 and _relc(s["pu_cov"][1:], ref["sP"][1:, :2, :2]) < 1e-6
    # EK0, FixedMVDiffusion, calibrated
    ref = dense_filter_smoother(f, u0, p, q, h, N, False, False)
    kw = dict(alg=oracle.EK0, diffusion=oracle.DIFF_FIXED_MV, calibrate=True, **base)
    s = oracle.solve(oracle.make_config(2, q, 4, smooth=True, **kw), rhs, u0, p)
    assert np.max(np.abs(s["pu_mu"] - ref["sm"][:, :2])) < 1e-10
    want = ref["sP"][1:, :2, :2] * np.sqrt(np.outer(ref["gd_mv"], ref["gd_mv"]))[None]
    assert _relc(s["pu_cov"][1:], want) < 1e-8


@pytest.mark.parametrize("dynamic", [False, True], ids=["fixed", "dynamic"])
def test_dense_output(oracle, dynamic):
    """sol(t) between the saved points (src/solution.jl:330-398, src/filtering/smooth.jl:43-65) against the 60-digit evaluation of the same
    two-stage formula: means 1e-9, covariances 1e-6 (dynamic diffusion) / 1e-8."""
    mp.mp.dps = 120
    f, u0, p, q, h, N = problems.lotka_volterra, [1.0, 1.0], [1.5, 1.0, 3.0, 1.0], 3, "0.02", 30
    ts = ["0.005", "0.031", "0.3", "0.4137", "0.59"]
    ref = dense_filter_smoother(f, u0, p, q, h, N, True, dynamic, dense_ts=ts)
    rhs = oracle.compile_rhs(tracer.trace(f, 2, 4), q)
    cfg = oracle.make_config(2, q, 4, alg=oracle.EK1, diffusion=oracle.DIFF_DYNAMIC if dynamic else oracle.DIFF_FIXED, calibrate=False, smooth=True,
                             adaptive=False, dt=float(h), t0=0.0, t1=float(h) * N, save_everystep=True)
    s = oracle.solve(cfg, rhs, u0, p)
    for k, tv in enumerate(ts):
        mu, cov = oracle.interpolate(cfg, s, float(tv))
        assert np.max(np.abs(mu - ref["dm"][k, :2])) < 1e-9
        assert np.linalg.norm(cov - ref["dP"][k, :2, :2]) / np.linalg.norm(ref["dP"][k, :2, :2]) < (1e-6 if dynamic else 1e-8)


def test_variable_steps_teacher_forced(oracle):
    """The accepted step sequence of an ADAPTIVE oracle solve, replayed by the 60-digit filter / smoother with one transition pair per step: pins the
    variable-step algebra (preconditioner and transition matrices recomputed per step, src/perform_step.jl:37-50) independently of the step-size
    controller, which is third-party code restated from SURVEY.md A.6."""
    mp.mp.dps = 120
    f, u0, p, q = problems.lotka_volterra, [1.0, 1.0], [1.5, 1.0, 3.0, 1.0], 3
    rhs = oracle.compile_rhs(tracer.trace(f, 2, 4), q)
    s = oracle.solve(oracle.make_config(2, q, 4, alg=oracle.EK1, diffusion=oracle.DIFF_DYNAMIC, smooth=True, adaptive=True, t0=0.0, t1=3.0), rhs, u0, p)
    assert s["retcode"] == "Success" and 20 < s["naccept"] < 200 and np.ptp(np.diff(s["t"])) > 1e-3      # genuinely variable steps
    hs = np.diff(s["t"])
    ref = dense_filter_smoother(f, u0, p, q, hs, len(hs), True, True, tols=(1e-6, 1e-3))
    # every accepted step has an error estimate below one, and -- no attempt of this solve was rejected -- each step size is the proposal of
    # the PI controller (OrdinaryDiffEqCore, restated in SURVEY.md A.6: q = EEst^b1 / qold^b2 / gamma clamped to [1/qmax, 1/qmin], b1 = 7/(10 order),
    # b2 = 2/(5 order), qold = max(EEst_prev, 1e-4)) evaluated here from the 60-digit error estimates; the last step is clipped to the end point
    E = ref["eest"]
    assert s["nreject"] == 0 and np.all(E < 1.0)
    b1, b2, qold = 0.7 / q, 0.4 / q, 1e-4
    for k in range(len(hs) - 2):
        fac = min(1.0 / 0.2, max(1.0 / 10.0, E[k] ** b1 / qold ** b2 / 0.9))
        qold = max(E[k], 1e-4)
        assert hs[k + 1] == pytest.approx(hs[k] / fac, rel=1e-6), k
    # the oracle accumulates t += h in floating point; the replay uses the differences of its saved times, so the two step sequences agree to an ulp
    assert np.max(np.abs(s["diffusions"][:len(hs)] - ref["sig"]) / ref["sig"]) < 1e-5
    assert np.max(np.abs(s["pu_mu"] - ref["sm"][:, :2])) < 1e-8
    assert _relc(s["pu_cov"][1:], ref["sP"][1:, :2, :2]) < 1e-5


def test_rosenbrock_exp_ek_trajectory(oracle):
    """EK1(prior = IOUP(3, update_rate_parameter = true)) = RosenbrockExpEK (src/algorithms.jl:458-459, src/perform_step.jl:88-96): before every step
    the rate parameter becomes the Jacobian at the last filtered mean and the transition pair is rebuilt; the smoother uses each step's own pair."""
    mp.mp.dps = 120
    f, u0, p, q, h, N = problems.lotka_volterra, [1.0, 1.0], [1.5, 1.0, 3.0, 1.0], 3, "0.02", 40
    ref = dense_filter_smoother(f, u0, p, q, h, N, True, False, rate=np.zeros((2, 2)), rate_update=True)
    rhs = oracle.compile_rhs(tracer.trace(f, 2, 4), q)
    kw = dict(alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, calibrate=False, adaptive=False, dt=float(h), t0=0.0, t1=float(h) * N, save_everystep=True,
              prior=oracle.PRIOR_IOUP, rate_parameter=None, update_rate_parameter=True)
    _check(oracle.solve(oracle.make_config(2, q, 4, smooth=False, **kw), rhs, u0, p), ref, 2, None)
    s = oracle.solve(oracle.make_config(2, q, 4, smooth=True, **kw), rhs, u0, p)
    assert np.max(np.abs(s["pu_mu"] - ref["sm"][:, :2])) < 1e-10
    assert _relc(s["pu_cov"][1:], ref["sP"][1:, :2, :2]) < 1e-8
