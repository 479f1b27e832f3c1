"""The oracle against an independent implementation that shares nothing with it: the textbook (covariance-form) Gaussian ODE filter and
RTS smoother in plain coordinates -- no square roots, no preconditioning, closed-form IWP transition matrices -- evaluated with 60
significant digits (mpmath), exact initial derivatives by symbolic differentiation.  In double precision that formulation loses the
solution after a few dozen steps (the process-noise covariance has entries h^(2q+1)); with 60 digits it is the exact answer the
square-root algorithm of the reference (src/filtering/predict.jl, update.jl, markov_kernel.jl) approximates.

Pins, over whole fixed-step trajectories: filtered and smoothed means and covariances, the local diffusions, the running global
diffusion MLE with the reference's divisor (src/diffusions/calibration.jl:56-62) and its a-posteriori scaling of the covariances
(src/integrator_utils.jl:46-65), the log-likelihood (src/filtering/update.jl:140-148) and, for the EK0, the reference's Kronecker
log-determinant as written (update.jl:147: log det of the 1 x 1 factor, not multiplied by d)."""
import mpmath as mp
import numpy as np
import pytest
import sympy as sp

import probnumdiffeq_jl_b200  # noqa: F401
from probnumdiffeq_jl_b200 import problems, tracer


def _kron_identity(B, d):
    M = mp.zeros(B.rows * d, B.cols * d)
    for i in range(B.rows):
        for j in range(B.cols):
            for a in range(d):
                M[i * d + a, j * d + a] = B[i, j]
    return M


def dense_filter_smoother(f, u0, p, q, h, nsteps, ek1, dynamic):
    """Returns numpy arrays: filtered / smoothed means [N+1, D] and covariances [N+1, D, D], local diffusions [N], the running MLE of the
    global diffusion, the log-likelihood, and the log-likelihood with the scalar (Kronecker-factor) log-determinant."""
    d = len(u0)
    n, D = q + 1, d * (q + 1)
    us = list(sp.symbols(f"u0:{d}"))
    du = [None] * d
    f(du, us, [sp.Rational(str(x)) for x in p], 0)
    fvec = sp.Matrix(du)
    derivs = [sp.Matrix(us), fvec]                       # autonomous problems: u^(k+1) = (d u^(k) / d u) f
    for _ in range(2, n):
        derivs.append(derivs[-1].jacobian(us) * fvec)
    sub = dict(zip(us, [sp.Rational(str(x)) for x in u0]))
    m = mp.matrix([mp.mpf(sp.N(e.subs(sub), 80)) for dk in derivs for e in dk])
    h = mp.mpf(str(h))
    A1, Q1 = mp.matrix(n, n), mp.matrix(n, n)
    for i in range(n):
        for j in range(n):
            A1[i, j] = h ** (j - i) / mp.factorial(j - i) if j >= i else 0
            Q1[i, j] = h ** (2 * q + 1 - i - j) / ((2 * q + 1 - i - j) * mp.factorial(q - i) * mp.factorial(q - j))
    A, Q = _kron_identity(A1, d), _kron_identity(Q1, d)
    e0, e1 = mp.zeros(1, n), mp.zeros(1, n)
    e0[0, 0] = e1[0, 1] = 1
    E0, E1 = _kron_identity(e0, d), _kron_identity(e1, d)
    f_l = sp.lambdify([us], fvec, "mpmath")
    J_l = sp.lambdify([us], fvec.jacobian(us), "mpmath")
    P = mp.zeros(D, D)
    fm, fP, pm, pP, sig = [m.copy()], [P.copy()], [], [], []
    ll = ll_kron = gd = mp.mpf(0)
    for k in range(nsteps):
        mpred = A * m
        up = [(E0 * mpred)[i] for i in range(d)]
        z = E1 * mpred - mp.matrix(f_l(up))
        H = E1 - mp.matrix(J_l(up)) * E0 if ek1 else E1
        W = H * Q * H.T
        s2 = (z.T * mp.lu_solve(W, z))[0] / d if dynamic else mp.mpf(1)
        Pp = A * P * A.T + s2 * Q
        S = H * Pp * H.T
        Sinv = S ** -1
        K = Pp * H.T * Sinv
        quad = (z.T * Sinv * z)[0]
        m = mpred - K * z
        P = Pp - K * S * K.T
        inc = quad / d
        gd = inc if k == 0 else gd + (inc - gd) / k      # the divisor is the number of steps accepted BEFORE this one
        ll += -quad / 2 - mp.mpf(d) / 2 * mp.log(2 * mp.pi) - mp.log(mp.det(S)) / 2
        ll_kron += -quad / 2 - mp.mpf(d) / 2 * mp.log(2 * mp.pi) - mp.log(S[0, 0]) / 2
        fm.append(m.copy()), fP.append(P.copy()), pm.append(mpred), pP.append(Pp), sig.append(s2)
    sm, sP = [None] * (nsteps + 1), [None] * (nsteps + 1)
    sm[nsteps], sP[nsteps] = fm[nsteps], fP[nsteps]
    for k in range(nsteps - 1, -1, -1):
        G = fP[k] * A.T * pP[k] ** -1 if k > 0 else mp.zeros(D, D)      # P_0 = 0: G = 0, the initial state stays exact
        sm[k] = fm[k] + G * (sm[k + 1] - pm[k])
        sP[k] = fP[k] + G * (sP[k + 1] - pP[k]) * G.T
    vec = lambda x: np.array([float(v) for v in x])
    mat = lambda M: np.array([[float(M[i, j]) for j in range(M.cols)] for i in range(M.rows)])
    return dict(fm=np.array([vec(x) for x in fm]), fP=np.array([mat(x) for x in fP]), sm=np.array([vec(x) for x in sm]),
                sP=np.array([mat(x) for x in sP]), sig=vec(sig), gd=float(gd), ll=float(ll), ll_kron=float(ll_kron))


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
    mp.mp.dps = 60
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
    mp.mp.dps = 60
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
