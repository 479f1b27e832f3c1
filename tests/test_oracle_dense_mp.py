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


def _to_mp(M):
    M = np.atleast_2d(np.asarray(M, dtype=object))
    return mp.matrix([[mp.mpf(str(x)) for x in row] for row in M])


def dense_filter_smoother(f, u0, p, q, h, nsteps, ek1, dynamic, rate=None, mass=None, second_order=False, obs_noise=None, diag_jac=False,
                          mv=False, dense_ts=None, tols=None):
    """Returns numpy arrays: filtered / smoothed means [N+1, D] and covariances [N+1, D, D], local diffusions [N], the running MLE of the
    global diffusion, the log-likelihood, and the log-likelihood with the scalar (Kronecker-factor) log-determinant.

    rate          IOUP prior with this d x d rate matrix (src/priors/ioup.jl:66-101): transition pair from the matrix exponential of the Van Loan
                  block matrix, in plain coordinates
    mass          nonsingular mass matrix M: z = M E1 x - f(E0 x), H = M E1 - J E0 (src/measurement_models.jl:11-20); initial state as the
                  reference builds it (see below)
    second_order  f is the first-order form F(du, u) = (f1(du, u), du) of u'' = f1(u', u): z = E2 x - f1(E1 x, E0 x),
                  H = E2 - J_du E1 - J_u E0 (src/measurement_models.jl:29-49); u0 = (du0, u0)
    obs_noise     pn_observation_noise R (d x d): S = H P^- H' + R (src/perform_step.jl:182-188)
    diag_jac      DiagonalEK1: H = E1 - diag(J) E0 (src/derivative_utils.jl:14-29); the components then decouple and the dense covariance is the
                  reference's BlocksOfDiagonals covariance exactly
    mv            multivariate diffusion: dynamic -> sigma^2_i = z_i^2 / (H Q H')_ii scales component i of Q (calibration.jl:151-180,
                  apply_diffusion.jl); fixed -> the running MLE z_i^2 / S[1,1] per component (calibration.jl:88-107), returned as `gd_mv`"""
    dim = len(u0)
    d = dim // 2 if second_order else dim
    n, D = q + 1, d * (q + 1)
    xs = list(sp.symbols(f"u0:{dim}"))
    dx = [None] * dim
    f(dx, xs, [sp.Rational(str(x)) for x in p], 0)
    fvec = sp.Matrix(dx)
    rat = lambda M: sp.Matrix(np.atleast_2d(np.asarray(M, dtype=float)).tolist()).applyfunc(lambda v: sp.Rational(str(v)))
    if second_order:                                     # state x = (du, u); positions are derivative 0, velocities derivative 1
        flow = fvec
        derivs = [sp.Matrix(xs[d:]), sp.Matrix(xs[:d]), sp.Matrix(fvec[:d])]
    else:
        flow = fvec
        derivs = [sp.Matrix(xs), flow]
    while len(derivs) < n:                               # autonomous problems: d/dt g(x) = (d g / d x) x'
        derivs.append(derivs[-1].jacobian(xs) * flow)
    if mass is not None:
        # src/initialization/autodiffinit.jl:16-48 as written: the Taylor coefficients are those of u' = f(u) -- the mass matrix is not
        # part of the Taylor recursion -- and the state is conditioned on M x^(o) = d^o u / dt^o for o >= 1.  For M != I the derivatives
        # of order >= 2 are therefore not those of the solution of M u' = f; the filter is started from this state all the same.
        Minv = rat(mass).inv()
        derivs = [derivs[0]] + [Minv * dk for dk in derivs[1:]]
    sub = dict(zip(xs, [sp.Rational(str(x)) for x in u0]))
    m = mp.matrix([mp.mpf(sp.N(e.subs(sub), 80)) for dk in derivs[:n] for e in dk])
    hs = None
    if isinstance(h, (list, tuple, np.ndarray)):      # variable steps: one transition pair per step (exact binary values of the doubles)
        hs = [mp.mpf(float(x)) for x in h]
        h = hs[0]
    else:
        h = mp.mpf(str(h))
    def transition(h):
        if rate is None:
            A1, Q1 = mp.matrix(n, n), mp.matrix(n, n)
            for i in range(n):
                for j in range(n):
                    A1[i, j] = h ** (j - i) / mp.factorial(j - i) if j >= i else 0
                    Q1[i, j] = h ** (2 * q + 1 - i - j) / ((2 * q + 1 - i - j) * mp.factorial(q - i) * mp.factorial(q - j))
            return _kron_identity(A1, d), _kron_identity(Q1, d)
        # dx = F x dt + L dW, F = shift (x) I + e_q e_q' (x) rate, L = e_q (x) I
        F = mp.zeros(D, D)
        for jj in range(q):
            for a in range(d):
                F[jj * d + a, (jj + 1) * d + a] = 1
        Rm = _to_mp(rate)
        for a in range(d):
            for b in range(d):
                F[q * d + a, q * d + b] = Rm[a, b]
        LLt = mp.zeros(D, D)
        for a in range(d):
            LLt[q * d + a, q * d + a] = 1
        blk = mp.zeros(2 * D, 2 * D)
        blk[:D, :D], blk[:D, D:], blk[D:, D:] = -F * h, LLt * h, F.T * h
        E = mp.expm(blk)
        At = E[D:, D:].T
        return At, At * E[:D, D:]

    A, Q = transition(h)
    As = []
    sel = lambda k: _kron_identity(mp.matrix([[1 if c == k else 0 for c in range(n)]]), d)
    E0, E1 = sel(0), sel(1)
    E2 = sel(2) if second_order else None
    f_l = sp.lambdify([xs], fvec, "mpmath")
    J_l = sp.lambdify([xs], fvec.jacobian(xs), "mpmath")
    Mm = _to_mp(mass) if mass is not None else None
    Rn = _to_mp(obs_noise) if obs_noise is not None else mp.zeros(d, d)
    P = mp.zeros(D, D)
    fm, fP, pm, pP, sig = [m.copy()], [P.copy()], [], [], []
    ll = ll_kron = gd = mp.mpf(0)
    gd_mv = [mp.mpf(0)] * d
    eest = []
    for k in range(nsteps):
        if hs is not None:
            A, Q = transition(hs[k])
        As.append(A)
        mpred = A * m
        if second_order:
            xin = list(E1 * mpred) + list(E0 * mpred)
            fx, Jx = mp.matrix(f_l(xin)), mp.matrix(J_l(xin))
            z = E2 * mpred - fx[:d, 0]
            H = E2 - Jx[:d, :d] * E1 - Jx[:d, d:] * E0 if ek1 else E2
        else:
            xin = list(E0 * mpred)
            lhs = Mm * E1 if Mm is not None else E1
            z = lhs * mpred - mp.matrix(f_l(xin))
            Jx = mp.matrix(J_l(xin))
            if diag_jac:
                Jx = mp.diag([Jx[a, a] for a in range(d)])
            H = lhs - Jx * E0 if ek1 else lhs
        W = H * Q * H.T
        if tols is not None:
            # local error estimate (src/perform_step.jl:199-247): e_i = h sqrt(sigma^2_loc W_ii), scaled by abstol + max(|u^-_i|, |uprev_i|) reltol, RMS over i
            s2l = (z.T * mp.lu_solve(W, z))[0] / d
            hk = hs[k] if hs is not None else h
            acc = mp.mpf(0)
            for a in range(d):
                sc = tols[0] + max(abs((E0 * mpred)[a]), abs((E0 * m)[a])) * tols[1]
                acc += (hk * mp.sqrt(s2l * W[a, a]) / sc) ** 2
            eest.append(mp.sqrt(acc / d))
        if mv and dynamic:
            s2 = [z[a] ** 2 / W[a, a] for a in range(d)]
            Qs = mp.zeros(D, D)
            for i in range(D):
                for j in range(D):
                    if i % d == j % d:
                        Qs[i, j] = s2[i % d] * Q[i, j]
            Pp = A * P * A.T + Qs
        else:
            s2 = (z.T * mp.lu_solve(W, z))[0] / d if dynamic else mp.mpf(1)
            Pp = A * P * A.T + s2 * Q
        S = H * Pp * H.T + Rn
        Sinv = S ** -1
        K = Pp * H.T * Sinv
        quad = (z.T * Sinv * z)[0]
        m = mpred - K * z
        P = Pp - K * S * K.T
        inc = quad / d
        gd = inc if k == 0 else gd + (inc - gd) / k      # the divisor is the number of steps accepted BEFORE this one
        inc_mv = [z[a] ** 2 / S[0, 0] for a in range(d)]
        gd_mv = inc_mv if k == 0 else [g + (i_ - g) / k for g, i_ in zip(gd_mv, inc_mv)]
        ll += -quad / 2 - mp.mpf(d) / 2 * mp.log(2 * mp.pi) - mp.log(mp.det(S)) / 2
        ll_kron += -quad / 2 - mp.mpf(d) / 2 * mp.log(2 * mp.pi) - mp.log(S[0, 0]) / 2
        fm.append(m.copy()), fP.append(P.copy()), pm.append(mpred), pP.append(Pp), sig.append(s2)
    sm, sP = [None] * (nsteps + 1), [None] * (nsteps + 1)
    sm[nsteps], sP[nsteps] = fm[nsteps], fP[nsteps]
    for k in range(nsteps - 1, -1, -1):
        G = fP[k] * As[k].T * pP[k] ** -1 if k > 0 else mp.zeros(D, D)      # P_0 = 0: G = 0, the initial state stays exact
        sm[k] = fm[k] + G * (sm[k + 1] - pm[k])
        sP[k] = fP[k] + G * (sP[k + 1] - pP[k]) * G.T
    vec = lambda x: np.array([float(v) for v in x])
    mat = lambda M: np.array([[float(M[i, j]) for j in range(M.cols)] for i in range(M.rows)])
    dense = {}
    if dense_ts is not None:
        # sol(t) (src/solution.jl:330-398): predict the filtering estimate of the last saved point over h1 with that interval's diffusion, then
        # one smoothing step against the next smoothed estimate over h2
        dm, dP = [], []
        for tv in dense_ts:
            tv = mp.mpf(str(tv))
            k = int(mp.floor(tv / h))
            h1, h2 = tv - k * h, (k + 1) * h - tv
            A1_, Q1_ = transition(h1)
            A2_, Q2_ = transition(h2)
            sk = sig[k]
            m_p, P_p = A1_ * fm[k], A1_ * fP[k] * A1_.T + sk * Q1_
            m_2, P_2 = A2_ * m_p, A2_ * P_p * A2_.T + sk * Q2_
            G = P_p * A2_.T * P_2 ** -1
            dm.append(m_p + G * (sm[k + 1] - m_2))
            dP.append(P_p + G * (sP[k + 1] - P_2) * G.T)
        dense = dict(dm=np.array([vec(x) for x in dm]), dP=np.array([mat(x) for x in dP]))
    dense["eest"] = np.array([float(x) for x in eest])
    return dict(**dense, fm=np.array([vec(x) for x in fm]), fP=np.array([mat(x) for x in fP]), sm=np.array([vec(x) for x in sm]),
                sP=np.array([mat(x) for x in sP]), sig=np.array([vec(x) if isinstance(x, list) else float(x) for x in sig]), gd=float(gd),
                gd_mv=vec(gd_mv), ll=float(ll), ll_kron=float(ll_kron))


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
    mp.mp.dps = 60
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
    mp.mp.dps = 60
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
    mp.mp.dps = 60
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
    mp.mp.dps = 60
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
    mp.mp.dps = 60
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
    mp.mp.dps = 60
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
    mp.mp.dps = 60
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
