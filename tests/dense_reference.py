"""Independent reference for the oracle and for the CUDA path: the textbook (covariance-form) Gaussian ODE filter and RTS smoother in plain
coordinates -- no square roots, no preconditioning, closed-form IWP transition matrices, IOUP pairs from the matrix exponential of the Van Loan
block matrix -- evaluated in extended precision (mpmath; the caller sets mp.mp.dps -- the covariance form loses digits with every step, so long runs need hundreds), exact initial derivatives by symbolic differentiation.  Shares no code and no
algorithmic choice with oracle/ or with the kernels.  Used by tests/test_oracle_dense_mp.py (CPU) and by tests/golden/make_mp_golden.py, whose
fixtures the GPU tests compare the CUDA path with."""
import mpmath as mp
import numpy as np
import sympy as sp


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
                          mv=False, dense_ts=None, tols=None, rate_update=False):
    """Returns numpy arrays: filtered / smoothed means [N+1, D] and covariances [N+1, D, D], local diffusions [N], the running MLE of the
    global diffusion, the log-likelihood, and the log-likelihood with the scalar (Kronecker-factor) log-determinant.

    rate_update   IOUP(q, update_rate_parameter=true) (RosenbrockExpEK): the rate matrix is re-set to the Jacobian before every step
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
    def transition(h, rate=rate):
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
        Rm = rate if isinstance(rate, mp.matrix) else _to_mp(rate)
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
        if rate_update:      # src/perform_step.jl:88-96: the rate parameter is the Jacobian at the last filtered mean, before every step
            A, Q = transition(hs[k] if hs is not None else h, rate=mp.matrix(J_l(list(E0 * m))))
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
