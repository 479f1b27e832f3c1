"""IOUP priors with a vector / matrix rate parameter and the Jacobian-updated rate (RosenbrockExpEK) -- src/priors/ioup.jl:81-131,
src/perform_step.jl:88-96 -- on the general-prior path (pn_gen.cuh filter, pn_smooth.cuh sweep, pn_lti.cuh transition pairs by
scaling and doubling) against the oracle, whose transition matrices come from an independent route (Pade matrix exponential of the
Van Loan block matrix, pinned against scipy quadrature and against the thresholds of test/ioup.jl and test/correctness.jl:53,84 in
tests/test_oracle_pins.py)."""
import numpy as np
import pytest

import probnumdiffeq_jl_b200  # noqa: F401
from probnumdiffeq_jl_b200 import api, lib, problems, tracer

import floor as rounding
from test_gpu_parity import _gpu, _gpu_smooth, _inputs, _rel, _relcov

pytestmark = pytest.mark.gpu

RATE = np.array([[-1.0, 0.3], [-0.2, -0.5]])


def _osc(du, u, p, t):
    du[0] = -p[0] * u[0] - 2 * np.pi * u[1]
    du[1] = 2 * np.pi * u[0] - p[0] * u[1]


@pytest.mark.parametrize("alg,rate", [(lib.EK1, RATE), (lib.EK0, RATE), (lib.EK1, np.array([-1.0, -0.4]))])
def test_matrix_and_vector_rate_fixed_steps_static_diffusion(oracle, alg, rate):
    """Fixed steps, FixedDiffusion(calibrate=false): every filtered and every smoothed point at 1e-10 (means, covariances), the
    log-likelihood at 1e-9; EK1 and EK0 (ExpEK), matrix and vector rates."""
    pd, u0, p = _inputs("lotka_volterra", 3, 501, du0=0.1, dp=0.1)
    kw = dict(adaptive=False, dt=0.02, t0=0.0, t1=2.0, calibrate=False, diffusion=lib.DIFF_FIXED, prior=lib.PRIOR_IOUP, rate_parameter=rate)
    g = _gpu(pd, u0, p, 3, builtin=False, alg=alg, save_everystep=True, **kw)
    gs = _gpu_smooth(pd, u0, p, 3, alg=alg, **kw)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    okw = dict(alg=alg, diffusion=oracle.DIFF_FIXED, calibrate=False, adaptive=False, dt=0.02, t0=0.0, t1=2.0, prior=oracle.PRIOR_IOUP,
               rate_parameter=rate, cov_backend=oracle.COV_REF)
    for i in range(3):
        o = oracle.solve(oracle.make_config(2, 3, 4, smooth=True, **okw), rhs, u0[i], p[i])
        of = oracle.solve(oracle.make_config(2, 3, 4, smooth=False, **okw), rhs, u0[i], p[i])
        a, b = g["off"][i], g["off"][i + 1]
        assert b - a == o["n"] == 101 and g["retcode"][i] == 0
        assert _rel(g["U"][a:b], of["pu_mu"]) < 1e-10
        assert _relcov(g["C"][a + 1:b], of["pu_cov"][1:]) < 1e-10
        assert g["loglik"][i] == pytest.approx(of["loglik"], rel=1e-9)
        a, b = gs["off"][i], gs["off"][i + 1]
        assert _rel(gs["U"][a:b], o["pu_mu"]) < 1e-10
        assert _relcov(gs["C"][a + 1:b], o["pu_cov"][1:]) < 1e-10
        Sg = np.einsum("kra,krb->kab", gs["XR"][a:b], gs["XR"][a:b])
        So = np.einsum("kra,krb->kab", o["x_smooth_R"], o["x_smooth_R"])
        assert _relcov(Sg[1:], So[1:]) < 1e-9


def test_isotropic_rate_matches_the_scalar_rate_kernels(oracle):
    """IOUP(3, [r, r]) and IOUP(3, r I) are the SDE of IOUP(3, r) (test/ioup.jl:48-55): the general-prior path (scaling and doubling of
    the D x D pair) against the register-resident scalar-rate kernels (phi-functions + Gauss-Legendre, pn_ioup.cuh), same accept /
    reject sequence, end points at 1e-9."""
    n = 16
    pd, u0, p = _inputs("lotka_volterra", n, 502, du0=0.1, dp=0.1)
    kw = dict(builtin=False, adaptive=True, t0=0.0, t1=5.0, prior=lib.PRIOR_IOUP)
    gs = _gpu(pd, u0, p, 3, rate_parameter=-0.7, **kw)
    for r in ([-0.7, -0.7], -0.7 * np.eye(2)):
        gv = _gpu(pd, u0, p, 3, rate_parameter=r, **kw)
        assert np.all(gv["retcode"] == 0)
        assert np.array_equal(gv["naccept"], gs["naccept"]) and np.array_equal(gv["nreject"], gs["nreject"])
        assert _rel(gv["u"], gs["u"]) < 1e-9
        assert _relcov(gv["cov"], gs["cov"]) < 1e-6


@pytest.mark.parametrize("diffusion", [lib.DIFF_FIXED, lib.DIFF_DYNAMIC])
def test_rate_update_adaptive_matches_oracle_step_sequences(oracle, diffusion):
    """RosenbrockExpEK = EK1(prior=IOUP(3, update_rate_parameter=true)), adaptive: identical accept / reject counts and nf per trajectory.
    End points: 1e-8 with FixedDiffusion.  With DynamicDiffusion this prior makes the residual z -- and with it sigma^2 = z'(HQH')^-1 z,
    which steers the step sizes -- almost pure rounding noise (the Jacobian-updated prior nearly solves the ODE), so the bound is
    max(1e-8, 20 x the oracle's own floor) with the floor measured per trajectory against its rounding-level variants (tests/floor.py)."""
    n = 16
    pd, u0, p = _inputs("lotka_volterra", n, 503, du0=0.1, dp=0.1)
    g = _gpu(pd, u0, p, 3, builtin=False, adaptive=True, t0=0.0, t1=5.0, prior=lib.PRIOR_IOUP, rate_parameter=None,
             update_rate_parameter=True, diffusion=diffusion)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    kwo = dict(alg=oracle.EK1, smooth=False, adaptive=True, t0=0.0, t1=5.0, save_everystep=False, diffusion=diffusion,
               cov_backend=oracle.COV_REF, prior=oracle.PRIOR_IOUP, rate_parameter=None, update_rate_parameter=True)
    o = oracle.solve_ensemble(oracle.make_config(2, 3, 4, **kwo), rhs, u0, p)
    assert np.all(g["retcode"] == 0)
    assert np.array_equal(g["naccept"], o["naccept"]) and np.array_equal(g["nreject"], o["nreject"])
    assert np.array_equal(g["nf"], o["nf"])
    if diffusion == lib.DIFF_FIXED:
        assert _rel(g["u"], o["u_final"]) < 1e-8
        assert _relcov(g["cov"], o["cov_final"]) < 1e-6
        assert _rel(g["loglik"], o["loglik"]) < 1e-6
        return
    worst = 0.0
    for i in range(n):
        oi = oracle.solve(oracle.make_config(2, 3, 4, **kwo), rhs, u0[i], p[i])
        e = _rel(g["u"][i], oi["pu_mu"][-1])
        fl = rounding.floor(oracle, rhs, 2, 3, 4, u0[i], p[i], kwo, oi, lambda s_, b_: _rel(s_["pu_mu"][-1], b_["pu_mu"][-1]))
        assert e < max(1e-8, 20 * fl), (i, e, fl)
        worst = max(worst, e)
    assert worst < 1e-3   # the solver tolerance is reltol = 1e-3 (measured oracle-vs-oracle floors: 2e-11 ... 3.5e-5)


def test_rate_update_smoother_fixed_steps(oracle):
    """Smoothing with the Jacobian-updated rate: the backward kernel of a step uses the rate matrix the step was taken with (saved per
    step by the filter).  Fixed steps, static diffusion: smoothed means / covariances at 1e-10."""
    pd, u0, p = _inputs("fitzhugh_nagumo", 2, 504, dp=0.05)
    kw = dict(adaptive=False, dt=0.05, t0=0.0, t1=5.0, diffusion=lib.DIFF_FIXED, calibrate=True, prior=lib.PRIOR_IOUP, rate_parameter=None,
              update_rate_parameter=True)
    gs = _gpu_smooth(pd, u0, p, 3, **kw)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    for i in range(2):
        o = oracle.solve(oracle.make_config(2, 3, 3, alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, calibrate=True, smooth=True, adaptive=False,
                                            dt=0.05, t0=0.0, t1=5.0, cov_backend=oracle.COV_REF, prior=oracle.PRIOR_IOUP, rate_parameter=None,
                                            update_rate_parameter=True), rhs, u0[i], p[i])
        a, b = gs["off"][i], gs["off"][i + 1]
        assert b - a == o["n"] == 101
        assert _rel(gs["U"][a:b], o["pu_mu"]) < 1e-10
        assert _relcov(gs["C"][a + 1:b], o["pu_cov"][1:]) < 1e-10


def test_exact_rate_matrix_on_the_linear_oscillator_and_api(oracle):
    """test/ioup.jl through the Python mirror of the reference API: on u' = A u the prior IOUP(3, A) is exact up to rounding
    (|error| < 5e-10 with a handful of steps, :28-31), a perturbed A still beats the IWP's nf (:22-26), and RosenbrockExpEK() finds A by
    itself; ExpEK(L = A) is the EK0 with that prior (src/algorithms.jl:405-425)."""
    a = 0.1
    A = np.array([[-a, -2 * np.pi], [2 * np.pi, -a]])
    w, V = np.linalg.eig(A)
    exact = np.real(V @ np.diag(np.exp(2.0 * w)) @ np.linalg.solve(V, np.array([1.0, 1.0])))
    prob = api.ODEProblem(_osc, [1.0, 1.0], (0.0, 2.0), [a])
    sol_iwp = api.solve(prob, api.EK1(smooth=False))
    sol_exact = api.solve(prob, api.EK1(prior=api.IOUP(3, A), smooth=False))
    sol_noisy = api.solve(prob, api.EK1(prior=api.IOUP(3, A + 1e-3 * np.random.default_rng(42).standard_normal((2, 2))), smooth=False))
    sol_ros = api.solve(prob, api.RosenbrockExpEK(smooth=False))
    sol_exp = api.solve(prob, api.ExpEK(L=A, smooth=True))
    err = lambda s: np.linalg.norm(np.asarray(s.u[-1]) - exact)
    assert err(sol_iwp) < 1e-5 and err(sol_noisy) < 2e-5 and err(sol_exact) < 5e-10 and err(sol_ros) < 5e-10
    assert sol_exact.stats.nf < sol_noisy.stats.nf < sol_iwp.stats.nf
    assert err(sol_exp) < 1e-3 and sol_exp.retcode == "Success"
    with pytest.raises(ValueError, match="Do not manually set the `rate_parameter`"):
        api.IOUP(3, A, update_rate_parameter=True)


def test_exponential_integrators_scalar_problem(oracle):
    """test/exponential_integrators.jl:6-41 through the Python mirror of the reference API, on the device: u' = p u (d = 1, the smallest
    state: D = 4), p = -1, t in [0, 10], default tolerances, adaptive steps, smoothing on.  EK0 runs on the Kronecker layout, EK1 on the
    dense one, ExpEK(L = p) on the scalar-rate IOUP kernels, RosenbrockExpEK on the general-prior path.  The reference's own assertions
    (end-point errors 1e-7 / 1e-9 / 2e-10 / 2e-12, their ordering, fewer steps and f evaluations for the exponential integrators) and
    parity with the oracle (same accept / reject counts, same nf, end point 1e-8 of the solution scale)."""

    def lin(du, u, p, t):
        du[0] = p[0] * u[0]

    prob = api.ODEProblem(lin, [1.0], (0.0, 10.0), [-1.0])
    rhs = oracle.compile_rhs(tracer.trace(lin, 1, 1), 3)
    uend = np.exp(-10.0)
    res = {}
    for name, alg, okw in (("ek0", api.EK0(order=3), dict(alg=oracle.EK0)), ("ek1", api.EK1(order=3), dict(alg=oracle.EK1)),
                           ("exp", api.ExpEK(L=-1.0, order=3), dict(alg=oracle.EK0, prior=oracle.PRIOR_IOUP, rate_parameter=-1.0)),
                           ("ros", api.RosenbrockExpEK(order=3),
                            dict(alg=oracle.EK1, prior=oracle.PRIOR_IOUP, rate_parameter=None, update_rate_parameter=True))):
        sol = api.solve(prob, alg)
        o = oracle.solve(oracle.make_config(1, 3, 1, smooth=True, adaptive=True, t0=0, t1=10, **okw), rhs, [1.0], [-1.0])
        assert sol.retcode == "Success" and o["retcode"] == "Success"
        assert (sol.stats.naccept, sol.stats.nreject, sol.stats.nf) == (o["naccept"], o["nreject"], o["nf"]), name
        assert len(sol.t) == o["n"] and sol.t[-1] == 10.0
        # smoothed trajectory against the oracle on the scale of the solution (u decays from 1 to 4.5e-5)
        assert np.max(np.abs(np.asarray(sol.u)[:, 0] - o["pu_mu"][:, 0])) < 1e-8, name
        res[name] = (abs(sol.u[-1][0] - uend), len(sol.t), sol.stats.nf)
    (e0, n0, nf0), (e1, n1, nf1), (ee, ne, nfe), (er, nr, nfr) = res["ek0"], res["ek1"], res["exp"], res["ros"]
    assert e0 < 1e-7 and e1 < 1e-9 and ee < 2e-10 and er < 2e-12
    assert er < ee < e1 < e0
    assert ne < n0 and nfe < nf0 and ne < n1 and nfe < nf1 and nr < n0 and nfr < nf0 and nr < n1 and nfr < nf1


def test_large_state_through_the_global_workspace(oracle):
    """D = d(q+1) = 28 * 2 = 56 (Pleiades, order 1): the per-warp workspace (10 D^2 doubles and change, 0.3 MB) does not fit in shared
    memory and the filter runs out of a global buffer; diagonal rate, fixed steps, static diffusion, every point against the oracle."""
    pd, u0, p = _inputs("pleiades", 2, 505, du0=1e-3)
    rate = -0.5 + 0.01 * np.arange(28)
    kw = dict(adaptive=False, dt=0.01, t0=0.0, t1=0.1, diffusion=lib.DIFF_FIXED, calibrate=False, prior=lib.PRIOR_IOUP, rate_parameter=rate)
    g = _gpu(pd, u0, p, 1, builtin=False, save_everystep=True, **kw)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 1)
    for i in range(2):
        o = oracle.solve(oracle.make_config(28, 1, 0, alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, calibrate=False, smooth=False, adaptive=False,
                                            dt=0.01, t0=0.0, t1=0.1, cov_backend=oracle.COV_REF, prior=oracle.PRIOR_IOUP, rate_parameter=rate),
                         rhs, u0[i], p[i])
        a, b = g["off"][i], g["off"][i + 1]
        assert b - a == o["n"] == 11 and g["retcode"][i] == 0
        assert _rel(g["U"][a:b], o["pu_mu"]) < 1e-10
        assert _relcov(g["C"][a + 1:b], o["pu_cov"][1:]) < 1e-9


def test_scalar_rate_beyond_the_register_kernels(oracle):
    """IOUP(1, r) on Pleiades (D = 56 > 24, the limit of the register-resident scalar-rate kernels): the library hands the same SDE to
    the general-prior path as IOUP(1, r I) (pnode_api.cu::promote_scalar_rate; test/ioup.jl:48-55).  Against the oracle's scalar-rate
    discretisation, fixed steps, static diffusion, every point; and bit-identical to the solve that spells the rate as r I."""
    pd, u0, p = _inputs("pleiades", 2, 506, du0=1e-3)
    kw = dict(adaptive=False, dt=0.01, t0=0.0, t1=0.1, diffusion=lib.DIFF_FIXED, calibrate=False, prior=lib.PRIOR_IOUP)
    g = _gpu(pd, u0, p, 1, builtin=False, save_everystep=True, rate_parameter=-0.5, **kw)
    gm = _gpu(pd, u0, p, 1, builtin=False, save_everystep=True, rate_parameter=-0.5 * np.eye(28), **kw)
    assert np.array_equal(g["U"], gm["U"]) and np.array_equal(g["C"], gm["C"])
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 1)
    for i in range(2):
        o = oracle.solve(oracle.make_config(28, 1, 0, alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, calibrate=False, smooth=False, adaptive=False,
                                            dt=0.01, t0=0.0, t1=0.1, cov_backend=oracle.COV_REF, prior=oracle.PRIOR_IOUP, rate_parameter=-0.5),
                         rhs, u0[i], p[i])
        a, b = g["off"][i], g["off"][i + 1]
        assert b - a == o["n"] == 11 and g["retcode"][i] == 0
        assert _rel(g["U"][a:b], o["pu_mu"]) < 1e-10
        assert _relcov(g["C"][a + 1:b], o["pu_cov"][1:]) < 1e-9
