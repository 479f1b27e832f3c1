"""The thread-per-trajectory filter kernel (pn_thread.cuh; lanes_per_traj = 1, and the default for EK1 / IWP filter-only solves with
D <= 12) against the oracle and against the group-per-trajectory kernel (pn_small.cuh, lanes_per_traj = 2)."""
import numpy as np
import pytest

import probnumdiffeq_jl_b200  # noqa: F401
from probnumdiffeq_jl_b200 import lib, problems, tracer

from test_gpu_parity import _gpu, _inputs, _rel, _relcov

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,q,t1,dt", [("fitzhugh_nagumo", 3, 5.0, 0.01), ("rober", 3, 0.02, 1e-4), ("vanderpol", 5, 0.05, 1e-4)])
def test_thread_kernel_fixed_steps_static_diffusion(oracle, name, q, t1, dt):
    """Fixed steps, FixedDiffusion: end state (mean, covariance, full state factor as R'R), log-likelihood and the calibrated
    diffusion at 1e-10 against BOTH oracle back-ends; the kernel follows the reference's Gram + Cholesky route."""
    pd, u0, p = _inputs(name, 4, 601, dp=0.05)
    g = _gpu(pd, u0, p, q, builtin=False, lanes_per_traj=1, adaptive=False, dt=dt, t0=0.0, t1=t1, diffusion=lib.DIFF_FIXED, calibrate=True,
             save_state=True)
    assert g["info"].lanes_per_traj == 1
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), q)
    for backend in (oracle.COV_REF, oracle.COV_QR):
        cfg = oracle.make_config(pd.d, q, pd.n_p, alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, calibrate=True, smooth=False, adaptive=False,
                                 dt=dt, t0=0.0, t1=t1, save_everystep=False, cov_backend=backend)
        for i in range(4):
            o = oracle.solve(cfg, rhs, u0[i], p[i])
            assert g["retcode"][i] == 0 and g["naccept"][i] == o["naccept"]
            assert _rel(g["u"][i], o["pu_mu"][-1]) < 1e-10
            assert _relcov(g["cov"][i], o["pu_cov"][-1]) < 1e-9
            assert g["loglik"][i] == pytest.approx(o["loglik"], rel=1e-9)
            assert g["diffusion"][i] == pytest.approx(o["diffusions"][0], rel=1e-9)
            So = o["x_filt_R"][-1].T @ o["x_filt_R"][-1]
            Sg = g["xR"][i].T @ g["xR"][i]
            assert np.linalg.norm(Sg - So) / np.linalg.norm(So) < 1e-8
            assert _rel(g["x"][i][:pd.d], o["x_filt_mu"][-1][:pd.d]) < 1e-10


def test_thread_kernel_adaptive_step_sequences(oracle):
    """Adaptive, DynamicDiffusion, non-stiff (Lotka-Volterra) and ROBER: accept / reject counts and nf identical to the oracle's per
    trajectory, end points at 1e-8; and identical counts / 1e-10 end points against the group-per-trajectory kernel."""
    for name, t1 in (("lotka_volterra", 10.0), ("rober", 100.0)):
        n = 64
        pd, u0, p = _inputs(name, n, 602, du0=0.1 if name == "lotka_volterra" else 0.0, dp=0.1)
        g1 = _gpu(pd, u0, p, 3, builtin=True, lanes_per_traj=1, adaptive=True, t0=0.0, t1=t1)
        g2 = _gpu(pd, u0, p, 3, builtin=True, lanes_per_traj=2, adaptive=True, t0=0.0, t1=t1)
        assert g1["info"].lanes_per_traj == 1 and g2["info"].lanes_per_traj == 2
        rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
        cfg = oracle.make_config(pd.d, 3, pd.n_p, alg=oracle.EK1, smooth=False, adaptive=True, t0=0.0, t1=t1, save_everystep=False,
                                 cov_backend=oracle.COV_REF)
        o = oracle.solve_ensemble(cfg, rhs, u0, p)
        assert np.all(g1["retcode"] == 0)
        assert np.array_equal(g1["naccept"], o["naccept"]) and np.array_equal(g1["nreject"], o["nreject"]) and np.array_equal(g1["nf"], o["nf"])
        assert _rel(g1["u"], o["u_final"]) < 1e-8
        assert np.array_equal(g1["naccept"], g2["naccept"]) and np.array_equal(g1["nreject"], g2["nreject"])
        assert _rel(g1["u"], g2["u"]) < 1e-10


def test_thread_kernel_zero_diffusion_takes_the_triangularize_route(oracle):
    """predict.jl:71-74 / :90: with zero diffusion (here forced on every third accepted step through the test hook) the Gram + Cholesky
    route does not apply; the kernel brings X = R Ah' to triangular form by Householder in local memory.  Every quantity still agrees
    with the oracle at 1e-10."""
    pd, u0, p = _inputs("lotka_volterra", 3, 603, du0=0.05)
    nst = 100
    forced = np.where(np.arange(nst) % 3 == 2, 0.0, 1.0 + 0.1 * np.arange(nst))
    kw = dict(adaptive=False, dt=0.02, t0=0.0, t1=2.0, forced_diffusion=forced)
    g = _gpu(pd, u0, p, 3, builtin=False, lanes_per_traj=1, **kw)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    cfg = oracle.make_config(2, 3, 4, alg=oracle.EK1, smooth=False, save_everystep=False, cov_backend=oracle.COV_REF, **kw)
    for i in range(3):
        o = oracle.solve(cfg, rhs, u0[i], p[i])
        assert g["naccept"][i] == o["naccept"] == nst
        assert _rel(g["u"][i], o["pu_mu"][-1]) < 1e-10
        assert _relcov(g["cov"][i], o["pu_cov"][-1]) < 1e-9
        assert g["loglik"][i] == pytest.approx(o["loglik"], rel=1e-9)


def test_thread_kernel_scope(oracle):
    """lanes_per_traj = 1 outside the kernel's scope is refused (no silent fallback to another mapping)."""
    pd = problems.PROBLEMS["lotka_volterra"]
    src = tracer.cuda_source(tracer.trace(pd.f, pd.d, pd.n_p), "UserProb")
    for kw in (dict(save_everystep=True), dict(alg=lib.EK0), dict(prior=lib.PRIOR_IOUP, rate_parameter=-1.0)):
        with pytest.raises(lib.PnodeError, match="thread per trajectory"):
            lib.Solver(lib.make_config(d=2, q=3, n_p=4, rhs_name="UserProb", rhs_src=src, lanes_per_traj=1, t0=0.0, t1=1.0, **kw))
