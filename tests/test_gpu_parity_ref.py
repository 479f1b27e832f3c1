"""GPU parity against the oracle's REFERENCE-ROUTE covariance back-end (`COV_REF`: Gram matrix + Cholesky with the QR
fallback -- what `predict_cov!` itself does, /root/reference src/filtering/predict.jl:84-91), not only against the
oracle's all-QR back-end that tests/test_gpu_parity.py uses.

The two oracle back-ends agree to <= 7e-13 under static diffusion on every configuration below (measured on the CPU;
SURVEY.md H0), so the 1e-10 bar of BASELINE.json's north_star is asserted against the reference's own route:
means, covariances (as R'R -- factors are unique only up to an orthogonal factor), log-likelihood, diffusion MLE.

Second half: the cfg3 smoother instantiations `pn_smooth_col_kernel<2,5,8>` / `pn_smooth_reg_kernel<2,5,8>` (Van der Pol,
order 5) against the ORACLE (round 1 only cross-checked them against the shared-memory sweep on the same device).
"""
import numpy as np
import pytest

import probnumdiffeq_jl_b200  # noqa: F401
from probnumdiffeq_jl_b200 import lib, problems, tracer

import floor as rounding
from test_gpu_parity import _gpu, _gpu_ek0, _gpu_smooth, _inputs, _rel, _relcov

pytestmark = pytest.mark.gpu


def _state_cov(R):
    return np.einsum("...ra,...rb->...ab", R, R)


# ------------------------------------------------------------------------------------------------------------------
# COV_REF legs: fixed steps, FixedDiffusion -> 1e-10 on everything
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("calibrate", [True, False])
def test_cfg1_fhn_fixed_steps_vs_reference_route(oracle, calibrate):
    """BASELINE cfg1 (FitzHugh-Nagumo EK1(3), dt = 0.01, 2000 steps): filter AND smoother against the Gram + Cholesky oracle."""
    pd, u0, p = _inputs("fitzhugh_nagumo", 2, 101, dp=0.05)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    kw = dict(adaptive=False, dt=0.01, t0=0.0, t1=20.0, diffusion=lib.DIFF_FIXED, calibrate=calibrate)
    g = _gpu(pd, u0, p, 3, builtin=False, save_everystep=True, save_state=True, **kw)
    gs = _gpu_smooth(pd, u0, p, 3, **kw)
    for i in range(u0.shape[0]):
        ko = dict(alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, calibrate=calibrate, adaptive=False, dt=0.01, t0=0.0, t1=20.0,
                  save_everystep=True, cov_backend=oracle.COV_REF)
        o = oracle.solve(oracle.make_config(2, 3, 3, smooth=False, **ko), rhs, u0[i], p[i])
        a, b = g["off"][i], g["off"][i + 1]
        assert b - a == o["n"] == 2001 and np.array_equal(g["T"][a:b], o["t"])
        assert _rel(g["U"][a:b], o["pu_mu"]) < 1e-10
        assert _relcov(g["C"][a + 1:b], o["pu_cov"][1:]) < 1e-10
        assert g["loglik"][i] == pytest.approx(o["loglik"], rel=1e-10)
        assert g["diffusion"][i] == pytest.approx(o["diffusions"][0], rel=1e-10)
        assert _relcov(_state_cov(g["xR"][i]), _state_cov(o["x_filt_R"][-1])) < 1e-10
        os_ = oracle.solve(oracle.make_config(2, 3, 3, smooth=True, **ko), rhs, u0[i], p[i])
        a, b = gs["off"][i], gs["off"][i + 1]
        assert _rel(gs["U"][a:b], os_["pu_mu"]) < 1e-10
        assert _relcov(gs["C"][a + 1:b], os_["pu_cov"][1:]) < 1e-10
        assert _relcov(_state_cov(gs["XR"][a + 1:b]), _state_cov(os_["x_smooth_R"][1:])) < 1e-9


def test_cfg2_lv_ek0_fixed_steps_vs_reference_route(oracle):
    """BASELINE cfg2 shape (Lotka-Volterra EK0(3), IsometricKronecker layout), fixed steps, FixedDiffusion."""
    pd, u0, p = _inputs("lotka_volterra", 3, 102, du0=0.1)
    g = _gpu_ek0(pd, u0, p, 3, adaptive=False, dt=0.01, t0=0.0, t1=10.0, diffusion=lib.DIFF_FIXED, calibrate=True,
                 save_everystep=True, save_state=True)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    for i in range(3):
        o = oracle.solve(oracle.make_config(2, 3, 4, alg=oracle.EK0, diffusion=oracle.DIFF_FIXED, calibrate=True, smooth=False,
                                            adaptive=False, dt=0.01, t0=0.0, t1=10.0, save_everystep=True,
                                            cov_backend=oracle.COV_REF), rhs, u0[i], p[i])
        a, b = g["off"][i], g["off"][i + 1]
        assert b - a == o["n"] == 1001
        assert _rel(g["U"][a:b], o["pu_mu"]) < 1e-10
        assert _relcov(g["C"][a + 1:b], o["pu_cov"][1:]) < 1e-10
        assert g["loglik"][i] == pytest.approx(o["loglik"], rel=1e-10)
        assert g["diffusion"][i] == pytest.approx(o["diffusions"][0], rel=1e-10)
        assert _relcov(_state_cov(g["xR"][i]), _state_cov(o["x_filt_R"][-1])) < 1e-10


def test_cfg5_rober_fixed_steps_vs_reference_route(oracle):
    """BASELINE cfg5 problem (ROBER EK1(3), the headline kernel instantiation <PnRober,3,4>), dt = 1e-3 on [0, 1]."""
    pd, u0, p = _inputs("rober", 3, 103)
    g = _gpu(pd, u0, p, 3, builtin=True, adaptive=False, dt=1e-3, t0=0.0, t1=1.0, diffusion=lib.DIFF_FIXED, calibrate=True,
             save_everystep=True, save_state=True)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    for i in range(3):
        o = oracle.solve(oracle.make_config(3, 3, 3, alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, calibrate=True, smooth=False,
                                            adaptive=False, dt=1e-3, t0=0.0, t1=1.0, save_everystep=True,
                                            cov_backend=oracle.COV_REF), rhs, u0[i], p[i])
        a, b = g["off"][i], g["off"][i + 1]
        assert b - a == o["n"] == 1001 and g["retcode"][i] == 0
        assert _rel(g["U"][a:b], o["pu_mu"]) < 1e-10
        assert _relcov(g["C"][a + 1:b], o["pu_cov"][1:]) < 1e-10
        assert g["loglik"][i] == pytest.approx(o["loglik"], rel=1e-10)
        assert g["diffusion"][i] == pytest.approx(o["diffusions"][0], rel=1e-10)
        assert _relcov(_state_cov(g["xR"][i]), _state_cov(o["x_filt_R"][-1])) < 1e-10


def test_cfg4_pleiades_fixed_steps_vs_reference_route(oracle):
    """BASELINE cfg4 (Pleiades, D = 112, CTA kernel): 50 fixed steps, FixedDiffusion, whole 112 x 112 state covariance."""
    pd, u0, p = _inputs("pleiades", 1, 104)
    p = np.zeros((1, 0))
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    g = _gpu(pd, u0, p, 3, builtin=False, adaptive=False, dt=0.01, t0=0.0, t1=0.5, diffusion=lib.DIFF_FIXED, save_state=True)
    o = oracle.solve(oracle.make_config(28, 3, 0, alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, smooth=False, adaptive=False, dt=0.01,
                                        t0=0.0, t1=0.5, save_everystep=False, cov_backend=oracle.COV_REF), rhs, u0[0], [])
    assert _rel(g["u"][0], o["pu_mu"][-1]) < 1e-10
    assert _relcov(g["cov"][0], o["pu_cov"][-1]) < 1e-10
    assert _relcov(_state_cov(g["xR"][0]), _state_cov(o["x_filt_R"][-1])) < 1e-10
    assert g["loglik"][0] == pytest.approx(o["loglik"], rel=1e-10)
    assert g["diffusion"][0] == pytest.approx(o["diffusions"][0], rel=1e-10)


# ------------------------------------------------------------------------------------------------------------------
# cfg3: Van der Pol, EK1(order=5), D = 12 -- the <2,5,8> smoother instantiations against the oracle
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["col", "reg"])
@pytest.mark.parametrize("t1", [0.5, 2.0])
def test_cfg3_vdp_q5_smoother_vs_oracle(oracle, monkeypatch, kind, t1):
    """Step-teacher-forced (the kernel replays the oracle's accepted grid through tstops, as the stiff filter test does):
    (i) sigma^2 forced as well -> pure linear algebra of the backward sweep, means / covariances within
        max(1e-10, 10 x floor) of BOTH oracle back-ends, the floor being the oracle against one-ulp-perturbed / FMA /
        other-back-end copies of itself, measured here;
    (ii) DynamicDiffusion free sigma^2 -> the same with 20 x floor.
    t1 = 2.0 is the full BASELINE cfg3 interval (583 saved points through the relaxation jump), t1 = 0.5 a short one
    whose floor is small enough for the 1e-10 bar itself to bind on the means."""
    if kind == "reg":
        monkeypatch.setenv("PNODE_SMOOTH_KIND", "0")
    pd, u0, p = _inputs("vanderpol", 1, 105, dp=0.0)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 5)
    o = oracle.solve(oracle.make_config(2, 5, 1, alg=oracle.EK1, smooth=True, adaptive=True, t0=0.0, t1=t1, save_everystep=True,
                                        cov_backend=oracle.COV_QR), rhs, u0[0], p[0])
    grid, sig = o["t"][1:-1], o["diffusions"][:-1]
    base = dict(alg=oracle.EK1, smooth=True, adaptive=False, dt=t1, t0=0.0, t1=t1, tstops=grid, save_everystep=True)

    def m_mu(s_, b_):
        return _rel(s_["pu_mu"], b_["pu_mu"])

    def m_cov(s_, b_):
        return _relcov(s_["pu_cov"][1:], b_["pu_cov"][1:])

    for forced, k in ((True, 10.0), (False, 20.0)):
        kw = dict(base, forced_diffusion=sig) if forced else dict(base)
        g = _gpu_smooth(pd, u0, p, 5, adaptive=False, dt=t1, t0=0.0, t1=t1, tstops=grid, **({"forced_diffusion": sig} if forced else {}))
        assert g["info"].lanes_per_traj == 2                       # filter <.,5,2>; the sweep runs 8 lanes per trajectory
        for backend in (oracle.COV_QR, oracle.COV_REF):
            kwb = dict(kw, cov_backend=backend)
            ob = oracle.solve(oracle.make_config(2, 5, 1, **kwb), rhs, u0[0], p[0])
            assert np.array_equal(g["T"], ob["t"]) and np.array_equal(ob["t"], o["t"])
            fm = rounding.floor(oracle, rhs, 2, 5, 1, u0[0], p[0], kwb, ob, m_mu)
            fc = rounding.floor(oracle, rhs, 2, 5, 1, u0[0], p[0], kwb, ob, m_cov)
            em, ec = _rel(g["U"], ob["pu_mu"]), _relcov(g["C"][1:], ob["pu_cov"][1:])
            assert em < max(1e-10, k * fm), (kind, t1, forced, backend, em, fm)
            assert ec < max(1e-10, k * fc), (kind, t1, forced, backend, ec, fc)


def test_cfg3_vdp_q5_smoother_adaptive_ensemble_vs_oracle(oracle):
    """Free-running adaptive cfg3 ensemble (ragged lengths, idle groups in the sweep's warps).  The stiff Van der Pol step
    pattern is rounding-sensitive: the oracle itself changes its number of steps under rounding-level
    perturbations (tests/floor.py: the other covariance back-end, fused multiply-adds, ONE ulp on one initial derivative) on
    most of these trajectories.  So: every trajectory must end within 1e-5 of the oracle and take a number of steps inside
    the oracle's own envelope (+-1); trajectories whose envelope is a single value must reproduce it (one miss allowed);
    and wherever the step count equals the oracle's, the smoothed means are compared point by point."""
    import floor as F
    n = 12
    pd, u0, p = _inputs("vanderpol", n, 106, dp=0.2)
    u0[:, 0] += 0.1 * np.random.default_rng(7).uniform(-1, 1, n)
    g = _gpu_smooth(pd, u0, p, 5, adaptive=True, t0=0.0, t1=0.5)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 5)
    kw = dict(alg=oracle.EK1, smooth=True, adaptive=True, t0=0.0, t1=0.5, save_everystep=True, cov_backend=oracle.COV_REF)
    same = stable = stable_hit = 0
    for i in range(n):
        o = oracle.solve(oracle.make_config(2, 5, 1, **kw), rhs, u0[i], p[i])
        env = {o["n"]} | {s["n"] for _, s in F.variants(oracle, rhs, 2, 5, 1, u0[i], p[i], kw)}
        a, b = g["off"][i], g["off"][i + 1]
        scale = np.abs(o["pu_mu"]).max(axis=0)
        assert min(env) - 1 <= b - a <= max(env) + 1, (i, b - a, sorted(env))
        assert np.max(np.abs(g["U"][b - 1] - o["pu_mu"][-1]) / scale) < 1e-5
        if len(env) == 1:
            stable += 1
            stable_hit += int(b - a == o["n"])
        if b - a == o["n"]:   # same accept pattern; the step times drift by rounding through sigma^2 (oracle QR-vs-Cholesky: <= 4e-3)
            same += 1
            dT = np.abs(g["T"][a:b] - o["t"])
            du = np.abs(np.gradient(o["pu_mu"], o["t"], axis=0)).max(axis=0)
            assert np.all(np.abs(g["U"][a:b] - o["pu_mu"]) <= 1e-6 * scale + 4.0 * du * dT[:, None])
    assert stable_hit >= stable - 1 and same >= 1, (same, stable, stable_hit)


# ------------------------------------------------------------------------------------------------------------------
# smaller reference behaviours fixed in round 2
# ------------------------------------------------------------------------------------------------------------------
def test_fenrir_last_observation_before_the_end_point_is_skipped_like_the_reference(oracle):
    """fenrir.jl:87-99 sets `data_idx = length(data.u) - 1` unconditionally: with data.t[end] < sol.t[end] the last observation
    never enters.  Register sweep and shared-memory sweep against the oracle (pinned in tests/test_oracle_pins.py)."""
    import os
    from test_gpu_parity import _gpu_fenrir
    rng = np.random.default_rng(121)
    data_t = np.array([2.5, 5.0, 7.5])
    data_u = np.array([[1.0, 1.0]]) + 0.3 * rng.standard_normal((3, 2))
    pd, u0, p = _inputs("lotka_volterra", 5, 122, dp=0.05)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, 2, 4), 3)
    cfg = oracle.make_config(2, 3, 4, alg=oracle.EK1, calibrate=False, smooth=True, adaptive=False, dt=2e-2, t0=0.0, t1=10.0,
                             tstops=data_t, cov_backend=oracle.COV_QR)
    want = np.array([oracle.fenrir_data_loglik(cfg, oracle.solve(cfg, rhs, u0[i], p[i]), data_t, data_u, observation_noise_cov=0.09)
                     for i in range(5)])
    for generic in ("0", "1"):
        os.environ["PNODE_FENRIR_GENERIC"] = generic
        try:
            g = _gpu_fenrir(pd, u0, p, 3, data_t, data_u, alg=lib.EK1, calibrate=False, adaptive=False, dt=2e-2, t0=0.0, t1=10.0,
                            observation_noise_cov=0.09)
        finally:
            del os.environ["PNODE_FENRIR_GENERIC"]
        assert np.allclose(g["ll"], want, rtol=1e-8, atol=0), (generic, g["ll"], want)


@pytest.mark.parametrize("smooth", [False, True])
def test_fixed_diffusion_calibration_with_a_non_unit_initial_diffusion(oracle, smooth):
    """calibrate_solution! (src/integrator_utils.jl:46-65): covariances are scaled by the MLE sigma_hat^2, sol.diffusions
    become sigma_hat^2 * initial_diffusion.  (Round 1 rejected save_everystep + calibrate with initial_diffusion != 1.)"""
    pd, u0, p = _inputs("fitzhugh_nagumo", 2, 123, dp=0.05)
    kw = dict(adaptive=False, dt=0.01, t0=0.0, t1=5.0, diffusion=lib.DIFF_FIXED, calibrate=True, initial_diffusion=2.5)
    g = _gpu_smooth(pd, u0, p, 3, **kw) if smooth else _gpu(pd, u0, p, 3, builtin=False, save_everystep=True, **kw)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    for i in range(2):
        o = oracle.solve(oracle.make_config(2, 3, 3, alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, calibrate=True, initial_diffusion=2.5,
                                            smooth=smooth, adaptive=False, dt=0.01, t0=0.0, t1=5.0, save_everystep=True,
                                            cov_backend=oracle.COV_REF), rhs, u0[i], p[i])
        a, b = g["off"][i], g["off"][i + 1]
        assert _rel(g["U"][a:b], o["pu_mu"]) < 1e-10
        assert _relcov(g["C"][a + 1:b], o["pu_cov"][1:]) < 1e-10
        assert _rel(g["Df"][a:b - 1], o["diffusions"][:-1]) < 1e-10
