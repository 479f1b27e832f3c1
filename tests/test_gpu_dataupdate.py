"""Forward data updates on the device: `DataUpdateCallback` (src/callbacks/dataupdate.jl:39-84), `filtering_data_loglik`
(src/data_likelihoods/filtering.jl) and `dalton_data_loglik` (src/data_likelihoods/dalton.jl) over an ensemble of parameters against
one data set.  The oracle's restatement is pinned in tests/test_oracle_pins.py on the reference's own identity (filtering == fenrir at
rtol 1e-6, test/data_likelihoods.jl:31-60)."""
import numpy as np
import pytest

import probnumdiffeq_jl_b200  # noqa: F401
from probnumdiffeq_jl_b200 import api, lib, problems, tracer

from test_gpu_parity import _inputs, _lv_data, _rel, _relcov

pytestmark = pytest.mark.gpu


def _gpu_forward(pd, u0, p, q, data_t, data_u, **kw):
    src = tracer.cuda_source(tracer.trace(pd.f, pd.d, pd.n_p), "UserProb")
    cfg = lib.make_config(d=pd.d, q=q, n_p=pd.n_p, rhs_name="UserProb", rhs_src=src, data_t=data_t, data_u=data_u, data_forward=True, **kw)
    s = lib.Solver(cfg)
    r = s.solve(u0, p)
    n = u0.shape[0]
    out = dict(dll=r.field(lib.F_DATA_LOGLIK), ll=r.field(lib.F_LOGLIK), u=r.field(lib.F_U_FINAL).reshape(n, pd.d),
               cov=r.field(lib.F_U_FINAL_COV).reshape(n, pd.d, pd.d), naccept=r.field(lib.F_NACCEPT), nreject=r.field(lib.F_NREJECT),
               retcode=r.field(lib.F_RETCODE))
    if r.has(lib.F_STEP_OFFSETS):
        out.update(off=r.field(lib.F_STEP_OFFSETS), T=r.field(lib.F_T), U=r.field(lib.F_U).reshape(-1, pd.d), C=r.field(lib.F_U_COV).reshape(-1, pd.d, pd.d))
    r.free()
    s.close()
    return out


@pytest.mark.parametrize("obs,noise", [(None, 0.05), (np.array([[1.0, 0.0]]), 0.05), (np.array([[0.5, 0.5], [0.0, 1.0]]), np.array([[0.05, 0.01], [0.01, 0.08]]))])
def test_forward_data_updates_fixed_steps(oracle, obs, noise):
    """Fixed steps, FixedDiffusion(calibrate=false): data log-likelihood, PN log-likelihood, every saved point (the state is saved after
    the data update, as callbacks run before savevalues!) at 1e-10 / 1e-9 against the oracle.  Full / partial observations, scalar /
    matrix noise."""
    n = 4
    pd, u0, p = _inputs("lotka_volterra", n, 401, dp=0.1)
    rng = np.random.default_rng(402)
    data_t = np.array([0.5, 1.0, 1.5, 2.0])
    y = _lv_data(rng, 0.2, data_t)
    data_u = y if obs is None else y @ obs.T
    kw = dict(alg=lib.EK1, diffusion=lib.DIFF_FIXED, calibrate=False, adaptive=False, dt=0.02, t0=0.0, t1=2.0, smooth=False)
    g = _gpu_forward(pd, u0, p, 3, data_t, data_u, observation_matrix=obs, observation_noise_cov=noise, save_everystep=True, **kw)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    cfg = oracle.make_config(2, 3, 4, alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, calibrate=False, smooth=False, adaptive=False, dt=0.02,
                             t0=0.0, t1=2.0, cov_backend=oracle.COV_REF, data=(data_t, data_u), observation_matrix=obs, observation_noise_cov=noise)
    for i in range(n):
        o = oracle.solve(cfg, rhs, u0[i], p[i])
        a, b = g["off"][i], g["off"][i + 1]
        assert b - a == o["n"] == 101 and g["retcode"][i] == 0
        assert g["dll"][i] == pytest.approx(o["data_loglik"], rel=1e-9)
        assert g["ll"][i] == pytest.approx(o["loglik"], rel=1e-9)
        assert _rel(g["U"][a:b], o["pu_mu"]) < 1e-10
        assert _relcov(g["C"][a + 1:b], o["pu_cov"][1:]) < 1e-9


def test_filtering_and_dalton_loglik_over_a_parameter_ensemble(oracle):
    """api.filtering_data_loglik (adaptive, default tolerances) and api.dalton_data_loglik (fixed steps) for 24 parameter sets against
    one data set, vs the oracle per trajectory; and the reference's compare_data_likelihoods identity on the device:
    filtering == fenrir at rtol 1e-6 (test/data_likelihoods.jl:31-60)."""
    n = 24
    pd, u0, p = _inputs("lotka_volterra", n, 403, dp=0.1)
    rng = np.random.default_rng(404)
    data_t = np.array([2.0, 4.0, 6.0, 8.0, 10.0])
    data_u = _lv_data(rng, 0.1, data_t)
    prob = api.ODEProblem(pd.f, list(pd.u0), (0.0, 10.0), list(pd.p))
    ens = api.EnsembleProblem(prob, prob_func=lambda pr, i, rep: api.remake(pr, u0=list(u0[i]), p=list(p[i])))
    alg = api.EK1(order=3, smooth=True)
    fl = api.filtering_data_loglik(ens, alg, data=(data_t, data_u), observation_noise_cov=0.01, trajectories=n)
    fe = api.fenrir_data_loglik(ens, alg, data=(data_t, data_u), observation_noise_cov=0.01, trajectories=n)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    cfg = oracle.make_config(2, 3, 4, alg=oracle.EK1, smooth=False, adaptive=True, t0=0.0, t1=10.0, save_everystep=False, cov_backend=oracle.COV_REF,
                             data=(data_t, data_u), observation_noise_cov=0.01)
    for i in range(n):
        o = oracle.solve(cfg, rhs, u0[i], p[i])
        assert fl[i] == pytest.approx(o["data_loglik"], rel=1e-6), i
    # test/data_likelihoods.jl:16-35 (`compare_data_likelihoods`: fixed steps dt = 2e-2, sigma = 0.5): the data likelihoods agree.  With a
    # time-fixed, uncalibrated diffusion the forward (filtering) and the backward (Fenrir) evaluation factorise the same Gaussian
    # marginal likelihood in opposite orders -- exactly so for a linear ODE; for Lotka-Volterra the EK1 re-linearises at the conditioned
    # state after every observation, which moves the value by a few 1e-6 (measured 3.5e-6 over 24 parameter sets x 5 observations; the
    # reference's test passes rtol 1e-6 with two observations of one parameter set).  With the default DynamicDiffusion conditioning
    # also changes the later local diffusion estimates, so the identity is looser still.
    data_n = _lv_data(rng, 0.5, data_t)
    ckw = dict(data=(data_t, data_n), observation_noise_cov=0.25, trajectories=n, adaptive=False, dt=0.02)
    fixed = api.EK1(order=3, smooth=True, diffusionmodel=api.FixedDiffusion(calibrate=False))
    fl2 = api.filtering_data_loglik(ens, fixed, **ckw)
    fe2 = api.fenrir_data_loglik(ens, fixed, **ckw)
    assert np.allclose(fl2, fe2, rtol=2e-5), np.max(np.abs(fl2 - fe2) / np.abs(fe2))
    fl3 = api.filtering_data_loglik(ens, alg, **ckw)
    fe3 = api.fenrir_data_loglik(ens, alg, **ckw)
    da3 = api.dalton_data_loglik(ens, alg, **ckw)
    assert np.allclose(fl3, fe3, rtol=5e-2) and np.allclose(da3, fl3, rtol=5e-2)
    assert not np.allclose(fl, fe, rtol=1e-9)   # adaptive steps: conditioning on the data changes the grid as well
    kw = dict(adaptive=False, dt=0.01)
    dal = api.dalton_data_loglik(ens, api.EK1(order=3, smooth=False, diffusionmodel=api.FixedDiffusion(calibrate=False)),
                                 data=(data_t, data_u), observation_noise_cov=0.01, trajectories=n, **kw)
    okw = dict(alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, calibrate=False, smooth=False, adaptive=False, dt=0.01, t0=0.0, t1=10.0,
               save_everystep=False, cov_backend=oracle.COV_REF)
    for i in range(0, n, 6):
        w = oracle.solve(oracle.make_config(2, 3, 4, data=(data_t, data_u), observation_noise_cov=0.01, **okw), rhs, u0[i], p[i])
        wo = oracle.solve(oracle.make_config(2, 3, 4, tstops=data_t[:-1], **okw), rhs, u0[i], p[i])
        # DALTON's value is a difference of two PN log-likelihoods of 1000 steps each (magnitude 5e6 here) plus the data term; the O(0.1)
        # result inherits the rounding of the two sums (1e-13 relative = a few hundred ulps over 1000 summands), measured 5e-8 absolute
        floor = 1e-13 * (abs(w["loglik"]) + abs(wo["loglik"]) + abs(w["data_loglik"]))
        assert dal[i] == pytest.approx(w["data_loglik"] + w["loglik"] - wo["loglik"], rel=1e-8, abs=floor), i
    with pytest.raises(ValueError, match="only works with fixed step sizes"):
        api.dalton_data_loglik(ens, alg, data=(data_t, data_u), observation_noise_cov=0.01, trajectories=n)
