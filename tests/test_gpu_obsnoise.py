"""`pn_observation_noise` (EK0 / EK1 / DiagonalEK1 keyword, src/algorithms.jl:195-393): S += R (src/perform_step.jl:182-188), the
filtered covariance gains K R K' and is re-factorised (src/filtering/update.jl:125-136).  The CUDA kernels on all three covariance
layouts against the oracle, whose step is pinned against the textbook Kalman update in tests/test_oracle_pins.py.  Noise shapes as in
test/observation_noise.jl: a number, a Diagonal, a full matrix."""
import numpy as np
import pytest

import probnumdiffeq_jl_b200  # noqa: F401
from probnumdiffeq_jl_b200 import lib, problems, tracer

from test_gpu_parity import _gpu, _inputs, _rel, _relcov

pytestmark = pytest.mark.gpu

FULL = np.array([[0.1, 0.01], [0.01, 0.1]])


@pytest.mark.parametrize("alg,diffusion,noise", [
    (lib.EK1, lib.DIFF_FIXED, 0.1), (lib.EK1, lib.DIFF_FIXED, np.array([0.1, 0.2])), (lib.EK1, lib.DIFF_FIXED, FULL),
    (lib.EK0, lib.DIFF_FIXED, 0.1), (lib.EK0, lib.DIFF_FIXED_MV, np.array([0.1, 0.2])), (lib.DIAGONAL_EK1, lib.DIFF_FIXED, np.array([0.1, 0.2])),
    (lib.DIAGONAL_EK1, lib.DIFF_FIXED_MV, 0.3)])
def test_observation_noise_fixed_steps_static_diffusion(oracle, alg, diffusion, noise):
    """Every saved point of a fixed-step solve with FixedDiffusion(calibrate=false): means and covariances at 1e-10, the
    log-likelihood at 1e-9, on the dense, Kronecker and blocks layouts."""
    pd, u0, p = _inputs("lotka_volterra", 3, 301, du0=0.1, dp=0.1)
    kw = dict(adaptive=False, dt=0.02, t0=0.0, t1=2.0, calibrate=False)
    g = _gpu(pd, u0, p, 3, builtin=False, alg=alg, diffusion=diffusion, save_everystep=True, pn_observation_noise=noise, **kw)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    for backend in (oracle.COV_REF, oracle.COV_QR):
        cfg = oracle.make_config(2, 3, 4, alg=alg, diffusion=diffusion, smooth=False, cov_backend=backend, pn_observation_noise=noise, **kw)
        for i in range(3):
            o = oracle.solve(cfg, rhs, u0[i], p[i])
            a, b = g["off"][i], g["off"][i + 1]
            assert b - a == o["n"] == 101 and g["retcode"][i] == 0
            assert _rel(g["U"][a:b], o["pu_mu"]) < 1e-10
            assert _relcov(g["C"][a + 1:b], o["pu_cov"][1:]) < 1e-10
            assert g["loglik"][i] == pytest.approx(o["loglik"], rel=1e-9)
    # the noise matters: the noise-free solve differs
    g0 = _gpu(pd, u0, p, 3, builtin=False, alg=alg, diffusion=diffusion, save_everystep=True, **kw)
    assert _relcov(g["C"][1:], g0["C"][1:]) > 1e-3


@pytest.mark.parametrize("alg,noise", [(lib.EK1, FULL), (lib.EK0, 0.05), (lib.DIAGONAL_EK1, np.array([0.05, 0.1]))])
def test_observation_noise_adaptive_dynamic_diffusion(oracle, alg, noise):
    """Adaptive solves with the default DynamicDiffusion: identical accept / reject sequences as the oracle, end points at 1e-8;
    smoothing on top of a noisy filter runs through the unchanged sweep (the backward kernels do not involve R)."""
    n = 16
    pd, u0, p = _inputs("lotka_volterra", n, 302, du0=0.1, dp=0.1)
    g = _gpu(pd, u0, p, 3, builtin=False, alg=alg, adaptive=True, t0=0.0, t1=5.0, pn_observation_noise=noise)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    cfg = oracle.make_config(2, 3, 4, alg=alg, smooth=False, adaptive=True, t0=0.0, t1=5.0, save_everystep=False, cov_backend=oracle.COV_REF,
                             pn_observation_noise=noise)
    o = oracle.solve_ensemble(cfg, rhs, u0, p)
    assert np.all(g["retcode"] == 0)
    assert np.array_equal(g["naccept"], o["naccept"]) and np.array_equal(g["nreject"], o["nreject"])
    assert _rel(g["u"], o["u_final"]) < 1e-8
    assert _relcov(g["cov"], o["cov_final"]) < 1e-6
    gs = _gpu(pd, u0[:4], p[:4], 3, builtin=False, alg=alg, adaptive=True, t0=0.0, t1=5.0, smooth=True, save_everystep=True,
              pn_observation_noise=noise)
    cfgs = oracle.make_config(2, 3, 4, alg=alg, smooth=True, adaptive=True, t0=0.0, t1=5.0, cov_backend=oracle.COV_REF, pn_observation_noise=noise)
    for i in range(4):
        os_ = oracle.solve(cfgs, rhs, u0[i], p[i])
        a, b = gs["off"][i], gs["off"][i + 1]
        assert b - a == os_["n"]
        assert _rel(gs["U"][a:b], os_["pu_mu"]) < 1e-7
