"""The structured layouts beyond the register budget of their thread-per-trajectory kernels: EK0 on the IsometricKronecker layout at
orders above 7 (test/correctness.jl:57-87 runs EK0(order=8)) and the BlocksOfDiagonals layout (DiagonalEK1, EK0 with multivariate
diffusions; src/algorithms.jl:132-151, :343-393) at d = 28 (Pleiades: the O(d) algorithms are meant for large d).  The kernels are
pn_kron.cuh / pn_blocks.cuh with their factors in thread-local memory; the arithmetic is the one the small-state tests check."""
import numpy as np
import pytest

import probnumdiffeq_jl_b200  # noqa: F401
from probnumdiffeq_jl_b200 import lib, problems, tracer

from test_gpu_parity import _gpu, _inputs, _rel, _relcov

pytestmark = pytest.mark.gpu


def test_ek0_order_8_on_the_kronecker_layout(oracle):
    """EK0(order=8), Lotka-Volterra on [0, 1].  Adaptive: the oracle's accept / reject counts per trajectory, end points 1e-7 (the
    reference's accuracy threshold for this row, test/correctness.jl:57-87, is pinned on the oracle in tests/test_oracle_pins.py).
    One compilation only: the order-8 kernel keeps its factor in thread-local memory and takes NVRTC 25 s."""
    n = 4
    pd, u0, p = _inputs("lotka_volterra", n, 601, du0=0.1, dp=0.1)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 8)
    g = _gpu(pd, u0, p, 8, builtin=False, alg=lib.EK0, adaptive=True, t0=0.0, t1=1.0)
    o = oracle.solve_ensemble(oracle.make_config(2, 8, 4, alg=oracle.EK0, smooth=False, save_everystep=False, adaptive=True, t0=0.0, t1=1.0,
                                                 cov_backend=oracle.COV_REF), rhs, u0, p)
    assert np.all(g["retcode"] == 0)
    assert np.array_equal(g["naccept"], o["naccept"]) and np.array_equal(g["nreject"], o["nreject"])
    assert _rel(g["u"], o["u_final"]) < 1e-7


def test_blocks_layout_at_d_28(oracle):
    """Pleiades (d = 28, order 3: 448 doubles of block factors per trajectory).  DiagonalEK1, fixed steps, static diffusion: every saved
    mean 1e-10, every covariance 1e-9, log-likelihood 1e-9.  EK0 with DynamicMVDiffusion, adaptive: the oracle's step counts, end
    points 1e-8."""
    n = 3
    pd, u0, p = _inputs("pleiades", n, 602, du0=1e-3)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    kw = dict(adaptive=False, dt=0.01, t0=0.0, t1=0.1, diffusion=lib.DIFF_FIXED, calibrate=False)
    g = _gpu(pd, u0, p, 3, builtin=False, alg=lib.DIAGONAL_EK1, save_everystep=True, **kw)
    for i in range(n):
        o = oracle.solve(oracle.make_config(28, 3, 0, alg=oracle.DIAGONAL_EK1, diffusion=oracle.DIFF_FIXED, calibrate=False, smooth=False,
                                            adaptive=False, dt=0.01, t0=0.0, t1=0.1, cov_backend=oracle.COV_REF), rhs, u0[i], p[i])
        lo, hi = g["off"][i], g["off"][i + 1]
        assert hi - lo == o["n"] == 11 and g["retcode"][i] == 0
        assert _rel(g["U"][lo:hi], o["pu_mu"]) < 1e-10
        assert _relcov(g["C"][lo + 1:hi], o["pu_cov"][1:]) < 1e-9
        assert g["loglik"][i] == pytest.approx(o["loglik"], rel=1e-9)
    g = _gpu(pd, u0, p, 3, builtin=False, alg=lib.EK0, diffusion=lib.DIFF_DYNAMIC_MV, adaptive=True, t0=0.0, t1=0.5)
    o = oracle.solve_ensemble(oracle.make_config(28, 3, 0, alg=oracle.EK0, diffusion=oracle.DIFF_DYNAMIC_MV, smooth=False, save_everystep=False,
                                                 adaptive=True, t0=0.0, t1=0.5, cov_backend=oracle.COV_REF), rhs, u0, p)
    assert np.all(g["retcode"] == 0)
    assert np.array_equal(g["naccept"], o["naccept"]) and np.array_equal(g["nreject"], o["nreject"])
    assert _rel(g["u"], o["u_final"]) < 1e-8
    assert _relcov(g["cov"], o["cov_final"]) < 1e-6
