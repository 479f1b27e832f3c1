"""CTA-per-trajectory kernel (pn_cta.cuh) beyond first-order ODEs: second-order problems (`SecondOrderODEProblem`,
src/measurement_models.jl:29-49, src/derivative_utils.jl:35-53, src/caches.jl:100-126) and mass matrices
(src/measurement_models.jl:11-20, src/derivative_utils.jl:43-50) -- SURVEY.md section 8f rank 3 for D > 32, in particular
Pleiades in the reference's own second-order form (benchmarks/pleiades.jmd: 14 second-order equations, D = 56 at order 3).
The kernel is size-generic, so the small cases run the very same code with `lanes_per_traj=256`."""
import numpy as np
import pytest

import probnumdiffeq_jl_b200  # noqa: F401
from probnumdiffeq_jl_b200 import lib, problems, tracer

from test_gpu_parity import _gpu, _gpu_second, _inputs, _mass_inputs, _rel, _relcov, _second_inputs, _twobody_first_order

pytestmark = pytest.mark.gpu


def test_cta_second_order_fixed_steps_and_adaptive(oracle):
    """Two-body problem (test/secondorderode.jl:15-19) through the CTA kernel: fixed steps + FixedDiffusion at 1e-10 on every
    saved sol.u = (du, u), its 4 x 4 covariance, diffusions, log-likelihood; adaptive: identical accept / reject / nf counts."""
    u0, p = _second_inputs(3, 171)
    rhs = oracle.compile_rhs(tracer.trace(_twobody_first_order, 4, 1), 3)
    g = _gpu_second(u0, p, 3, alg=lib.EK1, adaptive=False, dt=1e-3, t0=0.0, t1=0.1, diffusion=lib.DIFF_FIXED, calibrate=True, lanes_per_traj=256)
    for backend in (oracle.COV_QR, oracle.COV_REF):
        cfg = oracle.make_config(2, 3, 1, alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, calibrate=True, smooth=False, adaptive=False, dt=1e-3,
                                 t0=0.0, t1=0.1, cov_backend=backend, second_order=True)
        for i in range(3):
            o = oracle.solve(cfg, rhs, u0[i], p[i])
            a, b = g["off"][i], g["off"][i + 1]
            assert b - a == o["n"] == 101 and g["retcode"][i] == 0
            assert np.array_equal(g["U"][a], u0[i]) and not g["C"][a].any()
            assert _rel(g["U"][a:b], o["pu_mu"]) < 1e-10
            assert _relcov(g["C"][a + 1:b], o["pu_cov"][1:]) < 1e-10
            assert g["loglik"][i] == pytest.approx(o["loglik"], rel=1e-10)
            assert _rel(g["Df"][a:b - 1], o["diffusions"][:-1]) < 1e-10
            assert g["nf"][i] == o["nf"]
    n = 8
    u0, p = _second_inputs(n, 172)
    g = _gpu_second(u0, p, 3, alg=lib.EK1, adaptive=True, t0=0.0, t1=0.1, lanes_per_traj=256)
    cfg = oracle.make_config(2, 3, 1, alg=oracle.EK1, smooth=False, adaptive=True, t0=0.0, t1=0.1, cov_backend=oracle.COV_REF, second_order=True)
    for i in range(n):
        o = oracle.solve(cfg, rhs, u0[i], p[i])
        a, b = g["off"][i], g["off"][i + 1]
        assert g["naccept"][i] == o["naccept"] and g["nreject"][i] == o["nreject"] and g["nf"][i] == o["nf"], i
        assert np.linalg.norm(g["U"][b - 1] - o["pu_mu"][-1]) <= 1e-8 * np.linalg.norm(o["pu_mu"][-1])


def test_cta_mass_matrix_fixed_steps_and_rober_dae(oracle):
    """A dense non-symmetric M on fixed steps with FixedDiffusion (1e-10), and the Robertson DAE M = diag(1, 1, 0)
    (benchmarks/rober.jmd:37-46) on a short adaptive run, through the CTA kernel."""
    pd, u0, p = _mass_inputs(3, 162)
    u0 = u0 + np.array([0.0, 0.1, 0.2])
    p = p * np.array([1.0, 1e-6, 1e-3])
    M = np.array([[1.0, 0.2, 0.0], [0.0, 1.0, 0.0], [0.1, 0.0, 0.5]])
    kw = dict(adaptive=False, dt=1e-3, t0=0.0, t1=0.2, diffusion=lib.DIFF_FIXED, calibrate=True, mass_matrix=M)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    cfg = oracle.make_config(3, 3, 3, alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, calibrate=True, smooth=False, adaptive=False,
                             dt=1e-3, t0=0.0, t1=0.2, cov_backend=oracle.COV_REF, mass_matrix=M)
    g = _gpu(pd, u0, p, 3, builtin=False, save_everystep=True, lanes_per_traj=256, **kw)
    for i in range(u0.shape[0]):
        o = oracle.solve(cfg, rhs, u0[i], p[i])
        a, b = g["off"][i], g["off"][i + 1]
        assert b - a == o["n"] == 201
        assert _rel(g["U"][a:b], o["pu_mu"]) < 1e-10
        assert _relcov(g["C"][a + 1:b], o["pu_cov"][1:]) < 1e-9
        assert g["loglik"][i] == pytest.approx(o["loglik"], rel=1e-9)
    # the same configuration through the register kernel (lanes auto): the two kernels agree to rounding
    g2 = _gpu(pd, u0, p, 3, builtin=False, save_everystep=True, **kw)
    assert _rel(g["U"], g2["U"]) < 1e-11 and _relcov(g["C"][1:], g2["C"][1:]) < 1e-9
    pd, u0, p = _mass_inputs(4, 163)
    Md = np.diag([1.0, 1.0, 0.0])
    g = _gpu(pd, u0, p, 3, builtin=False, adaptive=True, t0=0.0, t1=1e-2, mass_matrix=Md, lanes_per_traj=256)
    g2 = _gpu(pd, u0, p, 3, builtin=False, adaptive=True, t0=0.0, t1=1e-2, mass_matrix=Md)
    assert np.all(g["retcode"] == 0)
    assert np.array_equal(g["naccept"], g2["naccept"]) and np.array_equal(g["nreject"], g2["nreject"]) and np.array_equal(g["nf"], g2["nf"])
    assert _rel(g["u"], g2["u"]) < 1e-8


def _gpu_pleiades_second(u0, q=3, **kw):
    pd = problems.PROBLEMS["pleiades"]
    cfg = lib.make_config(d=14, q=q, n_p=0, rhs_name="builtin:pleiades", second_order=True, **kw)
    s = lib.Solver(cfg)
    n = u0.shape[0]
    r = s.solve(u0, np.zeros((n, 0)))
    out = dict(info=r.info, u=r.field(lib.F_U_FINAL).reshape(n, 28), cov=r.field(lib.F_U_FINAL_COV).reshape(n, 28, 28),
               naccept=r.field(lib.F_NACCEPT), nreject=r.field(lib.F_NREJECT), nf=r.field(lib.F_NF), retcode=r.field(lib.F_RETCODE),
               loglik=r.field(lib.F_LOGLIK), t_end=r.field(lib.F_T_END))
    if r.has(lib.F_X_FINAL):
        out.update(x=r.field(lib.F_X_FINAL).reshape(n, 56), xR=r.field(lib.F_X_FINAL_COVFAC).reshape(n, 56, 56))
    r.free()
    s.close()
    return out


def test_pleiades_as_14_second_order_equations_d56(oracle):
    """benchmarks/pleiades.jmd `prob2 = SecondOrderODEProblem(...)`: d = 14, D = 56 at order 3.  The first-order function of
    problems.py is the first-order form F([du; u]) = [f1(u); du] the second-order interface takes.  Fixed steps +
    FixedDiffusion: end point, its 28 x 28 covariance, the 56 x 56 state covariance and the log-likelihood at 1e-10; adaptive
    (default tolerances): accept / reject counts within the slack of the first-order cfg4 test, end points 1e-8 on matching
    patterns; and the second-order form takes far less time per step than the first-order one (the reason it exists)."""
    n = 3
    pd, u0, _ = _inputs("pleiades", n, 2)
    u0[:, 14:] += 1e-3 * np.random.default_rng(2).uniform(-1, 1, (n, 14))
    rhs = oracle.compile_rhs(tracer.trace(pd.f, 28, 0), 3)
    gf = _gpu_pleiades_second(u0[:1], adaptive=False, dt=0.01, t0=0.0, t1=0.5, diffusion=lib.DIFF_FIXED, save_state=True)
    for backend in (oracle.COV_QR, oracle.COV_REF):
        of = oracle.solve(oracle.make_config(14, 3, 0, alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, smooth=False, adaptive=False, dt=0.01,
                                             t0=0.0, t1=0.5, save_everystep=False, cov_backend=backend, second_order=True), rhs, u0[0], [])
        assert np.linalg.norm(gf["u"][0] - of["pu_mu"][-1]) / np.linalg.norm(of["pu_mu"][-1]) < 1e-10
        assert _relcov(gf["cov"][0], of["pu_cov"][-1]) < 1e-10
        Sg = gf["xR"][0].T @ gf["xR"][0]
        So = of["x_filt_R"][-1].T @ of["x_filt_R"][-1]
        assert np.linalg.norm(Sg - So) / np.linalg.norm(So) < 1e-10
        assert gf["loglik"][0] == pytest.approx(of["loglik"], rel=1e-9)
    g = _gpu_pleiades_second(u0, adaptive=True, t0=0.0, t1=3.0)
    cfg = oracle.make_config(14, 3, 0, alg=oracle.EK1, smooth=False, adaptive=True, t0=0.0, t1=3.0, save_everystep=False,
                             cov_backend=oracle.COV_REF, second_order=True)
    assert np.all(g["retcode"] == 0) and np.all(g["t_end"] == 3.0)
    for i in range(n):
        o = oracle.solve(cfg, rhs, u0[i], [])
        scale = np.abs(o["pu_mu"][-1]).max()
        assert abs(int(g["naccept"][i]) - o["naccept"]) <= 2 and abs(int(g["nreject"][i]) - o["nreject"]) <= 2, (g["naccept"][i], o["naccept"])
        tol = 1e-8 if (g["naccept"][i] == o["naccept"] and g["nreject"][i] == o["nreject"]) else 1e-4
        assert np.max(np.abs(g["u"][i] - o["pu_mu"][-1])) / scale < tol
    # cost per step: 2048 trajectories in both forms
    m = 2048
    _, U0, _ = _inputs("pleiades", m, 3)
    U0[:, 14:] += 1e-3 * np.random.default_rng(3).uniform(-1, 1, (m, 14))
    g2 = _gpu_pleiades_second(U0, adaptive=True, t0=0.0, t1=3.0)
    g1 = _gpu(pd, U0, np.zeros((m, 0)), 3, builtin=True, adaptive=True, t0=0.0, t1=3.0)
    per2 = g2["info"].kernel_ms / g2["info"].total_accepted
    per1 = g1["info"].kernel_ms / g1["info"].total_accepted
    print(f"pleiades second-order D=56: {1e3 / per2:.3e} steps/s; first-order D=112: {1e3 / per1:.3e} steps/s")
    assert np.all(g2["retcode"] == 0) and per2 < 0.5 * per1


def test_cta_kernel_with_smoothing_small_and_pleiades_d112(oracle):
    """Filter + smoother for states beyond the register sweeps (SURVEY.md section 8d cfg4 asks for both numbers): the CTA kernel
    saves its filter records, the warp-per-trajectory sweep (reference formulas literally, pn_smooth.cuh) runs them backwards --
    out of shared memory for mid-sized states, out of an L2-resident global workspace for Pleiades (D = 112, 7 D^2 doubles per
    warp).  Fixed steps + FixedDiffusion: smoothed means 1e-10, covariances 1e-9 against the oracle, every saved point."""
    from test_gpu_parity import _gpu_smooth
    pd, u0, p = _inputs("fitzhugh_nagumo", 2, 181, dp=0.05)
    kw = dict(adaptive=False, dt=0.01, t0=0.0, t1=2.0, diffusion=lib.DIFF_FIXED, calibrate=True)
    g = _gpu_smooth(pd, u0, p, 3, lanes_per_traj=256, **kw)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    cfg = oracle.make_config(2, 3, 3, alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, calibrate=True, smooth=True, adaptive=False, dt=0.01,
                             t0=0.0, t1=2.0, cov_backend=oracle.COV_REF)
    for i in range(2):
        o = oracle.solve(cfg, rhs, u0[i], p[i])
        a, b = g["off"][i], g["off"][i + 1]
        assert b - a == o["n"] == 201
        assert _rel(g["U"][a:b], o["pu_mu"]) < 1e-10
        assert _relcov(g["C"][a + 1:b], o["pu_cov"][1:]) < 1e-9
    n = 2
    pd, u0, _ = _inputs("pleiades", n, 2)
    u0[:, 14:] += 1e-3 * np.random.default_rng(2).uniform(-1, 1, (n, 14))
    p = np.zeros((n, 0))
    kw = dict(adaptive=False, dt=0.01, t0=0.0, t1=0.3, diffusion=lib.DIFF_FIXED, calibrate=True)
    g = _gpu_smooth(pd, u0, p, 3, **kw)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, 28, 0), 3)
    cfg = oracle.make_config(28, 3, 0, alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, calibrate=True, smooth=True, adaptive=False, dt=0.01,
                             t0=0.0, t1=0.3, cov_backend=oracle.COV_REF)
    for i in range(n):
        o = oracle.solve(cfg, rhs, u0[i], [])
        a, b = g["off"][i], g["off"][i + 1]
        assert b - a == o["n"] == 31
        assert _rel(g["U"][a:b], o["pu_mu"]) < 1e-10
        assert _relcov(g["C"][a + 1:b], o["pu_cov"][1:]) < 1e-9
    # adaptive ensemble: ragged lengths through the pilot-sized save pass; smoothed end point = filtered end point
    m = 64
    _, U0, _ = _inputs("pleiades", m, 5)
    U0[:, 14:] += 1e-3 * np.random.default_rng(5).uniform(-1, 1, (m, 14))
    gs = _gpu_smooth(pd, U0, np.zeros((m, 0)), 3, adaptive=True, t0=0.0, t1=1.0)
    gf = _gpu(pd, U0, np.zeros((m, 0)), 3, builtin=False, adaptive=True, t0=0.0, t1=1.0)
    ends = gs["off"][1:] - 1
    assert np.array_equal(np.diff(gs["off"]) - 1, gf["naccept"])
    assert _rel(gs["U"][ends], gf["u"]) < 1e-12
    print("pleiades D=112 filter+smoother, 64 trajectories to t=1: kernel_ms", gs["info"].kernel_ms, "steps", int(gf["naccept"].sum()))


@pytest.mark.parametrize("smooth", [True, False])
def test_dense_output_beyond_the_register_kernels(oracle, smooth):
    """sol(saveat) (src/solution.jl:330-398) where the dense-output register kernels end (D > 32): warp-per-pair predict
    (pn_dense_gen.cuh) + one backward step of the shared-memory sweep on the two-point virtual trajectories.  Pleiades, D = 112, fixed
    steps, FixedDiffusion (calibrated): against the oracle's `interpolate` at grid points inside steps, on saved points, at t0 and t1;
    and Lotka-Volterra, order 6, forced onto the CTA filter (lanes_per_traj = 256): register predict kernel, shared-memory sweep."""
    from test_gpu_parity import _gpu_dense
    n = 2
    pd, u0, _ = _inputs("pleiades", n, 7)
    u0[:, 14:] += 1e-3 * np.random.default_rng(7).uniform(-1, 1, (n, 14))
    p = np.zeros((n, 0))
    saveat = np.array([0.0, 0.0042, 0.05, 0.1234, 0.2])
    kw = dict(adaptive=False, dt=0.01, t0=0.0, t1=0.2, diffusion=lib.DIFF_FIXED, calibrate=True)
    g = _gpu_dense(pd, u0, p, 3, saveat, smooth=smooth, **kw)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, 28, 0), 3)
    cfg = oracle.make_config(28, 3, 0, alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, calibrate=True, smooth=smooth, adaptive=False, dt=0.01,
                             t0=0.0, t1=0.2, cov_backend=oracle.COV_REF)
    for i in range(n):
        o = oracle.solve(cfg, rhs, u0[i], [])
        for j, t in enumerate(saveat):
            m, c = oracle.interpolate(cfg, o, float(t))
            assert _rel(g["U"][i, j], m) < 1e-10, (i, j)
            if np.linalg.norm(c) > 0:
                assert np.linalg.norm(g["C"][i, j] - c) / np.linalg.norm(c) < 1e-8, (i, j)
            else:
                assert not g["C"][i, j].any()
    # CTA filter + register predict + shared-memory sweep over the virtual trajectories
    pd, u0, p = _inputs("lotka_volterra", 2, 8, du0=0.05)
    saveat = np.array([0.0, 0.013, 0.4, 0.77777, 1.0])
    kw = dict(adaptive=False, dt=0.02, t0=0.0, t1=1.0, diffusion=lib.DIFF_FIXED, calibrate=True)
    g = _gpu_dense(pd, u0, p, 6, saveat, smooth=smooth, lanes_per_traj=256, **kw)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 6)
    okw = dict(alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, calibrate=True, smooth=smooth, adaptive=False, dt=0.02, t0=0.0, t1=1.0)
    cfg = oracle.make_config(2, 6, 4, cov_backend=oracle.COV_REF, **okw)
    cfg_qr = oracle.make_config(2, 6, 4, cov_backend=oracle.COV_QR, **okw)
    for i in range(2):
        o = oracle.solve(cfg, rhs, u0[i], p[i])
        o_qr = oracle.solve(cfg_qr, rhs, u0[i], p[i])
        for j, t in enumerate(saveat):
            m, c = oracle.interpolate(cfg, o, float(t))
            _, c_qr = oracle.interpolate(cfg_qr, o_qr, float(t))
            assert _rel(g["U"][i, j], m) < 1e-10, (i, j)
            if np.linalg.norm(c) > 0:
                # order 6: the 7 x 7 preconditioned process-noise matrix has condition number 5e8, so the covariances (1e-25 here) carry a
                # relative rounding floor of eps * cond ~ 1e-7 (measured: 1.4e-7 against the oracle, 7e-9 ... 1e-7 between the oracle's own
                # two covariance back-ends); the bound is ten times that
                floor = np.linalg.norm(c_qr - c) / np.linalg.norm(c)
                assert np.linalg.norm(g["C"][i, j] - c) / np.linalg.norm(c) < max(1e-6, 20 * floor), (i, j, floor)
