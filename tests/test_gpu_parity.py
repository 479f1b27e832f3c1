"""GPU parity: the CUDA path (through the C ABI of libpnode_cuda.so) against the CPU oracle on the same seeded
inputs, plus size-independent properties at BASELINE.json's full ensemble sizes.

Tolerances (BASELINE.json north_star; SURVEY.md H0 for the measured rounding floor):
  * fixed steps, static diffusion (FixedDiffusion with/without calibration): means and covariances rtol 1e-10;
  * fixed steps, DynamicDiffusion: sigma^2 = z'(HQH')^-1 z amplifies one-ulp differences (oracle-vs-oracle floor
    5e-11 on means, 6e-7 on covariances), so means and covariances are held to max(1e-10, 20 x floor) with the floor
    measured in the same test (tests/floor.py), and the sigma^2-teacher-forced run (kernel consumes the oracle's
    sigma^2 sequence) must reach 1e-10 on everything;
  * adaptive, non-stiff: identical accept/reject counts, |dt| sequences to 1e-7, solutions rtol 1e-8;
  * adaptive, stiff (Van der Pol, OREGO): step-teacher-forced (kernel replays the oracle's accepted grid) rtol 1e-8,
    plus free-running step counts within 3 %.
"""
import numpy as np
import pytest

import probnumdiffeq_jl_b200  # noqa: F401
from probnumdiffeq_jl_b200 import lib, problems, tracer

import floor as rounding

pytestmark = pytest.mark.gpu


def _inputs(name, n, seed, du0=0.0, dp=0.1):
    pd = problems.PROBLEMS[name]
    rng = np.random.default_rng(seed)
    u0 = np.array(pd.u0, dtype=float)[None, :] * (1 + du0 * rng.uniform(-1, 1, (n, pd.d)))
    p = np.array(pd.p, dtype=float)[None, :] * (1 + dp * rng.uniform(-1, 1, (n, pd.n_p)))
    return pd, np.ascontiguousarray(u0), np.ascontiguousarray(p)


def _gpu(pd, u0, p, q, *, builtin=True, **kw):
    if builtin:
        cfg = lib.make_config(d=pd.d, q=q, n_p=pd.n_p, rhs_name="builtin:" + pd.name, **kw)
    else:
        src = tracer.cuda_source(tracer.trace(pd.f, pd.d, pd.n_p), "UserProb")
        cfg = lib.make_config(d=pd.d, q=q, n_p=pd.n_p, rhs_name="UserProb", rhs_src=src, **kw)
    s = lib.Solver(cfg)
    r = s.solve(u0, p)
    n, d = u0.shape
    out = dict(info=r.info, u=r.field(lib.F_U_FINAL).reshape(n, d), cov=r.field(lib.F_U_FINAL_COV).reshape(n, d, d),
               naccept=r.field(lib.F_NACCEPT), nreject=r.field(lib.F_NREJECT), nf=r.field(lib.F_NF),
               retcode=r.field(lib.F_RETCODE), loglik=r.field(lib.F_LOGLIK), t_end=r.field(lib.F_T_END),
               diffusion=r.field(lib.F_DIFFUSION))
    if r.has(lib.F_STEP_OFFSETS):
        out.update(off=r.field(lib.F_STEP_OFFSETS), T=r.field(lib.F_T), U=r.field(lib.F_U).reshape(-1, d),
                   C=r.field(lib.F_U_COV).reshape(-1, d, d), Df=r.field(lib.F_DIFFUSIONS))
    if r.has(lib.F_X_FINAL):
        D = pd.d * (q + 1)
        out.update(x=r.field(lib.F_X_FINAL).reshape(n, D), xR=r.field(lib.F_X_FINAL_COVFAC).reshape(n, D, D))
    r.free()
    s.close()
    return out


def _rel(a, b):
    """Norm-wise relative difference (Julia `isapprox` semantics, as the reference's tests use): for 2-d inputs the
    norm is taken per row (per time point / per trajectory), so zero crossings of one component do not blow it up."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    if a.ndim >= 2:
        nb = np.linalg.norm(b, axis=-1)
        return np.max(np.linalg.norm(a - b, axis=-1) / (nb + 1e-300))
    return np.max(np.abs(a - b) / (np.abs(b) + 1e-300))


def _relcov(a, b):
    nb = np.linalg.norm(b, axis=(-2, -1))
    ok = nb > 0
    return np.max(np.linalg.norm(a - b, axis=(-2, -1))[ok] / nb[ok]) if ok.any() else 0.0


# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("calibrate", [True, False])
def test_cfg1_fixed_steps_static_diffusion(oracle, calibrate):
    """BASELINE cfg1: FitzHugh-Nagumo EK1(order=3), dt = 0.01 on [0, 20] (2000 steps), every step compared."""
    pd, u0, p = _inputs("fitzhugh_nagumo", 3, 0, dp=0.05)
    g = _gpu(pd, u0, p, 3, builtin=False, adaptive=False, dt=0.01, t0=0.0, t1=20.0, diffusion=lib.DIFF_FIXED,
             calibrate=calibrate, save_everystep=True, save_state=True)
    tr = tracer.trace(pd.f, pd.d, pd.n_p)
    rhs = oracle.compile_rhs(tr, 3)
    for i in range(u0.shape[0]):
        o = oracle.solve(oracle.make_config(2, 3, 3, alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, calibrate=calibrate, smooth=False,
                                            adaptive=False, dt=0.01, t0=0.0, t1=20.0, save_everystep=True,
                                            cov_backend=oracle.COV_QR), rhs, u0[i], p[i])
        a, b = g["off"][i], g["off"][i + 1]
        assert b - a == o["n"] == 2001
        assert np.array_equal(g["T"][a:b], o["t"])
        assert _rel(g["U"][a:b], o["pu_mu"]) < 1e-10
        assert _relcov(g["C"][a + 1:b], o["pu_cov"][1:]) < 1e-10   # calibrate_solution! applied to saved steps
        assert _rel(g["Df"][a:b - 1], o["diffusions"][:-1]) < 1e-10
        assert _relcov(g["cov"][i], o["pu_cov"][-1]) < 1e-10
        assert g["loglik"][i] == pytest.approx(o["loglik"], rel=1e-10)
        assert g["diffusion"][i] == pytest.approx(o["diffusions"][0], rel=1e-10)
        # full state: x_filt mean and R'R
        assert np.linalg.norm(g["x"][i][:2] - o["x_filt_mu"][-1][:2]) / np.linalg.norm(o["x_filt_mu"][-1][:2]) < 1e-10
        Sg = g["xR"][i].T @ g["xR"][i]
        So = o["x_filt_R"][-1].T @ o["x_filt_R"][-1]
        assert np.linalg.norm(Sg - So) / np.linalg.norm(So) < 1e-10


def test_cfg1_fixed_steps_dynamic_diffusion_and_teacher_forcing(oracle):
    pd, u0, p = _inputs("fitzhugh_nagumo", 1, 1, dp=0.0)
    tr = tracer.trace(pd.f, pd.d, pd.n_p)
    rhs = oracle.compile_rhs(tr, 3)
    kw = dict(alg=oracle.EK1, diffusion=oracle.DIFF_DYNAMIC, smooth=False, adaptive=False, dt=0.01, t0=0.0, t1=20.0,
              save_everystep=True)
    o = oracle.solve(oracle.make_config(2, 3, 3, cov_backend=oracle.COV_QR, **kw), rhs, u0[0], p[0])
    g = _gpu(pd, u0, p, 3, builtin=False, adaptive=False, dt=0.01, t0=0.0, t1=20.0, save_everystep=True)
    assert np.array_equal(g["T"], o["t"])
    # free sigma^2: hold the kernel to 20 x the oracle's own rounding floor (Cholesky-vs-QR, FMA, one-ulp variants), measured here
    kwq = dict(cov_backend=oracle.COV_QR, **kw)
    f_mu = rounding.floor(oracle, rhs, 2, 3, 3, u0[0], p[0], kwq, o, lambda s_, b_: _rel(s_["pu_mu"], b_["pu_mu"]))
    f_cov = rounding.floor(oracle, rhs, 2, 3, 3, u0[0], p[0], kwq, o, lambda s_, b_: _relcov(s_["pu_cov"][1:], b_["pu_cov"][1:]))
    e_mu, e_cov = _rel(g["U"], o["pu_mu"]), _relcov(g["C"][1:], o["pu_cov"][1:])
    assert e_mu < max(1e-10, 20 * f_mu), (e_mu, f_mu)           # measured floor ~1e-10
    assert e_cov < max(1e-10, 20 * f_cov), (e_cov, f_cov)       # floor 6e-7
    # sigma^2 itself is rounding noise wherever z crosses zero (and at the first steps, where z is *only* rounding
    # noise): compare the sequence norm-wise against the oracle's own floor under one-ulp perturbations
    def m_sig(s_, b_):
        return np.linalg.norm(s_["diffusions"][:-1] - b_["diffusions"][:-1]) / np.linalg.norm(b_["diffusions"][:-1])

    fl = rounding.floor(oracle, rhs, 2, 3, 3, u0[0], p[0], kwq, o, m_sig)
    assert m_sig(dict(diffusions=g["Df"]), o) < max(1e-6, 20 * fl)
    # sigma^2 teacher forcing: the kernel consumes the oracle's sigma^2 sequence -> everything reaches 1e-10
    sig = o["diffusions"][:-1]
    o2 = oracle.solve(oracle.make_config(2, 3, 3, forced_diffusion=sig, cov_backend=oracle.COV_QR, **kw), rhs, u0[0], p[0])
    g2 = _gpu(pd, u0, p, 3, builtin=False, adaptive=False, dt=0.01, t0=0.0, t1=20.0, save_everystep=True,
              forced_diffusion=sig)
    assert _rel(g2["U"], o2["pu_mu"]) < 1e-10
    assert _relcov(g2["C"][1:], o2["pu_cov"][1:]) < 1e-10


@pytest.mark.parametrize("name,q,t1", [("lotka_volterra", 3, 10.0), ("rober", 3, 100.0), ("fitzhugh_nagumo", 3, 20.0),
                                       ("lotka_volterra", 5, 10.0), ("fitzhugh_nagumo", 2, 20.0)])
def test_adaptive_nonstiff_identical_step_sequences(oracle, name, q, t1):
    n = 24
    pd, u0, p = _inputs(name, n, 2, du0=0.05 if name != "rober" else 0.0)
    g = _gpu(pd, u0, p, q, builtin=(q == 3), adaptive=True, t0=0.0, t1=t1, save_everystep=True)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), q)
    cfg = oracle.make_config(pd.d, q, pd.n_p, alg=oracle.EK1, smooth=False, adaptive=True, t0=0.0, t1=t1, save_everystep=True,
                             cov_backend=oracle.COV_QR)
    for i in range(n):
        o = oracle.solve(cfg, rhs, u0[i], p[i])
        a, b = g["off"][i], g["off"][i + 1]
        assert g["naccept"][i] == o["naccept"] and g["nreject"][i] == o["nreject"] and g["nf"][i] == o["nf"]
        assert g["retcode"][i] == 0 and o["retcode"] == "Success"
        # identical accept/reject pattern; step times drift only by accumulated rounding through the controller
        # (measured: <= 4e-7 relative, oracle-vs-oracle floor 2e-9)
        dT = np.abs(g["T"][a:b] - o["t"])
        scale = np.abs(o["pu_mu"]).max(axis=0)
        e_t = np.max(dT[1:] / o["t"][1:])
        e_u = np.max(np.abs(g["u"][i] - o["pu_mu"][-1]) / scale)
        assert g["T"][b - 1] == t1
        if e_t >= 2e-6 or e_u >= 1e-8:
            # a trajectory that is sensitive to rounding (an error estimate close to 1 somewhere): compare with the
            # measured floor of the oracle against one-ulp-perturbed copies of itself
            kwo = dict(alg=oracle.EK1, smooth=False, adaptive=True, t0=0.0, t1=t1, save_everystep=True, cov_backend=oracle.COV_QR)
            f_t = rounding.floor(oracle, rhs, pd.d, q, pd.n_p, u0[i], p[i], kwo, o,
                                 lambda s_, b_: np.max(np.abs(s_["t"] - b_["t"])[1:] / b_["t"][1:]))
            f_u = rounding.floor(oracle, rhs, pd.d, q, pd.n_p, u0[i], p[i], kwo, o,
                                 lambda s_, b_: np.max(np.abs(s_["pu_mu"][-1] - b_["pu_mu"][-1]) / scale))
            assert e_t < max(2e-6, 20 * f_t) and e_u < max(1e-8, 20 * f_u), (i, e_t, f_t, e_u, f_u)
            continue
        # intermediate saved points sit at slightly different times: allow |u'| |dT| on top of 1e-8
        du = np.abs(np.gradient(o["pu_mu"], o["t"], axis=0)).max(axis=0) if len(o["t"]) > 2 else 0.0
        assert np.all(np.abs(g["U"][a:b] - o["pu_mu"]) <= 1e-8 * scale + 4.0 * du * dT[:, None] + 1e-300)
        # covariances carry sigma^2 (rounding-amplified) and are sampled at slightly different times
        kwo = dict(alg=oracle.EK1, smooth=False, adaptive=True, t0=0.0, t1=t1, save_everystep=True, cov_backend=oracle.COV_QR)
        f_c = rounding.floor(oracle, rhs, pd.d, q, pd.n_p, u0[i], p[i], kwo, o,
                             lambda s_, b_: _relcov(s_["pu_cov"][1:], b_["pu_cov"][1:]), ulp_components=[0, pd.d])
        e_c = _relcov(g["C"][a + 1:b], o["pu_cov"][1:])
        assert e_c < (max(1e-10, 20 * f_c) if f_c > 0 else 1e-3), (i, e_c, f_c)   # f_c == 0: every variant changed its step pattern


@pytest.mark.parametrize("name,q,t1", [("vanderpol", 5, 2.0), ("orego", 3, 30.0)])
def test_adaptive_stiff_step_teacher_forced(oracle, name, q, t1):
    """Stiff trajectories are chaotic in the rounding: free-running step sequences cannot be identical
    between any two implementations (not even the oracle's own Cholesky and QR back-ends).  So (i) the kernel replays
    the oracle's accepted grid (step-teacher-forced) and must stay within the oracle-vs-oracle rounding floor measured in
    the same test; (ii) with sigma^2 forced as well (pure linear-algebra parity) it must reach rtol 1e-8; (iii) the
    free-running kernel does the same amount of work and lands on the same solution to solver accuracy."""
    pd, u0, p = _inputs(name, 1, 3, dp=0.0)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), q)
    o = oracle.solve(oracle.make_config(pd.d, q, pd.n_p, alg=oracle.EK1, smooth=False, adaptive=True, t0=0.0, t1=t1,
                                        save_everystep=True, cov_backend=oracle.COV_QR), rhs, u0[0], p[0])
    grid = o["t"][1:-1]
    kw = dict(alg=oracle.EK1, smooth=False, adaptive=False, dt=t1, t0=0.0, t1=t1, save_everystep=True, tstops=grid)
    of = oracle.solve(oracle.make_config(pd.d, q, pd.n_p, cov_backend=oracle.COV_QR, **kw), rhs, u0[0], p[0])
    g = _gpu(pd, u0, p, q, builtin=False, adaptive=False, dt=t1, t0=0.0, t1=t1, save_everystep=True, tstops=grid)
    assert np.array_equal(g["T"], of["t"]) and np.array_equal(of["t"], o["t"])
    scale = np.abs(of["pu_mu"]).max(axis=0)

    def m_mean(s_, b_):
        return np.max(np.abs(s_["pu_mu"] - b_["pu_mu"]) / scale)

    kwq = dict(cov_backend=oracle.COV_QR, **kw)
    fl = rounding.floor(oracle, rhs, pd.d, q, pd.n_p, u0[0], p[0], kwq, of, m_mean)
    err = np.max(np.abs(g["U"] - of["pu_mu"]) / scale)
    assert err < max(1e-8, 10.0 * fl), (err, fl)
    # (ii) both forcings: sigma^2 no longer amplifies; what is left is the dynamics' own sensitivity
    sig = of["diffusions"][:-1]
    kwf = dict(forced_diffusion=sig, **kwq)
    of2 = oracle.solve(oracle.make_config(pd.d, q, pd.n_p, **kwf), rhs, u0[0], p[0])
    g2 = _gpu(pd, u0, p, q, builtin=False, adaptive=False, dt=t1, t0=0.0, t1=t1, save_everystep=True, tstops=grid,
              forced_diffusion=sig)
    fl2 = rounding.floor(oracle, rhs, pd.d, q, pd.n_p, u0[0], p[0], kwf, of2, m_mean)
    fc2 = rounding.floor(oracle, rhs, pd.d, q, pd.n_p, u0[0], p[0], kwf, of2,
                         lambda s_, b_: _relcov(s_["pu_cov"][1:], b_["pu_cov"][1:]))
    assert np.max(np.abs(g2["U"] - of2["pu_mu"]) / scale) < max(1e-8, 10.0 * fl2)
    assert _relcov(g2["C"][1:], of2["pu_cov"][1:]) < max(1e-8, 10.0 * fc2)
    # (iii) free running
    gf = _gpu(pd, u0, p, q, adaptive=True, t0=0.0, t1=t1)
    assert abs(int(gf["naccept"][0]) - o["naccept"]) <= 0.03 * o["naccept"]
    assert gf["retcode"][0] == 0
    assert np.max(np.abs(gf["u"][0] - o["pu_mu"][-1]) / scale) < 1e-3


def test_nvrtc_and_ahead_of_time_kernels_agree():
    pd, u0, p = _inputs("rober", 40, 4)
    a = _gpu(pd, u0, p, 3, builtin=True, t0=0.0, t1=100.0)
    b = _gpu(pd, u0, p, 3, builtin=False, t0=0.0, t1=100.0)
    assert np.array_equal(a["naccept"], b["naccept"])
    assert _rel(a["u"], b["u"]) < 1e-12


@pytest.mark.parametrize("lanes", [2, 4, 8])
def test_lane_group_sizes_agree(oracle, lanes):
    pd, u0, p = _inputs("lotka_volterra", 13, 5, du0=0.1)   # 13: not a multiple of any group count
    g = _gpu(pd, u0, p, 3, builtin=False, adaptive=True, t0=0.0, t1=10.0, lanes_per_traj=lanes)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    o = oracle.solve_ensemble(oracle.make_config(2, 3, 4, alg=oracle.EK1, smooth=False, adaptive=True, t0=0.0, t1=10.0,
                                                 save_everystep=False, cov_backend=oracle.COV_QR), rhs, u0, p)
    assert np.array_equal(g["naccept"], o["naccept"]) and np.array_equal(g["nreject"], o["nreject"])
    assert _rel(g["u"], o["u_final"]) < 1e-8
    assert g["info"].lanes_per_traj == lanes


def test_edge_cases(oracle):
    pd, u0, p = _inputs("lotka_volterra", 1, 6)
    # single trajectory
    g = _gpu(pd, u0, p, 3, t0=0.0, t1=10.0)
    assert g["retcode"][0] == 0 and g["t_end"][0] == 10.0
    # maxiters -> MaxIters retcode (sol.retcode), partial progress reported
    g = _gpu(pd, u0, p, 3, t0=0.0, t1=10.0, maxiters=10)
    assert lib.RETCODES[int(g["retcode"][0])] == "MaxIters" and g["t_end"][0] < 10.0
    # user-provided initial dt with adaptive stepping, and tstops inside the interval
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    ts = np.array([2.5, 5.0, 7.5])
    o = oracle.solve(oracle.make_config(2, 3, 4, alg=oracle.EK1, smooth=False, adaptive=True, dt=0.05, t0=0.0, t1=10.0, tstops=ts,
                                        cov_backend=oracle.COV_QR), rhs, u0[0], p[0])
    g = _gpu(pd, u0, p, 3, builtin=False, adaptive=True, dt=0.05, t0=0.0, t1=10.0, tstops=ts, save_everystep=True)
    assert g["naccept"][0] == o["naccept"] and g["nreject"][0] == o["nreject"]
    assert set(ts).issubset(set(g["T"]))
    assert _rel(g["u"][0], o["pu_mu"][-1]) < 1e-8
    # the host API mirror (solve(EnsembleProblem, EK1(), EnsembleB200()))
    from probnumdiffeq_jl_b200 import api

    prob = api.ODEProblem.from_library("lotka_volterra")
    ens = api.EnsembleProblem(prob, prob_func=lambda pr, i, rep: api.remake(pr, u0=[1.0 + 0.01 * i, 1.0]))
    sol = api.solve(ens, api.EK1(order=3, smooth=False), api.EnsembleB200(), trajectories=5)
    assert len(sol) == 5 and sol.converged
    assert all(len(s.t) == s.stats.naccept + 1 and s.t[-1] == 10.0 for s in sol.u)   # test/solution.jl:35-42
    assert np.array_equal(sol[0].u[0], [1.0, 1.0]) and not sol[0].pu_cov[0].any()      # test/solution.jl:59-62


@pytest.mark.parametrize("alg", [lib.EK0, lib.EK1])
def test_scalar_problem_convergence_orders(oracle, alg):
    """test/convergence.jl:19-46 on the device: d = 1 (the smallest state the kernels are instantiated for), u' = 1.01 u,
    orders 1..4, fixed steps 2^-7 ... 2^-4: convergence order q + 1 within 0.3 and agreement with the oracle."""
    def lin(du, u, p, t):
        du[0] = p[0] * u[0]

    tr = tracer.trace(lin, 1, 1)
    src = tracer.cuda_source(tr, "UserProb")
    u0, p = np.array([[0.5]]), np.array([[1.01]])
    for q in (1, 2, 3, 4):
        rhs = oracle.compile_rhs(tr, q)
        errs = []
        for k in (7, 6, 5, 4):
            dt = 0.5 ** k
            cfg = lib.make_config(d=1, q=q, n_p=1, rhs_name="UserProb", rhs_src=src, alg=alg, adaptive=False, dt=dt, t0=0.0, t1=1.0)
            s = lib.Solver(cfg)
            r = s.solve(u0, p)
            uf = r.field(lib.F_U_FINAL)[0]
            r.free()
            s.close()
            o = oracle.solve(oracle.make_config(1, q, 1, alg=alg, smooth=False, adaptive=False, dt=dt, t0=0, t1=1, save_everystep=False,
                                                cov_backend=oracle.COV_QR), rhs, [0.5], [1.01])
            assert abs(uf - o["pu_mu"][-1][0]) <= 1e-9 * abs(uf)
            errs.append(abs(uf - 0.5 * np.exp(1.01)))
        orders = [np.log2(errs[i + 1] / errs[i]) for i in range(3)]
        assert abs(np.mean(orders) - (q + 1)) < 0.3, (q, errs)


def test_full_size_properties_rober_1m():
    """BASELINE cfg5 at full size (2^20 trajectories): properties that need no oracle."""
    n = 1 << 20
    pd, u0, p = _inputs("rober", n, 3)
    g = _gpu(pd, u0, p, 3, t0=0.0, t1=100.0)
    assert np.all(g["retcode"] == 0) and np.all(g["t_end"] == 100.0)
    # ROBER conserves u1 + u2 + u3 = 1 (linear invariant: the filter mean preserves it up to the tolerance)
    assert np.max(np.abs(g["u"].sum(axis=1) - 1.0)) < 1e-6
    assert np.all(np.linalg.eigvalsh(g["cov"]) > -1e-18)
    # determinism + permutation equivariance: results do not depend on which warp / neighbours a trajectory gets
    perm = np.random.default_rng(0).permutation(n)
    g2 = _gpu(pd, u0[perm], p[perm], 3, t0=0.0, t1=100.0)
    assert np.array_equal(g2["naccept"], g["naccept"][perm])
    assert np.array_equal(g2["u"], g["u"][perm])
    assert np.array_equal(g2["cov"], g["cov"][perm])
    # identical inputs -> identical outputs
    u0c = np.tile(u0[:1], (4096, 1))
    pc = np.tile(p[:1], (4096, 1))
    g3 = _gpu(pd, u0c, pc, 3, t0=0.0, t1=100.0)
    assert np.all(g3["u"] == g3["u"][0]) and np.all(g3["naccept"] == g3["naccept"][0])
    assert np.array_equal(g3["u"][0], g["u"][0])


def test_fixed_steps_skip_the_counting_pass_and_fall_back_when_a_trajectory_fails():
    """Fixed-step solves compute the CSR offsets on the host (no counting pass).  If a trajectory stops early the solve is
    repeated with the counting pass: here the EK0 with dt = 0.02 is unstable on the stiff Van der Pol members (mu ~ 1e3)
    of a mixed ensemble (Unstable retcode after a few steps) while the mild ones (mu ~ 1) finish."""
    pd = problems.PROBLEMS["vanderpol"]
    n = 12
    u0 = np.tile(np.array(pd.u0, dtype=float), (n, 1))
    p = np.ones((n, 1))
    p[::3, 0] = 1e3
    src = tracer.cuda_source(tracer.trace(pd.f, pd.d, pd.n_p), "UserProb")

    def run(uu, pp):
        cfg = lib.make_config(d=2, q=3, n_p=1, rhs_name="UserProb", rhs_src=src, alg=lib.EK0, adaptive=False, dt=0.02, t0=0.0,
                              t1=4.0, save_everystep=True, smooth=True)
        s = lib.Solver(cfg)
        r = s.solve(uu, pp)
        out = dict(off=r.field(lib.F_STEP_OFFSETS), na=r.field(lib.F_NACCEPT), rc=r.field(lib.F_RETCODE), T=r.field(lib.F_T),
                   U=r.field(lib.F_U).reshape(-1, 2), nl=r.info.n_launches)
        r.free()
        s.close()
        return out

    g = run(u0, p)
    assert np.array_equal(np.diff(g["off"]), g["na"] + 1)                  # exact CSR also for the failed members
    bad = g["rc"] != 0
    assert bad[::3].all() and not bad[1::3].any() and not bad[2::3].any()
    ok = run(u0[~bad], p[~bad])                                            # no failure: single pass (one launch fewer)
    assert ok["nl"] < g["nl"] and (ok["na"] == 200).all()
    for j, i in enumerate(np.flatnonzero(~bad)):
        a, b = g["off"][i], g["off"][i + 1]
        a2, b2 = ok["off"][j], ok["off"][j + 1]
        assert np.array_equal(g["T"][a:b], ok["T"][a2:b2]) and np.array_equal(g["U"][a:b], ok["U"][a2:b2])


# ------------------------------------------------------------------------------------------------------------------
# smoother (BASELINE cfg1 / cfg3: filter + RTS smoother)
# ------------------------------------------------------------------------------------------------------------------
def _gpu_smooth(pd, u0, p, q, **kw):
    src = tracer.cuda_source(tracer.trace(pd.f, pd.d, pd.n_p), "UserProb")
    cfg = lib.make_config(d=pd.d, q=q, n_p=pd.n_p, rhs_name="UserProb", rhs_src=src, smooth=True, save_everystep=True,
                          save_state=True, **kw)
    s = lib.Solver(cfg)
    r = s.solve(u0, p)
    n, d = u0.shape
    D = d * (q + 1)
    out = dict(off=r.field(lib.F_STEP_OFFSETS), T=r.field(lib.F_T), U=r.field(lib.F_U).reshape(-1, d),
               C=r.field(lib.F_U_COV).reshape(-1, d, d), X=r.field(lib.F_X).reshape(-1, D),
               XR=r.field(lib.F_X_COVFAC).reshape(-1, D, D), Df=r.field(lib.F_DIFFUSIONS), naccept=r.field(lib.F_NACCEPT),
               u=r.field(lib.F_U_FINAL).reshape(n, d), info=r.info)
    r.free()
    s.close()
    return out


@pytest.mark.parametrize("calibrate", [True, False])
def test_cfg1_smoother_fixed_steps_static_diffusion(oracle, calibrate):
    """BASELINE cfg1 with FixedDiffusion: smoothed means and covariances of all 2001 points, rtol 1e-10."""
    pd, u0, p = _inputs("fitzhugh_nagumo", 2, 7, dp=0.05)
    g = _gpu_smooth(pd, u0, p, 3, adaptive=False, dt=0.01, t0=0.0, t1=20.0, diffusion=lib.DIFF_FIXED, calibrate=calibrate)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    for i in range(u0.shape[0]):
        o = oracle.solve(oracle.make_config(2, 3, 3, alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, calibrate=calibrate, smooth=True,
                                            adaptive=False, dt=0.01, t0=0.0, t1=20.0, cov_backend=oracle.COV_QR), rhs, u0[i], p[i])
        a, b = g["off"][i], g["off"][i + 1]
        assert b - a == o["n"] == 2001
        assert _rel(g["U"][a:b], o["pu_mu"]) < 1e-10
        assert _relcov(g["C"][a + 1:b], o["pu_cov"][1:]) < 1e-10
        assert not g["C"][a].any()                                   # exact initial value (test/solution.jl:59-62)
        # full smoothed state: means norm-wise on the u-block, covariances as R'R
        Sg = np.einsum("kra,krb->kab", g["XR"][a:b], g["XR"][a:b])
        So = np.einsum("kra,krb->kab", o["x_smooth_R"], o["x_smooth_R"])
        assert _relcov(Sg[1:], So[1:]) < 1e-9
        # higher-derivative blocks of the state are conditioned much worse than the u-block
        assert _rel(g["X"][a:b], o["x_smooth_mu"]) < 1e-6
        assert _rel(g["X"][a:b, :2], o["x_smooth_mu"][:, :2]) < 1e-10
        assert _rel(g["Df"][a:b - 1], o["diffusions"][:-1]) < 1e-10


def test_smoother_dynamic_diffusion_teacher_forced_and_free(oracle):
    pd, u0, p = _inputs("fitzhugh_nagumo", 1, 8, dp=0.0)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    kw = dict(alg=oracle.EK1, diffusion=oracle.DIFF_DYNAMIC, smooth=True, adaptive=False, dt=0.01, t0=0.0, t1=20.0)
    o = oracle.solve(oracle.make_config(2, 3, 3, cov_backend=oracle.COV_QR, **kw), rhs, u0[0], p[0])
    g = _gpu_smooth(pd, u0, p, 3, adaptive=False, dt=0.01, t0=0.0, t1=20.0)
    kwq = dict(cov_backend=oracle.COV_QR, **kw)
    f_mu = rounding.floor(oracle, rhs, 2, 3, 3, u0[0], p[0], kwq, o, lambda s_, b_: _rel(s_["pu_mu"], b_["pu_mu"]))
    f_cov = rounding.floor(oracle, rhs, 2, 3, 3, u0[0], p[0], kwq, o, lambda s_, b_: _relcov(s_["pu_cov"][1:], b_["pu_cov"][1:]))
    e_mu, e_cov = _rel(g["U"], o["pu_mu"]), _relcov(g["C"][1:], o["pu_cov"][1:])
    assert e_mu < max(1e-10, 20 * f_mu), (e_mu, f_mu)       # measured oracle-vs-oracle floor 5e-10
    assert e_cov < max(1e-10, 20 * f_cov), (e_cov, f_cov)   # floor 8e-7
    sig = o["diffusions"][:-1]
    o2 = oracle.solve(oracle.make_config(2, 3, 3, cov_backend=oracle.COV_QR, forced_diffusion=sig, **kw), rhs, u0[0], p[0])
    g2 = _gpu_smooth(pd, u0, p, 3, adaptive=False, dt=0.01, t0=0.0, t1=20.0, forced_diffusion=sig)
    # the forced sigma^2 sequence spans 40 orders of magnitude (its first entries are rounding noise, ~1e35), which
    # degrades the conditioning of the smoothing gain: hold the kernel to 1e-10 or to the oracle's own floor
    kwf = dict(cov_backend=oracle.COV_QR, forced_diffusion=sig, **kw)
    fm = rounding.floor(oracle, rhs, 2, 3, 3, u0[0], p[0], kwf, o2, lambda s_, b_: _rel(s_["pu_mu"], b_["pu_mu"]), ulp_components=[])
    fc = rounding.floor(oracle, rhs, 2, 3, 3, u0[0], p[0], kwf, o2, lambda s_, b_: _relcov(s_["pu_cov"][1:], b_["pu_cov"][1:]),
                        ulp_components=[])
    assert _rel(g2["U"], o2["pu_mu"]) < max(1e-10, 10 * fm), (fm,)
    assert _relcov(g2["C"][1:], o2["pu_cov"][1:]) < max(1e-10, 10 * fc), (fc,)


@pytest.mark.parametrize("name,q,t1", [("lotka_volterra", 3, 10.0), ("rober", 3, 100.0)])
def test_smoother_adaptive(oracle, name, q, t1):
    n = 6
    pd, u0, p = _inputs(name, n, 9, du0=0.05 if name != "rober" else 0.0)
    g = _gpu_smooth(pd, u0, p, q, adaptive=True, t0=0.0, t1=t1)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), q)
    for i in range(n):
        o = oracle.solve(oracle.make_config(pd.d, q, pd.n_p, alg=oracle.EK1, smooth=True, adaptive=True, t0=0.0, t1=t1,
                                            cov_backend=oracle.COV_QR), rhs, u0[i], p[i])
        a, b = g["off"][i], g["off"][i + 1]
        assert b - a == o["n"]
        scale = np.abs(o["pu_mu"]).max(axis=0)
        dT = np.abs(g["T"][a:b] - o["t"])
        du = np.abs(np.gradient(o["pu_mu"], o["t"], axis=0)).max(axis=0)
        assert np.all(np.abs(g["U"][a:b] - o["pu_mu"]) <= 1e-7 * scale + 4.0 * du * dT[:, None])
        assert np.max(np.abs(g["U"][b - 1] - o["pu_mu"][-1]) / scale) < 1e-8


def test_smoother_register_and_shared_memory_sweeps_agree(monkeypatch):
    """The register-resident sweep (pn_smooth_reg.cuh: one joint QR, G never formed) and the shared-memory sweep
    (pn_smooth.cuh: the reference formulas literally) smooth the same filtered records: ragged adaptive ensemble
    (idle groups, different lengths per warp), BASELINE cfg3 shape (Van der Pol, order 5, D = 12)."""
    n = 203
    pd, u0, p = _inputs("vanderpol", n, 21, dp=0.2)
    u0[:, 0] += 0.1 * np.random.default_rng(5).uniform(-1, 1, n)
    g = _gpu_smooth(pd, u0, p, 5, adaptive=True, t0=0.0, t1=0.5)
    monkeypatch.setenv("PNODE_SMOOTH_GENERIC", "1")
    r = _gpu_smooth(pd, u0, p, 5, adaptive=True, t0=0.0, t1=0.5)
    assert np.array_equal(g["off"], r["off"]) and np.array_equal(g["T"], r["T"])
    assert len(set(np.diff(g["off"]))) > 3                          # ragged
    scale = np.abs(r["U"]).max(axis=0)
    assert np.max(np.abs(g["U"] - r["U"]) / scale) < 1e-9
    assert _relcov(g["C"], r["C"]) < 1e-6
    # (full-state covariances of this stiff order-5 problem have entries up to 1e20 in the high-derivative blocks;
    #  they are compared on the non-stiff problem below)
    monkeypatch.delenv("PNODE_SMOOTH_GENERIC")
    pd, u0, p = _inputs("lotka_volterra", 37, 22, du0=0.05)
    g = _gpu_smooth(pd, u0, p, 3, adaptive=True, t0=0.0, t1=10.0)
    monkeypatch.setenv("PNODE_SMOOTH_GENERIC", "1")
    r = _gpu_smooth(pd, u0, p, 3, adaptive=True, t0=0.0, t1=10.0)
    assert np.array_equal(g["off"], r["off"])
    assert _rel(g["U"], r["U"]) < 1e-10 and _relcov(g["C"], r["C"]) < 1e-9
    Sg = np.einsum("kra,krb->kab", g["XR"], g["XR"])
    Sr = np.einsum("kra,krb->kab", r["XR"], r["XR"])
    assert _relcov(Sg, Sr) < 1e-8
    assert _rel(g["X"], r["X"]) < 1e-8


@pytest.mark.parametrize("smooth,diffusion", [(False, lib.DIFF_DYNAMIC), (True, lib.DIFF_DYNAMIC), (True, lib.DIFF_FIXED)])
def test_pilot_sized_single_save_pass_equals_two_pass(monkeypatch, smooth, diffusion):
    """Adaptive save_everystep solves of >= 1024 trajectories size their segments from a pilot run and do ONE full pass;
    the result must be bit-identical to the counting-pass layout (PNODE_TWO_PASS=1), also when the pilot's estimate is too
    small and the solve falls back (PNODE_PILOT_MARGIN_PCT=50), and with dense output on a saveat grid."""
    n = 2048
    pd, u0, p = _inputs("lotka_volterra", n, 61, du0=0.3, dp=0.3)
    saveat = np.linspace(0.0, 10.0, 5)

    def run():
        cfg = lib.make_config(d=2, q=3, n_p=4, rhs_name="builtin:lotka_volterra", adaptive=True, t0=0.0, t1=10.0, smooth=smooth,
                              save_everystep=True, saveat=saveat, diffusion=diffusion)
        s = lib.Solver(cfg)
        r = s.solve(u0, p)
        out = {f: r.field(f) for f in (lib.F_STEP_OFFSETS, lib.F_T, lib.F_U, lib.F_U_COV, lib.F_DIFFUSIONS, lib.F_NACCEPT,
                                       lib.F_U_FINAL, lib.F_SAVEAT_U, lib.F_SAVEAT_U_COV)}
        out["launches"] = r.info.n_launches
        out["saved"] = r.info.total_saved
        r.free()
        s.close()
        return out

    a = run()
    monkeypatch.setenv("PNODE_TWO_PASS", "1")
    b = run()
    monkeypatch.delenv("PNODE_TWO_PASS")
    monkeypatch.setenv("PNODE_PILOT_MARGIN_PCT", "50")
    c = run()
    assert len(set(np.diff(a[lib.F_STEP_OFFSETS]))) > 5                     # ragged
    assert a["saved"] == b["saved"] == c["saved"] == int(a[lib.F_NACCEPT].sum()) + n
    used = np.ones(a["saved"], dtype=bool)
    used[a[lib.F_STEP_OFFSETS][1:] - 1] = False      # sol.diffusions has one entry per STEP: the last slot of a segment is unused
    for other in (b, c):
        for f in a:
            if f == lib.F_DIFFUSIONS:
                assert np.array_equal(a[f][used], other[f][used])
            elif f not in ("launches", "saved"):
                assert np.array_equal(a[f], other[f]), f
    assert c["launches"] == b["launches"] != a["launches"]                   # the fallback ended on the counting-pass layout


def test_smoother_properties_at_scale():
    """BASELINE cfg3 shape (Van der Pol mu ~ 1e3, EK1 order 5, adaptive) on 20 000 trajectories / 11.6 M saved points:
    size-independent properties of the RTS smoother -- identical time grids and final point with and without smoothing,
    and a smoothed marginal variance that never exceeds the filtered one (conditioning on more data)."""
    n = 20000
    pd = problems.PROBLEMS["vanderpol"]
    rng = np.random.default_rng(1)
    u0 = np.tile(np.array(pd.u0, dtype=float), (n, 1))
    u0[:, 0] += 0.1 * rng.uniform(-1, 1, n)
    p = np.array(pd.p, dtype=float)[None, :] * (1 + 0.2 * rng.uniform(-1, 1, (n, 1)))
    out = {}
    for smooth in (False, True):
        cfg = lib.make_config(d=2, q=5, n_p=1, rhs_name="builtin:vanderpol", adaptive=True, t0=0.0, t1=2.0, smooth=smooth,
                              save_everystep=True)
        s = lib.Solver(cfg)
        r = s.solve(u0, p)
        out[smooth] = dict(off=r.field(lib.F_STEP_OFFSETS), T=r.field(lib.F_T), U=r.field(lib.F_U).reshape(-1, 2),
                           C=r.field(lib.F_U_COV).reshape(-1, 2, 2), rc=r.field(lib.F_RETCODE))
        r.free()
        s.close()
    f, sm = out[False], out[True]
    assert (f["rc"] == 0).all() and (sm["rc"] == 0).all()
    assert np.array_equal(f["off"], sm["off"]) and np.array_equal(f["T"], sm["T"])
    last = f["off"][1:] - 1
    assert np.array_equal(f["U"][last], sm["U"][last])                       # x_smooth[N] = x_filt[N]
    assert np.allclose(f["C"][last], sm["C"][last], rtol=1e-13, atol=0)
    vf = np.einsum("kii->ki", f["C"])
    vs = np.einsum("kii->ki", sm["C"])
    assert np.isfinite(vs).all() and (vs >= 0).all()
    assert np.all(vs <= vf * (1 + 1e-6) + 1e-300)
    interior = np.ones(len(vf), dtype=bool)
    interior[last] = False
    interior[f["off"][:-1]] = False
    assert np.mean(vs[interior] < vf[interior]) > 0.99                        # and it is a strict improvement almost everywhere


# ------------------------------------------------------------------------------------------------------------------
# EK0 / IsometricKronecker layout (BASELINE cfg2)
# ------------------------------------------------------------------------------------------------------------------
def _gpu_ek0(pd, u0, p, q, builtin=False, **kw):
    if builtin:
        cfg = lib.make_config(d=pd.d, q=q, n_p=pd.n_p, rhs_name="builtin:" + pd.name, alg=lib.EK0, **kw)
    else:
        src = tracer.cuda_source(tracer.trace(pd.f, pd.d, pd.n_p), "UserProb")
        cfg = lib.make_config(d=pd.d, q=q, n_p=pd.n_p, rhs_name="UserProb", rhs_src=src, alg=lib.EK0, **kw)
    s = lib.Solver(cfg)
    r = s.solve(u0, p)
    n, d = u0.shape
    out = dict(u=r.field(lib.F_U_FINAL).reshape(n, d), cov=r.field(lib.F_U_FINAL_COV).reshape(n, d, d),
               naccept=r.field(lib.F_NACCEPT), nreject=r.field(lib.F_NREJECT), nf=r.field(lib.F_NF),
               retcode=r.field(lib.F_RETCODE), loglik=r.field(lib.F_LOGLIK), diffusion=r.field(lib.F_DIFFUSION), info=r.info)
    if r.has(lib.F_STEP_OFFSETS):
        out.update(off=r.field(lib.F_STEP_OFFSETS), T=r.field(lib.F_T), U=r.field(lib.F_U).reshape(-1, d),
                   C=r.field(lib.F_U_COV).reshape(-1, d, d), Df=r.field(lib.F_DIFFUSIONS))
    if r.has(lib.F_X_FINAL):
        D = d * (q + 1)
        out.update(x=r.field(lib.F_X_FINAL).reshape(n, D), xR=r.field(lib.F_X_FINAL_COVFAC).reshape(n, D, D))
    r.free()
    s.close()
    return out


@pytest.mark.parametrize("diffusion,calibrate", [(lib.DIFF_FIXED, True), (lib.DIFF_FIXED, False)])
def test_ek0_fixed_steps_static_diffusion(oracle, diffusion, calibrate):
    pd, u0, p = _inputs("lotka_volterra", 3, 11, du0=0.1)
    g = _gpu_ek0(pd, u0, p, 3, adaptive=False, dt=0.01, t0=0.0, t1=10.0, diffusion=diffusion, calibrate=calibrate,
                 save_everystep=True, save_state=True)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    for i in range(3):
        o = oracle.solve(oracle.make_config(2, 3, 4, alg=oracle.EK0, diffusion=oracle.DIFF_FIXED, calibrate=calibrate, smooth=False,
                                            adaptive=False, dt=0.01, t0=0.0, t1=10.0, save_everystep=True,
                                            cov_backend=oracle.COV_QR), rhs, u0[i], p[i])
        a, b = g["off"][i], g["off"][i + 1]
        assert b - a == o["n"]
        assert np.array_equal(g["T"][a:b], o["t"])
        assert _rel(g["U"][a:b], o["pu_mu"]) < 1e-10
        assert _relcov(g["C"][a + 1:b], o["pu_cov"][1:]) < 1e-10
        assert g["loglik"][i] == pytest.approx(o["loglik"], rel=1e-10)      # incl. the logdet quirk (update.jl:147)
        assert g["diffusion"][i] == pytest.approx(o["diffusions"][0], rel=1e-10)
        Sg = g["xR"][i].T @ g["xR"][i]
        So = o["x_filt_R"][-1].T @ o["x_filt_R"][-1]
        assert np.linalg.norm(Sg - So) / np.linalg.norm(So) < 1e-10           # kron(R_B, I_d) dense form


@pytest.mark.parametrize("q", [2, 3, 5])
def test_ek0_adaptive_identical_step_sequences(oracle, q):
    """BASELINE cfg2 (small ensemble): Lotka-Volterra EK0, adaptive, filter only."""
    n = 40
    pd, u0, p = _inputs("lotka_volterra", n, 12, du0=0.1)
    g = _gpu_ek0(pd, u0, p, q, builtin=(q == 3), adaptive=True, t0=0.0, t1=10.0)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), q)
    o = oracle.solve_ensemble(oracle.make_config(2, q, 4, alg=oracle.EK0, smooth=False, adaptive=True, t0=0.0, t1=10.0,
                                                 save_everystep=False, cov_backend=oracle.COV_QR), rhs, u0, p)
    same = (g["naccept"] == o["naccept"]) & (g["nreject"] == o["nreject"]) & (g["nf"] == o["nf"])
    assert np.all(g["retcode"] == 0)
    if q <= 3:
        assert same.all()
    else:
        # EK0(order=5) takes >400 steps with ~10 % rejections on this problem; a decision with EEst within rounding of 1
        # flips in about one trajectory out of 40 (it also flips between the oracle's own Cholesky and QR back-ends)
        assert same.mean() >= 0.9
        assert np.max(np.abs(g["naccept"] - o["naccept"])) <= 2
        assert _rel(g["u"], o["u_final"]) < 1e-5
    assert _rel(g["u"][same], o["u_final"][same]) < (1e-8 if q <= 3 else 1e-5)
    assert _relcov(g["cov"][same], o["cov_final"][same]) < (1e-4 if q <= 3 else 1e-2)
    # isotropic covariance of the Kronecker layout
    assert np.allclose(g["cov"][:, 0, 1], 0.0) and np.allclose(g["cov"][:, 0, 0], g["cov"][:, 1, 1], rtol=0, atol=0)


def test_ek0_smoother(oracle):
    pd, u0, p = _inputs("lotka_volterra", 2, 13, du0=0.1)
    src = tracer.cuda_source(tracer.trace(pd.f, pd.d, pd.n_p), "UserProb")
    cfg = lib.make_config(d=2, q=3, n_p=4, rhs_name="UserProb", rhs_src=src, alg=lib.EK0, smooth=True, save_everystep=True,
                          adaptive=False, dt=0.02, t0=0.0, t1=10.0, diffusion=lib.DIFF_FIXED)
    s = lib.Solver(cfg)
    r = s.solve(u0, p)
    off, U, Cv = r.field(lib.F_STEP_OFFSETS), r.field(lib.F_U).reshape(-1, 2), r.field(lib.F_U_COV).reshape(-1, 2, 2)
    r.free()
    s.close()
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    for i in range(2):
        o = oracle.solve(oracle.make_config(2, 3, 4, alg=oracle.EK0, diffusion=oracle.DIFF_FIXED, smooth=True, adaptive=False, dt=0.02,
                                            t0=0.0, t1=10.0, cov_backend=oracle.COV_QR), rhs, u0[i], p[i])
        a, b = off[i], off[i + 1]
        assert _rel(U[a:b], o["pu_mu"]) < 1e-10
        assert _relcov(Cv[a + 1:b], o["pu_cov"][1:]) < 1e-9


# ------------------------------------------------------------------------------------------------------------------
# BlocksOfDiagonals layout: EK0 + multivariate diffusion, DiagonalEK1 (src/algorithms.jl:132-151)
# ------------------------------------------------------------------------------------------------------------------
def _gpu_blocks(pd, u0, p, q, alg, diffusion, **kw):
    src = tracer.cuda_source(tracer.trace(pd.f, pd.d, pd.n_p), "UserProb")
    cfg = lib.make_config(d=pd.d, q=q, n_p=pd.n_p, rhs_name="UserProb", rhs_src=src, alg=alg, diffusion=diffusion,
                          save_everystep=True, save_state=True, **kw)
    s = lib.Solver(cfg)
    r = s.solve(u0, p)
    n, d = u0.shape
    D = d * (q + 1)
    out = dict(off=r.field(lib.F_STEP_OFFSETS), T=r.field(lib.F_T), U=r.field(lib.F_U).reshape(-1, d),
               C=r.field(lib.F_U_COV).reshape(-1, d, d), naccept=r.field(lib.F_NACCEPT), nreject=r.field(lib.F_NREJECT),
               loglik=r.field(lib.F_LOGLIK), retcode=r.field(lib.F_RETCODE), xR=r.field(lib.F_X_FINAL_COVFAC).reshape(n, D, D),
               x=r.field(lib.F_X_FINAL).reshape(n, D), cov=r.field(lib.F_U_FINAL_COV).reshape(n, d, d), Df=r.field(lib.F_DIFFUSIONS))
    if r.has(lib.F_DIFFUSIONS_MV):
        out.update(DfMV=r.field(lib.F_DIFFUSIONS_MV).reshape(-1, d), dMV=r.field(lib.F_DIFFUSION_MV).reshape(n, d))
    if r.has(lib.F_X):
        out.update(X=r.field(lib.F_X).reshape(-1, D), XR=r.field(lib.F_X_COVFAC).reshape(-1, D, D))
    r.free()
    s.close()
    return out


_BLOCKS_STATIC = [(lib.EK0, lib.DIFF_FIXED_MV, True), (lib.EK0, lib.DIFF_FIXED_MV, False), (lib.DIAGONAL_EK1, lib.DIFF_FIXED, True),
                  (lib.DIAGONAL_EK1, lib.DIFF_FIXED, False), (lib.DIAGONAL_EK1, lib.DIFF_FIXED_MV, True)]


@pytest.mark.parametrize("alg,diffusion,calibrate", _BLOCKS_STATIC)
@pytest.mark.parametrize("smooth", [False, True])
def test_blocks_fixed_steps_static_diffusion(oracle, alg, diffusion, calibrate, smooth):
    """Static diffusion models on the BlocksOfDiagonals layout: every mean, covariance and diffusion of all 1001
    points to 1e-10 (filter) / 1e-9 (smoother covariances), FitzHugh-Nagumo, dt = 0.01."""
    pd, u0, p = _inputs("fitzhugh_nagumo", 2, 31, dp=0.05)
    g = _gpu_blocks(pd, u0, p, 3, alg, diffusion, calibrate=calibrate, smooth=smooth, adaptive=False, dt=0.01, t0=0.0, t1=10.0)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    for i in range(u0.shape[0]):
        o = oracle.solve(oracle.make_config(2, 3, 3, alg=alg, diffusion=diffusion, calibrate=calibrate, smooth=smooth, adaptive=False,
                                            dt=0.01, t0=0.0, t1=10.0, cov_backend=oracle.COV_QR), rhs, u0[i], p[i])
        a, b = g["off"][i], g["off"][i + 1]
        assert b - a == o["n"] == 1001 and g["retcode"][i] == 0
        assert np.array_equal(g["T"][a:b], o["t"])
        assert _rel(g["U"][a:b], o["pu_mu"]) < 1e-10
        assert _relcov(g["C"][a + 1:b], o["pu_cov"][1:]) < (1e-9 if smooth else 1e-10)
        assert g["loglik"][i] == pytest.approx(o["loglik"], rel=1e-10)
        if "DfMV" in g:
            assert _rel(g["DfMV"][a:b - 1], o["diffusions_mv"][:-1]) < 1e-10
            assert _rel(g["dMV"][i], o["diffusions_mv"][0]) < 1e-10
        else:
            assert _rel(g["Df"][a:b - 1], o["diffusions"][:-1]) < 1e-10
        if smooth:
            Sg = np.einsum("kra,krb->kab", g["XR"][a:b], g["XR"][a:b])
            So = np.einsum("kra,krb->kab", o["x_smooth_R"], o["x_smooth_R"])
            assert _relcov(Sg[1:], So[1:]) < 1e-9
            assert _rel(g["X"][a:b, :2], o["x_smooth_mu"][:, :2]) < 1e-10
        else:
            Sg = g["xR"][i].T @ g["xR"][i]
            So = o["x_filt_R"][-1].T @ o["x_filt_R"][-1]
            assert np.linalg.norm(Sg - So) / np.linalg.norm(So) < 1e-10


@pytest.mark.parametrize("alg,diffusion", [(lib.EK0, lib.DIFF_DYNAMIC_MV), (lib.DIAGONAL_EK1, lib.DIFF_DYNAMIC),
                                           (lib.DIAGONAL_EK1, lib.DIFF_DYNAMIC_MV), (lib.EK0, lib.DIFF_FIXED_MV)])
def test_blocks_adaptive_identical_step_sequences(oracle, alg, diffusion):
    """Adaptive runs on the Blocks layout: identical accept/reject counts, step times to 1e-7, solutions rtol 1e-8."""
    n = 12
    pd, u0, p = _inputs("lotka_volterra", n, 33, du0=0.1)
    g = _gpu_blocks(pd, u0, p, 3, alg, diffusion, adaptive=True, t0=0.0, t1=10.0)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    for i in range(n):
        o = oracle.solve(oracle.make_config(2, 3, 4, alg=alg, diffusion=diffusion, smooth=False, adaptive=True, t0=0.0, t1=10.0,
                                            cov_backend=oracle.COV_QR), rhs, u0[i], p[i])
        a, b = g["off"][i], g["off"][i + 1]
        assert g["naccept"][i] == o["naccept"] and g["nreject"][i] == o["nreject"]
        tol_t, tol_u = 1e-7, 1e-8
        if diffusion == lib.DIFF_DYNAMIC_MV:
            # sigma_i^2 = z_i^2 / HQH_i has no averaging over components: a residual component near a zero crossing moves
            # its sigma_i^2 -- and through the error estimate the step size -- by O(1) per ulp.  Hold the kernel to the
            # oracle's own rounding floor (fma / Cholesky-vs-QR / one-ulp variants with the same step pattern).
            kwf = dict(alg=alg, diffusion=diffusion, smooth=False, adaptive=True, t0=0.0, t1=10.0, cov_backend=oracle.COV_QR)
            ft = rounding.floor(oracle, rhs, 2, 3, 4, u0[i], p[i], kwf, o,
                                lambda s_, b_: np.max(np.abs(s_["t"] - b_["t"]) / np.maximum(b_["t"], 1e-3)), ulp_components=[0, 3])
            fu = rounding.floor(oracle, rhs, 2, 3, 4, u0[i], p[i], kwf, o, lambda s_, b_: _rel(s_["pu_mu"][-1], b_["pu_mu"][-1]),
                                ulp_components=[0, 3])
            tol_t, tol_u = max(tol_t, 20 * ft), max(tol_u, 20 * fu)
        assert np.max(np.abs(g["T"][a:b] - o["t"]) / np.maximum(o["t"], 1e-3)) < tol_t
        assert _rel(g["U"][b - 1], o["pu_mu"][-1]) < tol_u


# ------------------------------------------------------------------------------------------------------------------
# IOUP prior with a scalar rate parameter (src/priors/ioup.jl; test/correctness.jl:50-51,82-83 use IOUP(3, -1))
# ------------------------------------------------------------------------------------------------------------------
def _gpu_ioup(pd, u0, p, q, alg, rate, **kw):
    src = tracer.cuda_source(tracer.trace(pd.f, pd.d, pd.n_p), "UserProb")
    cfg = lib.make_config(d=pd.d, q=q, n_p=pd.n_p, rhs_name="UserProb", rhs_src=src, alg=alg, prior=lib.PRIOR_IOUP,
                          rate_parameter=rate, save_everystep=True, save_state=True, **kw)
    s = lib.Solver(cfg)
    r = s.solve(u0, p)
    n, d = u0.shape
    D = d * (q + 1)
    out = dict(off=r.field(lib.F_STEP_OFFSETS), T=r.field(lib.F_T), U=r.field(lib.F_U).reshape(-1, d),
               C=r.field(lib.F_U_COV).reshape(-1, d, d), naccept=r.field(lib.F_NACCEPT), nreject=r.field(lib.F_NREJECT),
               loglik=r.field(lib.F_LOGLIK), retcode=r.field(lib.F_RETCODE), Df=r.field(lib.F_DIFFUSIONS))
    if r.has(lib.F_X):
        out.update(X=r.field(lib.F_X).reshape(-1, D), XR=r.field(lib.F_X_COVFAC).reshape(-1, D, D))
    r.free()
    s.close()
    return out


@pytest.mark.parametrize("alg", [lib.EK0, lib.EK1])
@pytest.mark.parametrize("smooth", [False, True])
def test_ioup_fixed_steps_static_diffusion(oracle, alg, smooth):
    """IOUP(3, -1) with FixedDiffusion, fixed steps: the per-step transition (Gauss-Legendre process noise + series /
    recurrence for the phi functions on the device, matrix-fraction decomposition with a Pade exponential in the oracle)
    and the dense EK0 measurement.  Means 1e-10; covariances 1e-9 (two independent routes to Qh)."""
    pd, u0, p = _inputs("lotka_volterra", 2, 41, du0=0.05)
    g = _gpu_ioup(pd, u0, p, 3, alg, -1.0, diffusion=lib.DIFF_FIXED, smooth=smooth, adaptive=False, dt=0.01, t0=0.0, t1=5.0)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    for i in range(u0.shape[0]):
        o = oracle.solve(oracle.make_config(2, 3, 4, alg=alg, diffusion=oracle.DIFF_FIXED, smooth=smooth, adaptive=False, dt=0.01,
                                            t0=0.0, t1=5.0, cov_backend=oracle.COV_QR, prior=oracle.PRIOR_IOUP, rate_parameter=-1.0),
                         rhs, u0[i], p[i])
        a, b = g["off"][i], g["off"][i + 1]
        assert b - a == o["n"] == 501 and g["retcode"][i] == 0
        assert _rel(g["U"][a:b], o["pu_mu"]) < 1e-10
        assert _relcov(g["C"][a + 1:b], o["pu_cov"][1:]) < 1e-9
        assert g["loglik"][i] == pytest.approx(o["loglik"], rel=1e-9)
        assert _rel(g["Df"][a:b - 1], o["diffusions"][:-1]) < 1e-9
        if smooth:
            Sg = np.einsum("kra,krb->kab", g["XR"][a:b], g["XR"][a:b])
            So = np.einsum("kra,krb->kab", o["x_smooth_R"], o["x_smooth_R"])
            assert _relcov(Sg[1:], So[1:]) < 1e-8
            assert _rel(g["X"][a:b, :2], o["x_smooth_mu"][:, :2]) < 1e-10


@pytest.mark.parametrize("alg", [lib.EK0, lib.EK1])
def test_ioup_adaptive_identical_step_sequences(oracle, alg):
    n = 10
    pd, u0, p = _inputs("lotka_volterra", n, 43, du0=0.1)
    g = _gpu_ioup(pd, u0, p, 3, alg, -1.0, adaptive=True, t0=0.0, t1=10.0)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    for i in range(n):
        o = oracle.solve(oracle.make_config(2, 3, 4, alg=alg, smooth=False, adaptive=True, t0=0.0, t1=10.0, cov_backend=oracle.COV_QR,
                                            prior=oracle.PRIOR_IOUP, rate_parameter=-1.0), rhs, u0[i], p[i])
        a, b = g["off"][i], g["off"][i + 1]
        assert g["naccept"][i] == o["naccept"] and g["nreject"][i] == o["nreject"]
        assert np.max(np.abs(g["T"][a:b] - o["t"]) / np.maximum(o["t"], 1e-3)) < 1e-7
        assert _rel(g["U"][b - 1], o["pu_mu"][-1]) < 1e-8


def test_ioup_rate_zero_is_the_iwp():
    """IOUP with rate 0 is the IWP (to_sde, src/priors/ioup.jl:66-79): the two code paths of the kernel agree."""
    pd, u0, p = _inputs("lotka_volterra", 3, 45, du0=0.05)
    a = _gpu_ioup(pd, u0, p, 3, lib.EK1, 0.0, diffusion=lib.DIFF_FIXED, adaptive=False, dt=0.02, t0=0.0, t1=4.0)
    b = _gpu(pd, u0, p, 3, builtin=False, diffusion=lib.DIFF_FIXED, adaptive=False, dt=0.02, t0=0.0, t1=4.0, save_everystep=True)
    assert _rel(a["U"], b["U"]) < 1e-11
    assert _relcov(a["C"][1:], b["C"][1:]) < 1e-9


# ------------------------------------------------------------------------------------------------------------------
# dense output sol(t) on a saveat grid (src/solution.jl:330-398)
# ------------------------------------------------------------------------------------------------------------------
def _gpu_dense(pd, u0, p, q, saveat, builtin=False, **kw):
    src = tracer.cuda_source(tracer.trace(pd.f, pd.d, pd.n_p), "UserProb")
    cfg = lib.make_config(d=pd.d, q=q, n_p=pd.n_p, rhs_name="UserProb", rhs_src=src, save_everystep=True, saveat=saveat, **kw)
    s = lib.Solver(cfg)
    r = s.solve(u0, p)
    n, d = u0.shape
    out = dict(U=r.field(lib.F_SAVEAT_U).reshape(n, len(saveat), d), C=r.field(lib.F_SAVEAT_U_COV).reshape(n, len(saveat), d, d),
               off=r.field(lib.F_STEP_OFFSETS), T=r.field(lib.F_T), info=r.info)
    r.free()
    s.close()
    return out


@pytest.mark.parametrize("alg,smooth,diffusion,calibrate", [(lib.EK1, True, lib.DIFF_FIXED, True), (lib.EK1, False, lib.DIFF_FIXED, True),
                                                            (lib.EK1, True, lib.DIFF_FIXED, False), (lib.EK0, True, lib.DIFF_FIXED, True),
                                                            (lib.EK1, True, lib.DIFF_DYNAMIC, True)])
def test_dense_output_on_saveat_grid(oracle, alg, smooth, diffusion, calibrate):
    """sol(saveat) from the device against the oracle's `interpolate` on the same trajectory: grid points inside steps,
    exactly on saved points (t = 0.4 is a multiple of dt), at t0 and at t1.  Static diffusion: means 1e-10,
    covariances 1e-8; DynamicDiffusion: the rounding floor of SURVEY.md H0 (1e-8 / 1e-4)."""
    pd, u0, p = _inputs("lotka_volterra", 2, 51, du0=0.05)
    saveat = np.array([0.0, 0.013, 0.4, 0.77777, 1.23456, 2.0])
    g = _gpu_dense(pd, u0, p, 3, saveat, alg=alg, smooth=smooth, diffusion=diffusion, calibrate=calibrate, adaptive=False, dt=0.02,
                   t0=0.0, t1=2.0)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    static = diffusion == lib.DIFF_FIXED
    for i in range(u0.shape[0]):
        cfg = oracle.make_config(2, 3, 4, alg=alg, diffusion=diffusion, calibrate=calibrate, smooth=smooth, adaptive=False, dt=0.02,
                                 t0=0.0, t1=2.0, cov_backend=oracle.COV_QR)
        o = oracle.solve(cfg, rhs, u0[i], p[i])
        for j, t in enumerate(saveat):
            m, c = oracle.interpolate(cfg, o, float(t))
            assert _rel(g["U"][i, j], m) < (1e-10 if static else 1e-8), (i, j)
            if np.linalg.norm(c) > 0:
                assert np.linalg.norm(g["C"][i, j] - c) / np.linalg.norm(c) < (1e-8 if static else 1e-4), (i, j)
            else:
                assert not g["C"][i, j].any()


@pytest.mark.parametrize("alg,smooth,diffusion,calibrate", [(lib.EK0, True, lib.DIFF_FIXED_MV, True), (lib.EK0, False, lib.DIFF_FIXED_MV, True),
                                                            (lib.DIAGONAL_EK1, True, lib.DIFF_FIXED_MV, False), (lib.EK0, True, lib.DIFF_DYNAMIC_MV, True)])
def test_dense_output_multivariate_diffusions(oracle, alg, smooth, diffusion, calibrate):
    """sol(saveat) with the multivariate diffusion models (blocks layout): `diffusions[idx]` is a Diagonal and scales block i of Qh
    (src/solution.jl:351, apply_diffusion on BlocksOfDiagonals); FixedMVDiffusion: 1e-10 / 1e-8, DynamicMVDiffusion: the rounding floor
    of the scalar dynamic case (1e-8 / 1e-4)."""
    pd, u0, p = _inputs("lotka_volterra", 2, 52, du0=0.05)
    saveat = np.array([0.0, 0.013, 0.4, 0.77777, 1.23456, 2.0])
    g = _gpu_dense(pd, u0, p, 3, saveat, alg=alg, smooth=smooth, diffusion=diffusion, calibrate=calibrate, adaptive=False, dt=0.02,
                   t0=0.0, t1=2.0)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    static = diffusion == lib.DIFF_FIXED_MV
    for i in range(u0.shape[0]):
        cfg = oracle.make_config(2, 3, 4, alg=alg, diffusion=diffusion, calibrate=calibrate, smooth=smooth, adaptive=False, dt=0.02,
                                 t0=0.0, t1=2.0, cov_backend=oracle.COV_QR)
        o = oracle.solve(cfg, rhs, u0[i], p[i])
        if calibrate or not static:
            assert not np.allclose(o["diffusions_mv"][1, 0], o["diffusions_mv"][1, 1], rtol=1e-3)   # the components really differ
        for j, t in enumerate(saveat):
            m, c = oracle.interpolate(cfg, o, float(t))
            assert _rel(g["U"][i, j], m) < (1e-10 if static else 1e-8), (i, j)
            if np.linalg.norm(c) > 0:
                assert np.linalg.norm(g["C"][i, j] - c) / np.linalg.norm(c) < (1e-8 if static else 1e-4), (i, j)
            else:
                assert not g["C"][i, j].any()


def test_dense_output_adaptive_ensemble(oracle):
    """Adaptive ensemble with ragged step counts, common saveat grid (what the NCCL ensemble statistics consume)."""
    n = 40
    pd, u0, p = _inputs("lotka_volterra", n, 53, du0=0.1)
    saveat = np.linspace(0.0, 10.0, 17)
    g = _gpu_dense(pd, u0, p, 3, saveat, alg=lib.EK1, smooth=True, adaptive=True, t0=0.0, t1=10.0)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    assert len(set(np.diff(g["off"]))) > 3
    for i in (0, 7, 39):
        cfg = oracle.make_config(2, 3, 4, alg=oracle.EK1, smooth=True, adaptive=True, t0=0.0, t1=10.0, cov_backend=oracle.COV_QR)
        o = oracle.solve(cfg, rhs, u0[i], p[i])
        want = np.array([oracle.interpolate(cfg, o, float(t))[0] for t in saveat])
        assert _rel(g["U"][i], want) < 1e-7


def test_full_size_properties_lv_ek0_1m():
    """BASELINE cfg2 at full size: 2^20 Lotka-Volterra trajectories, EK0(order=3), adaptive, filter only."""
    n = 1 << 20
    pd, u0, p = _inputs("lotka_volterra", n, 0, du0=0.1)
    g = _gpu_ek0(pd, u0, p, 3, builtin=True, adaptive=True, t0=0.0, t1=10.0)
    assert np.all(g["retcode"] == 0)
    assert 150 < g["naccept"].mean() < 300
    assert np.all(g["u"] > 0)                       # populations stay positive
    perm = np.random.default_rng(1).permutation(n)
    g2 = _gpu_ek0(pd, u0[perm], p[perm], 3, builtin=True, adaptive=True, t0=0.0, t1=10.0)
    assert np.array_equal(g2["u"], g["u"][perm]) and np.array_equal(g2["naccept"], g["naccept"][perm])
    print("EK0 cfg2 kernel ms", g["info"].kernel_ms, "steps/s", g["info"].total_accepted / (g["info"].kernel_ms * 1e-3))


# ------------------------------------------------------------------------------------------------------------------
# CTA-per-trajectory kernel (BASELINE cfg4: Pleiades d = 28, D = 112)
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("qr_only", [False, True])
def test_cta_kernel_matches_oracle_on_small_problems(oracle, monkeypatch, qr_only):
    """The shared-memory CTA kernel is size-generic: run it on small problems where the oracle is fast.  Both
    covariance-prediction routes of predict.jl:84-91 are covered: Gram + Cholesky (default) and the Householder
    fallback `triangularize!` (forced with the developer flag PN_CTA_QR_ONLY)."""
    if qr_only:
        monkeypatch.setenv("PNODE_NVRTC_FLAGS", "-DPN_CTA_QR_ONLY")
    pd, u0, p = _inputs("fitzhugh_nagumo", 2, 14, dp=0.05)
    g = _gpu(pd, u0, p, 3, builtin=False, adaptive=False, dt=0.01, t0=0.0, t1=5.0, diffusion=lib.DIFF_FIXED,
             save_everystep=True, save_state=True, lanes_per_traj=256)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    for i in range(2):
        o = oracle.solve(oracle.make_config(2, 3, 3, alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, smooth=False, adaptive=False, dt=0.01,
                                            t0=0.0, t1=5.0, save_everystep=True, cov_backend=oracle.COV_QR), rhs, u0[i], p[i])
        a, b = g["off"][i], g["off"][i + 1]
        assert b - a == o["n"]
        assert _rel(g["U"][a:b], o["pu_mu"]) < 1e-10
        assert _relcov(g["C"][a + 1:b], o["pu_cov"][1:]) < 1e-10
        assert g["loglik"][i] == pytest.approx(o["loglik"], rel=1e-10)
    pd, u0, p = _inputs("rober", 5, 15)
    g = _gpu(pd, u0, p, 3, builtin=False, adaptive=True, t0=0.0, t1=100.0, lanes_per_traj=256)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    o = oracle.solve_ensemble(oracle.make_config(3, 3, 3, alg=oracle.EK1, smooth=False, adaptive=True, t0=0.0, t1=100.0,
                                                 save_everystep=False, cov_backend=oracle.COV_QR), rhs, u0, p)
    assert np.array_equal(g["naccept"], o["naccept"]) and np.array_equal(g["nreject"], o["nreject"])
    assert _rel(g["u"], o["u_final"]) < 1e-8


def test_cfg4_pleiades_cta_kernel(oracle):
    """Pleiades first-order form (benchmarks/pleiades.jmd:35-61), EK1(order=3), D = 112, adaptive, filter only."""
    n = 3
    pd, u0, p = _inputs("pleiades", n, 2)
    rng = np.random.default_rng(2)
    u0[:, 14:] += 1e-3 * rng.uniform(-1, 1, (n, 14))      # perturb the initial positions
    p = np.zeros((n, 0))
    g = _gpu(pd, u0, p, 3, builtin=True, adaptive=True, t0=0.0, t1=3.0, save_state=True)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    cfg = oracle.make_config(28, 3, 0, alg=oracle.EK1, smooth=False, adaptive=True, t0=0.0, t1=3.0, save_everystep=False,
                             cov_backend=oracle.COV_QR)
    assert np.all(g["retcode"] == 0) and np.all(g["t_end"] == 3.0)
    for i in range(n):
        o = oracle.solve(cfg, rhs, u0[i], [])
        scale = np.abs(o["pu_mu"][-1]).max()
        # ~600 adaptive steps of a 112-dimensional filter: the accept/reject pattern is compared with a small slack
        assert abs(int(g["naccept"][i]) - o["naccept"]) <= 2 and abs(int(g["nreject"][i]) - o["nreject"]) <= 2, (g["naccept"][i], o["naccept"])
        if g["naccept"][i] == o["naccept"] and g["nreject"][i] == o["nreject"]:
            assert np.max(np.abs(g["u"][i] - o["pu_mu"][-1])) / scale < 1e-8
        else:
            assert np.max(np.abs(g["u"][i] - o["pu_mu"][-1])) / scale < 1e-4
    # fixed steps + static diffusion: tight parity on the whole 112-dimensional state covariance
    gf = _gpu(pd, u0[:1], p[:1], 3, builtin=False, adaptive=False, dt=0.01, t0=0.0, t1=0.5, diffusion=lib.DIFF_FIXED, save_state=True)
    of = oracle.solve(oracle.make_config(28, 3, 0, alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, smooth=False, adaptive=False, dt=0.01,
                                         t0=0.0, t1=0.5, save_everystep=False, cov_backend=oracle.COV_QR), rhs, u0[0], [])
    assert np.linalg.norm(gf["u"][0] - of["pu_mu"][-1]) / np.linalg.norm(of["pu_mu"][-1]) < 1e-10
    assert _relcov(gf["cov"][0], of["pu_cov"][-1]) < 1e-10
    Sg = gf["xR"][0].T @ gf["xR"][0]
    So = of["x_filt_R"][-1].T @ of["x_filt_R"][-1]
    assert np.linalg.norm(Sg - So) / np.linalg.norm(So) < 1e-10
    assert gf["loglik"][0] == pytest.approx(of["loglik"], rel=1e-9)
    print("pleiades cta kernel_ms", g["info"].kernel_ms, "accepted", g["info"].total_accepted)


# ------------------------------------------------------------------------------------------------------------------
# mass-matrix problems M u' = f: z = M E1 x - f, H = M E1 - J E0
# ------------------------------------------------------------------------------------------------------------------
def _mass_inputs(n, seed, dp=0.1):
    pd = problems.MASS_MATRIX_PROBLEMS["rober_dae"]
    rng = np.random.default_rng(seed)
    u0 = np.tile(np.array(pd.u0, dtype=float), (n, 1))
    p = np.array(pd.p, dtype=float)[None, :] * (1 + dp * rng.uniform(-1, 1, (n, pd.n_p)))
    return pd, np.ascontiguousarray(u0), np.ascontiguousarray(p)


@pytest.mark.parametrize("t1", [1e-2, 10.0])
def test_mass_matrix_rober_dae_adaptive(oracle, t1):
    """ROBER as the reference benchmarks it (benchmarks/rober.jmd:37-46; test/mass_matrix.jl:64-88): singular
    M = Diagonal([1, 1, 0]), adaptive EK1(order=3).  The algebraic row makes the early residuals pure cancellation
    (z_3 = 1 - sum(u^-) ~ 1e-14 with 1e-16 rounding noise, sigma^2 ~ 1e16), so the free-running step sequence carries
    percent-level rounding noise in ANY two implementations -- the same situation as the stiff problems above, and the same
    protocol: (i) the kernel replays the oracle's accepted grid and stays within the oracle-vs-oracle floor, (ii) with
    sigma^2 forced as well it reaches 1e-8, (iii) free running: same work, same end point norm-wise to 1e-8 (Julia
    `isapprox`, what test/mass_matrix.jl asserts), first step 1e-6, constraint satisfied on every saved point."""
    n = 8
    pd, u0, p = _mass_inputs(n, 61)
    M = np.array(pd.mass_matrix)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    # (iii) free running ensemble
    g = _gpu(pd, u0, p, 3, builtin=False, adaptive=True, t0=0.0, t1=t1, save_everystep=True, mass_matrix=M)
    kwo = dict(alg=oracle.EK1, smooth=False, adaptive=True, t0=0.0, t1=t1, save_everystep=True, cov_backend=oracle.COV_QR,
               mass_matrix=M)
    cfg = oracle.make_config(3, 3, 3, **kwo)
    sols = []
    for i in range(n):
        o = oracle.solve(cfg, rhs, u0[i], p[i])
        sols.append(o)
        a, b = g["off"][i], g["off"][i + 1]
        assert g["retcode"][i] == 0 and o["retcode"] == "Success"
        assert abs(int(g["naccept"][i]) - o["naccept"]) <= max(2, 0.03 * o["naccept"]), i
        assert g["nf"][i] == 3 + g["naccept"][i] + g["nreject"][i]        # no f evaluations in the initial-step choice
        assert g["T"][a + 1] - g["T"][a] == o["t"][1] - o["t"][0] == 1e-6   # initial step with a mass matrix
        assert g["T"][b - 1] == t1
        assert np.linalg.norm(g["u"][i] - o["pu_mu"][-1]) <= 1e-8 * np.linalg.norm(o["pu_mu"][-1]), i
        assert np.max(np.abs(g["U"][a:b].sum(axis=1) - 1.0)) < 1e-9
    # (i) step-teacher-forced, (ii) sigma^2 forced as well
    for i in (0, 6):
        o = sols[i]
        grid = o["t"][1:-1]
        kw = dict(alg=oracle.EK1, smooth=False, adaptive=False, dt=t1, t0=0.0, t1=t1, save_everystep=True, tstops=grid,
                  cov_backend=oracle.COV_QR, mass_matrix=M)
        of = oracle.solve(oracle.make_config(3, 3, 3, **kw), rhs, u0[i], p[i])
        gi = _gpu(pd, u0[i:i + 1], p[i:i + 1], 3, builtin=False, adaptive=False, dt=t1, t0=0.0, t1=t1, save_everystep=True,
                  tstops=grid, mass_matrix=M)
        assert np.array_equal(gi["T"], of["t"]) and np.array_equal(of["t"], o["t"])
        scale = np.abs(of["pu_mu"]).max(axis=0)

        def m_mean(s_, b_):
            return np.max(np.abs(s_["pu_mu"] - b_["pu_mu"]) / scale)

        fl = rounding.floor(oracle, rhs, 3, 3, 3, u0[i], p[i], kw, of, m_mean)
        err = np.max(np.abs(gi["U"] - of["pu_mu"]) / scale)
        assert err < max(1e-8, 10.0 * fl), (i, err, fl)
        sig = of["diffusions"][:-1]
        kwf = dict(forced_diffusion=sig, **kw)
        of2 = oracle.solve(oracle.make_config(3, 3, 3, **kwf), rhs, u0[i], p[i])
        g2 = _gpu(pd, u0[i:i + 1], p[i:i + 1], 3, builtin=False, adaptive=False, dt=t1, t0=0.0, t1=t1, save_everystep=True,
                  tstops=grid, forced_diffusion=sig, mass_matrix=M)
        fl2 = rounding.floor(oracle, rhs, 3, 3, 3, u0[i], p[i], kwf, of2, m_mean)
        fc2 = rounding.floor(oracle, rhs, 3, 3, 3, u0[i], p[i], kwf, of2,
                             lambda s_, b_: _relcov(s_["pu_cov"][1:], b_["pu_cov"][1:]))
        assert np.max(np.abs(g2["U"] - of2["pu_mu"]) / scale) < max(1e-8, 10.0 * fl2), (i, fl2)
        assert _relcov(g2["C"][1:], of2["pu_cov"][1:]) < max(1e-8, 10.0 * fc2), (i, fc2)
    if t1 == 1e-2:   # test/mass_matrix.jl:81-82 (nominal parameters): implicit reference solution, rtol 1e-8
        from scipy.integrate import solve_ivp
        k1, k2, k3 = pd.p
        ref = solve_ivp(lambda t, u: [-k1 * u[0] + k3 * u[1] * u[2], k1 * u[0] - k3 * u[1] * u[2] - k2 * u[1] ** 2, k2 * u[1] ** 2],
                        (0, t1), pd.u0, method="Radau", rtol=1e-12, atol=1e-14).y[:, -1]
        g1 = _gpu(pd, u0[:1], np.array([pd.p]), 3, builtin=False, adaptive=True, t0=0.0, t1=t1, mass_matrix=M)
        assert np.linalg.norm(g1["u"][0] - ref) <= 1e-8 * np.linalg.norm(ref)


@pytest.mark.parametrize("smooth", [False, True])
def test_mass_matrix_general_fixed_steps_static_diffusion(oracle, smooth):
    """A dense, non-symmetric, non-singular M on fixed steps with FixedDiffusion: filter (and smoother) means and
    covariances of every saved point against the oracle, rtol 1e-10; plus the dense output on a saveat grid."""
    pd, u0, p = _mass_inputs(3, 62)
    u0 = u0 + np.array([0.0, 0.1, 0.2])
    p = p * np.array([1.0, 1e-6, 1e-3])
    M = np.array([[1.0, 0.2, 0.0], [0.0, 1.0, 0.0], [0.1, 0.0, 0.5]])
    kw = dict(adaptive=False, dt=1e-3, t0=0.0, t1=0.2, diffusion=lib.DIFF_FIXED, calibrate=True, mass_matrix=M)
    saveat = np.array([0.0, 0.0123, 0.1, 0.2])
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    cfg = oracle.make_config(3, 3, 3, alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, calibrate=True, smooth=smooth, adaptive=False,
                             dt=1e-3, t0=0.0, t1=0.2, cov_backend=oracle.COV_QR, mass_matrix=M)
    if smooth:
        g = _gpu_smooth(pd, u0, p, 3, **kw)
    else:
        g = _gpu(pd, u0, p, 3, builtin=False, save_everystep=True, **kw)
    gd = _gpu_dense(pd, u0, p, 3, saveat, smooth=smooth, **kw)
    for i in range(u0.shape[0]):
        o = oracle.solve(cfg, rhs, u0[i], p[i])
        a, b = g["off"][i], g["off"][i + 1]
        assert b - a == o["n"] == 201
        assert _rel(g["U"][a:b], o["pu_mu"]) < 1e-10
        assert _relcov(g["C"][a + 1:b], o["pu_cov"][1:]) < 1e-9
        for j, t in enumerate(saveat):
            m, c = oracle.interpolate(cfg, o, float(t))
            assert _rel(gd["U"][i, j], m) < 1e-10, (i, j)
            if np.linalg.norm(c) > 0:
                assert np.linalg.norm(gd["C"][i, j] - c) / np.linalg.norm(c) < 1e-8, (i, j)


def test_mass_matrix_uniform_scaling_and_errors(oracle):
    """test/mass_matrix.jl:7-17: M = -100 I, f(u) = u: adaptive EK1(order=3) end point against exp(-t/100), rtol 1e-10;
    EK0 on the Kronecker layout raises the reference's ArgumentError (src/caches.jl:111-118)."""
    def vf(du, u, p, t):
        du[0] = u[0]
        du[1] = u[1]

    pd = problems.ProblemDef("vf", vf, [1.0, 2.0], [], (0.0, 1.0))
    u0 = np.array([[1.0, 2.0], [0.5, -1.0]])
    p = np.zeros((2, 0))
    M = -100.0 * np.eye(2)
    g = _gpu(pd, u0, p, 3, builtin=False, adaptive=True, t0=0.0, t1=1.0, mass_matrix=M)
    assert np.allclose(g["u"], u0 * np.exp(-0.01), rtol=1e-10, atol=0)
    rhs = oracle.compile_rhs(tracer.trace(vf, 2, 0), 3)
    o = oracle.solve(oracle.make_config(2, 3, 0, alg=oracle.EK1, smooth=False, adaptive=True, t0=0.0, t1=1.0, cov_backend=oracle.COV_QR,
                                        mass_matrix=M), rhs, u0[0], [])
    assert abs(int(g["naccept"][0]) - o["naccept"]) <= 2 and g["nf"][0] == 3 + g["naccept"][0] + g["nreject"][0]
    assert _rel(g["u"][0], o["pu_mu"][-1]) < 1e-10
    with pytest.raises(lib.PnodeError) as e:   # not a UniformScaling
        _gpu(pd, u0, p, 3, builtin=False, alg=lib.EK0, adaptive=True, t0=0.0, t1=1.0, mass_matrix=np.diag([1.0, 2.0]))
    assert e.value.kind == "ArgumentError"


@pytest.mark.parametrize("alg,diffusion", [(lib.EK0, lib.DIFF_FIXED), (lib.EK0, lib.DIFF_FIXED_MV), (lib.DIAGONAL_EK1, lib.DIFF_FIXED)])
@pytest.mark.parametrize("smooth", [False, True])
def test_mass_matrix_structured_layouts_fixed_steps(oracle, alg, diffusion, smooth):
    """test/mass_matrix.jl:19-52 ("Kronecker working"): M = lambda I with the EK0 (Kronecker layout), and diagonal M on the
    BlocksOfDiagonals layout (EK0 + FixedMV, DiagonalEK1); fixed steps, static diffusion: every saved mean and covariance
    to 1e-10 against the oracle, and the end point against the exact solution of lambda u' = -u."""
    def vf(du, u, p, t):
        du[0] = -p[0] * u[0]
        du[1] = -p[1] * u[1] + 0.1 * u[0]

    pd = problems.ProblemDef("vf", vf, [1.0, 2.0], [1.0, 0.5], (0.0, 1.0))
    u0 = np.array([[1.0, 2.0], [0.5, -1.0]])
    p = np.array([[1.0, 0.5], [1.3, 0.7]])
    kron = (alg == lib.EK0 and diffusion == lib.DIFF_FIXED)
    M = -3.0 * np.eye(2) if kron else np.diag([-3.0, 0.5])
    g = _gpu_blocks(pd, u0, p, 3, alg, diffusion, calibrate=True, smooth=smooth, adaptive=False, dt=0.01, t0=0.0, t1=1.0,
                    mass_matrix=M)
    rhs = oracle.compile_rhs(tracer.trace(vf, 2, 2), 3)
    for i in range(2):
        o = oracle.solve(oracle.make_config(2, 3, 2, alg=alg, diffusion=diffusion, calibrate=True, smooth=smooth, adaptive=False,
                                            dt=0.01, t0=0.0, t1=1.0, cov_backend=oracle.COV_QR, mass_matrix=M), rhs, u0[i], p[i])
        a, b = g["off"][i], g["off"][i + 1]
        assert b - a == o["n"] == 101 and g["retcode"][i] == 0
        assert _rel(g["U"][a:b], o["pu_mu"]) < 1e-10
        assert _relcov(g["C"][a + 1:b], o["pu_cov"][1:]) < (1e-9 if smooth else 1e-10)
        assert g["loglik"][i] == pytest.approx(o["loglik"], rel=1e-10)
        # the initial higher derivatives are those of u' = f(u) divided by M (autodiffinit.jl:33-47), hence only 1e-4
        assert abs(g["U"][b - 1, 0] - u0[i, 0] * np.exp(-p[i, 0] / M[0, 0])) < 1e-4 * abs(u0[i, 0])


def test_mass_matrix_rober_dae_diagonal_ek1(oracle):
    """test/mass_matrix.jl:87-88: DiagonalEK1(order=3) on the Robertson DAE, rtol 1e-8 against an implicit reference
    solution; step counts as the oracle's (free running, see test_mass_matrix_rober_dae_adaptive for the protocol)."""
    from scipy.integrate import solve_ivp
    pd = problems.MASS_MATRIX_PROBLEMS["rober_dae"]
    M = np.array(pd.mass_matrix)
    k1, k2, k3 = pd.p
    ref = solve_ivp(lambda t, u: [-k1 * u[0] + k3 * u[1] * u[2], k1 * u[0] - k3 * u[1] * u[2] - k2 * u[1] ** 2, k2 * u[1] ** 2],
                    (0, 1e-2), pd.u0, method="Radau", rtol=1e-12, atol=1e-14).y[:, -1]
    u0, p = np.array([pd.u0]), np.array([pd.p])
    g = _gpu_blocks(pd, u0, p, 3, lib.DIAGONAL_EK1, lib.DIFF_DYNAMIC, adaptive=True, t0=0.0, t1=1e-2, mass_matrix=M)
    assert g["retcode"][0] == 0
    assert np.linalg.norm(g["U"][-1] - ref) <= 1e-8 * np.linalg.norm(ref)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, 3, 3), 3)
    o = oracle.solve(oracle.make_config(3, 3, 3, alg=oracle.DIAGONAL_EK1, smooth=False, adaptive=True, t0=0.0, t1=1e-2,
                                        cov_backend=oracle.COV_QR, mass_matrix=M), rhs, u0[0], p[0])
    assert abs(int(g["naccept"][0]) - o["naccept"]) <= 2
    assert g["T"][1] - g["T"][0] == 1e-6
    assert np.linalg.norm(g["U"][-1] - o["pu_mu"][-1]) <= 1e-8 * np.linalg.norm(o["pu_mu"][-1])


# ------------------------------------------------------------------------------------------------------------------
# second-order problems u'' = f1(u', u): z = E2 x - f1(E1 x, E0 x), sol.u = (du, u)
# ------------------------------------------------------------------------------------------------------------------
def _twobody_first_order(du, u, p, t):
    """First-order form F([du; u]) = [f1(du, u); du] of test/secondorderode.jl:15-19 with a gravitational parameter p[0]."""
    r3 = (u[2] * u[2] + u[3] * u[3]) ** 1.5
    du[0] = -p[0] * u[2] / r3
    du[1] = -p[0] * u[3] / r3
    du[2] = u[0]
    du[3] = u[1]


def _gpu_second(u0, p, q, *, smooth=False, saveat=None, **kw):
    src = tracer.cuda_source(tracer.trace(_twobody_first_order, 4, 1), "UserProb")
    cfg = lib.make_config(d=2, q=q, n_p=1, rhs_name="UserProb", rhs_src=src, second_order=True, smooth=smooth, save_everystep=True,
                          save_state=True, saveat=saveat, **kw)
    s = lib.Solver(cfg)
    r = s.solve(u0, p)
    n, D = u0.shape[0], 2 * (q + 1)
    out = dict(off=r.field(lib.F_STEP_OFFSETS), T=r.field(lib.F_T), U=r.field(lib.F_U).reshape(-1, 4), C=r.field(lib.F_U_COV).reshape(-1, 4, 4),
               u=r.field(lib.F_U_FINAL).reshape(n, 4), cov=r.field(lib.F_U_FINAL_COV).reshape(n, 4, 4), naccept=r.field(lib.F_NACCEPT),
               nreject=r.field(lib.F_NREJECT), nf=r.field(lib.F_NF), retcode=r.field(lib.F_RETCODE), loglik=r.field(lib.F_LOGLIK),
               Df=r.field(lib.F_DIFFUSIONS))
    if r.has(lib.F_X):
        out.update(X=r.field(lib.F_X).reshape(-1, D), XR=r.field(lib.F_X_COVFAC).reshape(-1, D, D))
    if saveat is not None:
        out.update(SU=r.field(lib.F_SAVEAT_U).reshape(n, len(saveat), 4), SC=r.field(lib.F_SAVEAT_U_COV).reshape(n, len(saveat), 4, 4))
    r.free()
    s.close()
    return out


def _second_inputs(n, seed):
    rng = np.random.default_rng(seed)
    u0 = np.array([0.0, 2.0, 0.4, 0.0])[None, :] * (1 + 0.05 * rng.uniform(-1, 1, (n, 4)))   # (du0, u0)
    p = 1.0 + 0.1 * rng.uniform(-1, 1, (n, 1))
    return np.ascontiguousarray(u0), np.ascontiguousarray(p)


@pytest.mark.parametrize("alg", [lib.EK1, lib.EK0])
@pytest.mark.parametrize("smooth", [False, True])
def test_second_order_fixed_steps_static_diffusion(oracle, alg, smooth):
    """Two-body problem as a SecondOrderODEProblem, fixed steps, FixedDiffusion: every saved sol.u = (du, u), its 4 x 4
    covariance, the diffusions and the log-likelihood against the oracle at 1e-10 (smoother covariances 1e-9); the dense
    output on a saveat grid at 1e-10 / 1e-8."""
    u0, p = _second_inputs(3, 71)
    saveat = np.array([0.0, 0.0123, 0.05, 0.1])
    kw = dict(alg=alg, adaptive=False, dt=1e-3, t0=0.0, t1=0.1, diffusion=lib.DIFF_FIXED, calibrate=True)
    g = _gpu_second(u0, p, 3, smooth=smooth, saveat=saveat, **kw)
    rhs = oracle.compile_rhs(tracer.trace(_twobody_first_order, 4, 1), 3)
    cfg = oracle.make_config(2, 3, 1, alg=alg, diffusion=oracle.DIFF_FIXED, calibrate=True, smooth=smooth, adaptive=False, dt=1e-3,
                             t0=0.0, t1=0.1, cov_backend=oracle.COV_QR, second_order=True)
    for i in range(3):
        o = oracle.solve(cfg, rhs, u0[i], p[i])
        a, b = g["off"][i], g["off"][i + 1]
        assert b - a == o["n"] == 101 and g["retcode"][i] == 0
        assert np.array_equal(g["U"][a], u0[i]) and not g["C"][a].any()        # exact initial values (test/solution.jl:59-62)
        assert _rel(g["U"][a:b], o["pu_mu"]) < 1e-10
        assert _relcov(g["C"][a + 1:b], o["pu_cov"][1:]) < (1e-9 if smooth else 1e-10)
        assert g["loglik"][i] == pytest.approx(o["loglik"], rel=1e-10)
        assert _rel(g["Df"][a:b - 1], o["diffusions"][:-1]) < 1e-10
        assert g["nf"][i] == o["nf"]
        for j, t in enumerate(saveat):
            m, c = oracle.interpolate(cfg, o, float(t))
            assert _rel(g["SU"][i, j], m) < 1e-10, (i, j)
            if np.linalg.norm(c) > 0:
                assert np.linalg.norm(g["SC"][i, j] - c) / np.linalg.norm(c) < 1e-8, (i, j)


@pytest.mark.parametrize("alg", [lib.EK1, lib.EK0])
def test_second_order_adaptive(oracle, alg):
    """test/secondorderode.jl:29-52 on an ensemble: adaptive EK0 / EK1 (default tolerances, smoothing), identical
    accept / reject / nf counts as the oracle, end points 1e-8; sol.u[end] and sol(T/pi) within rtol 1e-3 of an accurate
    solution (the reference's own assertion)."""
    from scipy.integrate import solve_ivp
    n = 12
    u0, p = _second_inputs(n, 72)
    tq = 0.1 / np.pi
    g = _gpu_second(u0, p, 3, smooth=True, saveat=np.array([tq]), alg=alg, adaptive=True, t0=0.0, t1=0.1)
    rhs = oracle.compile_rhs(tracer.trace(_twobody_first_order, 4, 1), 3)
    cfg = oracle.make_config(2, 3, 1, alg=alg, smooth=True, adaptive=True, t0=0.0, t1=0.1, cov_backend=oracle.COV_QR, second_order=True)
    for i in range(n):
        o = oracle.solve(cfg, rhs, u0[i], p[i])
        a, b = g["off"][i], g["off"][i + 1]
        assert g["retcode"][i] == 0 and o["retcode"] == "Success"
        assert g["naccept"][i] == o["naccept"] and g["nreject"][i] == o["nreject"] and g["nf"][i] == o["nf"], i
        assert np.max(np.abs(g["T"][a:b] - o["t"])[1:] / o["t"][1:]) < 2e-6
        assert np.linalg.norm(g["U"][b - 1] - o["pu_mu"][-1]) <= 1e-8 * np.linalg.norm(o["pu_mu"][-1])
        m, _ = oracle.interpolate(cfg, o, tq)
        assert np.linalg.norm(g["SU"][i, 0] - m) <= 1e-7 * np.linalg.norm(m)
        mu_i = p[i, 0]
        tr = solve_ivp(lambda t, x: [-mu_i * x[2] / (x[2] ** 2 + x[3] ** 2) ** 1.5, -mu_i * x[3] / (x[2] ** 2 + x[3] ** 2) ** 1.5, x[0], x[1]],
                       (0, 0.1), u0[i], method="DOP853", rtol=1e-12, atol=1e-12, dense_output=True)
        assert np.linalg.norm(g["U"][b - 1] - tr.y[:, -1]) <= 1e-3 * np.linalg.norm(tr.y[:, -1])
        assert np.linalg.norm(g["SU"][i, 0] - tr.sol(tq)) <= 1e-3 * np.linalg.norm(tr.sol(tq))


# ------------------------------------------------------------------------------------------------------------------
# Fenrir data log-likelihood over an ensemble of parameters (src/data_likelihoods/fenrir.jl)
# ------------------------------------------------------------------------------------------------------------------
def _gpu_fenrir(pd, u0, p, q, data_t, data_u, *, second_order=False, f=None, **kw):
    fn = f or pd.f
    dim = u0.shape[1]
    src = tracer.cuda_source(tracer.trace(fn, dim, p.shape[1]), "UserProb")
    cfg = lib.make_config(d=dim // 2 if second_order else dim, q=q, n_p=p.shape[1], rhs_name="UserProb", rhs_src=src, smooth=True,
                          save_everystep=True, second_order=second_order, data_t=data_t, data_u=data_u, **kw)
    s = lib.Solver(cfg)
    r = s.solve(u0, p)
    out = dict(ll=r.field(lib.F_FENRIR_LOGLIK), off=r.field(lib.F_STEP_OFFSETS), T=r.field(lib.F_T), retcode=r.field(lib.F_RETCODE),
               U=r.field(lib.F_U).reshape(-1, dim))
    r.free()
    s.close()
    return out


def _lv_data(rng, sigma, data_t):
    from scipy.integrate import solve_ivp
    truth = solve_ivp(lambda t, u: [1.5 * u[0] - u[0] * u[1], -3 * u[1] + u[0] * u[1]], (0, 10), [1.0, 1.0], rtol=1e-10, atol=1e-10,
                      t_eval=data_t).y.T
    return truth + sigma * rng.standard_normal(truth.shape)


@pytest.mark.parametrize("alg,diffusion", [(lib.EK1, lib.DIFF_DYNAMIC), (lib.EK1, lib.DIFF_FIXED), (lib.EK0, lib.DIFF_DYNAMIC)])
def test_fenrir_fixed_steps(oracle, alg, diffusion):
    """test/data_likelihoods.jl:17-60 over an ensemble of parameters: Lotka-Volterra, noisy observations, fixed steps
    dt = 2e-2 with the data times as tstops; the device's backward pass against the oracle's fit_pnsolution_to_data!
    (rtol 1e-8; the reference compares its three likelihood evaluations at 1e-6); full and partial observations, scalar and
    matrix noise covariances; the maximum over the sweep sits at the data-generating parameters."""
    rng = np.random.default_rng(81)
    sigma = 0.5
    data_t = np.array([2.5, 5.0, 7.5, 10.0])
    data_u = _lv_data(rng, sigma, data_t)
    n = 9
    pd, u0, p = _inputs("lotka_volterra", n, 82, dp=0.0)
    p[:, 0] = 1.5 * (1 + np.linspace(-0.2, 0.2, n))     # sweep of the first rate constant, the middle one is the truth
    rhs = oracle.compile_rhs(tracer.trace(pd.f, 2, 4), 3)
    Rm = np.array([[sigma ** 2, 0.05], [0.05, 2 * sigma ** 2]])
    cases = [dict(), dict(observation_matrix=np.array([[1.0, 0.0]])), dict(observation_noise_cov=Rm)]
    for case in cases:
        H = case.get("observation_matrix")
        du_ = data_u if H is None else data_u @ H.T
        cov = case.get("observation_noise_cov", sigma ** 2)
        if alg == lib.EK0 and case:
            # Kronecker layout: the reference updates on the n x n factors -- full observations (dataupdate.jl:69-71) and
            # isotropic noise (to_factorized_matrix, covariance_structure.jl:47-58) only
            with pytest.raises(lib.PnodeError, match="Partial observations only work with the EK1|multiple of the identity"):
                _gpu_fenrir(pd, u0, p, 3, data_t, du_, alg=alg, diffusion=diffusion, calibrate=False, adaptive=False, dt=2e-2, t0=0.0,
                            t1=10.0, observation_matrix=H, observation_noise_cov=cov)
            continue
        g = _gpu_fenrir(pd, u0, p, 3, data_t, du_, alg=alg, diffusion=diffusion, calibrate=False, adaptive=False, dt=2e-2, t0=0.0,
                        t1=10.0, observation_matrix=H, observation_noise_cov=cov)
        cfg = oracle.make_config(2, 3, 4, alg=alg, diffusion=diffusion, calibrate=False, smooth=True, adaptive=False, dt=2e-2, t0=0.0,
                                 t1=10.0, tstops=data_t, cov_backend=oracle.COV_QR)
        # EK0: the reference's Kronecker log-density, logdet of the 1 x 1 factor of S without the factor d (update.jl:140-148, 182-194)
        want = np.array([oracle.fenrir_data_loglik(cfg, oracle.solve(cfg, rhs, u0[i], p[i]), data_t, du_, observation_matrix=H,
                                                   observation_noise_cov=cov, kronecker_logdet_quirk=(alg == lib.EK0)) for i in range(n)])
        assert np.all(g["retcode"] == 0)
        assert np.allclose(g["ll"], want, rtol=1e-8, atol=0), (case.keys(), np.max(np.abs(g["ll"] / want - 1)))
        if not case:
            assert int(np.argmax(g["ll"])) in (n // 2 - 1, n // 2, n // 2 + 1)
            # the U fields stay the filtering estimates (the reference's step! skips the postamble): same as a filter-only solve
            gf = _gpu(pd, u0[:1], p[:1], 3, builtin=False, alg=alg, diffusion=diffusion, calibrate=False, adaptive=False, dt=2e-2,
                      t0=0.0, t1=10.0, save_everystep=True, tstops=data_t)
            assert np.array_equal(g["U"][g["off"][0]:g["off"][1]], gf["U"])


def test_fenrir_adaptive_and_second_order(oracle):
    """Adaptive steps (ragged step counts, data times hit through tstops) and a second-order problem (observation matrix
    applied to sol.u = (du, u): positions only) against the oracle, rtol 1e-6; a failed solve gives -Inf (fenrir.jl:55-59)."""
    rng = np.random.default_rng(83)
    sigma = 0.1
    data_t = np.array([3.0, 6.0, 10.0])
    data_u = _lv_data(rng, sigma, data_t)
    n = 6
    pd, u0, p = _inputs("lotka_volterra", n, 84, dp=0.05)
    g = _gpu_fenrir(pd, u0, p, 3, data_t, data_u, alg=lib.EK1, adaptive=True, t0=0.0, t1=10.0, observation_noise_cov=sigma ** 2)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, 2, 4), 3)
    cfg = oracle.make_config(2, 3, 4, alg=oracle.EK1, smooth=True, adaptive=True, t0=0.0, t1=10.0, tstops=data_t, cov_backend=oracle.COV_QR)
    want = np.array([oracle.fenrir_data_loglik(cfg, oracle.solve(cfg, rhs, u0[i], p[i]), data_t, data_u, observation_noise_cov=sigma ** 2)
                     for i in range(n)])
    assert len(set(np.diff(g["off"]))) > 1
    assert np.allclose(g["ll"], want, rtol=1e-6, atol=0), np.max(np.abs(g["ll"] / want - 1))
    gm = _gpu_fenrir(pd, u0, p, 3, data_t, data_u, alg=lib.EK1, adaptive=True, t0=0.0, t1=10.0, observation_noise_cov=sigma ** 2,
                     maxiters=20)
    assert np.all(gm["retcode"] != 0) and np.all(np.isneginf(gm["ll"]))
    # second-order two-body problem, positions observed
    u2, p2 = _second_inputs(4, 85)
    dt2 = np.array([0.05, 0.1])
    from scipy.integrate import solve_ivp
    tr = solve_ivp(lambda t, x: [-x[2] / (x[2] ** 2 + x[3] ** 2) ** 1.5, -x[3] / (x[2] ** 2 + x[3] ** 2) ** 1.5, x[0], x[1]], (0, 0.1),
                   [0.0, 2.0, 0.4, 0.0], rtol=1e-12, atol=1e-12, t_eval=dt2).y.T
    Hpos = np.hstack([np.zeros((2, 2)), np.eye(2)])
    obs = tr @ Hpos.T + 1e-3 * rng.standard_normal((2, 2))
    g2 = _gpu_fenrir(None, u2, p2, 3, dt2, obs, second_order=True, f=_twobody_first_order, alg=lib.EK1, adaptive=False, dt=1e-3, t0=0.0,
                     t1=0.1, diffusion=lib.DIFF_FIXED, calibrate=False, observation_matrix=Hpos, observation_noise_cov=1e-6)
    rhs2 = oracle.compile_rhs(tracer.trace(_twobody_first_order, 4, 1), 3)
    cfg2 = oracle.make_config(2, 3, 1, alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, calibrate=False, smooth=True, adaptive=False, dt=1e-3,
                              t0=0.0, t1=0.1, tstops=dt2, cov_backend=oracle.COV_QR, second_order=True)
    want2 = np.array([oracle.fenrir_data_loglik(cfg2, oracle.solve(cfg2, rhs2, u2[i], p2[i]), dt2, obs, observation_matrix=Hpos,
                                                observation_noise_cov=1e-6) for i in range(4)])
    assert np.allclose(g2["ll"], want2, rtol=1e-7, atol=0), np.max(np.abs(g2["ll"] / want2 - 1))


@pytest.mark.parametrize("alg,diffusion", [(lib.DIAGONAL_EK1, lib.DIFF_FIXED), (lib.EK0, lib.DIFF_FIXED_MV), (lib.DIAGONAL_EK1, lib.DIFF_FIXED_MV)])
@pytest.mark.parametrize("smooth", [False, True])
def test_second_order_blocks_layout_fixed_steps(oracle, alg, diffusion, smooth):
    """Second-order problems on the BlocksOfDiagonals layout (DiagonalEK1: H_i = e_2' - Jdu_ii e_1' - Ju_ii e_0',
    src/derivative_utils.jl:14-28; EK0 with multivariate diffusion): fixed steps, static diffusion, every saved
    sol.u = (du, u) with its covariance against the oracle at 1e-10 / 1e-9."""
    u0, p = _second_inputs(2, 91)
    kw = dict(alg=alg, adaptive=False, dt=1e-3, t0=0.0, t1=0.1, diffusion=diffusion, calibrate=True)
    g = _gpu_second(u0, p, 3, smooth=smooth, **kw)
    rhs = oracle.compile_rhs(tracer.trace(_twobody_first_order, 4, 1), 3)
    cfg = oracle.make_config(2, 3, 1, alg=alg, diffusion=diffusion, calibrate=True, smooth=smooth, adaptive=False, dt=1e-3, t0=0.0, t1=0.1,
                             cov_backend=oracle.COV_QR, second_order=True)
    for i in range(2):
        o = oracle.solve(cfg, rhs, u0[i], p[i])
        a, b = g["off"][i], g["off"][i + 1]
        assert b - a == o["n"] == 101 and g["retcode"][i] == 0
        assert _rel(g["U"][a:b], o["pu_mu"]) < 1e-10
        assert _relcov(g["C"][a + 1:b], o["pu_cov"][1:]) < (1e-9 if smooth else 1e-10)
        assert g["loglik"][i] == pytest.approx(o["loglik"], rel=1e-10)
        assert _relcov(g["cov"][i][None], o["pu_cov"][-1][None]) < (1e-9 if smooth else 1e-10)


@pytest.mark.parametrize("alg,diffusion", [(lib.DIAGONAL_EK1, lib.DIFF_DYNAMIC), (lib.EK0, lib.DIFF_DYNAMIC_MV), (lib.DIAGONAL_EK1, lib.DIFF_DYNAMIC_MV)])
def test_second_order_blocks_layout_adaptive(oracle, alg, diffusion):
    """test/secondorderode.jl:30-52, the DiagonalEK1 / multivariate-diffusion rows: adaptive, identical accept / reject /
    nf counts as the oracle, end point 1e-8, and rtol 1e-3 against an accurate solution."""
    from scipy.integrate import solve_ivp
    n = 6
    u0, p = _second_inputs(n, 92)
    g = _gpu_second(u0, p, 3, smooth=True, alg=alg, diffusion=diffusion, adaptive=True, t0=0.0, t1=0.1)
    rhs = oracle.compile_rhs(tracer.trace(_twobody_first_order, 4, 1), 3)
    cfg = oracle.make_config(2, 3, 1, alg=alg, diffusion=diffusion, smooth=True, adaptive=True, t0=0.0, t1=0.1, cov_backend=oracle.COV_QR,
                             second_order=True)
    for i in range(n):
        o = oracle.solve(cfg, rhs, u0[i], p[i])
        a, b = g["off"][i], g["off"][i + 1]
        assert g["naccept"][i] == o["naccept"] and g["nreject"][i] == o["nreject"] and g["nf"][i] == o["nf"], i
        assert np.linalg.norm(g["U"][b - 1] - o["pu_mu"][-1]) <= 1e-8 * np.linalg.norm(o["pu_mu"][-1])
        mu_i = p[i, 0]
        tr = solve_ivp(lambda t, x: [-mu_i * x[2] / (x[2] ** 2 + x[3] ** 2) ** 1.5, -mu_i * x[3] / (x[2] ** 2 + x[3] ** 2) ** 1.5, x[0], x[1]],
                       (0, 0.1), u0[i], method="DOP853", rtol=1e-12, atol=1e-12)
        assert np.linalg.norm(g["U"][b - 1] - tr.y[:, -1]) <= 1e-3 * np.linalg.norm(tr.y[:, -1])


def test_new_paths_at_scale_agree_with_small_runs(oracle):
    """Size-independent property for the mass-matrix, second-order and Fenrir paths: a large adaptive ensemble goes through
    the pilot-sized single save pass (over-allocated segments + compaction), a small one through the counting pass; the
    first trajectories of the large run must equal the small run bit for bit (every trajectory's arithmetic depends on
    its own history only), and every trajectory of the large run must succeed."""
    n_big, n_small = 4096, 6
    # mass matrix (ROBER DAE)
    pd, u0, p = _mass_inputs(n_big, 101)
    M = np.array(pd.mass_matrix)
    kw = dict(builtin=False, adaptive=True, t0=0.0, t1=1.0, save_everystep=True, mass_matrix=M)
    big, small = _gpu(pd, u0, p, 3, **kw), _gpu(pd, u0[:n_small], p[:n_small], 3, **kw)
    assert np.all(big["retcode"] == 0) and big["info"].total_saved == big["off"][-1]
    e = small["off"][-1]
    assert np.array_equal(big["off"][:n_small + 1], small["off"]) and np.array_equal(big["T"][:e], small["T"])
    assert np.array_equal(big["U"][:e], small["U"]) and np.array_equal(big["C"][:e], small["C"])
    assert np.max(np.abs(big["U"].sum(axis=1) - 1.0)) < 1e-9          # the algebraic constraint, every point of every trajectory
    # second-order problem, smoothed, with dense output
    u2, p2 = _second_inputs(n_big, 102)
    sv = np.array([0.02, 0.1 / np.pi, 0.1])
    kw2 = dict(smooth=True, saveat=sv, alg=lib.EK1, adaptive=True, t0=0.0, t1=0.1)
    b2, s2 = _gpu_second(u2, p2, 3, **kw2), _gpu_second(u2[:n_small], p2[:n_small], 3, **kw2)
    e = s2["off"][-1]
    assert np.all(b2["retcode"] == 0) and np.array_equal(b2["off"][:n_small + 1], s2["off"])
    assert np.array_equal(b2["U"][:e], s2["U"]) and np.array_equal(b2["C"][:e], s2["C"])
    assert np.array_equal(b2["SU"][:n_small], s2["SU"]) and np.array_equal(b2["SC"][:n_small], s2["SC"])
    # energy of the two-body problem along the smoothed solution: conserved to solver accuracy
    E = 0.5 * (b2["U"][:, 0] ** 2 + b2["U"][:, 1] ** 2) - np.repeat(p2[:, 0], np.diff(b2["off"])) / np.hypot(b2["U"][:, 2], b2["U"][:, 3])
    E0 = np.repeat(E[b2["off"][:-1]], np.diff(b2["off"]))
    assert np.max(np.abs(E - E0) / np.abs(E0)) < 2e-2      # default tolerances (reltol 1e-3): measured 6e-3
    # Fenrir log-likelihoods of a parameter sweep
    rng = np.random.default_rng(103)
    data_t = np.array([2.0, 4.0, 6.0, 8.0, 10.0])
    data_u = _lv_data(rng, 0.2, data_t)
    pdl, u0l, pl = _inputs("lotka_volterra", n_big, 104, dp=0.1)
    kwf = dict(alg=lib.EK1, adaptive=True, t0=0.0, t1=10.0, observation_noise_cov=0.04)
    bf, sf = _gpu_fenrir(pdl, u0l, pl, 3, data_t, data_u, **kwf), _gpu_fenrir(pdl, u0l[:n_small], pl[:n_small], 3, data_t, data_u, **kwf)
    assert np.all(bf["retcode"] == 0) and np.all(np.isfinite(bf["ll"]))
    assert np.array_equal(bf["ll"][:n_small], sf["ll"])
    # the data were generated at the nominal parameters: the likelihood decreases with the distance from them
    dist = np.linalg.norm(pl / np.array(pdl.p) - 1.0, axis=1)
    near, far = bf["ll"][dist < 0.05], bf["ll"][dist > 0.12]
    assert near.size > 10 and far.size > 10 and np.median(near) > np.median(far)


def test_fenrir_register_and_shared_memory_sweeps_agree(monkeypatch):
    """The Fenrir backward pass exists twice: in the column-distributed register sweep (default) and in the shared-memory
    warp-per-trajectory sweep (fallback for observation vectors longer than sol.u; PNODE_FENRIR_GENERIC=1).  Same
    formulas, different reduction orders: the log-likelihoods agree to 1e-9."""
    rng = np.random.default_rng(111)
    data_t = np.array([1.0, 2.5, 4.0, 10.0])
    data_u = _lv_data(rng, 0.3, data_t)
    pd, u0, p = _inputs("lotka_volterra", 40, 112, dp=0.1)
    kw = dict(alg=lib.EK1, adaptive=True, t0=0.0, t1=10.0, observation_noise_cov=np.array([[0.09, 0.01], [0.01, 0.2]]))
    a = _gpu_fenrir(pd, u0, p, 3, data_t, data_u, **kw)
    monkeypatch.setenv("PNODE_FENRIR_GENERIC", "1")
    b = _gpu_fenrir(pd, u0, p, 3, data_t, data_u, **kw)
    assert np.all(np.isfinite(a["ll"])) and np.allclose(a["ll"], b["ll"], rtol=1e-9, atol=0)
