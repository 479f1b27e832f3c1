"""Small ensembles through every kernel family, for compute-sanitizer (racecheck / memcheck):

    compute-sanitizer --tool racecheck python scripts/sanitize_small.py

256 trajectories each: ROBER adaptive filter (pn_small, 2 lanes per trajectory, shared-memory exchange of the factor rows), Van der
Pol filter + smoother (pn_small SAVE=2, column-distributed register sweep with cp.async record staging), Lotka-Volterra EK0
(pn_kron), FitzHugh-Nagumo DiagonalEK1 (pn_blocks), Lotka-Volterra with a matrix-rate IOUP prior (pn_gen, warp per trajectory, and
the shared-memory sweep), Pleiades on the CTA kernel (8 trajectories)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import probnumdiffeq_jl_b200  # noqa: F401
from probnumdiffeq_jl_b200 import lib, problems, tracer

N = int(os.environ.get("SAN_N", "256"))


def run(tag, name, q, t1, n=N, builtin=True, **kw):
    pd = problems.PROBLEMS[name]
    rng = np.random.default_rng(1)
    u0 = np.tile(np.array(pd.u0, dtype=float), (n, 1))
    p = np.array(pd.p, dtype=float)[None, :] * (1 + 0.1 * rng.uniform(-1, 1, (n, pd.n_p))) if pd.n_p else np.zeros((n, 0))
    if builtin:
        cfg = lib.make_config(d=pd.d, q=q, n_p=pd.n_p, rhs_name="builtin:" + name, t0=0.0, t1=t1, **kw)
    else:
        src = tracer.cuda_source(tracer.trace(pd.f, pd.d, pd.n_p), "SanProb")
        cfg = lib.make_config(d=pd.d, q=q, n_p=pd.n_p, rhs_name="SanProb", rhs_src=src, t0=0.0, t1=t1, **kw)
    s = lib.Solver(cfg)
    r = s.solve(np.ascontiguousarray(u0), np.ascontiguousarray(p))
    rc = r.field(lib.F_RETCODE)
    print(f"{tag}: {n} trajectories, {int(r.info.total_accepted)} accepted steps, launches {int(r.info.n_launches)}, all Success: {bool((rc == 0).all())}", flush=True)
    r.free()
    s.close()


run("pn_small rober adaptive filter", "rober", 3, 100.0)
run("pn_small + pn_smooth_col vanderpol filter+smoother", "vanderpol", 5, 0.5, smooth=True, save_everystep=True)
run("pn_kron lotka_volterra EK0", "lotka_volterra", 3, 5.0, alg=lib.EK0)
run("pn_blocks fitzhugh_nagumo DiagonalEK1", "fitzhugh_nagumo", 3, 5.0, builtin=False, alg=lib.DIAGONAL_EK1)
run("pn_gen + pn_smooth lotka_volterra IOUP matrix rate", "lotka_volterra", 3, 2.0, n=32, builtin=False, prior=lib.PRIOR_IOUP,
    rate_parameter=np.array([[-1.0, 0.3], [-0.2, -0.5]]), smooth=True, save_everystep=True)
run("pn_cta pleiades", "pleiades", 3, 0.3, n=8)
