"""First GPU bring-up: CUDA path vs oracle on a few small cases (scratch; real tests live in tests/)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import probnumdiffeq_jl_b200 as P
from probnumdiffeq_jl_b200 import lib, tracer, problems
from oracle import oracle as O

def run(name, q, adaptive, diffusion, dt, t1, n=64, builtin=True, seed=0, lanes=0, **kw):
    pd = problems.PROBLEMS[name]
    tr = tracer.trace(pd.f, pd.d, pd.n_p)
    rng = np.random.default_rng(seed)
    u0 = np.array(pd.u0)[None, :] * (1 + 0.0 * rng.uniform(-1, 1, (n, pd.d)))
    p = np.array(pd.p)[None, :] * (1 + 0.1 * rng.uniform(-1, 1, (n, pd.n_p)))
    if builtin:
        cfg = lib.make_config(d=pd.d, q=q, n_p=pd.n_p, rhs_name="builtin:" + name, adaptive=adaptive,
                              diffusion=diffusion, dt=dt, t0=0.0, t1=t1, save_state=True, lanes_per_traj=lanes, **kw)
    else:
        cfg = lib.make_config(d=pd.d, q=q, n_p=pd.n_p, rhs_name="UserProb", rhs_src=tracer.cuda_source(tr, "UserProb"),
                              adaptive=adaptive, diffusion=diffusion, dt=dt, t0=0.0, t1=t1, save_state=True,
                              lanes_per_traj=lanes, **kw)
    t = time.time(); s = lib.Solver(cfg); tc = time.time() - t
    r = s.solve(u0, p)
    uf = r.field(lib.F_U_FINAL).reshape(n, pd.d); cov = r.field(lib.F_U_FINAL_COV).reshape(n, pd.d, pd.d)
    na = r.field(lib.F_NACCEPT); nr = r.field(lib.F_NREJECT); rc = r.field(lib.F_RETCODE); nf = r.field(lib.F_NF)
    ll = r.field(lib.F_LOGLIK)
    rhs = O.compile_rhs(tr, q)
    ocfg = O.make_config(pd.d, q, pd.n_p, alg=O.EK1, diffusion=diffusion, smooth=False, adaptive=adaptive,
                         save_everystep=False, dt=dt, t0=0.0, t1=t1, cov_backend=O.COV_QR)
    o = O.solve_ensemble(ocfg, rhs, u0, p)
    du = np.max(np.abs(uf - o["u_final"]) / (np.abs(o["u_final"]) + 1e-300))
    dc = np.max(np.linalg.norm(cov - o["cov_final"], axis=(1, 2)) / (np.linalg.norm(o["cov_final"], axis=(1, 2)) + 1e-300))
    print(f"{name} q={q} adaptive={adaptive} diff={diffusion} builtin={builtin}: create {tc:.2f}s kernel {r.info.kernel_ms:.3f} ms | "
          f"naccept gpu {na[:4]} oracle {o['naccept'][:4]} same={np.array_equal(na, o['naccept'])} nrej same={np.array_equal(nr, o['nreject'])} "
          f"nf same={np.array_equal(nf, o['nf'])} retcodes {np.unique(rc)} | max rel du {du:.2e} dcov {dc:.2e} dll {np.max(np.abs(ll-o['loglik'])/np.abs(o['loglik'])):.2e}",
          flush=True)

if __name__ == "__main__":
    print("devices", lib.device_count(), flush=True)
    run("fitzhugh_nagumo", 3, True, lib.DIFF_DYNAMIC, 0.0, 5.0)
    run("fitzhugh_nagumo", 3, False, lib.DIFF_DYNAMIC, 0.01, 5.0, builtin=False)
    run("fitzhugh_nagumo", 3, False, lib.DIFF_FIXED, 0.01, 5.0, builtin=False)
    run("lotka_volterra", 3, True, lib.DIFF_DYNAMIC, 0.0, 10.0)
    run("rober", 3, True, lib.DIFF_DYNAMIC, 0.0, 100.0)
    run("rober", 3, True, lib.DIFF_DYNAMIC, 0.0, 100.0, builtin=False)
    run("orego", 3, True, lib.DIFF_DYNAMIC, 0.0, 30.0)
    run("vanderpol", 5, True, lib.DIFF_DYNAMIC, 0.0, 2.0)
    # throughput probe
    for n in (1 << 14, 1 << 17, 1 << 20):
        pd = problems.PROBLEMS["rober"]
        rng = np.random.default_rng(3)
        u0 = np.tile(np.array(pd.u0), (n, 1)); p = np.array(pd.p)[None, :] * (1 + 0.1 * rng.uniform(-1, 1, (n, 3)))
        cfg = lib.make_config(d=3, q=3, n_p=3, rhs_name="builtin:rober", t0=0.0, t1=100.0)
        s = lib.Solver(cfg)
        for rep in range(2):
            t = time.time(); r = s.solve(u0, p); wall = time.time() - t
            print(f"rober n={n}: kernel {r.info.kernel_ms:.2f} ms, wall {wall*1e3:.1f} ms, accepted {r.info.total_accepted}, rejected {r.info.total_rejected}, "
                  f"{r.info.total_accepted / (r.info.kernel_ms * 1e-3):.3e} steps/s (kernel)", flush=True)
            r.free()
