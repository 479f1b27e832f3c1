import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import probnumdiffeq_jl_b200
from probnumdiffeq_jl_b200 import lib, problems
name = sys.argv[1] if len(sys.argv) > 1 else "vanderpol"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
q = {"vanderpol": 5, "fitzhugh_nagumo": 3, "rober": 3, "lotka_volterra": 3}[name]
t1 = {"vanderpol": 2.0, "fitzhugh_nagumo": 20.0, "rober": 100.0, "lotka_volterra": 10.0}[name]
pd = problems.PROBLEMS[name]
rng = np.random.default_rng(1)
u0 = np.tile(np.array(pd.u0, dtype=float), (n, 1))
p = np.array(pd.p, dtype=float)[None, :] * (1 + 0.2 * rng.uniform(-1, 1, (n, pd.n_p)))
if name == "vanderpol": u0[:, 0] += 0.1 * rng.uniform(-1, 1, n)
for smooth in (False, True):
    cfg = lib.make_config(d=pd.d, q=q, n_p=pd.n_p, rhs_name="builtin:" + name, adaptive=True, t0=0.0, t1=t1, smooth=smooth,
                          save_everystep=True)
    s = lib.Solver(cfg)
    for rep in range(2):
        t = time.time(); r = s.solve(u0, p); wall = time.time() - t
        print(f"{name} n={n} smooth={smooth}: kernel {r.info.kernel_ms:.1f} ms wall {wall*1e3:.1f} ms launches {r.info.n_launches} accepted {r.info.total_accepted} "
              f"saved {r.info.total_saved} -> {r.info.total_accepted/(r.info.kernel_ms*1e-3):.3e} steps/s", flush=True)
        r.free()
    s.close()
