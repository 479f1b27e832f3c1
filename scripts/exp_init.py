import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import probnumdiffeq_jl_b200
from probnumdiffeq_jl_b200 import lib, problems
n = 262144
pd = problems.PROBLEMS["rober"]
rng = np.random.Generator(np.random.PCG64(3))
u0 = np.tile(np.array(pd.u0, dtype=float), (n, 1))
p = np.array(pd.p, dtype=float)[None, :] * (1 + 0.1 * rng.uniform(-1, 1, (n, 3)))
for dt in (0.0, 1.2e-4):
    cfg = lib.make_config(d=3, q=3, n_p=3, rhs_name="builtin:rober", adaptive=True, t0=0.0, t1=100.0, dt=dt)
    s = lib.Solver(cfg)
    for rep in range(3):
        r = s.solve(u0, p)
        acc, rej, ms = r.info.total_accepted, r.info.total_rejected, r.info.kernel_ms
        r.free()
    print(f"dt0={dt}: kernel {ms:.2f} ms accepted {acc} rejected {rej} -> {(acc)/(ms*1e-3):.4e} steps/s, attempts/s {(acc+rej)/(ms*1e-3):.4e}")
    s.close()
