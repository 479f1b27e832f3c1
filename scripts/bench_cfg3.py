import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import probnumdiffeq_jl_b200
from probnumdiffeq_jl_b200 import lib, problems
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
pd = problems.PROBLEMS["vanderpol"]
rng = np.random.Generator(np.random.PCG64(1))
u0 = np.array(pd.u0, dtype=float)[None, :] * np.ones((n, 1))
p = np.array(pd.p, dtype=float)[None, :] * (1 + 0.2 * rng.uniform(-1, 1, (n, 1)))
cfg = lib.make_config(d=2, q=5, n_p=1, rhs_name="builtin:vanderpol", adaptive=True, t0=0.0, t1=2.0, smooth=True, save_everystep=True)
s = lib.Solver(cfg)
for rep in range(2):
    r = s.solve(u0, p)
    print(f"cfg3 n={n}: kernel {r.info.kernel_ms:.1f} ms accepted {r.info.total_accepted} launches {r.info.n_launches}", flush=True)
    r.free()
