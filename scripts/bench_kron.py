import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import probnumdiffeq_jl_b200
from probnumdiffeq_jl_b200 import lib, problems
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
pd = problems.PROBLEMS["lotka_volterra"]
rng = np.random.Generator(np.random.PCG64(0))
u0 = np.array(pd.u0, dtype=float)[None, :] * (1 + 0.1 * rng.uniform(-1, 1, (n, 2)))
p = np.array(pd.p, dtype=float)[None, :] * (1 + 0.1 * rng.uniform(-1, 1, (n, 4)))
cfg = lib.make_config(d=2, q=3, n_p=4, rhs_name="builtin:lotka_volterra", alg=lib.EK0, adaptive=True, t0=0.0, t1=10.0)
s = lib.Solver(cfg)
for rep in range(3):
    t = time.perf_counter(); r = s.solve(u0, p); w = time.perf_counter() - t
    print(f"cfg2 LV EK0 n={n}: kernel {r.info.kernel_ms:.2f} ms wall {w*1e3:.2f} ms accepted {r.info.total_accepted} -> {r.info.total_accepted/(r.info.kernel_ms*1e-3):.3e} steps/s", flush=True)
    r.free()
