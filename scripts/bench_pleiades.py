import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import probnumdiffeq_jl_b200
from probnumdiffeq_jl_b200 import lib, problems
n = int(sys.argv[1]) if len(sys.argv) > 1 else 592
pd = problems.PROBLEMS["pleiades"]
rng = np.random.default_rng(2)
u0 = np.tile(np.array(pd.u0, dtype=float), (n, 1)); u0[:, 14:] += 1e-3 * rng.uniform(-1, 1, (n, 14))
p = np.zeros((n, 0))
cfg = lib.make_config(d=28, q=3, n_p=0, rhs_name="builtin:pleiades", adaptive=True, t0=0.0, t1=3.0)
s = lib.Solver(cfg)
for rep in range(2):
    r = s.solve(u0, p)
    acc = r.info.total_accepted; ms = r.info.kernel_ms
    F = (22/3)*112**3 + 8*112**2*28 + 6*112**2 + 6*112*28**2 + (2/3)*28**3
    print(f"pleiades n={n}: kernel {ms:.1f} ms, accepted {acc}, rejected {r.info.total_rejected}, {acc/(ms*1e-3):.3e} steps/s, {acc*F/(ms*1e-3)/1e12:.2f} TFLOP/s algorithmic", flush=True)
    r.free()
