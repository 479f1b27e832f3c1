import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import probnumdiffeq_jl_b200
from probnumdiffeq_jl_b200 import lib, problems, tracer
from oracle import oracle as O
from test_gpu_parity import _gpu, _inputs, _rel, _relcov
name, q, t1 = "vanderpol", 5, 2.0
pd, u0, p = _inputs(name, 1, 3, dp=0.0)
rhs = O.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), q)
o = O.solve(O.make_config(pd.d, q, pd.n_p, alg=O.EK1, smooth=False, adaptive=True, t0=0.0, t1=t1, save_everystep=True, cov_backend=O.COV_QR), rhs, u0[0], p[0])
grid = o["t"][1:-1]
kw = dict(alg=O.EK1, smooth=False, adaptive=False, dt=t1, t0=0.0, t1=t1, save_everystep=True, tstops=grid, cov_backend=O.COV_QR)
of = O.solve(O.make_config(pd.d, q, pd.n_p, **kw), rhs, u0[0], p[0])
O.use_variant("fma"); om = O.solve(O.make_config(pd.d, q, pd.n_p, **kw), rhs, u0[0], p[0]); O.use_variant("")
for lanes in (2, 4, 8):
    g = _gpu(pd, u0, p, q, builtin=False, adaptive=False, dt=t1, t0=0.0, t1=t1, save_everystep=True, tstops=grid, lanes_per_traj=lanes)
    scale = np.abs(of["pu_mu"]).max(axis=0)
    rs_k = np.abs(g["Df"][:-1] - of["diffusions"][:-1]) / np.abs(of["diffusions"][:-1])
    rs_m = np.abs(om["diffusions"][:-1] - of["diffusions"][:-1]) / np.abs(of["diffusions"][:-1])
    em_k = np.max(np.abs(g["U"] - of["pu_mu"]) / scale, axis=1)
    em_m = np.max(np.abs(om["pu_mu"] - of["pu_mu"]) / scale, axis=1)
    print("lanes", lanes, "max mean err", em_k.max(), "fma", em_m.max())
    if lanes == 4:
        for k in list(range(0, 40)) + list(range(40, len(rs_k), 20)):
            print(k, "t=%.5f h=%.2e" % (of["t"][k+1], of["t"][k+1]-of["t"][k]), "sig2 %.3e" % of["diffusions"][k], "relerr kernel %.2e fma %.2e" % (rs_k[k], rs_m[k]), "mean err kernel %.2e fma %.2e" % (em_k[k+1], em_m[k+1]))
