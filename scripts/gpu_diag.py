import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import probnumdiffeq_jl_b200
from probnumdiffeq_jl_b200 import lib, problems, tracer
from oracle import oracle as O
from test_gpu_parity import _gpu, _inputs, _rel, _relcov

# 1: dynamic diffusion fixed steps
pd, u0, p = _inputs("fitzhugh_nagumo", 1, 1, dp=0.0)
tr = tracer.trace(pd.f, pd.d, pd.n_p); rhs = O.compile_rhs(tr, 3)
kw = dict(alg=O.EK1, diffusion=O.DIFF_DYNAMIC, smooth=False, adaptive=False, dt=0.01, t0=0.0, t1=20.0, save_everystep=True)
o = O.solve(O.make_config(2, 3, 3, cov_backend=O.COV_QR, **kw), rhs, u0[0], p[0])
oc = O.solve(O.make_config(2, 3, 3, cov_backend=O.COV_REF, **kw), rhs, u0[0], p[0])
g = _gpu(pd, u0, p, 3, builtin=False, adaptive=False, dt=0.01, t0=0.0, t1=20.0, save_everystep=True)
print("dyn: T equal", np.array_equal(g["T"], o["t"]), "U", _rel(g["U"], o["pu_mu"]), "C", _relcov(g["C"][1:], o["pu_cov"][1:]),
      "Df", np.max(np.abs(g["Df"][:-1]-o["diffusions"][:-1])/np.abs(o["diffusions"][:-1])))
print("floor (oracle chol vs qr): U", _rel(oc["pu_mu"], o["pu_mu"]), "C", _relcov(oc["pu_cov"][1:], o["pu_cov"][1:]),
      "Df", np.max(np.abs(oc["diffusions"][:-1]-o["diffusions"][:-1])/np.abs(o["diffusions"][:-1])))
sig = o["diffusions"][:-1]
o2 = O.solve(O.make_config(2, 3, 3, cov_backend=O.COV_QR, forced_diffusion=sig, **kw), rhs, u0[0], p[0])
g2 = _gpu(pd, u0, p, 3, builtin=False, adaptive=False, dt=0.01, t0=0.0, t1=20.0, save_everystep=True, forced_diffusion=sig)
print("forced: U", _rel(g2["U"], o2["pu_mu"]), "C", _relcov(g2["C"][1:], o2["pu_cov"][1:]))

# 3: LV q=5 adaptive
n = 24
pd, u0, p = _inputs("lotka_volterra", n, 2, du0=0.05)
g = _gpu(pd, u0, p, 5, builtin=False, adaptive=True, t0=0.0, t1=10.0, save_everystep=True)
rhs = O.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 5)
cfg = O.make_config(2, 5, 4, alg=O.EK1, smooth=False, adaptive=True, t0=0.0, t1=10.0, save_everystep=True, cov_backend=O.COV_QR)
cfgc = O.make_config(2, 5, 4, alg=O.EK1, smooth=False, adaptive=True, t0=0.0, t1=10.0, save_everystep=True, cov_backend=O.COV_REF)
for i in range(n):
    o = O.solve(cfg, rhs, u0[i], p[i]); oc = O.solve(cfgc, rhs, u0[i], p[i])
    a, b = g["off"][i], g["off"][i + 1]
    same = g["naccept"][i] == o["naccept"] and g["nreject"][i] == o["nreject"]
    samec = oc["naccept"] == o["naccept"] and oc["nreject"] == o["nreject"]
    dt = np.max(np.abs(g["T"][a:b] - o["t"])[1:] / o["t"][1:]) if same else -1
    dtc = np.max(np.abs(oc["t"] - o["t"])[1:] / o["t"][1:]) if samec else -1
    print(i, "gpu acc/rej", g["naccept"][i], g["nreject"][i], "oracle", o["naccept"], o["nreject"], "chol-oracle", oc["naccept"], oc["nreject"],
          "dT %.2e floor %.2e" % (dt, dtc), "final du %.2e floor %.2e" % (_rel(g["u"][i:i+1], o["pu_mu"][-1:]), _rel(oc["pu_mu"][-1:], o["pu_mu"][-1:])))
