import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import probnumdiffeq_jl_b200
from probnumdiffeq_jl_b200 import lib, problems, tracer
from oracle import oracle as O
from test_gpu_parity import _gpu, _inputs, _rel, _relcov

for name, q, t1 in [("vanderpol", 5, 2.0), ("orego", 3, 30.0)]:
    pd, u0, p = _inputs(name, 1, 3, dp=0.0)
    rhs = O.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), q)
    o = O.solve(O.make_config(pd.d, q, pd.n_p, alg=O.EK1, smooth=False, adaptive=True, t0=0.0, t1=t1, save_everystep=True, cov_backend=O.COV_QR), rhs, u0[0], p[0])
    grid = o["t"][1:-1]
    kw = dict(alg=O.EK1, smooth=False, adaptive=False, dt=t1, t0=0.0, t1=t1, save_everystep=True, tstops=grid)
    of = O.solve(O.make_config(pd.d, q, pd.n_p, cov_backend=O.COV_QR, **kw), rhs, u0[0], p[0])
    oc = O.solve(O.make_config(pd.d, q, pd.n_p, cov_backend=O.COV_REF, **kw), rhs, u0[0], p[0])
    g = _gpu(pd, u0, p, q, builtin=False, adaptive=False, dt=t1, t0=0.0, t1=t1, save_everystep=True, tstops=grid)
    scale = np.abs(of["pu_mu"]).max(axis=0)
    print(name, "grid equal", np.array_equal(g["T"], of["t"]), np.array_equal(of["t"], o["t"]), len(o["t"]))
    print("  step-forced: kernel err", np.max(np.abs(g["U"] - of["pu_mu"]) / scale), "floor", np.max(np.abs(oc["pu_mu"] - of["pu_mu"]) / scale),
          "cov kernel", _relcov(g["C"][1:], of["pu_cov"][1:]), "cov floor", _relcov(oc["pu_cov"][1:], of["pu_cov"][1:]))
    sig = of["diffusions"][:-1]
    of2 = O.solve(O.make_config(pd.d, q, pd.n_p, cov_backend=O.COV_QR, forced_diffusion=sig, **kw), rhs, u0[0], p[0])
    oc2 = O.solve(O.make_config(pd.d, q, pd.n_p, cov_backend=O.COV_REF, forced_diffusion=sig, **kw), rhs, u0[0], p[0])
    g2 = _gpu(pd, u0, p, q, builtin=False, adaptive=False, dt=t1, t0=0.0, t1=t1, save_everystep=True, tstops=grid, forced_diffusion=sig)
    print("  both forced: kernel err", np.max(np.abs(g2["U"] - of2["pu_mu"]) / scale), "floor", np.max(np.abs(oc2["pu_mu"] - of2["pu_mu"]) / scale),
          "cov kernel", _relcov(g2["C"][1:], of2["pu_cov"][1:]), "cov floor", _relcov(oc2["pu_cov"][1:], of2["pu_cov"][1:]))
    gf = _gpu(pd, u0, p, q, adaptive=True, t0=0.0, t1=t1)
    print("  free: naccept", gf["naccept"][0], o["naccept"], "final", np.max(np.abs(gf["u"][0] - o["pu_mu"][-1]) / scale), gf["retcode"][0])

n = 24
pd, u0, p = _inputs("fitzhugh_nagumo", n, 2, du0=0.05)
g = _gpu(pd, u0, p, 3, builtin=True, adaptive=True, t0=0.0, t1=20.0, save_everystep=True)
rhs = O.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
cfg = O.make_config(2, 3, 3, alg=O.EK1, smooth=False, adaptive=True, t0=0.0, t1=20.0, save_everystep=True, cov_backend=O.COV_QR)
for i in range(n):
    o = O.solve(cfg, rhs, u0[i], p[i])
    a, b = g["off"][i], g["off"][i + 1]
    same = g["naccept"][i] == o["naccept"] and g["nreject"][i] == o["nreject"]
    if not same:
        print(i, "MISMATCH", g["naccept"][i], g["nreject"][i], o["naccept"], o["nreject"]); continue
    dT = np.abs(g["T"][a:b] - o["t"])
    scale = np.abs(o["pu_mu"]).max(axis=0)
    du = np.abs(np.gradient(o["pu_mu"], o["t"], axis=0)).max(axis=0)
    exc = np.max(np.abs(g["U"][a:b] - o["pu_mu"]) / (1e-8 * scale + 4.0 * du * dT[:, None] + 1e-300))
    print(i, "dT %.2e" % np.max(dT[1:] / o["t"][1:]), "final %.2e" % np.max(np.abs(g["u"][i] - o["pu_mu"][-1]) / scale), "interm ratio %.2f" % exc,
          "cov %.2e" % _relcov(g["C"][a + 1:b], o["pu_cov"][1:]))
