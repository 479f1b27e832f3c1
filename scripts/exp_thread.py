"""Thread-per-trajectory kernel (pn_thread.cuh, lanes_per_traj = 1) against the group-per-trajectory kernel (pn_small.cuh,
PNODE_THREAD_KERNEL=0): same ensemble, ahead-of-time builds, device-event kernel time, accept / reject counts and end points.

    python scripts/exp_thread.py [n_traj]      -> one JSON line per (problem, kernel) (profiles/r2*_thread_vs_group.jsonl)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import probnumdiffeq_jl_b200  # noqa: F401
from probnumdiffeq_jl_b200 import lib, problems

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
for name, q, t1, dp in (("rober", 3, 100.0, 0.1), ("orego", 3, 30.0, 0.1), ("vanderpol", 5, 2.0, 0.2), ("fitzhugh_nagumo", 3, 20.0, 0.05),
                        ("lotka_volterra", 3, 10.0, 0.1)):
    pd = problems.PROBLEMS[name]
    rng = np.random.default_rng(3)
    u0 = np.tile(np.array(pd.u0, dtype=float), (n, 1))
    p = np.array(pd.p, dtype=float)[None, :] * (1 + dp * rng.uniform(-1, 1, (n, pd.n_p)))
    res = {}
    for kind in ("group", "thread"):
        os.environ["PNODE_THREAD_KERNEL"] = "0" if kind == "group" else "1"
        cfg = lib.make_config(d=pd.d, q=q, n_p=pd.n_p, rhs_name="builtin:" + name, t0=0.0, t1=t1, lanes_per_traj=0 if kind == "group" else 1)
        try:
            s = lib.Solver(cfg)
            best = None
            for _ in range(3):
                r = s.solve(u0, p)
                ms = r.info.kernel_ms
                out = dict(na=r.field(lib.F_NACCEPT), nr=r.field(lib.F_NREJECT), u=r.field(lib.F_U_FINAL).reshape(n, pd.d), rc=r.field(lib.F_RETCODE),
                           cov=r.field(lib.F_U_FINAL_COV).reshape(n, pd.d, pd.d), ll=r.field(lib.F_LOGLIK), acc=int(r.info.total_accepted),
                           lanes=int(r.info.lanes_per_traj))
                r.free()
                best = ms if best is None else min(best, ms)
            s.close()
        except lib.PnodeError as e:
            print(json.dumps(dict(problem=name, kernel=kind, error=str(e)[:300])), flush=True)
            continue
        res[kind] = out
        rec = dict(problem=name, kernel=kind, lanes=out["lanes"], n_traj=n, kernel_ms=best, accepted=out["acc"], steps_per_s=out["acc"] / (best * 1e-3),
                   ok=bool((out["rc"] == 0).all()))
        if kind == "thread" and "group" in res:
            g = res["group"]
            same = (out["na"] == g["na"]) & (out["nr"] == g["nr"])
            sc = np.abs(g["u"]).max(axis=0)
            rel = np.abs(out["u"] - g["u"]) / sc
            rec.update(same_step_counts_frac=float(same.mean()), max_rel_diff_same=float(rel[same].max()) if same.any() else None,
                       max_rel_diff_all=float(rel.max()),
                       max_rel_cov_same=float((np.abs(out["cov"] - g["cov"])[same].reshape(same.sum(), -1).max(axis=1) /
                                               (np.abs(g["cov"])[same].reshape(same.sum(), -1).max(axis=1) + 1e-300)).max()) if same.any() else None,
                       max_rel_loglik_same=float((np.abs(out["ll"] - g["ll"])[same] / np.abs(g["ll"])[same]).max()) if same.any() else None)
        print(json.dumps(rec), flush=True)
