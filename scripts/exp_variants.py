"""Kernel-variant experiments on one GPU: the same ensemble through NVRTC builds of the step kernel with different
compile-time switches / CTA sizes (PNODE_NVRTC_FLAGS, PNODE_THREADS, PNODE_MIN_BLOCKS), timed with the solver's device
events and checked against the first variant (accept / reject counts identical, end points to 1e-8).

    python scripts/exp_variants.py rober 262144 100.0 3 "name|threads|flags[|lanes[|min_blocks]]" ...

One JSON line per variant (profiles/r2*_variants.jsonl)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import probnumdiffeq_jl_b200  # noqa: F401
from probnumdiffeq_jl_b200 import lib, problems


def main():
    name, n, t1, q = sys.argv[1], int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4])
    variants = [v.split("|") for v in sys.argv[5:]]
    pd = problems.PROBLEMS[name]
    rng = np.random.default_rng(3)
    u0 = np.tile(np.array(pd.u0, dtype=float), (n, 1))
    p = np.array(pd.p, dtype=float)[None, :] * (1 + 0.1 * rng.uniform(-1, 1, (n, pd.n_p)))
    if name == "pleiades":   # initial positions perturbed by 1e-3
        u0[:, 14:] += 1e-3 * rng.uniform(-1, 1, (n, 14))
    smooth = os.environ.get("EXP_SMOOTH", "0") == "1"
    second = os.environ.get("EXP_SECOND", "0") == "1"   # the problem's first-order form as a SecondOrderODEProblem
    ref = None
    for var in variants:
        vname, threads, flags = var[:3]
        lanes = int(var[3]) if len(var) > 3 and var[3] else 0
        os.environ["PNODE_MIN_BLOCKS"] = var[4] if len(var) > 4 and var[4] else "1"
        os.environ["PNODE_FORCE_NVRTC"] = "1"
        os.environ["PNODE_THREADS"] = threads
        os.environ["PNODE_CTA_THREADS"] = threads          # CTA-per-trajectory kernel (D > 32)
        os.environ["PNODE_CTA_MIN_BLOCKS"] = var[4] if len(var) > 4 and var[4] else ""
        if not os.environ["PNODE_CTA_MIN_BLOCKS"]:
            del os.environ["PNODE_CTA_MIN_BLOCKS"]
        os.environ["PNODE_NVRTC_FLAGS"] = flags
        try:
            cfg = lib.make_config(d=pd.d // 2 if second else pd.d, second_order=second, q=q, n_p=pd.n_p, rhs_name="builtin:" + name, t0=0.0, t1=t1, smooth=smooth, save_everystep=smooth, lanes_per_traj=lanes)
            s = lib.Solver(cfg)
            best = None
            for _ in range(3):
                r = s.solve(u0, p)
                ms = r.info.kernel_ms
                out = dict(na=r.field(lib.F_NACCEPT), nr=r.field(lib.F_NREJECT), u=r.field(lib.F_U_FINAL).reshape(n, pd.d),
                           rc=r.field(lib.F_RETCODE), acc=int(r.info.total_accepted))
                r.free()
                best = ms if best is None else min(best, ms)
            s.close()
        except lib.PnodeError as e:
            print(json.dumps(dict(variant=vname, error=str(e)[:300])), flush=True)
            continue
        rec = dict(variant=vname, threads=int(threads), flags=flags, lanes=lanes, problem=name, n_traj=n, kernel_ms=best, accepted=out["acc"],
                   steps_per_s=out["acc"] / (best * 1e-3), ok=bool((out["rc"] == 0).all()))
        if ref is None:
            ref = out
        else:
            same = (out["na"] == ref["na"]) & (out["nr"] == ref["nr"])
            rec["same_step_counts_frac"] = float(same.mean())
            sc = np.abs(ref["u"]).max(axis=0)
            rec["max_rel_diff_same"] = float(np.max(np.abs(out["u"] - ref["u"])[same] / sc)) if same.any() else None
            rec["max_rel_diff_all"] = float(np.max(np.abs(out["u"] - ref["u"]) / sc))
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
