"""Generate the committed golden fixtures of tests/golden/.  Run in the build container (needs /root/reference for part 1):

    python tests/golden/make_golden.py

Part 1 -- reference known-answer values.  The reference is Julia and cannot be executed here, so its
golden vectors are the literal matrices of its own test-suite: this script *parses* the `NAME = [ ... ]` blocks of
/root/reference/test/core/priors.jl:49-127 ("Test IWP (d=2,q=2)"), evaluates the entries (h, sigma as set in that file) and
applies the PERM conjugation written there.  -> reference_iwp_d2_q2.json

Part 2 -- oracle trajectories on small seeded problems (the oracle is pinned against part 1 and the reference's other
tests in tests/test_oracle_pins.py).  They freeze the oracle's behaviour (a regression guard for the checker itself) and
give the GPU tests a fixture that does not depend on the oracle being rebuilt on the GPU box.  -> oracle_*.npz
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

REF = "/root/reference/test/core/priors.jl"


def parse_reference_constants():
    src = open(REF).read()
    h = float(re.search(r"^h\s*=\s*([0-9.eE+-]+)", src, re.M).group(1))
    blk = src[src.index('@testset "Test IWP (d=2,q=2)"'):src.index('@testset "Test SDE"')]
    sigma = float(re.search(r"σ\s*=\s*([0-9.eE+-]+)", blk).group(1))
    env = {"h": h, "sigma": sigma}

    def matrix(name):
        m = re.search(name + r"\s*=\s*(?:σ\^2\s*\.?\*\s*)?\[\n(.*?)\n\s*\]", blk, re.S)
        scaled = bool(re.search(name + r"\s*=\s*σ\^2", blk))
        rows = []
        for line in m.group(1).strip().splitlines():
            rows.append([eval(tok.replace("^", "**"), {}, env) for tok in line.split()])
        a = np.array(rows, dtype=float)
        return sigma**2 * a if scaled else a

    perm = matrix("PERM")
    out = {"source": "test/core/priors.jl:49-127 (parsed, not retyped)", "h": h, "sigma": sigma, "d": 2, "q": 2,
           "PERM": perm.tolist()}
    for name in ("F", "AH_22_IBM", "QH_22_IBM", "AH_22_PRE", "QH_22_PRE"):
        out[name] = (perm @ matrix(name) @ perm.T).tolist()
    out["L"] = (perm @ matrix("L")).tolist()
    return out


CASES = {
    # name: (problem, q, oracle config kwargs, u0 scale, p scale)
    "oracle_fhn_ek1_q3_fixed_smooth": ("fitzhugh_nagumo", 3, dict(alg=1, diffusion=1, calibrate=True, smooth=True, adaptive=False, dt=0.01, t0=0.0, t1=2.0)),
    "oracle_fhn_ek1_q3_dynamic_filter": ("fitzhugh_nagumo", 3, dict(alg=1, diffusion=0, smooth=False, adaptive=False, dt=0.01, t0=0.0, t1=2.0)),
    "oracle_lv_ek0_q3_adaptive": ("lotka_volterra", 3, dict(alg=0, diffusion=0, smooth=False, adaptive=True, t0=0.0, t1=10.0)),
    "oracle_lv_ek1_q3_adaptive_smooth": ("lotka_volterra", 3, dict(alg=1, diffusion=0, smooth=True, adaptive=True, t0=0.0, t1=10.0)),
    "oracle_lv_diagek1_q3_fixedmv": ("lotka_volterra", 3, dict(alg=2, diffusion=3, calibrate=True, smooth=True, adaptive=False, dt=0.02, t0=0.0, t1=4.0)),
    "oracle_lv_ek1_q3_ioup_fixed": ("lotka_volterra", 3, dict(alg=1, diffusion=1, calibrate=True, smooth=True, adaptive=False, dt=0.02, t0=0.0, t1=4.0,
                                                               prior=1, rate_parameter=-1.0)),
    # IOUP with a matrix rate, and with the Jacobian-updated rate (RosenbrockExpEK): the general-prior path
    "oracle_lv_ek1_q3_ioup_matrix_fixed": ("lotka_volterra", 3, dict(alg=1, diffusion=1, calibrate=True, smooth=True, adaptive=False, dt=0.02, t0=0.0,
                                                                      t1=4.0, prior=1, rate_parameter=[[-1.0, 0.3], [-0.2, -0.5]])),
    "oracle_fhn_ek1_q3_ioup_rateupdate_fixed": ("fitzhugh_nagumo", 3, dict(alg=1, diffusion=1, calibrate=True, smooth=True, adaptive=False, dt=0.02,
                                                                           t0=0.0, t1=2.0, prior=1, rate_parameter=None, update_rate_parameter=True)),
    "oracle_rober_ek1_q3_adaptive": ("rober", 3, dict(alg=1, diffusion=0, smooth=False, adaptive=True, t0=0.0, t1=100.0)),
    # mass matrix (dense, non-symmetric), second-order problem, Fenrir data log-likelihood
    "oracle_robermass_ek1_q3_fixed_smooth": ("rober_dae", 3, dict(alg=1, diffusion=1, calibrate=True, smooth=True, adaptive=False, dt=1e-3, t0=0.0,
                                                                  t1=0.2, mass_matrix=[[1.0, 0.2, 0.0], [0.0, 1.0, 0.0], [0.1, 0.0, 0.5]]),
                                             [1.0, 0.1, 0.2], [0.04, 30.0, 10.0]),
    "oracle_twobody2_ek1_q3_fixed_smooth": ("twobody", 3, dict(alg=1, diffusion=1, calibrate=True, smooth=True, adaptive=False, dt=1e-3, t0=0.0,
                                                               t1=0.1, second_order=True)),
    "oracle_lv_ek1_q3_fenrir": ("lotka_volterra", 3, dict(alg=1, diffusion=1, calibrate=False, smooth=True, adaptive=False, dt=0.02, t0=0.0,
                                                          t1=10.0, tstops=[2.5, 5.0, 7.5, 10.0])),
}
FENRIR_DATA = {  # data set of the Fenrir fixture: times, observations (seeded noise around the solution), noise variance
    "oracle_lv_ek1_q3_fenrir": dict(data_t=[2.5, 5.0, 7.5, 10.0], sigma=0.5, seed=7),
}


def oracle_cases(only=None):
    import probnumdiffeq_jl_b200  # noqa: F401
    from probnumdiffeq_jl_b200 import problems, tracer
    from oracle import oracle as O

    for name, case in CASES.items():
        if only and name not in only:
            continue
        prob, q, kw = case[:3]
        pd = problems.lookup(prob)
        rhs = O.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), q)
        u0 = np.array(case[3] if len(case) > 3 else pd.u0, dtype=float)
        p = np.array(case[4] if len(case) > 4 else pd.p, dtype=float)
        d_model = pd.d // 2 if kw.get("second_order") else pd.d
        cfg = O.make_config(d_model, q, pd.n_p, cov_backend=O.COV_QR, **kw)
        s = O.solve(cfg, rhs, u0, p)
        extra = {}
        if name in FENRIR_DATA:
            fd = FENRIR_DATA[name]
            t = np.array(fd["data_t"])
            idx = [int(np.where(s["t"] == tv)[0][0]) for tv in t]
            obs = s["pu_mu"][idx] + fd["sigma"] * np.random.default_rng(fd["seed"]).standard_normal((len(t), pd.d))
            extra = dict(data_t=t, data_u=obs, noise_var=fd["sigma"] ** 2,
                         fenrir_ll=O.fenrir_data_loglik(cfg, s, t, obs, observation_noise_cov=fd["sigma"] ** 2))
        keep = dict(**extra, problem=prob, q=q, config=json.dumps(kw), u0=u0, p=p, t=s["t"], pu_mu=s["pu_mu"], pu_cov=s["pu_cov"],
                    diffusions=s.get("diffusions_mv", s["diffusions"]), loglik=s["loglik"], naccept=s["naccept"], nreject=s["nreject"])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **keep)
        print(name, s["retcode"], s["n"], "points")


if __name__ == "__main__":
    if len(sys.argv) > 1:   # regenerate the named oracle fixtures only
        oracle_cases(set(sys.argv[1:]))
        sys.exit(0)
    if os.path.exists(REF):
        json.dump(parse_reference_constants(), open(os.path.join(HERE, "reference_iwp_d2_q2.json"), "w"), indent=1)
        print("reference_iwp_d2_q2.json")
    else:
        print("skipping part 1: no", REF)
    oracle_cases()
