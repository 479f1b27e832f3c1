"""Golden trajectories from the extended-precision textbook reference (tests/dense_reference.py) -- independent of the oracle and of the kernels:

    python tests/golden/make_mp_golden.py        ->  tests/golden/mp_*.npz   (a few seconds per case)

Fixed steps, FixedDiffusion(calibrate=false), so every number is a deterministic function of (problem, u0, p, order, dt): filtered and smoothed
solution means / covariances (sol.u, sol.pu) at every step and the log-likelihood.  tests/test_zz_golden_mp.py compares the oracle (CPU leg) and the
CUDA path through the C ABI (GPU leg) with them."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

import mpmath as mp  # noqa: E402

import probnumdiffeq_jl_b200  # noqa: E402,F401
from dense_reference import dense_filter_smoother  # noqa: E402
from probnumdiffeq_jl_b200 import problems  # noqa: E402

# name: (problem, u0, p, order, alg, dt, steps)
CASES = {
    "mp_fhn_ek1_q3": ("fitzhugh_nagumo", 3, "EK1", "0.01", 200),          # BASELINE cfg1's shape (D = 8), first 200 steps
    "mp_lv_ek0_q3": ("lotka_volterra", 3, "EK0", "0.02", 100),            # cfg2's shape on the Kronecker layout
    "mp_vdp_ek1_q5": ("vanderpol", 5, "EK1", "0.0001", 60),               # cfg3's shape (D = 12, stiff: mu = 1e3)
    "mp_rober_ek1_q3": ("rober", 3, "EK1", "0.001", 80),                  # cfg5's shape (D = 12, d = 3, stiff)
}


def converged(pd, q, alg, dt, n, dps=120):
    """The covariance form is unstable in finite precision (the lost digits grow with the number of steps: 60 digits last about 60 steps of these
    problems), so the working precision is doubled until two successive precisions agree to 1e-25 in every returned number."""
    prev = None
    while True:
        mp.mp.dps = dps
        ref = dense_filter_smoother(pd.f, list(pd.u0), list(pd.p), q, dt, n, alg == "EK1", False)
        if prev is not None:
            err = max(float(np.max(np.abs(ref[k] - prev[k]) / (np.abs(ref[k]) + 1e-300))) for k in ("fm", "fP", "sm", "sP", "ll", "ll_kron", "gd")
                      if np.size(ref[k]))
            # relative to each entry would blow up on structural zeros: compare against the largest entry of the same array instead
            err = max(float(np.max(np.abs(np.asarray(ref[k]) - np.asarray(prev[k]))) / (np.max(np.abs(np.asarray(ref[k]))) + 1e-300))
                      for k in ("fm", "fP", "sm", "sP", "ll", "ll_kron", "gd"))
            if err < 1e-25:
                return ref, dps
        prev, dps = ref, 2 * dps
        if dps > 4000:
            raise RuntimeError("the reference does not converge")


def main():
    for name, (prob, q, alg, dt, n) in CASES.items():
        pd = problems.PROBLEMS[prob]
        ref, dps = converged(pd, q, alg, dt, n)
        d = pd.d
        cfg = dict(problem=prob, q=q, alg=alg, dt=float(dt), steps=n, t1=float(dt) * n, digits=dps)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), config=json.dumps(cfg), u0=np.array(pd.u0, dtype=float), p=np.array(pd.p, dtype=float),
                            filt_u=ref["fm"][:, :d], filt_cov=ref["fP"][:, :d, :d], smooth_u=ref["sm"][:, :d], smooth_cov=ref["sP"][:, :d, :d],
                            loglik=ref["ll"], loglik_kronecker=ref["ll_kron"], global_diffusion=ref["gd"])
        print(name, "written, converged at", dps, "digits")


if __name__ == "__main__":
    main()
