"""Measured rounding floor of the oracle against legitimately perturbed copies of itself.

The ODE filter amplifies one-ulp differences: the local diffusion sigma^2 = z'(HQH')^-1 z is built from a residual z
that is pure cancellation (at the first step z is *only* rounding noise, because the initial state carries the exact
derivatives), and stiff problems amplify further through the dynamics.  No two implementations that are not
bit-identical can agree better than this floor, so parity tests on such configurations assert
`kernel-vs-oracle <= factor * floor` with the floor measured in the same test from
  (a) the oracle's Cholesky vs QR covariance back-ends (what the reference itself switches between, predict.jl:84-91),
  (b) the oracle compiled with and without fused multiply-adds,
  (c) the oracle started from initial derivatives perturbed by ONE ulp in one component.
"""
import ctypes as C

import numpy as np


def variants(oracle, rhs, d, q, n_p, u0, p, cfg_kwargs, ulp_components=None):
    """Yield solutions of the same configuration under rounding-level perturbations."""
    base_backend = cfg_kwargs.get("cov_backend", oracle.COV_QR)
    other = oracle.COV_REF if base_backend == oracle.COV_QR else oracle.COV_QR
    yield "chol-vs-qr", oracle.solve(oracle.make_config(d, q, n_p, **{**cfg_kwargs, "cov_backend": other}), rhs, u0, p)
    oracle.use_variant("fma")
    try:
        yield "fma", oracle.solve(oracle.make_config(d, q, n_p, **cfg_kwargs), rhs, u0, p)
    finally:
        oracle.use_variant("")
    proto = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double, C.c_int)
    orig = proto(("oracle_user_taylor", rhs.so))
    comps = range(d * (q + 1)) if ulp_components is None else ulp_components
    for comp in comps:
        def tay(out, u, pp, t, qq, comp=comp):
            orig(out, u, pp, t, qq)
            out[comp] = np.nextafter(out[comp], np.inf)

        cb = oracle.TAYLOR_FN(tay)

        class _R:
            pass

        r2 = _R()
        r2.f, r2.jac, r2.taylor, r2.so = rhs.f, rhs.jac, C.cast(cb, C.c_void_p), rhs.so
        yield f"ulp[{comp}]", oracle.solve(oracle.make_config(d, q, n_p, **cfg_kwargs), r2, u0, p)


def floor(oracle, rhs, d, q, n_p, u0, p, cfg_kwargs, base, metric, ulp_components=None):
    """max over variants of metric(variant_solution, base_solution); variants whose step pattern differs are skipped."""
    worst = 0.0
    for _, s in variants(oracle, rhs, d, q, n_p, u0, p, cfg_kwargs, ulp_components):
        if s["n"] != base["n"]:
            continue
        worst = max(worst, float(metric(s, base)))
    return worst
