"""The headline kernel's SOURCE on the CPU: csrc/kernels/pn_thread.cuh (one thread per trajectory: Gram matrix in registers, packed factor,
in-register Cholesky, update, controller -- DESIGN.md section 4.0) is compiled for the host with g++ under tests/emulation/cuda_shim.h and run on
single trajectories.  The kernel gives a whole trajectory to one thread and exchanges nothing with other threads, so a warp of one is a faithful
execution of its arithmetic (the hardware reciprocal / reciprocal-square-root seeds are replaced by seeds of the same 23-bit accuracy; products
and sums are not contracted into FMAs here, so the agreement with the device is to rounding, not bit for bit).

Checked against the oracle (fixed steps 1e-10, adaptive: identical accept / reject counts) and against the extended-precision fixtures
(tests/golden/mp_*.npz).  This runs in the CPU suite: a logic error in the kernel shows up without a GPU.  It is not a product path -- nothing in
probnumdiffeq.jl_b200/ can reach it, and the library still refuses to create a solver without a device."""
import ctypes as C
import json
import os
import subprocess
import tempfile
from math import factorial, sqrt

import numpy as np
import pytest

import probnumdiffeq_jl_b200  # noqa: F401
from probnumdiffeq_jl_b200 import problems, tracer

# a warp / CTA of host threads that disagrees about its collectives deadlocks instead of failing: bound every test
pytestmark = pytest.mark.timeout(600)

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "probnumdiffeq.jl_b200", "csrc")
GOLD = os.path.join(HERE, "golden")
STRUCT = {"pleiades": "PnPleiades", "rober": "PnRober", "fitzhugh_nagumo": "PnFitzHughNagumo", "vanderpol": "PnVanDerPol", "lotka_volterra": "PnLotkaVolterra",
          "orego": "PnOrego"}
dp = C.POINTER(C.c_double)


KERNELS = {"thread": (0, 0, 0), "kron": (1, 0, 0), "blocks_ek0_mv": (2, 0, 1), "blocks_diag1": (2, 1, 0)}   # PN_EMU_KERNEL, _DIAG1, _MV


def _build(problem, q, adaptive, dynamic, kernel="thread"):
    kk, diag1, mv = KERNELS[kernel]
    out = os.path.join(tempfile.gettempdir(), f"pn_emu_{os.getuid()}", f"{kernel}_{problem}_q{q}_a{int(adaptive)}_d{int(dynamic)}.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    src = os.path.join(HERE, "emulation", "pn_thread_host.cpp")
    deps = [src, os.path.join(HERE, "emulation", "cuda_shim.h")] + [os.path.join(CSRC, "kernels", f) for f in
                                                                     ("pn_thread.cuh", "pn_kron.cuh", "pn_blocks.cuh", "pn_small.cuh", "pn_params.h", "pn_builtin_problems.cuh")]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(f) for f in deps):
        subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-ffp-contract=off", "-I", CSRC, f"-DPN_EMU_PROBLEM={STRUCT[problem]}",
                        f"-DPN_EMU_Q={q}", f"-DPN_EMU_ADAPTIVE={int(adaptive)}", f"-DPN_EMU_DYNAMIC={int(dynamic)}", f"-DPN_EMU_KERNEL={kk}",
                        f"-DPN_EMU_DIAG1={diag1}", f"-DPN_EMU_MV={mv}", src, "-o", out], check=True)
    lib = C.CDLL(out)
    lib.pn_emu_solve.restype = None
    return lib


def _iwp_constants(q):
    """L of src/priors/iwp.jl:84-92 (lower triangular), Qb = L'L, and an upper-triangular U with U'U = L'L -- computed here, not taken from
    pnode_api.cu."""
    n = q + 1
    L = np.zeros((n, n))
    for m in range(n):
        for c in range(m + 1):
            L[m, c] = sqrt(2 * q - 2 * m + 1) * factorial(q - c) ** 2 / factorial(m - c) / factorial(2 * q - c - m + 1)
    return L, np.triu(np.linalg.qr(L, mode="r")), L.T @ L


def _run(problem, q, u_init, p, *, t1, dt, adaptive=False, dynamic=False, calibrate=False, abstol=1e-6, reltol=1e-3, forced=None, kernel="thread"):
    lib = _build(problem, q, adaptive, dynamic, kernel)
    d, n = lib.pn_emu_d(), q + 1
    D = d * n
    L, U, Qb = _iwp_constants(q)
    scal = np.array([0.0, t1, dt, abstol, reltol, 0.0, t1, 0.7 / q, 0.4 / q, 0.2, 10.0, 0.9, 1.0, 1.0, 1e-4, 1.0])
    um, uc, xm, xr = np.zeros(d), np.zeros(d * d), np.zeros(D), np.zeros(D * D)
    so, cnt, dmv = np.zeros(4), np.zeros(4, dtype=np.int64), np.zeros(d)
    mu0 = np.ascontiguousarray(u_init, dtype=float)
    pp = np.ascontiguousarray(p if len(p) else [0.0], dtype=float)
    fd = None if forced is None else np.ascontiguousarray(forced, dtype=float)
    A = lambda a: a.ctypes.data_as(dp)  # noqa: E731
    lib.pn_emu_solve(A(scal), int(calibrate), A(np.ascontiguousarray(L.ravel())), A(np.ascontiguousarray(U.ravel())),
                     A(np.ascontiguousarray(Qb.ravel())), A(mu0), A(pp), A(um), A(uc), A(xm), A(xr), A(so), cnt.ctypes.data_as(C.POINTER(C.c_longlong)),
                     A(fd) if fd is not None else None, C.c_longlong(0 if fd is None else len(fd)), A(dmv))
    return dict(diffusion_mv=dmv, u=um, cov=uc.reshape(d, d), x=xm, xR=xr.reshape(D, D), t_end=so[0], loglik=so[1], diffusion=so[2], dt_last=so[3],
                naccept=int(cnt[0]), nreject=int(cnt[1]), nf=int(cnt[2]), retcode=int(cnt[3]))


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b)


@pytest.mark.parametrize("problem,q,t1,dt", [("fitzhugh_nagumo", 3, 2.0, 0.01), ("rober", 3, 0.08, 1e-3), ("vanderpol", 5, 0.006, 1e-4),
                                             ("lotka_volterra", 3, 2.0, 0.02)])
@pytest.mark.parametrize("dynamic", [False, True])
def test_fixed_steps_against_the_oracle(oracle, problem, q, t1, dt, dynamic):
    """Fixed steps; FixedDiffusion(calibrate=false): end state (all derivatives), covariance, log-likelihood at 1e-10.  DynamicDiffusion: the local
    diffusions are ratios of rounding-sensitive residuals, so means 1e-8 and covariances 1e-4 -- the bounds of the GPU tests."""
    pd = problems.PROBLEMS[problem]
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), q)
    o = oracle.solve(oracle.make_config(pd.d, q, pd.n_p, alg=oracle.EK1, diffusion=oracle.DIFF_DYNAMIC if dynamic else oracle.DIFF_FIXED, calibrate=False,
                                        smooth=False, adaptive=False, dt=dt, t0=0.0, t1=t1, save_everystep=True, cov_backend=oracle.COV_REF),
                     rhs, pd.u0, pd.p)
    o2 = oracle.solve(oracle.make_config(pd.d, q, pd.n_p, alg=oracle.EK1, diffusion=oracle.DIFF_DYNAMIC if dynamic else oracle.DIFF_FIXED, calibrate=False,
                                         smooth=False, adaptive=False, dt=dt, t0=0.0, t1=t1, save_everystep=True, cov_backend=oracle.COV_QR),
                      rhs, pd.u0, pd.p)
    k = _run(problem, q, o["x_filt_mu"][0], pd.p, t1=t1, dt=dt, dynamic=dynamic)
    assert k["retcode"] == 0 and k["naccept"] == o["naccept"] and k["t_end"] == o["t"][-1]
    So = o["x_filt_R"][-1].T @ o["x_filt_R"][-1]
    So2 = o2["x_filt_R"][-1].T @ o2["x_filt_R"][-1]
    # the highest derivatives of a stiff problem (Van der Pol: x^(5) ~ 4e9) carry the rounding of the whole run: the full state and its covariance are
    # compared against the oracle's own floor (its two covariance back-ends), the solution components at the fixed bound
    fx, fS = _rel(o2["x_filt_mu"][-1], o["x_filt_mu"][-1]), _rel(So2, So)
    assert _rel(k["u"], o["pu_mu"][-1]) < (1e-8 if dynamic else 1e-10)
    assert _rel(k["x"], o["x_filt_mu"][-1]) < max(1e-8 if dynamic else 1e-10, 20 * fx)
    assert _rel(k["cov"], o["pu_cov"][-1]) < (1e-4 if dynamic else 1e-9)
    assert _rel(k["xR"].T @ k["xR"], So) < max(1e-4 if dynamic else 1e-9, 20 * fS)
    assert k["loglik"] == pytest.approx(o["loglik"], rel=1e-6 if dynamic else 1e-10)


@pytest.mark.parametrize("name", ["mp_fhn_ek1_q3", "mp_rober_ek1_q3", "mp_vdp_ek1_q5"])
def test_fixed_steps_against_the_extended_precision_fixtures(oracle, name):
    """The same kernel source against trajectories that depend on neither the oracle nor the kernels (tests/golden/make_mp_golden.py)."""
    z = np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)
    c = json.loads(str(z["config"]))
    pd = problems.PROBLEMS[c["problem"]]
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), c["q"])
    x0 = oracle.solve(oracle.make_config(pd.d, c["q"], pd.n_p, adaptive=False, dt=c["dt"], t0=0.0, t1=c["dt"], smooth=False, save_everystep=True),
                      rhs, z["u0"], z["p"])["x_filt_mu"][0]          # exact initial derivatives (symbolic)
    k = _run(c["problem"], c["q"], x0, z["p"], t1=c["t1"], dt=c["dt"])
    assert k["retcode"] == 0 and k["naccept"] == c["steps"]
    assert _rel(k["u"], z["filt_u"][-1]) < 1e-10
    assert _rel(k["cov"], z["filt_cov"][-1]) < 1e-8
    assert k["loglik"] == pytest.approx(float(z["loglik"]), rel=1e-10)


@pytest.mark.parametrize("problem,q,t1,dt0", [("rober", 3, 10.0, 1e-4), ("lotka_volterra", 3, 5.0, 1e-2), ("orego", 3, 1.0, 1e-4)])
def test_adaptive_steps_against_the_oracle(oracle, problem, q, t1, dt0):
    """Adaptive steps from a given first step (error estimate, PI controller, reject path -- the kernel's own scalar phase with its reciprocal-based
    bookkeeping, PN_FAST_SCALAR): the oracle's accept / reject counts, end point 1e-8."""
    pd = problems.PROBLEMS[problem]
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), q)
    o = oracle.solve(oracle.make_config(pd.d, q, pd.n_p, alg=oracle.EK1, diffusion=oracle.DIFF_DYNAMIC, smooth=False, adaptive=True, dt=dt0, t0=0.0, t1=t1,
                                        save_everystep=True, cov_backend=oracle.COV_REF), rhs, pd.u0, pd.p)
    k = _run(problem, q, o["x_filt_mu"][0], pd.p, t1=t1, dt=dt0, adaptive=True, dynamic=True)
    assert k["retcode"] == 0 and o["retcode"] == "Success"
    assert (k["naccept"], k["nreject"]) == (o["naccept"], o["nreject"])
    assert k["t_end"] == t1
    assert _rel(k["u"], o["pu_mu"][-1]) < 1e-8


def test_zero_diffusion_takes_the_triangularize_route(oracle):
    """predict.jl:71-74 / :90 (the CPU twin of tests/test_gpu_thread.py::test_thread_kernel_zero_diffusion_takes_the_triangularize_route): with zero
    diffusion, forced on every third accepted step through the kernel's test hook, the Gram + Cholesky route does not apply and the kernel's cold
    path -- Householder triangularisation of the stack in local memory -- runs; everything still agrees with the oracle at 1e-10."""
    pd = problems.PROBLEMS["lotka_volterra"]
    nst = 100
    forced = np.where(np.arange(nst) % 3 == 2, 0.0, 1.0 + 0.1 * np.arange(nst))
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), 3)
    o = oracle.solve(oracle.make_config(2, 3, 4, alg=oracle.EK1, smooth=False, save_everystep=True, cov_backend=oracle.COV_REF, adaptive=False, dt=0.02,
                                        t0=0.0, t1=2.0, forced_diffusion=forced), rhs, pd.u0, pd.p)
    k = _run("lotka_volterra", 3, o["x_filt_mu"][0], pd.p, t1=2.0, dt=0.02, dynamic=True, forced=forced)
    assert k["retcode"] == 0 and k["naccept"] == o["naccept"] == nst
    assert _rel(k["u"], o["pu_mu"][-1]) < 1e-10
    assert _rel(k["cov"], o["pu_cov"][-1]) < 1e-9
    assert k["loglik"] == pytest.approx(o["loglik"], rel=1e-9)


# ---- the structured layouts: pn_kron.cuh (EK0, IsometricKronecker) and pn_blocks.cuh (BlocksOfDiagonals), also one thread per trajectory ----
@pytest.mark.parametrize("q", [3, 8])
def test_kron_kernel_orders(oracle, q):
    """EK0 on the Kronecker layout at order 3 and 8 (beyond order 7 the device build keeps the factor in thread-local memory; the arithmetic is
    the same code; the oracle's symbolic initial derivatives make order 11 too slow for the suite): fixed steps with static diffusion against the oracle -- means 1e-9, covariance against the oracle's own Cholesky-vs-QR floor (order
    q scales the state by h^(q + 1/2)), the log-likelihood with the reference's Kronecker log-determinant -- and an adaptive run: accept / reject counts."""
    pd = problems.PROBLEMS["lotka_volterra"]
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), q)
    kw = dict(alg=oracle.EK0, diffusion=oracle.DIFF_FIXED, calibrate=False, smooth=False, adaptive=False, dt=0.05, t0=0.0, t1=1.0, save_everystep=True)
    o = oracle.solve(oracle.make_config(2, q, 4, cov_backend=oracle.COV_REF, **kw), rhs, pd.u0, pd.p)
    o2 = oracle.solve(oracle.make_config(2, q, 4, cov_backend=oracle.COV_QR, **kw), rhs, pd.u0, pd.p)
    k = _run("lotka_volterra", q, o["x_filt_mu"][0], pd.p, t1=1.0, dt=0.05, kernel="kron")
    assert k["retcode"] == 0 and k["naccept"] == o["naccept"] == 20
    assert _rel(k["u"], o["pu_mu"][-1]) < 1e-9
    assert _rel(k["cov"], o["pu_cov"][-1]) < max(1e-9, 20 * _rel(o2["pu_cov"][-1], o["pu_cov"][-1]))
    assert k["loglik"] == pytest.approx(o["loglik"], rel=1e-8)
    if True:
        o = oracle.solve(oracle.make_config(2, q, 4, alg=oracle.EK0, diffusion=oracle.DIFF_DYNAMIC, smooth=False, adaptive=True, dt=1e-2, t0=0.0, t1=1.0,
                                            save_everystep=True, cov_backend=oracle.COV_REF), rhs, pd.u0, pd.p)
        k = _run("lotka_volterra", q, o["x_filt_mu"][0], pd.p, t1=1.0, dt=1e-2, adaptive=True, dynamic=True, kernel="kron")
        assert k["retcode"] == 0 and k["t_end"] == 1.0
        if q == 3:
            assert (k["naccept"], k["nreject"]) == (o["naccept"], o["nreject"])
            assert _rel(k["u"], o["pu_mu"][-1]) < 1e-7
        else:
            # order 8: a quarter of the attempts is rejected and the controller's exponent 7 / (10 q) moves the step size by a few per cent per step, so one
            # accept / reject decision that falls the other way at rounding level (this build has no FMA contraction, the device has) changes every count
            # after it; both runs are valid solves: same end point within the solver tolerance, step counts of the same size
            assert _rel(k["u"], o["pu_mu"][-1]) < 1e-4 and 0.5 * o["naccept"] < k["naccept"] < 2 * o["naccept"]


@pytest.mark.parametrize("problem,q,t1,dt", [("lotka_volterra", 3, 1.0, 0.02), ("lotka_volterra", 8, 0.5, 0.05), ("pleiades", 3, 0.1, 0.01)])
def test_blocks_kernel_shapes(oracle, problem, q, t1, dt):
    """The blocks layout from the register-resident shape (d (q+1)^2 = 32) to those that live in thread-local memory on the device (162, and
    Pleiades: 448): DiagonalEK1 with FixedDiffusion, fixed steps -- means 1e-10 / covariance 1e-9 / log-likelihood 1e-9 (order 8: the oracle's own
    floor) -- and the EK0 with DynamicMVDiffusion, adaptive: the oracle's step counts and per-component diffusions."""
    pd = problems.PROBLEMS[problem]
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), q)
    kw = dict(alg=oracle.DIAGONAL_EK1, diffusion=oracle.DIFF_FIXED, calibrate=False, smooth=False, adaptive=False, dt=dt, t0=0.0, t1=t1, save_everystep=True)
    o = oracle.solve(oracle.make_config(pd.d, q, pd.n_p, cov_backend=oracle.COV_REF, **kw), rhs, pd.u0, pd.p)
    o2 = oracle.solve(oracle.make_config(pd.d, q, pd.n_p, cov_backend=oracle.COV_QR, **kw), rhs, pd.u0, pd.p)
    k = _run(problem, q, o["x_filt_mu"][0], pd.p, t1=t1, dt=dt, kernel="blocks_diag1")
    assert k["retcode"] == 0 and k["naccept"] == o["naccept"]
    assert _rel(k["u"], o["pu_mu"][-1]) < 1e-10
    assert _rel(k["cov"], o["pu_cov"][-1]) < max(1e-9, 20 * _rel(o2["pu_cov"][-1], o["pu_cov"][-1]))
    # order 8 with the uncalibrated unit diffusion: the log-likelihood (-9.5e8) is z' S^-1 z with a residual z that is a cancellation of O(1) terms; the
    # kernel's preconditioned arithmetic and the oracle's round z differently in the 9th digit
    assert k["loglik"] == pytest.approx(o["loglik"], rel=1e-9 if q <= 3 else 1e-7)
    # first step of the adaptive run: at order 8 a step of dt / 10 = 0.005 puts the residual z ~ h^8 below the rounding level of f, so the per-component
    # diffusion of that step is rounding noise (a 1-ulp change of u0 moves the oracle's own value by 4x) and every later step size with it; start at dt
    dt0 = dt / 10 if q <= 3 else dt
    o = oracle.solve(oracle.make_config(pd.d, q, pd.n_p, alg=oracle.EK0, diffusion=oracle.DIFF_DYNAMIC_MV, smooth=False, adaptive=True, dt=dt0, t0=0.0,
                                        t1=t1, save_everystep=True, cov_backend=oracle.COV_REF), rhs, pd.u0, pd.p)
    k = _run(problem, q, o["x_filt_mu"][0], pd.p, t1=t1, dt=dt0, adaptive=True, dynamic=True, kernel="blocks_ek0_mv")
    assert (k["naccept"], k["nreject"], k["retcode"]) == (o["naccept"], o["nreject"], 0)
    assert _rel(k["u"], o["pu_mu"][-1]) < 1e-8
    assert _rel(k["diffusion_mv"], o["diffusions_mv"][o["naccept"] - 1]) < 1e-5


# ---- the group-per-trajectory kernels as ONE WARP of host threads (cuda_shim.h, PN_SHIM_WARP): pn_small.cuh, pn_smooth_col.cuh, pn_smooth_reg.cuh ----
def _table(name, T):
    """constexpr lookup table in the form pnode_api.cu generates for the NVRTC build (mass_preamble / obsn_preamble)"""
    T = np.asarray(T, dtype=float)
    body = "".join(f"(a == {a} && i == {i}) ? {float(T[a, i])!r} : " for a in range(T.shape[0]) for i in range(T.shape[1]) if T[a, i] != 0.0)
    return f"constexpr double {name}(int a, int i) {{\n  return {body}0.0;\n}}\n"


def _build_warp(problem, q, G, GS, adaptive, dynamic, save, tsan=False, cta=False, defs=(), preamble=None):
    tag = "".join("_" + x.replace("=", "") for x in defs)
    pre = []
    if preamble is not None:  # generated source that precedes the kernel headers (what pnode_api.cu prepends to the NVRTC program)
        import hashlib
        h = hashlib.sha1(preamble.encode()).hexdigest()[:10]
        tag += "_pre" + h
        ppath = os.path.join(tempfile.gettempdir(), f"pn_emu_{os.getuid()}", f"preamble_{h}.h")
        os.makedirs(os.path.dirname(ppath), exist_ok=True)
        if not os.path.exists(ppath):
            with open(ppath, "w") as fh:
                fh.write(preamble)
        pre = ["-include", ppath]
    if problem not in STRUCT:  # a traced problem of the library (mass-matrix / second-order tables): generated source, struct UserProb
        pdef = problems.lookup(problem)
        upath = os.path.join(tempfile.gettempdir(), f"pn_emu_{os.getuid()}", f"user_{problem}.cuh")
        os.makedirs(os.path.dirname(upath), exist_ok=True)
        text = tracer.cuda_source(tracer.trace(pdef.f, len(pdef.u0), len(pdef.p)), "UserProb")
        if not os.path.exists(upath) or open(upath).read() != text:
            with open(upath, "w") as fh:
                fh.write(text)
        pre += [f'-DPN_EMU_USER_SOURCE="{upath}"']
    out = os.path.join(tempfile.gettempdir(), f"pn_emu_{os.getuid()}",
                       f"{('warp', 'cta', 'gen')[int(cta)]}_{problem}_q{q}_g{G}_gs{GS}_a{int(adaptive)}_d{int(dynamic)}_s{save}{tag}{'_tsan' if tsan else ''}" + (".exe" if tsan else ".so"))
    os.makedirs(os.path.dirname(out), exist_ok=True)
    src = os.path.join(HERE, "emulation", "pn_warp_host.cpp")
    deps = [src, os.path.join(HERE, "emulation", "cuda_shim.h")] + [os.path.join(CSRC, "kernels", f) for f in
                                                                     ("pn_small.cuh", "pn_cta.cuh", "pn_smooth_col.cuh", "pn_smooth_reg.cuh", "pn_smooth.cuh",
                                                                      "pn_params.h", "pn_builtin_problems.cuh")]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(f) for f in deps):
        mode = ["-fsanitize=thread", "-g", "-DPN_EMU_MAIN=1"] if tsan else ["-fPIC", "-shared"]
        subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", *mode, "-I", CSRC, f"-DPN_EMU_PROBLEM={STRUCT.get(problem, 'UserProb')}", f"-DPN_EMU_Q={q}",
                        f"-DPN_EMU_G={G}", f"-DPN_EMU_GS={GS}", f"-DPN_EMU_ADAPTIVE={int(adaptive)}", f"-DPN_EMU_DYNAMIC={int(dynamic)}",
                        f"-DPN_EMU_SAVE={save}", f"-DPN_EMU_KERNEL={int(cta)}", *["-D" + x for x in defs], *pre, src, "-o", out, "-lpthread"], check=True)
    return out


def _run_warp(problem, q, G, GS, init_mu, ps, offsets, *, t1, dt, adaptive=False, dynamic=False, calibrate=False, save=2, smooth=0, write_x=1, lanes=32,
              abstol=1e-6, reltol=1e-3, cta=False, defs=(), rate=None, update_rate=False, ioup_rate=None, fenrir=None, preamble=None, data=None, saveat=None):
    lib = C.CDLL(_build_warp(problem, q, G, GS, adaptive, dynamic, save, cta=cta, defs=defs, preamble=preamble))
    lib.pn_emu_warp_solve.restype = None
    if rate is not None:
        rate = np.ascontiguousarray(rate, dtype=float)
        lib.pn_emu_set_rate(rate.ctypes.data_as(dp), int(update_rate))
    elif update_rate:
        lib.pn_emu_set_rate(None, 1)
    keep = []
    if fenrir is not None:  # (data_t, data_u [n_data][n_obs], H [n_obs][D], Rn [n_obs][n_obs]) -> o["fenrir_ll"]
        ft, fu, fH, fR = [np.ascontiguousarray(a, dtype=float) for a in fenrir]
        fll = np.full(len(init_mu), np.nan)
        keep += [ft, fu, fH, fR, fll]
        lib.pn_emu_set_fenrir.argtypes = [dp, dp, C.c_longlong, C.c_int, dp, dp, dp, C.c_int]
        lib.pn_emu_set_fenrir(ft.ctypes.data_as(dp), fu.ctypes.data_as(dp), len(ft), fu.shape[1], fH.ctypes.data_as(dp), fR.ctypes.data_as(dp),
                              fll.ctypes.data_as(dp), 0)
    if data is not None:  # forward data updates: (data_t, data_u [n_data][n_obs]) -> o["data_ll"]; the build needs a PN_DATA preamble
        dt_, du_ = [np.ascontiguousarray(a, dtype=float) for a in data]
        dll = np.full(len(init_mu), np.nan)
        keep += [dt_, du_, dll]
        lib.pn_emu_set_data.argtypes = [dp, dp, C.c_longlong, dp]
        lib.pn_emu_set_data(dt_.ctypes.data_as(dp), du_.ctypes.data_as(dp), len(dt_), dll.ctypes.data_as(dp))
    if saveat is not None:  # dense output of a filter-only solve -> o["dense_u"] [N][n_saveat][d], o["dense_cov"]
        sa = np.ascontiguousarray(saveat, dtype=float)
        dd_ = lib.pn_emu_d()
        du_out, dc_out = np.full((len(init_mu), len(sa), dd_), np.nan), np.full((len(init_mu), len(sa), dd_, dd_), np.nan)
        keep += [sa, du_out, dc_out]
        lib.pn_emu_set_saveat.argtypes = [dp, C.c_longlong, dp, dp]
        lib.pn_emu_set_saveat(sa.ctypes.data_as(dp), len(sa), du_out.ctypes.data_as(dp), dc_out.ctypes.data_as(dp))
    if ioup_rate is not None:  # scalar-rate IOUP of the register kernels (defs must hold PN_IOUP=1): 16-point Gauss-Legendre rule on [0, 1]
        x, w = np.polynomial.legendre.leggauss(16)
        lib.pn_emu_set_ioup.argtypes = [C.c_double, dp, dp]
        lib.pn_emu_set_ioup(float(ioup_rate), np.ascontiguousarray(0.5 * (x + 1)).ctypes.data_as(dp), np.ascontiguousarray(0.5 * w).ctypes.data_as(dp))
    d, n = lib.pn_emu_d(), q + 1            # d: length of sol.u (for a second-order problem (du, u): twice the modelled dimension)
    D = len(init_mu[0])
    N = len(init_mu)
    L, U, Qb = _iwp_constants(q)
    scal = np.array([0.0, t1, dt, abstol, reltol, 0.0, t1, 0.7 / q, 0.4 / q, 0.2, 10.0, 0.9, 1.0, 1.0, 1e-4, 1.0])
    off = np.ascontiguousarray(offsets, dtype=np.int64)
    T = int(off[-1])
    o = dict(u=np.zeros((N, d)), cov=np.zeros((N, d, d)), scal=np.zeros((N, 4)), counts=np.zeros((N, 4), dtype=np.int64), s_t=np.full(T, np.nan),
             s_u=np.full((T, d), np.nan), s_cov=np.full((T, d, d), np.nan), s_diff=np.full(T, np.nan), s_x=np.full((T, D), np.nan),
             s_xR=np.full((T, D, D), np.nan), s_h=np.zeros(T))
    mu0 = np.ascontiguousarray(init_mu, dtype=float)
    dt0 = np.full(N, float(dt))
    pp = np.ascontiguousarray(ps, dtype=float)
    A = lambda a: a.ctypes.data_as(dp)  # noqa: E731
    I = lambda a: a.ctypes.data_as(C.POINTER(C.c_longlong))  # noqa: E731
    lib.pn_emu_warp_solve(C.c_longlong(N), int(lanes), A(scal), int(calibrate), A(np.ascontiguousarray(L.ravel())), A(np.ascontiguousarray(U.ravel())),
                          A(np.ascontiguousarray(Qb.ravel())), A(mu0), A(dt0), A(pp), A(o["u"]), A(o["cov"]), A(o["scal"]), I(o["counts"]), I(off),
                          A(o["s_t"]), A(o["s_u"]), A(o["s_cov"]), A(o["s_diff"]), A(o["s_x"]), A(o["s_xR"]), A(o["s_h"]), int(smooth), int(write_x))
    if fenrir is not None:
        o["fenrir_ll"] = keep[4]
    if data is not None:
        o["data_ll"] = dll
    if saveat is not None:
        o["dense_u"], o["dense_cov"] = du_out, dc_out
    return o


def _ensemble(pd, N, seed=0):
    rng = np.random.default_rng(seed)
    u0 = np.array(pd.u0, dtype=float)[None, :] * (1.0 + 0.05 * rng.standard_normal((N, pd.d)))
    p = np.array(pd.p, dtype=float)[None, :] * (1.0 + 0.05 * rng.standard_normal((N, pd.n_p)))
    return u0, p


@pytest.mark.parametrize("problem,q,G,GS,t1,dt,N", [("fitzhugh_nagumo", 3, 4, 8, 1.0, 0.05, 11), ("vanderpol", 5, 8, 8, 0.01, 2.5e-4, 6)])
@pytest.mark.parametrize("sweep", [1, 2, 3])
def test_group_step_kernel_and_register_sweeps_as_a_warp(oracle, problem, q, G, GS, t1, dt, N, sweep):
    """pn_small.cuh (saving pass, SAVE = 2) and the register-resident sweeps pn_smooth_col.cuh (sweep 1, the default) / pn_smooth_reg.cuh (sweep 2) /
    the warp-per-trajectory sweep with the reference's formulas out of shared memory, pn_smooth.cuh (sweep 3: four warps of host threads), on the
    BASELINE shapes cfg1 (FitzHugh-Nagumo, EK1(3), G = 4 / 8) and cfg3 (Van der Pol, EK1(5): the instantiation <d = 2, q = 5, G = 8>), run as one
    warp of 32 host threads over an ensemble larger than the 32 / G groups (so the groups pull from the work queue and drain raggedly), fixed steps,
    FixedDiffusion without calibration: every filtered and smoothed mean 1e-10 against the oracle, covariances against the oracle's own
    Cholesky-vs-QR floor, and the log-likelihood."""
    pd = problems.PROBLEMS[problem]
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), q)
    u0, p = _ensemble(pd, N)
    kw = dict(alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, calibrate=False, adaptive=False, dt=dt, t0=0.0, t1=t1, save_everystep=True)
    of, os_, os2 = [], [], []
    for i in range(N):
        of.append(oracle.solve(oracle.make_config(pd.d, q, pd.n_p, smooth=False, cov_backend=oracle.COV_REF, **kw), rhs, u0[i], p[i]))
        os_.append(oracle.solve(oracle.make_config(pd.d, q, pd.n_p, smooth=True, cov_backend=oracle.COV_REF, **kw), rhs, u0[i], p[i]))
        os2.append(oracle.solve(oracle.make_config(pd.d, q, pd.n_p, smooth=True, cov_backend=oracle.COV_QR, **kw), rhs, u0[i], p[i]))
    ns = of[0]["n"]
    off = np.arange(N + 1) * ns
    mu0 = np.array([o["x_filt_mu"][0] for o in of])
    kf = _run_warp(problem, q, G, GS, mu0, p, off, t1=t1, dt=dt, smooth=0)
    ks = _run_warp(problem, q, G, GS, mu0, p, off, t1=t1, dt=dt, smooth=sweep)
    for i in range(N):
        sl = slice(off[i], off[i + 1])
        assert tuple(kf["counts"][i]) == (ns - 1, 0, of[i]["nf"], 0)
        assert np.allclose(kf["s_t"][sl], of[i]["t"], rtol=0, atol=1e-13)
        assert _rel(kf["s_u"][sl], of[i]["pu_mu"]) < 1e-10 and _rel(kf["s_x"][sl], of[i]["x_filt_mu"]) < 1e-10
        assert _rel(kf["s_cov"][sl][1:], of[i]["pu_cov"][1:]) < 1e-9
        assert kf["scal"][i, 1] == pytest.approx(of[i]["loglik"], rel=1e-10)
        assert _rel(ks["s_u"][sl], os_[i]["pu_mu"]) < 1e-10
        floor = _rel(os2[i]["pu_cov"][1:], os_[i]["pu_cov"][1:])
        assert _rel(ks["s_cov"][sl][1:], os_[i]["pu_cov"][1:]) < max(1e-9, 20 * floor)
        # the smoothed state records (write_x): means, and the factors through R'R
        assert _rel(ks["s_x"][sl], os_[i]["x_smooth_mu"]) < max(1e-9, 20 * _rel(os2[i]["x_smooth_mu"], os_[i]["x_smooth_mu"]))


def test_group_step_kernel_adaptive_ragged_ensemble(oracle):
    """pn_small.cuh, adaptive with DynamicDiffusion, nothing saved, on ROBER (G = 4: eight trajectories in flight, 19 in the queue with different
    parameters and therefore different step counts): per-trajectory accept / reject / nf counts identical with the oracle's, end states 1e-8."""
    pd = problems.PROBLEMS["rober"]
    q, N = 3, 19
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), q)
    u0, p = _ensemble(pd, N, seed=3)
    u0[:] = np.array(pd.u0)[None, :]
    oo = [oracle.solve(oracle.make_config(pd.d, q, pd.n_p, alg=oracle.EK1, diffusion=oracle.DIFF_DYNAMIC, smooth=False, adaptive=True, dt=1e-4, t0=0.0, t1=10.0,
                                          save_everystep=True, cov_backend=oracle.COV_REF), rhs, u0[i], p[i]) for i in range(N)]
    mu0 = np.array([o["x_filt_mu"][0] for o in oo])
    k = _run_warp("rober", q, 4, 8, mu0, p, np.arange(N + 1), t1=10.0, dt=1e-4, adaptive=True, dynamic=True, save=0)
    assert len({o["naccept"] for o in oo}) > 3  # ragged
    for i in range(N):
        assert tuple(k["counts"][i]) == (oo[i]["naccept"], oo[i]["nreject"], oo[i]["nf"], 0)
        assert k["scal"][i, 0] == 10.0
        assert _rel(k["u"][i], oo[i]["pu_mu"][-1]) < 1e-8
        assert k["scal"][i, 1] == pytest.approx(oo[i]["loglik"], rel=1e-8)


def _tsan_run(problem, q, G, GS, *, N, t1, dt, adaptive, dynamic, save, smooth, oracle, drop_syncwarp=False, lanes=32, cta=False, defs=(), okw=None,
              extra=(), preamble=None, p_override=None):
    pd = problems.lookup(problem)
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), q)
    u0, p = _ensemble(pd, N, seed=5)
    if p_override is not None:
        u0, p = p_override
    if pd.n_p == 0:
        p = np.zeros((N, 1))
    kw = dict(alg=oracle.EK1, diffusion=oracle.DIFF_DYNAMIC if dynamic else oracle.DIFF_FIXED, calibrate=False, adaptive=adaptive, dt=dt, t0=0.0, t1=t1,
              save_everystep=True, smooth=False, **(okw or {}))
    assert not (adaptive and save)  # ragged CSR segments are not laid out here
    oo = [oracle.solve(oracle.make_config(pd.d, q, pd.n_p, **kw), rhs, u0[i], p[i] if pd.n_p else []) for i in range(N)]
    L, U, Qb = _iwp_constants(q)
    scal = np.array([0.0, t1, dt, 1e-6, 1e-3, 0.0, t1, 0.7 / q, 0.4 / q, 0.2, 10.0, 0.9, 1.0, 1.0, 1e-4, 1.0])
    blob = np.concatenate([[N, lanes, smooth, 1, oo[0]["n"] if save else 1], scal, L.ravel(), U.ravel(), Qb.ravel(),
                           np.array([o["x_filt_mu"][0] for o in oo]).ravel(), np.full(N, dt), p.ravel(),
                           *[np.asarray(e, dtype=float).ravel() for e in extra]]).astype(np.float64)
    exe = _build_warp(problem, q, G, GS, adaptive, dynamic, save, tsan=True, cta=cta, defs=defs, preamble=preamble)
    if drop_syncwarp:
        exe2 = exe.replace(".exe", "_nosync.exe")
        if not os.path.exists(exe2) or os.path.getmtime(exe2) < os.path.getmtime(exe):
            subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-fsanitize=thread", "-g", "-DPN_EMU_MAIN=1", "-DPN_SHIM_DROP_SYNCWARP=1", "-I", CSRC,
                            f"-DPN_EMU_PROBLEM={STRUCT.get(problem, 'UserProb')}", f"-DPN_EMU_Q={q}", f"-DPN_EMU_G={G}", f"-DPN_EMU_GS={GS}",
                            f"-DPN_EMU_ADAPTIVE={int(adaptive)}", f"-DPN_EMU_DYNAMIC={int(dynamic)}", f"-DPN_EMU_SAVE={save}",
                            os.path.join(HERE, "emulation", "pn_warp_host.cpp"), "-o", exe2, "-lpthread"], check=True)
        exe = exe2
    path = os.path.join(os.path.dirname(exe), f"tsan_in_{problem}_{q}_{save}_{smooth}_{int(cta)}{len(extra)}.bin")
    blob.tofile(path)
    r = subprocess.run([exe, path], capture_output=True, text=True, timeout=900, env=dict(os.environ, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0"))
    return r, sum(o["naccept"] for o in oo)


@pytest.mark.parametrize("problem,q,G,GS,t1,dt,smooth", [("fitzhugh_nagumo", 3, 4, 8, 0.5, 0.05, 1), ("fitzhugh_nagumo", 3, 4, 8, 0.5, 0.05, 2),
                                                       ("fitzhugh_nagumo", 3, 4, 8, 0.5, 0.05, 3), ("vanderpol", 5, 8, 8, 2.5e-3, 2.5e-4, 1)])
def test_threadsanitizer_racecheck_of_the_exchange_protocol(oracle, problem, q, G, GS, t1, dt, smooth):
    """Racecheck on the host (compute-sanitizer is closed on the GPU pool): the step kernel (saving pass: Gram + Cholesky through the group's exchange
    region) and the register sweeps (record staging with cp.async double buffering, smoothed-record write-back) as one warp of host threads under
    ThreadSanitizer.  Every cross-lane ordering the device guarantees (shuffles, votes, __syncwarp) is a pthread barrier, so a shared- or global-memory
    access pair of two lanes that the kernel does not separate by one of them is reported.  Clean run: exit code 0, no report, all steps accepted."""
    r, nacc = _tsan_run(problem, q, G, GS, N=9, t1=t1, dt=dt, adaptive=False, dynamic=False, save=2, smooth=smooth, oracle=oracle)
    assert "ThreadSanitizer" not in r.stderr, r.stderr[:4000]
    assert r.returncode == 0 and f"accepted {nacc} " in r.stdout


def test_threadsanitizer_sees_a_missing_syncwarp(oracle):
    """Negative control: the same program with the kernels' __syncwarp() calls compiled out must be reported as racy (the check has teeth)."""
    r, _ = _tsan_run("fitzhugh_nagumo", 3, 4, 8, N=9, t1=0.5, dt=0.05, adaptive=False, dynamic=False, save=2, smooth=1, oracle=oracle, drop_syncwarp=True)
    assert "ThreadSanitizer: data race" in r.stderr and r.returncode != 0


# ---- the CTA-per-trajectory kernel (pn_cta.cuh: D > 32, Pleiades D = 112) as a CTA of 256 host threads ----
def test_cta_kernel_pleiades_as_256_host_threads(oracle):
    """pn_cta.cuh on Pleiades EK1(3), D = 112 (BASELINE cfg4): Gram + blocked Cholesky on register tiles out of shared memory, a chain of
    __syncthreads() per step.  Three trajectories through the work queue, fixed steps with FixedDiffusion: end state 1e-10 / covariance 1e-9 /
    log-likelihood 1e-9 against the oracle (the reference's route, COV_REF); adaptive with DynamicDiffusion: the oracle's accept / reject counts."""
    pd = problems.PROBLEMS["pleiades"]
    q, N = 3, 3
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), q)
    u0, p = _ensemble(pd, N, seed=7)
    p = np.zeros((N, 1))
    kw = dict(alg=oracle.EK1, smooth=False, t0=0.0, save_everystep=True, cov_backend=oracle.COV_REF)
    oo = [oracle.solve(oracle.make_config(pd.d, q, pd.n_p, diffusion=oracle.DIFF_FIXED, calibrate=False, adaptive=False, dt=0.01, t1=0.06, **kw), rhs, u0[i], [])
          for i in range(N)]
    mu0 = np.array([o["x_filt_mu"][0] for o in oo])
    k = _run_warp("pleiades", q, 4, 8, mu0, p, np.arange(N + 1), t1=0.06, dt=0.01, save=0, cta=True)
    for i in range(N):
        assert tuple(k["counts"][i]) == (oo[i]["naccept"], 0, oo[i]["nf"], 0)
        assert _rel(k["u"][i], oo[i]["pu_mu"][-1]) < 1e-10
        assert _rel(k["cov"][i], oo[i]["pu_cov"][-1]) < 1e-9
        assert k["scal"][i, 1] == pytest.approx(oo[i]["loglik"], rel=1e-9)
    oa = [oracle.solve(oracle.make_config(pd.d, q, pd.n_p, diffusion=oracle.DIFF_DYNAMIC, adaptive=True, dt=1e-3, t1=0.3, **kw), rhs, u0[i], []) for i in range(N)]
    k = _run_warp("pleiades", q, 4, 8, mu0, p, np.arange(N + 1), t1=0.3, dt=1e-3, adaptive=True, dynamic=True, save=0, cta=True)
    for i in range(N):
        assert tuple(k["counts"][i]) == (oa[i]["naccept"], oa[i]["nreject"], oa[i]["nf"], 0)
        assert _rel(k["u"][i], oa[i]["pu_mu"][-1]) < 1e-8


def test_threadsanitizer_racecheck_of_the_cta_kernel(oracle):
    """pn_cta.cuh (Pleiades, D = 112) as a CTA of 256 host threads under ThreadSanitizer: every hand-over between threads of the step (panel
    factorisation -> trailing update, C2 / S / gain, the scalar bookkeeping in PnCtaShared) must be separated by one of the kernel's __syncthreads()."""
    r, nacc = _tsan_run("pleiades", 3, 4, 8, N=2, t1=0.03, dt=0.01, adaptive=False, dynamic=False, save=0, smooth=0, oracle=oracle, cta=True)
    assert "ThreadSanitizer" not in r.stderr, r.stderr[:4000]
    assert r.returncode == 0 and f"accepted {nacc} " in r.stdout
    # the saving pass (filtered records written by the CTA) followed by the large-state sweep out of the global workspace (four warps)
    r, nacc = _tsan_run("pleiades", 3, 4, 8, N=5, t1=0.03, dt=0.01, adaptive=False, dynamic=False, save=2, smooth=3, oracle=oracle, cta=True)
    assert "ThreadSanitizer" not in r.stderr, r.stderr[:4000]
    assert r.returncode == 0 and f"accepted {nacc} " in r.stdout


@pytest.mark.parametrize("problem,cta,N,t1,dt", [("rober", False, 11, 1.0, 1e-4), ("pleiades", True, 2, 0.1, 1e-3)])
def test_threadsanitizer_racecheck_adaptive(oracle, problem, cta, N, t1, dt):
    """The adaptive route (rejected attempts, the controller, DynamicDiffusion; ROBER on pn_small: a ragged ensemble draining through the work queue;
    Pleiades on pn_cta) under ThreadSanitizer: clean, and the total of accepted steps is the oracle's."""
    r, nacc = _tsan_run(problem, 3, 4, 8, N=N, t1=t1, dt=dt, adaptive=True, dynamic=True, save=0, smooth=0, oracle=oracle, cta=cta)
    assert "ThreadSanitizer" not in r.stderr, r.stderr[:4000]
    assert r.returncode == 0 and f"accepted {nacc} " in r.stdout


def test_cta_kernel_saving_pass_and_large_state_sweep(oracle):
    """BASELINE cfg4 with smoothing: Pleiades EK1(3), D = 112.  pn_cta.cuh's saving pass (SAVE = 2: the filtered records) as 256 host threads, then
    pn_smooth.cuh as four warps sweeping out of a global workspace (7 D^2 doubles per warp do not fit in shared memory), fixed steps with
    FixedDiffusion: filtered and smoothed means 1e-10, covariances 1e-9 / the oracle's own Cholesky-vs-QR floor."""
    pd = problems.PROBLEMS["pleiades"]
    q, N, dt, t1 = 3, 5, 0.01, 0.04
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), q)
    u0, _ = _ensemble(pd, N, seed=11)
    p = np.zeros((N, 1))
    kw = dict(alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, calibrate=False, adaptive=False, dt=dt, t0=0.0, t1=t1, save_everystep=True)
    of = [oracle.solve(oracle.make_config(pd.d, q, pd.n_p, smooth=False, cov_backend=oracle.COV_REF, **kw), rhs, u0[i], []) for i in range(N)]
    os_ = [oracle.solve(oracle.make_config(pd.d, q, pd.n_p, smooth=True, cov_backend=oracle.COV_REF, **kw), rhs, u0[i], []) for i in range(N)]
    os2 = [oracle.solve(oracle.make_config(pd.d, q, pd.n_p, smooth=True, cov_backend=oracle.COV_QR, **kw), rhs, u0[i], []) for i in range(N)]
    ns = of[0]["n"]
    off = np.arange(N + 1) * ns
    mu0 = np.array([o["x_filt_mu"][0] for o in of])
    kf = _run_warp("pleiades", q, 4, 8, mu0, p, off, t1=t1, dt=dt, save=2, smooth=0, cta=True)
    ks = _run_warp("pleiades", q, 4, 8, mu0, p, off, t1=t1, dt=dt, save=2, smooth=3, cta=True)
    for i in range(N):
        sl = slice(off[i], off[i + 1])
        assert tuple(kf["counts"][i]) == (ns - 1, 0, of[i]["nf"], 0)
        assert _rel(kf["s_u"][sl], of[i]["pu_mu"]) < 1e-10 and _rel(kf["s_x"][sl], of[i]["x_filt_mu"]) < 1e-10
        assert _rel(kf["s_cov"][sl][1:], of[i]["pu_cov"][1:]) < 1e-9
        assert _rel(ks["s_u"][sl], os_[i]["pu_mu"]) < 1e-10
        assert _rel(ks["s_cov"][sl][1:], os_[i]["pu_cov"][1:]) < max(1e-9, 20 * _rel(os2[i]["pu_cov"][1:], os_[i]["pu_cov"][1:]))
        assert _rel(ks["s_x"][sl], os_[i]["x_smooth_mu"]) < max(1e-9, 20 * _rel(os2[i]["x_smooth_mu"], os_[i]["x_smooth_mu"]))


def test_cta_kernel_pleiades_as_14_second_order_equations(oracle):
    """pn_cta.cuh built with PN_SECOND (benchmarks/pleiades.jmd's own form: SecondOrderODEProblem, d = 14, D = 56; H = [-J_u | -J_du | I | 0],
    measurement_models.jl:29-49) as 256 host threads: fixed steps 1e-10 on sol.u = (du, u) / 1e-9 on its covariance / log-likelihood, and the
    adaptive run's accept / reject counts, against the oracle's second-order route."""
    pd = problems.PROBLEMS["pleiades"]
    q, N = 3, 3
    rhs = oracle.compile_rhs(tracer.trace(pd.f, 28, 0), q)
    u0, _ = _ensemble(pd, N, seed=13)
    p = np.zeros((N, 1))
    kw = dict(alg=oracle.EK1, smooth=False, t0=0.0, save_everystep=True, cov_backend=oracle.COV_REF, second_order=True)
    oo = [oracle.solve(oracle.make_config(14, q, 0, diffusion=oracle.DIFF_FIXED, calibrate=False, adaptive=False, dt=0.01, t1=0.06, **kw), rhs, u0[i], [])
          for i in range(N)]
    mu0 = np.array([o["x_filt_mu"][0] for o in oo])
    assert mu0.shape == (N, 56)
    k = _run_warp("pleiades", q, 4, 8, mu0, p, np.arange(N + 1), t1=0.06, dt=0.01, save=0, cta=True, defs=("PN_SECOND=1",))
    for i in range(N):
        assert tuple(k["counts"][i]) == (oo[i]["naccept"], 0, oo[i]["nf"], 0)
        assert _rel(k["u"][i], oo[i]["pu_mu"][-1]) < 1e-10
        assert _rel(k["cov"][i], oo[i]["pu_cov"][-1]) < 1e-9
        assert k["scal"][i, 1] == pytest.approx(oo[i]["loglik"], rel=1e-9)
    oa = [oracle.solve(oracle.make_config(14, q, 0, diffusion=oracle.DIFF_DYNAMIC, adaptive=True, dt=1e-3, t1=0.3, **kw), rhs, u0[i], []) for i in range(N)]
    k = _run_warp("pleiades", q, 4, 8, mu0, p, np.arange(N + 1), t1=0.3, dt=1e-3, adaptive=True, dynamic=True, save=0, cta=True, defs=("PN_SECOND=1",))
    for i in range(N):
        assert tuple(k["counts"][i]) == (oa[i]["naccept"], oa[i]["nreject"], oa[i]["nf"], 0)
        assert _rel(k["u"][i], oa[i]["pu_mu"][-1]) < 1e-8


@pytest.mark.parametrize("rate,update", [(np.array([[-1.0, 0.3], [-0.2, -0.5]]), False), (np.diag([-1.0, -0.4]), False), (None, True)])
def test_general_prior_kernel_ioup_rates(oracle, rate, update):
    """pn_gen.cuh + pn_lti.cuh (IOUP with a matrix rate, a vector rate as its diagonal matrix, and the Jacobian-updated rate of the
    RosenbrockExpEK; src/priors/ioup.jl:81-131, perform_step.jl:88-96) as four warps of host threads, one trajectory per warp: fixed steps with
    FixedDiffusion against the oracle (whose transition pairs come from a Pade exponential of the Van Loan block matrix, not from scaling and
    doubling) -- every filtered point 1e-10 / covariance 1e-9 / log-likelihood 1e-9, and with a constant rate the smoothed points through
    pn_smooth.cuh's dense-transition mode."""
    pd = problems.PROBLEMS["lotka_volterra"]
    q, N, dt, t1 = 3, 6, 0.02, 0.4
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), q)
    u0, p = _ensemble(pd, N, seed=17)
    okw = dict(alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, calibrate=False, adaptive=False, dt=dt, t0=0.0, t1=t1, prior=oracle.PRIOR_IOUP,
               cov_backend=oracle.COV_REF, save_everystep=True)
    if update:
        okw["update_rate_parameter"] = True
    else:
        okw["rate_parameter"] = rate if rate[0, 1] != 0 else np.diag(rate).copy()
    of = [oracle.solve(oracle.make_config(2, q, 4, smooth=False, **okw), rhs, u0[i], p[i]) for i in range(N)]
    ns = of[0]["n"]
    off = np.arange(N + 1) * ns
    mu0 = np.array([o["x_filt_mu"][0] for o in of])
    kf = _run_warp("lotka_volterra", q, 4, 8, mu0, p, off, t1=t1, dt=dt, save=2, smooth=0, cta=2, rate=rate, update_rate=update)
    for i in range(N):
        sl = slice(off[i], off[i + 1])
        assert tuple(kf["counts"][i][[0, 1, 3]]) == (ns - 1, 0, 0)
        assert _rel(kf["s_u"][sl], of[i]["pu_mu"]) < 1e-10
        assert _rel(kf["s_cov"][sl][1:], of[i]["pu_cov"][1:]) < 1e-9
        assert kf["scal"][i, 1] == pytest.approx(of[i]["loglik"], rel=1e-9)
    if not update:
        os_ = [oracle.solve(oracle.make_config(2, q, 4, smooth=True, **okw), rhs, u0[i], p[i]) for i in range(N)]
        ks = _run_warp("lotka_volterra", q, 4, 8, mu0, p, off, t1=t1, dt=dt, save=2, smooth=3, cta=2, rate=rate)
        for i in range(N):
            sl = slice(off[i], off[i + 1])
            assert _rel(ks["s_u"][sl], os_[i]["pu_mu"]) < 1e-10
            assert _rel(ks["s_cov"][sl][1:], os_[i]["pu_cov"][1:]) < 1e-9


@pytest.mark.parametrize("sweep", [1, 2])
def test_scalar_rate_ioup_on_the_register_kernels(oracle, sweep):
    """IOUP(3, -1) (src/priors/ioup.jl; pn_ioup.cuh: phi-functions + 16-point Gauss-Legendre per step inside pn_small.cuh and the sweeps, PN_IOUP=1)
    as a warp of host threads: filtered and smoothed points (column- and row-distributed register sweeps; the warp-per-trajectory sweep serves the
    IWP and the general-prior path only, pnode_api.cu launch_sweep) 1e-10 / 1e-9 against the oracle's Van Loan route, Lotka-Volterra, fixed steps."""
    pd = problems.PROBLEMS["lotka_volterra"]
    q, N, dt, t1 = 3, 9, 0.02, 0.4
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), q)
    u0, p = _ensemble(pd, N, seed=19)
    okw = dict(alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, calibrate=False, adaptive=False, dt=dt, t0=0.0, t1=t1, prior=oracle.PRIOR_IOUP,
               rate_parameter=-1.0, cov_backend=oracle.COV_REF, save_everystep=True)
    of = [oracle.solve(oracle.make_config(2, q, 4, smooth=False, **okw), rhs, u0[i], p[i]) for i in range(N)]
    os_ = [oracle.solve(oracle.make_config(2, q, 4, smooth=True, **okw), rhs, u0[i], p[i]) for i in range(N)]
    ns = of[0]["n"]
    off = np.arange(N + 1) * ns
    mu0 = np.array([o["x_filt_mu"][0] for o in of])
    kf = _run_warp("lotka_volterra", q, 4, 8, mu0, p, off, t1=t1, dt=dt, save=2, smooth=0, defs=("PN_IOUP=1",), ioup_rate=-1.0)
    ks = _run_warp("lotka_volterra", q, 4, 8, mu0, p, off, t1=t1, dt=dt, save=2, smooth=sweep, defs=("PN_IOUP=1",), ioup_rate=-1.0)
    for i in range(N):
        sl = slice(off[i], off[i + 1])
        assert _rel(kf["s_u"][sl], of[i]["pu_mu"]) < 1e-10
        assert _rel(kf["s_cov"][sl][1:], of[i]["pu_cov"][1:]) < 1e-9
        assert kf["scal"][i, 1] == pytest.approx(of[i]["loglik"], rel=1e-9)
        assert _rel(ks["s_u"][sl], os_[i]["pu_mu"]) < 1e-10
        assert _rel(ks["s_cov"][sl][1:], os_[i]["pu_cov"][1:]) < 1e-9


@pytest.mark.parametrize("sweep,data_k", [(1, [10, 20]), (3, [10, 20]), (1, [5, 15]), (3, [5, 15])])
def test_fenrir_backward_pass_as_a_warp(oracle, sweep, data_k):
    """fenrir_data_loglik (src/data_likelihoods/fenrir.jl:68-137): the backward pass with data updates inside the column sweep (pn_smooth_col.cuh built
    with PN_FENRIR, pn_fenrir_update.cuh: sweep 1) and inside the warp-per-trajectory sweep (pn_smooth.cuh: sweep 3) on FitzHugh-Nagumo EK1(3),
    observations of both components with noise variance 0.25: the oracle's value at 1e-10 per trajectory -- with the last observation at the final time
    and with it before the final time (fenrir.jl:87-99: the sweep then starts at the second to last observation)."""
    pd = problems.PROBLEMS["fitzhugh_nagumo"]
    q, N, dt, t1 = 3, 9, 0.05, 1.0
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), q)
    u0, p = _ensemble(pd, N, seed=23)
    data_u = np.array([[1.9, 0.2], [1.7, 0.6]])
    cfg = oracle.make_config(2, q, pd.n_p, alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, calibrate=False, adaptive=False, dt=dt, t0=0.0, t1=t1,
                             save_everystep=True, smooth=True, cov_backend=oracle.COV_REF)
    sols = [oracle.solve(cfg, rhs, u0[i], p[i]) for i in range(N)]
    data_t = sols[0]["t"][data_k]  # grid points of the fixed-step solve (the reference makes the data times tstops)
    want = [oracle.fenrir_data_loglik(cfg, s, data_t, data_u, observation_noise_cov=0.25) for s in sols]
    ns = sols[0]["n"]
    off = np.arange(N + 1) * ns
    mu0 = np.array([s["x_filt_mu"][0] for s in sols])
    H = np.eye(8)[:2]
    k = _run_warp("fitzhugh_nagumo", q, 4, 8, mu0, p, off, t1=t1, dt=dt, save=2, smooth=sweep, defs=("PN_FENRIR=1",),
                  fenrir=(data_t, data_u, H, 0.5 * np.eye(2)))
    assert np.all(np.isfinite(want))
    assert np.allclose(k["fenrir_ll"], want, rtol=1e-10, atol=0)


def test_threadsanitizer_racecheck_general_prior_ioup_and_fenrir(oracle):
    """The remaining lane-exchanging builds under ThreadSanitizer: pn_gen + pn_lti with the Jacobian-updated rate and with a matrix rate followed by the
    dense-transition sweep; pn_small + the column sweep with PN_IOUP; the Fenrir backward pass in the column sweep.  All clean."""
    rate = np.array([[-1.0, 0.3], [-0.2, -0.5]])
    common = dict(N=6, t1=0.2, dt=0.02, adaptive=False, dynamic=False, oracle=oracle)
    runs = [
        _tsan_run("lotka_volterra", 3, 4, 8, save=2, smooth=0, cta=2, okw=dict(prior=oracle.PRIOR_IOUP, update_rate_parameter=True), extra=([1, 1, -1], np.zeros(4)), **common),
        _tsan_run("lotka_volterra", 3, 4, 8, save=2, smooth=3, cta=2, okw=dict(prior=oracle.PRIOR_IOUP, rate_parameter=rate), extra=([1, 0, 1], rate), **common),
        _tsan_run("lotka_volterra", 3, 4, 8, save=2, smooth=1, defs=("PN_IOUP=1",), okw=dict(prior=oracle.PRIOR_IOUP, rate_parameter=-1.0),
                  extra=([2, -1.0], 0.5 * (np.polynomial.legendre.leggauss(16)[0] + 1), 0.5 * np.polynomial.legendre.leggauss(16)[1]), **common),
        _tsan_run("fitzhugh_nagumo", 3, 4, 8, save=2, smooth=1, defs=("PN_FENRIR=1",),
                  extra=([3, 2, 2], [0.1, 0.2], [[1.9, 0.2], [1.7, 0.6]], np.eye(8)[:2], 0.5 * np.eye(2)), **common),
    ]
    for r, nacc in runs:
        assert "ThreadSanitizer" not in r.stderr, r.stderr[:4000]
        assert r.returncode == 0 and f"accepted {nacc} " in r.stdout, (r.returncode, r.stdout, r.stderr[:2000])


def test_observation_noise_and_mass_matrix_builds_as_a_warp(oracle):
    """Two generated-preamble builds of pn_small.cuh as a warp of host threads.  PN_OBSN: `pn_observation_noise = 0.01 I` in the step update
    (S += R, Joseph term K R K' re-factorised; src/filtering/update.jl:125-136, perform_step.jl:185-187) on FitzHugh-Nagumo.  PN_MASS: ROBER as the
    index-1 DAE the reference benchmarks (M = diag(1, 1, 0), H = [-J | M | 0]; measurement_models.jl:11-20), adaptive with DynamicDiffusion.  Tables as
    pnode_api.cu generates them; against the oracle: fixed steps 1e-10 / 1e-9, adaptive accept / reject counts."""
    pd = problems.PROBLEMS["fitzhugh_nagumo"]
    q, N, dt, t1 = 3, 9, 0.05, 1.0
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), q)
    u0, p = _ensemble(pd, N, seed=29)
    Rn = 0.01 * np.eye(2)
    oo = [oracle.solve(oracle.make_config(2, q, pd.n_p, alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, calibrate=False, adaptive=False, dt=dt, t0=0.0, t1=t1,
                                          save_everystep=True, smooth=False, cov_backend=oracle.COV_REF, pn_observation_noise=Rn), rhs, u0[i], p[i])
          for i in range(N)]
    ns = oo[0]["n"]
    off = np.arange(N + 1) * ns
    pre = "#define PN_OBSN 1\n" + _table("pn_obsn_cov", Rn) + _table("pn_obsn_rt", np.linalg.cholesky(Rn).T)
    k = _run_warp("fitzhugh_nagumo", q, 4, 8, np.array([o["x_filt_mu"][0] for o in oo]), p, off, t1=t1, dt=dt, save=2, preamble=pre)
    for i in range(N):
        sl = slice(off[i], off[i + 1])
        assert _rel(k["s_u"][sl], oo[i]["pu_mu"]) < 1e-10
        assert _rel(k["s_cov"][sl][1:], oo[i]["pu_cov"][1:]) < 1e-9
        assert k["scal"][i, 1] == pytest.approx(oo[i]["loglik"], rel=1e-9)

    pd = problems.lookup("rober_dae")
    N = 10
    rhs = oracle.compile_rhs(tracer.trace(pd.f, 3, 3), q)
    rng = np.random.default_rng(31)
    p = np.array(pd.p)[None, :] * (1.0 + 0.05 * rng.standard_normal((N, 3)))
    u0 = np.tile(np.array(pd.u0, dtype=float), (N, 1))
    M = np.array(pd.mass_matrix)
    oo = [oracle.solve(oracle.make_config(3, q, 3, alg=oracle.EK1, diffusion=oracle.DIFF_DYNAMIC, adaptive=True, dt=1e-4, t0=0.0, t1=10.0,
                                          save_everystep=True, smooth=False, cov_backend=oracle.COV_REF, mass_matrix=M), rhs, u0[i], p[i]) for i in range(N)]
    assert all(o["retcode"] == "Success" and o["naccept"] > 5 for o in oo)
    pre = "#define PN_MASS 1\n" + _table("pn_mass", M) + _table("pn_mass_inv", np.linalg.inv(M + 1e-20 * np.eye(3)))
    k = _run_warp("rober_dae", q, 4, 8, np.array([o["x_filt_mu"][0] for o in oo]), p, np.arange(N + 1), t1=10.0, dt=1e-4, adaptive=True, dynamic=True, save=0,
                  preamble=pre)
    for i in range(N):
        assert tuple(k["counts"][i][[0, 1, 3]]) == (oo[i]["naccept"], oo[i]["nreject"], 0)
        assert _rel(k["u"][i], oo[i]["pu_mu"][-1]) < 1e-8


def test_forward_data_updates_as_a_warp(oracle):
    """filtering_data_loglik's forward pass (DataUpdateCallback, src/callbacks/dataupdate.jl:39-84; pn_small.cuh built with the PN_DATA preamble
    pnode_api.cu generates): adaptive EK1(3) with DynamicDiffusion on FitzHugh-Nagumo, observations of both components at t = 0.5 and t = 1 (the first
    one a tstop inside the interval), noise variance 0.25 -- accept / reject counts and the sum of the updates' log-likelihoods against the oracle."""
    pd = problems.PROBLEMS["fitzhugh_nagumo"]
    q, N, t1 = 3, 10, 1.0
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), q)
    u0, p = _ensemble(pd, N, seed=37)
    data_t, data_u = np.array([0.5, 1.0]), np.array([[1.9, 0.2], [1.7, 0.6]])
    oo = [oracle.solve(oracle.make_config(2, q, pd.n_p, alg=oracle.EK1, diffusion=oracle.DIFF_DYNAMIC, adaptive=True, dt=1e-2, t0=0.0, t1=t1, smooth=False,
                                          save_everystep=True, cov_backend=oracle.COV_REF, data=(data_t, data_u), observation_noise_cov=0.25), rhs, u0[i], p[i])
          for i in range(N)]
    pre = ("#define PN_DATA 1\n#define PN_DATA_O 2\n" + _table("pn_data_H", np.eye(8)[:2]) + _table("pn_data_cov", 0.25 * np.eye(2)) +
           _table("pn_data_rt", 0.5 * np.eye(2)))
    k = _run_warp("fitzhugh_nagumo", q, 4, 8, np.array([o["x_filt_mu"][0] for o in oo]), p, np.arange(N + 1), t1=t1, dt=1e-2, adaptive=True, dynamic=True,
                  save=0, preamble=pre, data=(data_t, data_u))
    for i in range(N):
        assert tuple(k["counts"][i][[0, 1, 3]]) == (oo[i]["naccept"], oo[i]["nreject"], 0)
        assert k["data_ll"][i] == pytest.approx(oo[i]["data_loglik"], rel=1e-9)
        assert _rel(k["u"][i], oo[i]["pu_mu"][-1]) < 1e-9


@pytest.mark.parametrize("dynamic,calibrate", [(False, False), (False, True), (True, False)])
def test_dense_output_kernel_as_a_warp(oracle, dynamic, calibrate):
    """sol(t) of a filter-only solve on a saveat grid (src/solution.jl:330-398: predict from the last saved point; pn_dense.cuh, G = 2 lanes per
    (trajectory, grid point) pair): FitzHugh-Nagumo EK1(3), fixed steps, grid points between, on and after saved points -- FixedDiffusion with and
    without calibration (the marginals are scaled by the global MLE), DynamicDiffusion -- against the oracle's `interpolate`, 1e-10 / 1e-9."""
    pd = problems.PROBLEMS["fitzhugh_nagumo"]
    q, N, dt, t1 = 3, 7, 0.05, 1.0
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), q)
    u0, p = _ensemble(pd, N, seed=41)
    cfg = oracle.make_config(2, q, pd.n_p, alg=oracle.EK1, diffusion=oracle.DIFF_DYNAMIC if dynamic else oracle.DIFF_FIXED, calibrate=calibrate, adaptive=False,
                             dt=dt, t0=0.0, t1=t1, save_everystep=True, smooth=False, cov_backend=oracle.COV_REF)
    sols = [oracle.solve(cfg, rhs, u0[i], p[i]) for i in range(N)]
    ns = sols[0]["n"]
    off = np.arange(N + 1) * ns
    grid = np.concatenate([[0.0, 0.013, 0.31, 0.5 * (sols[0]["t"][7] + sols[0]["t"][8])], sols[0]["t"][[3, 11]], [0.987, 1.0]])
    k = _run_warp("fitzhugh_nagumo", q, 4, 8, np.array([s["x_filt_mu"][0] for s in sols]), p, off, t1=t1, dt=dt, dynamic=dynamic, calibrate=calibrate, save=2,
                  saveat=grid)
    for i in range(N):
        for j, tv in enumerate(grid):
            um, uc = oracle.interpolate(cfg, sols[i], float(tv))
            assert _rel(k["dense_u"][i, j], um) < 1e-10
            if tv > 0:
                assert _rel(k["dense_cov"][i, j], uc) < 1e-9
            else:
                assert np.all(k["dense_cov"][i, j] == 0.0)


def test_threadsanitizer_racecheck_generated_preamble_builds(oracle):
    """The builds of pn_small.cuh with a generated preamble under ThreadSanitizer: pn_observation_noise (fixed steps, saving pass), the mass-matrix
    build on ROBER as a DAE (adaptive) and forward data updates (adaptive, tstops).  All clean."""
    Rn = 0.01 * np.eye(2)
    M = np.diag([1.0, 1.0, 0.0])
    pdr = problems.lookup("rober_dae")
    rng = np.random.default_rng(43)
    pr = np.array(pdr.p)[None, :] * (1.0 + 0.05 * rng.standard_normal((6, 3)))
    data_t, data_u = np.array([0.5, 1.0]), np.array([[1.9, 0.2], [1.7, 0.6]])
    runs = [
        _tsan_run("fitzhugh_nagumo", 3, 4, 8, N=9, t1=0.5, dt=0.05, adaptive=False, dynamic=False, save=2, smooth=0, oracle=oracle,
                  okw=dict(pn_observation_noise=Rn), preamble="#define PN_OBSN 1\n" + _table("pn_obsn_cov", Rn) + _table("pn_obsn_rt", np.linalg.cholesky(Rn).T)),
        _tsan_run("rober_dae", 3, 4, 8, N=6, t1=1.0, dt=1e-4, adaptive=True, dynamic=True, save=0, smooth=0, oracle=oracle, okw=dict(mass_matrix=M),
                  p_override=(np.tile(np.array(pdr.u0, dtype=float), (6, 1)), pr),
                  preamble="#define PN_MASS 1\n" + _table("pn_mass", M) + _table("pn_mass_inv", np.linalg.inv(M + 1e-20 * np.eye(3)))),
        _tsan_run("fitzhugh_nagumo", 3, 4, 8, N=9, t1=1.0, dt=1e-2, adaptive=True, dynamic=True, save=0, smooth=0, oracle=oracle,
                  okw=dict(data=(data_t, data_u), observation_noise_cov=0.25), extra=([4, 2, 2], data_t, data_u),
                  preamble="#define PN_DATA 1\n#define PN_DATA_O 2\n" + _table("pn_data_H", np.eye(8)[:2]) + _table("pn_data_cov", 0.25 * np.eye(2)) +
                  _table("pn_data_rt", 0.5 * np.eye(2))),
    ]
    for r, nacc in runs:
        assert "ThreadSanitizer" not in r.stderr, r.stderr[:4000]
        assert r.returncode == 0 and f"accepted {nacc} " in r.stdout, (r.returncode, r.stdout, r.stderr[:2000])


def test_second_order_problem_on_the_group_kernel_and_column_sweep(oracle):
    """SecondOrderODEProblem (test/secondorderode.jl:15-19, the two-body problem; d = 2 positions, D = 8 at order 3, sol.u = (du, u)) on pn_small.cuh and
    pn_smooth_col.cuh built with PN_SECOND, traced at run time: filtered and smoothed points 1e-10 / 1e-9 against the oracle's second-order route."""
    pd = problems.lookup("twobody")
    q, N, dt, t1 = 3, 9, 0.005, 0.1
    rhs = oracle.compile_rhs(tracer.trace(pd.f, 4, 1), q)
    rng = np.random.default_rng(47)
    u0 = np.array(pd.u0)[None, :] * (1.0 + 0.02 * rng.standard_normal((N, 4)))
    p = np.array(pd.p)[None, :] * (1.0 + 0.05 * rng.standard_normal((N, 1)))
    kw = dict(alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, calibrate=False, adaptive=False, dt=dt, t0=0.0, t1=t1, save_everystep=True, second_order=True)
    of = [oracle.solve(oracle.make_config(2, q, 1, smooth=False, cov_backend=oracle.COV_REF, **kw), rhs, u0[i], p[i]) for i in range(N)]
    os_ = [oracle.solve(oracle.make_config(2, q, 1, smooth=True, cov_backend=oracle.COV_REF, **kw), rhs, u0[i], p[i]) for i in range(N)]
    os2 = [oracle.solve(oracle.make_config(2, q, 1, smooth=True, cov_backend=oracle.COV_QR, **kw), rhs, u0[i], p[i]) for i in range(N)]
    ns = of[0]["n"]
    off = np.arange(N + 1) * ns
    mu0 = np.array([o["x_filt_mu"][0] for o in of])
    assert mu0.shape == (N, 8)
    kf = _run_warp("twobody", q, 4, 8, mu0, p, off, t1=t1, dt=dt, save=2, smooth=0, defs=("PN_SECOND=1",))
    ks = _run_warp("twobody", q, 4, 8, mu0, p, off, t1=t1, dt=dt, save=2, smooth=1, defs=("PN_SECOND=1",))
    for i in range(N):
        sl = slice(off[i], off[i + 1])
        assert _rel(kf["s_u"][sl], of[i]["pu_mu"]) < 1e-10
        assert _rel(kf["s_cov"][sl][1:], of[i]["pu_cov"][1:]) < 1e-9
        assert kf["scal"][i, 1] == pytest.approx(of[i]["loglik"], rel=1e-9)
        assert _rel(ks["s_u"][sl], os_[i]["pu_mu"]) < 1e-10
        assert _rel(ks["s_cov"][sl][1:], os_[i]["pu_cov"][1:]) < max(1e-9, 20 * _rel(os2[i]["pu_cov"][1:], os_[i]["pu_cov"][1:]))


@pytest.mark.parametrize("sweep", [1, 2, 3])
def test_calibrated_fixed_diffusion_through_the_sweeps(oracle, sweep):
    """FixedDiffusion(calibrate = true) (calibrate_solution!, src/integrator_utils.jl:46-57; global MLE, src/diffusions/calibration.jl:49-66): the step kernel
    reports sigma_hat^2 and the sweeps scale the smoothed covariances by it -- FitzHugh-Nagumo EK1(3), fixed steps, all three sweeps, against the oracle:
    global diffusion 1e-10, smoothed means 1e-10, calibrated covariances 1e-9 / floor."""
    pd = problems.PROBLEMS["fitzhugh_nagumo"]
    q, N, dt, t1 = 3, 9, 0.05, 1.0
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), q)
    u0, p = _ensemble(pd, N, seed=53)
    kw = dict(alg=oracle.EK1, diffusion=oracle.DIFF_FIXED, calibrate=True, adaptive=False, dt=dt, t0=0.0, t1=t1, save_everystep=True, smooth=True)
    os_ = [oracle.solve(oracle.make_config(2, q, pd.n_p, cov_backend=oracle.COV_REF, **kw), rhs, u0[i], p[i]) for i in range(N)]
    os2 = [oracle.solve(oracle.make_config(2, q, pd.n_p, cov_backend=oracle.COV_QR, **kw), rhs, u0[i], p[i]) for i in range(N)]
    ns = os_[0]["n"]
    off = np.arange(N + 1) * ns
    ks = _run_warp("fitzhugh_nagumo", q, 4, 8, np.array([o["x_filt_mu"][0] for o in os_]), p, off, t1=t1, dt=dt, calibrate=True, save=2, smooth=sweep)
    for i in range(N):
        sl = slice(off[i], off[i + 1])
        assert ks["scal"][i, 2] == pytest.approx(os_[i]["global_diffusion"], rel=1e-10)
        assert _rel(ks["s_u"][sl], os_[i]["pu_mu"]) < 1e-10
        assert _rel(ks["s_cov"][sl][1:], os_[i]["pu_cov"][1:]) < max(1e-9, 20 * _rel(os2[i]["pu_cov"][1:], os_[i]["pu_cov"][1:]))


@pytest.mark.parametrize("problem,q,G,GS,t1,dt", [("fitzhugh_nagumo", 3, 4, 8, 1.0, 0.05), ("vanderpol", 5, 8, 8, 0.01, 2.5e-4)])
@pytest.mark.parametrize("sweep", [1, 2])
def test_dynamic_diffusion_through_the_register_sweeps(oracle, problem, q, G, GS, t1, dt, sweep):
    """DynamicDiffusion (the reference's default; src/diffusions/calibration.jl:123-133): the local estimate of every step is saved and scales that step's
    process noise in the backward sweep.  cfg1 / cfg3 shapes, fixed steps, both register sweeps against the oracle; tolerances are 20 x the oracle's own
    Cholesky-vs-QR floor measured here (with a per-step diffusion the covariances of the two routes sit further apart than under FixedDiffusion), at
    least 1e-10 (means) / 1e-9 (covariances)."""
    pd = problems.PROBLEMS[problem]
    N = 6
    rhs = oracle.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), q)
    u0, p = _ensemble(pd, N, seed=59)
    kw = dict(alg=oracle.EK1, diffusion=oracle.DIFF_DYNAMIC, adaptive=False, dt=dt, t0=0.0, t1=t1, save_everystep=True, smooth=True)
    os_ = [oracle.solve(oracle.make_config(pd.d, q, pd.n_p, cov_backend=oracle.COV_REF, **kw), rhs, u0[i], p[i]) for i in range(N)]
    os2 = [oracle.solve(oracle.make_config(pd.d, q, pd.n_p, cov_backend=oracle.COV_QR, **kw), rhs, u0[i], p[i]) for i in range(N)]
    ns = os_[0]["n"]
    off = np.arange(N + 1) * ns
    ks = _run_warp(problem, q, G, GS, np.array([o["x_filt_mu"][0] for o in os_]), p, off, t1=t1, dt=dt, dynamic=True, save=2, smooth=sweep)
    for i in range(N):
        sl = slice(off[i], off[i + 1])
        assert _rel(ks["s_diff"][sl][:-1], os_[i]["diffusions"][:ns - 1]) < 1e-8
        assert _rel(ks["s_u"][sl], os_[i]["pu_mu"]) < max(1e-10, 20 * _rel(os2[i]["pu_mu"], os_[i]["pu_mu"]))
        assert _rel(ks["s_cov"][sl][1:], os_[i]["pu_cov"][1:]) < max(1e-9, 20 * _rel(os2[i]["pu_cov"][1:], os_[i]["pu_cov"][1:]))
