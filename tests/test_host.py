"""CPU-side checks of the product: the C-ABI library loads and exports every symbol include/pnode.h declares,
argument errors mirror the reference's exceptions, NVRTC compiles traced right-hand sides for sm_100a without a
device, and the device Taylor arithmetic (compiled for the host here) reproduces symbolic derivatives."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest
import sympy as sp

import probnumdiffeq_jl_b200  # noqa: F401
from probnumdiffeq_jl_b200 import lib, problems, tracer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(pnlib):
    hdr = open(os.path.join(ROOT, "include", "pnode.h")).read()
    declared = set(re.findall(r"^(?:int|void|double|int64_t|const char\*|const void\*)\s+(pnode_[a-z0-9_]+)\s*\(", hdr, re.M))
    assert declared, "no declarations found"
    L = pnlib.load()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in pnode.h but not exported"
    assert declared == set(pnlib.EXPORTS)
    assert L.pnode_abi_version() == 1
    # struct layout agreement between pnode.h and the ctypes mirror
    src = '#include <stdio.h>\n#include "pnode.h"\nint main(){printf("%zu %zu", sizeof(pnode_config), sizeof(pnode_result_info));return 0;}'
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        a, b = map(int, subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split())
    assert a == C.sizeof(pnlib.Config) and b == C.sizeof(pnlib.ResultInfo)


def test_no_device_fails_loudly(pnlib):
    if pnlib.device_count() > 0:
        pytest.skip("a CUDA device is present")
    cfg = pnlib.make_config(d=3, q=3, n_p=3, rhs_name="builtin:rober", t0=0.0, t1=1.0)
    with pytest.raises(pnlib.PnodeError) as e:
        pnlib.Solver(cfg)
    assert e.value.kind == "NoDevice"


def _src(name="rober"):
    pd = problems.PROBLEMS[name]
    return pd, tracer.cuda_source(tracer.trace(pd.f, pd.d, pd.n_p), "UserProb")


def test_argument_errors_mirror_reference(pnlib):
    pd, src = _src()
    base = dict(d=3, q=3, n_p=3, rhs_name="UserProb", rhs_src=src, t0=0.0, t1=1.0)

    def err(**kw):
        with pytest.raises(pnlib.PnodeError) as e:
            pnlib.compile_check(pnlib.make_config(**{**base, **kw}))
        return e.value

    # test/errors_thrown.jl:6-16: adaptive=false without dt -> ArgumentError
    assert err(adaptive=False, dt=0.0).kind == "ArgumentError"
    # test/errors_thrown.jl:23-26 / src/checks.jl:18-20
    assert err(smooth=True, save_everystep=False).kind == "ArgumentError"
    assert err(q=0).kind == "ArgumentError"
    assert err(t1=-1.0).kind == "ArgumentError"
    assert err(lanes_per_traj=3).kind == "ArgumentError"
    assert err(rhs_src=None).kind == "ArgumentError"
    # src/algorithms.jl:101-110: Kronecker covariance only with the EK0
    assert err(cov=pnlib.COV_KRONECKER).kind == "ArgumentError"
    # ekargcheck (src/algorithms.jl:93-107; test/errors_thrown.jl:52-56): EK1 with multivariate diffusion models
    assert err(diffusion=pnlib.DIFF_DYNAMIC_MV).kind == "ArgumentError"
    assert err(diffusion=pnlib.DIFF_FIXED_MV, calibrate=True).kind == "ArgumentError"
    assert err(alg=7).kind == "ArgumentError"
    # the BlocksOfDiagonals kernels (EK0 + multivariate diffusion, DiagonalEK1) compile for sm_100a
    for kw in (dict(alg=pnlib.EK0, diffusion=pnlib.DIFF_DYNAMIC_MV), dict(alg=pnlib.EK0, diffusion=pnlib.DIFF_FIXED_MV),
               dict(alg=pnlib.DIAGONAL_EK1), dict(alg=pnlib.DIAGONAL_EK1, diffusion=pnlib.DIFF_FIXED, smooth=True, save_everystep=True)):
        assert pnlib.compile_check(pnlib.make_config(**{**base, **kw})) > 10_000
    bad = pnlib.make_config(**base)
    bad.struct_size = 8
    with pytest.raises(pnlib.PnodeError):
        pnlib.compile_check(bad)


def test_nvrtc_compiles_traced_rhs_for_sm100a(pnlib):
    pd, src = _src("vanderpol")
    n = pnlib.compile_check(pnlib.make_config(d=2, q=5, n_p=1, rhs_name="UserProb", rhs_src=src, t0=0.0, t1=2.0))
    assert n > 10_000
    with pytest.raises(pnlib.PnodeError) as e:
        pnlib.compile_check(pnlib.make_config(d=2, q=5, n_p=1, rhs_name="UserProb", rhs_src=src + "\n#error boom\n", t0=0.0, t1=2.0))
    assert e.value.kind == "CompileError" and "boom" in str(e.value)


def test_nvrtc_compiles_smoother_and_dense_output_kernels(pnlib, monkeypatch):
    """The RTS sweeps (column- and row-distributed), the dense-output kernel and the IOUP variants compile for sm_100a
    without a device (pnode_compile_check compiles every kernel the configuration needs)."""
    pd, src = _src("vanderpol")
    base = dict(d=2, q=5, n_p=1, rhs_name="UserProb", rhs_src=src, t0=0.0, t1=2.0, smooth=True, save_everystep=True)
    step_only = pnlib.compile_check(pnlib.make_config(**{**base, "smooth": False}))
    with_sweep = pnlib.compile_check(pnlib.make_config(**base))
    assert with_sweep > step_only
    monkeypatch.setenv("PNODE_SMOOTH_KIND", "0")
    assert pnlib.compile_check(pnlib.make_config(**base)) > step_only
    monkeypatch.delenv("PNODE_SMOOTH_KIND")
    pd, src = _src("lotka_volterra")
    kw = dict(d=2, q=3, n_p=4, rhs_name="UserProb", rhs_src=src, t0=0.0, t1=1.0, smooth=True, save_everystep=True,
              saveat=[0.0, 0.5, 1.0])
    assert pnlib.compile_check(pnlib.make_config(**kw)) > 10_000
    assert pnlib.compile_check(pnlib.make_config(prior=pnlib.PRIOR_IOUP, rate_parameter=-1.0, **kw)) > 10_000
    with pytest.raises(pnlib.PnodeError) as e:   # src/solution.jl:340-342
        pnlib.compile_check(pnlib.make_config(**{**kw, "saveat": [-0.1, 0.5]}))
    assert "Invalid t<t0" in str(e.value)


def test_hand_written_problem_struct_compiles_for_every_kernel(pnlib):
    """The ABI takes CUDA C++ source: a minimal struct (static d, n_p, f<T>, jac -- what the Julia host's Symbolics tracer
    emits, without the optional f_part / jac_part slices) must compile for the group-per-trajectory, the
    thread-per-trajectory and the CTA-per-trajectory kernels."""
    src = """
struct Mini {
  static constexpr int d = 2; static constexpr int n_p = 1;
  template <class T> __device__ __forceinline__ static void f(T* du, const T* u, const double* p, const T& t) {
    du[0] = u[1]; du[1] = -p[0] * u[0];
  }
  __device__ __forceinline__ static void jac(double* J, const double* u, const double* p, double t) {
    J[0] = 0.0; J[1] = -p[0]; J[2] = 1.0; J[3] = 0.0;
  }
};"""
    base = dict(d=2, q=3, n_p=1, rhs_name="Mini", rhs_src=src, t0=0.0, t1=1.0)
    for kw in (dict(), dict(alg=pnlib.EK0), dict(alg=pnlib.DIAGONAL_EK1), dict(lanes_per_traj=256)):
        assert pnlib.compile_check(pnlib.make_config(**{**base, **kw})) > 10_000


def test_tracer_matches_python_rhs():
    for name, pd in problems.PROBLEMS.items():
        tr = tracer.trace(pd.f, pd.d, pd.n_p)
        subs = {s: v for s, v in zip(tr.u, pd.u0)}
        subs.update({s: v for s, v in zip(tr.p, pd.p)})
        du = [0.0] * pd.d
        pd.f(du, list(map(float, pd.u0)), list(map(float, pd.p)), 0.0)
        got = [float(e.subs(subs)) for e in tr.f]
        assert np.allclose(got, du, rtol=1e-13, atol=1e-300), name
        assert "struct X" in tracer.cuda_source(tr, "X")


_TAYLOR_MAIN = r"""
#include <cmath>
#include <cstdio>
#define PN_HD inline
#define __device__
#define __forceinline__ inline
#include "pn_taylor.cuh"
%(problem)s
int main() {
  double u0[] = {%(u0)s};
  double p[] = {%(p)s};
  double out[(%(q)d + 1) * %(d)d];
  pn_taylor_init<UserProb, %(q)d>(out, u0, p, %(t0)s);
  for (int i = 0; i < (%(q)d + 1) * %(d)d; i++) printf("%%.17g\n", out[i]);
  return 0;
}
"""


def _device_taylor_on_host(f, d, n_p, u0, p, q, t0=0.0):
    tr = tracer.trace(f, d, n_p)
    src = _TAYLOR_MAIN % dict(problem=tracer.cuda_source(tr, "UserProb"), u0=",".join(map(repr, map(float, u0))),
                              p=",".join(map(repr, map(float, p))) or "0.0", q=q, d=d, t0=repr(float(t0)))
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "t.cpp")
        open(c, "w").write(src)
        exe = os.path.join(td, "t")
        subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "probnumdiffeq.jl_b200", "csrc", "kernels"), c, "-o", exe],
                       check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    return tr, np.array(list(map(float, out))).reshape(q + 1, d)


def test_taylor_mode_init_matches_symbolic_derivatives(oracle):
    """Product: truncated Taylor arithmetic (pn_taylor.cuh).  Oracle: symbolic Lie derivatives.  Two different
    methods must give the same u^(k)(t0) (reference pin: test/state_init.jl:9-47)."""
    for name, q in [("fitzhugh_nagumo", 5), ("vanderpol", 5), ("rober", 3), ("lotka_volterra", 4)]:
        pd = problems.PROBLEMS[name]
        tr, got = _device_taylor_on_host(pd.f, pd.d, pd.n_p, pd.u0, pd.p, q)
        ys = oracle.symbolic_derivatives(tr, q)
        subs = {s: v for s, v in zip(tr.u, pd.u0)}
        subs.update({s: v for s, v in zip(tr.p, pd.p)})
        want = np.array([[float(ys[k][i].subs(subs)) for i in range(pd.d)] for k in range(q + 1)])
        assert np.allclose(got, want, rtol=1e-12, atol=1e-12 * np.abs(want).max()), name


def test_taylor_elementary_functions():
    """exp / log / sin / cos / sqrt / tanh / pow recurrences against SymPy."""
    def f(du, u, p, t):
        du[0] = tracer.exp(-u[0]) * tracer.sin(u[1]) + tracer.sqrt(1 + u[0] ** 2) / (2 + tracer.cos(t))
        du[1] = tracer.log(2 + u[1] ** 2) * tracer.tanh(u[0]) - (1 + u[0] ** 2) ** sp.Rational(-3, 2) * p[0]

    u0, p, q = [0.3, -0.8], [1.7], 4
    tr, got = _device_taylor_on_host(f, 2, 1, u0, p, q, t0=0.25)
    from oracle import oracle as O

    ys = O.symbolic_derivatives(tr, q)
    subs = {tr.u[0]: u0[0], tr.u[1]: u0[1], tr.p[0]: p[0], tr.t: 0.25}
    want = np.array([[float(ys[k][i].subs(subs)) for i in range(2)] for k in range(q + 1)])
    assert np.allclose(got, want, rtol=1e-11, atol=1e-12)


def test_mass_matrix_kernels_compile_and_argument_checks(pnlib):
    """Mass-matrix problems (src/measurement_models.jl:11-20, src/derivative_utils.jl:43-50): the EK1 step kernel with M
    baked in compiles for sm_100a; EK0 on the Kronecker layout raises the reference's ArgumentError (src/caches.jl:111-118);
    an identity M is the plain path; a matrix whose initial conditioning is singular is rejected."""
    pd, src = _src("rober")
    base = dict(d=3, q=3, n_p=3, rhs_name="UserProb", rhs_src=src, t0=0.0, t1=1e-2)
    plain = pnlib.compile_check(pnlib.make_config(**base))
    assert pnlib.compile_check(pnlib.make_config(mass_matrix=np.eye(3), **base)) == plain
    assert pnlib.compile_check(pnlib.make_config(mass_matrix=np.diag([1.0, 1.0, 0.0]), **base)) > 10_000
    assert pnlib.compile_check(pnlib.make_config(mass_matrix=[[1, 0.2, 0], [0, 1, 0], [0.1, 0, 0.5]], smooth=True,
                                                 save_everystep=True, saveat=[0.0, 5e-3], **base)) > 10_000
    # EK0 / Kronecker: only a UniformScaling (M == lambda I) passes; DiagonalEK1 / blocks: diagonal M
    assert pnlib.compile_check(pnlib.make_config(mass_matrix=-100 * np.eye(3), alg=pnlib.EK0, **base)) > 10_000
    assert pnlib.compile_check(pnlib.make_config(mass_matrix=np.diag([1.0, 1.0, 0.0]), alg=pnlib.DIAGONAL_EK1, **base)) > 10_000
    for M in (np.diag([1.0, 2.0, 3.0]), np.diag([1.0, 1.0, 0.0]), [[1, 0.2, 0], [0, 1, 0], [0, 0, 1]]):
        with pytest.raises(pnlib.PnodeError) as e:
            pnlib.compile_check(pnlib.make_config(mass_matrix=M, alg=pnlib.EK0, **base))
        assert e.value.kind == "ArgumentError" and "incompatible with the provided mass matrix" in str(e.value)
    with pytest.raises(pnlib.PnodeError) as e:
        pnlib.compile_check(pnlib.make_config(mass_matrix=[[1, 0.2, 0], [0, 1, 0], [0, 0, 1]], alg=pnlib.DIAGONAL_EK1, **base))
    assert e.value.kind == "Unsupported"
    with pytest.raises(pnlib.PnodeError) as e:   # EK0 + multivariate diffusion (blocks) with a singular M
        pnlib.compile_check(pnlib.make_config(mass_matrix=np.diag([1.0, 1.0, 0.0]), alg=pnlib.EK0, diffusion=pnlib.DIFF_DYNAMIC_MV, **base))
    assert e.value.kind == "ArgumentError"
    with pytest.raises(pnlib.PnodeError) as e:
        pnlib.compile_check(pnlib.make_config(mass_matrix=[[1, 1, 0], [1, 1, 0], [0, 0, 1]], **base))
    assert e.value.kind == "ArgumentError"


def _twobody_first_order(du, u, p, t):
    r3 = (u[2] * u[2] + u[3] * u[3]) ** 1.5
    du[0] = -u[2] / r3
    du[1] = -u[3] / r3
    du[2] = u[0]
    du[3] = u[1]


def test_second_order_kernels_compile_and_argument_checks(pnlib):
    """SecondOrderODEProblem path (src/measurement_models.jl:29-49): the step kernels (EK1 dense, EK0 Kronecker), the
    smoother sweep and the dense-output kernel compile with PN_SECOND for sm_100a; unsupported combinations are rejected."""
    src = tracer.cuda_source(tracer.trace(_twobody_first_order, 4, 0), "UserProb")
    base = dict(d=2, q=3, n_p=0, rhs_name="UserProb", rhs_src=src, t0=0.0, t1=0.1, second_order=True)
    for alg in (pnlib.EK1, pnlib.EK0):
        a = pnlib.compile_check(pnlib.make_config(alg=alg, **base))
        b = pnlib.compile_check(pnlib.make_config(alg=alg, smooth=True, save_everystep=True, saveat=[0.0, 0.03, 0.1], **base))
        assert b > a > 10_000
    for kw in (dict(alg=pnlib.DIAGONAL_EK1), dict(alg=pnlib.EK0, diffusion=pnlib.DIFF_DYNAMIC_MV)):   # blocks layout
        assert pnlib.compile_check(pnlib.make_config(**{**base, **kw})) > 10_000
    for kw, kind in ((dict(q=1), "ArgumentError"),
                     (dict(mass_matrix=2 * np.eye(2)), "ArgumentError"),
                     (dict(prior=pnlib.PRIOR_IOUP, rate_parameter=-1.0), "Unsupported")):
        with pytest.raises(pnlib.PnodeError) as e:
            pnlib.compile_check(pnlib.make_config(**{**base, **kw}))
        assert e.value.kind == kind, kw


def test_fenrir_configuration_checks(pnlib):
    """fenrir_data_loglik (src/data_likelihoods/fenrir.jl:30-66): needs smoothing (ArgumentError, :42-44), a consistent data
    set and a positive (definite) observation noise; unsupported combinations are named."""
    pd, src = _src("lotka_volterra")
    base = dict(d=2, q=3, n_p=4, rhs_name="UserProb", rhs_src=src, t0=0.0, t1=2.0, smooth=True, save_everystep=True,
                data_t=[1.0, 2.0], data_u=[[1.0, 1.0], [2.0, 2.0]], observation_noise_cov=0.25)
    assert pnlib.compile_check(pnlib.make_config(**base)) > 10_000
    assert pnlib.compile_check(pnlib.make_config(**{**base, "observation_noise_cov": [[0.25, 0.01], [0.01, 0.5]]})) > 10_000

    def err(**kw):
        with pytest.raises(pnlib.PnodeError) as e:
            pnlib.compile_check(pnlib.make_config(**{**base, **kw}))
        return e.value

    e = err(smooth=False)
    assert e.kind == "ArgumentError" and "fenrir only works with smoothing" in str(e)
    assert err(data_t=[2.0, 1.0]).kind == "ArgumentError"
    assert err(data_t=[1.0, 3.0]).kind == "ArgumentError"
    assert err(observation_noise_cov=0.0).kind == "ArgumentError"
    assert err(data_u=[[1.0], [2.0]]).kind == "ArgumentError"                      # partial data without an observation matrix
    assert pnlib.compile_check(pnlib.make_config(**{**base, "data_u": [[1.0], [2.0]], "observation_matrix": [[1.0, 0.0]]})) > 10_000
    assert err(saveat=[0.5]).kind == "Unsupported"
    assert err(prior=pnlib.PRIOR_IOUP, rate_parameter=-1.0).kind == "Unsupported"
    assert err(alg=pnlib.EK0, diffusion=pnlib.DIFF_DYNAMIC_MV).kind == "Unsupported"
    # structured layouts: the reference's data update runs on the factors -- full observations only
    # (src/callbacks/dataupdate.jl:69-71) and a noise covariance of the layout's own shape (covariance_structure.jl:46-58)
    for alg in (pnlib.EK0, pnlib.DIAGONAL_EK1):
        assert pnlib.compile_check(pnlib.make_config(**{**base, "alg": alg})) > 10_000
        e = err(alg=alg, data_u=[[1.0], [2.0]], observation_matrix=[[1.0, 0.0]])
        assert e.kind == "ArgumentError" and "Partial observations only work with the EK1 right now" in str(e)
        assert err(alg=alg, observation_noise_cov=[[0.25, 0.01], [0.01, 0.5]]).kind == "ArgumentError"
    assert err(alg=pnlib.EK0, observation_noise_cov=[[0.25, 0.0], [0.0, 0.5]]).kind == "ArgumentError"       # not isotropic
    assert pnlib.compile_check(pnlib.make_config(**{**base, "alg": pnlib.EK0, "observation_noise_cov": [[0.25, 0.0], [0.0, 0.25]]})) > 10_000
    assert pnlib.compile_check(pnlib.make_config(**{**base, "alg": pnlib.DIAGONAL_EK1, "observation_noise_cov": [[0.25, 0.0], [0.0, 0.5]]})) > 10_000


def test_fenrir_update_header_against_reference_formulas(tmp_path):
    """csrc/kernels/pn_fenrir_update.cuh is plain C++: compile it with g++ and compare the single-thread data update (one
    orthogonal transformation of the pre-array) with the reference's formulas -- make_obscov_sqrt
    (src/callbacks/dataupdate.jl:86-87), update!(...; R) (src/filtering/update.jl:65-139) -- on random inputs: mean,
    covariance R'R and log-likelihood, full / partial observations, triangular and dense prior factors, Sigma == 0."""
    hdr = os.path.join(os.path.dirname(lib.__file__), "csrc", "kernels")
    src = tmp_path / "fen.cpp"
    src.write_text('#include <math.h>\n#define PN_FEN_FN\n#define PN_FEN_INF INFINITY\n#include "pn_fenrir_update.cuh"\n'
                   'extern "C" double upd_8_2(const double* H, const double* Rn, const double* u, int o, double* Sm, double* Ms) {\n'
                   '  return pn_fenrir_update_serial<8, 9, 2>(H, Rn, u, o, Sm, Ms);\n}\n'
                   'extern "C" double upd_12_4(const double* H, const double* Rn, const double* u, int o, double* Sm, double* Ms) {\n'
                   '  return pn_fenrir_update_serial<12, 13, 4>(H, Rn, u, o, Sm, Ms);\n}\n')
    so = tmp_path / "fen.so"
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-I", hdr, "-o", str(so), str(src)], check=True)
    L = C.CDLL(str(so))
    dp = C.POINTER(C.c_double)
    rng = np.random.default_rng(0)
    for fn, D, LD, o, dense in ((L.upd_8_2, 8, 9, 2, False), (L.upd_8_2, 8, 9, 1, True), (L.upd_12_4, 12, 13, 4, False),
                                (L.upd_12_4, 12, 13, 3, True)):
        fn.restype = C.c_double
        fn.argtypes = [dp, dp, dp, C.c_int, dp, dp]
        R = rng.standard_normal((D, D)) * 1e-2
        if not dense:
            R = np.triu(R)
        m = rng.standard_normal(D)
        H = np.ascontiguousarray(rng.standard_normal((o, D)))
        Rn = np.ascontiguousarray(np.triu(rng.standard_normal((o, o))) + 2 * np.eye(o))
        u = rng.standard_normal(o)
        # reference formulas
        z = H @ m - u
        Cm = R @ H.T
        Sr = np.linalg.qr(np.vstack([Cm, Rn]), mode="r")
        S = Sr.T @ Sr
        K = np.linalg.solve(S, (R.T @ Cm).T).T
        ll = -0.5 * z @ np.linalg.solve(S, z) - 0.5 * o * np.log(2 * np.pi) - np.sum(np.log(np.abs(np.diag(Sr))))
        m_ref = m - K @ z
        R2 = R @ (np.eye(D) - K @ H).T
        R3 = np.linalg.qr(np.vstack([R2, (K @ Rn.T).T]), mode="r")
        Sm = np.zeros((D, LD))
        Sm[:, :D] = R
        Ms = m.copy()
        got = fn(H.ctypes.data_as(dp), Rn.ctypes.data_as(dp), u.ctypes.data_as(dp), o, Sm.ctypes.data_as(dp), Ms.ctypes.data_as(dp))
        assert got == pytest.approx(ll, rel=1e-12)
        assert np.allclose(Ms, m_ref, rtol=1e-12, atol=1e-15)
        Sg, Sw = Sm[:, :D].T @ Sm[:, :D], R3.T @ R3
        assert np.linalg.norm(Sg - Sw) / np.linalg.norm(Sw) < 1e-12
        assert not Sm[:, D:].any()
    # Sigma == 0: +Inf if the residual is zero, -Inf otherwise, state untouched (update.jl:82-89)
    Sm, Ms = np.zeros((8, 9)), np.arange(8.0)
    H = np.ascontiguousarray(np.eye(8)[:2])
    Rn = np.ascontiguousarray(np.eye(2))
    for u, want in ((np.array([0.0, 1.0]), np.inf), (np.array([0.5, 1.0]), -np.inf)):
        got = L.upd_8_2(H.ctypes.data_as(dp), Rn.ctypes.data_as(dp), u.ctypes.data_as(dp), 2, Sm.ctypes.data_as(dp), Ms.ctypes.data_as(dp))
        assert got == want and np.array_equal(Ms, np.arange(8.0))


def test_python_mirror_argument_errors_of_the_new_entry_points():
    """api.SecondOrderODEProblem / api.fenrir_data_loglik raise before any device work, like the reference's constructors."""
    from probnumdiffeq_jl_b200 import api

    with pytest.raises(ValueError):   # DimensionMismatch
        api.SecondOrderODEProblem(lambda ddu, du, u, p, t: None, [0.0], [1.0, 2.0], (0.0, 1.0))
    prob = api.ODEProblem.from_library("lotka_volterra")
    with pytest.raises(ValueError) as e:   # src/data_likelihoods/fenrir.jl:42-44
        api.fenrir_data_loglik(prob, api.EK1(order=3, smooth=False), data=([1.0], [[1.0, 1.0]]), observation_noise_cov=0.1)
    assert "fenrir only works with smoothing" in str(e.value)
    two = api.SecondOrderODEProblem(lambda ddu, du, u, p, t: [ddu.__setitem__(0, -u[0])], [0.0], [1.0], (0.0, 1.0))
    assert two.second_order and two.u0 == [0.0, 1.0]
    dx = [0, 0]
    two.f(dx, [0.5, 2.0], [], 0.0)     # first-order form F([du; u]) = [f1(du, u); du]
    assert dx == [-2.0, 0.5]
    dae = api.ODEProblem.from_library("rober_dae")
    with pytest.raises(ValueError) as e:   # src/caches.jl:111-118
        api._config_for(dae, api.EK0(order=3), api.EnsembleB200(), adaptive=True, dt=None, abstol=1e-6, reltol=1e-3, save_everystep=True,
                        maxiters=10 ** 6, dtmin=0.0, dtmax=None, tstops=None, save_state=False)
    assert "incompatible with the provided mass matrix" in str(e.value)


def test_pn_observation_noise_argument_errors_and_compilation(pnlib):
    """ekargcheck (src/algorithms.jl:78-130) for `pn_observation_noise`, and the three noise-aware step kernels compile for
    sm_100a (test/observation_noise.jl's compatibility table: EK0 isotropic only, DiagonalEK1 diagonal only, EK1 anything)."""
    import numpy as np
    pd, src = _src("lotka_volterra")
    base = dict(d=2, q=3, n_p=4, rhs_name="UserProb", rhs_src=src, t0=0.0, t1=1.0)
    full, diag = np.array([[0.1, 0.01], [0.01, 0.1]]), np.array([0.1, 0.2])

    def err(**kw):
        with pytest.raises(pnlib.PnodeError) as e:
            pnlib.compile_check(pnlib.make_config(**{**base, **kw}))
        return e.value

    e = err(diffusion=pnlib.DIFF_FIXED, calibrate=True, pn_observation_noise=0.1)
    assert e.kind == "ArgumentError" and "Automatic calibration of global diffusion models is not possible" in str(e)
    e = err(alg=pnlib.EK0, pn_observation_noise=diag)
    assert e.kind == "ArgumentError" and "IsometricKroneckerCovariance" in str(e)
    e = err(alg=pnlib.EK0, pn_observation_noise=full)
    assert e.kind == "ArgumentError" and "IsometricKroneckerCovariance" in str(e)
    e = err(alg=pnlib.DIAGONAL_EK1, pn_observation_noise=full)
    assert e.kind == "ArgumentError" and "BlockDiagonalCovariance" in str(e)
    assert err(pn_observation_noise=np.array([[0.1, 0.2], [0.2, 0.1]])).kind == "ArgumentError"      # not positive definite
    assert err(pn_observation_noise=np.array([[0.1, 0.02], [0.01, 0.1]])).kind == "ArgumentError"    # not symmetric
    assert err(pn_observation_noise=-1.0).kind == "ArgumentError"
    plain = pnlib.compile_check(pnlib.make_config(**base))
    for kw in (dict(pn_observation_noise=0.1), dict(pn_observation_noise=diag), dict(pn_observation_noise=full),
               dict(alg=pnlib.EK0, pn_observation_noise=0.1), dict(alg=pnlib.EK0, pn_observation_noise=np.array([0.1, 0.1])),
               dict(alg=pnlib.DIAGONAL_EK1, pn_observation_noise=diag), dict(alg=pnlib.EK0, diffusion=pnlib.DIFF_DYNAMIC_MV, pn_observation_noise=diag)):
        assert pnlib.compile_check(pnlib.make_config(**{**base, **kw})) > 10_000
    assert pnlib.compile_check(pnlib.make_config(**{**base, "pn_observation_noise": full})) > plain   # the second Gram + Cholesky


def test_config_struct_is_the_same_in_header_python_and_julia():
    """`pnode_config` is declared three times -- include/pnode.h (the ABI), lib.Config (ctypes) and PnodeConfig in the Julia host
    (julia/ProbNumDiffEqB200/src/abi.jl): same field names in the same order, and the ctypes size equals what the library checks."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "pnode.h")).read()
    body = hdr[hdr.index("typedef struct pnode_config {"):hdr.index("} pnode_config;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    c_fields = []
    for stmt in body[body.index("{") + 1:].split(";"):
        m = re.match(r"\s*(?:const\s+)?(?:char|double|int32_t|int64_t)\s*\*?\s*(.+)$", stmt.strip(), flags=re.S)
        if m:
            c_fields += [n.strip().lstrip("*").strip() for n in m.group(1).split(",")]
    py_fields = [f[0] for f in lib.Config._fields_]
    jl = open(os.path.join(root, "julia", "ProbNumDiffEqB200", "src", "abi.jl")).read()
    jl_body = jl[jl.index("struct PnodeConfig"):]
    jl_body = jl_body[:jl_body.index("\nend")]
    jl_fields = re.findall(r"^\s*(\w+)::", jl_body, flags=re.M)
    assert c_fields == py_fields
    assert c_fields == jl_fields
    assert len(c_fields) > 50


def test_field_ids_and_enums_agree_between_header_and_julia():
    """Every `const PNODE_* = n` of the Julia binding has the value the header's enums give it (field ids, algorithm / prior / diffusion /
    covariance codes, error codes)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "pnode.h")).read(), flags=re.S)
    c_vals = {}
    for body in re.findall(r"enum\s*\w*\s*\{(.*?)\}", hdr, flags=re.S):
        nxt = 0
        for item in body.split(","):
            item = item.strip()
            if not item:
                continue
            m = re.match(r"(\w+)\s*(?:=\s*(-?\d+))?$", item)
            assert m, item
            nxt = int(m.group(2)) if m.group(2) is not None else nxt
            c_vals[m.group(1)] = nxt
            nxt += 1
    jl = open(os.path.join(root, "julia", "ProbNumDiffEqB200", "src", "abi.jl")).read()
    jl_vals = {}
    for names, vals in re.findall(r"^const\s+((?:PNODE_\w+\s*,\s*)*PNODE_\w+)\s*=\s*([-\d\s,]+)$", jl, flags=re.M):
        ns = [x.strip() for x in names.split(",")]
        vs = [int(x) for x in vals.split(",")]
        assert len(ns) == len(vs), names
        jl_vals.update(zip(ns, vs))
    assert len(jl_vals) > 30
    for k, v in jl_vals.items():
        assert k in c_vals, k
        assert c_vals[k] == v, (k, v, c_vals[k])


def test_scalar_rate_ioup_beyond_the_register_kernels_goes_to_the_general_path(pnlib):
    """IOUP(q, r) with 24 < D <= 64 compiles (as IOUP(q, r I) on the general-prior path, pnode_api.cu::promote_scalar_rate); beyond
    D = 64, or with dense output, the error names the limits."""
    pd, src = _src("pleiades")
    kw = dict(d=28, n_p=0, rhs_name="UserProb", rhs_src=src, t0=0.0, t1=0.1, prior=pnlib.PRIOR_IOUP, rate_parameter=-0.5)
    assert pnlib.compile_check(pnlib.make_config(q=1, **kw)) > 10_000
    for bad in (dict(q=3), dict(q=1, smooth=True, save_everystep=True, saveat=[0.0, 0.05])):
        with pytest.raises(pnlib.PnodeError) as e:
            pnlib.compile_check(pnlib.make_config(**{**kw, **bad}))
        assert "up to D = 64 on the general-prior path" in str(e.value)


def test_structured_layouts_beyond_the_register_budget_compile(pnlib):
    """EK0 at order 8 / 11 (Kronecker layout) and the blocks layout at d = 28 (DiagonalEK1, EK0 + DynamicMVDiffusion on Pleiades) compile:
    the factors then live in thread-local memory.  Beyond d (q+1)^2 = 512 the blocks kernel says so."""
    pd, src = _src("lotka_volterra")
    for q in (8, 11):
        assert pnlib.compile_check(pnlib.make_config(d=2, q=q, n_p=4, rhs_name="UserProb", rhs_src=src, t0=0.0, t1=1.0, alg=pnlib.EK0)) > 10_000
    pd, src = _src("pleiades")
    kw = dict(d=28, n_p=0, rhs_name="UserProb", rhs_src=src, t0=0.0, t1=0.1)
    assert pnlib.compile_check(pnlib.make_config(q=3, alg=pnlib.DIAGONAL_EK1, **kw)) > 10_000
    assert pnlib.compile_check(pnlib.make_config(q=3, alg=pnlib.EK0, diffusion=pnlib.DIFF_DYNAMIC_MV, **kw)) > 10_000
    with pytest.raises(pnlib.PnodeError) as e:
        pnlib.compile_check(pnlib.make_config(q=5, alg=pnlib.DIAGONAL_EK1, **kw))
    assert e.value.kind == "Unsupported" and "512" in str(e.value)


def test_nvrtc_results_are_cached_per_process(pnlib, monkeypatch):
    """A second solver for the same right-hand side and compile-time configuration does not compile again (pnode_api.cu::nvrtc_compile):
    run-time parameters (tspan, tolerances) are not part of the key, compile-time ones (order) are; PNODE_NVRTC_CACHE=0 turns it off."""
    import time
    pd, src = _src("orego")
    kw = dict(d=3, n_p=3, rhs_name="UserProb", rhs_src=src + "\n// cache test\n", t0=0.0)

    def timed(**extra):
        t = time.perf_counter()
        n = pnlib.compile_check(pnlib.make_config(**{**kw, **extra}))
        return n, time.perf_counter() - t

    n1, cold = timed(q=3, t1=1.0)
    n2, warm = timed(q=3, t1=2.0, abstol=1e-8)
    assert n1 == n2 and warm < 0.2 * cold
    n3, other = timed(q=2, t1=1.0)
    assert n3 != n1 and other > 5 * warm
    monkeypatch.setenv("PNODE_NVRTC_CACHE", "0")
    n4, off = timed(q=3, t1=1.0)
    assert n4 == n1 and off > 5 * warm


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the driver's reference arm; the oracle port on the host cores) on a small sample: one JSON line with the
    keys of the bench contract, `impl: reference`, the e2e object equal to the line's value, and the same metric / unit / workload names the
    GPU arm prints."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-sample", "512"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["dtype"] == "f64" and line["vs_baseline"] is None
    assert line["unit"] == "accepted trajectory-steps/s" and "rober" in line["config"]["workload"]
    assert line["value"] > 0 and line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
