"""Symbolic tracing of user right-hand sides into CUDA C++ device functions.

The reference takes an ordinary Julia function ``f(du, u, p, t)`` and (for the EK1) differentiates it
with ForwardDiff (``src/derivative_utils.jl:1-33``) and TaylorIntegration (``src/initialization/
autodiffinit.jl:52-67``).  north_star asks for the B200 build to trace ``f`` symbolically and hand the
result to NVRTC.  On the Julia host that is Symbolics.jl; in this container (no Julia) the same job is
done with SymPy: the user's Python ``f(du, u, p, t)`` is called once with symbols, and we print

* ``f``   as a device function templated on the scalar type, so the same code runs on ``double`` (every
          filter step) and on the truncated Taylor type of ``csrc/kernels/pn_taylor.cuh`` (initial
          derivatives, the TaylorModeInit of the reference), and
* ``jac`` as a plain ``double`` device function holding the analytic Jacobian (``df_a/du_i`` at
          ``J[a + d*i]``, column-major like Julia).

The C ABI takes the resulting source *string* (``pnode_config.rhs_src``), so a Symbolics.jl front-end
feeds the same boundary.
"""
from __future__ import annotations

import dataclasses
from typing import Callable, List, Sequence

import sympy as sp
from sympy.printing.c import C99CodePrinter


# Polymorphic elementary functions users can call inside f (work on floats and on symbols).
def _poly(name):
    spf = getattr(sp, name)

    def fn(x):
        return spf(x)

    fn.__name__ = name
    return fn


exp = _poly("exp")
log = _poly("log")
sin = _poly("sin")
cos = _poly("cos")
sqrt = _poly("sqrt")
tanh = _poly("tanh")


@dataclasses.dataclass
class TracedRHS:
    d: int
    n_p: int
    u: List[sp.Symbol]
    p: List[sp.Symbol]
    t: sp.Symbol
    f: List[sp.Expr]

    def jacobian(self) -> sp.Matrix:
        return sp.Matrix(self.f).jacobian(sp.Matrix(self.u))


def trace(f: Callable, d: int, n_p: int) -> TracedRHS:
    """Call ``f(du, u, p, t)`` (SciML in-place signature) with symbols and record ``du``."""
    u = [sp.Symbol(f"u{i}", real=True) for i in range(d)]
    p = [sp.Symbol(f"p{i}", real=True) for i in range(n_p)]
    t = sp.Symbol("t", real=True)
    du = [sp.Integer(0)] * d
    du = list(du)
    ret = f(du, u, p, t)
    if isinstance(ret, (list, tuple)):
        du = list(ret)
    exprs = [sp.sympify(e) for e in du]
    if len(exprs) != d:
        raise ValueError(f"f wrote {len(exprs)} components, expected d={d}")
    return TracedRHS(d=d, n_p=n_p, u=u, p=p, t=t, f=exprs)


class _DevicePrinter(C99CodePrinter):
    """Prints expressions using the pn_* helper overloads of csrc/kernels/pn_taylor.cuh."""

    _fn = {"exp": "pn_exp", "log": "pn_log", "sin": "pn_sin", "cos": "pn_cos", "tanh": "pn_tanh"}

    def _print_Pow(self, expr):
        base, e = expr.base, expr.exp
        b = self._print(base)
        if e.is_Integer:
            n = int(e)
            if n == 2:
                return f"pn_sq({b})"
            if n == -1:
                return f"(1.0/({b}))"
            if n > 0:
                return f"pn_ipow({b}, {n})"
            return f"(1.0/pn_ipow({b}, {-n}))"
        if e == sp.Rational(1, 2):
            return f"pn_sqrt({b})"
        if e == sp.Rational(-1, 2):
            return f"(1.0/pn_sqrt({b}))"
        if e.is_Rational and int(e.q) == 2:
            return f"pn_pow_half({b}, {int(e.p)})"   # x^(k/2) = sqrt(x)^k: no pow() call on the step path
        if e.is_Rational:
            return f"pn_pow({b}, {float(e)!r})"
        return f"pn_pow({b}, {self._print(e)})"

    def _print_Rational(self, expr):
        return f"({int(expr.p)}.0/{int(expr.q)}.0)"

    def _print_Integer(self, expr):
        return f"{int(expr)}.0"

    def _print_Function(self, expr):
        name = expr.func.__name__
        if name in self._fn:
            return f"{self._fn[name]}({self._print(expr.args[0])})"
        return super()._print_Function(expr)

    def _print_Float(self, expr):
        return repr(float(expr))

    def _pn1(self, name, expr):
        return f"{name}({self._print(expr.args[0])})"

    def _print_exp(self, expr):
        return self._pn1("pn_exp", expr)

    def _print_log(self, expr):
        return self._pn1("pn_log", expr)

    def _print_sin(self, expr):
        return self._pn1("pn_sin", expr)

    def _print_cos(self, expr):
        return self._pn1("pn_cos", expr)

    def _print_tanh(self, expr):
        return self._pn1("pn_tanh", expr)


def _subs_arrays(tr: TracedRHS):
    rep = {}
    for i, s in enumerate(tr.u):
        rep[s] = sp.Symbol(f"u[{i}]", real=True)
    for i, s in enumerate(tr.p):
        rep[s] = sp.Symbol(f"p[{i}]", real=True)
    return rep


def _emit_block(exprs: Sequence[sp.Expr], targets: Sequence[str], scalar: str, printer, prefix: str) -> str:
    repl, reduced = sp.cse(list(exprs), symbols=sp.numbered_symbols(prefix), optimizations="basic")
    lines = []
    for s, e in repl:
        lines.append(f"    const {scalar} {s} = {printer.doprint(e)};")
    for tgt, e in zip(targets, reduced):
        lines.append(f"    {tgt} = {printer.doprint(e)};")
    return "\n".join(lines)


def cuda_source(tr: TracedRHS, struct_name: str = "PnodeUserProblem", jac_parts: int = 0) -> str:
    """CUDA C++ source defining ``struct <struct_name>`` with static ``d``, ``n_p``, ``f<T>``, ``jac`` and
    ``f_part<PART>`` / ``jac_part<PART>`` (``jac_parts`` independent slices of f and of the Jacobian, each with its own common
    subexpressions, so that the CTA-per-trajectory kernel can evaluate a large right-hand side on several threads at once;
    0 = automatic = 1: on Pleiades, d = 28, slices were slower than one thread each for f and J).

    This is the string handed to ``pnode_config.rhs_src`` and compiled by NVRTC for sm_100a together
    with the step kernels (csrc/kernels/*.cuh)."""
    pr = _DevicePrinter()
    rep = _subs_arrays(tr)
    fex = [e.xreplace(rep) for e in tr.f]
    J = tr.jacobian()
    if jac_parts <= 0:
        jac_parts = 1   # measured on Pleiades (d = 28): slicing f / J over the warps of the CTA kernel does not pay (i-cache)
    fbody = _emit_block(fex, [f"du[{i}]" for i in range(tr.d)], "T", pr, "x")
    fparts = []
    for k in range(jac_parts):
        idx = [a for a in range(tr.d) if a % jac_parts == k]
        body = _emit_block([fex[a] for a in idx], [f"du[{a}]" for a in idx], "double", pr, f"x{k}_").replace("\n    ", "\n      ")
        fparts.append(f"    {'if' if k == 0 else 'else if'} constexpr (PART == {k}) {{\n  {body}\n    }}")
    fpbody = "\n".join(fparts)
    parts = []
    for k in range(jac_parts):
        jex, jt = [], []
        for a in range(tr.d):          # row a: rows are dealt round-robin, a row shares the most subexpressions
            if a % jac_parts != k:
                continue
            for i in range(tr.d):      # column i
                jex.append(J[a, i].xreplace(rep))
                jt.append(f"J[{a + tr.d * i}]")
        body = _emit_block(jex, jt, "double", pr, f"y{k}_").replace("\n    ", "\n      ")
        parts.append(f"    {'if' if k == 0 else 'else if'} constexpr (PART == {k}) {{\n  {body}\n    }}")
    pbody = "\n".join(parts)
    calls = "\n".join(f"    jac_part<{k}>(J, u, p, t);" for k in range(jac_parts))
    return f"""// generated by probnumdiffeq.jl_b200/tracer.py -- do not edit
struct {struct_name} {{
  static constexpr int d = {tr.d};
  static constexpr int n_p = {tr.n_p};
  static constexpr int jac_parts = {jac_parts};
  template <class T>
  __device__ __forceinline__ static void f(T* du, const T* u, const double* p, const T& t) {{
{fbody}
  }}
  // components a == PART (mod jac_parts) of f, double precision (the CTA-per-trajectory kernel spreads f and J over warps)
  template <int PART>
  __device__ __forceinline__ static void f_part(double* du, const double* u, const double* p, double t) {{
{fpbody}
  }}
  // rows a == PART (mod jac_parts) of J (column-major, J[a + d*i] = d f_a / d u_i)
  template <int PART>
  __device__ __forceinline__ static void jac_part(double* J, const double* u, const double* p, double t) {{
{pbody}
  }}
  __device__ __forceinline__ static void jac(double* J, const double* u, const double* p, double t) {{
{calls}
  }}
}};
"""


BUILTIN_STRUCTS = {"fitzhugh_nagumo": "PnFitzHughNagumo", "lotka_volterra": "PnLotkaVolterra", "vanderpol": "PnVanDerPol",
                   "rober": "PnRober", "orego": "PnOrego", "pleiades": "PnPleiades"}


def builtin_header() -> str:
    """csrc/kernels/pn_builtin_problems.cuh: the problems BASELINE.json names (problems.py), traced ahead of time."""
    from . import problems

    out = ["// pn_builtin_problems.cuh -- generated by probnumdiffeq.jl_b200/tracer.py from problems.py; do not edit.",
           "// (regenerate: python scripts/gen_builtin_problems.py)",
           '// Built-in right-hand sides named by BASELINE.json (rhs_name = "builtin:<name>").',
           "#ifndef PN_BUILTIN_PROBLEMS_CUH", "#define PN_BUILTIN_PROBLEMS_CUH", '#include "pn_taylor.cuh"', ""]
    for name, sname in BUILTIN_STRUCTS.items():
        pd = problems.PROBLEMS[name]
        src = cuda_source(trace(pd.f, pd.d, pd.n_p), sname)
        out.append(f"// builtin:{name}")
        out.append(src.split("\n", 1)[1])
    out.append("#endif /* PN_BUILTIN_PROBLEMS_CUH */")
    return "\n".join(out) + "\n"
