"""The ODE problems BASELINE.json names, written in SciML's in-place signature ``f(du, u, p, t)``.

Definitions follow the reference's own benchmark notebooks / tutorials (cited per function); the
functions are plain Python working on floats *and* on SymPy symbols, so they can be traced by
``tracer.trace`` into CUDA C++ for NVRTC.
"""
from __future__ import annotations

import dataclasses
from typing import Callable, Optional, Sequence, Tuple


@dataclasses.dataclass
class ProblemDef:
    name: str
    f: Callable
    u0: Sequence[float]
    p: Sequence[float]
    tspan: Tuple[float, float]
    mass_matrix: Optional[Sequence[Sequence[float]]] = None   # M of M u' = f (None = identity)
    second_order: bool = False   # f is the first-order form F([du; u]) = [f1(du, u); du] of a SecondOrderODEProblem, u0 = (du0, u0)

    @property
    def d(self) -> int:
        return len(self.u0)

    @property
    def n_p(self) -> int:
        return len(self.p)


def fitzhugh_nagumo(du, u, p, t):
    """README.md:35-43 / docs/src/tutorials/getting_started.md:35-43."""
    a, b, c = p[0], p[1], p[2]
    du[0] = c * (u[0] - u[0] ** 3 / 3 + u[1])
    du[1] = -(1 / c) * (u[0] - a - b * u[1])


def lotka_volterra(du, u, p, t):
    """benchmarks/lotkavolterra.jmd:46-53."""
    a, b, c, dd = p[0], p[1], p[2], p[3]
    du[0] = a * u[0] - b * u[0] * u[1]
    du[1] = -c * u[1] + dd * u[0] * u[1]


def vanderpol(du, u, p, t):
    """benchmarks/vanderpol.jmd:35-42."""
    du[0] = u[1]
    du[1] = p[0] * ((1 - u[0] ** 2) * u[1] - u[0])


def rober(du, u, p, t):
    """ROBER in ODE form (the reference notebook benchmarks/rober.jmd:37-46 uses the mass-matrix DAE
    form; SURVEY.md H9 -- the ODE form replaces the algebraic constraint by its derivative)."""
    k1, k2, k3 = p[0], p[1], p[2]
    du[0] = -k1 * u[0] + k3 * u[1] * u[2]
    du[1] = k1 * u[0] - k3 * u[1] * u[2] - k2 * u[1] ** 2
    du[2] = k2 * u[1] ** 2


def rober_dae(du, u, p, t):
    """ROBER as the reference benchmarks it: mass-matrix DAE, M = Diagonal([1, 1, 0]), third row the algebraic constraint
    (benchmarks/rober.jmd:37-46; test/mass_matrix.jl:64-79)."""
    k1, k2, k3 = p[0], p[1], p[2]
    du[0] = -k1 * u[0] + k3 * u[1] * u[2]
    du[1] = k1 * u[0] - k3 * u[1] * u[2] - k2 * u[1] ** 2
    du[2] = u[0] + u[1] + u[2] - 1


def twobody_first_order(du, u, p, t):
    """Two-body problem of test/secondorderode.jl:15-19 (with a gravitational parameter p[0]) in the first-order form of a
    SecondOrderODEProblem: x = (v1, v2, r1, r2)."""
    r3 = (u[2] * u[2] + u[3] * u[3]) ** 1.5
    du[0] = -p[0] * u[2] / r3
    du[1] = -p[0] * u[3] / r3
    du[2] = u[0]
    du[3] = u[1]


def orego(du, u, p, t):
    """benchmarks/orego.jmd:37-46."""
    p1, p2, p3 = p[0], p[1], p[2]
    du[0] = p1 * (u[1] + u[0] * (1 - p2 * u[0] - u[1]))
    du[1] = (u[2] - (1 + u[0]) * u[1]) / p1
    du[2] = p3 * (u[0] - u[2])


def pleiades(du, u, p, t):
    """benchmarks/pleiades.jmd:35-61, first-order form u = [x'; y'; x; y], d = 28."""
    x = u[14:21]
    y = u[21:28]
    for i in range(7):
        du[14 + i] = u[i]
        du[21 + i] = u[7 + i]
    for i in range(7):
        ax = 0
        ay = 0
        for j in range(7):
            if i != j:
                r2 = (x[i] - x[j]) ** 2 + (y[i] - y[j]) ** 2
                r = r2 * _sqrt(r2)
                ax = ax + (j + 1) * (x[j] - x[i]) / r
                ay = ay + (j + 1) * (y[j] - y[i]) / r
        du[i] = ax
        du[7 + i] = ay


def _sqrt(x):
    if isinstance(x, (int, float)):
        return float(x) ** 0.5
    import sympy as _sp

    return _sp.sqrt(x)


PROBLEMS = {
    "fitzhugh_nagumo": ProblemDef("fitzhugh_nagumo", fitzhugh_nagumo, [-1.0, 1.0], [0.2, 0.2, 3.0], (0.0, 20.0)),
    "lotka_volterra": ProblemDef("lotka_volterra", lotka_volterra, [1.0, 1.0], [1.5, 1.0, 3.0, 1.0], (0.0, 10.0)),
    "vanderpol": ProblemDef("vanderpol", vanderpol, [2.0, 0.0], [1e3], (0.0, 2.0)),
    "rober": ProblemDef("rober", rober, [1.0, 0.0, 0.0], [0.04, 3e7, 1e4], (0.0, 100.0)),
    "orego": ProblemDef("orego", orego, [1.0, 2.0, 3.0], [77.27, 8.375e-6, 0.161], (0.0, 30.0)),
    "pleiades": ProblemDef(
        "pleiades",
        pleiades,
        [0, 0, 0, 0, 0, 1.75, -1.5] + [0, 0, 0, -1.25, 1, 0, 0] + [3.0, 3.0, -1.0, -3.0, 2.0, -2.0, 2.0]
        + [3.0, -3.0, 2.0, 0, 0, -4.0, 4.0],
        [],
        (0.0, 3.0),
    ),
}

# Mass-matrix problems: traced and compiled at run time (the ahead-of-time kernel table holds ODE-form problems only).
MASS_MATRIX_PROBLEMS = {
    "rober_dae": ProblemDef("rober_dae", rober_dae, [1.0, 0.0, 0.0], [0.04, 3e7, 1e4], (0.0, 1e5),
                            mass_matrix=[[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 0.0]]),
}

# Second-order problems (SecondOrderODEProblem): `d` of the solver is len(u0) // 2
SECOND_ORDER_PROBLEMS = {
    "twobody": ProblemDef("twobody", twobody_first_order, [0.0, 2.0, 0.4, 0.0], [1.0], (0.0, 0.1), second_order=True),
}


def lookup(name: str) -> ProblemDef:
    for table in (PROBLEMS, MASS_MATRIX_PROBLEMS, SECOND_ORDER_PROBLEMS):
        if name in table:
            return table[name]
    raise KeyError(name)
