"""Host-side mirror of the reference's user-facing surface for the ODE-filter path.

Reference (Julia)                                         here (Python stand-in for the Julia host package)
--------------------------------------------------------  -------------------------------------------------
ODEProblem(f, u0, tspan, p)                                ODEProblem(f, u0, tspan, p)
EK0(; order, smooth, diffusionmodel) src/algorithms.jl:188  EK0(order=3, smooth=True, diffusionmodel=DynamicDiffusion())
EK1(; order, smooth, diffusionmodel) src/algorithms.jl:250  EK1(...)
DiagonalEK1(; ...)                   src/algorithms.jl:343  DiagonalEK1(...)   (BlocksOfDiagonals layout)
DynamicMVDiffusion(), FixedMVDiffusion(...)  src/diffusions/typedefs.jl   same names (EK0 / DiagonalEK1)
DynamicDiffusion(), FixedDiffusion(; initial_diffusion, calibrate)  src/diffusions/typedefs.jl   same names
EnsembleProblem(prob; prob_func) (SciMLBase)               EnsembleProblem(prob, prob_func=...)
solve(ensprob, alg, EnsembleThreads(); trajectories, ...)  solve(ensprob, alg, EnsembleB200(), trajectories=N, ...)
sol.t, sol.u, sol.pu, sol.stats, sol.retcode  src/solution.jl:33-56   ProbODESolution with the same field names

`EnsembleB200` gathers every trajectory's (u0, p), traces `f` once (tracer.py) and makes ONE call into
libpnode_cuda.so (include/pnode.h); nothing here computes filter arithmetic on the host.
Exceptions mirror the reference: ValueError <-> ArgumentError / DimensionMismatch, RuntimeError <-> ErrorException.
"""
from __future__ import annotations

import dataclasses
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import lib as _lib
from . import problems as _problems
from . import tracer as _tracer


# ---- diffusion models (src/diffusions/typedefs.jl) -------------------------------------------------
@dataclasses.dataclass(frozen=True)
class DynamicDiffusion:
    pass


@dataclasses.dataclass(frozen=True)
class FixedDiffusion:
    initial_diffusion: float = 1.0
    calibrate: bool = True


@dataclasses.dataclass(frozen=True)
class DynamicMVDiffusion:
    """Time-varying diagonal diffusion (src/diffusions/typedefs.jl); EK0 / DiagonalEK1 only."""
    pass


@dataclasses.dataclass(frozen=True)
class FixedMVDiffusion:
    """Time-fixed diagonal diffusion, optionally calibrated (src/diffusions/typedefs.jl); EK0 / DiagonalEK1 only."""
    initial_diffusion: float = 1.0
    calibrate: bool = True


@dataclasses.dataclass(frozen=True)
class IWP:
    """Integrated Wiener process prior (src/priors/iwp.jl)."""
    num_derivatives: int = 3


@dataclasses.dataclass(frozen=True)
class IOUP:
    """Integrated Ornstein-Uhlenbeck prior (src/priors/ioup.jl): `IOUP(num_derivatives, rate_parameter)` with a number, a length-d
    vector or a d x d matrix, or `IOUP(num_derivatives, update_rate_parameter=True)` (rate := Jacobian before every step,
    src/perform_step.jl:88-96; the rate must then be left unset, src/caches.jl:139-149)."""
    num_derivatives: int = 3
    rate_parameter: object = None
    update_rate_parameter: bool = False

    def __post_init__(self):
        if self.rate_parameter is None and not self.update_rate_parameter:
            raise TypeError("IOUP(num_derivatives; update_rate_parameter) needs update_rate_parameter=True when no rate is given")  # ioup.jl:51-54
        if self.rate_parameter is not None and self.update_rate_parameter:
            raise ValueError("Do not manually set the `rate_parameter` of the IOUP prior when using the `update_rate_parameter=true` option."
                             "Reset the prior and try again.")  # caches.jl:139-149

    def __hash__(self):
        r = self.rate_parameter
        return hash((self.num_derivatives, None if r is None else np.asarray(r, dtype=float).tobytes(), self.update_rate_parameter))


def ExpEK(*, L, order=3, **kw):
    """`ExpEK(; L, order=3, kwargs...) = EK0(; prior=IOUP(order, L), kwargs...)` (src/algorithms.jl:405-425)."""
    return EK0(prior=IOUP(order, L), **kw)


def RosenbrockExpEK(*, order=3, **kw):
    """`RosenbrockExpEK(; order=3, kwargs...) = EK1(; prior=IOUP(order, update_rate_parameter=true), kwargs...)` (src/algorithms.jl:438-459)."""
    return EK1(prior=IOUP(order, update_rate_parameter=True), **kw)


@dataclasses.dataclass(frozen=True)
class _EK:
    order: int = 3
    smooth: bool = True
    diffusionmodel: object = DynamicDiffusion()
    prior: Optional[object] = None   # IWP(q) or IOUP(q, rate)
    # observation noise of the ODE measurement z = 0 (src/algorithms.jl:195-393): a number (R = x I), a 1-d array (Diagonal) or a
    # d x d covariance matrix; the layout rules of ekargcheck (src/algorithms.jl:78-130) are checked by the library
    pn_observation_noise: Optional[object] = None

    @property
    def q(self) -> int:
        return self.prior.num_derivatives if self.prior is not None else self.order


class EK0(_EK):
    """Gaussian ODE filter with zeroth-order linearisation (src/algorithms.jl:188-208)."""
    alg = _lib.EK0


class EK1(_EK):
    """Gaussian ODE filter with first-order linearisation (src/algorithms.jl:250-296)."""
    alg = _lib.EK1


class DiagonalEK1(_EK):
    """EK1 with a diagonal Jacobian approximation on the BlocksOfDiagonals layout (src/algorithms.jl:343-393)."""
    alg = _lib.DIAGONAL_EK1


@dataclasses.dataclass
class ODEProblem:
    f: Callable            # in-place SciML signature f(du, u, p, t)
    u0: Sequence[float]
    tspan: Tuple[float, float]
    p: Sequence[float] = ()
    builtin: Optional[str] = None   # name in problems.PROBLEMS -> ahead-of-time kernels when available
    mass_matrix: Optional[Sequence[Sequence[float]]] = None   # ODEFunction(f; mass_matrix=M): M u' = f(u, p, t), d x d, may be singular
    second_order: bool = False      # built by SecondOrderODEProblem: f is the first-order form on (du, u), u0 = (du0, u0)

    @staticmethod
    def from_library(name: str) -> "ODEProblem":
        if name in _problems.MASS_MATRIX_PROBLEMS:
            pd = _problems.MASS_MATRIX_PROBLEMS[name]
            return ODEProblem(pd.f, list(pd.u0), tuple(pd.tspan), list(pd.p), mass_matrix=pd.mass_matrix)
        pd = _problems.PROBLEMS[name]
        return ODEProblem(pd.f, list(pd.u0), tuple(pd.tspan), list(pd.p), builtin=name)


def SecondOrderODEProblem(f1: Callable, du0: Sequence[float], u0: Sequence[float], tspan: Tuple[float, float],
                          p: Sequence[float] = ()) -> ODEProblem:
    """`SecondOrderODEProblem(f1, du0, u0, tspan, p)` with the in-place signature `f1(ddu, du, u, p, t)`.  The filter models
    u and its derivatives and measures `E2 x - f1(E1 x, E0 x)` (src/measurement_models.jl:29-49); `sol.u` holds
    `(du, u)` as the reference's ArrayPartition does (SolProj = [E1; E0], src/caches.jl:126)."""
    d = len(u0)
    if len(du0) != d:
        raise ValueError("du0 and u0 must have the same length")  # DimensionMismatch

    def first_order_form(dx, x, p_, t):   # F([du; u]) = [f1(du, u); du], what DynamicalODEFunction evaluates
        ddu = [0] * d
        f1(ddu, x[:d], x[d:], p_, t)
        for i in range(d):
            dx[i] = ddu[i]
            dx[d + i] = x[i]

    return ODEProblem(first_order_form, list(du0) + list(u0), tuple(tspan), list(p), second_order=True)


def remake(prob: ODEProblem, **kw) -> ODEProblem:
    return dataclasses.replace(prob, **kw)


@dataclasses.dataclass
class EnsembleProblem:
    prob: ODEProblem
    prob_func: Optional[Callable] = None   # (prob, i, repeat) -> prob   (SciMLBase default: identity)


@dataclasses.dataclass(frozen=True)
class EnsembleB200:
    """Ensemble algorithm: all trajectories in one persistent-kernel launch per B200.

    `devices` (a tuple of device ordinals, or "all") shards the ensemble over several GPUs of the node: contiguous trajectory
    ranges (`shard_range`), one solver + host thread + stream per device inside the library
    (`pnode_solve_ensemble_sharded`), no exchange between the shards; the solutions come back concatenated in trajectory order
    (SURVEY.md section 8e).  Default: the single device `device`."""
    device: int = 0
    lanes_per_traj: int = 0
    devices: object = None

    def device_list(self) -> Tuple[int, ...]:
        if self.devices is None:
            return (self.device,)
        if isinstance(self.devices, str):
            if self.devices != "all":
                raise ValueError('devices must be a sequence of device ordinals or "all"')
            return tuple(range(max(_lib.device_count(), 1)))
        devs = tuple(int(x) for x in self.devices)
        if not devs or len(set(devs)) != len(devs):
            raise ValueError("devices must be a non-empty sequence of distinct device ordinals")
        return devs


@dataclasses.dataclass
class Stats:
    nf: int
    naccept: int
    nreject: int


@dataclasses.dataclass
class Gaussian:
    mu: np.ndarray      # mean
    Sigma: np.ndarray   # dense covariance


@dataclasses.dataclass
class ProbODESolution:
    """Field names follow src/solution.jl:33-56."""
    t: np.ndarray
    u: np.ndarray                 # [n, d]
    pu_cov: np.ndarray            # [n, d, d]   Matrix(sol.pu[i].Sigma)
    diffusions: Optional[np.ndarray]
    stats: Stats
    log_likelihood: float
    retcode: str
    saveat_t: Optional[np.ndarray] = None     # dense output on the requested grid: mean(sol(t)), Matrix(sol(t).Sigma)
    saveat_u: Optional[np.ndarray] = None
    saveat_cov: Optional[np.ndarray] = None

    @property
    def pu(self) -> List[Gaussian]:
        return [Gaussian(m, c) for m, c in zip(self.u, self.pu_cov)]


@dataclasses.dataclass
class EnsembleSolution:
    u: List[ProbODESolution]
    elapsedTime: float
    converged: bool
    info: object = None

    def __len__(self):
        return len(self.u)

    def __getitem__(self, i):
        return self.u[i]


def _check_alg(alg, adaptive, dt, save_everystep, dense):
    if not isinstance(alg, _EK):
        raise ValueError("alg must be EK0(...), EK1(...) or DiagonalEK1(...)")
    if not adaptive and not (dt and dt > 0):
        # test/errors_thrown.jl:6-8
        raise ValueError("Fixed timestep methods require a choice of dt or choosing the tstops")
    if dense and not alg.smooth:
        raise RuntimeError("To use `dense=true` you need to set `smooth=true`!")  # src/checks.jl:15-17
    if not save_everystep and alg.smooth:
        raise RuntimeError("If you set `save_everystep=false` also set `smooth=false` in the alg!")  # checks.jl:18-20
    if not isinstance(alg.diffusionmodel, (DynamicDiffusion, FixedDiffusion, DynamicMVDiffusion, FixedMVDiffusion)):
        raise ValueError("unsupported diffusion model")
    if isinstance(alg, EK1):  # ekargcheck, src/algorithms.jl:93-107
        if isinstance(alg.diffusionmodel, FixedMVDiffusion) and alg.diffusionmodel.calibrate:
            raise ValueError("The `EK1` algorithm does not support automatic global calibration of multivariate diffusion "
                             "models.")
        if isinstance(alg.diffusionmodel, DynamicMVDiffusion):
            raise ValueError("The `EK1` algorithm does not support automatic calibration of local multivariate diffusion "
                             "models.")


def _config_for(prob: ODEProblem, alg, ens: EnsembleB200, *, adaptive, dt, abstol, reltol, save_everystep, maxiters,
                dtmin, dtmax, tstops, save_state, forced_diffusion=None, saveat=None, fenrir=None):
    d, n_p = len(prob.u0), len(prob.p)
    if d < 1:
        raise RuntimeError("We currently don't support scalar-valued problems")  # src/caches.jl:96-98
    dim_rhs = d
    if prob.second_order:
        d = d // 2   # positions; the traced right-hand side keeps the first-order form of dimension 2d
    diff = alg.diffusionmodel
    kw = dict(second_order=prob.second_order, d=d, q=alg.q, n_p=n_p, alg=alg.alg, smooth=alg.smooth, adaptive=adaptive, dt=dt or 0.0,
              abstol=abstol, reltol=reltol, save_everystep=save_everystep, maxiters=maxiters, dtmin=dtmin,
              dtmax=dtmax or 0.0, t0=prob.tspan[0], t1=prob.tspan[1], tstops=tstops, save_state=save_state,
              device=ens.device, lanes_per_traj=ens.lanes_per_traj, forced_diffusion=forced_diffusion, saveat=saveat)
    if isinstance(diff, FixedDiffusion):
        kw.update(diffusion=_lib.DIFF_FIXED, initial_diffusion=diff.initial_diffusion, calibrate=diff.calibrate)
    elif isinstance(diff, FixedMVDiffusion):
        kw.update(diffusion=_lib.DIFF_FIXED_MV, initial_diffusion=diff.initial_diffusion, calibrate=diff.calibrate)
    elif isinstance(diff, DynamicMVDiffusion):
        kw.update(diffusion=_lib.DIFF_DYNAMIC_MV)
    else:
        kw.update(diffusion=_lib.DIFF_DYNAMIC)
    if isinstance(alg.prior, IOUP):
        kw.update(prior=_lib.PRIOR_IOUP, rate_parameter=alg.prior.rate_parameter if alg.prior.rate_parameter is not None else 0.0,
                  update_rate_parameter=alg.prior.update_rate_parameter)
    if alg.pn_observation_noise is not None:
        kw.update(pn_observation_noise=alg.pn_observation_noise)
    if prob.mass_matrix is not None:
        M = np.asarray(prob.mass_matrix, dtype=float)
        if M.shape != (d, d):
            raise ValueError("mass_matrix must be d x d")  # DimensionMismatch
        if not np.array_equal(M, np.eye(d)):
            uniform = np.array_equal(M, M[0, 0] * np.eye(d)) and M[0, 0] != 0.0
            if isinstance(alg, EK0) and (alg.prior is None or isinstance(alg.prior, IWP)) \
                    and not isinstance(diff, (FixedMVDiffusion, DynamicMVDiffusion)) \
                    and not uniform:
                # src/caches.jl:111-118: only a UniformScaling (here: M == lambda I) passes on the Kronecker layout
                raise ValueError("The selected algorithm uses an efficient Kronecker-factorized implementation which is "
                                 "incompatible with the provided mass matrix. Try using the `EK1` instead.")
            kw.update(mass_matrix=M)
    if fenrir is not None:
        kw.update(**fenrir)
    if prob.builtin is not None:
        kw.update(rhs_name="builtin:" + prob.builtin)
    else:
        tr = _tracer.trace(prob.f, dim_rhs, n_p)
        kw.update(rhs_name="PnodeUserProblem", rhs_src=_tracer.cuda_source(tr, "PnodeUserProblem"))
    return _lib.make_config(**kw)


_solver_cache = {}


def solve(prob, alg, ensemblealg: Optional[EnsembleB200] = None, *, trajectories: Optional[int] = None,
          adaptive: bool = True, dt: Optional[float] = None, abstol: float = 1e-6, reltol: float = 1e-3,
          save_everystep: bool = True, dense: Optional[bool] = None, maxiters: int = 1_000_000, dtmin: float = 0.0,
          dtmax: Optional[float] = None, tstops=None, save_state: bool = False, forced_diffusion=None, saveat=None):
    """`solve(prob, EK1())` for one trajectory, or `solve(EnsembleProblem(...), EK1(), EnsembleB200(); trajectories=N)`."""
    import time

    ens = ensemblealg or EnsembleB200()
    single = isinstance(prob, ODEProblem)
    eprob = EnsembleProblem(prob) if single else prob
    n = 1 if single else trajectories
    if n is None or n < 1:
        raise ValueError("trajectories must be a positive integer")
    if dense is None:
        dense = bool(alg.smooth and save_everystep)
    _check_alg(alg, adaptive, dt, save_everystep, dense)
    base = eprob.prob
    d, n_p = len(base.u0), len(base.p)
    # prob_func(prob, i, repeat) -> prob  (SciMLBase); only u0 / p may vary across trajectories
    u0 = np.empty((n, d))
    p = np.empty((n, max(n_p, 0)))
    for i in range(n):
        pi = eprob.prob_func(base, i, 0) if eprob.prob_func else base
        if len(pi.u0) != d or len(pi.p) != n_p:
            raise ValueError("prob_func must not change the problem dimensions")  # DimensionMismatch
        u0[i] = pi.u0
        if n_p:
            p[i] = pi.p
    cfg = _config_for(base, alg, ens, adaptive=adaptive, dt=dt, abstol=abstol, reltol=reltol,
                      save_everystep=save_everystep, maxiters=maxiters, dtmin=dtmin, dtmax=dtmax, tstops=tstops,
                      save_state=save_state, forced_diffusion=forced_diffusion, saveat=saveat)
    t0 = time.time()
    devs = ens.device_list()
    if len(devs) > n:
        devs = devs[:n]
    solvers = []
    try:
        for dev in devs:   # one solver (kernels, stream, work queue) per device, same configuration
            cfg.device = dev
            solvers.append(_lib.Solver(cfg))
        results = [solvers[0].solve(u0, p)] if len(solvers) == 1 else _lib.solve_sharded(solvers, u0, p)
        sols, infos = [], []
        for k, res in enumerate(results):   # shard k holds the trajectories shard_range(n, k, len(devs)), in order
            lo, hi = _lib.shard_range(n, k, len(results))
            part = _unpack(res, hi - lo, d, save_everystep)
            if saveat is not None and len(saveat):
                # dense output sol(saveat) evaluated on the device (src/solution.jl:330-398)
                su = res.field(_lib.F_SAVEAT_U).reshape(hi - lo, len(saveat), d)
                sc = res.field(_lib.F_SAVEAT_U_COV).reshape(hi - lo, len(saveat), d, d)
                for i, sol_i in enumerate(part):
                    sol_i.saveat_t, sol_i.saveat_u, sol_i.saveat_cov = np.asarray(saveat, dtype=float), su[i], sc[i]
            sols += part
            infos.append(res.info)
            res.free()
    finally:
        for sv in solvers:
            sv.close()
    out = EnsembleSolution(sols, time.time() - t0, all(s.retcode == "Success" for s in sols), infos[0] if len(infos) == 1 else infos)
    return out.u[0] if single else out


def fenrir_data_loglik(prob, alg, ensemblealg: Optional[EnsembleB200] = None, *, data, observation_noise_cov,
                       observation_matrix=None, trajectories: Optional[int] = None, adaptive: bool = True,
                       dt: Optional[float] = None, abstol: float = 1e-6, reltol: float = 1e-3, maxiters: int = 1_000_000,
                       dtmin: float = 0.0, dtmax: Optional[float] = None, tstops=None):
    """`fenrir_data_loglik(prob, alg; data=(t, u), observation_matrix, observation_noise_cov, kwargs...)`
    (src/data_likelihoods/fenrir.jl:30-66): the Fenrir approximate log-likelihood of `data` -- a pair (t, u) -- under the PN
    posterior of the ODE solution.  For an `EnsembleProblem` (a parameter sweep against ONE data set) it returns one value
    per trajectory; the solves and the backward passes of the whole ensemble run on the device."""
    if not alg.smooth:
        raise ValueError("fenrir only works with smoothing. Set `smooth=true`.")   # fenrir.jl:42-44
    ens = ensemblealg or EnsembleB200()
    single = isinstance(prob, ODEProblem)
    eprob = EnsembleProblem(prob) if single else prob
    n = 1 if single else trajectories
    if n is None or n < 1:
        raise ValueError("trajectories must be a positive integer")
    _check_alg(alg, adaptive, dt, True, False)
    base = eprob.prob
    d, n_p = len(base.u0), len(base.p)
    u0 = np.empty((n, d))
    p = np.empty((n, max(n_p, 0)))
    for i in range(n):
        pi = eprob.prob_func(base, i, 0) if eprob.prob_func else base
        u0[i] = pi.u0
        if n_p:
            p[i] = pi.p
    data_t, data_u = data
    fen = dict(data_t=np.asarray(data_t, dtype=float), data_u=np.asarray(data_u, dtype=float),
               observation_matrix=observation_matrix, observation_noise_cov=observation_noise_cov)
    cfg = _config_for(base, alg, ens, adaptive=adaptive, dt=dt, abstol=abstol, reltol=reltol, save_everystep=True,
                      maxiters=maxiters, dtmin=dtmin, dtmax=dtmax, tstops=tstops, save_state=False, fenrir=fen)
    solver = _lib.Solver(cfg)
    try:
        res = solver.solve(u0, p)
        ll = res.field(_lib.F_FENRIR_LOGLIK).copy()
        res.free()
    finally:
        solver.close()
    return float(ll[0]) if single else ll


def _ensemble_inputs(prob, trajectories):
    single = isinstance(prob, ODEProblem)
    eprob = EnsembleProblem(prob) if single else prob
    n = 1 if single else trajectories
    if n is None or n < 1:
        raise ValueError("trajectories must be a positive integer")
    base = eprob.prob
    d, n_p = len(base.u0), len(base.p)
    u0 = np.empty((n, d))
    p = np.empty((n, max(n_p, 0)))
    for i in range(n):
        pi = eprob.prob_func(base, i, 0) if eprob.prob_func else base
        u0[i] = pi.u0
        if n_p:
            p[i] = pi.p
    return single, base, n, u0, p


def _forward_data_solve(prob, alg, ens, *, data, observation_noise_cov, observation_matrix, trajectories, with_data, adaptive, dt, abstol,
                        reltol, maxiters, dtmin, dtmax, tstops):
    """One ensemble solve with save_everystep=false and tstops = union(data.t, tstops); with `with_data` the DataUpdateCallback is
    active.  Returns (data log-likelihoods, PN log-likelihoods, single)."""
    single, base, n, u0, p = _ensemble_inputs(prob, trajectories)
    data_t, data_u = np.asarray(data[0], dtype=float), np.asarray(data[1], dtype=float)
    alg_f = dataclasses.replace(alg, smooth=False)   # the reference only warns that smoothing is wasted here; save_everystep=false
    _check_alg(alg_f, adaptive, dt, False, False)
    if with_data:
        extra = dict(data_t=data_t, data_u=data_u, observation_matrix=observation_matrix, observation_noise_cov=observation_noise_cov,
                     data_forward=True)
        ts = tstops
    else:
        extra = None
        ts = np.union1d(data_t, [] if tstops is None else np.asarray(tstops, dtype=float))
        ts = ts[(ts > base.tspan[0]) & (ts < base.tspan[1])]
    cfg = _config_for(base, alg_f, ens, adaptive=adaptive, dt=dt, abstol=abstol, reltol=reltol, save_everystep=False, maxiters=maxiters,
                      dtmin=dtmin, dtmax=dtmax, tstops=ts, save_state=False, fenrir=extra)
    solver = _lib.Solver(cfg)
    try:
        res = solver.solve(u0, p)
        dll = res.field(_lib.F_DATA_LOGLIK).copy() if with_data else np.zeros(n)
        pll = res.field(_lib.F_LOGLIK).copy()
        res.free()
    finally:
        solver.close()
    return dll, pll, single


def filtering_data_loglik(prob, alg, ensemblealg: Optional[EnsembleB200] = None, *, data, observation_noise_cov, observation_matrix=None,
                          trajectories: Optional[int] = None, adaptive: bool = True, dt: Optional[float] = None, abstol: float = 1e-6,
                          reltol: float = 1e-3, maxiters: int = 1_000_000, dtmin: float = 0.0, dtmax: Optional[float] = None, tstops=None):
    """`filtering_data_loglik(prob, alg; data=(t, u), observation_matrix, observation_noise_cov, kwargs...)`
    (src/data_likelihoods/filtering.jl): the data are conditioned on during the FORWARD pass (`DataUpdateCallback`,
    src/callbacks/dataupdate.jl:39-84) and the updates' log-likelihoods are summed.  One value per trajectory for an
    `EnsembleProblem` (a parameter sweep against one data set)."""
    ens = ensemblealg or EnsembleB200()
    dll, _, single = _forward_data_solve(prob, alg, ens, data=data, observation_noise_cov=observation_noise_cov,
                                         observation_matrix=observation_matrix, trajectories=trajectories, with_data=True, adaptive=adaptive,
                                         dt=dt, abstol=abstol, reltol=reltol, maxiters=maxiters, dtmin=dtmin, dtmax=dtmax, tstops=tstops)
    return float(dll[0]) if single else dll


def dalton_data_loglik(prob, alg, ensemblealg: Optional[EnsembleB200] = None, *, data, observation_noise_cov, observation_matrix=None,
                       trajectories: Optional[int] = None, adaptive: Optional[bool] = None, dt: Optional[float] = None, abstol: float = 1e-6,
                       reltol: float = 1e-3, maxiters: int = 1_000_000, dtmin: float = 0.0, dtmax: Optional[float] = None, tstops=None):
    """`dalton_data_loglik` (src/data_likelihoods/dalton.jl:23-78): data log-likelihood of the solve WITH data updates, plus the PN
    log-likelihood of that solve, minus the PN log-likelihood of the same solve WITHOUT data (both through the data times)."""
    if adaptive is None:
        raise ValueError("`dalton_nll` only works with fixed step sizes. Set `adaptive=false`.")   # dalton.jl:41-44 (keyword must be passed)
    ens = ensemblealg or EnsembleB200()
    kw = dict(data=data, observation_noise_cov=observation_noise_cov, observation_matrix=observation_matrix, trajectories=trajectories,
              adaptive=adaptive, dt=dt, abstol=abstol, reltol=reltol, maxiters=maxiters, dtmin=dtmin, dtmax=dtmax, tstops=tstops)
    dll, pll_with, single = _forward_data_solve(prob, alg, ens, with_data=True, **kw)
    _, pll_without, _ = _forward_data_solve(prob, alg, ens, with_data=False, **kw)
    out = dll + pll_with - pll_without
    return float(out[0]) if single else out


def _unpack(res, n, d, save_everystep) -> List[ProbODESolution]:
    L = _lib
    na, nr, nf = res.field(L.F_NACCEPT), res.field(L.F_NREJECT), res.field(L.F_NF)
    rc, ll = res.field(L.F_RETCODE), res.field(L.F_LOGLIK)
    sols = []
    if save_everystep and res.has(L.F_STEP_OFFSETS):
        off = res.field(L.F_STEP_OFFSETS)
        T = res.field(L.F_T)
        U = res.field(L.F_U).reshape(-1, d)
        Cv = res.field(L.F_U_COV).reshape(-1, d, d)
        Df = res.field(L.F_DIFFUSIONS_MV).reshape(-1, d) if res.has(L.F_DIFFUSIONS_MV) else res.field(L.F_DIFFUSIONS)
        for i in range(n):
            a, b = off[i], off[i + 1]
            sols.append(ProbODESolution(T[a:b], U[a:b], Cv[a:b], Df[a:b - 1], Stats(int(nf[i]), int(na[i]), int(nr[i])),
                                        float(ll[i]), L.RETCODES[int(rc[i])]))
    else:
        te = res.field(L.F_T_END)
        uf = res.field(L.F_U_FINAL).reshape(n, d)
        cf = res.field(L.F_U_FINAL_COV).reshape(n, d, d)
        for i in range(n):
            sols.append(ProbODESolution(np.array([te[i]]), uf[i:i + 1], cf[i:i + 1], None,
                                        Stats(int(nf[i]), int(na[i]), int(nr[i])), float(ll[i]), L.RETCODES[int(rc[i])]))
    return sols


# ---- multi-GPU: trajectory sharding + optional NCCL statistics ---------------------------------------------------
def shard_range(n_traj: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous trajectory range [lo, hi) of `rank`: floor split, remainder to the low ranks."""
    base, rem = divmod(n_traj, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def ensemble_statistics(u: np.ndarray, group=None):
    """Ensemble mean and covariance over ALL ranks of per-trajectory values `u`: shape (n, d) (e.g. the final solution
    values) or (n, m, d) (the dense output on a common `saveat` grid of m points, `sol.saveat_u`).

    The only collective of the whole path: one all-reduce (sum) of `1 + m (d + d*d)` doubles (count, first and second
    moments per grid point).  With `torch.distributed` initialised on the `nccl` backend the moments are reduced on the
    GPUs over NVLink/NVSwitch; with `gloo` (CPU tests) on the host; without an initialised process group it is the local
    result.  Returns (n_total, mean, cov) with mean (d,) / (m, d) and cov (d, d) / (m, d, d)."""
    import torch
    import torch.distributed as dist

    u = np.ascontiguousarray(u, dtype=np.float64)
    grid = u.ndim == 3
    if not grid:
        u = u[:, None, :]
    n_loc, m, d = u.shape
    s1 = u.sum(axis=0)                                   # (m, d)
    s2 = np.einsum("nma,nmb->mab", u, u)                 # (m, d, d)
    mom = np.concatenate([[float(n_loc)], s1.ravel(), s2.ravel()])
    if dist.is_available() and dist.is_initialized():
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        t = torch.from_numpy(mom).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        mom = t.cpu().numpy()
    n = mom[0]
    mean = mom[1:1 + m * d].reshape(m, d) / n
    cov = mom[1 + m * d:].reshape(m, d, d) / n - np.einsum("ma,mb->mab", mean, mean)
    if not grid:
        mean, cov = mean[0], cov[0]
    return int(n), mean, cov
