"""Library-level multi-GPU: `EnsembleB200(devices=...)` / `pnode_solve_ensemble_sharded` (SURVEY.md section 8b `n_gpus`, 8e:
one host thread + solver + stream per device, contiguous trajectory ranges, results concatenated in CSR order, no exchange
on the step path).  The sharded call is also exercised with two solvers on ONE device, so the threading path is covered on
a single-GPU box; the multi-device test proper skips below two devices."""
import numpy as np
import pytest

import probnumdiffeq_jl_b200  # noqa: F401
from probnumdiffeq_jl_b200 import api, lib, problems

from test_gpu_parity import _inputs

pytestmark = pytest.mark.gpu

FIELDS = (lib.F_T_END, lib.F_U_FINAL, lib.F_U_FINAL_COV, lib.F_LOGLIK, lib.F_NACCEPT, lib.F_NREJECT, lib.F_NF, lib.F_RETCODE)
SAVED = (lib.F_T, lib.F_U, lib.F_U_COV, lib.F_DIFFUSIONS)


def _run_sharded(devices, u0, p, **kw):
    pd = problems.PROBLEMS["lotka_volterra"]
    solvers = [lib.Solver(lib.make_config(d=pd.d, q=3, n_p=pd.n_p, rhs_name="builtin:lotka_volterra", t0=0.0, t1=10.0, device=dev, **kw))
               for dev in devices]
    res = lib.solve_sharded(solvers, u0, p) if len(solvers) > 1 else [solvers[0].solve(u0, p)]
    out = {f: np.concatenate([r.field(f) for r in res]) for f in FIELDS}
    if kw.get("save_everystep"):
        out.update({f: np.concatenate([r.field(f) for r in res]) for f in SAVED})
        counts = np.concatenate([np.diff(r.field(lib.F_STEP_OFFSETS)) for r in res])
        out["off"] = np.concatenate([[0], np.cumsum(counts)])
    for r in res:
        r.free()
    for s in solvers:
        s.close()
    return out


@pytest.mark.parametrize("kw", [dict(), dict(save_everystep=True, smooth=True)])
def test_two_shards_on_one_device_equal_the_single_solve(kw):
    n = 1501   # odd: uneven shards (751 + 750)
    pd, u0, p = _inputs("lotka_volterra", n, 201, du0=0.2, dp=0.2)
    one = _run_sharded([0], u0, p, **kw)
    two = _run_sharded([0, 0], u0, p, **kw)
    for f, v in one.items():
        assert np.array_equal(v, two[f]), f
    assert lib.shard_range(n, 0, 2) == (0, 751) and lib.shard_range(n, 1, 2) == (751, 1501)


def test_sharded_argument_errors():
    pd, u0, p = _inputs("lotka_volterra", 4, 202)
    s = lib.Solver(lib.make_config(d=2, q=3, n_p=4, rhs_name="builtin:lotka_volterra", t0=0.0, t1=1.0))
    with pytest.raises(lib.PnodeError, match="one solver per device"):
        lib.solve_sharded([s, s], u0, p)
    s2 = lib.Solver(lib.make_config(d=2, q=3, n_p=4, rhs_name="builtin:lotka_volterra", t0=0.0, t1=1.0))
    with pytest.raises(lib.PnodeError, match="at least the number of devices"):
        lib.solve_sharded([s, s2], u0[:1], p[:1])
    s.close()
    s2.close()


def test_ensemble_b200_over_all_devices_is_bit_identical_to_one_device():
    ndev = lib.device_count()
    if ndev < 2:
        pytest.skip("needs at least two CUDA devices")
    prob = api.ODEProblem.from_library("lotka_volterra")
    rng = np.random.default_rng(203)
    pert = 1 + 0.1 * rng.uniform(-1, 1, (257, 2))
    ens = api.EnsembleProblem(prob, prob_func=lambda pr, i, rep: api.remake(pr, u0=list(np.array(pr.u0) * pert[i])))
    a = api.solve(ens, api.EK1(order=3, smooth=True), api.EnsembleB200(device=0), trajectories=257, saveat=[2.5, 7.5])
    b = api.solve(ens, api.EK1(order=3, smooth=True), api.EnsembleB200(devices="all"), trajectories=257, saveat=[2.5, 7.5])
    assert len(a) == len(b) == 257 and b.converged and len(b.info) == ndev
    for sa, sb in zip(a.u, b.u):
        assert np.array_equal(sa.t, sb.t) and np.array_equal(sa.u, sb.u) and np.array_equal(sa.pu_cov, sb.pu_cov)
        assert np.array_equal(sa.saveat_u, sb.saveat_u) and sa.stats == sb.stats
