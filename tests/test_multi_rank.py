"""World-size-2 host logic on CPU (gloo): trajectory sharding and the optional ensemble-statistics all-reduce
(the only collective of the path).  The GPU solve itself is rank-local and is covered by -m gpu tests."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_traj, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import probnumdiffeq_jl_b200  # noqa: F401
    from probnumdiffeq_jl_b200 import api

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(123)
    u_all = rng.normal(size=(n_traj, 3))          # stands in for the per-trajectory final values
    lo, hi = api.shard_range(n_traj, rank, world)
    n, mean, cov = api.ensemble_statistics(u_all[lo:hi])
    g_all = rng.normal(size=(n_traj, 5, 2))       # ... and for the dense output on a common saveat grid of 5 points
    _, gmean, gcov = api.ensemble_statistics(g_all[lo:hi])
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), n=n, mean=mean, cov=cov, lo=lo, hi=hi, gmean=gmean, gcov=gcov)
    dist.destroy_process_group()


def test_shard_ranges_partition_the_ensemble():
    import probnumdiffeq_jl_b200  # noqa: F401
    from probnumdiffeq_jl_b200 import api

    for n in (1, 7, 8, 1000, 1 << 20, 10_000_001):
        for w in (1, 2, 4, 8):
            r = [api.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_ensemble_statistics_allreduce_gloo_world2(tmp_path):
    n_traj, world = 1001, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_traj, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(123)
    u_all = rng.normal(size=(n_traj, 3))
    res = [np.load(tmp_path / f"r{k}.npz") for k in range(world)]
    assert res[0]["hi"] == res[1]["lo"] and res[1]["hi"] == n_traj
    for r in res:
        assert int(r["n"]) == n_traj
        assert np.allclose(r["mean"], u_all.mean(axis=0), rtol=1e-12, atol=1e-14)
        assert np.allclose(r["cov"], np.cov(u_all.T, bias=True), rtol=1e-10, atol=1e-13)
    g_all = rng.normal(size=(n_traj, 5, 2))
    for r in res:
        assert np.allclose(r["gmean"], g_all.mean(axis=0), rtol=1e-12, atol=1e-14)
        for j in range(5):
            assert np.allclose(r["gcov"][j], np.cov(g_all[:, j].T, bias=True), rtol=1e-10, atol=1e-13)
