"""BASELINE configs[4] sweep: ROBER EK1(3) adaptive, filter only, 10k ... 10M trajectories IN TOTAL (strong scaling), on the
ranks of one torchrun launch (or one process).  One process group, one solver per rank, all sizes in one run:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/sweep_cfg5.py

Per size: the trajectories are split into contiguous ranges (pnode_shard_range), inputs are resident in HBM, 1 warm-up and
`--steps` timed solves bracketed by barrier + synchronize, MAX time over ranks, SUM of accepted steps; rank 0 prints one JSON line
per size (profiles/r2*_cfg5_sweep_Ngpu.jsonl).  SURVEY.md section 8d: the 10M point runs ROBER to T = 100 like the others."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", type=str, default="10000,100000,1000000,10000000")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--problem", type=str, default="rober")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import probnumdiffeq_jl_b200  # noqa: F401
    from probnumdiffeq_jl_b200 import lib, problems

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pd = problems.PROBLEMS[args.problem]
    t1 = {"rober": 100.0, "orego": 30.0}[args.problem]
    cfg = lib.make_config(d=pd.d, q=3, n_p=pd.n_p, rhs_name="builtin:" + args.problem, adaptive=True, t0=0.0, t1=t1, device=local)
    solver = lib.Solver(cfg)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for total in [int(x) for x in args.sizes.split(",")]:
        lo, hi = lib.shard_range(total, rank, world)
        n = hi - lo
        rng = np.random.Generator(np.random.PCG64(3))
        # the same ensemble at every world size: parameters drawn for the whole ensemble, this rank takes its range
        pall = np.array(pd.p, dtype=np.float64)[None, :] * (1.0 + 0.1 * rng.uniform(-1.0, 1.0, (total, pd.n_p)))
        p = np.ascontiguousarray(pall[lo:hi])
        del pall
        u0 = np.tile(np.array(pd.u0, dtype=np.float64), (max(n, 1), 1))[:n]
        du0, dp = torch.from_numpy(u0).cuda(), torch.from_numpy(p).cuda()

        def one():
            flush.zero_()
            torch.cuda.synchronize()
            r = solver.solve_device(n, du0.data_ptr(), dp.data_ptr())
            info = r.info
            out = (info.total_accepted, info.kernel_ms)
            r.free()
            return out

        if n > 0:
            one()
        barrier()
        t0 = time.perf_counter()
        acc, kms = 0, 0.0
        for _ in range(args.steps):
            if n > 0:
                a, ms = one()
                acc += a
                kms += ms
        barrier()
        wall = time.perf_counter() - t0
        v = torch.tensor([wall, kms * 1e-3], dtype=torch.float64, device="cuda")
        sm = torch.tensor([float(acc)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        if rank == 0:
            wall_m, k_m = v.tolist()
            print(json.dumps({"workload": f"cfg5 {args.problem} EK1(3) adaptive filter only, strong scaling", "trajectories_total": total,
                              "n_gpus": world, "steps": args.steps, "accepted_per_solve": sm.item() / args.steps,
                              "steps_per_s_wall": sm.item() / wall_m, "steps_per_s_kernel": sm.item() / k_m if k_m > 0 else None,
                              "ms_per_solve_wall": 1e3 * wall_m / args.steps, "ms_per_solve_kernel_max_rank": 1e3 * k_m / args.steps,
                              "timing": "barrier + synchronize both sides, max over ranks; 256 MiB L2 flush before every solve (inside the wall time)"}),
                  flush=True)
    solver.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
