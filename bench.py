#!/usr/bin/env python
"""bench.py -- ODE-filter trajectory-steps/sec (EK1, FP64) on N B200s, beside the CPU EnsembleThreads-style baseline.

Workload (BASELINE.json configs[4], the configuration the metric is quoted on): ROBER in ODE form,
EK1(order=3, smooth=false), adaptive (abstol 1e-6, reltol 1e-3), t in [0, 100], 2^20 trajectories PER GPU
(weak scaling; trajectories are independent, no collective on the step path), rate constants perturbed +-10 %,
PCG64 seed 3 (+rank).  A "step" of this benchmark is one ensemble solve (one persistent-kernel launch).

  python bench.py --gpus N --steps K --warmup W            # product arm (libpnode_cuda.so)
  python bench.py --impl reference ...                      # CPU arm: the oracle port of the reference algorithm
                                                            # on all host cores (Julia cannot run here)
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ODE-filter trajectory-steps/sec (EK1, FP64)"
UNIT = "accepted trajectory-steps/s"
WORKLOADS = {
    # name: (problem, q, t1, default trajectories per GPU)
    "rober": ("rober", 3, 100.0, 1 << 20),
    "orego": ("orego", 3, 30.0, 1 << 17),
    "vanderpol": ("vanderpol", 5, 2.0, 1 << 17),
    "lotka_volterra": ("lotka_volterra", 3, 10.0, 1 << 20),
    "fitzhugh_nagumo": ("fitzhugh_nagumo", 3, 20.0, 1 << 20),
}


def flops_per_step(D, d):
    """Algorithmic FLOPs of one accepted / one rejected step (dense operands, QR route, no
    credit for the structure the kernel exploits; f and J excluded)."""
    acc = (22.0 / 3.0) * D**3 + 8.0 * D * D * d + 6.0 * D * D + 6.0 * D * d * d + (2.0 / 3.0) * d**3
    rej = 5.0 * D * D + 2.0 * D * D * d + 2.0 * D * d * d + d**3 / 3.0
    return acc, rej


def make_inputs(problem, n, seed):
    from probnumdiffeq_jl_b200 import problems

    pd = problems.PROBLEMS[problem]
    rng = np.random.Generator(np.random.PCG64(seed))
    u0 = np.tile(np.array(pd.u0, dtype=np.float64), (n, 1))
    p = np.array(pd.p, dtype=np.float64)[None, :] * (1.0 + 0.1 * rng.uniform(-1.0, 1.0, (n, pd.n_p)))
    return pd, np.ascontiguousarray(u0), np.ascontiguousarray(p)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.rows = []

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for k, nm in enumerate(names) if any(len(r) > 3 + k and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "power_w_max": max((float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()), default=None),
                "samples": len(self.rows), "reasons": reasons}


def cpu_reference_run(problem, q, t1, n_sample, seed, threads=0):
    """The oracle (port of the reference algorithm: Gram+Cholesky with QR fallback, src/filtering/predict.jl:84-91) on the
    host cores, one trajectory per OpenMP task -- the stand-in for EnsembleThreads (Julia is not installable)."""
    from oracle import oracle as O
    from probnumdiffeq_jl_b200 import tracer

    pd, u0, p = make_inputs(problem, n_sample, seed)
    rhs = O.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), q)
    cfg = O.make_config(pd.d, q, pd.n_p, alg=O.EK1, smooth=False, adaptive=True, save_everystep=False, t0=0.0, t1=t1,
                        cov_backend=O.COV_REF)
    # torchrun exports OMP_NUM_THREADS=1 to every rank: ask for the host's cores explicitly (affinity-aware)
    nthr = threads or max(O.max_threads(), len(os.sched_getaffinity(0)))
    t = time.perf_counter()
    r = O.solve_ensemble(cfg, rhs, u0, p, n_threads=nthr)
    dt = time.perf_counter() - t
    return int(r["naccept"].sum()), dt, nthr, r


def run_reference(args):
    problem, q, t1, _ = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = args.cpu_sample
    # warm-up (also compiles the RHS) then K bounded samples
    for w in range(max(args.warmup, 1)):
        cpu_reference_run(problem, q, t1, max(64, n_sample // 16), 1000 + w)
    tot_steps, tot_t, nthr = 0, 0.0, 1
    for k in range(args.steps):
        s, dt, nthr, _ = cpu_reference_run(problem, q, t1, n_sample, 3 + k)
        tot_steps += s
        tot_t += dt
    val = tot_steps / tot_t
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{problem} EK1(order={q}) adaptive filter-only t1={t1}", "sample_trajectories_per_step": n_sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": nthr, "kind": "port",
                         "sample": f"{n_sample} trajectories per step (same distribution as the GPU ensemble), oracle port of the reference "
                                   f"algorithm (Gram+Cholesky predict, QR fallback), OpenMP {nthr} threads, gcc -O3 -march=native"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="rober", choices=sorted(WORKLOADS))
    ap.add_argument("--trajectories", type=int, default=0, help="trajectories per GPU (default: workload's BASELINE size)")
    ap.add_argument("--cpu-sample", type=int, default=65536, help="trajectories per CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lanes", type=int, default=0)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import probnumdiffeq_jl_b200  # noqa: F401
    from probnumdiffeq_jl_b200 import lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    problem, q, t1, n_default = WORKLOADS[args.workload]
    n = args.trajectories or n_default
    pd, u0, p = make_inputs(problem, n, 3 + rank)  # shard = this rank's trajectories; no data-path collective
    d, D = pd.d, pd.d * (q + 1)
    cfg = lib.make_config(d=d, q=q, n_p=pd.n_p, rhs_name="builtin:" + problem, adaptive=True, t0=0.0, t1=t1, device=local,
                          lanes_per_traj=args.lanes)
    solver = lib.Solver(cfg)

    # device-resident inputs (value) and pinned host inputs (e2e)
    du0 = torch.from_numpy(u0).cuda()
    dp = torch.from_numpy(p).cuda()
    hu0 = torch.from_numpy(u0).pin_memory()
    hp = torch.from_numpy(p).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        flush.zero_()
        torch.cuda.synchronize()
        r = solver.solve_device(n, du0.data_ptr(), dp.data_ptr())
        info = r.info
        out = (info.total_accepted, info.total_rejected, info.kernel_ms, info.n_launches)
        r.free()
        return out

    # pinned destinations of the results a caller reads (full-rate D2H, as the pinned inputs are for H2D)
    o_uf = torch.empty(n * d, dtype=torch.float64).pin_memory().numpy()
    o_cf = torch.empty(n * d * d, dtype=torch.float64).pin_memory().numpy()
    o_na = torch.empty(n, dtype=torch.int64).pin_memory().numpy()
    o_rc = torch.empty(n, dtype=torch.int32).pin_memory().numpy()

    def step_e2e():
        r = solver.solve(hu0.numpy(), hp.numpy())
        # the result a caller reads: final means, covariances and the per-trajectory stats / retcodes
        uf = r.field_into(lib.F_U_FINAL, o_uf)
        cf = r.field_into(lib.F_U_FINAL_COV, o_cf)
        na = r.field_into(lib.F_NACCEPT, o_na)
        rc = r.field_into(lib.F_RETCODE, o_rc)
        info = r.info
        nbytes = uf.nbytes + cf.nbytes + na.nbytes + rc.nbytes
        r.free()
        return int(na.sum()), info.h2d_bytes, nbytes + 16, bool((rc == 0).all())

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t_start = time.perf_counter()
    acc = rej = 0
    kms = 0.0
    launches = 0
    flush_s = 0.0
    for _ in range(args.steps):
        tf = time.perf_counter()
        a, r_, ms, nl = step_device()
        acc += a
        rej += r_
        kms += ms
        launches += nl
    barrier()
    wall = time.perf_counter() - t_start
    sampler.stop_flag.set()
    sampler.join()

    # e2e: same metric through the host-buffer call (H2D + kernel + D2H inside the timed region)
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    t_e = time.perf_counter()
    acc_e = 0
    h2d = d2h = 0
    ok = True
    for _ in range(args.steps):
        a, h2d, d2h, okk = step_e2e()
        acc_e += a
        ok = ok and okk
    barrier()
    wall_e = time.perf_counter() - t_e

    # reduce over ranks: MAX time, SUM units
    vals = torch.tensor([wall, wall_e, kms * 1e-3], dtype=torch.float64, device="cuda")
    sums = torch.tensor([acc, rej, acc_e, launches], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    wall_m, wall_em, ksec_m = vals.tolist()
    acc_t, rej_t, acc_et, launches_t = sums.tolist()

    if rank == 0:
        f_acc, f_rej = flops_per_step(D, d)
        # roofline of the dominant (only) kernel: algorithmic FLOPs per launch / CUDA-event duration of the launch
        fp64_peak = lib.measure_fp64_peak(local)
        flops_launch = (acc * f_acc + rej * f_rej) / args.steps
        dur = (kms * 1e-3) / args.steps
        achieved = flops_launch / dur / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        # per trajectory: u0, p (read by the init pre-pass), the pre-pass' initial state mu0[D] + dt0 (read by the step kernel),
        # final mean / covariance, five doubles and three counters + retcode
        alg_bytes = n * (8 * (d + pd.n_p) + 8 * (D + 1) + 8 * (d + d * d) + 8 * 5 + 3 * 8 + 4)
        line = {
            "metric": METRIC, "value": acc_t / wall_m, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall_m / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"BASELINE configs[4]: {problem} (ODE form) EK1(order={q}, smooth=false) adaptive abstol=1e-6 reltol=1e-3 "
                            f"t in [0,{t1}], filter only",
                "trajectories_per_gpu": n, "D": D, "d": d, "lanes_per_trajectory": int(solver.cfg.lanes_per_traj) or None,
                "parallelism": f"trajectory-sharded x{world}, no collective on the step path",
                "l2": "256 MiB buffer written between timed iterations (inputs+outputs are also re-staged per step)",
                "kernel_ms_per_step_device_events": 1e3 * ksec_m / args.steps,
                "accepted_steps_per_step": acc_t / args.steps, "rejected_steps_per_step": rej_t / args.steps,
                "all_retcodes_success": ok,
            },
            "e2e": {"value": acc_et / wall_em, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches_t),
            "roofline": {
                "bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, one ncu --set full capture of this very
                # command at its default size (profiles/r1r_pn_small_rober_summary.txt: 146.9 MB read + 123.3 MB written, the
                # write side includes the L2 write-back of the init pre-pass' buffers); null for other workloads / sizes
                "traffic": 270.2e6 if (args.workload == "rober" and n == (1 << 20)) else None,
                "peak_source": "measured live by pnode_measure_fp64_peak (DFMA probe); MEASURED_PEAKS.json has no FP64 figure",
                # for reference: 64 DFMA / clk / SM x 148 SMs x sm_max_mhz (the DFMA probe reaches ~91 % of it)
                "peak_nominal": 2.0 * 64 * 148 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12,
                "frac_of_nominal": achieved / (2.0 * 64 * 148 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12),
                "algorithmic_flops_per_accepted_step": f_acc, "algorithmic_flops_per_rejected_step": f_rej,
                "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / dur / 1e9, "peak_gbs": hbm_peak,
                        "frac": alg_bytes / dur / 1e9 / hbm_peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback"},
            },
            "clocks": sampler.summary(),
        }
        if not args.no_cpu_baseline:
            nsm = args.cpu_sample
            cpu_reference_run(problem, q, t1, 256, 999)  # warm-up / RHS compile
            s, dtc, nthr, _ = cpu_reference_run(problem, q, t1, nsm, 3)
            line["cpu_baseline"] = {
                "value": s / dtc, "unit": UNIT, "cores": nthr, "kind": "port",
                "sample": f"{nsm} trajectories of the same ensemble distribution ({s} accepted steps in {dtc:.1f} s), oracle port of the "
                          f"reference algorithm on {nthr} OpenMP threads (CPU restatement of EnsembleThreads, not Julia)"}
        print(json.dumps(line), flush=True)
    solver.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
