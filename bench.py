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
    # name: (problem, q, t1, default trajectories per GPU, algorithm, smooth, BASELINE.json config)
    "rober": ("rober", 3, 100.0, 1 << 20, "EK1", False, "configs[4]"),
    "orego": ("orego", 3, 30.0, 1 << 17, "EK1", False, "configs[4]"),
    "vanderpol": ("vanderpol", 5, 2.0, 100000, "EK1", True, "configs[2]"),          # cfg3: filter + smoother, every step saved
    "lotka_volterra": ("lotka_volterra", 3, 10.0, 1 << 20, "EK0", False, "configs[1]"),  # cfg2: EK0 on the Kronecker layout
    "pleiades": ("pleiades", 3, 3.0, 1 << 14, "EK1", False, "configs[3]"),          # cfg4: D = 112, one CTA per trajectory
}


def flops_per_step(D, d, alg="EK1", smooth=False, n=None):
    """Algorithmic FLOPs of one accepted / one rejected step (dense operands, QR route, no
    credit for the structure the kernel exploits; f and J excluded).  EK0 on the Kronecker layout: the same formula on the
    n x n factors with a scalar S.  Smoothing adds F_bk + F_sm per accepted step."""
    if alg == "EK0":
        acc = (22.0 / 3.0) * n**3 + 4.0 * n * n * d + 6.0 * n * n
        rej = 5.0 * n * n + 2.0 * n * n * d
        return acc, rej
    acc = (22.0 / 3.0) * D**3 + 8.0 * D * D * d + 6.0 * D * D + 6.0 * D * d * d + (2.0 / 3.0) * d**3
    rej = 5.0 * D * D + 2.0 * D * D * d + 2.0 * D * d * d + d**3 / 3.0
    if smooth:
        acc += 10.0 * D**3 + (22.0 / 3.0) * D**3 + 2.0 * D * D * d + 2.0 * D * D
    return acc, rej


def committed_ncu(workload, n):
    """Numbers that only a profiler can give -- DRAM traffic of the dominant kernel, its FP64-pipe utilisation and executed FP64
    instruction count -- read from the committed ncu capture of THIS command at THIS size (profiles/ncu_captures.json, written by
    scripts/ncu_summary.py from the .ncu-rep); None when no capture exists for the workload / size."""
    try:
        tab = json.load(open(os.path.join(ROOT, "profiles", "ncu_captures.json")))
    except Exception:
        return None
    return tab.get(f"{workload}:{n}")


def executed_flops_per_accepted_step(workload, kernel_prefix=""):
    """Executed FP64 operations (2 DFMA + DMUL + DADD, predicated-on thread instructions counted by ncu) per accepted step of the
    dominant kernel, from the committed capture of this workload that carries its own accepted-step count (any ensemble size: the
    per-step count does not depend on it).  Rejected attempts and the identity steps of finished groups are inside the count."""
    try:
        tab = json.load(open(os.path.join(ROOT, "profiles", "ncu_captures.json")))
    except Exception:
        return None, None
    best = None
    for key, e in tab.items():
        if key.split(":")[0] == workload and e.get("accepted_steps") and kernel_prefix in e.get("kernel", ""):
            if best is None or e["accepted_steps"] > best[1]["accepted_steps"]:
                best = (key, e)
    if best is None:
        return None, None
    return best[1]["executed"]["fp64_flops"] / best[1]["accepted_steps"], best[0]


def make_inputs(problem, n, seed):
    from probnumdiffeq_jl_b200 import problems

    pd = problems.PROBLEMS[problem]
    rng = np.random.Generator(np.random.PCG64(seed))
    u0 = np.tile(np.array(pd.u0, dtype=np.float64), (n, 1))
    p = np.array(pd.p, dtype=np.float64)[None, :] * (1.0 + 0.1 * rng.uniform(-1.0, 1.0, (n, pd.n_p)))
    if problem == "pleiades":      # initial positions perturbed by 1e-3
        u0[:, 14:] += 1e-3 * rng.uniform(-1.0, 1.0, (n, 14))
    elif problem == "vanderpol":   # cfg3: mu_i = 1e3 (1 + 0.2 xi), u0_i = (2 + 0.1 xi, 0)
        p = np.array(pd.p, dtype=np.float64)[None, :] * (1.0 + 0.2 * rng.uniform(-1.0, 1.0, (n, pd.n_p)))
        u0[:, 0] += 0.1 * rng.uniform(-1.0, 1.0, n)
    elif problem == "lotka_volterra":  # cfg2: u0_i = (1,1)(1 + 0.1 xi)
        u0 = u0 * (1.0 + 0.1 * rng.uniform(-1.0, 1.0, (n, pd.d)))
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


def cpu_reference_run(problem, q, t1, n_sample, seed, threads=0, alg="EK1", smooth=False):
    """The oracle (port of the reference algorithm: Gram+Cholesky with QR fallback, src/filtering/predict.jl:84-91) on the
    host cores, one trajectory per OpenMP task -- the stand-in for EnsembleThreads (Julia is not installable)."""
    from oracle import oracle as O
    from probnumdiffeq_jl_b200 import tracer

    pd, u0, p = make_inputs(problem, n_sample, seed)
    rhs = O.compile_rhs(tracer.trace(pd.f, pd.d, pd.n_p), q)
    cfg = O.make_config(pd.d, q, pd.n_p, alg=O.EK0 if alg == "EK0" else O.EK1, smooth=smooth, adaptive=True, save_everystep=smooth,
                        t0=0.0, t1=t1, cov_backend=O.COV_REF)
    # torchrun exports OMP_NUM_THREADS=1 to every rank: ask for the host's cores explicitly (affinity-aware)
    nthr = threads or max(O.max_threads(), len(os.sched_getaffinity(0)))
    t = time.perf_counter()
    r = O.solve_ensemble(cfg, rhs, u0, p, n_threads=nthr)
    dt = time.perf_counter() - t
    return int(r["naccept"].sum()), dt, nthr, r


def run_reference(args):
    problem, q, t1, _, alg, smooth, cfgname = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = args.cpu_sample
    if problem == "pleiades":
        n_sample = min(n_sample, 32)   # D = 112: ~0.2 s per trajectory and core
    # warm-up (also compiles the RHS) then K bounded samples
    for w in range(max(args.warmup, 1)):
        cpu_reference_run(problem, q, t1, max(16, n_sample // 16), 1000 + w, alg=alg, smooth=smooth)
    tot_steps, tot_t, nthr = 0, 0.0, 1
    for k in range(args.steps):
        s, dt, nthr, _ = cpu_reference_run(problem, q, t1, n_sample, 3 + k, alg=alg, smooth=smooth)
        tot_steps += s
        tot_t += dt
    val = tot_steps / tot_t
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload), "sample_trajectories_per_step": n_sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": nthr, "kind": "port",
                         "sample": f"{n_sample} trajectories per step (same distribution as the GPU ensemble), oracle port of the reference "
                                   f"algorithm (Gram+Cholesky predict, QR fallback), OpenMP {nthr} threads, gcc -O3 -march=native"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_name(w):
    problem, q, t1, _, alg, smooth, cfgname = WORKLOADS[w]
    form = " (ODE form)" if problem == "rober" else ""
    return (f"BASELINE {cfgname}: {problem}{form} {alg}(order={q}, smooth={'true' if smooth else 'false'}) adaptive abstol=1e-6 "
            f"reltol=1e-3 t in [0,{t1}], {'filter + smoother, every step saved' if smooth else 'filter only'}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="rober", choices=sorted(WORKLOADS))
    ap.add_argument("--trajectories", type=int, default=0, help="trajectories per GPU (default: workload's BASELINE size)")
    ap.add_argument("--total-trajectories", type=int, default=0,
                    help="strong scaling: this many trajectories in total, split over the ranks (BASELINE configs[4] sweep 10k ... 10M)")
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

    problem, q, t1, n_default, alg, smooth, cfgname = WORKLOADS[args.workload]
    strong = args.total_trajectories > 0
    if strong:   # contiguous trajectory ranges per rank (floor split, remainder to the low ranks), as pnode_shard_range
        lo, hi = lib.shard_range(args.total_trajectories, rank, world)
        n = hi - lo
    else:
        n = args.trajectories or n_default
    pd, u0, p = make_inputs(problem, n, 3 + rank)  # shard = this rank's trajectories; no data-path collective
    d, D = pd.d, pd.d * (q + 1)
    cfg = lib.make_config(d=d, q=q, n_p=pd.n_p, rhs_name="builtin:" + problem, adaptive=True, t0=0.0, t1=t1, device=local,
                          lanes_per_traj=args.lanes, alg=lib.EK0 if alg == "EK0" else lib.EK1, smooth=smooth, save_everystep=smooth)
    solver = lib.Solver(cfg)

    # device-resident inputs (value) and pinned host inputs (e2e)
    du0 = torch.from_numpy(u0).cuda()
    dp = torch.from_numpy(p).cuda() if p.size else torch.zeros(1, dtype=torch.float64, device="cuda")
    hu0 = torch.from_numpy(u0).pin_memory()
    hp = torch.from_numpy(p).pin_memory() if p.size else torch.from_numpy(p)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    lanes_used = [0]   # lanes per trajectory of the step kernel that ran (1: pn_thread, 2 ... 32: pn_small, 256: pn_cta)

    def step_device():
        flush.zero_()
        torch.cuda.synchronize()
        r = solver.solve_device(n, du0.data_ptr(), dp.data_ptr())
        info = r.info
        out = (info.total_accepted, info.total_rejected, info.kernel_ms, info.n_launches)
        lanes_used[0] = int(info.lanes_per_traj)
        r.free()
        return out

    # pinned destinations of the results a caller reads (full-rate D2H, as the pinned inputs are for H2D)
    o_uf = torch.empty(n * d, dtype=torch.float64).pin_memory().numpy()
    o_cf = torch.empty(n * d * d, dtype=torch.float64).pin_memory().numpy()
    o_na = torch.empty(n, dtype=torch.int64).pin_memory().numpy()
    o_rc = torch.empty(n, dtype=torch.int32).pin_memory().numpy()
    o_saved = {}

    def step_e2e():
        r = solver.solve(hu0.numpy(), hp.numpy())
        # the result a caller reads: final means, covariances and the per-trajectory stats / retcodes
        uf = r.field_into(lib.F_U_FINAL, o_uf)
        cf = r.field_into(lib.F_U_FINAL_COV, o_cf)
        na = r.field_into(lib.F_NACCEPT, o_na)
        rc = r.field_into(lib.F_RETCODE, o_rc)
        info = r.info
        nbytes = uf.nbytes + cf.nbytes + na.nbytes + rc.nbytes
        if smooth:   # the smoothed solution itself: sol.t, sol.u, sol.pu of every saved point (CSR over trajectories)
            for f, w, dt in ((lib.F_STEP_OFFSETS, 0, np.int64), (lib.F_T, 1, np.float64), (lib.F_U, d, np.float64), (lib.F_U_COV, d * d, np.float64)):
                cnt = (n + 1) if f == lib.F_STEP_OFFSETS else info.total_saved * w
                if f not in o_saved or o_saved[f].size < cnt:
                    o_saved[f] = torch.empty(int(cnt * 1.1) + 16, dtype=torch.int64 if dt is np.int64 else torch.float64).pin_memory().numpy()
                nbytes += r.field_into(f, o_saved[f][:cnt]).nbytes
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
    for _ in range(args.steps):
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
    sums = torch.tensor([acc, rej, acc_e, launches, n], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    wall_m, wall_em, ksec_m = vals.tolist()
    acc_t, rej_t, acc_et, launches_t, n_t = sums.tolist()

    if rank == 0:
        f_acc, f_rej = flops_per_step(D, d, alg, smooth, q + 1)
        # roofline of the dominant kernel: algorithmic FLOPs per launch / CUDA-event duration of the launch(es) of one solve.
        # The denominator is a DFMA-saturating probe run here, with the SM clock it ran at; without a driver-measured FP64
        # peak the nominal 64 DFMA / clk / SM x 148 SMs x max clock is reported beside it (frac_of_nominal).
        probe_clk = ClockSampler(local)
        probe_clk.start()
        fp64_peak = max(lib.measure_fp64_peak(local) for _ in range(3))
        probe_clk.stop_flag.set()
        probe_clk.join()
        flops_launch = (acc * f_acc + rej * f_rej) / args.steps
        dur = (kms * 1e-3) / args.steps
        achieved = flops_launch / dur / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        nominal = 2.0 * 64 * 148 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12
        # per trajectory: u0, p (read by the init pre-pass), the pre-pass' initial state mu0[D] + dt0 (read by the step kernel),
        # final mean / covariance, five doubles and three counters + retcode; smoothing: write + read of the filter records
        # (D^2 + D + 2 doubles per saved point) and the saved sol.t / sol.u / sol.pu
        alg_bytes = n * (8 * (d + pd.n_p) + 8 * (D + 1) + 8 * (d + d * d) + 8 * 5 + 3 * 8 + 4)
        if smooth:
            alg_bytes += (acc / args.steps) * 8 * (2 * (D * D + D + 2) + 1 + d + d * d)
        kprefix = ("pn_kron_kernel" if alg == "EK0" else
                   {1: "pn_thread_kernel", 256: "pn_cta_kernel"}.get(lanes_used[0], "pn_small_kernel"))
        cap = committed_ncu(args.workload, n)
        if cap and kprefix not in cap.get("kernel", ""):
            cap = None   # the committed capture is of another step kernel than the one that ran
        ex_per_step, ex_key = executed_flops_per_accepted_step(args.workload, kprefix)
        executed_tflops = (acc / args.steps) * ex_per_step / dur / 1e12 if ex_per_step else None
        line = {
            "metric": METRIC, "value": acc_t / wall_m, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall_m / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": workload_name(args.workload),
                "trajectories_per_gpu": n, "trajectories_total": int(n_t), "D": D, "d": d,
                "lanes_per_trajectory": lanes_used[0],
                "step_kernel": ("pn_kron_kernel (one thread per trajectory)" if alg == "EK0" else
                                {1: "pn_thread_kernel (one thread per trajectory)", 256: "pn_cta_kernel (one CTA per trajectory)"}.get(
                                    lanes_used[0], f"pn_small_kernel ({lanes_used[0]} lanes per trajectory)")),
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
                # dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu --set full capture of
                # this command at this size (profiles/ncu_captures.json); null when there is none
                "traffic": cap.get("dram_bytes") if cap else None,
                "peak_source": "measured here by pnode_measure_fp64_peak (DFMA probe, best of 3); MEASURED_PEAKS.json has no FP64 figure",
                "peak_probe_sm_mhz": probe_clk.summary().get("sm_mhz"),
                # 64 DFMA / clk / SM x 148 SMs x sm_max_mhz: the figure to read when no driver-measured FP64 peak exists
                "peak_nominal": nominal, "frac_of_nominal": achieved / nominal,
                "algorithmic_flops_per_accepted_step": f_acc, "algorithmic_flops_per_rejected_step": f_rej,
                # the algorithmic count is SURVEY.md section 8d's (dense operands, Householder route); the kernel follows the
                # reference's own Gram + Cholesky route and uses the Kronecker / triangular structure, so it EXECUTES fewer FP64
                # operations than that: the executed count and the FP64-pipe utilisation below come from the ncu capture
                "executed": cap.get("executed") if cap else None,
                # what the FP64 pipe really does: executed FP64 operations per accepted step (ncu capture `executed_from`) x the
                # accepted steps of THIS run / the CUDA-event duration of THIS run.  `frac` above can exceed 1 because its numerator
                # credits the dense-operand count of SURVEY.md section 8d (17 658 FLOP per step at D = 12) while the kernel uses the
                # structure of Ah, Qh and H; `frac_executed` is the pipe-level figure and is the one to judge the kernel by.
                "executed_flops_per_accepted_step": ex_per_step, "executed_from": ex_key,
                "executed_tflops": executed_tflops, "frac_executed": executed_tflops / fp64_peak if executed_tflops else None,
                "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / dur / 1e9, "peak_gbs": hbm_peak,
                        "frac": alg_bytes / dur / 1e9 / hbm_peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback"},
            },
            "clocks": sampler.summary(),
        }
        if not args.no_cpu_baseline:
            nsm = 32 if problem == "pleiades" else args.cpu_sample
            cpu_reference_run(problem, q, t1, 16 if problem == "pleiades" else 256, 999, alg=alg, smooth=smooth)  # warm-up / RHS compile
            s, dtc, nthr, _ = cpu_reference_run(problem, q, t1, nsm, 3, alg=alg, smooth=smooth)
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
