"""Throughput of every BASELINE.json configuration on one GPU (device-event kernel time, inputs resident in HBM is not
attempted here: this goes through the host-buffer C-ABI call and reports both the kernel time and the wall time).
Writes one JSON line per configuration; used to fill the table in DESIGN.md / profiles/."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import probnumdiffeq_jl_b200  # noqa: F401
from probnumdiffeq_jl_b200 import lib, problems

PEAK = lib.measure_fp64_peak(0)


def F_filt(D, d):
    return (22 / 3) * D**3 + 8 * D * D * d + 6 * D * D + 6 * D * d * d + (2 / 3) * d**3


def F_smooth(D, d):
    return 10 * D**3 + (22 / 3) * D**3 + 2 * D * D * d + 2 * D * D


def run(tag, name, n, q, t1, alg=lib.EK1, seed=0, reps=2, du0=0.0, dp=0.1, second=False, **kw):
    pd = problems.PROBLEMS.get(name) or problems.MASS_MATRIX_PROBLEMS[name]
    rng = np.random.Generator(np.random.PCG64(seed))
    u0 = np.array(pd.u0, dtype=float)[None, :] * (1 + du0 * rng.uniform(-1, 1, (n, pd.d)))
    p = np.array(pd.p, dtype=float)[None, :] * (1 + dp * rng.uniform(-1, 1, (n, pd.n_p))) if pd.n_p else np.zeros((n, 0))
    if name == "pleiades":
        u0[:, 14:] += 1e-3 * rng.uniform(-1, 1, (n, 14))
    if pd.mass_matrix is not None:   # mass-matrix problems are traced and NVRTC-compiled (M baked into the kernel)
        from probnumdiffeq_jl_b200 import tracer
        src = tracer.cuda_source(tracer.trace(pd.f, pd.d, pd.n_p), "UserProb")
        cfg = lib.make_config(d=pd.d, q=q, n_p=pd.n_p, rhs_name="UserProb", rhs_src=src, alg=alg, t0=0.0, t1=t1,
                              mass_matrix=pd.mass_matrix, **kw)
    else:
        cfg = lib.make_config(d=pd.d // 2 if second else pd.d, second_order=second, q=q, n_p=pd.n_p, rhs_name="builtin:" + name, alg=alg,
                              t0=0.0, t1=t1, **kw)
    s = lib.Solver(cfg)
    best = None
    for _ in range(reps + (1 if pd.mass_matrix is not None else 0)):   # first call of an NVRTC solver compiles
        t = time.perf_counter()
        r = s.solve(np.ascontiguousarray(u0), np.ascontiguousarray(p))
        wall = time.perf_counter() - t
        info = r.info
        rc = r.field(lib.F_RETCODE)
        rec = dict(kernel_ms=info.kernel_ms, wall_ms=wall * 1e3, accepted=int(info.total_accepted), rejected=int(info.total_rejected),
                   launches=int(info.n_launches), ok=bool((rc == 0).all()))
        r.free()
        if best is None or rec["kernel_ms"] < best["kernel_ms"]:
            best = rec
    s.close()
    dm = pd.d // 2 if second else pd.d
    D = dm * (q + 1)
    fl = F_filt(D, dm) + (F_smooth(D, pd.d) if kw.get("smooth") else 0.0)
    if alg == lib.EK0:
        nn = q + 1
        fl = (22 / 3) * nn**3 + 4 * nn * nn * pd.d
    passes = 2 if kw.get("save_everystep") else 1   # the counting pass repeats the filter arithmetic
    best.update(tag=tag, problem=name, n_traj=n, D=D, d=pd.d, steps_per_s_kernel=best["accepted"] / (best["kernel_ms"] * 1e-3),
                steps_per_s_wall=best["accepted"] / (best["wall_ms"] * 1e-3),
                algorithmic_tflops=best["accepted"] * fl / (best["kernel_ms"] * 1e-3) / 1e12, fp64_peak_tflops=PEAK,
                note=f"{passes} filter pass(es) inside kernel_ms" if passes > 1 else "")
    best["roofline_frac"] = best["algorithmic_tflops"] / PEAK
    print(json.dumps(best), flush=True)


which = sys.argv[1:] or ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"]
if "cfg1" in which:
    run("cfg1 FHN EK1(3) dt=0.01 filter+smoother, 1 trajectory", "fitzhugh_nagumo", 1, 3, 20.0, adaptive=False, dt=0.01, smooth=True, save_everystep=True)
    run("cfg1 shape, 65536 trajectories", "fitzhugh_nagumo", 65536, 3, 20.0, adaptive=False, dt=0.01, smooth=True, save_everystep=True, dp=0.05)
if "cfg2" in which:
    run("cfg2 LV EK0(3) adaptive filter only, 1M", "lotka_volterra", 1 << 20, 3, 10.0, alg=lib.EK0, du0=0.1)
if "cfg3" in which:
    run("cfg3 VdP mu=1e3 EK1(5) adaptive filter+smoother, 100k", "vanderpol", 100000, 5, 2.0, smooth=True, save_everystep=True, seed=1, dp=0.2)
    run("cfg3 filter only", "vanderpol", 100000, 5, 2.0, seed=1, dp=0.2)
if "cfg4" in which:
    run("cfg4 Pleiades EK1(3) D=112 filter only, 16k", "pleiades", 16384, 3, 3.0, seed=2)
    run("cfg4 Pleiades as 14 second-order equations (benchmarks/pleiades.jmd prob2), EK1(3) D=56 filter only, 16k", "pleiades", 16384, 3, 3.0,
        seed=2, second=True)
    # filter + smoother at D = 112: 100 KB of filter record per step -- 512 trajectories x ~770 record slots = 40 GB on one GPU
    # (16 384 trajectories need ~1 TB: trajectory-sharded over 8 GPUs by storage, in chunks)
    run("cfg4 Pleiades EK1(3) D=112 filter+smoother, 512 trajectories", "pleiades", 512, 3, 3.0, seed=2, smooth=True, save_everystep=True, reps=1)
if "cfg5" in which:
    for n in (10_000, 100_000, 1_000_000, 10_000_000):
        run(f"cfg5 ROBER EK1(3) adaptive filter only, {n}", "rober", n, 3, 100.0, seed=3)
    run("cfg5 OREGO EK1(3) adaptive filter only, 131072", "orego", 131072, 3, 30.0, seed=3)
if "cfg5dae" in which:   # ROBER as the reference benchmarks it: mass-matrix DAE form (benchmarks/rober.jmd:37-46)
    for n in (100_000, 1_000_000):
        run(f"cfg5 ROBER-DAE (M = diag(1,1,0)) EK1(3) adaptive filter only, {n}", "rober_dae", n, 3, 100.0, seed=3)
if "ioup" in which:   # IOUP priors: scalar rate on the register kernels, matrix / Jacobian-updated rate on the general-prior path (pn_gen.cuh)
    run("IOUP(3, -1) scalar rate, LV EK1 adaptive filter only, 262144 (register kernels)", "lotka_volterra", 262144, 3, 10.0, du0=0.1,
        prior=lib.PRIOR_IOUP, rate_parameter=-1.0)
    run("IOUP(3, 2x2 matrix rate), LV EK1 adaptive filter only, 65536 (general-prior path, warp per trajectory)", "lotka_volterra", 65536, 3, 10.0,
        du0=0.1, prior=lib.PRIOR_IOUP, rate_parameter=np.array([[-1.0, 0.3], [-0.2, -0.5]]))
    run("RosenbrockExpEK (IOUP(3, update_rate_parameter=true)), LV EK1 adaptive filter only, 65536", "lotka_volterra", 65536, 3, 10.0, du0=0.1,
        prior=lib.PRIOR_IOUP, rate_parameter=None, update_rate_parameter=True)
if "fenrir" in which:   # Fenrir data log-likelihood of a parameter sweep against one data set (src/data_likelihoods/fenrir.jl)
    from probnumdiffeq_jl_b200 import tracer as _tr
    pd = problems.PROBLEMS["lotka_volterra"]
    for n in (16_384, 131_072):
        rng = np.random.Generator(np.random.PCG64(7))
        u0 = np.tile(np.array(pd.u0, dtype=float), (n, 1))
        p = np.array(pd.p, dtype=float)[None, :] * (1 + 0.1 * rng.uniform(-1, 1, (n, pd.n_p)))
        data_t = np.linspace(0.5, 10.0, 20)
        data_u = 1.0 + rng.uniform(0, 2, (20, 2))
        src = _tr.cuda_source(_tr.trace(pd.f, 2, 4), "UserProb")
        cfg = lib.make_config(d=2, q=3, n_p=4, rhs_name="UserProb", rhs_src=src, smooth=True, save_everystep=True, t0=0.0, t1=10.0,
                              data_t=data_t, data_u=data_u, observation_noise_cov=0.25)
        s = lib.Solver(cfg)
        best = None
        for _ in range(3):
            t = time.perf_counter()
            r = s.solve(np.ascontiguousarray(u0), np.ascontiguousarray(p))
            wall = time.perf_counter() - t
            info = r.info
            ll = r.field(lib.F_FENRIR_LOGLIK)
            rec = dict(kernel_ms=info.kernel_ms, wall_ms=wall * 1e3, accepted=int(info.total_accepted), launches=int(info.n_launches),
                       ok=bool(np.isfinite(ll).all()))
            r.free()
            if best is None or rec["kernel_ms"] < best["kernel_ms"]:
                best = rec
        s.close()
        best.update(tag=f"Fenrir log-likelihood, LV EK1(3) adaptive, 20 observations, {n} parameter sets", n_traj=n,
                    loglik_per_s_kernel=n / (best["kernel_ms"] * 1e-3), steps_per_s_kernel=best["accepted"] / (best["kernel_ms"] * 1e-3))
        print(json.dumps(best), flush=True)
