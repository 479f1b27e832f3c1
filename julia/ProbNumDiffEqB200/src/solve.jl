# `solve(EnsembleProblem(prob; prob_func), EK1(...), EnsembleB200(); trajectories = N, kwargs...)`

"""
    EnsembleB200(; devices = 0, lanes_per_traj = 0)

Ensemble algorithm: all trajectories in one persistent-kernel launch per B200.  `devices` is a device ordinal, a collection of
ordinals, or `:all`; with several devices the ensemble is split into contiguous trajectory ranges (`pnode_shard_range`), one
`pnode_solver` per device, one host thread and stream per device inside `pnode_solve_ensemble_sharded`, no exchange between the
shards (SURVEY.md section 8e).
"""
struct EnsembleB200 <: SciMLBase.EnsembleAlgorithm
    devices::Vector{Int32}
    lanes_per_traj::Int32
end
function EnsembleB200(; devices=0, lanes_per_traj=0)
    devs = devices === :all ? collect(Int32, 0:max(device_count() - 1, 0)) : collect(Int32, devices)
    isempty(devs) && throw(ArgumentError("EnsembleB200 needs at least one device"))
    allunique(devs) || throw(ArgumentError("devices must be distinct"))
    return EnsembleB200(devs, Int32(lanes_per_traj))
end

alg_code(alg) = alg isa ProbNumDiffEq.EK0 ? PNODE_EK0 : alg isa ProbNumDiffEq.DiagonalEK1 ? PNODE_DIAGONAL_EK1 : PNODE_EK1
prior_code(p) = p isa ProbNumDiffEq.IOUP ? PNODE_PRIOR_IOUP : PNODE_PRIOR_IWP
"""
`(scalar rate, row-major d x d rate matrix or nothing, update flag)` of the prior (src/priors/ioup.jl:33-101): a Number stays a
number, a vector becomes `Diagonal(r)` and a matrix is passed as is -- exactly what `to_sde` builds --, and
`IOUP(q; update_rate_parameter=true)` (rate_parameter === missing) asks the library for the Jacobian-updated rate.
"""
function rate_of(p, d::Int)
    p isa ProbNumDiffEq.IOUP || return (0.0, nothing, 0)
    r = p.rate_parameter
    if p.update_rate_parameter
        ismissing(r) || throw(ArgumentError("Do not manually set the `rate_parameter` of the IOUP prior when using the `update_rate_parameter=true` option.Reset the prior and try again."))
        return (0.0, nothing, 1)
    end
    r isa Number && return (Float64(r), nothing, 0)
    R = r isa AbstractVector ? Matrix{Float64}(LinearAlgebra.Diagonal(r)) : Matrix{Float64}(r)
    size(R) == (d, d) || throw(DimensionMismatch("the IOUP rate parameter must be a number, a vector of length $d or a $d x $d matrix"))
    return (0.0, collect(vec(permutedims(R))), 0)        # row-major for the C side
end
"`initial_diffusion` of a FixedMVDiffusion (a number, a vector or a Diagonal, src/diffusions/typedefs.jl:83-99) as the ONE number the ABI takes."
function uniform_initial_diffusion(x)
    x isa Number && return Float64(x)
    v = x isa LinearAlgebra.Diagonal ? x.diag : x
    (x isa AbstractMatrix && !(x isa LinearAlgebra.Diagonal)) && throw(ArgumentError(
        "Invalid `initial_diffusion`. The `FixedMVDiffusion` assumes a dense diagonal diffusion model. So, pass either a Number, a Vector, or a `Diagonal` matrix!"))
    all(==(first(v)), v) || throw(ArgumentError("libpnode_cuda takes one initial diffusion for all components: a FixedMVDiffusion with " *
                                                "different per-component initial values is not supported by EnsembleB200"))
    return Float64(first(v))
end
function diffusion_fields(m)
    m isa ProbNumDiffEq.FixedMVDiffusion && return (PNODE_DIFF_FIXED_MV, m.calibrate ? 1 : 0, uniform_initial_diffusion(m.initial_diffusion))
    m isa ProbNumDiffEq.DynamicMVDiffusion && return (PNODE_DIFF_DYNAMIC_MV, 0, 1.0)
    m isa ProbNumDiffEq.FixedDiffusion && return (PNODE_DIFF_FIXED, m.calibrate ? 1 : 0, Float64(m.initial_diffusion))
    return (PNODE_DIFF_DYNAMIC, 0, 1.0)
end

"Row-major d x d copy of a constant mass matrix, `nothing` for the identity (UniformScaling λI becomes the dense λ·I)."
function mass_of(f, d::Int)
    M = f.mass_matrix
    M === LinearAlgebra.I && return nothing
    Md = M isa UniformScaling ? Matrix{Float64}(M, d, d) : Matrix{Float64}(M)
    size(Md) == (d, d) || throw(DimensionMismatch("mass_matrix must be $d x $d"))
    Md == Matrix{Float64}(LinearAlgebra.I, d, d) && return nothing
    return collect(vec(permutedims(Md)))
end

"pn_observation_noise -> (row-major d x d covariance or nothing, scalar variance); cov2psdmatrix's accepted types."
function obsnoise_of(R, d::Int)
    (R === nothing || R === missing) && return (nothing, 0.0)
    R isa Number && return (nothing, Float64(R))
    R isa UniformScaling && return (nothing, Float64(R.λ))
    C = R isa ProbNumDiffEq.PSDMatrix ? Matrix{Float64}(R.R' * R.R) : Matrix{Float64}(R isa AbstractMatrix ? R : Diagonal(R))
    size(C) == (d, d) || throw(DimensionMismatch("pn_observation_noise must be $d x $d"))
    return (collect(vec(permutedims(C))), 0.0)
end

ptr_or_null(v) = v === nothing || isempty(v) ? Ptr{Float64}(C_NULL) : pointer(v)

"""
Everything `pnode_config` points to, kept alive together with the struct: the strings, the optional arrays.
"""
struct ConfigBundle
    cfg::Base.RefValue{PnodeConfig}
    keep::Vector{Any}
end

function make_config(prob, alg, device::Integer, lanes::Integer; adaptive, dt, abstol, reltol, save_everystep, save_state, maxiters, dtmin,
                     dtmax, tstops, saveat, beta1, beta2, qmin, qmax, gamma, qsteady_min, qsteady_max, qoldinit, fenrir=nothing,
                     data_forward=false, smooth=alg.smooth)
    second = prob.f isa SciMLBase.DynamicalODEFunction
    dim = length(prob.u0)
    d = second ? dim ÷ 2 : dim
    n_p = prob.p isa SciMLBase.NullParameters ? 0 : length(prob.p)
    q = ProbNumDiffEq.num_derivatives(alg.prior)
    diffusion, calibrate, initial_diffusion = diffusion_fields(alg.diffusionmodel)
    src = trace_to_cuda(prob)
    name = "PnodeUserProblem"
    mm = second ? nothing : mass_of(prob.f, d)
    noise, noise_var = obsnoise_of(alg.pn_observation_noise, d)
    rate, rate_mat, rate_update = rate_of(alg.prior, d)
    ts = tstops === nothing ? Float64[] : sort!(collect(Float64, tstops))
    sa = saveat === nothing ? Float64[] : collect(Float64, saveat)
    data_t, data_u, n_obs, obs_H, obs_cov, obs_var = Float64[], Float64[], 0, nothing, nothing, 0.0
    if fenrir !== nothing
        data_t, data_u, n_obs, obs_H, obs_cov, obs_var = fenrir
    end
    keep = Any[src, name, mm, noise, ts, sa, data_t, data_u, obs_H, obs_cov, rate_mat]
    cfg = PnodeConfig(
        sizeof(PnodeConfig), device, d, q, n_p, alg_code(alg), 0 #= PNODE_COV_AUTO =#, prior_code(alg.prior), diffusion, calibrate,
        smooth ? 1 : 0, adaptive ? 1 : 0, save_everystep ? 1 : 0, save_state ? 1 : 0, lanes, 0, maxiters, initial_diffusion,
        prob.tspan[1], prob.tspan[2], dt, abstol, reltol, dtmin, dtmax,
        beta1, beta2, qmin, qmax, gamma, qsteady_min, qsteady_max, qoldinit,
        Base.unsafe_convert(Cstring, src), Base.unsafe_convert(Cstring, name),
        ptr_or_null(ts), length(ts), Ptr{Float64}(C_NULL), 0, rate, ptr_or_null(sa), length(sa),
        ptr_or_null(mm), second ? 1 : 0, 0,
        ptr_or_null(data_t), length(data_t), ptr_or_null(data_u), n_obs, data_forward ? 1 : 0, ptr_or_null(obs_H), ptr_or_null(obs_cov), obs_var,
        ptr_or_null(noise), noise_var, ptr_or_null(rate_mat), rate_update, 0)
    return ConfigBundle(Ref(cfg), keep)
end

"""
Checks the reference performs before a solve starts, with its exception types and messages: `ErrorException`s of src/checks.jl:1-21 and
src/caches.jl:96-98, the `ArgumentError` for fixed steps without `dt` (raised upstream; test/errors_thrown.jl:6-8).
"""
function check_problem(prob, alg; adaptive, dt, save_everystep, dense)
    prob.u0 isa Number && error("We currently don't support scalar-valued problems")
    if prob.f isa SciMLBase.DynamicalODEFunction && !(prob.problem_type isa SciMLBase.SecondOrderODEProblem)
        error("The given problem is a `DynamicalODEProblem`, but not a `SecondOrderODEProblem`.\n" *
              "This can not be handled by ProbNumDiffEq.jl right now. Please check if the\n" *
              "problem can be formulated as a second order ODE. If not, please open a new\ngithub issue!")
    end
    (dense && !alg.smooth) && error("To use `dense=true` you need to set `smooth=true`!")
    (!save_everystep && alg.smooth) && error("If you set `save_everystep=false` also set `smooth=false` in the alg!")
    (!adaptive && !(dt > 0)) && throw(ArgumentError("Fixed timestep methods require a choice of dt or choosing the tstops"))
    return nothing
end

function SciMLBase.__solve(ep::SciMLBase.AbstractEnsembleProblem, alg::ProbNumDiffEq.AbstractEK, ens::EnsembleB200;
                           trajectories::Integer, adaptive=true, dt=0.0, abstol=1e-6, reltol=1e-3, save_everystep=true,
                           dense=save_everystep && alg.smooth, save_state=false, maxiters=1_000_000, dtmin=0.0, dtmax=0.0, tstops=nothing, saveat=nothing,
                           beta1=0.0, beta2=0.0, qmin=0.2, qmax=10.0, gamma=0.9, qsteady_min=1.0, qsteady_max=1.0, qoldinit=1e-4,
                           kwargs...)
    t_start = time()
    N = Int(trajectories)
    probs = [ep.prob_func(deepcopy(ep.prob), i, 1) for i in 1:N]
    prob = ep.prob
    check_problem(prob, alg; adaptive, dt, save_everystep, dense)
    second = prob.f isa SciMLBase.DynamicalODEFunction
    dsol = length(prob.u0)                               # length of sol.u: (du, u) for second-order problems
    d = second ? dsol ÷ 2 : dsol
    n_p = prob.p isa SciMLBase.NullParameters ? 0 : length(prob.p)
    D = d * (ProbNumDiffEq.num_derivatives(alg.prior) + 1)
    u0 = Matrix{Float64}(undef, dsol, N)                 # column-major d x N == C's [N][d]
    ps = Matrix{Float64}(undef, max(n_p, 1), N)
    for i in 1:N
        u0[:, i] .= vec(collect(probs[i].u0))
        n_p > 0 && (ps[:, i] .= collect(probs[i].p))
    end
    ndev = min(length(ens.devices), N)
    bundles = [make_config(prob, alg, ens.devices[k], ens.lanes_per_traj; adaptive, dt=Float64(dt), abstol=Float64(abstol),
                           reltol=Float64(reltol), save_everystep, save_state, maxiters, dtmin=Float64(dtmin), dtmax=Float64(dtmax),
                           tstops, saveat, beta1, beta2, qmin, qmax, gamma, qsteady_min, qsteady_max, qoldinit) for k in 1:ndev]
    solvers = fill(Ptr{Cvoid}(C_NULL), ndev)
    results = fill(Ptr{Cvoid}(C_NULL), ndev)
    sols = Vector{Any}(undef, N)
    try
        for k in 1:ndev
            b = bundles[k]
            h = Ref{Ptr{Cvoid}}(C_NULL)
            GC.@preserve b check(ccall((:pnode_create, LIB), Cint, (Ref{PnodeConfig}, Ref{Ptr{Cvoid}}), b.cfg, h))
            solvers[k] = h[]
        end
        pp = n_p > 0 ? pointer(ps) : Ptr{Float64}(C_NULL)
        GC.@preserve u0 ps begin
            if ndev == 1
                h = Ref{Ptr{Cvoid}}(C_NULL)
                check(ccall((:pnode_solve_ensemble, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ref{Ptr{Cvoid}}),
                            solvers[1], N, u0, pp, h))
                results[1] = h[]
            else
                check(ccall((:pnode_solve_ensemble_sharded, LIB), Cint,
                            (Ptr{Ptr{Cvoid}}, Int32, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Ptr{Cvoid}}), solvers, ndev, N, u0, pp, results))
            end
        end
        ns = saveat === nothing ? 0 : length(saveat)
        for k in 1:ndev
            lo, hi = shard_range(N, k - 1, ndev)          # 0-based [lo, hi)
            f = copy_fields(results[k]; dsol, D, n=hi - lo, ns)
            for i in 1:(hi - lo)
                sols[lo + i] = build_b200_solution(probs[lo + i], alg, f, i; d, dsol, D, smooth=alg.smooth, save_everystep, saveat)
            end
        end
    finally
        for r in results
            r == C_NULL || ccall((:pnode_result_free, LIB), Cvoid, (Ptr{Cvoid},), r)
        end
        for s in solvers
            s == C_NULL || ccall((:pnode_destroy, LIB), Cvoid, (Ptr{Cvoid},), s)
        end
    end
    out = [ep.output_func(sols[i], i)[1] for i in 1:N]
    converged = all(s -> s.retcode == SciMLBase.ReturnCode.Success, sols)
    return SciMLBase.EnsembleSolution(out, time() - t_start, converged)
end
