# The data likelihoods of src/data_likelihoods/ for an ensemble of parameters against ONE data set: one value per trajectory.
#   fenrir_data_loglik     (fenrir.jl:30-137)    backward pass with data updates            -> PNODE_F_FENRIR_LOGLIK
#   filtering_data_loglik  (filtering.jl)        DataUpdateCallback in the forward pass     -> PNODE_F_DATA_LOGLIK
#   dalton_data_loglik     (dalton.jl)           data_ll + pn_ll(with data) - pn_ll(without), two solves

"Observation model of the three likelihoods as the C side takes it: (data_t, data_u row-major, n_obs, H row-major | nothing, cov | nothing, var)."
function data_model(data, observation_matrix, observation_noise_cov)
    data_t = collect(Float64, data.t)
    ut = reduce(hcat, (collect(Float64, y) for y in data.u))      # n_obs x n_data column-major == C's [n_data][n_obs]
    n_obs = size(ut, 1)
    obs_H = observation_matrix === LinearAlgebra.I ? nothing : collect(vec(permutedims(Matrix{Float64}(observation_matrix))))
    obs_cov, obs_var = observation_noise_cov isa Number ? (nothing, Float64(observation_noise_cov)) :
                       observation_noise_cov isa UniformScaling ? (nothing, Float64(observation_noise_cov.λ)) :
                       (collect(vec(permutedims(Matrix{Float64}(observation_noise_cov)))), 0.0)
    return (data_t, collect(vec(ut)), n_obs, obs_H, obs_cov, obs_var)
end

"One ensemble solve with (or, `model === nothing`, without) the data set; returns the requested per-trajectory fields."
function data_solve(ep, alg, ens::EnsembleB200, model, fields; trajectories::Integer, forward::Bool, smooth::Bool, save_everystep::Bool,
                    adaptive=true, dt=0.0, abstol=1e-6, reltol=1e-3, maxiters=1_000_000, dtmin=0.0, dtmax=0.0, tstops=nothing, kwargs...)
    N = Int(trajectories)
    prob = ep.prob
    probs = [ep.prob_func(deepcopy(prob), i, 1) for i in 1:N]
    dsol = length(prob.u0)
    n_p = prob.p isa SciMLBase.NullParameters ? 0 : length(prob.p)
    b = make_config(prob, alg, ens.devices[1], ens.lanes_per_traj; adaptive, dt=Float64(dt), abstol=Float64(abstol), reltol=Float64(reltol),
                    save_everystep, save_state=false, maxiters, dtmin=Float64(dtmin), dtmax=Float64(dtmax), tstops, saveat=nothing,
                    beta1=0.0, beta2=0.0, qmin=0.2, qmax=10.0, gamma=0.9, qsteady_min=1.0, qsteady_max=1.0, qoldinit=1e-4,
                    fenrir=model, data_forward=forward, smooth)
    u0 = Matrix{Float64}(undef, dsol, N)
    ps = Matrix{Float64}(undef, max(n_p, 1), N)
    for i in 1:N
        u0[:, i] .= vec(collect(probs[i].u0))
        n_p > 0 && (ps[:, i] .= collect(probs[i].p))
    end
    solver, res = Ref{Ptr{Cvoid}}(C_NULL), Ref{Ptr{Cvoid}}(C_NULL)
    try
        GC.@preserve b check(ccall((:pnode_create, LIB), Cint, (Ref{PnodeConfig}, Ref{Ptr{Cvoid}}), b.cfg, solver))
        GC.@preserve u0 ps check(ccall((:pnode_solve_ensemble, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ref{Ptr{Cvoid}}),
                                       solver[], N, u0, n_p > 0 ? pointer(ps) : Ptr{Float64}(C_NULL), res))
        return [field(res[], f, Float64) for f in fields]
    finally
        res[] == C_NULL || ccall((:pnode_result_free, LIB), Cvoid, (Ptr{Cvoid},), res[])
        solver[] == C_NULL || ccall((:pnode_destroy, LIB), Cvoid, (Ptr{Cvoid},), solver[])
    end
end

"""
    fenrir_data_loglik(ensembleprob, alg, EnsembleB200(); data = (t = ..., u = ...), observation_noise_cov, observation_matrix = I,
                       trajectories, kwargs...)

Vector of Fenrir log-likelihoods, one per trajectory (`-Inf` where the solve did not succeed, fenrir.jl:55-59).
"""
function ProbNumDiffEq.fenrir_data_loglik(ep::SciMLBase.AbstractEnsembleProblem, alg::ProbNumDiffEq.AbstractEK, ens::EnsembleB200;
                                          data::NamedTuple{(:t, :u)}, observation_noise_cov, observation_matrix=LinearAlgebra.I,
                                          trajectories::Integer, kwargs...)
    alg.smooth || throw(ArgumentError("fenrir only works with smoothing. Set `smooth=true`."))   # fenrir.jl:42-44
    model = data_model(data, observation_matrix, observation_noise_cov)
    return data_solve(ep, alg, ens, model, (PNODE_F_FENRIR_LOGLIK,); trajectories, forward=false, smooth=true, save_everystep=true, kwargs...)[1]
end

"""
    filtering_data_loglik(ensembleprob, alg, EnsembleB200(); data, observation_noise_cov, observation_matrix = I, trajectories, kwargs...)

`DataUpdateLogLikelihood.ll` of a solve with the `DataUpdateCallback` at `data.t` (src/data_likelihoods/filtering.jl), per trajectory.
"""
function ProbNumDiffEq.filtering_data_loglik(ep::SciMLBase.AbstractEnsembleProblem, alg::ProbNumDiffEq.AbstractEK, ens::EnsembleB200;
                                             data::NamedTuple{(:t, :u)}, observation_noise_cov, observation_matrix=LinearAlgebra.I,
                                             trajectories::Integer, kwargs...)
    model = data_model(data, observation_matrix, observation_noise_cov)
    return data_solve(ep, alg, ens, model, (PNODE_F_DATA_LOGLIK,); trajectories, forward=true, smooth=false, save_everystep=false, kwargs...)[1]
end

"""
    dalton_data_loglik(ensembleprob, alg, EnsembleB200(); data, observation_noise_cov, observation_matrix = I, trajectories,
                       adaptive = false, dt, kwargs...)

DALTON: `data_ll + pn_ll(with data) - pn_ll(without data)` (src/data_likelihoods/dalton.jl), per trajectory; fixed steps only.
"""
function ProbNumDiffEq.dalton_data_loglik(ep::SciMLBase.AbstractEnsembleProblem, alg::ProbNumDiffEq.AbstractEK, ens::EnsembleB200;
                                          data::NamedTuple{(:t, :u)}, observation_noise_cov, observation_matrix=LinearAlgebra.I,
                                          trajectories::Integer, kwargs...)
    (:adaptive in keys(kwargs)) || throw(ArgumentError("`dalton_nll` only works with fixed step sizes. Set `adaptive=false`."))   # dalton.jl
    model = data_model(data, observation_matrix, observation_noise_cov)
    tstops = union(collect(Float64, data.t), get(kwargs, :tstops, Float64[]))
    kw = Base.structdiff((; kwargs...), (tstops=nothing,))
    data_ll, with_ll = data_solve(ep, alg, ens, model, (PNODE_F_DATA_LOGLIK, PNODE_F_LOGLIK); trajectories, forward=true, smooth=false,
                                  save_everystep=false, tstops, kw...)
    without_ll = data_solve(ep, alg, ens, nothing, (PNODE_F_LOGLIK,); trajectories, forward=false, smooth=false, save_everystep=false,
                            tstops, kw...)[1]
    return data_ll .+ with_ll .- without_ll
end
