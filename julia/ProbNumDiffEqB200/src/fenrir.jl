# `fenrir_data_loglik` (src/data_likelihoods/fenrir.jl:30-137) for an ensemble of parameters against one data set: one
# log-likelihood per trajectory comes back in PNODE_F_FENRIR_LOGLIK.

"""
    fenrir_data_loglik(ensembleprob, alg, EnsembleB200(); data = (t = ..., u = ...), observation_noise_cov, observation_matrix = I,
                       trajectories, kwargs...)

Vector of Fenrir log-likelihoods, one per trajectory (`-Inf` where the solve did not succeed, fenrir.jl:55-59).
"""
function ProbNumDiffEq.fenrir_data_loglik(ep::SciMLBase.AbstractEnsembleProblem, alg::ProbNumDiffEq.AbstractEK, ens::EnsembleB200;
                                          data::NamedTuple{(:t, :u)}, observation_noise_cov, observation_matrix=LinearAlgebra.I,
                                          trajectories::Integer, adaptive=true, dt=0.0, abstol=1e-6, reltol=1e-3, maxiters=1_000_000,
                                          dtmin=0.0, dtmax=0.0, tstops=nothing, kwargs...)
    alg.smooth || throw(ArgumentError("fenrir only works with smoothing. Set `smooth=true`."))   # fenrir.jl:42-44
    N = Int(trajectories)
    prob = ep.prob
    probs = [ep.prob_func(deepcopy(prob), i, 1) for i in 1:N]
    dsol = length(prob.u0)
    n_p = prob.p isa SciMLBase.NullParameters ? 0 : length(prob.p)
    data_t = collect(Float64, data.t)
    ut = reduce(hcat, (collect(Float64, y) for y in data.u))      # n_obs x n_data column-major == C's [n_data][n_obs]
    n_obs = size(ut, 1)
    obs_H = observation_matrix === LinearAlgebra.I ? nothing : collect(vec(permutedims(Matrix{Float64}(observation_matrix))))
    obs_cov, obs_var = observation_noise_cov isa Number ? (nothing, Float64(observation_noise_cov)) :
                       observation_noise_cov isa UniformScaling ? (nothing, Float64(observation_noise_cov.λ)) :
                       (collect(vec(permutedims(Matrix{Float64}(observation_noise_cov)))), 0.0)
    b = make_config(prob, alg, ens.devices[1], ens.lanes_per_traj; adaptive, dt=Float64(dt), abstol=Float64(abstol), reltol=Float64(reltol),
                    save_everystep=true, save_state=false, maxiters, dtmin=Float64(dtmin), dtmax=Float64(dtmax), tstops, saveat=nothing,
                    beta1=0.0, beta2=0.0, qmin=0.2, qmax=10.0, gamma=0.9, qsteady_min=1.0, qsteady_max=1.0, qoldinit=1e-4,
                    fenrir=(data_t, collect(vec(ut)), n_obs, obs_H, obs_cov, obs_var))
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
        return field(res[], PNODE_F_FENRIR_LOGLIK, Float64)
    finally
        res[] == C_NULL || ccall((:pnode_result_free, LIB), Cvoid, (Ptr{Cvoid},), res[])
        solver[] == C_NULL || ccall((:pnode_destroy, LIB), Cvoid, (Ptr{Cvoid},), solver[])
    end
end
