# Solution container with the field names of ProbNumDiffEq.ProbODESolution (src/solution.jl:33-56), filled from the PNODE_F_*
# result fields.  Not a ProbODESolution: that type carries an `EKCache` and an `ODEFilterPosterior` for CPU-side interpolation,
# which an ensemble solved on the device does not have; dense output comes from the device instead (`saveat`, fields
# PNODE_F_SAVEAT_U / _COV) and is stored in `saveat_t`, `saveat_u`, `saveat_pu`.

"`sol.pnstats` (src/solution.jl:9-23)."
mutable struct B200PNStats
    log_likelihood::Float64
end

"""
    B200ODESolution

Field-compatible stand-in for `ProbNumDiffEq.ProbODESolution`: `u, pu, t, x_filt, x_smooth, diffusions, pnstats, prob, alg, dense,
stats, retcode`.  `pu[k]` and the entries of `x_filt` / `x_smooth` are `ProbNumDiffEq.Gaussian`s with `PSDMatrix` covariances
(`Σ.R` is the right factor the library returns).
"""
struct B200ODESolution{uType,puType,tType,xfType,xsType,diffType,P,A,S} <: SciMLBase.AbstractODESolution{Float64,2,uType}
    u::uType
    pu::puType
    u_analytic::Nothing
    errors::Nothing
    t::tType
    k::Nothing
    x_filt::xfType      # Vector of Gaussians, or `nothing`: the library returns ONE state record per point (smoothed if alg.smooth)
    x_smooth::xsType
    diffusions::diffType
    backward_kernels::Nothing
    pnstats::B200PNStats
    prob::P
    alg::A
    interp::Nothing
    cache::Nothing
    dense::Bool
    tslocation::Int
    stats::S
    retcode::SciMLBase.ReturnCode.T
    saveat_t::Vector{Float64}
    saveat_u::Vector{Vector{Float64}}
    saveat_pu::Vector{Matrix{Float64}}
end

"pnode_retcode -> SciMLBase.ReturnCode (include/pnode.h; OrdinaryDiffEqCore's check_error!)."
function retcode_of(code::Integer)
    code == 0 && return SciMLBase.ReturnCode.Success
    code == 1 && return SciMLBase.ReturnCode.MaxIters
    code == 2 && return SciMLBase.ReturnCode.DtLessThanMin
    return SciMLBase.ReturnCode.Unstable
end

"Gaussian(mean, PSDMatrix(R)) of one saved state; `R` arrives row-major (D x D right factor)."
function state_gaussian(mu::AbstractVector, Rrowmajor::AbstractVector, D::Int)
    R = permutedims(reshape(collect(Rrowmajor), D, D))      # row-major bytes -> Julia matrix
    return ProbNumDiffEq.Gaussian(collect(mu), ProbNumDiffEq.PSDMatrix(R))
end

"Gaussian of sol.pu[k]: the library returns the dense d x d covariance; its upper Cholesky factor serves as Σ.R (zeros if singular)."
function pu_gaussian(mu::AbstractVector, covflat::AbstractVector, d::Int)
    C = Symmetric(permutedims(reshape(collect(covflat), d, d)))
    F = cholesky(C; check=false)
    R = issuccess(F) ? Matrix(F.U) : Matrix(sqrt(Symmetric(Matrix(C) + 0I)))   # PSD square root when the covariance is singular (t0)
    return ProbNumDiffEq.Gaussian(collect(mu), ProbNumDiffEq.PSDMatrix(R))
end

"""
Assemble the solution of trajectory `i` (1-based within the shard `res`) from the CSR fields.  `fields` holds the shard's copied
arrays (see `copy_fields`).
"""
function build_b200_solution(prob, alg, fields, i::Int; d::Int, dsol::Int, D::Int, smooth::Bool, save_everystep::Bool, saveat)
    retcode = retcode_of(fields.retcode[i])
    stats = SciMLBase.DEStats(0)
    stats.naccept = fields.naccept[i]
    stats.nreject = fields.nreject[i]
    stats.nf = fields.nf[i]
    pnstats = B200PNStats(fields.loglik[i])
    if save_everystep
        lo, hi = fields.offsets[i] + 1, fields.offsets[i + 1]        # 0-based CSR -> 1-based range
        t = fields.t[lo:hi]
        u = [fields.u[:, k] for k in lo:hi]
        pu = [pu_gaussian(fields.u[:, k], view(fields.ucov, :, k), dsol) for k in lo:hi]
        diffusions = fields.diffusions[lo:(hi - 1)]                   # slot k belongs to the step k -> k+1
        xs = isempty(fields.x) ? nothing : [state_gaussian(view(fields.x, :, k), view(fields.xfac, :, k), D) for k in lo:hi]
    else
        t = [prob.tspan[1], fields.t_end[i]]
        u = [collect(prob.u0), fields.u_final[:, i]]
        pu = [pu_gaussian(collect(prob.u0), zeros(dsol * dsol), dsol), pu_gaussian(fields.u_final[:, i], view(fields.ucov_final, :, i), dsol)]
        diffusions = [fields.diffusion[i]]
        xs = nothing
    end
    x_filt = smooth ? nothing : xs          # the library returns ONE state record per point: smoothed if smooth, else filtered
    x_smooth = smooth ? xs : nothing
    sa_t = saveat === nothing ? Float64[] : collect(Float64, saveat)
    ns = length(sa_t)
    sa_u = ns == 0 ? Vector{Float64}[] : [fields.saveat_u[:, j, i] for j in 1:ns]
    sa_pu = ns == 0 ? Matrix{Float64}[] : [permutedims(reshape(fields.saveat_cov[:, j, i], dsol, dsol)) for j in 1:ns]
    return B200ODESolution(u, pu, nothing, nothing, t, nothing, x_filt, x_smooth, diffusions, nothing, pnstats, prob, alg, nothing, nothing,
                           false, 0, stats, retcode, sa_t, sa_u, sa_pu)
end

"Copy every field a solution needs out of one result handle (one shard)."
function copy_fields(res::Ptr{Cvoid}; dsol::Int, D::Int, n::Int, ns::Int)
    r2(v, rows) = isempty(v) ? reshape(v, rows, 0) : reshape(v, rows, :)
    sa_u = field(res, PNODE_F_SAVEAT_U, Float64)
    sa_c = field(res, PNODE_F_SAVEAT_U_COV, Float64)
    return (
        t_end=field(res, PNODE_F_T_END, Float64), u_final=r2(field(res, PNODE_F_U_FINAL, Float64), dsol),
        ucov_final=r2(field(res, PNODE_F_U_FINAL_COV, Float64), dsol * dsol), loglik=field(res, PNODE_F_LOGLIK, Float64),
        diffusion=field(res, PNODE_F_DIFFUSION, Float64), naccept=field(res, PNODE_F_NACCEPT, Int64),
        nreject=field(res, PNODE_F_NREJECT, Int64), nf=field(res, PNODE_F_NF, Int64), retcode=field(res, PNODE_F_RETCODE, Int32),
        offsets=field(res, PNODE_F_STEP_OFFSETS, Int64), t=field(res, PNODE_F_T, Float64), u=r2(field(res, PNODE_F_U, Float64), dsol),
        ucov=r2(field(res, PNODE_F_U_COV, Float64), dsol * dsol), diffusions=field(res, PNODE_F_DIFFUSIONS, Float64),
        x=r2(field(res, PNODE_F_X, Float64), D), xfac=r2(field(res, PNODE_F_X_COVFAC, Float64), D * D),
        saveat_u=ns == 0 ? zeros(dsol, 0, n) : reshape(sa_u, dsol, ns, n),
        saveat_cov=ns == 0 ? zeros(dsol * dsol, 0, n) : reshape(sa_c, dsol * dsol, ns, n),
        fenrir=field(res, PNODE_F_FENRIR_LOGLIK, Float64),
    )
end
