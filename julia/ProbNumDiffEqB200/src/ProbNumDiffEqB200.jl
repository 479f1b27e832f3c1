# See INTEGRATION.md. Host-side binding of include/pnode.h (unexecuted: no Julia in the build image).
module ProbNumDiffEqB200
using ProbNumDiffEq, SciMLBase, Symbolics, LinearAlgebra
const LIB = "libpnode_cuda"

struct EnsembleB200 <: SciMLBase.EnsembleAlgorithm
    device::Int32
end
EnsembleB200() = EnsembleB200(0)

# mirrors `struct pnode_config` (include/pnode.h) field for field
struct PnodeConfig
    struct_size::Int32; device::Int32; d::Int32; q::Int32; n_p::Int32; alg::Int32; cov::Int32; prior::Int32
    diffusion::Int32; calibrate::Int32; smooth::Int32; adaptive::Int32; save_everystep::Int32; save_state::Int32
    lanes_per_traj::Int32; reserved0::Int32; maxiters::Int64; initial_diffusion::Float64
    t0::Float64; t1::Float64; dt::Float64; abstol::Float64; reltol::Float64; dtmin::Float64; dtmax::Float64
    beta1::Float64; beta2::Float64; qmin::Float64; qmax::Float64; gamma::Float64
    qsteady_min::Float64; qsteady_max::Float64; qoldinit::Float64
    rhs_src::Cstring; rhs_name::Cstring
    tstops::Ptr{Float64}; n_tstops::Int64; forced_diffusion::Ptr{Float64}; n_forced::Int64
    rate_parameter::Float64                      # IOUP(q, rate) prior; ignored for the IWP
    saveat::Ptr{Float64}; n_saveat::Int64        # dense output sol(saveat) evaluated on the device
    mass_matrix::Ptr{Float64}                    # d × d row-major, C_NULL = identity (ODEFunction(f; mass_matrix=M))
    second_order::Int32; reserved1::Int32        # SecondOrderODEProblem: rhs is the first-order form on (du, u)
    data_t::Ptr{Float64}; n_data::Int64; data_u::Ptr{Float64}; n_obs::Int32; reserved2::Int32   # fenrir_data_loglik
    observation_matrix::Ptr{Float64}; observation_noise_cov::Ptr{Float64}; observation_noise_var::Float64
end

# enum values of include/pnode.h
alg_code(alg) = alg isa EK0 ? 0 : alg isa DiagonalEK1 ? 2 : 1
prior_code(p) = p isa ProbNumDiffEq.IOUP ? 1 : 0
rate_of(p) = p isa ProbNumDiffEq.IOUP ? Float64(p.rate_parameter) : 0.0
# row-major copy of a constant mass matrix; `nothing` for I (UniformScaling λI becomes the dense λ·I)
mass_of(f, d) = f.mass_matrix === LinearAlgebra.I ? nothing : collect(transpose(Matrix{Float64}(f.mass_matrix * LinearAlgebra.I(d))))
diffusion_code(m) = m isa FixedMVDiffusion ? 3 : m isa DynamicMVDiffusion ? 2 : m isa FixedDiffusion ? 1 : 0

check(rc) = rc == 0 || throw(ArgumentError(unsafe_string(ccall((:pnode_last_error, LIB), Cstring, ()))))

function SciMLBase.__solve(ep::SciMLBase.AbstractEnsembleProblem, alg::ProbNumDiffEq.AbstractEK, ens::EnsembleB200;
                           trajectories, adaptive=true, dt=0.0, abstol=1e-6, reltol=1e-3, save_everystep=true, kwargs...)
    probs = [ep.prob_func(deepcopy(ep.prob), i, 1) for i in 1:trajectories]
    d, n_p = length(ep.prob.u0), length(ep.prob.p)
    u0 = reduce(hcat, (p.u0 for p in probs))          # d × N column-major == [N][d] row-major
    ps = reduce(hcat, (collect(p.p) for p in probs))
    src = trace_to_cuda(ep.prob.f, d, n_p)            # Symbolics → "struct PnodeUserProblem { f<T>, jac }"
    mm = mass_of(ep.prob.f, d)
    GC.@preserve src mm begin
        cfg = Ref(PnodeConfig(sizeof(PnodeConfig), ens.device, d, ProbNumDiffEq.num_derivatives(alg.prior), n_p,
                              alg_code(alg), 0, prior_code(alg.prior), diffusion_code(alg.diffusionmodel), 1,
                              alg.smooth, adaptive, save_everystep, 0, 0, 0, 1_000_000, 1.0,
                              ep.prob.tspan[1], ep.prob.tspan[2], dt, abstol, reltol, 0.0, 0.0,
                              0, 0, 0.2, 10.0, 0.9, 1.0, 1.0, 1e-4,
                              pointer(src), pointer("PnodeUserProblem"), C_NULL, 0, C_NULL, 0, rate_of(alg.prior), C_NULL, 0,
                              mm === nothing ? Ptr{Float64}(C_NULL) : pointer(mm),
                              0, 0, C_NULL, 0, C_NULL, 0, 0, C_NULL, C_NULL, 0.0))   # first-order problem, no data set
        solver = Ref{Ptr{Cvoid}}()
        check(ccall((:pnode_create, LIB), Cint, (Ref{PnodeConfig}, Ref{Ptr{Cvoid}}), cfg, solver))
        res = Ref{Ptr{Cvoid}}()
        check(ccall((:pnode_solve_ensemble, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ref{Ptr{Cvoid}}),
                    solver[], trajectories, u0, ps, res))
    end
    off = field(res[], PNODE_F_STEP_OFFSETS, Int64); t = field(res[], PNODE_F_T, Float64)   # pnode_result_copy
    u   = reshape(field(res[], PNODE_F_U, Float64), d, :)
    sols = [build_b200_solution(probs[i], alg, t[off[i]+1:off[i+1]], u[:, off[i]+1:off[i+1]], res[], i) for i in 1:trajectories]
    ccall((:pnode_result_free, LIB), Cvoid, (Ptr{Cvoid},), res[]); ccall((:pnode_destroy, LIB), Cvoid, (Ptr{Cvoid},), solver[])
    return SciMLBase.EnsembleSolution(sols, 0.0, true)
end
end

# ---- helpers referenced above ------------------------------------------------------------------
const PNODE_F_STEP_OFFSETS, PNODE_F_T, PNODE_F_U, PNODE_F_U_COV = 12, 13, 14, 15

function field(res::Ptr{Cvoid}, f::Integer, ::Type{T}) where {T}
    nb = ccall((:pnode_result_field_bytes, LIB), Int64, (Ptr{Cvoid}, Cint), res, f)
    out = Vector{T}(undef, nb ÷ sizeof(T))
    check(ccall((:pnode_result_copy, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Csize_t), res, f, out, nb))
    return out
end

"Trace `f(du,u,p,t)` with Symbolics and print `struct PnodeUserProblem` (same shape as tracer.py emits)."
function trace_to_cuda(f, d, n_p)
    Symbolics.@variables u[1:d] p[1:n_p] t
    du = Vector{Symbolics.Num}(undef, d)
    f(du, collect(u), collect(p), t)
    J = Symbolics.jacobian(du, collect(u))
    fbody = join(["    du[$(i-1)] = $(cexpr(du[i]));" for i in 1:d], "\n")
    jbody = join(["    J[$((a-1) + d*(i-1))] = $(cexpr(J[a, i]));" for i in 1:d for a in 1:d], "\n")
    return """
    struct PnodeUserProblem {
      static constexpr int d = $d; static constexpr int n_p = $n_p;
      template <class T> __device__ __forceinline__ static void f(T* du, const T* u, const double* p, const T& t) {
    $fbody
      }
      __device__ __forceinline__ static void jac(double* J, const double* u, const double* p, double t) {
    $jbody
      }
    };"""
end
cexpr(ex) = replace(string(Symbolics.toexpr(ex)), r"u\[(\d+)\]" => s -> "u[" * string(parse(Int, match(r"\d+", s).match) - 1) * "]",
                    r"p\[(\d+)\]" => s -> "p[" * string(parse(Int, match(r"\d+", s).match) - 1) * "]", "^" => "pn_pow")

function build_b200_solution(prob, alg, t, u, res, i)
    # thin container with the reference's field names (src/solution.jl:33-56)
    return (; t, u=[u[:, k] for k in axes(u, 2)], prob, alg, retcode=SciMLBase.ReturnCode.Success)
end
