# Binding of include/pnode.h.  Every struct mirrors its C counterpart field for field (same order, same widths).

"Name or path of the shared library; override with ENV[\"PNODE_LIBRARY\"] before loading the package."
const LIB = get(ENV, "PNODE_LIBRARY", "libpnode_cuda")

# struct pnode_config (include/pnode.h)
struct PnodeConfig
    struct_size::Int32
    device::Int32
    d::Int32
    q::Int32
    n_p::Int32
    alg::Int32
    cov::Int32
    prior::Int32
    diffusion::Int32
    calibrate::Int32
    smooth::Int32
    adaptive::Int32
    save_everystep::Int32
    save_state::Int32
    lanes_per_traj::Int32
    reserved0::Int32
    maxiters::Int64
    initial_diffusion::Float64
    t0::Float64
    t1::Float64
    dt::Float64
    abstol::Float64
    reltol::Float64
    dtmin::Float64
    dtmax::Float64
    beta1::Float64
    beta2::Float64
    qmin::Float64
    qmax::Float64
    gamma::Float64
    qsteady_min::Float64
    qsteady_max::Float64
    qoldinit::Float64
    rhs_src::Cstring
    rhs_name::Cstring
    tstops::Ptr{Float64}
    n_tstops::Int64
    forced_diffusion::Ptr{Float64}
    n_forced::Int64
    rate_parameter::Float64
    saveat::Ptr{Float64}
    n_saveat::Int64
    mass_matrix::Ptr{Float64}
    second_order::Int32
    reserved1::Int32
    data_t::Ptr{Float64}
    n_data::Int64
    data_u::Ptr{Float64}
    n_obs::Int32
    data_forward::Int32                # 0: Fenrir backward pass; 1: DataUpdateCallback (forward updates at the data times)
    observation_matrix::Ptr{Float64}
    observation_noise_cov::Ptr{Float64}
    observation_noise_var::Float64
    pn_observation_noise::Ptr{Float64}
    pn_observation_noise_var::Float64
    rate_matrix::Ptr{Float64}          # IOUP with a vector / matrix rate: d x d row-major (NULL: scalar rate_parameter)
    update_rate_parameter::Int32       # IOUP(q; update_rate_parameter=true): rate := Jacobian before every step
    reserved2::Int32
end

# struct pnode_result_info
struct PnodeResultInfo
    n_traj::Int64
    total_saved::Int64
    total_accepted::Int64
    total_rejected::Int64
    d::Int32
    q::Int32
    D::Int32
    lanes_per_traj::Int32
    kernel_ms::Float64
    h2d_ms::Float64
    d2h_ms::Float64
    n_launches::Int64
    h2d_bytes::Int64
    d2h_bytes::Int64
end

# enum pnode_field
const PNODE_F_T_END = 0
const PNODE_F_U_FINAL = 1
const PNODE_F_U_FINAL_COV = 2
const PNODE_F_X_FINAL = 3
const PNODE_F_X_FINAL_COVFAC = 4
const PNODE_F_LOGLIK = 5
const PNODE_F_DIFFUSION = 6
const PNODE_F_DT_LAST = 7
const PNODE_F_NACCEPT = 8
const PNODE_F_NREJECT = 9
const PNODE_F_NF = 10
const PNODE_F_RETCODE = 11
const PNODE_F_STEP_OFFSETS = 12
const PNODE_F_T = 13
const PNODE_F_U = 14
const PNODE_F_U_COV = 15
const PNODE_F_DIFFUSIONS = 16
const PNODE_F_X = 17
const PNODE_F_X_COVFAC = 18
const PNODE_F_DIFFUSION_MV = 19
const PNODE_F_DIFFUSIONS_MV = 20
const PNODE_F_SAVEAT_U = 21
const PNODE_F_SAVEAT_U_COV = 22
const PNODE_F_FENRIR_LOGLIK = 23
const PNODE_F_DATA_LOGLIK = 24

# enum pnode_alg / pnode_prior / pnode_diffusion / pnode_retcode, status codes
const PNODE_EK0, PNODE_EK1, PNODE_DIAGONAL_EK1 = 0, 1, 2
const PNODE_PRIOR_IWP, PNODE_PRIOR_IOUP = 0, 1
const PNODE_DIFF_DYNAMIC, PNODE_DIFF_FIXED, PNODE_DIFF_DYNAMIC_MV, PNODE_DIFF_FIXED_MV = 0, 1, 2, 3
const PNODE_ERR_ARGUMENT, PNODE_ERR_DIMENSION = -1, -2

last_error() = unsafe_string(ccall((:pnode_last_error, LIB), Cstring, ()))

"Status of a C call -> the exception type the reference raises for the same mistake (SURVEY.md section 8b, errors)."
function check(rc::Integer)
    rc == 0 && return nothing
    msg = last_error()
    rc == PNODE_ERR_ARGUMENT && throw(ArgumentError(msg))
    rc == PNODE_ERR_DIMENSION && throw(DimensionMismatch(msg))
    error(msg)
end

"Copy one result field out of a `pnode_result` handle (pnode_result_copy); empty vector if the field was not produced."
function field(res::Ptr{Cvoid}, f::Integer, ::Type{T}) where {T}
    nb = ccall((:pnode_result_field_bytes, LIB), Int64, (Ptr{Cvoid}, Cint), res, f)
    out = Vector{T}(undef, nb ÷ sizeof(T))
    nb == 0 && return out
    check(ccall((:pnode_result_copy, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Csize_t), res, f, out, nb))
    return out
end

function result_info(res::Ptr{Cvoid})
    info = Ref{PnodeResultInfo}()
    check(ccall((:pnode_result_info_get, LIB), Cint, (Ptr{Cvoid}, Ref{PnodeResultInfo}), res, info))
    return info[]
end

device_count() = Int(ccall((:pnode_device_count, LIB), Cint, ()))

"[lo, hi) (0-based) of shard `k` (0-based) out of `n` shards: pnode_shard_range."
function shard_range(n_traj::Integer, k::Integer, n::Integer)
    lo, hi = Ref{Int64}(0), Ref{Int64}(0)
    ccall((:pnode_shard_range, LIB), Cvoid, (Int64, Int32, Int32, Ref{Int64}, Ref{Int64}), n_traj, k, n, lo, hi)
    return lo[], hi[]
end
