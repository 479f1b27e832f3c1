# Symbolic trace of the user's right-hand side into the CUDA C++ source string of `pnode_config.rhs_src`
# (north_star: "f and its Jacobian are traced with Symbolics into CUDA C device functions and compiled with NVRTC").
# The emitted `struct PnodeUserProblem` has the shape tracer.py emits: static d, n_p, jac_parts, f<T> (templated on the scalar so
# that the Taylor-mode initialisation of pn_taylor.cuh runs the same code on truncated Taylor series), f_part / jac_part / jac.
# Expressions are printed with the pn_* helper overloads of csrc/kernels/pn_taylor.cuh -- by walking the expression tree, not by
# editing a string (the first version replaced "^" textually, which is not C).

# SymbolicUtils renamed `istree` to `iscall` in v2; support both.
@static if isdefined(SymbolicUtils, :iscall)
    _iscall(x) = SymbolicUtils.iscall(x)
else
    _iscall(x) = SymbolicUtils.istree(x)
end

const _UNARY = Dict{Any,String}(exp => "pn_exp", log => "pn_log", sin => "pn_sin", cos => "pn_cos", tanh => "pn_tanh",
                                sqrt => "pn_sqrt")

_lit(x::Integer) = string(x, ".0")
_lit(x::Rational) = string("(", numerator(x), ".0/", denominator(x), ".0)")
_lit(x::AbstractFloat) = repr(Float64(x))
_lit(x::Real) = repr(Float64(x))

"C expression of `x`; `names` maps the (unwrapped) symbolic variables to C lvalues such as \"u[0]\"."
function cexpr(x, names::AbstractDict)
    x = Symbolics.unwrap(x)
    x isa Real && return _lit(x)
    haskey(names, x) && return names[x]
    _iscall(x) || throw(ArgumentError("cannot trace $(x): not a function of u, p, t"))
    op = SymbolicUtils.operation(x)
    args = SymbolicUtils.arguments(x)
    c(a) = cexpr(a, names)
    if op === (+)
        return "(" * join(map(c, args), " + ") * ")"
    elseif op === (*)
        return "(" * join(map(c, args), " * ") * ")"
    elseif op === (-)
        return length(args) == 1 ? "(-" * c(args[1]) * ")" : "(" * c(args[1]) * " - " * c(args[2]) * ")"
    elseif op === (/)
        return "(" * c(args[1]) * " / " * c(args[2]) * ")"
    elseif op === (^)
        return cpow(c(args[1]), Symbolics.unwrap(args[2]), names)
    elseif haskey(_UNARY, op)
        return _UNARY[op] * "(" * c(args[1]) * ")"
    end
    throw(ArgumentError("cannot trace a call of $(op): supported are + - * / ^ exp log sin cos tanh sqrt"))
end

"x^e with the printer rules of tracer.py::_print_Pow (no pow() call on the step path for integer and half-integer exponents)."
function cpow(b::String, e, names)
    if e isa Integer
        e == 2 && return "pn_sq(" * b * ")"
        e == -1 && return "(1.0/(" * b * "))"
        e > 0 && return "pn_ipow(" * b * ", " * string(e) * ")"
        return "(1.0/pn_ipow(" * b * ", " * string(-e) * "))"
    elseif e isa Rational && denominator(e) == 2
        numerator(e) == 1 && return "pn_sqrt(" * b * ")"
        numerator(e) == -1 && return "(1.0/pn_sqrt(" * b * "))"
        return "pn_pow_half(" * b * ", " * string(numerator(e)) * ")"
    elseif e isa Real
        isinteger(e) && return cpow(b, Int(e), names)
        2e == round(2e) && return cpow(b, Int(round(2e)) // 2, names)
        return "pn_pow(" * b * ", " * repr(Float64(e)) * ")"
    end
    return "pn_pow(" * b * ", " * cexpr(e, names) * ")"
end

"""
    trace_first_order(prob) -> (exprs::Vector, u, p, t, names)

Call the problem's right-hand side once with symbols.  In-place `f(du, u, p, t)` and out-of-place `f(u, p, t)` problems; for a
`SecondOrderODEProblem` (DynamicalODEFunction) the FIRST-ORDER FORM `F([du; u]) = [f1(du, u); du]` of dimension 2d that the ABI takes
(include/pnode.h, `second_order`).
"""
function trace_first_order(prob)
    second = prob.f isa SciMLBase.DynamicalODEFunction
    dim = length(prob.u0)                      # ArrayPartition(du0, u0) has length 2d
    n_p = prob.p isa SciMLBase.NullParameters ? 0 : length(prob.p)
    u = Symbolics.variables(:u, 1:dim)
    p = Symbolics.variables(:p, 1:max(n_p, 1))[1:n_p]
    t = Symbolics.variable(:t)
    names = Dict{Any,String}(Symbolics.unwrap(t) => "t")
    for i in 1:dim
        names[Symbolics.unwrap(u[i])] = "u[$(i - 1)]"
    end
    for i in 1:n_p
        names[Symbolics.unwrap(p[i])] = "p[$(i - 1)]"
    end
    exprs = Vector{Symbolics.Num}(undef, dim)
    if second
        d = dim ÷ 2
        du_, u_ = u[1:d], u[(d + 1):dim]
        if SciMLBase.isinplace(prob)
            acc = Vector{Symbolics.Num}(undef, d)
            prob.f.f1(acc, du_, u_, p, t)
            exprs[1:d] .= acc
        else
            exprs[1:d] .= prob.f.f1(du_, u_, p, t)
        end
        exprs[(d + 1):dim] .= du_
    elseif SciMLBase.isinplace(prob)
        prob.f.f(exprs, u, p, t)
    else
        exprs .= prob.f.f(u, p, t)
    end
    return exprs, u, p, t, names
end

"CUDA C++ source of `struct PnodeUserProblem` for `prob` (the string handed to `pnode_config.rhs_src`)."
function trace_to_cuda(prob; struct_name::String="PnodeUserProblem")
    exprs, u, p, t, names = trace_first_order(prob)
    dim, n_p = length(u), length(p)
    J = Symbolics.jacobian(exprs, u)            # dense symbolic Jacobian; zeros print as 0.0 and fold away in nvcc
    fbody = join(("    du[$(i - 1)] = " * cexpr(exprs[i], names) * ";" for i in 1:dim), "\n")
    jbody = join(("    J[$((a - 1) + dim * (i - 1))] = " * cexpr(J[a, i], names) * ";" for i in 1:dim for a in 1:dim), "\n")
    return """
    // generated by ProbNumDiffEqB200/src/trace.jl
    struct $struct_name {
      static constexpr int d = $dim;
      static constexpr int n_p = $n_p;
      static constexpr int jac_parts = 1;
      template <class T>
      __device__ __forceinline__ static void f(T* du, const T* u, const double* p, const T& t) {
    $fbody
      }
      template <int PART>
      __device__ __forceinline__ static void f_part(double* du, const double* u, const double* p, double t) { f<double>(du, u, p, t); }
      template <int PART>
      __device__ __forceinline__ static void jac_part(double* J, const double* u, const double* p, double t) {
    $jbody
      }
      __device__ __forceinline__ static void jac(double* J, const double* u, const double* p, double t) { jac_part<0>(J, u, p, t); }
    };
    """
end
