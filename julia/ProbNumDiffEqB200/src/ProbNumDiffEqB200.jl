# ProbNumDiffEqB200 -- Julia host side of libpnode_cuda.so (include/pnode.h).
#
# Defines the ensemble algorithm `EnsembleB200` and `SciMLBase.__solve(::AbstractEnsembleProblem, ::AbstractEK, ::EnsembleB200)`:
# the drop-in for `solve(EnsembleProblem(prob; prob_func), EK1(), EnsembleThreads(); trajectories = N)` on this path
# (SURVEY.md section 8b).  NOT EXECUTED in the build image (no Julia, no network); the same ABI is driven by the Python mirror
# probnumdiffeq.jl_b200/{lib,api}.py in every GPU test.  File layout:
#   abi.jl       struct PnodeConfig / PnodeResultInfo, field ids, `check`, `field`
#   trace.jl     Symbolics trace of f -> CUDA C++ `struct PnodeUserProblem { f<T>, jac }` (the printer of tracer.py)
#   solve.jl     keyword mapping, `__solve`, multi-GPU sharding
#   solution.jl  `B200ODESolution` with the field names of ProbNumDiffEq.ProbODESolution (src/solution.jl:33-56)
#   fenrir.jl    `fenrir_data_loglik`, `filtering_data_loglik`, `dalton_data_loglik` over an ensemble of parameters
module ProbNumDiffEqB200

using LinearAlgebra
using ProbNumDiffEq
using RecursiveArrayTools
using SciMLBase
using SymbolicUtils
using Symbolics

export EnsembleB200, B200ODESolution

include("abi.jl")
include("trace.jl")
include("solution.jl")
include("solve.jl")
include("fenrir.jl")

end # module
