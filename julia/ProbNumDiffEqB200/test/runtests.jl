# Run with `julia --project -e 'using Pkg; Pkg.test()'` on a machine with Julia, a B200 and libpnode_cuda.so on the loader path
# (or ENV["PNODE_LIBRARY"]).  Mirrors tests/test_gpu_parity.py::test_adaptive_nonstiff_identical_step_sequences at the API level:
# the ensemble solved on the device agrees with the reference's own EnsembleThreads solve.
using Test, ProbNumDiffEq, ProbNumDiffEqB200, SciMLBase

@testset "trace printer" begin
    f!(du, u, p, t) = (du[1] = p[1] * u[1] - u[1]^2 * u[2]; du[2] = -u[2]^3 / sqrt(u[1]) + exp(-t))
    prob = ODEProblem(f!, [1.0, 2.0], (0.0, 1.0), [0.5])
    src = ProbNumDiffEqB200.trace_to_cuda(prob)
    @test occursin("pn_sq(u[0])", src) && occursin("pn_ipow(u[1], 3)", src) && occursin("pn_sqrt(u[0])", src)
    @test !occursin("^", src)
end

@testset "EnsembleB200 vs EnsembleThreads" begin
    function lv!(du, u, p, t)
        du[1] = p[1] * u[1] - p[2] * u[1] * u[2]
        du[2] = -p[3] * u[2] + p[4] * u[1] * u[2]
    end
    prob = ODEProblem(lv!, [1.0, 1.0], (0.0, 10.0), [1.5, 1.0, 3.0, 1.0])
    ens = EnsembleProblem(prob; prob_func=(pr, i, rep) -> remake(pr; u0=pr.u0 .* (1 + 0.01 * i)))
    ref = solve(ens, EK1(order=3, smooth=false), EnsembleThreads(); trajectories=8, dense=false)
    b200 = solve(ens, EK1(order=3, smooth=false), EnsembleB200(); trajectories=8)
    for (a, b) in zip(ref, b200)
        @test b.retcode == SciMLBase.ReturnCode.Success
        @test length(a.t) == length(b.t)
        @test isapprox(a.u[end], b.u[end]; rtol=1e-8)
    end
end
