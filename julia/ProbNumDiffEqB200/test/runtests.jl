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

# The parity contract of BASELINE.json's north_star against the LIVE reference (what cannot be run in the build image; the Python test-suite
# checks the same statements against the CPU restatement under oracle/): fixed steps -> filter and smoother means and covariances rtol 1e-10;
# adaptive steps -> identical accepted step sequences, solutions rtol 1e-8.
@testset "north_star parity: fixed steps, filter + smoother (BASELINE cfg1)" begin
    function fhn!(du, u, p, t)
        a, b, c = p
        du[1] = c * (u[1] - u[1]^3 / 3 + u[2])
        du[2] = -(1 / c) * (u[1] - a - b * u[2])
    end
    prob = ODEProblem(fhn!, [-1.0, 1.0], (0.0, 20.0), [0.2, 0.2, 3.0])
    ens = EnsembleProblem(prob)
    for alg in (EK1(order=3, diffusionmodel=FixedDiffusion()), EK1(order=3, diffusionmodel=FixedDiffusion(), smooth=false),
                EK0(order=3, diffusionmodel=FixedDiffusion()), DiagonalEK1(order=3, diffusionmodel=FixedDiffusion()))
        ref = solve(prob, alg; adaptive=false, dt=0.01, dense=alg.smooth)
        b200 = solve(ens, alg, EnsembleB200(); trajectories=1, adaptive=false, dt=0.01)[1]
        @test b200.t == ref.t
        @test all(isapprox(a, b; rtol=1e-10) for (a, b) in zip(ref.u, b200.u))
        @test all(isapprox(Matrix(a.Σ), Matrix(b.Σ); rtol=1e-10, atol=1e-300) for (a, b) in zip(ref.pu[2:end], b200.pu[2:end]))
        @test isapprox(ref.pnstats.log_likelihood, b200.pnstats.log_likelihood; rtol=1e-9)
    end
end

@testset "north_star parity: adaptive steps, identical step sequences (BASELINE cfg2 / cfg5 shapes)" begin
    function lv!(du, u, p, t)
        du[1] = p[1] * u[1] - p[2] * u[1] * u[2]
        du[2] = -p[3] * u[2] + p[4] * u[1] * u[2]
    end
    function rober!(du, u, p, t)
        du[1] = -p[1] * u[1] + p[3] * u[2] * u[3]
        du[2] = p[1] * u[1] - p[3] * u[2] * u[3] - p[2] * u[2]^2
        du[3] = p[2] * u[2]^2
    end
    cases = ((ODEProblem(lv!, [1.0, 1.0], (0.0, 10.0), [1.5, 1.0, 3.0, 1.0]), EK0(order=3, smooth=false)),
             (ODEProblem(lv!, [1.0, 1.0], (0.0, 10.0), [1.5, 1.0, 3.0, 1.0]), EK1(order=3, smooth=true)),
             (ODEProblem(rober!, [1.0, 0.0, 0.0], (0.0, 100.0), [0.04, 3e7, 1e4]), EK1(order=3, smooth=false)))
    for (prob, alg) in cases
        ens = EnsembleProblem(prob; prob_func=(pr, i, rep) -> remake(pr; p=pr.p .* (1 + 0.01 * (i - 1))))
        ref = solve(ens, alg, EnsembleThreads(); trajectories=4, save_everystep=alg.smooth, dense=alg.smooth)
        b200 = solve(ens, alg, EnsembleB200(); trajectories=4, save_everystep=alg.smooth)
        for (a, b) in zip(ref, b200)
            @test (a.stats.naccept, a.stats.nreject, a.stats.nf) == (b.stats.naccept, b.stats.nreject, b.stats.nf)
            @test isapprox(a.t, b.t; rtol=1e-12)
            @test isapprox(a.u[end], b.u[end]; rtol=1e-8)
        end
    end
end

@testset "error paths keep the reference's exception types (test/errors_thrown.jl)" begin
    lin!(du, u, p, t) = (du[1] = p[1] * u[1])
    ens = EnsembleProblem(ODEProblem(lin!, [1.0], (0.0, 1.0), [-1.0]))
    @test_throws ArgumentError solve(ens, EK0(), EnsembleB200(); trajectories=2, adaptive=false)
    @test_throws ErrorException solve(ens, EK0(smooth=false), EnsembleB200(); trajectories=2, dense=true)
    @test_throws ErrorException solve(ens, EK0(smooth=true), EnsembleB200(); trajectories=2, save_everystep=false)
    @test_throws ArgumentError solve(ens, EK1(diffusionmodel=FixedMVDiffusion()), EnsembleB200(); trajectories=2)
end
