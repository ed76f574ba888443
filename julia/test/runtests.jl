# runtests.jl — what a maintainer with Julia, HighDimPDE.jl and a B200 runs to accept the glue.
# NOT EXECUTED IN THIS REPOSITORY'S CI (the build image has no Julia); the Python twin of every check below runs in
# tests/ (test_gpu_parity.py, test_gpu_baseline_configs.py, test_golden.py).
#
#     julia --project=/path/to/HighDimPDE.jl -e 'using Pkg; Pkg.add(["NPZ"])'                 # once
#     HDPDE_B200_LIB=/path/to/libhdpde_b200.so julia --project=/path/to/HighDimPDE.jl julia/test/runtests.jl
#
# 1. loading the glue creates no method ambiguity with the reference (the defect round 1 shipped);
# 2. the shapes of the reference's own test/MLP.jl:19-317, through unchanged `solve(prob, MLP(...))` calls, with the
#    reference's own tolerances;
# 3. the golden tapes of tests/golden through the glue (`mlp_b200(...; tapes)`, `_reflect`, `sde_loop_b200!`): the CUDA
#    path against the oracle's stored outputs, 1e-12 / 1e-5;
# 4. the fall-back: problems whose functions are plain closures still run the stock CPU recursion.
using Test, Random, Statistics, LinearAlgebra
using HighDimPDE, NPZ
include(joinpath(@__DIR__, "..", "HighDimPDEB200.jl"))
using .HighDimPDEB200
const B = HighDimPDEB200
const GOLD = joinpath(@__DIR__, "..", "..", "tests", "golden")
rel(a, b) = abs(a - b) / abs(b)

@testset "HighDimPDEB200" begin
    @testset "no ambiguities" begin
        amb = Test.detect_ambiguities(HighDimPDE, HighDimPDEB200; recursive = true)
        @test isempty(filter(p -> any(m -> m.module === HighDimPDEB200, p), amb))
        # the calls the stock CPU path makes still resolve to the reference's methods
        a, b = zeros(3), [0.7, -0.2, 0.1]
        @test which(HighDimPDE._reflect, typeof.((a, b, fill(-0.5, 3), fill(0.5, 3)))).module === HighDimPDE
        @test which(HighDimPDE._reflect, typeof.((zeros(3, 4), zeros(3, 4), -0.5, 0.5))).module === HighDimPDEB200
    end

    @testset "test/MLP.jl shapes through solve(prob, MLP())" begin
        Random.seed!(100)
        # heat equation, d = 10 (test/MLP.jl:19-47): u = sum(x.^2) + d σ² T
        d = 10
        prob = ParabolicPDEProblem(B.ConstCoef(0.0), B.ConstCoef(0.1), fill(2.0, d), (0.0, 0.5); g = B.SumSq(0.0), f = B.FZero())
        for mt in (false, true)
            sol = solve(prob, MLP(); multithreading = mt)
            @test rel(sol.us[end], 40.05) < 0.1
            @test sol.us[1] == 40.0 && sol.ts == [0.0, 0.5]
        end
        # with Neumann reflection and Float32 time span / σ promoted by the glue (test/MLP.jl:80-105)
        prob = ParabolicPDEProblem(B.ConstCoef(0.0), B.ConstCoef(0.1), fill(0.3, 5), (0.0, 0.5); g = B.SumSq(0.0), f = B.FZero(),
                                   neumann_bc = [fill(-0.5, 5), fill(0.5, 5)])
        u1 = solve(prob, MLP(); multithreading = false).us[end]
        u2 = solve(prob, MLP(); multithreading = false).us[end]
        @test rel(u1, u2) < 0.1
        # exponential (test/MLP.jl:107-134): u = g e^{rT}
        prob = ParabolicPDEProblem(B.ConstCoef(0.0), B.ConstCoef(0.0), zeros(3), (0.0, 0.5); g = B.SumSq(2.0), f = B.FLinear(0.1))
        @test rel(solve(prob, MLP(); multithreading = false).us[end], 2.0 * exp(0.05)) < 1e-2
        # Allen–Cahn (test/MLP.jl:136-160, literature value 0.30879 with the test's tolerance)
        prob = ParabolicPDEProblem(B.ConstCoef(0.0), B.ConstCoef(1.0), zeros(10), (0.3, 0.6); g = B.InvSumSq(2.0, 0.4), f = B.FAllenCahn(-1.0))
        @test rel(solve(prob, MLP(); multithreading = false).us[end], 0.30879) < 0.5
        # Allen–Cahn with Neumann reflection, BASELINE C2 (test/MLP.jl:162-186 at d = 100): 0.2037 ± 0.004 per solve
        d = 100
        prob = ParabolicPDEProblem(B.ConstCoef(0.0), B.ConstCoef(1.0), zeros(d), (0.0, 0.5); g = B.ExpSumSq(0.25), f = B.FAllenCahn(1.0),
                                   neumann_bc = [fill(-0.5, d), fill(0.5, d)])
        @test abs(solve(prob, MLP(); multithreading = false).us[end] - 0.2037) < 0.03
        # Black–Scholes with default risk, d = 20 (test/MLP.jl:188-230)
        prob = ParabolicPDEProblem(B.LinearCoef(0.02), B.LinearCoef(0.2), fill(100.0, 20), (0.0, 1.0); g = B.MinCoord(),
                                   f = B.FDefaultRisk(2 / 3, 0.02, 50.0, 70.0, 0.2, 0.02))
        @test rel(solve(prob, MLP(); multithreading = false).us[end], 60.781) < 0.1
        # replicator–mutator, non-local, NormalSampling (test/MLP.jl:236-286), d = 1: analytic 3.2659 at the test's tolerance
        ss0 = 1e-2
        prob = PIDEProblem(B.ConstCoef(0.0), B.ConstCoef(0.1), zeros(1), (0.0, 0.5), B.GaussDensity(ss0), B.FNLReplicatorMutator(1.0))
        @test rel(solve(prob, MLP(M = 4, L = 4, K = 10, mc_sample = NormalSampling(1.0)); multithreading = false).us[end], 3.2659) < 0.1
        # non-local with reflection, run-to-run agreement (test/MLP.jl:288-317)
        prob = PIDEProblem(B.ConstCoef(0.0), B.ConstCoef(0.1), zeros(5), (0.0, 0.5), B.ExpSumSq(0.25), B.FNLAllenCahnDiff();
                           neumann_bc = [fill(-0.5, 5), fill(0.5, 5)])
        alg = MLP(M = 4, L = 4, K = 10, mc_sample = UniformSampling(fill(-0.5, 5), fill(0.5, 5)))
        @test rel(solve(prob, alg; multithreading = false).us[end], solve(prob, alg; multithreading = false).us[end]) < 0.1
        # MLP requires x0_sample = NoSampling (src/MLP.jl:65-66): the error is the reference's, raised before the glue
        prob = ParabolicPDEProblem(B.ConstCoef(0.0), B.ConstCoef(0.1), zeros(3), (0.0, 0.5); g = B.SumSq(0.0), f = B.FZero(),
                                   x0_sample = UniformSampling(fill(-1.0, 3), fill(1.0, 3)))
        @test_throws ErrorException solve(prob, MLP())
        # a user-defined g through NVRTC next to its Julia closure
        prob = ParabolicPDEProblem(B.ConstCoef(0.0), B.ConstCoef(0.1), fill(2.0, d), (0.0, 0.5);
                                   g = B.UserG("ssq + p[0]", (1.0, 0.0, 0.0, 0.0), x -> sum(x .^ 2) + 1.0), f = B.FZero())
        @test rel(solve(prob, MLP(); multithreading = false).us[end], 4.0 * d + 1.0 + d * 0.01 * 0.5) < 0.1
    end

    @testset "golden tapes through the glue" begin
        G1 = npzread(joinpath(GOLD, "golden_v1.npz"))
        cases = Dict(
            "heat_d5" => (ParabolicPDEProblem(B.ConstCoef(0.0), B.ConstCoef(0.1), fill(2.0, 5), (0.0, 0.5); g = B.SumSq(0.0), f = B.FZero()),
                          MLP(M = 3, L = 3, K = 2)),
            "allen_cahn_neumann_d6" => (ParabolicPDEProblem(B.ConstCoef(0.0), B.ConstCoef(1.0), zeros(6), (0.0, 0.5); g = B.ExpSumSq(0.25),
                                                             f = B.FAllenCahn(1.0), neumann_bc = [fill(-0.5, 6), fill(0.5, 6)]),
                                        MLP(M = 3, L = 3, K = 2)),
            "black_scholes_d4" => (ParabolicPDEProblem(B.LinearCoef(0.02), B.LinearCoef(0.2), fill(100.0, 4), (0.0, 1.0); g = B.MinCoord(),
                                                        f = B.FDefaultRisk(2 / 3, 0.02, 50.0, 70.0, 0.2, 0.02)), MLP(M = 3, L = 3, K = 2)),
            "nonlocal_uniform_neumann_d3" => (PIDEProblem(B.ConstCoef(0.0), B.ConstCoef(0.1), zeros(3), (0.0, 0.5), B.ExpSumSq(0.25),
                                                           B.FNLAllenCahnDiff(); neumann_bc = [fill(-0.5, 3), fill(0.5, 3)]),
                                              MLP(M = 2, L = 3, K = 2, mc_sample = UniformSampling(fill(-0.5, 3), fill(0.5, 3)))),
            "replicator_normal_d2" => (PIDEProblem(B.ConstCoef(0.0), B.ConstCoef(0.1), zeros(2), (0.0, 0.5), B.GaussDensity(1e-2),
                                                    B.FNLReplicatorMutator(1.0)), MLP(M = 2, L = 3, K = 2, mc_sample = NormalSampling(1.0))),
            "fisher_shifted_normal_d3" => (PIDEProblem(B.ConstCoef(0.0), B.ConstCoef(0.1), zeros(3), (0.0, 0.5), B.ExpSumSq(1.0), B.FNLFisher()),
                                           MLP(M = 2, L = 3, K = 2, mc_sample = NormalSampling(0.2, true))))
        for (name, (prob, alg)) in cases
            s = haskey(G1, "mlp_$(name)_s") ? permutedims(G1["mlp_$(name)_s"]) : nothing
            u = B.mlp_b200(alg.M, alg.L, alg.K, prob.x, alg.mc_sample!, prob, prob.neumann_bc;
                           tapes = (G1["mlp_$(name)_r"], permutedims(G1["mlp_$(name)_n"]), s))
            @test rel(u, G1["mlp_$(name)_u"][1]) <= 1e-12
        end
        # `_reflect`, Matrix method, above the size where the overlay takes over
        rng = MersenneTwister(1)
        d, Bn = 10, 4096
        a = rand(rng, d, Bn) .- 0.5
        b = a .+ randn(rng, d, Bn)
        got = HighDimPDE._reflect(a, b, -0.5, 0.5)
        want = invoke(HighDimPDE._reflect, Tuple{T, T, Any, Any} where {T <: AbstractArray}, a, b, -0.5, 0.5)
        @test maximum(abs.(got .- want)) <= 1e-12 && all(-0.5 .<= got .<= 0.5)
        # sde_loop! on the injected increments of the golden file
        for (tag, T) in (("f64", Float64), ("f32", Float32))
            dW = G1["sde_$(tag)_dW"]
            d, Bn, N = size(dW)
            prob = ParabolicPDEProblem(B.ConstCoef(T(0.05)), B.ConstCoef(T(1.0)), zeros(T, d), (T(0), T(1)); g = B.SumSq(T(0)), f = B.FZero())
            y1 = zeros(T, d, Bn); y0 = similar(y1)
            B.sde_loop_b200!(y0, y1, N, T(0.1), prob, (T(-1), T(1)), NoSampling(); dWall = dW)
            tol = T === Float64 ? 1e-12 : 1e-5
            @test maximum(abs.(y1 .- G1["sde_$(tag)_y1"])) <= tol && maximum(abs.(y0 .- G1["sde_$(tag)_y0"])) <= tol
        end
    end

    @testset "fall-back to the stock CPU path" begin
        # plain closures: `accelerated(prob)` is false for g, so dispatch never reaches the overlay
        prob = ParabolicPDEProblem((x, p, t) -> 0.0, (x, p, t) -> 0.1, fill(2.0, 4), (0.0, 0.5); g = x -> sum(x .^ 2), f = (y, v, ∇v, p, t) -> 0.0 .* v)
        @test rel(solve(prob, MLP(M = 3, L = 3, K = 2); multithreading = false).us[end], 16.0 + 4 * 0.01 * 0.5) < 0.2
        # a kind g with a closure f: the overlay is entered, finds the problem not accelerated, and `invoke`s the reference
        prob = ParabolicPDEProblem(B.ConstCoef(0.0), B.ConstCoef(0.1), fill(2.0, 4), (0.0, 0.5); g = B.SumSq(0.0), f = (y, v, ∇v, p, t) -> 0.0 .* v)
        @test rel(solve(prob, MLP(M = 3, L = 3, K = 2); multithreading = true).us[end], 16.02) < 0.2
    end
end
