# replay_reference.jl — pins this repository's CPU oracle (oracle/) to the STOCK reference.
#
# The build image has no Julia, so the parity chain of this repository ends at "oracle == two independent
# restatements of the reference source" (DESIGN.md §2: parity unpinned).  This script is the missing link for
# anyone who has Julia: it feeds the random tapes stored in tests/golden/golden_v1.npz / golden_v2.npz to the
# UNMODIFIED HighDimPDE.jl functions of the hot path
#     _ml_picard (src/MLP.jl:102-208) with `_mlt_sde_loop!` (:406-421) and the Vector `_reflect` (src/reflect.jl:8-55),
#     _reflect, both methods (src/reflect.jl:8-82),
#     UniformSampling / NormalSampling call operators (src/MCSample.jl:28-33, 63-71),
#     the `sde_loop!` closure of DeepSplitting (src/DeepSplitting.jl:153-166; transcribed below because a closure
#     cannot be called from outside — or `HighDimPDE._sde_loop!` when julia/patches/deepsplitting_sde_loop_hook.patch
#     is applied),
# by overriding `rand(T)`, `randn!`, `rand!` for plain Arrays with tape readers, and compares with the outputs the
# oracle produced on the same tapes: <= 1e-12 relative in Float64, <= 1e-5 in Float32 (north_star's tolerances).
#
#     julia --project=/path/to/HighDimPDE.jl -e 'using Pkg; Pkg.add("NPZ")'        # once
#     julia --project=/path/to/HighDimPDE.jl tests/golden/replay_reference.jl [dir with the .npz files]
#
# A pass turns every "oracle" in this repository's parity tests into "reference".  A failure names the array.
using HighDimPDE, NPZ, Random, Test, LinearAlgebra
import HighDimPDE: _ml_picard, _reflect, UniformSampling, NormalSampling, NoSampling, MCSampling

const DIR = length(ARGS) >= 1 ? ARGS[1] : @__DIR__
const G1 = npzread(joinpath(DIR, "golden_v1.npz"))
const G2 = npzread(joinpath(DIR, "golden_v2.npz"))

# ------------------------------------------------------------------ tapes in the reference's consumption order
mutable struct Tape
    data::Vector{Float64}
    pos::Int
end
Tape(a::AbstractArray) = Tape(Float64.(vec(a)), 0)
const TAPES = Dict{Symbol, Tape}()
const MODE = Ref(:n)                 # which tape `randn!` / `rand!` read: :n step normals, :s sampler draws, :dW
function take!(k::Symbol, n::Int)
    t = TAPES[k]
    t.pos + n <= length(t.data) || error("tape $k exhausted: the reference consumed more draws than the oracle")
    v = view(t.data, (t.pos + 1):(t.pos + n))
    t.pos += n
    v
end
# the three calls of the path (src/MLP.jl:132 `rand(tType)`; :416 and src/MCSample.jl:64, src/DeepSplitting.jl:154
# `randn!`; src/MCSample.jl:30 `rand!`).  Type piracy on purpose: this is a test harness.
Random.rand(::Type{T}) where {T <: Union{Float32, Float64}} = T(take!(:r, 1)[1])
function Random.randn!(A::Array{T}) where {T <: Union{Float32, Float64}}
    A .= reshape(T.(take!(MODE[], length(A))), size(A))   # column-major, like the oracle's tapes
end
function Random.rand!(A::Array{T}) where {T <: Union{Float32, Float64}}
    A .= reshape(T.(take!(:s, length(A))), size(A))
end
# a sampler that routes its draws to the sampler tape and then calls the STOCK sampler
struct TapedSampler{T, S} <: MCSampling{T}
    inner::S
end
TapedSampler(s::MCSampling{T}) where {T} = TapedSampler{T, typeof(s)}(s)
function (s::TapedSampler)(x_mc, x...)
    MODE[] = :s
    r = s.inner(x_mc, x...)
    MODE[] = :n
    r
end
# the oracle tapes are [ordinal][coordinate] row-major = column-major d×n: flatten in storage order
tape_of(a) = ndims(a) == 2 ? Tape(permutedims(a)) : Tape(a)

rel(a, b) = abs(a - b) / max(abs(b), floatmin(Float64))

# ------------------------------------------------------------------ MLP problems of make_golden.py, with closures
function mlp_cases()
    c = Dict{String, Any}()
    heat_g(x) = sum(x .^ 2)
    zero_f(y, v, ∇v, p, t) = 0 .* v
    c["heat_d5"] = (ParabolicPDEProblem((x, p, t) -> 0.0, (x, p, t) -> 0.1, fill(2.0, 5), (0.0, 0.5); g = heat_g, f = zero_f),
                    MLP(M = 3, L = 3, K = 2))
    ac_f(y, v, ∇v, p, t) = 1.0 .* (v .- v .^ 3)
    c["allen_cahn_neumann_d6"] = (ParabolicPDEProblem((x, p, t) -> 0.0, (x, p, t) -> 1.0, zeros(6), (0.0, 0.5);
                                                       g = x -> exp(-0.25 * sum(x .^ 2)), f = ac_f,
                                                       neumann_bc = [fill(-0.5, 6), fill(0.5, 6)]),
                                  MLP(M = 3, L = 3, K = 2))
    δ, R, vh, vl, γh, γl = 2 / 3, 0.02, 50.0, 70.0, 0.2, 0.02           # test/MLP.jl:201-215
    Q(u) = (u < vh) * γh + (u >= vl) * γl + ((u >= vh) * (u < vl)) * (((γh - γl) / (vh - vl)) * (u - vh) + γh)
    c["black_scholes_d4"] = (ParabolicPDEProblem((x, p, t) -> 0.02 * x, (x, p, t) -> 0.2 * x, fill(100.0, 4), (0.0, 1.0);
                                                  g = x -> minimum(x), f = (y, v, ∇v, p, t) -> -(1 - δ) * Q(v) * v - R * v),
                             MLP(M = 3, L = 3, K = 2))
    a(u) = u - u^3
    c["nonlocal_uniform_neumann_d3"] = (PIDEProblem((x, p, t) -> 0.0, (x, p, t) -> 0.1, zeros(3), (0.0, 0.5),
                                                     x -> exp(-0.25 * sum(x .^ 2)),
                                                     (y, z, vy, vz, ∇vy, ∇vz, p, t) -> a(vy) - a(vz);
                                                     neumann_bc = [fill(-0.5, 3), fill(0.5, 3)]),
                                        MLP(M = 2, L = 3, K = 2, mc_sample = UniformSampling(fill(-0.5, 3), fill(0.5, 3))))
    ss0, σs = 1e-2, 1.0                                                   # test/MLP.jl:236-286
    gd(x) = (2π)^(-length(x) / 2) * ss0^(-length(x) * 0.5) * exp(-0.5 * sum(x .^ 2 / ss0))
    m(x) = -0.5 * sum(x .^ 2)
    frm(y, z, vy, vz, ∇vy, ∇vz, p, t) = max(0.0, vy) * (m(y) - max(0.0, vz) * m(z) * (2π)^(length(y) / 2) * σs^length(y) *
                                                          exp(0.5 * sum(z .^ 2) / σs^2))
    c["replicator_normal_d2"] = (PIDEProblem((x, p, t) -> 0.0, (x, p, t) -> 0.1, zeros(2), (0.0, 0.5), gd, frm),
                                 MLP(M = 2, L = 3, K = 2, mc_sample = NormalSampling(1.0)))
    c["fisher_shifted_normal_d3"] = (PIDEProblem((x, p, t) -> 0.0, (x, p, t) -> 0.1, zeros(3), (0.0, 0.5),
                                                  x -> exp(-1.0 * sum(x .^ 2)),
                                                  (y, z, vy, vz, ∇vy, ∇vz, p, t) -> max(0.0, vy) * (1 - max(0.0, vz))),
                                     MLP(M = 2, L = 3, K = 2, mc_sample = NormalSampling(0.2, true)))
    c["heat_neumann_f32_d3"] = (ParabolicPDEProblem((x, p, t) -> 0.0f0, (x, p, t) -> 0.1f0, fill(0.3f0, 3), (0.0f0, 0.5f0);
                                                     g = heat_g, f = zero_f, neumann_bc = [fill(-0.5f0, 3), fill(0.5f0, 3)]),
                                MLP(M = 3, L = 3, K = 2))
    c
end

@testset "oracle == stock HighDimPDE.jl on the golden tapes" begin
    @testset "_ml_picard $name" for (name, (prob, alg)) in mlp_cases()
        TAPES[:r] = tape_of(G1["mlp_$(name)_r"])
        TAPES[:n] = tape_of(G1["mlp_$(name)_n"])
        haskey(G1, "mlp_$(name)_s") && (TAPES[:s] = tape_of(G1["mlp_$(name)_s"]))
        MODE[] = :n
        mc = alg.mc_sample! isa NoSampling ? alg.mc_sample! : TapedSampler(alg.mc_sample!)
        f = prob isa ParabolicPDEProblem ? ((y, z, vy, vz, ∇vy, ∇vz, p, t) -> prob.f(y, vy, ∇vy, p, t)) : prob.f  # src/MLP.jl:59-63
        u = _ml_picard(alg.M, alg.L, alg.K, prob.x, prob.tspan[1], prob.tspan[2], mc, prob.g, f, false, prob, prob.neumann_bc)
        want = G1["mlp_$(name)_u"][1]
        tol = eltype(prob.x) === Float32 ? 1e-5 : 1e-12
        @test rel(Float64(u), Float64(want)) <= tol
        # the reference consumed exactly the tapes the oracle consumed ("bit-exact indexing")
        @test TAPES[:r].pos == length(TAPES[:r].data) && TAPES[:n].pos == length(TAPES[:n].data)
        haskey(G1, "mlp_$(name)_s") && @test TAPES[:s].pos == length(TAPES[:s].data)
    end

    @testset "_reflect $tag" for (tag, T) in (("f64", Float64), ("f32", Float32))
        a, b, lo, hi = G1["reflect_$(tag)_a"], G1["reflect_$(tag)_b"], G1["reflect_$(tag)_lo"], G1["reflect_$(tag)_hi"]
        tol = T === Float64 ? 1e-12 : 4e-6
        @test maximum(abs.(_reflect(a, b, lo, hi) .- G1["reflect_$(tag)_array"])) <= tol        # src/reflect.jl:58-82
        for j in 1:size(a, 2)                                                                      # src/reflect.jl:8-55
            @test maximum(abs.(_reflect(a[:, j], b[:, j], lo, hi) .- G1["reflect_$(tag)_vector"][:, j])) <= tol
        end
    end

    @testset "samplers" begin
        y1 = G1["mc_y1"]
        TAPES[:s] = Tape(G1["mc_raw_u"]); MODE[] = :s
        z = similar(G1["mc_raw_u"])
        UniformSampling(G1["mc_a"], G1["mc_b"])(z, y1)
        @test z == G1["mc_uniform"]                      # same IEEE operations in the same order: bit-equal
        for (shifted, key) in ((false, "mc_normal"), (true, "mc_normal_shifted"))
            TAPES[:s] = Tape(G1["mc_raw_n"]); MODE[] = :s
            z = similar(G1["mc_raw_n"])
            NormalSampling(0.3, shifted)(z, y1)
            @test z == G1[key]
        end
        MODE[] = :n
    end

    # the Euler–Maruyama loop of DeepSplitting.  With the hook patch applied the stock method is called; otherwise
    # the closure body is transcribed here (src/DeepSplitting.jl:153-166) and drives the stock `_reflect`.
    function sde_loop_ref!(y0, y1, dWall, μ, σ, p, dt, ts, N, neumann_bc, x0_sample!)
        isdefined(HighDimPDE, :_sde_loop!) && return HighDimPDE._sde_loop!(y0, y1, dWall, μ, σ, p, dt, ts, N, neumann_bc, x0_sample!)
        randn!(dWall)
        x0_sample!(y1)
        for i in 1:size(dWall, 3)
            t = ts[N + 1 - i]
            dW = @view dWall[:, :, i]
            y0 .= y1
            y1 .= y0 .+ μ(y0, p, t) .* dt .+ σ(y0, p, t) .* sqrt(dt) .* dW
            if !isnothing(neumann_bc)
                y1 .= _reflect(y0, y1, neumann_bc[1], neumann_bc[2])
            end
        end
        return
    end
    function run_sde(G, key, T, μ, σ, bc, start)
        dW = G["$(key)_dW"]
        d, B, N = size(dW)
        TAPES[:dW] = Tape(dW); MODE[] = :dW
        y1 = start === nothing ? zeros(T, d, B) : copy(start)
        y0 = similar(y1)
        dt = T(0.1)
        ts = T(0):(dt - eps(T)):T(1)                                                  # src/DeepSplitting.jl:121
        sde_loop_ref!(y0, y1, similar(dW), μ, σ, nothing, dt, ts, length(ts) - 1, bc, NoSampling())
        MODE[] = :n
        tol = T === Float64 ? 1e-12 : 1e-5
        @test maximum(abs.(y1 .- G["$(key)_y1"])) <= tol
        @test maximum(abs.(y0 .- G["$(key)_y0"])) <= tol
    end
    @testset "sde_loop! $tag" for (tag, T) in (("f64", Float64), ("f32", Float32))
        run_sde(G1, "sde_$(tag)", T, (x, p, t) -> T(0.05), (x, p, t) -> T(1.0), (T(-1.0), T(1.0)), nothing)
        run_sde(G2, "sde_$(tag)_quad_box", T, (x, p, t) -> T(0.0), (x, p, t) -> T(0.4) .* x .^ 2, (T(-1.0), T(1.0)),
                G2["sde_$(tag)_quad_box_start"])
        run_sde(G2, "sde_$(tag)_black_scholes", T, (x, p, t) -> T(0.02) .* x, (x, p, t) -> T(0.2) .* x, nothing,
                G2["sde_$(tag)_black_scholes_start"])
    end
end
