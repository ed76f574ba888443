# HighDimPDEB200.jl — `ccall` glue between HighDimPDE.jl and libhdpde_b200.so.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia toolchain.
# The Python package `highdimpde.jl_b200/` is the 1:1 twin of this file (same
# descriptors, same entry points, same argument order) and IS exercised by
# `tests/`; keep the two in step.  See INTEGRATION.md.
#
# What it does: user-visible `solve(prob, MLP(...))` is unchanged.  When μ, σ, g, f
# are the callable *kind* structs below (ordinary Julia callables, so the stock CPU
# path keeps working), `HighDimPDE._ml_picard` / `_ml_picard_mlt`
# (src/MLP.jl:102-260) are overlaid by one `ccall` into `hdp_mlp_solve_f64/_f32`;
# `_reflect` (src/reflect.jl) and the samplers' call operators (src/MCSample.jl)
# get `ccall` methods for `Array`/`CuArray` arguments, and `sde_loop_b200!`
# replaces the `sde_loop!` closure of src/DeepSplitting.jl:153-166.
module HighDimPDEB200

using HighDimPDE
import HighDimPDE: _ml_picard, _ml_picard_mlt, _reflect, UniformSampling, NormalSampling, NoSampling

const libhdp = get(ENV, "HDPDE_B200_LIB", joinpath(@__DIR__, "..", "highdimpde.jl_b200", "libhdpde_b200.so"))

# ------------------------------------------------------------------ POD mirrors of include/hdpde.h
struct HdpSampler
    kind::Int32; shifted::Int32; bounds_are_scalar::Int32; _pad::Int32
    a_scalar::Float64; b_scalar::Float64; sigma::Float64
    a::Ptr{Float64}; b::Ptr{Float64}
end
struct HdpProblem
    d::Int32; mu_kind::Int32; sigma_kind::Int32; g_kind::Int32; f_kind::Int32
    has_bc::Int32; bc_is_scalar::Int32; _pad::Int32
    t0::Float64; t1::Float64
    mu_p::NTuple{2, Float64}; sigma_p::NTuple{2, Float64}
    g_p::NTuple{4, Float64}; f_p::NTuple{8, Float64}
    lo_scalar::Float64; hi_scalar::Float64
    lo::Ptr{Float64}; hi::Ptr{Float64}
end
struct HdpMlpCfg
    M::Int32; L::Int32; K::Int32; n_solves::Int32; n_x0::Int32; normal_mode::Int32
    compat_mlt_threads::Int32; target_units::Int32
    seed::UInt64
    sampler::HdpSampler
    r_tape::Ptr{Cvoid}; n_tape::Ptr{Cvoid}; s_tape::Ptr{Cvoid}
    user::Ptr{Cvoid}             # Ptr{HdpUserExprs} for the *_USER kinds, else C_NULL
end
struct HdpUserExprs                # CUDA-C expressions compiled by NVRTC (include/hdpde.h: hdp_user_exprs)
    mu::Cstring; sigma::Cstring; g::Cstring; f::Cstring
end

# ------------------------------------------------------------------ kinds: callable structs
# μ, σ (test/MLP.jl:24-25, 191-192)
struct ConstCoef{T}; c::T; end;   (k::ConstCoef)(x, p, t) = k.c
struct LinearCoef{T}; c::T; end;  (k::LinearCoef)(x, p, t) = k.c * x
struct QuadraticCoef{T}; c::T; end;  (k::QuadraticCoef)(x, p, t) = k.c * x .^ 2   # test/DeepSplitting.jl:355
# any other μ/σ: a CUDA-C expression over x, i, t, d, p[0..1] (compiled by NVRTC on first use) next to
# the Julia closure that the host-side code paths keep calling
struct UserCoef{F}; expr::String; p::NTuple{2, Float64}; fn::F; end;  (k::UserCoef)(x, p, t) = k.fn(x, p, t)
# g (test/MLP.jl:28,118,146,174,199,265)
struct SumSq{T}; c::T; end;            (k::SumSq)(x) = sum(x .^ 2) .+ k.c
struct InvSumSq{T}; a::T; b::T; end;   (k::InvSumSq)(x) = 1 ./ (k.a .+ k.b * sum(x .^ 2))
struct ExpSumSq{T}; c::T; end;         (k::ExpSumSq)(x) = exp.(-k.c * sum(x .^ 2))
struct MinCoord end;                   (k::MinCoord)(x) = minimum(x)
struct GaussDensity{T}; ss0::T; end
(k::GaussDensity)(x) = (2π)^(-length(x) / 2) * k.ss0^(-length(x) * 0.5) * exp.(-0.5 * sum(x .^ 2 / k.ss0))
# f, 5-argument local form (test/MLP.jl:35,122,148,176,209-215; docs/src/tutorials/mlp.md:20)
struct FZero end;                 (k::FZero)(y, v, ∇v, p, t) = 0 .* v
struct FLinear{T}; r::T; end;     (k::FLinear)(y, v, ∇v, p, t) = k.r * v
struct FAllenCahn{T}; c::T; end;  (k::FAllenCahn)(y, v, ∇v, p, t) = k.c .* (v .- v .^ 3)
struct FFisherKPP end;            (k::FFisherKPP)(y, v, ∇v, p, t) = max(0, v) * (1 - max(0, v))
struct FDefaultRisk{T}; δ::T; R::T; vh::T; vl::T; γh::T; γl::T; end
function (k::FDefaultRisk)(y, v, ∇v, p, t)
    Q(u) = (u < k.vh) * k.γh + (u >= k.vl) * k.γl +
           ((u >= k.vh) * (u < k.vl)) * (((k.γh - k.γl) / (k.vh - k.vl)) * (u - k.vh) + k.γh)
    -(1 - k.δ) * Q.(v) .* v .- k.R * v
end
# f, 8-argument non-local form (test/MLP.jl:268-274, 304; docs/src/tutorials/mlp.md:52)
struct FNLAllenCahnDiff end
(k::FNLAllenCahnDiff)(y, z, vy, vz, ∇vy, ∇vz, p, t) = (vy .- vy .^ 3) .- (vz .- vz .^ 3)
struct FNLFisher end
(k::FNLFisher)(y, z, vy, vz, ∇vy, ∇vz, p, t) = max(0, vy) * (1 - max(0, vz))
struct FNLReplicatorMutator{T}; σs::T; end
function (k::FNLReplicatorMutator)(y, z, vy, vz, ∇vy, ∇vz, p, t)
    d = length(y); m(x) = -0.5 * sum(x .^ 2)
    max(0, vy) * (m(y) - max(0, vz) * m(z) * (2π)^(d / 2) * k.σs^d * exp(0.5 * sum(z .^ 2) / k.σs^2))
end

kind_of(k::ConstCoef) = (Int32(0), (Float64(k.c), 0.0));  kind_of(k::LinearCoef) = (Int32(1), (Float64(k.c), 0.0))
kind_of(k::UserCoef) = (Int32(2), k.p);                    kind_of(k::QuadraticCoef) = (Int32(3), (Float64(k.c), 0.0))
gkind(k::SumSq) = (Int32(0), (Float64(k.c), 0.0, 0.0, 0.0))
gkind(k::InvSumSq) = (Int32(1), (Float64(k.a), Float64(k.b), 0.0, 0.0))
gkind(k::ExpSumSq) = (Int32(2), (Float64(k.c), 0.0, 0.0, 0.0))
gkind(::MinCoord) = (Int32(3), (0.0, 0.0, 0.0, 0.0))
gkind(k::GaussDensity) = (Int32(4), (Float64(k.ss0), 0.0, 0.0, 0.0))
_p8(v...) = ntuple(i -> i <= length(v) ? Float64(v[i]) : 0.0, 8)
fkind(::FZero) = (Int32(0), _p8());            fkind(k::FLinear) = (Int32(1), _p8(k.r))
fkind(k::FAllenCahn) = (Int32(2), _p8(k.c));   fkind(::FFisherKPP) = (Int32(4), _p8())
fkind(k::FDefaultRisk) = (Int32(3), _p8(k.δ, k.R, k.vh, k.vl, k.γh, k.γl))
fkind(::FNLAllenCahnDiff) = (Int32(16), _p8()); fkind(::FNLFisher) = (Int32(17), _p8())
fkind(k::FNLReplicatorMutator) = (Int32(18), _p8(k.σs))

const B200Coef = Union{ConstCoef, LinearCoef, QuadraticCoef, UserCoef}
const B200G = Union{SumSq, InvSumSq, ExpSumSq, MinCoord, GaussDensity}
const B200F = Union{FZero, FLinear, FAllenCahn, FFisherKPP, FDefaultRisk, FNLAllenCahnDiff, FNLFisher, FNLReplicatorMutator}

# ------------------------------------------------------------------ handle
mutable struct Handle
    ptr::Ptr{Cvoid}
    function Handle(device::Integer = -1)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:hdp_create, libhdp), Cint, (Ref{Ptr{Cvoid}}, Cint), out, device)
        rc == 0 || error("hdp_create: " * unsafe_string(ccall((:hdp_last_error, libhdp), Cstring, (Ptr{Cvoid},), C_NULL)))
        h = new(out[])
        finalizer(x -> ccall((:hdp_destroy, libhdp), Cint, (Ptr{Cvoid},), x.ptr), h)
    end
end
const _handles = Dict{Int, Handle}()          # one handle per Julia thread
handle() = get!(() -> Handle(), _handles, Threads.threadid())

function check(h::Handle, rc::Cint)
    rc == 0 && return
    msg = unsafe_string(ccall((:hdp_last_error, libhdp), Cstring, (Ptr{Cvoid},), h.ptr))
    rc == -2 && error("a not in hypercube: " * msg)            # src/reflect.jl:12
    rc == -3 && error("`MLP` algorithm can only be used with `x0_sample=NoSampling()`.")  # src/MLP.jl:65-66
    error("libhdpde_b200 error $rc: $msg")
end

# ------------------------------------------------------------------ descriptors
sampler_desc(::NoSampling, keep) = HdpSampler(0, 0, 1, 0, 0.0, 0.0, 0.0, C_NULL, C_NULL)
function sampler_desc(s::UniformSampling, keep)
    if s.a isa Number
        return HdpSampler(1, 0, 1, 0, Float64(s.a), Float64(s.b), 0.0, C_NULL, C_NULL)
    end
    a = Vector{Float64}(s.a); b = Vector{Float64}(s.b); push!(keep, a, b)
    HdpSampler(1, 0, 0, 0, 0.0, 0.0, 0.0, pointer(a), pointer(b))
end
sampler_desc(s::NormalSampling, keep) = HdpSampler(2, s.shifted, 1, 0, 0.0, 0.0, Float64(s.σ), C_NULL, C_NULL)

function problem_desc(prob, f, bc, keep; g = prob.g, d = length(prob.x))
    μk, μp = kind_of(prob.μ); σk, σp = kind_of(prob.σ)
    gk, gp = gkind(g);        fk, fp = fkind(f)
    d = Int32(d)
    if isnothing(bc)
        has, scal, lo_s, hi_s, lo, hi = 0, 1, 0.0, 0.0, C_NULL, C_NULL
    elseif bc[1] isa Number
        has, scal, lo_s, hi_s, lo, hi = 1, 1, Float64(bc[1]), Float64(bc[2]), C_NULL, C_NULL
    else
        l = Vector{Float64}(bc[1]); u = Vector{Float64}(bc[2]); push!(keep, l, u)
        has, scal, lo_s, hi_s, lo, hi = 1, 0, 0.0, 0.0, pointer(l), pointer(u)
    end
    HdpProblem(d, μk, σk, gk, fk, has, scal, 0, Float64(prob.tspan[1]), Float64(prob.tspan[2]),
               μp, σp, gp, fp, lo_s, hi_s, lo, hi)
end

# ------------------------------------------------------------------ MLP overlay (src/MLP.jl:68-98)
accelerated(prob) = prob.μ isa B200Coef && prob.σ isa B200Coef && prob.g isa B200G && prob.f isa B200F

function _mlp_b200(M, L, K, x::Vector{T}, mc_sample!, prob, neumann_bc, compat_threads) where {T <: Union{Float32, Float64}}
    keep = Any[]
    pb = problem_desc(prob, prob.f, neumann_bc, keep)
    cfg = HdpMlpCfg(M, L, K, 1, 1, 0, compat_threads, 0, rand(UInt64), sampler_desc(mc_sample!, keep),
                    C_NULL, C_NULL, C_NULL, C_NULL)
    u = Vector{T}(undef, 1)
    h = handle()
    sym = T === Float64 ? :hdp_mlp_solve_f64 : :hdp_mlp_solve_f32
    GC.@preserve keep x u begin
        rc = ccall((sym, libhdp), Cint,
                   (Ptr{Cvoid}, Ref{HdpProblem}, Ref{HdpMlpCfg}, Ptr{T}, Ptr{T}, Ptr{Cvoid}),
                   h.ptr, pb, cfg, x, u, C_NULL)
    end
    check(h, rc)
    u[1]
end

# `_ml_picard(M, L, K, x, s, t, mc_sample!, g, f, verbose, prob, neumann_bc)` at the top of the
# recursion (s == tspan[1], t == tspan[2]) — the calls made by solve(::MLP), src/MLP.jl:68-98.
function _ml_picard(M, L, K, x::Vector{T}, s, t, mc_sample!, g::B200G, f, verbose, prob, neumann_bc) where {T}
    if accelerated(prob) && s == prob.tspan[1] && t == prob.tspan[2]
        return _mlp_b200(M, L, K, x, mc_sample!, prob, neumann_bc, 0)
    end
    invoke(_ml_picard, Tuple{Any, Any, Any, typeof(x), Any, Any, Any, Any, Any, Any, Any, Any},
           M, L, K, x, s, t, mc_sample!, g, f, verbose, prob, neumann_bc)
end
function _ml_picard_mlt(M, L, K, x::Vector{T}, s, t, mc_sample!, g::B200G, f, verbose, prob, neumann_bc) where {T}
    accelerated(prob) || return invoke(_ml_picard_mlt, Tuple{Any, Any, Any, typeof(x), Any, Any, Any, Any, Any, Any, Any, Any},
                                       M, L, K, x, s, t, mc_sample!, g, f, verbose, prob, neumann_bc)
    # reproduce the `_get_loop_num` partition of the threaded estimator (src/MLP.jl:391-404)
    _mlp_b200(M, L, K, x, mc_sample!, prob, neumann_bc, Threads.nthreads())
end

# ------------------------------------------------------------------ _reflect (src/reflect.jl:8-82)
function _reflect(a::Array{T}, b::Array{T}, s, e) where {T <: Union{Float32, Float64}}
    size(a) == size(b) || error("a not same dim as b")
    d = size(a, 1); B = length(a) ÷ d
    out = similar(b)
    scalar = s isa Number
    lo = scalar ? T[s] : Vector{T}(s); hi = scalar ? T[e] : Vector{T}(e)
    h = handle()
    sym = T === Float64 ? :hdp_reflect_f64 : :hdp_reflect_f32
    rc = ccall((sym, libhdp), Cint,
               (Ptr{Cvoid}, Ptr{T}, Ptr{T}, Ptr{T}, Cint, Cint, Ptr{T}, Ptr{T}, Cint, Ptr{Int32}, Cuint),
               h.ptr, a, b, out, d, B, lo, hi, scalar, C_NULL, 0)
    check(h, rc)
    out
end

# ------------------------------------------------------------------ samplers (src/MCSample.jl:28-33, 63-71)
function _mc_sample!(s, x_mc::Array{T}, x) where {T <: Union{Float32, Float64}}
    keep = Any[]
    d = size(x_mc, 1); B = size(x_mc, 2); K = length(x_mc) ÷ (d * B)
    h = handle()
    sym = T === Float64 ? :hdp_mc_sample_f64 : :hdp_mc_sample_f32
    GC.@preserve keep begin
        rc = ccall((sym, libhdp), Cint,
                   (Ptr{Cvoid}, Ptr{T}, Ptr{T}, Cint, Cint, Cint, Ref{HdpSampler}, UInt64, Ptr{T}, Cuint),
                   h.ptr, x_mc, isnothing(x) ? C_NULL : pointer(x), d, B, K, sampler_desc(s, keep), rand(UInt64), C_NULL, 0)
    end
    check(h, rc)
    x_mc
end
b200_sample!(s::UniformSampling, x_mc, x = nothing) = _mc_sample!(s, x_mc, nothing)
b200_sample!(s::NormalSampling, x_mc, x = nothing) = _mc_sample!(s, x_mc, s.shifted ? x : nothing)

# ------------------------------------------------------------------ sde_loop! (src/DeepSplitting.jl:153-166)
# y1 holds x0_batch on entry (:182).  With `z` and `mc_sample!` the call also performs
# `mc_sample!(z, y1)` (:186-189) in the same kernel.
function sde_loop_b200!(y0::Array{T}, y1::Array{T}, nsteps, dt, prob, neumann_bc, x0_sample!;
                        mc_sample! = NoSampling(), z = nothing) where {T <: Union{Float32, Float64}}
    keep = Any[]
    pb = problem_desc(prob, prob.f, neumann_bc, keep)
    K = isnothing(z) ? 0 : size(z, 3)
    h = handle()
    sym = T === Float64 ? :hdp_sde_loop_f64 : :hdp_sde_loop_f32
    GC.@preserve keep begin
        rc = ccall((sym, libhdp), Cint,
                   (Ptr{Cvoid}, Ref{HdpProblem}, Ptr{T}, Ptr{T}, Cint, Cint, T, Ptr{T}, Ref{HdpSampler}, UInt64, Ptr{T},
                    Ref{HdpSampler}, Ptr{T}, Cint, Cint, Cuint),
                   h.ptr, pb, y0, y1, size(y1, 2), nsteps, T(dt), C_NULL, sampler_desc(x0_sample!, keep), rand(UInt64),
                   C_NULL, sampler_desc(mc_sample!, keep), isnothing(z) ? C_NULL : pointer(z), K, 0, 0)
    end
    check(h, rc)
    nothing
end
# CuArray arguments: same calls with `mem_flags = HDP_MEM_DEVICE (1)` and `pointer(::CuArray)`.

# μ/σ expressions of UserCoef kinds for the K1 entry points (hdp_set_user_exprs copies the strings);
# `ts` must then be passed to hdp_sde_loop_* (t = ts[N+1-i], src/DeepSplitting.jl:157).
function set_user_coefs!(h::Handle, prob)
    μ = prob.μ isa UserCoef ? prob.μ.expr : nothing;  σ = prob.σ isa UserCoef ? prob.σ.expr : nothing
    if isnothing(μ) && isnothing(σ)
        return check(h, ccall((:hdp_set_user_exprs, libhdp), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), h.ptr, C_NULL))
    end
    GC.@preserve μ σ begin
        ue = HdpUserExprs(isnothing(μ) ? C_NULL : pointer(μ), isnothing(σ) ? C_NULL : pointer(σ), C_NULL, C_NULL)
        check(h, ccall((:hdp_set_user_exprs, libhdp), Cint, (Ptr{Cvoid}, Ref{HdpUserExprs}), h.ptr, ue))
    end
end

# ------------------------------------------------------------------ EM ensembles (src/NNKolmogorov.jl:81-112, src/NNStopping.jl:90-99)
# Stands in for `solve(EnsembleProblem(SDEProblem(μ, σ, x0, tspan)[, prob_func]), EM(); dt, trajectories,
# adaptive = false[, saveat = dt])`: x is d×trajectories (start points in, final states out); with
# `saveat = true` returns `Array(sim)` (d × N+1 × trajectories), else x.
function em_ensemble_b200!(x::Matrix{T}, prob, dt; saveat = false) where {T <: Union{Float32, Float64}}
    keep = Any[]
    d, B = size(x)                      # prob.x is `nothing` for NNKolmogorov problems (src/NNKolmogorov.jl:19)
    pb = problem_desc(prob, FZero(), nothing, keep; g = SumSq(zero(T)), d = d)
    N = Int(round((prob.tspan[2] - prob.tspan[1]) / dt))
    traj = saveat ? Array{T}(undef, d, N + 1, B) : nothing
    h = handle(); set_user_coefs!(h, prob)
    sym = T === Float64 ? :hdp_em_ensemble_f64 : :hdp_em_ensemble_f32
    GC.@preserve keep begin
        rc = ccall((sym, libhdp), Cint,
                   (Ptr{Cvoid}, Ref{HdpProblem}, Ptr{T}, Cint, Cint, T, T, Ptr{T}, UInt64, Ptr{T}, Cuint),
                   h.ptr, pb, x, B, N, T(dt), T(prob.tspan[1]), saveat ? pointer(traj) : C_NULL, rand(UInt64), C_NULL, 0)
    end
    check(h, rc)
    saveat ? traj : x
end

end # module
