# HighDimPDEB200.jl — `ccall` glue between HighDimPDE.jl and libhdpde_b200.so.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia toolchain.
# The Python package `highdimpde.jl_b200/` is the 1:1 twin of this file (same
# descriptors, same entry points, same argument order) and IS exercised by
# `tests/`; keep the two in step.  `julia/test/runtests.jl` is what a maintainer
# with Julia runs; INTEGRATION.md lists every method added here next to the
# reference method it sits beside, and why the pair is unambiguous.
#
# What it does: user-visible `solve(prob, MLP(...))` is unchanged.  When μ, σ, g, f
# are the callable *kind* structs below (ordinary Julia callables, so the stock CPU
# path keeps working), `HighDimPDE._ml_picard` / `_ml_picard_mlt`
# (src/MLP.jl:102-260) are overlaid by one `ccall` into `hdp_mlp_solve_f64/_f32`.
# `solve(prob, DeepSplitting(...), dt)`: the Euler–Maruyama loop of that solver is a
# closure (src/DeepSplitting.jl:153-166) and cannot be overlaid from outside; with
# the two-line hook of `julia/patches/deepsplitting_sde_loop_hook.patch` applied it
# calls `HighDimPDE._sde_loop!`, for which this file (Array) and
# `ext/HighDimPDEB200CUDAExt.jl` (CuArray, no PCIe copy) add `ccall` methods.
# The Matrix method of `_reflect` (src/reflect.jl:58-82) gets a `ccall` method too.
module HighDimPDEB200

using HighDimPDE
using Random
import HighDimPDE: _ml_picard, _ml_picard_mlt, _reflect, UniformSampling, NormalSampling, NoSampling

const libhdp = get(ENV, "HDPDE_B200_LIB", joinpath(@__DIR__, "..", "highdimpde.jl_b200", "libhdpde_b200.so"))
const HDP_MEM_HOST, HDP_MEM_DEVICE, HDP_ASYNC = UInt32(0), UInt32(1), UInt32(2)
const HDP_MAX_LEVELS = 6

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
struct HdpUserExprs                # CUDA-C expressions compiled by NVRTC (include/hdpde.h: hdp_user_exprs)
    mu::Cstring; sigma::Cstring; g::Cstring; f::Cstring
end
struct HdpMlpCfg
    M::Int32; L::Int32; K::Int32; n_solves::Int32; n_x0::Int32; normal_mode::Int32
    compat_mlt_threads::Int32; target_units::Int32
    seed::UInt64
    sampler::HdpSampler
    r_tape::Ptr{Cvoid}; n_tape::Ptr{Cvoid}; s_tape::Ptr{Cvoid}
    user::Ptr{HdpUserExprs}      # for the *_USER kinds, else C_NULL
    group_size::Int32; mem_flags::UInt32
    unit_oversub::Float64
    std_err_out::Ptr{Cvoid}      # nullable: standard error per solve
end
struct HdpMlpCounts
    path_steps::UInt64; time_uniforms::UInt64; sampler_calls::UInt64; f_evals::UInt64; g_evals::UInt64
end
struct HdpMlpStats
    per_solve::HdpMlpCounts
    bounces::UInt64; n_units::UInt64; n_nodes::UInt64; n_sites::UInt64
    device_ms::Float64
    group_size::Int32; n_launches::Int32
    phase_ms::NTuple{4, Float64}
    level_n::NTuple{HDP_MAX_LEVELS + 1, UInt64}
    level_sum::NTuple{HDP_MAX_LEVELS + 1, Float64}
    level_sumsq::NTuple{HDP_MAX_LEVELS + 1, Float64}
    std_error::Float64
    exchange_bytes::UInt64
    exchange_kind::Int32; _pad::Int32   # 0 none, 1 ncclAllReduce, 2 exchange kernel over peer windows
end

# ------------------------------------------------------------------ kinds: callable structs
# μ, σ (test/MLP.jl:24-25, 191-192)
struct ConstCoef{T}; c::T; end;   (k::ConstCoef)(x, p, t) = k.c
struct LinearCoef{T}; c::T; end;  (k::LinearCoef)(x, p, t) = k.c * x
struct QuadraticCoef{T}; c::T; end;  (k::QuadraticCoef)(x, p, t) = k.c * x .^ 2   # test/DeepSplitting.jl:355
# Anything else: a CUDA-C expression (compiled by NVRTC on first use) next to the Julia closure that the
# host-side code paths keep calling.  μ, σ read x, i, t, d, p[0..1]; g reads ssq, mn, d, p[0..3];
# f reads vy, vz, ssq_y, ssq_z, t, d, p[0..7] (include/hdpde.h, hdp_user_exprs).
struct UserCoef{F}; expr::String; p::NTuple{2, Float64}; fn::F; end;  (k::UserCoef)(x, p, t) = k.fn(x, p, t)
struct UserG{F}; expr::String; p::NTuple{4, Float64}; fn::F; end;     (k::UserG)(x) = k.fn(x)
struct UserF{F}; expr::String; p::NTuple{8, Float64}; fn::F; end;     (k::UserF)(y, v, ∇v, p, t) = k.fn(y, v, ∇v, p, t)
struct UserFNL{F}; expr::String; p::NTuple{8, Float64}; fn::F; end
(k::UserFNL)(y, z, vy, vz, ∇vy, ∇vz, p, t) = k.fn(y, z, vy, vz, ∇vy, ∇vz, p, t)
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
gkind(k::UserG) = (Int32(64), k.p)
_p8(v...) = ntuple(i -> i <= length(v) ? Float64(v[i]) : 0.0, 8)
fkind(::FZero) = (Int32(0), _p8());            fkind(k::FLinear) = (Int32(1), _p8(k.r))
fkind(k::FAllenCahn) = (Int32(2), _p8(k.c));   fkind(::FFisherKPP) = (Int32(4), _p8())
fkind(k::FDefaultRisk) = (Int32(3), _p8(k.δ, k.R, k.vh, k.vl, k.γh, k.γl))
fkind(::FNLAllenCahnDiff) = (Int32(16), _p8()); fkind(::FNLFisher) = (Int32(17), _p8())
fkind(k::FNLReplicatorMutator) = (Int32(18), _p8(k.σs))
fkind(k::UserF) = (Int32(64), k.p);            fkind(k::UserFNL) = (Int32(65), k.p)

const B200Coef = Union{ConstCoef, LinearCoef, QuadraticCoef, UserCoef}
const B200G = Union{SumSq, InvSumSq, ExpSumSq, MinCoord, GaussDensity, UserG}
const B200F = Union{FZero, FLinear, FAllenCahn, FFisherKPP, FDefaultRisk, FNLAllenCahnDiff, FNLFisher,
                    FNLReplicatorMutator, UserF, UserFNL}
const HdpFloat = Union{Float32, Float64}

user_expr(k) = nothing
user_expr(k::Union{UserCoef, UserG, UserF, UserFNL}) = k.expr

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
# one handle per Julia thread; the table is shared, so it is only touched under a lock
const _handles = Dict{Int, Handle}()
const _handles_lock = ReentrantLock()
handle() = lock(() -> get!(() -> Handle(), _handles, Threads.threadid()), _handles_lock)
sync(h::Handle = handle()) = check(h, ccall((:hdp_sync, libhdp), Cint, (Ptr{Cvoid},), h.ptr))

function check(h::Handle, rc::Cint)
    rc == 0 && return
    msg = unsafe_string(ccall((:hdp_last_error, libhdp), Cstring, (Ptr{Cvoid},), h.ptr))
    rc == -2 && error("a not in hypercube: " * msg)            # src/reflect.jl:12
    rc == -3 && error("`MLP` algorithm can only be used with `x0_sample=NoSampling()`.")  # src/MLP.jl:65-66
    error("libhdpde_b200 error $rc: $msg")
end

# ------------------------------------------------------------------ one solve over several GPUs
# (include/hdpde.h: hdp_nccl_unique_id / hdp_comm_init_rank / hdp_peer_window_*; the bytes travel by whatever the
# host program uses between its processes — MPI.jl, Distributed.jl — see INTEGRATION.md §3)
function nccl_unique_id()
    id = Vector{UInt8}(undef, 128)
    rc = ccall((:hdp_nccl_unique_id, libhdp), Cint, (Ptr{UInt8},), id)
    rc == 0 || error("hdp_nccl_unique_id: " * unsafe_string(ccall((:hdp_last_error, libhdp), Cstring, (Ptr{Cvoid},), C_NULL)))
    id
end
comm_init_rank!(h::Handle, id::Vector{UInt8}, nranks::Integer, rank::Integer) =
    check(h, ccall((:hdp_comm_init_rank, libhdp), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint, Cint), h.ptr, id, nranks, rank))
"64-byte CUDA-IPC handle of this rank's exchange window (gather them in rank order, then `peer_window_attach!`)"
function peer_window_export(h::Handle)
    out = Vector{UInt8}(undef, 64)
    check(h, ccall((:hdp_peer_window_export, libhdp), Cint, (Ptr{Cvoid}, Ptr{UInt8}), h.ptr, out))
    out
end
"`handles`: nranks × 64 bytes in rank order.  Returns false (instead of throwing) when a peer cannot be mapped: the caller then detaches on every rank."
function peer_window_attach!(h::Handle, handles::Vector{UInt8}, nranks::Integer, rank::Integer)
    length(handles) == 64 * nranks || error("peer_window_attach!: expected $(64 * nranks) bytes")
    ccall((:hdp_peer_window_attach, libhdp), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint, Cint), h.ptr, handles, nranks, rank) == 0
end
peer_window_detach!(h::Handle) = check(h, ccall((:hdp_peer_window_detach, libhdp), Cint, (Ptr{Cvoid},), h.ptr))

# ------------------------------------------------------------------ descriptors
sampler_desc(::NoSampling, keep) = HdpSampler(0, 0, 1, 0, 0.0, 0.0, 0.0, C_NULL, C_NULL)
function sampler_desc(s::UniformSampling, keep)
    if s.a isa Number
        return HdpSampler(1, 0, 1, 0, Float64(s.a), Float64(s.b), 0.0, C_NULL, C_NULL)
    end
    a = Vector{Float64}(Array(s.a)); b = Vector{Float64}(Array(s.b)); push!(keep, a, b)
    HdpSampler(1, 0, 0, 0, 0.0, 0.0, 0.0, pointer(a), pointer(b))
end
sampler_desc(s::NormalSampling, keep) = HdpSampler(2, s.shifted, 1, 0, 0.0, 0.0, Float64(s.σ), C_NULL, C_NULL)

function problem_desc(μ, σ, g, f, tspan, bc, d, keep)
    μk, μp = kind_of(μ); σk, σp = kind_of(σ)
    gk, gp = gkind(g);   fk, fp = fkind(f)
    if isnothing(bc)
        has, scal, lo_s, hi_s, lo, hi = 0, 1, 0.0, 0.0, C_NULL, C_NULL
    elseif bc[1] isa Number
        has, scal, lo_s, hi_s, lo, hi = 1, 1, Float64(bc[1]), Float64(bc[2]), C_NULL, C_NULL
    else
        l = Vector{Float64}(Array(bc[1])); u = Vector{Float64}(Array(bc[2])); push!(keep, l, u)
        has, scal, lo_s, hi_s, lo, hi = 1, 0, 0.0, 0.0, pointer(l), pointer(u)
    end
    HdpProblem(Int32(d), μk, σk, gk, fk, has, scal, 0, Float64(tspan[1]), Float64(tspan[2]),
               μp, σp, gp, fp, lo_s, hi_s, lo, hi)
end
problem_desc(prob, f, bc, keep; g = prob.g, d = length(prob.x)) =
    problem_desc(prob.μ, prob.σ, g, f, prob.tspan, bc, d, keep)

# the expressions of the *_USER kinds as a C struct (strings kept alive in `keep`); C_NULL when none is used
function user_exprs_ref(μ, σ, g, f, keep)
    es = map(user_expr, (μ, σ, g, f))
    all(isnothing, es) && return nothing
    cs = map(e -> isnothing(e) ? nothing : Base.cconvert(Cstring, e), es)
    push!(keep, cs...)
    ptrs = map(c -> isnothing(c) ? Cstring(C_NULL) : Base.unsafe_convert(Cstring, c), cs)
    r = Ref(HdpUserExprs(ptrs...))
    push!(keep, r)
    r
end

# ------------------------------------------------------------------ MLP overlay (src/MLP.jl:68-98)
accelerated(prob) = prob.μ isa B200Coef && prob.σ isa B200Coef && prob.g isa B200G && prob.f isa B200F

"""
    mlp_b200(M, L, K, x, mc_sample!, prob, neumann_bc; compat_threads = 0, seed = rand(UInt64), std_err = false,
             tapes = nothing)

One `hdp_mlp_solve_*` call.  Returns `u`, or `(u, standard_error, stats)` with `std_err = true`.
`tapes = (r, n, s)`: injected random tapes in the reference's consumption order (parity mode; `s === nothing`
without a sampler) — `n` and `s` are d×count matrices or flat vectors, coordinate fastest.
"""
function mlp_b200(M, L, K, x::Vector{T}, mc_sample!, prob, neumann_bc; compat_threads = 0, seed = rand(UInt64),
                  std_err = false, tapes = nothing) where {T <: HdpFloat}
    keep = Any[]
    tr = tn = ts = Ptr{Cvoid}(C_NULL)
    if !isnothing(tapes)
        r, n, s = map(t -> isnothing(t) ? nothing : Vector{T}(vec(t)), tapes)
        push!(keep, r, n, s)
        tr, tn = Ptr{Cvoid}(pointer(r)), Ptr{Cvoid}(pointer(n))
        isnothing(s) || (ts = Ptr{Cvoid}(pointer(s)))
    end
    pb = problem_desc(prob, prob.f, neumann_bc, keep)
    ue = user_exprs_ref(prob.μ, prob.σ, prob.g, prob.f, keep)
    se = Vector{T}(undef, 1)
    u = Vector{T}(undef, 1)
    st = Ref{HdpMlpStats}()
    h = handle()
    sym = T === Float64 ? :hdp_mlp_solve_f64 : :hdp_mlp_solve_f32
    GC.@preserve keep x u se begin
        cfg = HdpMlpCfg(M, L, K, 1, 1, 0, compat_threads, 0, seed, sampler_desc(mc_sample!, keep),
                        tr, tn, ts,
                        isnothing(ue) ? Ptr{HdpUserExprs}(C_NULL) : Base.unsafe_convert(Ptr{HdpUserExprs}, ue),
                        0, HDP_MEM_HOST, 0.0, std_err ? Ptr{Cvoid}(pointer(se)) : C_NULL)
        rc = ccall((sym, libhdp), Cint,
                   (Ptr{Cvoid}, Ref{HdpProblem}, Ref{HdpMlpCfg}, Ptr{T}, Ptr{T}, Ref{HdpMlpStats}),
                   h.ptr, pb, cfg, x, u, st)
    end
    check(h, rc)
    std_err ? (u[1], se[1], st[]) : u[1]
end

# `_ml_picard(M, L, K, x, s, t, mc_sample!, g, f, verbose, prob, neumann_bc)` at the top of the
# recursion (s == tspan[1], t == tspan[2]) — the calls made by solve(::MLP), src/MLP.jl:68-98.
# Signature: `x::Vector{T}` and `g::B200G` are strictly narrower than the reference's untyped
# positional arguments (src/MLP.jl:102-115) and than its `x::Nothing` method (:210), so the methods
# below are more specific than both and ambiguous with neither.
function _ml_picard(M, L, K, x::Vector{T}, s, t, mc_sample!, g::B200G, f, verbose, prob, neumann_bc) where {T <: HdpFloat}
    if accelerated(prob) && s == prob.tspan[1] && t == prob.tspan[2]
        return mlp_b200(M, L, K, x, mc_sample!, prob, neumann_bc)
    end
    invoke(_ml_picard, Tuple{Any, Any, Any, Any, Any, Any, Any, Any, Any, Any, Any, Any},
           M, L, K, x, s, t, mc_sample!, g, f, verbose, prob, neumann_bc)
end
function _ml_picard_mlt(M, L, K, x::Vector{T}, s, t, mc_sample!, g::B200G, f, verbose, prob, neumann_bc) where {T <: HdpFloat}
    accelerated(prob) || return invoke(_ml_picard_mlt, Tuple{Any, Any, Any, Any, Any, Any, Any, Any, Any, Any, Any, Any},
                                       M, L, K, x, s, t, mc_sample!, g, f, verbose, prob, neumann_bc)
    # reproduce the `_get_loop_num` partition of the threaded estimator (src/MLP.jl:391-404)
    mlp_b200(M, L, K, x, mc_sample!, prob, neumann_bc; compat_threads = Threads.nthreads())
end

# ------------------------------------------------------------------ _reflect, batched method (src/reflect.jl:58-82)
# Only the Matrix method is overlaid.  The reference has
#   (A) _reflect(a::T, b::T, s::T, e::T) where {T <: Vector}        src/reflect.jl:8
#   (B) _reflect(a::T, b::T, s, e)       where {T <: AbstractArray}  src/reflect.jl:58
# `Matrix{T}` is not a `Vector`, so the method below is disjoint from (A) and a strict subset of (B):
# more specific than (B), ambiguous with nothing.  The Vector method (A) — called once per path-step
# by the CPU recursion, src/MLP.jl:419 — is deliberately left alone: it stays the stock code on the
# `invoke` fall-back path, and a kernel launch per d-vector would serve nobody.
const REFLECT_MIN_ELEMENTS = 1 << 14   # below this the stock broadcast is faster than a round trip
function _reflect(a::Matrix{T}, b::Matrix{T}, s, e) where {T <: HdpFloat}
    if length(a) < REFLECT_MIN_ELEMENTS
        return invoke(_reflect, Tuple{AbstractArray, AbstractArray, Any, Any}, a, b, s, e)
    end
    size(a) == size(b) || error("a not same dim as b")
    d = size(a, 1); B = length(a) ÷ d
    out = similar(b)
    scalar = s isa Number
    lo = scalar ? T[s] : Vector{T}(s); hi = scalar ? T[e] : Vector{T}(e)
    h = handle()
    sym = T === Float64 ? :hdp_reflect_f64 : :hdp_reflect_f32
    rc = ccall((sym, libhdp), Cint,
               (Ptr{Cvoid}, Ptr{T}, Ptr{T}, Ptr{T}, Cint, Cint, Ptr{T}, Ptr{T}, Cint, Ptr{Int32}, Cuint),
               h.ptr, a, b, out, d, B, lo, hi, scalar, C_NULL, HDP_MEM_HOST)
    check(h, rc)
    out
end

# ------------------------------------------------------------------ samplers (src/MCSample.jl:28-33, 63-71)
function _mc_sample!(s, x_mc::Array{T}, x) where {T <: HdpFloat}
    keep = Any[]
    d = size(x_mc, 1); B = size(x_mc, 2); K = length(x_mc) ÷ (d * B)
    h = handle()
    sym = T === Float64 ? :hdp_mc_sample_f64 : :hdp_mc_sample_f32
    GC.@preserve keep x begin
        rc = ccall((sym, libhdp), Cint,
                   (Ptr{Cvoid}, Ptr{T}, Ptr{T}, Cint, Cint, Cint, Ref{HdpSampler}, UInt64, Ptr{T}, Cuint),
                   h.ptr, x_mc, isnothing(x) ? C_NULL : pointer(x), d, B, K, sampler_desc(s, keep), rand(UInt64), C_NULL,
                   HDP_MEM_HOST)
    end
    check(h, rc)
    x_mc
end
b200_sample!(s::UniformSampling, x_mc, x = nothing) = _mc_sample!(s, x_mc, nothing)
b200_sample!(s::NormalSampling, x_mc, x = nothing) = _mc_sample!(s, x_mc, s.shifted ? x : nothing)

# ------------------------------------------------------------------ sde_loop! (src/DeepSplitting.jl:153-166)
# μ/σ expressions of UserCoef kinds for the K1 entry points (hdp_set_user_exprs copies the strings)
function set_user_coefs!(h::Handle, μ, σ)
    em = user_expr(μ); es = user_expr(σ)
    if isnothing(em) && isnothing(es)
        return check(h, ccall((:hdp_set_user_exprs, libhdp), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), h.ptr, C_NULL))
    end
    GC.@preserve em es begin
        ue = HdpUserExprs(isnothing(em) ? C_NULL : pointer(em), isnothing(es) ? C_NULL : pointer(es), C_NULL, C_NULL)
        check(h, ccall((:hdp_set_user_exprs, libhdp), Cint, (Ptr{Cvoid}, Ref{HdpUserExprs}), h.ptr, ue))
    end
end

# The raw call: pointers are host or device addresses according to `mem_flags`.  y1 holds x0_batch on
# entry (:182); with `z` it also performs `mc_sample!(z, y1)` (:186-189) in the same kernel.
function sde_loop_ccall!(::Type{T}, y0, y1, B, nsteps, dt, ts, μ, σ, tspan, d, neumann_bc, x0_sample!, mc_sample!, z, K,
                         mem_flags; seed = rand(UInt64), dWall_inject = Ptr{T}(C_NULL)) where {T <: HdpFloat}
    keep = Any[]
    pb = problem_desc(μ, σ, SumSq(0.0), FZero(), tspan, neumann_bc, d, keep)
    h = handle()
    set_user_coefs!(h, μ, σ)
    tsv = isnothing(ts) ? nothing : T[ts[length(ts) - i] for i in 1:nsteps]   # t = ts[N + 1 - i] (:157), 1-based
    sym = T === Float64 ? :hdp_sde_loop_f64 : :hdp_sde_loop_f32
    GC.@preserve keep tsv begin
        rc = ccall((sym, libhdp), Cint,
                   (Ptr{Cvoid}, Ref{HdpProblem}, Ptr{T}, Ptr{T}, Cint, Cint, T, Ptr{T}, Ref{HdpSampler}, UInt64, Ptr{T},
                    Ref{HdpSampler}, Ptr{T}, Cint, Cint, Cuint),
                   h.ptr, pb, y0, y1, B, nsteps, T(dt), isnothing(tsv) ? C_NULL : pointer(tsv),
                   sampler_desc(x0_sample!, keep), seed, dWall_inject, sampler_desc(mc_sample!, keep), z, K, 0, mem_flags)
    end
    check(h, rc)
    nothing
end

"""
    _sde_loop!(y0, y1, dWall, μ, σ, p, dt, ts, N, neumann_bc, x0_sample!)

The method `HighDimPDE._sde_loop!` gets once `julia/patches/deepsplitting_sde_loop_hook.patch` is applied to
src/DeepSplitting.jl (the patch moves the body of the `sde_loop!` closure, :153-166, into a generic method of
that name and leaves a one-line closure that forwards to it).  `Matrix{T}` states and kind coefficients are
narrower than the generic method's untyped arguments: more specific, unambiguous.  `dWall` is not filled —
the normals are generated in registers — only its third dimension (the number of steps) is read.
"""
function sde_loop_b200!(y0::Matrix{T}, y1::Matrix{T}, dWall::Array{T, 3}, μ::B200Coef, σ::B200Coef, p, dt, ts, N, neumann_bc,
                        x0_sample!) where {T <: HdpFloat}
    GC.@preserve y0 y1 begin
        sde_loop_ccall!(T, pointer(y0), pointer(y1), size(y1, 2), size(dWall, 3), dt, ts, μ, σ, (first(ts), last(ts)),
                        size(y1, 1), neumann_bc, x0_sample!, NoSampling(), Ptr{T}(C_NULL), 0, HDP_MEM_HOST)
    end
end
if isdefined(HighDimPDE, :_sde_loop!)   # the hook patch is applied: drop in
    @eval HighDimPDE._sde_loop!(y0::Matrix{T}, y1::Matrix{T}, dWall::Array{T, 3}, μ::B200Coef, σ::B200Coef, p, dt, ts, N,
                                neumann_bc, x0_sample!) where {T <: HdpFloat} =
        sde_loop_b200!(y0, y1, dWall, μ, σ, p, dt, ts, N, neumann_bc, x0_sample!)
end

# convenience form without the reference's calling convention (what the tests and the epoch batching use)
function sde_loop_b200!(y0::Matrix{T}, y1::Matrix{T}, nsteps::Integer, dt, prob, neumann_bc, x0_sample!;
                        mc_sample! = NoSampling(), z = nothing, seed = rand(UInt64), dWall = nothing) where {T <: HdpFloat}
    K = isnothing(z) ? 0 : size(z, 3)
    GC.@preserve y0 y1 z dWall begin
        sde_loop_ccall!(T, pointer(y0), pointer(y1), size(y1, 2), nsteps, dt, nothing, prob.μ, prob.σ, prob.tspan, size(y1, 1),
                        neumann_bc, x0_sample!, mc_sample!, isnothing(z) ? Ptr{T}(C_NULL) : pointer(z), K, HDP_MEM_HOST; seed,
                        dWall_inject = isnothing(dWall) ? Ptr{T}(C_NULL) : pointer(dWall))   # parity mode: the array after randn!
    end
end

# E epochs of paths per launch (hdp_sde_loop_epochs_*): y0, y1 are d×B×E, z optionally d×B×K×E; slice e is
# what sde_loop_b200!(…; seed = seed + e) returns.  DeepSplitting's paths do not depend on the network, so a
# training loop can simulate E epochs' worth at once and consume one slice per epoch (INTEGRATION.md §5).
function sde_loop_epochs_b200!(y0::Array{T, 3}, y1::Array{T, 3}, nsteps::Integer, dt, prob, neumann_bc, x0_sample!;
                               mc_sample! = NoSampling(), z = nothing, seed = rand(UInt64), mem_flags = HDP_MEM_HOST,
                               ptrs = nothing) where {T <: HdpFloat}
    keep = Any[]
    d, B, E = size(y1)
    pb = problem_desc(prob.μ, prob.σ, SumSq(0.0), FZero(), prob.tspan, neumann_bc, d, keep)
    K = isnothing(z) ? 0 : size(z, 3)
    h = handle(); set_user_coefs!(h, prob.μ, prob.σ)
    p0, p1, pz = isnothing(ptrs) ? (pointer(y0), pointer(y1), isnothing(z) ? Ptr{T}(C_NULL) : pointer(z)) : ptrs
    sym = T === Float64 ? :hdp_sde_loop_epochs_f64 : :hdp_sde_loop_epochs_f32
    GC.@preserve keep y0 y1 z begin
        rc = ccall((sym, libhdp), Cint,
                   (Ptr{Cvoid}, Ref{HdpProblem}, Ptr{T}, Ptr{T}, Cint, Cint, T, Ptr{T}, Ref{HdpSampler}, UInt64, Cint,
                    Ref{HdpSampler}, Ptr{T}, Cint, Cuint),
                   h.ptr, pb, p0, p1, B, nsteps, T(dt), C_NULL, sampler_desc(x0_sample!, keep), seed, E,
                   sampler_desc(mc_sample!, keep), pz, K, mem_flags)
    end
    check(h, rc)
    nothing
end

# ------------------------------------------------------------------ EM ensembles (src/NNKolmogorov.jl:81-112, src/NNStopping.jl:90-99)
# Stands in for `solve(EnsembleProblem(SDEProblem(μ, σ, x0, tspan)[, prob_func]), EM(); dt, trajectories,
# adaptive = false[, saveat = dt])`: x is d×trajectories (start points in, final states out); with
# `saveat = true` returns `Array(sim)` (d × N+1 × trajectories), else x.
function em_ensemble_b200!(x::Matrix{T}, prob, dt; saveat = false, seed = rand(UInt64)) where {T <: HdpFloat}
    keep = Any[]
    d, B = size(x)                      # prob.x is `nothing` for NNKolmogorov problems (src/NNKolmogorov.jl:19)
    pb = problem_desc(prob.μ, prob.σ, SumSq(zero(T)), FZero(), prob.tspan, nothing, d, keep)
    N = Int(round((prob.tspan[2] - prob.tspan[1]) / dt))
    traj = saveat ? Array{T}(undef, d, N + 1, B) : nothing
    h = handle(); set_user_coefs!(h, prob.μ, prob.σ)
    sym = T === Float64 ? :hdp_em_ensemble_f64 : :hdp_em_ensemble_f32
    GC.@preserve keep x traj begin
        rc = ccall((sym, libhdp), Cint,
                   (Ptr{Cvoid}, Ref{HdpProblem}, Ptr{T}, Cint, Cint, T, T, Ptr{T}, UInt64, Ptr{T}, Cuint),
                   h.ptr, pb, x, B, N, T(dt), T(prob.tspan[1]), saveat ? pointer(traj) : C_NULL, seed, C_NULL, HDP_MEM_HOST)
    end
    check(h, rc)
    saveat ? traj : x
end

end # module
