# HighDimPDEB200CUDAExt.jl — package extension loaded when CUDA.jl is present: the same entry points on
# `CuArray`s.  Device pointers go straight to libhdpde_b200 (`mem_flags = HDP_MEM_DEVICE`), so the paths of
# `solve(prob, DeepSplitting(nn), dt)` never cross PCIe: `y0`, `y1`, `z` stay where Flux reads them.
# NOT EXECUTED IN THIS REPOSITORY'S CI (no Julia toolchain in the build image); see INTEGRATION.md.
module HighDimPDEB200CUDAExt

using CUDA, HighDimPDE, HighDimPDEB200
import HighDimPDE: _reflect
import HighDimPDEB200: B200Coef, HdpFloat, HDP_MEM_DEVICE, HDP_ASYNC, sde_loop_ccall!, sde_loop_epochs_b200!, handle, check, libhdp

raw(a::CuArray{T}) where {T} = reinterpret(Ptr{T}, pointer(a))

# Drop-in for the hook of julia/patches/deepsplitting_sde_loop_hook.patch.  The library's stream is its own:
# the caller's CUDA.jl work on y1 must be complete before (CUDA.synchronize()), and the call below is synchronous
# (no HDP_ASYNC), so Flux sees finished paths.  `CuMatrix{T}` is narrower than the generic method's untyped
# arguments and disjoint from the `Matrix{T}` method of HighDimPDEB200: unambiguous.
function sde_loop_cuda!(y0::CuMatrix{T}, y1::CuMatrix{T}, dWall::CuArray{T, 3}, μ::B200Coef, σ::B200Coef, p, dt, ts, N, neumann_bc,
                        x0_sample!) where {T <: HdpFloat}
    CUDA.synchronize()
    GC.@preserve y0 y1 begin
        sde_loop_ccall!(T, raw(y0), raw(y1), size(y1, 2), size(dWall, 3), dt, ts, μ, σ, (first(ts), last(ts)), size(y1, 1),
                        neumann_bc, x0_sample!, HighDimPDE.NoSampling(), Ptr{T}(C_NULL), 0, HDP_MEM_DEVICE)
    end
end
if isdefined(HighDimPDE, :_sde_loop!)
    @eval HighDimPDE._sde_loop!(y0::CuMatrix{T}, y1::CuMatrix{T}, dWall::CuArray{T, 3}, μ::B200Coef, σ::B200Coef, p, dt, ts, N,
                                neumann_bc, x0_sample!) where {T <: HdpFloat} =
        sde_loop_cuda!(y0, y1, dWall, μ, σ, p, dt, ts, N, neumann_bc, x0_sample!)
end

# E epochs of paths per launch, asynchronously: enqueue, train on the previous slice meanwhile, `HighDimPDEB200.sync()`
function sde_loop_epochs_cuda!(y0::CuArray{T, 3}, y1::CuArray{T, 3}, nsteps, dt, prob, neumann_bc, x0_sample!; z = nothing,
                               mc_sample! = HighDimPDE.NoSampling(), seed = rand(UInt64), async = true) where {T <: HdpFloat}
    CUDA.synchronize()
    pz = isnothing(z) ? Ptr{T}(C_NULL) : raw(z)
    sde_loop_epochs_b200!(Array{T, 3}(undef, 0, 0, 0), Array{T, 3}(undef, size(y1)...), nsteps, dt, prob, neumann_bc, x0_sample!;
                          mc_sample!, z = isnothing(z) ? nothing : Array{T, 4}(undef, size(z)...), seed,
                          mem_flags = HDP_MEM_DEVICE | (async ? HDP_ASYNC : UInt32(0)), ptrs = (raw(y0), raw(y1), pz))
end

# `_reflect` on device matrices (src/reflect.jl:58-82 runs one broadcast pass and one host sync per bounce there)
function _reflect(a::CuMatrix{T}, b::CuMatrix{T}, s, e) where {T <: HdpFloat}
    size(a) == size(b) || error("a not same dim as b")
    out = similar(b)
    scalar = s isa Number
    lo = scalar ? CuArray(T[s]) : CuArray{T}(s); hi = scalar ? CuArray(T[e]) : CuArray{T}(e)
    h = handle()
    sym = T === Float64 ? :hdp_reflect_f64 : :hdp_reflect_f32
    CUDA.synchronize()
    GC.@preserve a b out lo hi begin
        rc = ccall((sym, libhdp), Cint,
                   (Ptr{Cvoid}, Ptr{T}, Ptr{T}, Ptr{T}, Cint, Cint, Ptr{T}, Ptr{T}, Cint, Ptr{Int32}, Cuint),
                   h.ptr, raw(a), raw(b), raw(out), size(a, 1), size(a, 2), raw(lo), raw(hi), scalar, C_NULL, HDP_MEM_DEVICE)
    end
    check(h, rc)
    out
end

end # module
