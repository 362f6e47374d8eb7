# LSMB200.jl — ccall glue that puts liblsm_b200.so behind LevelSetMethods.jl's own API.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build / GPU images have no Julia (SURVEY.md headline 3).
# It is written against include/lsm_b200.h and mirrors, line for line, what the Python host mirror
# (levelsetmethods.jl_b200/api.py) does through ctypes, which IS tested on the GPU.
#
# Usage:   using LevelSetMethods, LSMB200
#          eq = LevelSetEquation(; terms, ic = ϕ, bc, integrator = RK3())
#          LSMB200.integrate!(eq, tf)          # same semantics as LevelSetMethods.integrate!
#
# Seam (SURVEY.md §8b): the reference dispatches integrate! -> _integrate!(ls, ϕ, integrator, terms, tc, tf,
# Δt_max, prehook, posthook) (src/timestepping.jl:101).  This module provides that method for a device-backed
# state, plus compute_cfl / _advance! equivalents used when host hooks force a step-by-step loop.
module LSMB200

using LevelSetMethods
import LevelSetMethods as LSM
using StaticArrays

const LIB = get(ENV, "LSM_B200_LIB", joinpath(@__DIR__, "..", "liblsm_b200.so"))

# ---- enums of include/lsm_b200.h -------------------------------------------------------------------
const LSM_OK, LSM_ERR_ARG, LSM_ERR_CFL, LSM_ERR_TIME, LSM_ERR_BC = Int32(0), Int32(1), Int32(2), Int32(3), Int32(4)
const F32, F64 = Int32(0), Int32(1)
const BC_PERIODIC, BC_EXTRAP, BC_SYMMETRY = Int32(0), Int32(1), Int32(2)
const TERM_ADVECTION, TERM_NORMAL, TERM_CURVATURE, TERM_EIKONAL = Int32(0), Int32(1), Int32(2), Int32(3)
const COEF_CONST, COEF_FIELD, COEF_SEPARABLE, COEF_NONE = Int32(0), Int32(1), Int32(2), Int32(3)
const TS_NONE, TS_COS, TS_HOST = Int32(0), Int32(1), Int32(2)

struct CBC
    kind::Int32
    P::Int32
end

struct CTerm                      # lsm_term
    kind::Int32
    scheme::Int32
    coef_kind::Int32
    tscale_kind::Int32
    cval::NTuple{3, Float64}
    tparam::Float64
    field::Ptr{Cvoid}
end

last_error() = unsafe_string(ccall((:lsm_last_error, LIB), Cstring, ()))

# status -> the exception the reference throws at the cited line
function check(rc::Int32)
    rc == LSM_OK && return nothing
    msg = last_error()
    rc in (LSM_ERR_CFL, LSM_ERR_TIME, LSM_ERR_BC, LSM_ERR_ARG) && throw(ArgumentError(msg))   # levelsetterms.jl:26, levelsetequation.jl:196, boundaryconditions.jl:184
    error("lsm_b200 (status $rc): $msg")
end

# ---- context -----------------------------------------------------------------------------------------
mutable struct Context
    handle::Ptr{Cvoid}
end
function Context(device::Integer = 0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:lsm_ctx_create, LIB), Int32, (Int32, Ref{Ptr{Cvoid}}), device, h))
    ctx = Context(h[])
    finalizer(c -> ccall((:lsm_ctx_destroy, LIB), Int32, (Ptr{Cvoid},), c.handle), ctx)
    return ctx
end
# Multi-GPU: one Julia process per GPU; rank 0 creates the id and the host broadcasts its 128 bytes (MPI.jl ...).
nccl_unique_id() = (id = zeros(UInt8, 128); check(ccall((:lsm_nccl_unique_id, LIB), Int32, (Ptr{UInt8},), id)); id)
function Context(device::Integer, rank::Integer, nranks::Integer, id::Vector{UInt8})
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:lsm_ctx_create_rank, LIB), Int32, (Int32, Int32, Int32, Ptr{UInt8}, Ref{Ptr{Cvoid}}), device, rank, nranks, id, h))
    return Context(h[])
end
const DEFAULT = Ref{Union{Nothing, Context}}(nothing)
default_context() = something(DEFAULT[], (DEFAULT[] = Context(parse(Int, get(ENV, "LSM_B200_DEVICE", "0")))))

# ---- device mirror of a MeshField -----------------------------------------------------------------------
_dtype(::Type{Float32}) = F32
_dtype(::Type{Float64}) = F64
_cbc(::LSM.PeriodicBC) = CBC(BC_PERIODIC, 0)
_cbc(::LSM.ExtrapolationBC{P}) where {P} = CBC(BC_EXTRAP, P)
_cbc(::LSM.SymmetryBC) = CBC(BC_SYMMETRY, 0)

"Device field handle + the host MeshField it mirrors (the host array stays the user's view of `values(ϕ)`)."
mutable struct DeviceField{N}
    handle::Ptr{Cvoid}
    host::LSM.MeshField
end

function DeviceField(ctx::Context, ϕ::LSM.MeshField{N, T, V}) where {N, T, V}
    S = V <: Real ? V : eltype(V)                      # scalar type of a velocity field
    ncomp = V <: Real ? 1 : N
    g = LSM.mesh(ϕ)
    n = Int32[size(g)...]
    lc, hc = Float64[g.lc...], Float64[g.hc...]
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:lsm_field_create, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ref{Ptr{Cvoid}}),
                ctx.handle, N, n, _dtype(float(S)), ncomp, lc, hc, h))
    d = DeviceField{N}(h[], ϕ)
    finalizer(f -> ccall((:lsm_field_destroy, LIB), Int32, (Ptr{Cvoid},), f.handle), d)
    if LSM.has_boundary_conditions(ϕ)
        bcs = LSM.boundary_conditions(ϕ)
        flat = CBC[_cbc(bcs[dd][s]) for dd in 1:N for s in 1:2]
        check(ccall((:lsm_field_set_bc, LIB), Int32, (Ptr{Cvoid}, Ptr{CBC}), d.handle, flat))
    end
    upload!(d)
    return d
end

# Array{SVector{N,T},N} is memory-identical to the (N, n1, ..) AoS layout the ABI expects.
upload!(d::DeviceField) = (v = values(d.host); GC.@preserve v check(ccall((:lsm_field_upload, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), d.handle, pointer(v))); d)
download!(d::DeviceField) = (v = values(d.host); GC.@preserve v check(ccall((:lsm_field_download, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), d.handle, pointer(v))); d)

# ---- terms -> descriptors ---------------------------------------------------------------------------------
_isdefault(f) = f === nothing || (f isa Function && parentmodule(f) === LSM && occursin("#", string(nameof(f))))   # the no-op closure of the term constructors

struct Lowered
    terms::Vector{CTerm}
    keep::Vector{Any}         # device fields kept alive for the call
    device_only::Bool
end

function lower(ctx::Context, terms, ϕ::LSM.MeshField{N}, t) where {N}
    out, keep, device_only = CTerm[], Any[], true
    for term in terms
        kind, scheme, coef = if term isa LSM.AdvectionTerm
            TERM_ADVECTION, (LSM.scheme(term) isa LSM.WENO5 ? Int32(1) : Int32(0)), LSM.velocity(term)
        elseif term isa LSM.NormalMotionTerm
            TERM_NORMAL, Int32(0), LSM.speed(term)
        elseif term isa LSM.CurvatureTerm
            TERM_CURVATURE, Int32(0), LSM.coefficient(term)
        else
            TERM_EIKONAL, Int32(0), term.S₀
        end
        hasproperty(term, :update_func) && !_isdefault(term.update_func) && (device_only = false)
        ck, cval, fld = COEF_CONST, (0.0, 0.0, 0.0), C_NULL
        if coef === nothing
            ck = COEF_NONE
        elseif coef isa LSM.MeshField
            d = DeviceField(ctx, coef); push!(keep, d)
            ck, fld = COEF_FIELD, d.handle
        elseif coef isa Function
            # f(x, t): evaluated on the host at the stage time, like every host callback (slow path)
            device_only = false
            f = LSM.MeshField(x -> coef(x, t), LSM.mesh(ϕ))
            d = DeviceField(ctx, f); push!(keep, d)
            ck, fld = COEF_FIELD, d.handle
        else
            c = Float64[coef...]
            cval = ntuple(i -> i <= length(c) ? c[i] : 0.0, 3)
        end
        push!(out, CTerm(kind, scheme, ck, TS_NONE, cval, 1.0, fld))
    end
    return Lowered(out, keep, device_only)
end

# ---- the three entry points the reference's step loop needs -----------------------------------------------
"`compute_cfl(terms, ϕ, t)` (src/levelsetterms.jl:22-38)"
function compute_cfl(ctx::Context, dϕ::DeviceField, low::Lowered, t)
    dt = Ref{Float64}(0.0)
    check(ccall((:lsm_compute_cfl, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{CTerm}, Int32, Float64, Ptr{Float64}, Ref{Float64}),
                ctx.handle, dϕ.handle, low.terms, length(low.terms), t, C_NULL, dt))
    return dt[]
end

_code(::LSM.ForwardEuler) = Int32(0)
_code(::LSM.RK2) = Int32(1)
_code(::LSM.RK3) = Int32(2)
_stage_times(::LSM.ForwardEuler, tc, Δt) = (tc,)
_stage_times(::LSM.RK2, tc, Δt) = (tc, tc + Δt)
_stage_times(::LSM.RK3, tc, Δt) = (tc, tc + Δt, tc + 0.5Δt)

"`_advance!` (src/timestepping.jl:128-202), one lsm_stage per RK stage so update_term! can run in between."
function advance!(ctx::Context, integ, dϕ::DeviceField, terms, tc, Δt)
    for (s, ts) in enumerate(_stage_times(integ, tc, Δt))
        for term in terms
            LSM.update_term!(term, dϕ.host, ts)       # NB: a hook that reads the stage field needs a download first
        end
        low = lower(ctx, terms, dϕ.host, ts)
        check(ccall((:lsm_stage, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Cvoid}, Ptr{CTerm}, Int32, Float64, Float64, Ptr{Float64}),
                    ctx.handle, _code(integ), s, dϕ.handle, low.terms, length(low.terms), tc, Δt, C_NULL))
    end
end

"""
    integrate!(eq::LevelSetEquation, tf, Δt = Inf; prehook = identity, posthook = identity, ctx)

Drop-in for `LevelSetMethods.integrate!` (src/levelsetequation.jl:194-203) on a dense `MeshField` state.
Without host hooks the whole `_integrate!` loop (src/timestepping.jl:101-122) runs inside the library.
"""
function integrate!(eq::LSM.LevelSetEquation, tf, Δt = Inf; prehook = identity, posthook = identity, ctx::Context = default_context())
    tc = LSM.current_time(eq)
    tf >= tc || throw(ArgumentError("final time $tf must be ≥ initial time $tc: the level-set equation cannot be solved back in time"))
    ϕ = LSM.current_state(eq)
    ϕ isa LSM.MeshField || return LSM.integrate!(eq, tf, Δt; prehook, posthook)        # narrow band etc. stay on the host
    integ = LSM.time_integrator(eq)
    integ isa Union{LSM.ForwardEuler, LSM.RK2, LSM.RK3} || return LSM.integrate!(eq, tf, Δt; prehook, posthook)
    dϕ = DeviceField(ctx, ϕ)
    low = lower(ctx, eq.terms, ϕ, tc)
    if prehook === identity && posthook === identity && low.device_only
        t_out, steps = Ref{Float64}(tc), Ref{Int64}(0)
        rc = ccall((:lsm_integrate, LIB), Int32,
                   (Ptr{Cvoid}, Int32, Float64, Ptr{Cvoid}, Ptr{CTerm}, Int32, Float64, Float64, Float64, Int64, Ref{Float64}, Ref{Int64}),
                   ctx.handle, _code(integ), LSM.cfl(integ), dϕ.handle, low.terms, length(low.terms), tc, Float64(tf), Float64(Δt), -1, t_out, steps)
        download!(dϕ); eq.t = t_out[]
        check(rc)
        return eq
    end
    α = LSM.cfl(integ)
    while tc <= tf - eps(tc)                                            # src/timestepping.jl:104
        prehook !== identity && (download!(dϕ); prehook(eq); upload!(dϕ))   # the hook may mutate the state (levelsetequation.jl:180-185)
        for term in eq.terms
            LSM.update_term!(term, ϕ, tc)
        end
        low = lower(ctx, eq.terms, ϕ, tc)
        dt = min(Δt, α * compute_cfl(ctx, dϕ, low, tc), tf - tc)        # :111
        advance!(ctx, integ, dϕ, eq.terms, tc, dt)
        tc += dt; eq.t = tc
        posthook !== identity && (download!(dϕ); posthook(eq); upload!(dϕ))
    end
    download!(dϕ)
    eq.t = tf                                                           # :120
    return eq
end

"""
    volume(ϕ::MeshField; ctx) / perimeter(ϕ::MeshField; ctx)

Device reductions for `LevelSetMethods.volume` / `perimeter` (src/levelsetops.jl:27-33,139-149).
"""
function volume(ϕ::LSM.MeshField; ctx::Context = default_context())
    d = DeviceField(ctx, ϕ); out = Ref{Float64}(0)
    check(ccall((:lsm_volume, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{Float64}), ctx.handle, d.handle, out))
    return out[]
end
function perimeter(ϕ::LSM.MeshField; ctx::Context = default_context())
    d = DeviceField(ctx, ϕ); out = Ref{Float64}(0)
    check(ccall((:lsm_perimeter, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{Float64}), ctx.handle, d.handle, out))
    return out[]
end

"""
    extend_along_normals!(F, ϕ; nb_iters = 50, cfl = 0.45, frozen = nothing, interface_band = 1.5, min_norm = 1.0e-14, ctx)

Drop-in for `LevelSetMethods.extend_along_normals!` (src/velocityextension.jl:20-78) on dense fields.
"""
function extend_along_normals!(F::LSM.MeshField, ϕ::LSM.MeshField; nb_iters::Integer = 50, cfl::Real = 0.45, frozen = nothing,
                               interface_band::Real = 1.5, min_norm::Real = 1.0e-14, ctx::Context = default_context())
    size(values(F)) == size(values(ϕ)) || throw(ArgumentError("F must have the same size as ϕ"))
    frozen === nothing || size(frozen) == size(values(ϕ)) || throw(ArgumentError("frozen mask must have the same size as ϕ"))
    dF, dϕ = DeviceField(ctx, F), DeviceField(ctx, ϕ)
    mask = frozen === nothing ? UInt8[] : UInt8.(frozen)
    GC.@preserve mask check(ccall((:lsm_extend_along_normals, LIB), Int32,
                                  (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Float64, Ptr{UInt8}, Float64, Float64),
                                  ctx.handle, dF.handle, dϕ.handle, nb_iters, cfl, frozen === nothing ? C_NULL : pointer(mask),
                                  interface_band, min_norm))
    download!(dF)
    return F
end

# set operations on device fields (src/levelsetops.jl:253-325): 0 union!, 1 intersect!, 2 setdiff!, 3 complement!
function csg!(d1::DeviceField, d2::Union{DeviceField, Nothing}, op::Integer; ctx::Context = default_context())
    check(ccall((:lsm_field_csg, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32),
                ctx.handle, d1.handle, d2 === nothing ? C_NULL : d2.handle, op))
    return d1
end
Base.union!(a::DeviceField, b::DeviceField) = csg!(a, b, 0)
Base.intersect!(a::DeviceField, b::DeviceField) = csg!(a, b, 1)
Base.setdiff!(a::DeviceField, b::DeviceField) = csg!(a, b, 2)
complement!(a::DeviceField) = csg!(a, nothing, 3)

end # module
