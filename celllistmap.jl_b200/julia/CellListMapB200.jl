# CellListMapB200.jl -- Julia host package over libclm_b200.so (include/clm_b200.h).
#
# Keeps the CellListMap.jl 0.10 API surface for the cutoff-pair path -- ParticleSystem, pairwise!, update!,
# resize_output!, neighborlist, InPlaceNeighborList, neighborlist!, NeighborPair, get_computing_box -- with the pair
# function either one of the compiled-in catalogue (LJ / Coulomb energy and forces, distance histogram, mean pairwise
# velocity, minimum distance, neighbour list) or a CustomPairFunction: CUDA C++ source text that the library compiles at
# run time with NVRTC into the same sweep kernel (a Julia closure cannot run on the device).  All compute happens in
# hand-written sm_100a CUDA kernels behind the C ABI; there is no CUDA.jl codegen and no CPU fallback.
#
# STATUS: source-complete, NOT executed in the build environment (no Julia toolchain there or on the GPU box);
# the same ABI is exercised end to end by the Python twin celllistmap.jl_b200/api.py.
module CellListMapB200

using StaticArrays
using LinearAlgebra: Diagonal

export ParticleSystem, ParticleSystemPositions, pairwise!, update!, resize_output!, neighborlist, neighborlist!, InPlaceNeighborList,
       NeighborPair, get_computing_box, LJEnergy, LJForces, LJEnergyAndForces, CoulombEnergy, CoulombEnergyAndForces,
       DistanceHistogram, PairwiseVelocities, MinimumDistanceMap, MinimumDistance, EnergyAndForces,
       CustomPairFunction, CustomOutput

const libclm = get(ENV, "CLM_B200_LIB", joinpath(@__DIR__, "..", "libclm_b200.so"))

const CLM_F32, CLM_F64 = Cint(0), Cint(1)
const CLM_ORTHORHOMBIC, CLM_TRICLINIC, CLM_NONPERIODIC = Cint(0), Cint(1), Cint(2)
const CLM_RESET = Cint(1)

struct ClmBoxInfo           # clm_box_info
    dim::Int32; dtype::Int32; cell_type::Int32; lcell::Int32
    nc::NTuple{3,Int64}
    cutoff::Float64; cutoff_sqr::Float64
    input_unit_cell::NTuple{9,Float64}; aligned_unit_cell::NTuple{9,Float64}
    rotation::NTuple{9,Float64}; inv_rotation::NTuple{9,Float64}
    computing_box_min::NTuple{3,Float64}; computing_box_max::NTuple{3,Float64}
    cell_size::NTuple{3,Float64}; origin::NTuple{3,Float64}
end

# clm_status -> the exception the reference throws for the same condition (INTEGRATION.md, error table)
function _check(h::Ptr{Cvoid}, code::Cint)
    code == 0 && return nothing
    msg = unsafe_string(ccall((:clm_last_error, libclm), Cstring, (Ptr{Cvoid},), h))
    code in (1, 2, 3) && throw(ArgumentError(msg))
    code == 5 && throw(DimensionMismatch(msg))
    error(msg)
end

"""NeighborPair (src/API/NeighborPair.jl:19-33): what a pair function sees; consumed on the device by the catalogue."""
struct NeighborPair{N,T}
    i::Int; j::Int; x::SVector{N,T}; y::SVector{N,T}; d2::T
end
Base.getproperty(p::NeighborPair, s::Symbol) = s === :d ? sqrt(getfield(p, :d2)) : getfield(p, s)

# ---- catalogue -------------------------------------------------------------------------------------------
abstract type CatalogueFunction end
struct LJEnergy{T} <: CatalogueFunction; c6::T; c12::T; end
struct LJForces{T} <: CatalogueFunction; c6::T; c12::T; end
struct LJEnergyAndForces{T} <: CatalogueFunction; c6::T; c12::T; end
struct CoulombEnergy{T,V} <: CatalogueFunction; k::T; weights::V; weights_y::Union{V,Nothing}; end
struct CoulombEnergyAndForces{T,V} <: CatalogueFunction; k::T; weights::V; weights_y::Union{V,Nothing}; end
struct DistanceHistogram{T} <: CatalogueFunction; width::T; end
struct PairwiseVelocities{T,V} <: CatalogueFunction; rbins::Vector{T}; velocities::V; velocities_y::Union{V,Nothing}; end
struct MinimumDistanceMap <: CatalogueFunction end
struct MinimumDistance{T}; i::Int; j::Int; d::T; end
mutable struct EnergyAndForces{T,V}; energy::T; forces::V; end

# ---- ParticleSystemPositions (src/API/ParticleSystemPositions.jl:14-101) -------------------------------------------
"""Positions of one particle set: an owning `Vector{SVector{N,T}}` plus a flag that every mutation raises, so that
`sys.xpositions[i] = v`, `sys.xpositions .= frame`, `push!`, `resize!` ... are followed by a rebuild of the cell list at the
next `pairwise!` / `neighborlist!` without an explicit `update!`.  Views share the flag (they write through `setindex!`)."""
struct ParticleSystemPositions{N,T} <: AbstractVector{SVector{N,T}}
    x::Vector{SVector{N,T}}
    updated::Base.RefValue{Bool}
end
ParticleSystemPositions{N,T}(x) where {N,T} = ParticleSystemPositions{N,T}([SVector{N,T}(v) for v in x], Ref(true))
Base.size(p::ParticleSystemPositions) = size(p.x)
Base.IndexStyle(::Type{<:ParticleSystemPositions}) = IndexLinear()
Base.@propagate_inbounds Base.getindex(p::ParticleSystemPositions, i::Int) = p.x[i]
Base.@propagate_inbounds function Base.setindex!(p::ParticleSystemPositions{N,T}, v, i::Int) where {N,T}
    p.updated[] = true
    p.x[i] = SVector{N,T}(v)
    return p
end
Base.resize!(p::ParticleSystemPositions, n::Integer) = (p.updated[] = true; resize!(p.x, n); p)
Base.push!(p::ParticleSystemPositions{N,T}, v) where {N,T} = (p.updated[] = true; push!(p.x, SVector{N,T}(v)); p)
Base.append!(p::ParticleSystemPositions{N,T}, vs) where {N,T} = (p.updated[] = true; append!(p.x, [SVector{N,T}(v) for v in vs]); p)
Base.empty!(p::ParticleSystemPositions) = (p.updated[] = true; empty!(p.x); p)
# ccall: the AoS n x N memory of the owning vector is what the ABI takes (no conversion); the Vector is rooted by ccall
Base.cconvert(::Type{Ptr{SVector{N,T}}}, p::ParticleSystemPositions{N,T}) where {N,T} = p.x

# ---- ParticleSystem (src/API/ParticleSystem.jl:142-202, AbstractParticleSystem.jl:32-60) -------------------------
mutable struct ParticleSystem{N,T,O}
    handle::Ptr{Cvoid}
    xpositions::ParticleSystemPositions{N,T}        # owning copy + mutation flag
    ypositions::Union{ParticleSystemPositions{N,T},Nothing}
    unitcell::Union{SVector{N,T},SMatrix{N,N,T},Nothing}
    cutoff::T
    lcell::Int
    output::O
    output_name::Symbol
    parallel::Bool
    boxdirty::Bool
end

function ParticleSystem(; positions=nothing, xpositions=nothing, ypositions=nothing, unitcell=nothing, cutoff,
                        output, output_name::Symbol=:default_output_name, parallel::Bool=true,
                        nbatches::Tuple{Int,Int}=(0, 0), lcell=1, device::Integer=0)
    (isnothing(positions) == isnothing(xpositions)) &&
        throw(ArgumentError("Either `positions` OR `xpositions` must be defined."))
    x0 = isnothing(positions) ? xpositions : positions
    N = isnothing(unitcell) ? length(first(x0)) : size(unitcell, 1)
    T = eltype(first(x0)) == Float32 ? Float32 : Float64
    x = ParticleSystemPositions{N,T}(x0)
    y = isnothing(ypositions) ? nothing : ParticleSystemPositions{N,T}(ypositions)
    uc = isnothing(unitcell) ? nothing : (unitcell isa AbstractVector ? SVector{N,T}(unitcell) : SMatrix{N,N,T}(unitcell))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    code = ccall((:clm_create, libclm), Cint, (Ref{Ptr{Cvoid}}, Cint, Cint, Cint, Cint), h, N, T == Float32 ? CLM_F32 : CLM_F64, device, 1)
    code == 0 || error(unsafe_string(ccall((:clm_last_error, libclm), Cstring, (Ptr{Cvoid},), C_NULL)))
    sys = ParticleSystem{N,T,typeof(output)}(h[], x, y, uc, T(cutoff), lcell, output, output_name, parallel, true)
    finalizer(s -> ccall((:clm_destroy, libclm), Cint, (Ptr{Cvoid},), s.handle), sys)
    _sync!(sys)                              # the reference builds the cell list at construction
    return sys
end

# UpdateParticleSystem! (src/internals/ParticleSystem.jl:158-164, :209-224): rebuild only what changed
function _sync!(sys::ParticleSystem{N,T}) where {N,T}
    h = sys.handle
    if sys.boxdirty
        rc = Ref(sys.cutoff)
        if isnothing(sys.unitcell)
            _check(h, ccall((:clm_set_box, libclm), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint, Ref{T}, Cint), h, CLM_NONPERIODIC, C_NULL, 0, rc, sys.lcell))
        else   # sides -> orthorhombic, matrix -> triclinic; SMatrix memory is column-major with columns = lattice vectors: what the ABI takes
            uc = collect(T, vec(sys.unitcell))
            is_matrix = sys.unitcell isa SVector ? 0 : 1
            GC.@preserve uc _check(h, ccall((:clm_set_box, libclm), Cint, (Ptr{Cvoid}, Cint, Ptr{T}, Cint, Ref{T}, Cint),
                                            h, is_matrix == 1 ? CLM_TRICLINIC : CLM_ORTHORHOMBIC, uc, is_matrix, rc, sys.lcell))
        end
        sys.boxdirty = false; sys.xpositions.updated[] = true
    end
    xup = sys.xpositions.updated[]
    yup = !isnothing(sys.ypositions) && sys.ypositions.updated[]
    if xup              # Vector{SVector{N,T}} is the AoS n x N memory the ABI takes: no conversion, one H2D copy
        _check(h, ccall((:clm_set_positions, libclm), Cint, (Ptr{Cvoid}, Cint, Ptr{SVector{N,T}}, Int64, Cint), h, 0, sys.xpositions, length(sys.xpositions), 0))
    end
    if yup
        _check(h, ccall((:clm_set_positions, libclm), Cint, (Ptr{Cvoid}, Cint, Ptr{SVector{N,T}}, Int64, Cint), h, 1, sys.ypositions, length(sys.ypositions), 0))
    end
    if xup || yup
        _check(h, ccall((:clm_build, libclm), Cint, (Ptr{Cvoid},), h))      # UpdateCellList!
        sys.xpositions.updated[] = false
        isnothing(sys.ypositions) || (sys.ypositions.updated[] = false)
    end
    return sys
end

"""update!(sys; xpositions, ypositions, cutoff, unitcell, parallel) (src/API/updating.jl:165-187)"""
function update!(sys::ParticleSystem{N,T}; positions=nothing, xpositions=nothing, ypositions=nothing, cutoff=nothing,
                 unitcell=nothing, parallel=nothing) where {N,T}
    (!isnothing(positions) && !isnothing(xpositions)) && throw(ArgumentError("Either `positions` OR `xpositions` must be provided, not both."))
    x = isnothing(positions) ? xpositions : positions
    (!isnothing(ypositions) && isnothing(sys.ypositions)) && throw(ArgumentError("ypositions can only be set for a two-set particle system"))
    if !isnothing(x); resize!(sys.xpositions, length(x)); sys.xpositions .= x; end                      # every write raises the flag
    if !isnothing(ypositions); resize!(sys.ypositions, length(ypositions)); sys.ypositions .= ypositions; end
    if !isnothing(cutoff); sys.cutoff = T(cutoff); sys.boxdirty = true; end
    if !isnothing(unitcell)
        isnothing(sys.unitcell) && throw(ArgumentError("Manual updating of the unit cell of non-periodic systems is not allowed."))
        sys.unitcell = sys.unitcell isa SVector && unitcell isa AbstractVector ? SVector{N,T}(unitcell) :
                       SMatrix{N,N,T}(unitcell isa AbstractVector ? SMatrix{N,N,T}(Diagonal(unitcell)) : unitcell)
        sys.boxdirty = true
    end
    isnothing(parallel) || (sys.parallel = parallel)
    return sys
end

resize_output!(sys::ParticleSystem, n::Int) = (resize!(sys.output isa EnergyAndForces ? sys.output.forces : (sys.output isa CustomOutput ? sys.output.per_particle : sys.output), n); sys)

function get_computing_box(sys::ParticleSystem{N,T}) where {N,T}
    _sync!(sys)
    b = Ref{ClmBoxInfo}()
    _check(sys.handle, ccall((:clm_get_box, libclm), Cint, (Ptr{Cvoid}, Ref{ClmBoxInfo}), sys.handle, b))
    return (SVector{N,T}(b[].computing_box_min[1:N]), SVector{N,T}(b[].computing_box_max[1:N]))
end

# ---- pairwise! (src/API/pairwise.jl:48-63) ---------------------------------------------------------------------
_flags(reset) = reset ? CLM_RESET : Cint(0)

function pairwise!(f::LJEnergy, sys::ParticleSystem{N,T}; show_progress=false, reset=true) where {N,T}
    _sync!(sys)
    e = Ref(reset ? zero(T) : T(sys.output)); p = T[f.c6, f.c12]
    _check(sys.handle, ccall((:clm_map_lj, libclm), Cint, (Ptr{Cvoid}, Ptr{T}, Cint, Ref{T}, Ptr{Cvoid}), sys.handle, p, _flags(reset), e, C_NULL))
    return sys.output = e[]
end
function pairwise!(f::LJForces, sys::ParticleSystem{N,T}; show_progress=false, reset=true) where {N,T}
    _sync!(sys)
    length(sys.output) == length(sys.xpositions) || throw(DimensionMismatch("force output must have one entry per particle (resize_output!)"))
    e = Ref(zero(T)); p = T[f.c6, f.c12]    # Vector{SVector{N,T}} forces: written in place by the library
    _check(sys.handle, ccall((:clm_map_lj, libclm), Cint, (Ptr{Cvoid}, Ptr{T}, Cint, Ref{T}, Ptr{SVector{N,T}}), sys.handle, p, _flags(reset), e, sys.output))
    return sys.output
end
function pairwise!(f::LJEnergyAndForces, sys::ParticleSystem{N,T}; show_progress=false, reset=true) where {N,T}
    _sync!(sys)
    e = Ref(reset ? zero(T) : T(sys.output.energy)); p = T[f.c6, f.c12]
    _check(sys.handle, ccall((:clm_map_lj, libclm), Cint, (Ptr{Cvoid}, Ptr{T}, Cint, Ref{T}, Ptr{SVector{N,T}}), sys.handle, p, _flags(reset), e, sys.output.forces))
    sys.output.energy = e[]
    return sys.output
end
function pairwise!(f::CoulombEnergy, sys::ParticleSystem{N,T}; show_progress=false, reset=true) where {N,T}
    _sync!(sys)
    e = Ref(reset ? zero(T) : T(sys.output)); k = Ref(T(f.k))
    GC.@preserve f begin   # pointer(f.weights_y) is only valid while the array is rooted
        wy = isnothing(f.weights_y) ? Ptr{T}(C_NULL) : pointer(f.weights_y)
        _check(sys.handle, ccall((:clm_map_coulomb, libclm), Cint, (Ptr{Cvoid}, Ptr{T}, Ptr{T}, Ref{T}, Cint, Ref{T}, Ptr{Cvoid}), sys.handle, f.weights, wy, k, _flags(reset), e, C_NULL))
    end
    return sys.output = e[]
end
function pairwise!(f::CoulombEnergyAndForces, sys::ParticleSystem{N,T}; show_progress=false, reset=true) where {N,T}
    _sync!(sys)
    e = Ref(reset ? zero(T) : T(sys.output.energy)); k = Ref(T(f.k))
    GC.@preserve f begin
        wy = isnothing(f.weights_y) ? Ptr{T}(C_NULL) : pointer(f.weights_y)
        _check(sys.handle, ccall((:clm_map_coulomb, libclm), Cint, (Ptr{Cvoid}, Ptr{T}, Ptr{T}, Ref{T}, Cint, Ref{T}, Ptr{SVector{N,T}}), sys.handle, f.weights, wy, k, _flags(reset), e, sys.output.forces))
    end
    sys.output.energy = e[]
    return sys.output
end
function pairwise!(f::DistanceHistogram, sys::ParticleSystem{N,T}; show_progress=false, reset=true) where {N,T}
    _sync!(sys)                              # output::Vector{Int}
    _check(sys.handle, ccall((:clm_map_dist_hist, libclm), Cint, (Ptr{Cvoid}, Ref{T}, Cint, Cint, Ptr{Int64}), sys.handle, Ref(T(f.width)), length(sys.output), _flags(reset), sys.output))
    return sys.output
end
function pairwise!(f::PairwiseVelocities, sys::ParticleSystem{N,T}; show_progress=false, reset=true) where {N,T}
    _sync!(sys)                              # output = (counts::Vector{Int}, sums::Vector{T}); velocities::Vector{SVector{N,T}}
    counts, sums = sys.output
    GC.@preserve f begin
        vy = isnothing(f.velocities_y) ? Ptr{SVector{N,T}}(C_NULL) : pointer(f.velocities_y)
        _check(sys.handle, ccall((:clm_map_pairvel, libclm), Cint, (Ptr{Cvoid}, Ptr{SVector{N,T}}, Ptr{SVector{N,T}}, Ptr{T}, Cint, Cint, Ptr{Int64}, Ptr{T}),
                                 sys.handle, f.velocities, vy, f.rbins, length(f.rbins) - 1, _flags(reset), counts, sums))
    end
    return sys.output
end
function pairwise!(::MinimumDistanceMap, sys::ParticleSystem{N,T}; show_progress=false, reset=true) where {N,T}
    _sync!(sys)
    i = Ref(Int64(reset ? 0 : sys.output.i)); j = Ref(Int64(reset ? 0 : sys.output.j)); d = Ref(reset ? typemax(T) : T(sys.output.d))
    _check(sys.handle, ccall((:clm_map_mindist, libclm), Cint, (Ptr{Cvoid}, Cint, Ref{Int64}, Ref{Int64}, Ref{T}), sys.handle, _flags(reset), i, j, d))
    return sys.output = MinimumDistance{T}(i[], j[], d[])
end
# ---- user pair functions: CUDA C++ source compiled at run time (clm_custom_compile / clm_map_custom) ----------------
struct ClmCustomInfo; nscalar::Int32; npart::Int32; naux::Int32; hist::Int32; scalar_min_mask::Int32; scalar_max_mask::Int32; end
"""A user pair function: `source` defines a stateless struct `name` (interface: include/clm_b200.h).  `params` (<= 16)
reach the functor as par[]; `aux`/`aux_y` are per-particle side arrays (Vector{T} or Vector{SVector{NAUX,T}})."""
mutable struct CustomPairFunction
    source::String; name::String; params::Vector{Float64}; aux; aux_y
    compiled::Dict{Ptr{Cvoid},Tuple{Int32,ClmCustomInfo}}
end
CustomPairFunction(source, name; params=Float64[], aux=nothing, aux_y=nothing) =
    CustomPairFunction(source, name, collect(Float64, params), aux, aux_y, Dict{Ptr{Cvoid},Tuple{Int32,ClmCustomInfo}}())
"""Output of a CustomPairFunction, reduced with `+` like the reference's default reducer (src/API/parallel_custom.jl:213)."""
mutable struct CustomOutput{T,P}
    scalars::Vector{T}          # NSCALAR
    per_particle::P             # Vector{SVector{NPART,T}} (or Vector{T} for NPART == 1), one entry per particle of set x
    hist_counts::Vector{Int}
    hist_sums::Vector{T}
end
function _compiled(f::CustomPairFunction, sys::ParticleSystem)
    get!(f.compiled, sys.handle) do
        id = Ref{Int32}(-1); info = Ref{ClmCustomInfo}()
        _check(sys.handle, ccall((:clm_custom_compile, libclm), Cint, (Ptr{Cvoid}, Cstring, Cstring, Ref{Int32}, Ref{ClmCustomInfo}), sys.handle, f.source, f.name, id, info))
        (id[], info[])
    end
end
function pairwise!(f::CustomPairFunction, sys::ParticleSystem{N,T}; show_progress=false, reset=true) where {N,T}
    _sync!(sys)
    id, info = _compiled(f, sys)
    out = sys.output::CustomOutput
    info.nscalar > 0 && length(out.scalars) != info.nscalar && throw(DimensionMismatch("output.scalars must have NSCALAR entries"))
    info.npart > 0 && length(out.per_particle) != length(sys.xpositions) && throw(DimensionMismatch("output.per_particle must have one entry per particle (resize_output!)"))
    info.naux > 0 && isnothing(f.aux) && throw(ArgumentError("the functor reads per-particle side arrays: aux is required"))
    par = T.(f.params)
    ptr(a) = isnothing(a) ? C_NULL : Ptr{Cvoid}(pointer(a))
    GC.@preserve par f out _check(sys.handle, ccall((:clm_map_custom, libclm), Cint,
        (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int64}, Ptr{Cvoid}),
        sys.handle, id, isempty(par) ? C_NULL : ptr(par), length(par), ptr(f.aux), ptr(f.aux_y), info.hist != 0 ? length(out.hist_counts) : 0, _flags(reset),
        info.nscalar > 0 ? ptr(out.scalars) : C_NULL, info.npart > 0 ? ptr(out.per_particle) : C_NULL,
        info.hist != 0 ? pointer(out.hist_counts) : Ptr{Int64}(C_NULL), info.hist != 0 ? ptr(out.hist_sums) : C_NULL))
    return out
end
pairwise!(f::Function, sys::ParticleSystem; kw...) =
    throw(ArgumentError("a Julia closure cannot run on the device: use a catalogue functor (LJ/Coulomb, histogram, pair velocity, minimum distance, neighbour list) or a CustomPairFunction (CUDA C++ source compiled at run time)"))

# ---- neighbour lists (src/API/neighborlist.jl:12-15, :84-111, :159-169, :217-231, :314-340) ------------------------
mutable struct InPlaceNeighborList{N,T}
    sys::ParticleSystem{N,T,Nothing}
    list::Vector{Tuple{Int,Int,T}}
    n::Int
end
function InPlaceNeighborList(; x, y=nothing, cutoff, unitcell=nothing, parallel=true, show_progress=false, nbatches=(0, 0), lcell=1)
    sys = ParticleSystem(; xpositions=x, ypositions=y, unitcell, cutoff, output=nothing, output_name=:nb, parallel, nbatches, lcell)
    N, T = length(first(sys.xpositions)), eltype(first(sys.xpositions))
    return InPlaceNeighborList{N,T}(sys, Tuple{Int,Int,T}[], 0)
end
update!(nb::InPlaceNeighborList, x, y=nothing; cutoff=nothing, unitcell=nothing, parallel=nothing) =
    (update!(nb.sys; xpositions=x, ypositions=y, cutoff, unitcell, parallel); nb)
function neighborlist!(nb::InPlaceNeighborList{N,T}) where {N,T}
    _sync!(nb.sys)
    n = Ref{Int64}(0)
    _check(nb.sys.handle, ccall((:clm_neighborlist, libclm), Cint, (Ptr{Cvoid}, Cint, Ref{Int64}), nb.sys.handle, 0, n))
    resize!(nb.list, n[])                    # Tuple{Int,Int,T} is the 24-byte record the library writes in place
    n[] > 0 && _check(nb.sys.handle, ccall((:clm_neighborlist_copy, libclm), Cint, (Ptr{Cvoid}, Ptr{Tuple{Int,Int,T}}, Int64, Cint), nb.sys.handle, nb.list, n[], 0))
    nb.n = n[]
    return nb.list
end
function neighborlist(; xpositions=nothing, positions=nothing, ypositions=nothing, cutoff, unitcell=nothing, parallel=true,
                      show_progress=false, nbatches=(0, 0), lcell=1)
    x = isnothing(positions) ? xpositions : positions
    return copy(neighborlist!(InPlaceNeighborList(; x, y=ypositions, cutoff, unitcell, parallel, show_progress, nbatches, lcell)))
end

end # module
