# StencilsB200.jl — Julia shim: lowers the mutating entry points of rafaqz/Stencils.jl onto the C ABI of
# libstencils_b200.so (include/stencils_b200.h).
#
# STATUS: written against the header, NOT executed (no `julia` in the build image or on the GPU box). The
# executable consumer of the same ABI is the Python mirror in `stencils.jl_b200/` (ctypes), which the tests drive;
# this file is kept line-for-line equivalent to `stencils.jl_b200/ops.py` + `_desc.py`.
#
# What it overrides (signatures unchanged, see SURVEY §8b):
#   gatherstencil!(f, dest, source::AbstractStencilArray)      src/gatherstencil.jl:89-103  -> sb200_update_halo + sb200_gather
#   gatherstencil!(f, A::SwitchingStencilArray)                 src/gatherstencil.jl:77-83   -> same, then switch(A)
#   update_boundary!(A)                                         src/array.jl:195-233         -> sb200_update_halo
#   scatterstencil!(f, op, dest, source)                        src/scatterstencil.jl:36-45  -> sb200_scatter
# for parents that are `B200Array`s (device buffers owned through sb200_malloc) or plain `Array`s
# (host buffers -> sb200_gather_host). `mapstencil`/`gatherstencil` (allocating) keep the reference code and
# reach these methods through dispatch.
module StencilsB200

using Stencils
using Stencils: AbstractStencilArray, SwitchingStencilArray, StencilArray, Stencil, Kernel, Halo, Conditional,
                Remove, Wrap, Reflect, Use, boundary, padding, stencil, radius, offsets, source, dest, switch, kernel
using Statistics: mean

const LIB = joinpath(@__DIR__, "..", "stencils.jl_b200", "lib", "libstencils_b200.so")

# ---- enums (include/stencils_b200.h) ----
const ELTYPE = Dict(Bool => 0, UInt8 => 1, Int32 => 2, Int64 => 3, Float32 => 4, Float64 => 5)
const SB_REMOVE, SB_WRAP, SB_REFLECT, SB_USE = Int32(0), Int32(1), Int32(2), Int32(3)
const SB_SUM, SB_MEAN, SB_MIN, SB_MAX, SB_KERNELDOT, SB_LIFE, SB_DIFFUSION = Int32.(0:6)
const SB_EUNSUPPORTED, SB_ESIZE = 2, 3
# sb200_desc.flags (a binding normally leaves the *_STEP / *_BITS bits to sb200_iterate and the slab plans, which set them by themselves)
const SB_FLAG_ZERO_DEST, SB_FLAG_CELLS_01, SB_FLAG_ALLOW_FMA = Int32(2), Int32(8), Int32(128)
sb_flag_gens(n::Integer) = Int32(n == 2 ? 16 : n == 4 ? 32 : n == 8 ? 64 : (n == 3 || 5 <= n <= 7) ? n << 4 : 0)   # SB200_FLAG_GENS(n)
# Life state one bit per cell: the layout of a BitMatrix whose first dimension is a multiple of 128 (chunks of 64 bits, column-major),
# so `pointer(A.chunks)` of such a BitMatrix can be handed to sb200_gather with these flags and SB200_FLAG_GENS(2 .. 8)
const SB_FLAG_SRC_BITS, SB_FLAG_DST_BITS = Int32(256), Int32(512)

# struct sb200_desc — field order and sizes must match the header (248 bytes).
struct Desc
    struct_size::Int32
    ndim::Int32
    size::NTuple{3,Int64}
    src_ext::NTuple{3,Int64}
    dst_ext::NTuple{3,Int64}
    src_off::NTuple{3,Int32}
    dst_off::NTuple{3,Int32}
    boundary::NTuple{3,Int32}
    eltype::Int32
    out_eltype::Int32
    padval_bits::UInt64
    radius::Int32
    noffsets::Int32
    offsets_host::Ptr{Int32}
    reducer::Int32
    scatter_op::Int32
    scatter_rule::Int32
    born_mask::UInt32
    survive_mask::UInt32
    reserved0::Int32
    weights_host::Ptr{Cvoid}
    alpha::Float64
    region_lo::NTuple{3,Int64}
    region_hi::NTuple{3,Int64}
    flags::Int32
    reserved1::Int32
    mirror_parent::Ptr{Cvoid}   # fused ghost-plane push (slab runs): NULL = off
    mirror_lo::Int64
    mirror_hi::Int64
end

# ---- reducer menu: the only user functions that lower to CUDA kernels ----
struct Life
    born_mask::UInt32
    survive_mask::UInt32
end
Life(; born=(3,), survive=(2, 3)) = Life(reduce(|, (UInt32(1) << b for b in born)), reduce(|, (UInt32(1) << s for s in survive)))
struct Diffusion
    alpha::Float64
end

reducer_enum(::typeof(sum)) = SB_SUM
reducer_enum(::typeof(mean)) = SB_MEAN
reducer_enum(::typeof(minimum)) = SB_MIN
reducer_enum(::typeof(maximum)) = SB_MAX
reducer_enum(::typeof(Stencils.kernelproduct)) = SB_KERNELDOT
reducer_enum(::Life) = SB_LIFE
reducer_enum(::Diffusion) = SB_DIFFUSION
reducer_enum(f) = throw(ArgumentError("unsupported user function $f: only sum, mean, minimum, maximum, kernelproduct, " *
                                      "Life(...) and Diffusion(α) lower to CUDA kernels; there is no fallback path"))

bc_enum(::Remove) = SB_REMOVE
bc_enum(::Wrap) = SB_WRAP
bc_enum(::Reflect) = SB_REFLECT
bc_enum(::Use) = SB_USE

function check(status)
    status == 0 && return nothing
    msg = unsafe_string(ccall((:sb200_last_error, LIB), Cstring, ()))
    status in (1, SB_EUNSUPPORTED, SB_ESIZE) ? throw(ArgumentError(msg)) : error("libstencils_b200 status $status: $msg")
end

"Device buffer owned through the C ABI (column-major, like Array)."
mutable struct B200Array{T,N} <: AbstractArray{T,N}
    ptr::Ptr{Cvoid}
    dims::NTuple{N,Int}
    function B200Array{T}(::UndefInitializer, dims::NTuple{N,Int}) where {T,N}
        p = Ref{Ptr{Cvoid}}()
        check(ccall((:sb200_malloc, LIB), Int32, (Ptr{Ptr{Cvoid}}, Csize_t), p, prod(dims) * sizeof(T)))
        A = new{T,N}(p[], dims)
        finalizer(a -> ccall((:sb200_free, LIB), Int32, (Ptr{Cvoid},), a.ptr), A)
    end
end
Base.size(A::B200Array) = A.dims
Base.similar(A::B200Array, ::Type{T}, dims::Dims) where T = B200Array{T}(undef, dims)
function B200Array(h::Array{T,N}) where {T,N}
    A = B200Array{T}(undef, size(h))
    check(ccall((:sb200_memcpy_h2d, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}), A.ptr, h, sizeof(h), C_NULL))
    A
end
function Base.Array(A::B200Array{T,N}) where {T,N}
    h = Array{T,N}(undef, size(A))
    check(ccall((:sb200_memcpy_d2h, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}), h, A.ptr, sizeof(h), C_NULL))
    check(ccall((:sb200_stream_sync, LIB), Int32, (Ptr{Cvoid},), C_NULL))
    h
end

# ---- array plumbing (SURVEY §8f.3): what StencilArray / SwitchingStencilArray need from their parent ----
# The reference only ever calls size / similar / copyto! / fill! / parent-indexing on the parent array
# (src/array.jl:249-309 `similar` / `copy` rules, 367-385 halo views, 564-611 SwitchingStencilArray); scalar getindex on a
# device buffer is allowed (tests index single cells) but goes through a one-element D2H copy.
Base.IndexStyle(::Type{<:B200Array}) = IndexLinear()
Base.similar(A::B200Array{T}) where T = B200Array{T}(undef, size(A))
Base.similar(A::B200Array{T}, dims::Dims) where T = B200Array{T}(undef, dims)
Base.similar(A::B200Array, ::Type{T}) where T = B200Array{T}(undef, size(A))
function Base.copyto!(dst::B200Array{T}, src::Array{T}) where T
    length(dst) == length(src) || throw(DimensionMismatch("copyto!: $(size(dst)) vs $(size(src))"))
    check(ccall((:sb200_memcpy_h2d, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}), dst.ptr, src, sizeof(src), C_NULL))
    check(ccall((:sb200_stream_sync, LIB), Int32, (Ptr{Cvoid},), C_NULL))
    dst
end
function Base.copyto!(dst::Array{T}, src::B200Array{T}) where T
    length(dst) == length(src) || throw(DimensionMismatch("copyto!: $(size(dst)) vs $(size(src))"))
    check(ccall((:sb200_memcpy_d2h, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}), dst, src.ptr, sizeof(dst), C_NULL))
    check(ccall((:sb200_stream_sync, LIB), Int32, (Ptr{Cvoid},), C_NULL))
    dst
end
function Base.copyto!(dst::B200Array{T}, src::B200Array{T}) where T
    length(dst) == length(src) || throw(DimensionMismatch("copyto!: $(size(dst)) vs $(size(src))"))
    check(ccall((:sb200_memcpy_d2d, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}), dst.ptr, src.ptr, length(src) * sizeof(T), C_NULL))
    dst
end
Base.copy(A::B200Array) = copyto!(similar(A), A)
function Base.fill!(A::B200Array{T}, x) where T
    v = convert(T, x)
    if all(iszero, reinterpret(UInt8, [v]))   # the ABI has a byte memset; anything else goes through the host
        check(ccall((:sb200_memset, LIB), Int32, (Ptr{Cvoid}, Int32, Csize_t, Ptr{Cvoid}), A.ptr, 0, length(A) * sizeof(T), C_NULL))
    else
        copyto!(A, fill(v, size(A)))
    end
    A
end
function Base.getindex(A::B200Array{T}, i::Int) where T   # cold path: one element over PCIe
    @boundscheck checkbounds(A, i)
    r = Ref{T}()
    check(ccall((:sb200_memcpy_d2h, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}), r, A.ptr + (i - 1) * sizeof(T), sizeof(T), C_NULL))
    check(ccall((:sb200_stream_sync, LIB), Int32, (Ptr{Cvoid},), C_NULL))
    r[]
end
function Base.setindex!(A::B200Array{T}, x, i::Int) where T
    @boundscheck checkbounds(A, i)
    r = Ref{T}(convert(T, x))
    check(ccall((:sb200_memcpy_h2d, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}), A.ptr + (i - 1) * sizeof(T), r, sizeof(T), C_NULL))
    check(ccall((:sb200_stream_sync, LIB), Int32, (Ptr{Cvoid},), C_NULL))
    x
end
Base.:(==)(A::B200Array, B::AbstractArray) = Array(A) == B
Base.:(==)(A::AbstractArray, B::B200Array) = A == Array(B)
Base.:(==)(A::B200Array, B::B200Array) = Array(A) == Array(B)

pad3(t, fill) = ntuple(i -> i <= length(t) ? t[i] : fill, 3)

"Descriptor from a StencilArray pair; `keep` holds the arrays the pointers refer to."
function make_desc(f, dst_parent, dst_halo::Int, src::AbstractStencilArray{R,T,N}, src_parent; flags=Int32(0)) where {R,T,N}
    st = stencil(src)
    offs = zeros(Int32, 3, length(st))
    for (k, o) in enumerate(offsets(st)), a in 1:length(o)
        offs[a, k] = o[a]
    end
    sh = padding(src) isa Halo ? R : 0
    red = reducer_enum(f)
    out = Ref{Int32}()
    check(ccall((:sb200_out_eltype, LIB), Int32, (Int32, Int32, Ptr{Int32}), red, ELTYPE[T], out))
    w = st isa Kernel ? collect(T, vec(kernel(st))) : T[]
    bc = boundary(src)
    pv = bc isa Remove ? padbits(T, Stencils.padval(bc)) : UInt64(0)   # Remove() without padval (nothing) -> MethodError
    d = Desc(sizeof(Desc), N, pad3(size(src), 1), pad3(size(src_parent), 1), pad3(size(dst_parent), 1),
             pad3(ntuple(_ -> Int32(sh), N), Int32(0)), pad3(ntuple(_ -> Int32(dst_halo), N), Int32(0)),
             pad3(ntuple(_ -> bc_enum(bc), N), SB_REMOVE), ELTYPE[T], out[], pv, R, length(st), pointer(offs), red, 0, 0,
             f isa Life ? f.born_mask : UInt32(8), f isa Life ? f.survive_mask : UInt32(12), 0,
             isempty(w) ? C_NULL : pointer(w), f isa Diffusion ? f.alpha : 0.0, (0, 0, 0), (0, 0, 0), flags, 0,
             C_NULL, 0, 0)
    return Ref(d), (offs, w)
end
unsigned_of(::Type{T}) where T = sizeof(T) == 1 ? UInt8 : sizeof(T) == 4 ? UInt32 : UInt64
padbits(::Type{T}, v::Number) where T = UInt64(reinterpret(unsigned_of(T), convert(T, v)))
"Copy of a descriptor with some fields replaced."
with(d::Desc; kw...) = Desc((haskey(kw, k) ? kw[k] : getfield(d, k) for k in fieldnames(Desc))...)

dataptr(A::B200Array) = A.ptr
dataptr(A::Array) = Ptr{Cvoid}(pointer(A))

# gatherstencil!(f, dest, source) — src/gatherstencil.jl:89-103
function Stencils.gatherstencil!(f::F, dst, src::AbstractStencilArray{R,T,N,<:Union{B200Array,Array}}) where {F,R,T,N}
    Stencils._checksizes((dst, src))
    dpar, dh = dst isa AbstractStencilArray ? (parent(dst), padding(dst) isa Halo ? R : 0) : (dst, 0)
    d, keep = make_desc(f, dpar, dh, src, parent(src))
    GC.@preserve keep begin
        if parent(src) isa B200Array
            needs_halo = padding(src) isa Halo && !(boundary(src) isa Use)
            needs_halo && check(ccall((:sb200_update_halo, LIB), Int32, (Ref{Desc}, Ptr{Cvoid}, Ptr{Cvoid}), d, dataptr(parent(src)), C_NULL))
            check(ccall((:sb200_gather, LIB), Int32, (Ref{Desc}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), d, dataptr(parent(src)), dataptr(dpar), C_NULL))
            check(ccall((:sb200_stream_sync, LIB), Int32, (Ptr{Cvoid},), C_NULL))   # the reference call is synchronous (:100)
        else
            check(ccall((:sb200_gather_host, LIB), Int32, (Ref{Desc}, Ptr{Cvoid}, Ptr{Cvoid}), d, dataptr(parent(src)), dataptr(dpar)))
        end
    end
    return dst
end

# gatherstencil!(f, dest, A1, A2, ...) with extra array arguments — src/gatherstencil.jl:84-88, 95, 107, 112-113.
# The user function is spelled as a LinearCombination: f(h1, h2, ...) = c1*g1(h1) + c2*g2(h2) + ... (left to right,
# every operation rounded separately); `center` terms and plain-array arguments use the one-offset table {0}.
struct Term
    desc::Ptr{Desc}
    src_parent::Ptr{Cvoid}
    has_coef::Int32
    reserved::Int32
    coef::Float64
end
struct LinearCombination{T<:Tuple}
    terms::T            # one per array argument: g or (coef, g), g in (center, sum, mean, minimum, maximum, kernelproduct)
end
LinearCombination(terms...) = LinearCombination(terms)
term_parts(t::Tuple) = (true, Float64(t[1]), t[2])
term_parts(g) = (false, 0.0, g)

function Stencils.gatherstencil!(f::LinearCombination, dst, A1::AbstractStencilArray{R,T,N,<:B200Array}, args...) where {R,T,N}
    srcs = (A1, args...)
    length(srcs) == length(f.terms) || throw(ArgumentError("$(length(f.terms)) terms but $(length(srcs)) array arguments"))
    Stencils._checksizes((dst, srcs...))
    dpar, dh = dst isa AbstractStencilArray ? (parent(dst), padding(dst) isa Halo ? R : 0) : (dst, 0)
    descs = Vector{Base.RefValue{Desc}}(); keeps = Any[]; terms = Term[]
    for (src, t) in zip(srcs, f.terms)
        has, c, g = term_parts(t)
        sa = src isa AbstractStencilArray ? src : StencilArray(src, Positional(ntuple(_ -> 0, N)); boundary=Remove(zero(T)))
        src isa AbstractStencilArray || g === center || throw(ArgumentError("a plain array argument is indexed, not stencilled"))
        sa = g === center ? StencilArray(parent(sa), Positional(ntuple(_ -> 0, N)), boundary(sa), padding(sa)) : sa
        d, keep = make_desc(g === center ? sum : g, dpar, dh, sa, parent(sa))
        g === center && (d = with(d; radius=Int32(R)))          # ring thickness of the parent = the array's stencil radius
        push!(descs, Ref(d)); push!(keeps, keep)
        push!(terms, Term(Base.unsafe_convert(Ptr{Desc}, descs[end]), dataptr(parent(sa)), Int32(has), Int32(0), c))
    end
    GC.@preserve descs keeps terms begin
        check(ccall((:sb200_gather_multi, LIB), Int32, (Ptr{Term}, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                    terms, length(terms), dataptr(dpar), C_NULL, C_NULL))
        check(ccall((:sb200_stream_sync, LIB), Int32, (Ptr{Cvoid},), C_NULL))
    end
    return dst
end

# gatherstencil!(f, A::SwitchingStencilArray) — src/gatherstencil.jl:77-83
function Stencils.gatherstencil!(f::F, A::SwitchingStencilArray{R,T,N,<:Union{B200Array,Array}}) where {F,R,T,N}
    pd = padding(A) isa Halo ? Halo{:in}() : padding(A)
    src = StencilArray(source(A), stencil(A), boundary(A), pd)
    dst = StencilArray(dest(A), stencil(A), boundary(A), pd)
    Stencils.gatherstencil!(f, dst, src)
    return switch(A)
end

"`for _ in 1:n; A = mapstencil!(f, A); end` as one stream-ordered call (no host sync between steps)."
function iterate!(f, A::SwitchingStencilArray{R,T,N,<:B200Array}, nsteps::Integer) where {R,T,N}
    pd = padding(A) isa Halo ? Halo{:in}() : padding(A)
    src = StencilArray(source(A), stencil(A), boundary(A), pd)
    d, keep = make_desc(f, dest(A), padding(A) isa Halo ? R : 0, src, source(A))
    GC.@preserve keep check(ccall((:sb200_iterate, LIB), Int32, (Ref{Desc}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Cvoid}),
                                  d, dataptr(source(A)), dataptr(dest(A)), nsteps, C_NULL))
    check(ccall((:sb200_stream_sync, LIB), Int32, (Ptr{Cvoid},), C_NULL))
    return iseven(nsteps) ? A : switch(A)
end

# ---- the same loop over ALL the GPUs of the box: sb200_plan_* (include/stencils_b200.h, csrc/slab_plan.cu) ----
# A Julia session is one process, so it uses the single-process form: the library splits the array into slabs along its last
# axis, one per device, owns the slab buffers / mailboxes / streams and runs the exchange schedule (ghost planes over NVLink,
# boundary planes first, interior sweep overlapping the exchange). Results are bit-identical to iterate! on one GPU.
mutable struct SlabPlan
    handle::Ptr{Cvoid}
    size::Dims
    eltype::DataType
end

"SlabPlan(f, A::SwitchingStencilArray over a host Array; devices = 0:ndevices-1, ghost = 0 (library default))"
function SlabPlan(f, A::SwitchingStencilArray{R,T,N,<:Array}; devices = nothing, ghost::Integer = 0, flags::Integer = 0) where {R,T,N}
    padding(A) isa Halo && throw(ArgumentError("slab plans take Conditional padding: the ghost planes are the plan's own ring"))
    n = Ref{Int32}(0)
    check(ccall((:sb200_device_count, LIB), Int32, (Ptr{Int32},), n))
    devs = Int32.(collect(devices === nothing ? (0:n[]-1) : devices))
    src = StencilArray(source(A), stencil(A), boundary(A), padding(A))
    d, keep = make_desc(f, dest(A), 0, src, source(A))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep check(ccall((:sb200_plan_create, LIB), Int32, (Ref{Desc}, Int32, Ptr{Int32}, Int32, Int32, Ptr{Ptr{Cvoid}}),
                                  d, length(devs), devs, ghost, flags, h))
    plan = SlabPlan(h[], size(A), T)
    finalizer(p -> (p.handle != C_NULL && ccall((:sb200_plan_destroy, LIB), Int32, (Ptr{Cvoid},), p.handle); p.handle = C_NULL), plan)
    check(ccall((:sb200_plan_load_host, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), plan.handle, source(A)))
    return plan
end

"n generations on every GPU (returns when they are done; a dead neighbour is an error, not a hang)."
function iterate!(plan::SlabPlan, nsteps::Integer)
    check(ccall((:sb200_plan_iterate, LIB), Int32, (Ptr{Cvoid}, Int32), plan.handle, nsteps))
    check(ccall((:sb200_plan_sync, LIB), Int32, (Ptr{Cvoid},), plan.handle))
    return plan
end

"The current state as a host Array."
function Base.Array(plan::SlabPlan)
    out = Array{plan.eltype}(undef, plan.size)
    check(ccall((:sb200_plan_store_host, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), plan.handle, out))
    return out
end

"`for _ in 1:n; A = mapstencil!(f, A); end` over every GPU of the box: A is a SwitchingStencilArray over a host Array."
function iterate_multigpu!(f, A::SwitchingStencilArray{R,T,N,<:Array}, nsteps::Integer; kw...) where {R,T,N}
    plan = SlabPlan(f, A; kw...)
    iterate!(plan, nsteps)
    copyto!(source(A), Array(plan))
    finalize(plan)
    return A
end

# update_boundary!(A) — src/array.jl:195-233
function Stencils.update_boundary!(A::AbstractStencilArray{R,T,N,<:B200Array}) where {R,T,N}
    (padding(A) isa Halo && !(boundary(A) isa Use)) || return A
    d, keep = make_desc(sum, parent(A), R, A, parent(A))
    GC.@preserve keep check(ccall((:sb200_update_halo, LIB), Int32, (Ref{Desc}, Ptr{Cvoid}, Ptr{Cvoid}), d, dataptr(parent(A)), C_NULL))
    return A
end

# scatterstencil!(f, op, dest, source) — src/scatterstencil.jl:36-45. `f` must be one of the value rules below.
struct ScatterWeights{W}; w::W; end          # val_k = w_k
struct ScatterCenterWeights{W}; w::W; end    # val_k = center(hood) * w_k
scatter_op(::typeof(+)) = Int32(0)
scatter_op(::typeof(max)) = Int32(1)
scatter_op(::typeof(min)) = Int32(2)
scatter_op(op) = throw(ArgumentError("unsupported scatter op $op: use +, max or min"))
function Stencils.scatterstencil!(f::Union{ScatterWeights,ScatterCenterWeights}, op, dst::B200Array{T,2},
                                  src::AbstractStencilArray{R,T,2,<:B200Array}) where {R,T}
    Stencils._checksizes((dst, src))
    d, keep = make_desc(sum, dst, 0, src, parent(src))
    w = collect(T, f.w isa Number ? fill(f.w, length(stencil(src))) : f.w)
    rule = f isa ScatterWeights ? Int32(0) : Int32(1)
    d[] = with(d[]; scatter_op=scatter_op(op), scatter_rule=rule, weights_host=Ptr{Cvoid}(pointer(w)), out_eltype=ELTYPE[T])
    GC.@preserve keep w check(ccall((:sb200_scatter, LIB), Int32, (Ref{Desc}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                                     d, dataptr(parent(src)), dataptr(dst), C_NULL))
    check(ccall((:sb200_stream_sync, LIB), Int32, (Ptr{Cvoid},), C_NULL))
    return dst
end

end # module
