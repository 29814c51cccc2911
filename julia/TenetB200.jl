# TenetB200.jl — Julia glue for libtnb200.so (the B200 contraction engine behind Tenet.jl's hot path).
#
# UNTESTED: the build image has no Julia and Muscle/Tangles/EinExprs are not vendored in the reference checkout
# (/root/reference/Project.toml:6-18,29-45).  Upstream-internal names are marked [UPSTREAM-RECALL].  The Python
# package tenet.jl_b200/ binds the same C symbols with ctypes and is what the parity tests exercise.
#
# Pattern: a storage type (B200Array) selects the backend, exactly like `adapt(ConcreteRArray, ψ)` in
# docs/src/manual/reactant.md:23-36; `binary_einsum` and `contract` get methods for tensors backed by it.
module TenetB200

using Adapt
using Muscle          # Tensor, Index, inds, parent, binary_einsum            [UPSTREAM-RECALL]
using Tangles         # GenericTensorNetwork, tensors, contract               [UPSTREAM-RECALL]
using EinExprs        # einexpr, EinExpr (head, args)                         [UPSTREAM-RECALL]

const lib = get(ENV, "TNB200_LIB", "libtnb200.so")

const TNB_C128, TNB_C64, TNB_F64, TNB_F32 = Int32(0), Int32(1), Int32(2), Int32(3)
dtype_code(::Type{ComplexF64}) = TNB_C128
dtype_code(::Type{ComplexF32}) = TNB_C64
dtype_code(::Type{Float64}) = TNB_F64
dtype_code(::Type{Float32}) = TNB_F32

struct TnbTensor
    buf::Ptr{Cvoid}
    offset_elems::Int64
    dtype::Int32
    rank::Int32
    extent::Ptr{Int64}
    stride_elems::Ptr{Int64}
    mode::Ptr{Int32}
    conj::Int32
end

mutable struct Context
    h::Ptr{Cvoid}
end

function Context(device::Integer=0)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    rc = @ccall lib.tnb_ctx_create(device::Cint, r::Ptr{Ptr{Cvoid}})::Cint
    rc == 0 || error("tnb200: " * unsafe_string(@ccall lib.tnb_last_error(C_NULL::Ptr{Cvoid})::Cstring))
    c = Context(r[])
    finalizer(x -> (@ccall lib.tnb_ctx_destroy(x.h::Ptr{Cvoid})::Cint), c)
    return c
end

const _default = Ref{Union{Nothing,Context}}(nothing)
default_context() = something(_default[], (_default[] = Context(parse(Int, get(ENV, "LOCAL_RANK", "0")))))

function check(ctx::Context, rc)
    rc == 0 && return nothing
    msg = unsafe_string(@ccall lib.tnb_last_error(ctx.h::Ptr{Cvoid})::Cstring)
    rc == 1 ? throw(ArgumentError(msg)) : error("tnb200 error $rc: $msg")
end

# --- storage ------------------------------------------------------------------------------------------
mutable struct Buffer
    ctx::Context
    h::Ptr{Cvoid}
end

function Buffer(ctx::Context, bytes::Integer)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx, @ccall lib.tnb_alloc(ctx.h::Ptr{Cvoid}, max(bytes, 1)::Csize_t, r::Ptr{Ptr{Cvoid}})::Cint)
    b = Buffer(ctx, r[])
    # tnb_free only takes the context mutex and pushes to a free list: safe from finalizer threads
    finalizer(x -> (@ccall lib.tnb_free(x.ctx.h::Ptr{Cvoid}, x.h::Ptr{Cvoid})::Cint), b)
    return b
end

struct B200Array{T,N} <: AbstractArray{T,N}
    buf::Buffer
    dims::NTuple{N,Int}
    strides::NTuple{N,Int}     # element strides (column-major by default, like Array)
    offset::Int
end

Base.size(a::B200Array) = a.dims
Base.strides(a::B200Array) = a.strides
colmajor(dims) = ntuple(i -> prod(dims[1:(i - 1)]; init=1), length(dims))

function B200Array{T}(::UndefInitializer, dims::Int...; ctx=default_context()) where {T}
    B200Array{T,length(dims)}(Buffer(ctx, prod(dims; init=1) * sizeof(T)), dims, colmajor(dims), 0)
end

function upload(ctx::Context, a::Array{T,N}) where {T,N}
    d = B200Array{T}(undef, size(a)...; ctx)
    GC.@preserve a check(ctx, @ccall lib.tnb_upload(ctx.h::Ptr{Cvoid}, d.buf.h::Ptr{Cvoid}, 0::Csize_t,
                                                     pointer(a)::Ptr{Cvoid}, sizeof(a)::Csize_t)::Cint)
    check(ctx, @ccall lib.tnb_sync(ctx.h::Ptr{Cvoid})::Cint)    # `a` may be GC'd after return
    return d
end

function Base.Array(d::B200Array{T,N}) where {T,N}
    @assert d.strides == colmajor(d.dims) && d.offset == 0 "download of a strided view: copy it with binary_einsum first"
    a = Array{T,N}(undef, d.dims)
    check(d.buf.ctx, @ccall lib.tnb_download(d.buf.ctx.h::Ptr{Cvoid}, d.buf.h::Ptr{Cvoid}, 0::Csize_t,
                                             pointer(a)::Ptr{Cvoid}, sizeof(a)::Csize_t)::Cint)
    return a
end

# Lazy conjugate: `conj(tensor)` on device storage is a FLAG (tnb_tensor.conj), not a copy — overlap.jl:7,39 and
# DMRG.jl:80,96 materialise conj(ψ) on the CPU path; here the kernels conjugate at load time.
struct ConjB200{T,N} <: AbstractArray{T,N}
    parent::B200Array{T,N}
end
Base.size(a::ConjB200) = size(a.parent)
Base.conj(a::B200Array{<:Complex}) = ConjB200(a)
Base.conj(a::B200Array{<:Real}) = a
Base.conj(a::ConjB200) = a.parent
const B200Storage = Union{B200Array,ConjB200}
storage(a::B200Array) = (a, false)
storage(a::ConjB200) = (a.parent, true)

Adapt.adapt_storage(::Type{B200Array}, a::Array) = upload(default_context(), a)
Adapt.adapt_storage(::Type{Array}, a::B200Array) = Array(a)

# view(t, ind => i) / view(t, ind => a:b) stay metadata-only (compress.jl:46-58, evolve.jl:64-72)
function Base.view(a::B200Array{T,N}, I::Vararg{Union{Int,UnitRange{Int},Colon},N}) where {T,N}
    off, dims, st = a.offset, Int[], Int[]
    for (k, i) in enumerate(I)
        if i isa Int
            off += (i - 1) * a.strides[k]
        else
            r = i isa Colon ? (1:a.dims[k]) : i
            off += (first(r) - 1) * a.strides[k]
            push!(dims, length(r)); push!(st, a.strides[k])
        end
    end
    B200Array{T,length(dims)}(a.buf, Tuple(dims), Tuple(st), off)
end

# --- descriptors ----------------------------------------------------------------------------------------
struct Desc          # keeps the index arrays alive for the duration of a ccall
    t::TnbTensor
    keep::Tuple{Vector{Int64},Vector{Int64},Vector{Int32}}
end

Desc(a::ConjB200, labels::Vector{Int32}) = Desc(a.parent, labels; conj=true)
function Desc(a::B200Array{T}, labels::Vector{Int32}; conj::Bool=false) where {T}
    e, s = collect(Int64, a.dims), collect(Int64, a.strides)
    Desc(TnbTensor(a.buf.h, a.offset, dtype_code(T), length(e), pointer(e), pointer(s), pointer(labels), conj), (e, s, labels))
end

label_map(indlists...) = (m = Dict{Any,Int32}(); for l in indlists, i in l; get!(m, i, Int32(length(m))); end; m)

# --- Muscle.binary_einsum -----------------------------------------------------------------------------
# semantics fixed by the reference's call sites: default contracts all shared inds (overlap.jl:42,46), dims=Index[]
# keeps them as batch inds (canonize.jl:44, absorb.jl:31), rank-0 operands (DMRG.jl:60-61), rank-0 result (overlap.jl:49)
function Muscle.binary_einsum(a::Tensor{T,N,<:B200Storage}, b::Tensor{T,M,<:B200Storage};
                              dims=intersect(inds(a), inds(b)), out=nothing) where {T,N,M}
    ia, ib = collect(inds(a)), collect(inds(b))
    free_a = [i for i in ia if i ∉ ib && i ∉ dims]
    free_b = [i for i in ib if i ∉ ia && i ∉ dims]
    batch = [i for i in ia if i ∈ ib && i ∉ dims]
    ic = isnothing(out) ? vcat(free_a, free_b, batch) : collect(out)
    ext = merge(Dict(zip(ia, size(parent(a)))), Dict(zip(ib, size(parent(b)))))
    c = B200Array{T}(undef, (ext[i] for i in ic)...; ctx=storage(parent(a))[1].buf.ctx)
    m = label_map(ia, ib)
    da = Desc(parent(a), Int32[m[i] for i in ia])
    db = Desc(parent(b), Int32[m[i] for i in ib])
    dc = Desc(c, Int32[m[i] for i in ic])
    sm = Int32[m[i] for i in dims]
    ctx = c.buf.ctx
    GC.@preserve da db dc sm begin
        check(ctx, @ccall lib.tnb_binary_einsum(ctx.h::Ptr{Cvoid}, Ref(da.t)::Ptr{TnbTensor}, Ref(db.t)::Ptr{TnbTensor},
                                                Ref(dc.t)::Ptr{TnbTensor}, sm::Ptr{Int32}, length(sm)::Int32,
                                                C_NULL::Ptr{Cvoid}, C_NULL::Ptr{Cvoid})::Cint)
    end
    return Tensor(c, ic)
end

# --- contract(tn; path) ---------------------------------------------------------------------------------
# post-order walk of the EinExpr tree -> SSA pairs.  Leaves are matched to tensors by their index sets through a QUEUE
# per set: two tensors carrying identical index sets are interchangeable in the tree (either assignment contracts the
# same network), so they are handed out in order instead of colliding on one id (ADVICE r1).
function ssa_steps(path, leaf_queue::Dict)
    steps, next = Int32[], Ref(Int32(sum(length, values(leaf_queue); init=0)))
    function walk(node)
        isempty(node.args) && return popfirst!(leaf_queue[Set(node.head)])      # [UPSTREAM-RECALL] EinExpr fields
        ids = map(walk, node.args)
        acc = ids[1]
        for k in ids[2:end]
            push!(steps, acc, k); acc = next[]; next[] += 1
        end
        return acc
    end
    walk(path)
    return steps
end

# Own entry point (no method of Tangles.contract is overwritten — that would be type piracy): `TenetB200.contract(tn; ...)`.
# The one-line dispatch a maintainer adds upstream is in INTEGRATION.md §3 (a storage trait on the tensors of the network).
# `ngpus > 1` uses tnb_multi_contract_path: one host thread + context per device inside the library, slices dealt
# round-robin, one NCCL all-reduce — the single Julia task never needs an external launcher (SURVEY §8b).
function contract(tn; path=einexpr(tn), sliced=Index[], ngpus::Integer=1)
    ts = collect(tensors(tn))
    all(t -> parent(t) isa B200Storage, ts) || throw(ArgumentError("TenetB200.contract needs device-resident tensors: adapt(B200Array, tn) first"))
    T = eltype(parent(ts[1]))
    ctx = storage(parent(ts[1]))[1].buf.ctx
    m = label_map((inds(t) for t in ts)...)
    descs = [Desc(parent(t), Int32[m[i] for i in inds(t)]) for t in ts]
    leaf_queue = Dict{Any,Vector{Int32}}()
    for (k, t) in enumerate(ts)
        push!(get!(leaf_queue, Set(inds(t)), Int32[]), Int32(k - 1))
    end
    steps = ssa_steps(path, leaf_queue)
    iout = collect(path.head)
    ext = Dict(i => size(t, i) for t in ts for i in inds(t))
    out = B200Array{T}(undef, (ext[i] for i in iout)...; ctx)
    dout = Desc(out, Int32[m[i] for i in iout])
    sm = Int32[m[i] for i in sliced]
    nslices = prod((ext[i] for i in sliced); init=1)
    raw = [d.t for d in descs]
    GC.@preserve descs dout steps sm raw begin
        if ngpus > 1
            check(ctx, @ccall lib.tnb_multi_contract_path(ctx.h::Ptr{Cvoid}, raw::Ptr{TnbTensor}, length(raw)::Int32,
                                                          steps::Ptr{Int32}, (length(steps) ÷ 2)::Int32, sm::Ptr{Int32},
                                                          length(sm)::Int32, Ref(dout.t)::Ptr{TnbTensor}, ngpus::Int32)::Cint)
        else
            check(ctx, @ccall lib.tnb_contract_path(ctx.h::Ptr{Cvoid}, raw::Ptr{TnbTensor}, length(raw)::Int32,
                                                    steps::Ptr{Int32}, (length(steps) ÷ 2)::Int32, sm::Ptr{Int32},
                                                    length(sm)::Int32, 0::Int64, 1::Int64, nslices::Int64,
                                                    Ref(dout.t)::Ptr{TnbTensor})::Cint)
        end
    end
    return Tensor(out, iout)
end

# --- Muscle.tensor_qr_thin / tensor_svd_thin on device storage (canonize.jl:41,58,97; evolve.jl:62,92; DMRG.jl:338,437) ---
function Muscle.tensor_qr_thin(a::Tensor{T,N,<:B200Storage}; inds_q, inds_r=setdiff(inds(a), inds_q), ind_virtual) where {T,N}
    ia = collect(inds(a))
    ext = Dict(zip(ia, size(parent(a))))
    k = min(prod((ext[i] for i in inds_q); init=1), prod((ext[i] for i in inds_r); init=1))
    ctx = storage(parent(a))[1].buf.ctx
    q = B200Array{T}(undef, (ext[i] for i in inds_q)..., k; ctx)
    r = B200Array{T}(undef, k, (ext[i] for i in inds_r)...; ctx)
    m = label_map(ia, [ind_virtual])
    da = Desc(parent(a), Int32[m[i] for i in ia])
    dq = Desc(q, Int32[[m[i] for i in inds_q]; m[ind_virtual]])
    dr = Desc(r, Int32[m[ind_virtual]; [m[i] for i in inds_r]])
    rows = Int32[m[i] for i in inds_q]
    GC.@preserve da dq dr rows check(ctx, @ccall lib.tnb_qr_thin(ctx.h::Ptr{Cvoid}, Ref(da.t)::Ptr{TnbTensor}, rows::Ptr{Int32},
        length(rows)::Int32, m[ind_virtual]::Int32, Ref(dq.t)::Ptr{TnbTensor}, Ref(dr.t)::Ptr{TnbTensor})::Cint)
    return Tensor(q, [collect(inds_q); ind_virtual]), Tensor(r, [ind_virtual; collect(inds_r)])
end

function Muscle.tensor_svd_thin(a::Tensor{T,N,<:B200Storage}; inds_u, inds_v=setdiff(inds(a), inds_u), ind_s) where {T,N}
    ia = collect(inds(a))
    ext = Dict(zip(ia, size(parent(a))))
    k = min(prod((ext[i] for i in inds_u); init=1), prod((ext[i] for i in inds_v); init=1))
    ctx = storage(parent(a))[1].buf.ctx
    u = B200Array{T}(undef, (ext[i] for i in inds_u)..., k; ctx)
    sv = B200Array{T}(undef, k; ctx)
    v = B200Array{T}(undef, k, (ext[i] for i in inds_v)...; ctx)
    m = label_map(ia, [ind_s])
    da = Desc(parent(a), Int32[m[i] for i in ia])
    du = Desc(u, Int32[[m[i] for i in inds_u]; m[ind_s]])
    ds = Desc(sv, Int32[m[ind_s]])
    dv = Desc(v, Int32[m[ind_s]; [m[i] for i in inds_v]])
    rows = Int32[m[i] for i in inds_u]
    GC.@preserve da du ds dv rows check(ctx, @ccall lib.tnb_svd_thin(ctx.h::Ptr{Cvoid}, Ref(da.t)::Ptr{TnbTensor}, rows::Ptr{Int32},
        length(rows)::Int32, m[ind_s]::Int32, Ref(du.t)::Ptr{TnbTensor}, Ref(ds.t)::Ptr{TnbTensor}, Ref(dv.t)::Ptr{TnbTensor})::Cint)
    return Tensor(u, [collect(inds_u); ind_s]), Tensor(sv, [ind_s]), Tensor(v, [ind_s; collect(inds_v)])
end

end # module
