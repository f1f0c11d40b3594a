# TensorNetworkEvolveB200.jl -- the reference-side binding of libtne_b200.so (include/tne.h).
#
# NOT EXECUTED in the build environment (no Julia runtime in the image or on the GPU boxes).  It mirrors, call for
# call, the ctypes binding `tensornetworkevolve.jl_b200/_lib.py` + `array.py` that the tests drive, and every `ccall`
# below is checked mechanically against the prototypes of include/tne.h by tools/check_julia_ccalls.py (a CPU test).
# A maintainer of TensorNetworkEvolve.jl adds this file, `include`s it at the end of src/TensorNetworkEvolve.jl and
# converts a PEPS with `to_b200(peps)`; every method below hooks the SAME dispatch point the in-tree plug-ins use
# (`OMEinsum.einsum(code, xs, size_dict)`: src/TensorAD/rules.jl:1-8, src/cracker.jl:4-8).
module TensorNetworkEvolveB200

using OMEinsum, LinearAlgebra
using Graphs: edges, src, dst, degree
using Yao: put, mat, time_evolve
import ..TensorNetworkEvolve
using ..TensorNetworkEvolve: cterm, xterm, zterm, PEPS, SimplePEPS, VidalPEPS, alltensors, replace_tensors
using ..TensorNetworkEvolve.TensorAD: DiffTensor, difftensor, debug_info, projectto
using OMEinsum: DynamicEinCode, StaticEinCode, NestedEinsum, SlicedEinsum, getixsv, getiyv

const LIB = Ref{String}("libtne_b200.so")
const CTX = Ref{Ptr{Cvoid}}(C_NULL)

function check(rc::Cint)
    rc == 0 && return
    throw(ErrorException(unsafe_string(ccall((:tne_last_error, LIB[]), Cstring, (Ptr{Cvoid},), CTX[]))))
end

function __init__()
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:tne_ctx_create, LIB[]), Cint, (Cint, Ptr{Ptr{Cvoid}}), parse(Cint, get(ENV, "LOCAL_RANK", "0")), ctx)
    rc == 0 || error("tne_ctx_create failed: there is no CPU fallback, a B200 is required")
    CTX[] = ctx[]
end

# ---- device array: slots into `vertex_tensors::Vector{<:AbstractArray{T}}` (src/peps.jl:41) and DiffTensor ------
# Storage on the device is always ComplexF64; `T` only records whether the Julia-side element type is real (norms,
# singular values), exactly like `is_real` in array.py.
mutable struct B200Array{T,N} <: AbstractArray{T,N}
    handle::Int64
    dims::NTuple{N,Int}
    function B200Array{T,N}(h::Int64, dims::NTuple{N,Int}) where {T,N}
        x = new{T,N}(h, dims)
        finalizer(a -> ccall((:tne_free, LIB[]), Cint, (Ptr{Cvoid}, Int64), CTX[], a.handle), x)
        return x
    end
end
Base.size(x::B200Array) = x.dims
Base.IndexStyle(::Type{<:B200Array}) = IndexLinear()
# scalar indexing works (AbstractArray fallbacks such as `≈`, `iterate`, `collect` rely on it) but downloads the whole
# tensor: the hot path never calls it, it is there so that generic code and the REPL do not hit a MethodError
Base.getindex(x::B200Array, i::Int) = Array(x)[i]
Base.getindex(x::B200Array{T,0}) where T = Array(x)[]
Base.show(io::IO, x::B200Array{T,N}) where {T,N} = print(io, "B200Array{$T,$N}(handle=$(x.handle), size=$(x.dims))")
Base.show(io::IO, ::MIME"text/plain", x::B200Array) = show(io, x)

function B200Array(a::AbstractArray{T,N}) where {T,N}
    host = Array{ComplexF64,N}(a)                     # interleaved ComplexF64, column-major
    h = Ref{Int64}(0)
    check(ccall((:tne_alloc, LIB[]), Cint, (Ptr{Cvoid}, Cint, Ptr{Int64}, Ptr{Int64}), CTX[], N, collect(Int64, size(a)), h))
    GC.@preserve host check(ccall((:tne_upload, LIB[]), Cint, (Ptr{Cvoid}, Int64, Ptr{Cvoid}), CTX[], h[], pointer(host)))
    return B200Array{T,N}(h[], size(a))
end

function Base.Array(x::B200Array{T,N}) where {T,N}
    host = Array{ComplexF64,N}(undef, size(x))
    GC.@preserve host check(ccall((:tne_download, LIB[]), Cint, (Ptr{Cvoid}, Int64, Ptr{Cvoid}), CTX[], x.handle, pointer(host)))
    return T <: Real ? real.(host) : host
end

# ---- OMEinsum.einsum: unary and binary (peps.jl:334,344,373,376,392-393,429; rules.jl:2-5; timeevolve.jl:38) -----
# `size_dict` is forwarded for labels that only the output carries (OMEinsum's repeat rule, e.g. the pullback of a trace)
function OMEinsum.einsum(code::Union{DynamicEinCode,StaticEinCode}, xs::NTuple{N,B200Array}, size_dict::Dict) where N
    ixs, iy = getixsv(code), getiyv(code)
    flat = Int32[l for ix in ixs for l in ix]
    ranks = Int32[length(ix) for ix in ixs]
    extra = Int32[l for l in unique(iy) if all(ix -> !(l in ix), ixs)]
    sizes = Int64[size_dict[l] for l in extra]
    h = Ref{Int64}(0)
    check(ccall((:tne_einsum_sized, LIB[]), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Int32}, Ptr{Int64}, Ptr{Int32}, Ptr{Int32}, Cint, Cint, Ptr{Int32}, Ptr{Int64}, Ptr{Int64}),
                CTX[], N, flat, ranks, Int64[x.handle for x in xs], zeros(Int32, N), Int32.(iy), length(iy),
                length(extra), extra, sizes, h))
    T = promote_type(map(eltype, xs)...)
    return B200Array{T,length(iy)}(h[], ntuple(k -> Int(size_dict[iy[k]]), length(iy)))
end

# ---- (::NestedEinsum)(xs...) / (::SlicedEinsum)(xs...) fast path (peps.jl:109,321): one tne_contract_tree call ----
struct TneTree          # layout of `tne_tree` in include/tne.h
    n_leaves::Int32; n_nodes::Int32
    leaf_rank::Ptr{Int32}; leaf_labels::Ptr{Int32}; node_children::Ptr{Int32}
    node_out_rank::Ptr{Int32}; node_out_labels::Ptr{Int32}
    n_sliced::Int32; sliced_labels::Ptr{Int32}
    shard_rank::Int32; shard_world::Int32
    seg_shard::Int32; n_segments::Int32
    seg_first_node::Ptr{Int32}; seg_sliced_off::Ptr{Int32}; seg_sliced_labels::Ptr{Int32}
end

function serialize(ne::NestedEinsum, leaf_ixs)
    children, out_rank, out_labels = Int32[], Int32[], Int32[]
    counter = Ref(length(leaf_ixs))
    function walk(nd)
        OMEinsum.isleaf(nd) && return Int32(nd.tensorindex - 1)
        ids = [walk(a) for a in nd.args]
        append!(children, ids); push!(out_rank, length(getiyv(nd.eins))); append!(out_labels, Int32.(getiyv(nd.eins)))
        me = counter[]; counter[] += 1
        return Int32(me)
    end
    walk(ne)
    return children, out_rank, out_labels
end

function (ne::NestedEinsum)(xs::B200Array...; slicing = Int32[], conj_flags = zeros(Int32, length(xs)))
    leaf_ixs = OMEinsum.getixsv(ne)
    children, out_rank, out_labels = serialize(ne, leaf_ixs)
    lr = Int32[length(ix) for ix in leaf_ixs]; ll = Int32[l for ix in leaf_ixs for l in ix]
    h = Ref{Int64}(0)
    GC.@preserve lr ll children out_rank out_labels slicing begin
        t = Ref(TneTree(length(xs), length(out_rank), pointer(lr), pointer(ll), pointer(children), pointer(out_rank),
                        pointer(out_labels), length(slicing), pointer(slicing), 0, 1, 0, 0, C_NULL, C_NULL, C_NULL))
        check(ccall((:tne_contract_tree, LIB[]), Cint, (Ptr{Cvoid}, Ptr{TneTree}, Ptr{Int64}, Ptr{Int32}, Ptr{Int64}),
                    CTX[], t, Int64[x.handle for x in xs], conj_flags, h))
    end
    iy = getiyv(ne.eins)
    sd = OMEinsum.get_size_dict(leaf_ixs, xs)
    return B200Array{ComplexF64,length(iy)}(h[], ntuple(k -> Int(sd[iy[k]]), length(iy)))
end
(se::SlicedEinsum)(xs::B200Array...) = se.eins(xs...; slicing = Int32.(se.slicing))

# ---- Yao.expect(op::Add, pa, pb) fused fast path (src/peps.jl:469-477): ONE tne_expect_terms call ---------------------
# The reference contracts the whole network once per term; the library shares sub-contractions between the terms
# (term dynamic programme) and, with grad_leaves, returns the pullbacks of the bra leaves from the same programme
# (this is what the `einsum` rule of TensorAD sees as one instruction; mirror of peps.py::_expect_fused).
struct TneTerm          # layout of `tne_term` in include/tne.h
    n_repl::Int32
    leaf::NTuple{4,Int32}
    repl::NTuple{4,Int64}
    coef_re::Cdouble; coef_im::Cdouble
end
# terms: Vector of (coef, [(site, 2x2 matrix), ...]) -- one-site-product terms (kron / put of one-site blocks);
# the replaced ket leaf of a term is apply_onsite(pb, site, matrix) (src/peps.jl:338-347), built with tne_einsum
function expect_terms(pa::PEPS, pb::PEPS, terms; grad_leaves = Int32[])
    code = pa.code_inner_product
    ne = code isa SlicedEinsum ? code.eins : code
    slicing = code isa SlicedEinsum ? Int32.(code.slicing) : Int32[]
    nt = length(alltensors(pa))
    bra, ket = alltensors(pa), alltensors(pb)
    leaves = Int64[t.handle for t in vcat(bra, ket)]
    conj_flags = vcat(ones(Int32, nt), zeros(Int32, nt))          # bra leaves enter conjugated (src/peps.jl:107)
    keep = B200Array[]                                             # replaced leaves must outlive the call
    cterms = TneTerm[]
    for (coef, repl) in terms
        ls, hs = zeros(Int32, 4), zeros(Int64, 4)
        for (k, (site, m)) in enumerate(repl)
            t = TensorNetworkEvolve.apply_onsite(pb, site, B200Array(ComplexF64.(m))).vertex_tensors[site]
            push!(keep, t); ls[k] = nt + site - 1; hs[k] = t.handle
        end
        push!(cterms, TneTerm(length(repl), Tuple(ls), Tuple(hs), real(coef), imag(coef)))
    end
    leaf_ixs = OMEinsum.getixsv(ne)
    children, out_rank, out_labels = serialize(ne, leaf_ixs)
    lr = Int32[length(ix) for ix in leaf_ixs]; ll = Int32[l for ix in leaf_ixs for l in ix]
    val = Ref{Int64}(0); grads = zeros(Int64, max(1, length(grad_leaves)))
    GC.@preserve lr ll children out_rank out_labels slicing keep begin
        t = Ref(TneTree(2nt, length(out_rank), pointer(lr), pointer(ll), pointer(children), pointer(out_rank),
                        pointer(out_labels), length(slicing), pointer(slicing), 0, 1, 0, 0, C_NULL, C_NULL, C_NULL))
        check(ccall((:tne_expect_terms, LIB[]), Cint,
                    (Ptr{Cvoid}, Ptr{TneTree}, Ptr{Int64}, Ptr{Int32}, Cint, Ptr{TneTerm}, Cint, Ptr{Int32}, Cdouble, Cdouble, Ptr{Int64}, Ptr{Int64}),
                    CTX[], t, leaves, conj_flags, length(cterms), cterms, length(grad_leaves), grad_leaves, 1.0, 0.0, val, grads))
    end
    value = B200Array{ComplexF64,0}(val[], ())
    return isempty(grad_leaves) ? value : (value, [B200Array{ComplexF64,ndims(bra[g + 1])}(grads[k], size(bra[g + 1])) for (k, g) in enumerate(grad_leaves)])
end

# ---- Base / broadcast methods TensorAD's rules and src/peps.jl:283-295 call on tensors -----------------------------
# tne_map_op of include/tne.h
const MAP = Dict(:copy => 0, :conj => 1, :real => 2, :imag => 3, :abs => 4, :sqrt => 5, :inv => 6, :neg => 7, :scale => 8, :pow => 9,
                 :sin => 10, :cos => 11, :safe_inv => 12)
function tne_map(op::Symbol, x::B200Array{T,N}, s::Number = 0.0; RT = T) where {T,N}
    h = Ref{Int64}(0)
    check(ccall((:tne_map, LIB[]), Cint, (Ptr{Cvoid}, Cint, Int64, Cdouble, Cdouble, Ptr{Int64}), CTX[], MAP[op], x.handle, real(s), imag(s), h))
    return B200Array{RT,N}(h[], size(x))
end
Base.conj(x::B200Array) = tne_map(:conj, x)
Base.copy(x::B200Array) = tne_map(:copy, x)
Base.real(x::B200Array{T}) where T = tne_map(:real, x; RT = real(T))
Base.imag(x::B200Array{T}) where T = tne_map(:imag, x; RT = real(T))
Base.zero(x::B200Array) = tne_map(:scale, x, 0.0)
Base.:*(x::B200Array{T}, c::Number) where T = tne_map(:scale, x, c; RT = promote_type(T, typeof(c)))
Base.:*(c::Number, x::B200Array) = x * c
Base.:/(x::B200Array, c::Number) = x * inv(c)
Base.:-(x::B200Array) = tne_map(:neg, x)
Base.Broadcast.broadcasted(::typeof(abs), x::B200Array{T}) where T = tne_map(:abs, x; RT = real(T))
Base.Broadcast.broadcasted(::typeof(sqrt), x::B200Array) = tne_map(:sqrt, x)
Base.Broadcast.broadcasted(::typeof(inv), x::B200Array) = tne_map(:inv, x)
Base.Broadcast.broadcasted(::typeof(sin), x::B200Array) = tne_map(:sin, x)
Base.Broadcast.broadcasted(::typeof(cos), x::B200Array) = tne_map(:cos, x)
Base.Broadcast.broadcasted(::typeof(conj), x::B200Array) = tne_map(:conj, x)
Base.Broadcast.broadcasted(::typeof(real), x::B200Array{T}) where T = tne_map(:real, x; RT = real(T))
Base.Broadcast.broadcasted(::typeof(^), x::B200Array, p::Real) = tne_map(:pow, x, p)
Base.Broadcast.broadcasted(::typeof(Base.literal_pow), ::typeof(^), x::B200Array, ::Val{p}) where p = tne_map(:pow, x, p)
Base.Broadcast.broadcasted(::Type{Complex}, x::B200Array{T}) where T = tne_map(:copy, x; RT = complex(T))      # rules.jl:158-162
Base.Broadcast.broadcasted(::typeof(*), x::B200Array, c::Number) = x * c
Base.Broadcast.broadcasted(::typeof(*), c::Number, x::B200Array) = x * c
# tne_map2_op of include/tne.h: ADD 0, SUB 1, MUL 2, DIV 3 (an operand with one element broadcasts)
function tne_map2(op::Int, a::B200Array, b::B200Array)
    h = Ref{Int64}(0)
    check(ccall((:tne_map2, LIB[]), Cint, (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Int64}), CTX[], op, a.handle, b.handle, h))
    big = length(a) >= length(b) ? a : b
    return B200Array{promote_type(eltype(a), eltype(b)),ndims(big)}(h[], size(big))
end
Base.:+(a::B200Array, b::B200Array) = tne_map2(0, a, b)
Base.:-(a::B200Array, b::B200Array) = tne_map2(1, a, b)
Base.:*(a::B200Array{T,0}, b::B200Array) where T = tne_map2(2, a, b)      # rules.jl:79-88
Base.:*(a::B200Array, b::B200Array{T,0}) where T = tne_map2(2, a, b)
Base.:*(a::B200Array{T1,0}, b::B200Array{T2,0}) where {T1,T2} = tne_map2(2, a, b)
Base.Broadcast.broadcasted(::typeof(*), a::B200Array, b::B200Array) = tne_map2(2, a, b)
Base.Broadcast.broadcasted(::typeof(/), a::B200Array, b::B200Array) = tne_map2(3, a, b)
Base.Broadcast.broadcasted(::typeof(sign), x::B200Array) = tne_map2(3, x, tne_map(:abs, x))                     # x ./ abs.(x), rules.jl:155-157
function Base.sum(x::B200Array{T}) where T                                   # rules.jl:91-93 (kept as a 0-dim device array)
    h = Ref{Int64}(0)
    check(ccall((:tne_reduce_sum, LIB[]), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}), CTX[], x.handle, h))
    return B200Array{T,0}(h[], ())
end
function Base.fill(x::B200Array{T,0}, dims::Integer...) where T            # rules.jl:94-96 without a host round trip
    h = Ref{Int64}(0)
    check(ccall((:tne_fill, LIB[]), Cint, (Ptr{Cvoid}, Int64, Cint, Ptr{Int64}, Ptr{Int64}), CTX[], x.handle, length(dims), collect(Int64, dims), h))
    return B200Array{T,length(dims)}(h[], Tuple(Int.(dims)))
end
function Base.reshape(x::B200Array{T}, dims::Dims{N}) where {T,N}
    h = Ref{Int64}(0)
    check(ccall((:tne_reshape, LIB[]), Cint, (Ptr{Cvoid}, Int64, Cint, Ptr{Int64}, Ptr{Int64}), CTX[], x.handle, N, collect(Int64, dims), h))
    return B200Array{T,N}(h[], dims)
end
Base.vec(x::B200Array) = reshape(x, (length(x),))
function Base.getindex(x::B200Array{T,1}, r::UnitRange{Int}) where T     # variables[a:b], src/peps.jl:217
    h = Ref{Int64}(0)
    check(ccall((:tne_slice, LIB[]), Cint, (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Int64}), CTX[], x.handle, first(r) - 1, length(r), h))
    return B200Array{T,1}(h[], (length(r),))
end
# accum(zero(x), (a:b,), dy) of rules.jl:46-55 as one device op
function TensorNetworkEvolve.TensorAD.accum(x::B200Array{T,1}, indices::Tuple{UnitRange{Int}}, y::B200Array) where T
    h = Ref{Int64}(0)
    r = indices[1]
    check(ccall((:tne_embed, LIB[]), Cint, (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Int64}), CTX[], y.handle, first(r) - 1, length(x), h))
    return x + B200Array{T,1}(h[], size(x))
end
function Base.vcat(xs::B200Array...)                                       # variables(peps), src/peps.jl:202
    h = Ref{Int64}(0)
    check(ccall((:tne_concat, LIB[]), Cint, (Ptr{Cvoid}, Cint, Ptr{Int64}, Ptr{Int64}), CTX[], length(xs), Int64[x.handle for x in xs], h))
    return B200Array{promote_type(map(eltype, xs)...),1}(h[], (sum(length, xs),))
end
OMEinsum.asarray(x::Number, ::B200Array) = B200Array(fill(x))
OMEinsum.asarray(x::B200Array, ::B200Array) = x

# The rules that read a 0-dim operand on the host (`a.data[] * b.data`, `fill(a.data[], size...)`: rules.jl:79-96) get
# device versions, so that the normalisation chain of iloss2 never synchronises.
function Base.:(*)(a::DiffTensor{T1,0,<:B200Array}, b::DiffTensor{T2,N,<:B200Array}) where {T1,T2,N}
    difftensor(a.data * b.data, debug_info("*", a, b), a => dy -> projectto(a, sum(dy .* conj(b))), b => dy -> projectto(b, dy * conj(a)))
end
function Base.:(*)(a::DiffTensor{T1,0,<:B200Array}, b::DiffTensor{T2,0,<:B200Array}) where {T1,T2}
    difftensor(a.data * b.data, debug_info("*", a, b), a => dz -> projectto(a, conj(b) * dz), b => dz -> projectto(b, conj(a) * dz))
end
function Base.fill(a::DiffTensor{T,0,<:B200Array}, size::Integer...) where T
    difftensor(fill(a.data, size...), debug_info("fill", a), a => dz -> sum(dz))
end

# ---- LinearAlgebra.svd on the bond matrix (src/peps.jl:381) and the fused apply_onbond! -------------------------------
function LinearAlgebra.svd(a::B200Array{T,2}; Dmax = minimum(size(a)), ϵ = 0.0) where T
    u, s, v = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0); kept = Ref{Int32}(0); err = Ref{Cdouble}(0)
    check(ccall((:tne_svd_trunc, LIB[]), Cint, (Ptr{Cvoid}, Int64, Cint, Cdouble, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int32}, Ptr{Cdouble}),
                CTX[], a.handle, Dmax, ϵ, u, s, v, kept, err))
    k = Int(kept[])
    return (U = B200Array{T,2}(u[], (size(a, 1), k)), S = B200Array{real(T),1}(s[], (k,)), V = B200Array{T,2}(v[], (size(a, 2), k)))
end

# ---- rydberg_tebd_sweep! (src/tebd.jl:15-23): the whole sweep in ONE call (tne_tebd_sweep) ---------------------------
# The library levels the sequential gate list into wavefronts of site-disjoint gates, enqueues them without host
# synchronisation (kept rank assumed = cap, verified once; exact-rank redo otherwise) and returns the new handles.
struct SweepGate                                   # mirrors tne_sweep_gate (include/tne.h); isbits, passed by pointer
    site_i::Int32; site_j::Int32; axis_i::Int32; axis_j::Int32; bond::Int32
    gate::NTuple{32,Cdouble}
    lambda_i::NTuple{32,Int32}; lambda_j::NTuple{32,Int32}
    kept::Int32; trunc_err::Cdouble
end
on_device(peps::PEPS) = all(t -> t isa B200Array, alltensors(peps))
# `PEPS{T,NF,LT}` has three parameters (src/peps.jl:4) and the array type is not one of them: dispatch on the abstract type and
# look at the tensors; anything that is not on the device takes the reference's own method (src/tebd.jl:15-23)
function sweep_gates(peps::PEPS, g, C::Real, Δ::Real, Ω::Real, δt)
    vidal = peps isa VidalPEPS
    bondidx = Dict(l => Int32(k - 1) for (k, l) in enumerate(peps.virtual_labels))
    gates = SweepGate[]
    for e in edges(g)                                # the reference's order (src/tebd.jl:16)
        i, j = src(e), dst(e)
        hij = cterm(C) + put(2, 1 => xterm(Ω / degree(g, i))) + put(2, 2 => xterm(Ω / degree(g, j))) +
              put(2, 1 => zterm(Δ / degree(g, i))) + put(2, 2 => zterm(Δ / degree(g, j)))      # src/tebd.jl:19
        m = Matrix(mat(time_evolve(hij, δt)))      # M[o_i + 2 o_j, i_i + 2 i_j], column-major
        li, lj = peps.vertex_labels[i], peps.vertex_labels[j]
        shared = only(li ∩ lj)
        lam(ls) = ntuple(a -> (vidal && 1 < a <= length(ls)) ? bondidx[ls[a]] : Int32(-1), 32)
        flat = reinterpret(Cdouble, vec(ComplexF64.(m)))
        push!(gates, SweepGate(i - 1, j - 1, findfirst(==(shared), li) - 1, findfirst(==(shared), lj) - 1, bondidx[shared],
                               ntuple(k -> flat[k], 32), lam(li), lam(lj), 0, 0.0))
    end
    return gates
end

# `n_sweeps` identical sweeps in ONE library call (tne_tebd_sweeps): every wavefront of every sweep is enqueued without a host
# synchronisation, the kept ranks are verified once at the end
function run_sweeps!(peps::PEPS{T}, g, gates::Vector{SweepGate}, n_sweeps::Integer) where T
    vidal = peps isa VidalPEPS
    old_sites = Int64[t.handle for t in peps.vertex_tensors]
    sites = copy(old_sites)
    old_bonds = vidal ? Int64[t.handle for t in peps.bond_tensors] : Int64[0]
    bonds = copy(old_bonds)
    errs = zeros(Cdouble, length(gates) * n_sweeps)
    check(ccall((:tne_tebd_sweeps, LIB[]), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Int64}, Cint, Ptr{Int64}, Cint, Ptr{SweepGate}, Cint, Cdouble, Cint, Cint, Ptr{Int32}, Ptr{Cdouble}),
                CTX[], length(sites), sites, vidal ? length(bonds) : 0, bonds, length(gates), gates, peps.Dmax, peps.ϵ, vidal,
                n_sweeps, C_NULL, errs))
    for err in errs
        err >= 0 && @info "truncation error is $(err)"                                         # src/peps.jl:386
    end
    # final shapes first (the last gate on a bond decides its rank), then exactly ONE new wrapper per tensor whose handle
    # changed: two wrappers over one handle would free it twice (mirror of peps.py::apply_sweep_)
    shapes = [collect(size(t)) for t in peps.vertex_tensors]
    blen = Dict{Int,Int}()
    for (q, e) in enumerate(edges(g))
        gq = gates[q]
        shapes[src(e)][gq.axis_i + 1] = gq.kept
        shapes[dst(e)][gq.axis_j + 1] = gq.kept
        blen[gq.bond + 1] = gq.kept
    end
    for s in eachindex(sites)
        sites[s] == old_sites[s] && continue
        peps.vertex_tensors[s] = B200Array{T,length(shapes[s])}(sites[s], Tuple(shapes[s]))
    end
    if vidal
        for b in eachindex(bonds)
            bonds[b] == old_bonds[b] && continue
            peps.bond_tensors[b] = B200Array{T,1}(bonds[b], (blen[b],))
        end
    end
    return peps
end

function TensorNetworkEvolve.rydberg_tebd_sweep!(peps::PEPS{T}, g, C::Real, Δ::Real, Ω::Real, δt) where T
    on_device(peps) || return invoke(TensorNetworkEvolve.rydberg_tebd_sweep!, Tuple{Any,Any,Real,Real,Real,Any}, peps, g, C, Δ, Ω, δt)
    return run_sweeps!(peps, g, sweep_gates(peps, g, C, Δ, Ω, δt), 1)
end

# rydberg_tebd! (src/tebd.jl:25-38): its nstep - 1 sweeps (sic, src/tebd.jl:27) in one call
function TensorNetworkEvolve.rydberg_tebd!(peps::PEPS{T}, g::SimpleGraph; t::Real, C::Real, Δ::Real, Ω::Real, nstep::Int) where T
    on_device(peps) || return invoke(TensorNetworkEvolve.rydberg_tebd!, Tuple{Any,SimpleGraph}, peps, g; t = t, C = C, Δ = Δ, Ω = Ω, nstep = nstep)
    nstep > 1 && run_sweeps!(peps, g, sweep_gates(peps, g, C, Δ, Ω, t / nstep), nstep - 1)
    return peps
end

# R replica PEPS of ONE structure through ONE tne_tebd_sweep call (mirror of peps.py::apply_sweep_batch_): the site / bond tables
# are concatenated and the gate list is tiled with offset indices; the library's wavefront levelling then puts the R copies of a
# gate into the same launch.  TEBD does not shard inside one PEPS: this is how it fills a GPU (and N GPUs: R / N replicas per rank).
function TensorNetworkEvolve.rydberg_tebd_sweep!(pepses::AbstractVector{<:PEPS{T}}, g, C::Real, Δ::Real, Ω::Real, δt) where T
    all(on_device, pepses) || return map(p -> TensorNetworkEvolve.rydberg_tebd_sweep!(p, g, C, Δ, Ω, δt), pepses)
    p0 = first(pepses)
    vidal = p0 isa VidalPEPS
    nv, nb, R = length(p0.vertex_tensors), length(p0.virtual_labels), length(pepses)
    bondidx = Dict(l => Int32(k - 1) for (k, l) in enumerate(p0.virtual_labels))
    one = SweepGate[]
    for e in edges(g)
        i, j = src(e), dst(e)
        hij = cterm(C) + put(2, 1 => xterm(Ω / degree(g, i))) + put(2, 2 => xterm(Ω / degree(g, j))) +
              put(2, 1 => zterm(Δ / degree(g, i))) + put(2, 2 => zterm(Δ / degree(g, j)))
        m = Matrix(mat(time_evolve(hij, δt)))
        li, lj = p0.vertex_labels[i], p0.vertex_labels[j]
        shared = only(li ∩ lj)
        lam(ls) = ntuple(a -> (vidal && 1 < a <= length(ls)) ? bondidx[ls[a]] : Int32(-1), 32)
        flat = reinterpret(Cdouble, vec(ComplexF64.(m)))
        push!(one, SweepGate(i - 1, j - 1, findfirst(==(shared), li) - 1, findfirst(==(shared), lj) - 1, bondidx[shared],
                             ntuple(k -> flat[k], 32), lam(li), lam(lj), 0, 0.0))
    end
    shift(t, d) = map(x -> x >= 0 ? Int32(x + d) : x, t)
    gates = SweepGate[SweepGate(q.site_i + (r - 1) * nv, q.site_j + (r - 1) * nv, q.axis_i, q.axis_j, q.bond + (r - 1) * nb, q.gate,
                                shift(q.lambda_i, (r - 1) * nb), shift(q.lambda_j, (r - 1) * nb), 0, 0.0) for r in 1:R for q in one]
    old_sites = Int64[t.handle for p in pepses for t in p.vertex_tensors]
    sites = copy(old_sites)
    old_bonds = vidal ? Int64[t.handle for p in pepses for t in p.bond_tensors] : Int64[0]
    bonds = copy(old_bonds)
    check(ccall((:tne_tebd_sweep, LIB[]), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Int64}, Cint, Ptr{Int64}, Cint, Ptr{SweepGate}, Cint, Cdouble, Cint),
                CTX[], length(sites), sites, vidal ? length(bonds) : 0, bonds, length(gates), gates, p0.Dmax, p0.ϵ, vidal))
    ng = length(one)
    for (r, peps) in enumerate(pepses)
        shapes = [collect(size(t)) for t in peps.vertex_tensors]
        blen = Dict{Int,Int}()
        for (q, e) in enumerate(edges(g))
            gq = gates[(r - 1) * ng + q]
            gq.trunc_err >= 0 && @info "truncation error is $(gq.trunc_err)"
            shapes[src(e)][gq.axis_i + 1] = gq.kept
            shapes[dst(e)][gq.axis_j + 1] = gq.kept
            blen[one[q].bond + 1] = gq.kept
        end
        for s in 1:nv
            k = (r - 1) * nv + s
            sites[k] == old_sites[k] && continue
            peps.vertex_tensors[s] = B200Array{T,length(shapes[s])}(sites[k], Tuple(shapes[s]))
        end
        if vidal
            for b in 1:nb
                k = (r - 1) * nb + b
                bonds[k] == old_bonds[k] && continue
                peps.bond_tensors[b] = B200Array{T,1}(bonds[k], (blen[b],))
            end
        end
    end
    return pepses
end

# compiled-programme cache of the tree entry points (tne_plan_cache_info): (programmes held, calls served from the cache)
function plan_cache_info()
    n, h = Ref{Int64}(0), Ref{Int64}(0)
    check(ccall((:tne_plan_cache_info, LIB[]), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), CTX[], n, h))
    return n[], h[]
end

to_b200(peps::PEPS) = replace_tensors(peps, B200Array.(alltensors(peps)))

end # module
