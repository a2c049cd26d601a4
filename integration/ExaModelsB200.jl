# ExaModelsB200.jl — the reference-side binding of libexa_b200.so (include/exa_b200.h).
#
# UNVERIFIED: there is no Julia toolchain in the build image or on the GPU box, so this file has
# never been executed.  It is the stub a maintainer would add as `ext/ExaModelsB200.jl`, written
# against ExaModels v0.12.0: the extension point is `build_extension(c::ExaCore{T,VT,B}; prod)`
# (src/nlp.jl:898; KernelAbstractions method at ext/ExaModelsKernelAbstractions.jl:33-191) and the
# callback methods dispatch on the model's `E` parameter (`AbstractExaModel{T,VT,E}`, src/nlp.jl:702;
# KA methods ext:212-547).  Arrays are CUDA.jl `CuVector{Float64}`; every callback is one `ccall`.
module ExaModelsB200

import ExaModels, NLPModels
import ExaModels: Var, ParameterNode, DataSource, DataIndexed, Node1, Node2, Constant, Null, SumNode, ProdNode,
    VarSource, ParameterSource, ArgLeaf, SIMDFunction, Objective, Constraint, ConstraintAugmentation, ExaCore, AbstractExaModel
using CUDA

const LIB = get(ENV, "EXB_LIB", "libexa_b200.so")

struct B200Backend            # stored in ExaCore{T,VT,B}.backend
    device::Int
    rank::Int
    world::Int
end
B200Backend(; device = 0, rank = 0, world = 1) = B200Backend(device, rank, world)
ExaModels.convert_array(v, ::B200Backend) = CuArray(v)          # src/templates.jl:2-3
ExaModels.default_T(::B200Backend) = Float64

mutable struct B200Extension  # stored in ExaModel.ext
    handle::Ptr{Cvoid}
    keep::Vector{Any}         # iterator arrays referenced by the IR during exb_create
    compressed::Bool          # CompressedB200Model: jac / hess callbacks return the duplicate-free COO
end

check(rc) = rc == 0 || error("libexa_b200: ", unsafe_string(ccall((:exb_last_error, LIB), Cstring, ())))

# ---- IR emitter: walks the TYPE + fields of each pattern's tree (SURVEY.md Appendix A) -------------
const OP1 = Dict(f => i - 1 for (i, f) in enumerate(first.(ExaModels._UNIVARIATES)))   # src/functionlist.jl:6-60
const OP2 = Dict(f => i - 1 for (i, f) in enumerate(first.(ExaModels._BIVARIATES)))    # src/functionlist.jl:71-81
# SpecialFunctions extension (ext/functionlist.jl:6-126): codes continue after the base tables, in registration order
for (k, f) in enumerate((:erf, :erfc, :erfi, :erfcx, :digamma, :trigamma, :invdigamma, :gamma, :airyai, :airybi, :airyaiprime,
                         :airybiprime, :besselj0, :bessely0, :besselj1, :bessely1, :dawson, :erfinv, :erfcinv))
    OP1[f] = length(ExaModels._UNIVARIATES) + k - 1
end
OP2[:beta] = length(ExaModels._BIVARIATES); OP2[:logbeta] = length(ExaModels._BIVARIATES) + 1
const T_CONST_I, T_CONST_F, T_DATA_SELF, T_DATA_FIELD, T_VAR, T_PAR, T_NULL, T_OP1, T_OP2, T_VAL = 0:9

mutable struct Emitter
    rows::Vector{NTuple{4,Int64}}
    fields::Vector{Tuple{Int64,Int64}}      # (byte offset, type 0 i64 | 1 f64 | 2 i32 | 3 f32)
    eltype::Type
    isrange::Bool
end
push_row!(e, r) = (push!(e.rows, r); length(e.rows) - 1)
f64bits(v) = reinterpret(Int64, Float64(v))
ftype(::Type{Int64}) = 0; ftype(::Type{Float64}) = 1; ftype(::Type{Int32}) = 2; ftype(::Type{Float32}) = 3

emit!(e, v::Integer) = push_row!(e, (T_CONST_I, 0, 0, Int64(v)))
emit!(e, v::Real) = push_row!(e, (T_CONST_F, 0, 0, f64bits(v)))
emit!(e, ::Val{V}) where {V} = push_row!(e, (T_VAL, 0, 0, Int64(V)))
emit!(e, ::Constant{V}) where {V} = emit!(e, V)
emit!(e, n::Null) = push_row!(e, (T_NULL, 0, 0, f64bits(n.value === nothing ? 0.0 : n.value)))
emit!(e, n::Var) = push_row!(e, (T_VAR, emit!(e, n.i), 0, 0))
emit!(e, n::ParameterNode) = push_row!(e, (T_PAR, emit!(e, n.i), 0, 0))
emit!(e, n::Node1{F}) where {F} = push_row!(e, (T_OP1, emit!(e, n.inner), 0, OP1[Symbol(F.instance)]))
function emit!(e, n::Node2{F}) where {F}
    a = emit!(e, n.inner1); b = emit!(e, n.inner2)
    push_row!(e, (T_OP2, a, b, OP2[Symbol(F.instance)]))
end
# SumNode / ProdNode (src/graph.jl:520-567): the adjoint modes evaluate them as `reduce(+, ...)` / `reduce(*, ...)`, a LEFT
# FOLD of the registered binary operator, so the sparsity order is that of the chain ((c1 + c2) + c3) + ...; an empty tuple is
# the AdjointNull of zero(T) / one(T) (graph.jl:547-548,552-553)
function emit_fold!(e, inners, op::Symbol, empty)
    isempty(inners) && return push_row!(e, (T_NULL, 0, 0, f64bits(empty)))
    acc = emit!(e, inners[1])
    for k in 2:length(inners)
        b = emit!(e, inners[k])
        acc = push_row!(e, (T_OP2, acc, b, OP2[op]))
    end
    acc
end
emit!(e, n::SumNode) = emit_fold!(e, n.inners, :+, 0.0)
emit!(e, n::ProdNode) = emit_fold!(e, n.inners, :*, 1.0)
# sentinels that never survive into a built pattern tree (graph.jl:128,142): indexing them returns Var / ParameterNode;
# ArgLeaf (graph.jl:170) is substituted by a plain Real when the core is instantiated (graph.jl:148-169)
emit!(e, ::VarSource) = error("ExaModelsB200: a bare VarSource in a pattern tree (index it: x[i])")
emit!(e, ::ParameterSource) = error("ExaModelsB200: a bare ParameterSource in a pattern tree (index it: θ[i])")
emit!(e, n::ArgLeaf) = error("ExaModelsB200: ArgLeaf must be substituted (instantiate the recipe) before build_extension")
function emit!(e, n::Union{DataSource,DataIndexed})
    e.isrange && return push_row!(e, (T_DATA_SELF, 0, 0, 0))
    off, T = field_path(e.eltype, n)
    k = findfirst(==((off, ftype(T))), e.fields)
    k === nothing && (push!(e.fields, (off, ftype(T))); k = length(e.fields))
    push_row!(e, (T_DATA_FIELD, k - 1, 0, 0))
end
field_path(T, ::DataSource) = (0, T)
function field_path(T, n::DataIndexed{I,J}) where {I,J}     # J: Symbol or position (src/graph.jl:194-199)
    off, S = field_path(T, n.inner)
    k = J isa Symbol ? Base.fieldindex(S, J) : J
    (off + fieldoffset(S, k), fieldtype(S, k))
end

function emit_pattern!(words, bufs, p, kind, base_index)
    sf = p.f
    isrange = p.itr isa UnitRange
    e = Emitter(NTuple{4,Int64}[], Tuple{Int64,Int64}[], eltype(p.itr), isrange)
    tree = kind == 2 ? sf.f.second : sf.f                    # augmentation: f.f :: Pair(idx, expr)
    root = emit!(e, tree)
    # augmentation row index (src/nlp.jl:1986-2001): a node, an Int (`Pair{<:Integer}`, nlp.jl:1994-1997: every point adds to the
    # same row -- emitted as a CONST_I index expression), or a tuple of nodes linearised column-major over `dims`
    idx = kind == 2 ? (sf.f.first isa Tuple ? collect(sf.f.first) : Any[sf.f.first]) : Any[]
    idx_roots = [emit!(e, i) for i in idx]     # emit!(::Integer) -> CONST_I
    append!(words, (kind, length(p.itr)))
    if isrange
        append!(words, (0, first(p.itr), -1, 0))
    else
        host = Array(p.itr); push!(bufs, host)
        append!(words, (1, 0, length(bufs) - 1, sizeof(eltype(host))))
    end
    push!(words, length(e.fields)); foreach(f -> append!(words, f), e.fields)
    append!(words, (kind == 2 ? -1 : sf.o0, sf.o1, sf.o2, base_index, length(idx_roots)))
    append!(words, idx_roots)
    append!(words, kind == 2 ? collect(Int64, p.dims)[1:length(idx_roots)] : Int64[])
    push!(words, length(e.rows)); foreach(r -> append!(words, r), e.rows)
    push!(words, root)
    push!(words, length(sf.comp1.inner)); append!(words, sf.comp1.inner)   # cross-checked by the probe
    push!(words, length(sf.comp2.inner)); append!(words, sf.comp2.inner)
end

function ExaModels.build_extension(c::ExaCore{T,VT,B}; prod = false, kwargs...) where {T,VT,B<:B200Backend}
    # `prod` needs no work here: exb_jprod / exb_jtprod / exb_hprod are fused into the derivative sweep (nothing to build);
    # `sorted_products = true` selects the reference's COO + sorted-structure SpMV (EXB_FLAG_SORTED_PRODUCTS = 2, bitwise
    # reproducible), whose structure is built on first use
    # patterns in ADD ORDER: both lists are stored newest-first (src/nlp.jl:536); the shared nnzh
    # counter (f.o2) orders objectives against constraints
    # ADD ORDER.  The core keeps two lists (objectives; constraints and augmentations), each newest-first (src/nlp.jl:536), so
    # the order WITHIN each list is the core's own.  Only the interleaving of the two is not recorded; it is recovered by a
    # stable two-way MERGE on the shared nnzh counter o2 (a pattern added later never has a smaller o2).  Ties -- patterns
    # without Hessian slots, e.g. the four (1,0) augmentations of AC-OPF -- keep their list order, and whichever way an
    # objective / constraint tie is broken gives the same o0 / o1 / o2 (they count in different buffers).
    objs = [(o, 0) for o in reverse(collect(c.obj))]
    cons = [(k, k isa ConstraintAugmentation ? 2 : 1) for k in reverse(collect(c.cons))]
    pats = Any[]
    io, ic = 1, 1
    while io <= length(objs) || ic <= length(cons)
        take_obj = ic > length(cons) || (io <= length(objs) && objs[io][1].f.o2 < cons[ic][1].f.o2)
        if take_obj; push!(pats, objs[io]); io += 1; else; push!(pats, cons[ic]); ic += 1; end
    end
    words = Int64[0x0031425845, 1, c.nvar, c.npar, length(pats), 0]
    bufs = Any[]
    for (p, kind) in pats
        base = kind == 2 ? findfirst(q -> q[2] == 1 && q[1].f.o0 == p.f.o0, pats) - 1 : -1
        emit_pattern!(words, bufs, p, kind, base)
    end
    words[6] = length(bufs)
    flags = Int32(get(kwargs, :sorted_products, false) ? 2 : 0) | Int32(get(kwargs, :tune_at_create, true) ? 4 : 0)
    x0 = Array(c.x0)   # EXB_FLAG_TUNE_AT_CREATE ranks the launch-shape variants at x0, so that no callback synchronises
    opt = Ref((Int32(c.backend.device), Int32(c.backend.rank), Int32(c.backend.world), flags, Int64(0), pointer(x0)))   # exb_options (ABI 2)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    ptrs = Ptr{Cvoid}[pointer(b) for b in bufs]
    GC.@preserve bufs x0 check(ccall((:exb_create, LIB), Cint,
        (Ptr{Int64}, Csize_t, Ptr{Ptr{Cvoid}}, Cint, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}),
        words, 8 * length(words), ptrs, length(ptrs), opt, h))
    ext = B200Extension(h[], bufs, false)
    finalizer(e -> ccall((:exb_destroy, LIB), Cint, (Ptr{Cvoid},), e.handle), ext)
    c.npar > 0 && check(ccall((:exb_set_params, LIB), Cint, (Ptr{Cvoid}, CuPtr{Float64}, Ptr{Cvoid}), h[], c.θ, stream().handle))
    if c.backend.world > 1 && get(kwargs, :comm_id, nothing) !== nothing
        # multi-GPU: one Julia process per GPU; rank 0 made the 128-byte id with exb_comm_unique_id and sent it to the others
        # (MPI.Bcast!); every rank passes it here.  From then on obj / grad! / cons! complete themselves inside the library.
        id = kwargs[:comm_id]::Vector{UInt8}
        check(ccall((:exb_comm_init, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}), h[], id))
        check(ccall((:exb_comm_set_mode, LIB), Cint, (Ptr{Cvoid}, Cint), h[], get(kwargs, :comm_owner, false) ? 1 : 0))
    end
    ext
end
comm_unique_id() = (id = zeros(UInt8, 128); check(ccall((:exb_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id)); id)

# ---- callbacks: same signatures / return conventions as src/nlp.jl:1798-1940 -------------------------
const M{T,VT} = AbstractExaModel{T,VT,B200Extension}
st() = stream().handle

function NLPModels.obj(m::M, x::AbstractVector)
    out = Ref{Float64}(0)
    check(ccall((:exb_obj, LIB), Cint, (Ptr{Cvoid}, CuPtr{Float64}, Ptr{Float64}, Ptr{Cvoid}), m.ext.handle, x, out, st()))
    out[]
end
function NLPModels.grad!(m::M, x::AbstractVector, g::AbstractVector)
    check(ccall((:exb_grad, LIB), Cint, (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, Ptr{Cvoid}), m.ext.handle, x, g, st())); g
end
function NLPModels.cons_nln!(m::M, x::AbstractVector, c::AbstractVector)
    check(ccall((:exb_cons, LIB), Cint, (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, Ptr{Cvoid}), m.ext.handle, x, c, st())); c
end
function NLPModels.jac_coord!(m::M, x::AbstractVector, vals::AbstractVector)
    check(ccall((:exb_jac, LIB), Cint, (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, Ptr{Cvoid}), m.ext.handle, x, vals, st())); vals
end
function NLPModels.hess_coord!(m::M, x::AbstractVector, y::AbstractVector, vals::AbstractVector; obj_weight = one(eltype(x)))
    check(ccall((:exb_hess, LIB), Cint, (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, Cdouble, CuPtr{Float64}, Ptr{Cvoid}),
        m.ext.handle, x, y, obj_weight, vals, st())); vals
end
function NLPModels.hess_coord!(m::M, x::AbstractVector, vals::AbstractVector; obj_weight = one(eltype(x)))   # nlp.jl:1906-1915
    check(ccall((:exb_hess, LIB), Cint, (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, Cdouble, CuPtr{Float64}, Ptr{Cvoid}),
        m.ext.handle, x, CU_NULL, obj_weight, vals, st())); vals
end
function NLPModels.jprod_nln!(m::M, x::AbstractVector, v::AbstractVector, Jv::AbstractVector)      # nlp.jl:1882-1897 | ext:353-420
    check(ccall((:exb_jprod, LIB), Cint, (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Ptr{Cvoid}), m.ext.handle, x, v, Jv, st())); Jv
end
function NLPModels.jtprod_nln!(m::M, x::AbstractVector, v::AbstractVector, Jtv::AbstractVector)    # nlp.jl:1899-1904 | ext:421-440
    check(ccall((:exb_jtprod, LIB), Cint, (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Ptr{Cvoid}), m.ext.handle, x, v, Jtv, st())); Jtv
end
function NLPModels.hprod!(m::M, x::AbstractVector, y::AbstractVector, v::AbstractVector, Hv::AbstractVector; obj_weight = one(eltype(x)))   # nlp.jl:1942-1978 | ext:441-481
    check(ccall((:exb_hprod, LIB), Cint, (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Cdouble, CuPtr{Float64}, Ptr{Cvoid}),
        m.ext.handle, x, y, v, obj_weight, Hv, st())); Hv
end
for (name, sym64, sym32) in ((:jac_structure!, :exb_jac_structure64, :exb_jac_structure32),
                             (:hess_structure!, :exb_hess_structure64, :exb_hess_structure32))
    @eval function NLPModels.$name(m::M, rows::CuVector{I}, cols::CuVector{I}) where {I<:Union{Int64,Int32}}
        sym = I === Int64 ? $(QuoteNode(sym64)) : $(QuoteNode(sym32))
        check(ccall((sym, LIB), Cint, (Ptr{Cvoid}, CuPtr{I}, CuPtr{I}, Ptr{Cvoid}), m.ext.handle, rows, cols, st()))
        rows, cols
    end
end

# ---- one sweep for all five callbacks (exb_eval; the composition of src/nlp.jl:1827-1940) ------------------------------------
function eval_all!(m::M, x, y, obj::CuVector{Float64}, g, c, jac, hess; obj_weight = one(eltype(x)))
    check(ccall((:exb_eval, LIB), Cint, (Ptr{Cvoid}, Cuint, CuPtr{Float64}, CuPtr{Float64}, Cdouble, CuPtr{Float64}, CuPtr{Float64},
        CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Ptr{Cvoid}), m.ext.handle, 31, x, y, obj_weight, obj, g, c, jac, hess, st()))
    obj, g, c, jac, hess
end

# ---- CompressedNLPModel role (src/utils.jl:425-579): duplicate-free COO straight from the library ------------------------------
# `CompressedB200Model(m)` wraps an ExaModel built on the B200 backend; nnzj / nnzh become the unique counts and the four
# structure / value callbacks go to the exb_*_compressed entry points (the Hessian of a shift-indexed model is emitted
# duplicate-free by ONE launch).  Everything else forwards to the inner model, as src/utils.jl:512-530 does.
struct CompressedB200Model{T,VT,MT<:AbstractExaModel{T,VT,B200Extension}} <: NLPModels.AbstractNLPModel{T,VT}
    inner::MT
    meta::NLPModels.NLPModelMeta{T,VT}
    counters::NLPModels.Counters
end
function CompressedB200Model(m::AbstractExaModel{T,VT,B200Extension}) where {T,VT}
    nj = Ref{Int64}(0); nh = Ref{Int64}(0)
    check(ccall((:exb_compressed_dims, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), m.ext.handle, nj, nh))
    meta = NLPModels.NLPModelMeta(m.meta.nvar; ncon = m.meta.ncon, nnzj = nj[], nnzh = nh[], x0 = m.meta.x0, lvar = m.meta.lvar,
        uvar = m.meta.uvar, y0 = m.meta.y0, lcon = m.meta.lcon, ucon = m.meta.ucon, minimize = m.meta.minimize)
    CompressedB200Model{T,VT,typeof(m)}(m, meta, NLPModels.Counters())
end
NLPModels.obj(m::CompressedB200Model, x::AbstractVector) = NLPModels.obj(m.inner, x)
NLPModels.grad!(m::CompressedB200Model, x::AbstractVector, g::AbstractVector) = NLPModels.grad!(m.inner, x, g)
NLPModels.cons_nln!(m::CompressedB200Model, x::AbstractVector, c::AbstractVector) = NLPModels.cons_nln!(m.inner, x, c)
function NLPModels.jac_coord!(m::CompressedB200Model, x::AbstractVector, vals::AbstractVector)
    check(ccall((:exb_jac_compressed, LIB), Cint, (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, Ptr{Cvoid}), m.inner.ext.handle, x, vals, st())); vals
end
function NLPModels.hess_coord!(m::CompressedB200Model, x::AbstractVector, y::AbstractVector, vals::AbstractVector; obj_weight = one(eltype(x)))
    check(ccall((:exb_hess_compressed, LIB), Cint, (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, Cdouble, CuPtr{Float64}, Ptr{Cvoid}),
        m.inner.ext.handle, x, y, obj_weight, vals, st())); vals
end
for (name, sym64, sym32) in ((:jac_structure!, :exb_jac_structure_compressed64, :exb_jac_structure_compressed32),
                             (:hess_structure!, :exb_hess_structure_compressed64, :exb_hess_structure_compressed32))
    @eval function NLPModels.$name(m::CompressedB200Model, rows::CuVector{I}, cols::CuVector{I}) where {I<:Union{Int64,Int32}}
        sym = I === Int64 ? $(QuoteNode(sym64)) : $(QuoteNode(sym32))
        check(ccall((sym, LIB), Cint, (Ptr{Cvoid}, CuPtr{I}, CuPtr{I}, Ptr{Cvoid}), m.inner.ext.handle, rows, cols, st()))
        rows, cols
    end
end

end # module
