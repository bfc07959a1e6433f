# SciMLOperatorsB200Ext.jl -- package extension that routes the cached in-place `mul!` path of
# SciMLOperators.jl to libscimlop_b200.so for operators whose storage is a B200 device array.
#
# It sits next to ext/SciMLOperatorsLoopVectorizationExt.jl and is declared the same way in Project.toml:
#
#   [weakdeps]      CUDA = "052768ef-5323-5732-b1bb-66c8b64840ba"
#   [extensions]    SciMLOperatorsB200Ext = "CUDA"
#
# NOTE: Julia is not installed in the build image of this repository, so this file has never been executed
# there; the Python mirror (scimloperators.jl_b200/operators.py) is its executable twin and is what the
# parity tests drive.  Every `ccall` below matches include/scimlop_b200.h one to one.
module SciMLOperatorsB200Ext

using SciMLOperators
using SciMLOperators: TensorProductOperator, MatrixOperator, IdentityOperator, ScaledOperator,
                      AddedOperator, ComposedOperator, BatchedDiagonalOperator, getops, iscached
import SciMLOperators: has_tensor_outer_mul_fast, tensor_outer_mul_fast!, cache_self, update_coefficients!
import LinearAlgebra
import LinearAlgebra: mul!, Diagonal, Adjoint, Transpose
using CUDA

const LIB = get(ENV, "SCIMLOP_B200_LIB", "libscimlop_b200.so")
const DevMat{T} = Union{CuMatrix{T}, CuVector{T}}
const F = Union{Float32, Float64}

dtype(::Type{Float64}) = Cint(0)
dtype(::Type{Float32}) = Cint(1)
const DENSE, IDENTITY, DIAG, BDIAG, NULLOP, TRANS = Cint(0), Cint(1), Cint(2), Cint(3), Cint(4), Cint(0x100)

# ---- handle: one per device, created lazily ------------------------------------------------------------
const HANDLES = Dict{Int, Ptr{Cvoid}}()
function handle()
    dev = CUDA.deviceid(CUDA.device())
    get!(HANDLES, dev) do
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:scimlop_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint), h, dev)
        rc == 0 || error("scimlop_create: " * unsafe_string(ccall((:scimlop_status_string, LIB), Cstring, (Cint,), rc)))
        h[]
    end
end
function check(rc::Cint, what)
    rc == 0 && return
    buf = Vector{UInt8}(undef, 512)
    ccall((:scimlop_last_error, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Csize_t), handle(), buf, 512)
    error("$what: status $rc: " * unsafe_string(pointer(buf)))
end
stream_ptr() = Base.unsafe_convert(Ptr{Cvoid}, CUDA.stream().handle)
dptr(x) = reinterpret(Ptr{Cvoid}, pointer(x))

# ---- the reference's own plugin hook (src/tensor.jl:693,719): outer stage of a TPO ------------------------
has_tensor_outer_mul_fast(::MatrixOperator{T, <:CuMatrix{T}}) where {T <: F} = true

function tensor_outer_mul_fast!(w::DevMat{T}, outer::MatrixOperator{T, <:CuMatrix{T}}, C::DevMat{T},
                                mi, mo, no, k) where {T <: F}
    check(ccall((:scimlop_outer_mul_fast, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Int64, Int64, Int64,
                 Ptr{T}, Ptr{T}, Ptr{Cvoid}),
                handle(), dtype(T), 0, dptr(w), dptr(outer.A), stride(outer.A, 2), dptr(C), mi, mo, no, k,
                C_NULL, C_NULL, stream_ptr()), "scimlop_outer_mul_fast")
    w
end
function tensor_outer_mul_fast!(w::DevMat{T}, outer::MatrixOperator{T, <:CuMatrix{T}}, C::DevMat{T},
                                mi, mo, no, k, α, β) where {T <: F}
    check(ccall((:scimlop_outer_mul_fast, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Int64, Int64, Int64,
                 Ref{T}, Ref{T}, Ptr{Cvoid}),
                handle(), dtype(T), 0, dptr(w), dptr(outer.A), stride(outer.A, 2), dptr(C), mi, mo, no, k,
                T(α), T(β), stream_ptr()), "scimlop_outer_mul_fast")
    w
end

# ---- whole-TPO path: flatten nested products into native factors --------------------------------------------
# returns (kinds, ptrs, dims, lds, λ) or nothing when some factor is not native (then the generic
# src/tensor.jl code runs and only the hook above is used).
function flatten(L::TensorProductOperator, T)
    kinds = Cint[]; ptrs = Ptr{Cvoid}[]; dims = Int64[]; lds = Int64[]; λ = one(T)
    function walk(op)
        if op isa ScaledOperator
            λ *= convert(Number, op.λ); return walk(op.L)
        elseif op isa TensorProductOperator
            return all(walk, op.ops)
        elseif op isa IdentityOperator
            push!(kinds, IDENTITY); push!(ptrs, C_NULL); push!(lds, 0)
        elseif op isa MatrixOperator && op.A isa CuMatrix{T}
            push!(kinds, DENSE); push!(ptrs, dptr(op.A)); push!(lds, stride(op.A, 2))
        elseif op isa MatrixOperator && op.A isa Union{Adjoint{T, <:CuMatrix{T}}, Transpose{T, <:CuMatrix{T}}}
            push!(kinds, DENSE | TRANS); push!(ptrs, dptr(parent(op.A))); push!(lds, stride(parent(op.A), 2))
        elseif op isa MatrixOperator && op.A isa Diagonal{T, <:CuVector{T}}
            push!(kinds, DIAG); push!(ptrs, dptr(op.A.diag)); push!(lds, 0)
        else
            return false
        end
        append!(dims, size(op)); true
    end
    walk(L) ? (kinds, ptrs, dims, lds, λ) : nothing
end

# `cache_self` (src/tensor.jl:460-518): the native path needs at most two ping-pong buffers; they are stored
# in the first cache slot so that `iscached(L)` and `copy(L)` keep working unchanged.
function cache_self(L::TensorProductOperator, v::DevMat{T}) where {T <: F}
    fl = flatten(L, T)
    fl === nothing && return invoke(cache_self, Tuple{TensorProductOperator, AbstractVecOrMat}, L, v)
    kinds, _, dims, _, _ = fl
    nbytes = Ref{Csize_t}(0)
    rc = ccall((:scimlop_kron_workspace_bytes, LIB), Cint, (Cint, Cint, Ptr{Int64}, Ptr{Cint}, Int64, Ref{Csize_t}),
               dtype(T), length(kinds), dims, kinds, size(v, 2), nbytes)
    rc == 0 || error("scimlop_kron_workspace_bytes: status $rc")
    ws = CUDA.zeros(UInt8, max(nbytes[], 1))
    TensorProductOperator(L.ops, (ws, nothing, nothing, nothing, nothing, nothing, nothing))
end

for (args, αβ) in (((), (:(C_NULL), :(C_NULL))), ((:α, :β), (:(Ref(T(α))), :(Ref(T(β))))))
    @eval function mul!(w::DevMat{T}, L::TensorProductOperator, v::DevMat{T}, $(args...)) where {T <: F}
        @assert iscached(L) "cache needs to be set up for operator of type $(typeof(L)). Set up cache by calling `cache_operator(L, v)`"
        fl = flatten(L, T)
        fl === nothing && return invoke(mul!, Tuple{AbstractVecOrMat, TensorProductOperator, AbstractVecOrMat, $(map(_ -> Any, args)...)}, w, L, v, $(args...))
        kinds, ptrs, dims, lds, λ = fl
        ws = L.cache[1]
        a, b = $(αβ[1]), $(αβ[2])
        λ == one(T) || (a = Ref(T(λ) * ($(isempty(args)) ? one(T) : T(α))))
        check(ccall((:scimlop_kron_mul, LIB), Cint,
                    (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Ptr{Cvoid}}, Ptr{Int64}, Ptr{Int64}, Ptr{Cint}, Ptr{Cvoid}, Int64,
                     Ptr{Cvoid}, Int64, Int64, Ptr{T}, Ptr{T}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
                    handle(), dtype(T), 0, length(kinds), ptrs, dims, lds, kinds, dptr(v), size(v, 1), dptr(w),
                    size(w, 1), size(v, 2), a, b, dptr(ws), sizeof(ws), stream_ptr()), "scimlop_kron_mul")
        w
    end
end

# ---- leaves over device storage (src/matrix.jl:337-350, src/batch.jl:151-179) -------------------------------
# Methods are added for SciMLOperators' OWN types only (MatrixOperator over a CuMatrix, or over its lazy Adjoint /
# Transpose -- `adjoint(::MatrixOperator)` of a constant operator wraps the array, src/matrix.jl:174-175).  A method on
# the bare `mul!(::CuMatrix, ::CuMatrix, ::CuMatrix, α, β)` would be type piracy over CUDA.jl's CUBLAS method.
const DenseDev{T} = Union{CuMatrix{T}, Adjoint{T, <:CuMatrix{T}}, Transpose{T, <:CuMatrix{T}}}
_stored(A::CuMatrix) = (A, 0)
_stored(A::Union{Adjoint, Transpose}) = (parent(A), 1)

function mul!(w::DevMat{T}, L::MatrixOperator{T, <:DenseDev{T}}, v::DevMat{T}, α::Number, β::Number) where {T <: F}
    A, trans = _stored(L.A)
    check(ccall((:scimlop_matmul, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Cint, Int64, Int64, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64,
                 Ref{T}, Ref{T}, Ptr{Cvoid}),
                handle(), dtype(T), 0, trans, size(L, 1), size(L, 2), size(v, 2), dptr(A), stride(A, 2), dptr(v), size(v, 1),
                dptr(w), size(w, 1), T(α), T(β), stream_ptr()), "scimlop_matmul")
    w
end
mul!(w::DevMat{T}, L::MatrixOperator{T, <:DenseDev{T}}, v::DevMat{T}) where {T <: F} = mul!(w, L, v, true, false)

# `cache_operator` of a large Float32 MatrixOperator prepares the hi / lo image of the streamed tensor-core GEMM once
# (include/scimlop_b200.h: scimlop_split_attach), so that mul! stays allocation-free (test/Downstream/alloccheck.jl).
const SPLIT_IMAGES = IdDict{Any, Any}()
function cache_self(L::MatrixOperator{Float32, <:DenseDev{Float32}}, v::DevMat{Float32})
    A, _ = _stored(L.A)
    (minimum(size(A)) < 16 || maximum(size(A)) <= 128 || haskey(SPLIT_IMAGES, A)) && return L
    nbytes = Ref{Csize_t}(0)
    ccall((:scimlop_split_bytes, LIB), Cint, (Int64, Int64, Ref{Csize_t}), size(A, 1), size(A, 2), nbytes)
    img = CUDA.zeros(UInt8, nbytes[])
    check(ccall((:scimlop_split_attach, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
                handle(), dptr(A), stride(A, 2), size(A, 1), size(A, 2), dptr(img), nbytes[], stream_ptr()), "scimlop_split_attach")
    SPLIT_IMAGES[A] = img
    finalizer(a -> (ccall((:scimlop_split_detach, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), handle(), dptr(a)); delete!(SPLIT_IMAGES, a)), A)
    L
end
function update_coefficients!(L::MatrixOperator{Float32, <:DenseDev{Float32}}, u, p, t; kwargs...)
    invoke(update_coefficients!, Tuple{MatrixOperator, Any, Any, Any}, L, u, p, t; kwargs...)
    A, _ = _stored(L.A)
    haskey(SPLIT_IMAGES, A) && check(ccall((:scimlop_split_refresh, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                                           handle(), dptr(A), stream_ptr()), "scimlop_split_refresh")
    L
end

function mul!(w::DevMat{T}, L::BatchedDiagonalOperator, v::DevMat{T}, α::Number, β::Number) where {T <: F}
    check(ccall((:scimlop_diag_mul, LIB), Cint,
                (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Ref{T}, Ref{T}, Ptr{Cvoid}),
                handle(), dtype(T), size(v, 1), size(v, 2), dptr(L.diag), size(L.diag, 1), dptr(v), size(v, 1), dptr(w),
                size(w, 1), T(α), T(β), stream_ptr()), "scimlop_diag_mul")
    w
end

# ---- trees: AddedOperator / ComposedOperator / ScaledOperator -> one scimlop_tree_mul call -----------------
# A tree over native leaves is handed over flattened: a sum of terms, each term = (product of ScalarOperator
# values) * leaf_1 * leaf_2 * ... (what src/basic.jl:619-635 does for sums and :936-939 for products).
# The C structs are `scimlop_factor_t` / `scimlop_term_t` of include/scimlop_b200.h.
struct CFactor
    kind::Cint; trans::Cint; data::Ptr{Cvoid}; ld::Int64; rows::Int64; cols::Int64
end
struct CTerm
    scale::Cdouble; nfactors::Cint; factors::Ptr{CFactor}
end

# leaf -> CFactor, or nothing if the leaf is not native
function cfactor(op, T)
    m, n = size(op)
    if op isa IdentityOperator
        return CFactor(IDENTITY, 0, C_NULL, 0, m, n)
    elseif op isa SciMLOperators.NullOperator
        return CFactor(NULLOP, 0, C_NULL, 0, m, n)
    elseif op isa BatchedDiagonalOperator && op.diag isa CuArray{T}
        return CFactor(BDIAG, 0, dptr(op.diag), size(op.diag, 1), m, n)
    elseif op isa MatrixOperator
        A = op.A
        A isa CuMatrix{T} && return CFactor(DENSE, 0, dptr(A), stride(A, 2), m, n)
        A isa Union{Adjoint{T, <:CuMatrix{T}}, Transpose{T, <:CuMatrix{T}}} &&
            return CFactor(DENSE, 1, dptr(parent(A)), stride(parent(A), 2), m, n)
        A isa Diagonal{T, <:CuVector{T}} && return CFactor(DIAG, 0, dptr(A.diag), 0, m, n)
    end
    return nothing
end

# [(scalars::Vector, leaves::Vector{CFactor})] or nothing; a sum inside a product is not distributed
function flatten_tree(L, T, scalars = Any[])
    if L isa ScaledOperator
        return flatten_tree(L.L, T, vcat(scalars, Any[L.λ]))
    elseif L isa AddedOperator
        out = Tuple{Vector{Any}, Vector{CFactor}}[]
        for op in L.ops
            t = flatten_tree(op, T, scalars)
            t === nothing && return nothing
            append!(out, t)
        end
        return out
    elseif L isa ComposedOperator
        sc = copy(scalars); leaves = CFactor[]
        for op in L.ops
            t = flatten_tree(op, T, Any[])
            (t === nothing || length(t) != 1) && return nothing
            append!(sc, t[1][1]); append!(leaves, t[1][2])
        end
        return [(sc, leaves)]
    else
        f = cfactor(L, T)
        return f === nothing ? nothing : [(copy(scalars), CFactor[f])]
    end
end

# The plan lives next to the operator's own cache (a WeakKeyDict keeps `copy`/`adapt` semantics untouched):
# factor arrays, the term array and the device workspace are built once in `cache_operator`; `mul!` only
# refreshes the scales (ScalarOperators may have been updated, src/scalar.jl:267-270) and ccalls.
mutable struct TreePlan{T}
    terms::Vector{CTerm}
    keep::Vector{Vector{CFactor}}
    scalars::Vector{Vector{Any}}
    ws::CuVector{UInt8}
    k::Int
end
const PLANS = WeakKeyDict{Any, Any}()
const TreeOp = Union{AddedOperator, ComposedOperator, ScaledOperator}

function build_plan(L::TreeOp, v::DevMat{T}) where {T <: F}
    fl = flatten_tree(L, T)
    fl === nothing && return nothing
    keep = [t[2] for t in fl]
    terms = [CTerm(1.0, length(f), pointer(f)) for f in keep]
    nbytes = Ref{Csize_t}(0)
    rc = ccall((:scimlop_tree_workspace_bytes, LIB), Cint, (Cint, Ptr{CTerm}, Cint, Int64, Ref{Csize_t}),
               dtype(T), terms, length(terms), size(v, 2), nbytes)
    rc == 0 || error("scimlop_tree_workspace_bytes: status $rc")
    TreePlan{T}(terms, keep, [t[1] for t in fl], CUDA.zeros(UInt8, max(nbytes[], 1)), size(v, 2))
end

function SciMLOperators.cache_operator(L::TreeOp, v::DevMat{T}) where {T <: F}
    Lc = invoke(SciMLOperators.cache_operator, Tuple{SciMLOperators.AbstractSciMLOperator, AbstractVecOrMat}, L, v)
    plan = build_plan(Lc, v)
    plan === nothing || (PLANS[Lc] = plan)
    Lc
end

function tree_mul!(w::DevMat{T}, L::TreeOp, plan::TreePlan{T}, v::DevMat{T}, a, b) where {T <: F}
    for (i, sc) in enumerate(plan.scalars)      # refresh λ's: allocation-free (isbits CTerm, in-place setindex!)
        λ = 1.0
        for s in sc; λ *= Float64(convert(Number, s)); end
        plan.terms[i] = CTerm(λ, plan.terms[i].nfactors, plan.terms[i].factors)
    end
    GC.@preserve plan begin
        check(ccall((:scimlop_tree_mul, LIB), Cint,
                    (Ptr{Cvoid}, Cint, Cint, Ptr{CTerm}, Cint, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Int64, Ptr{T}, Ptr{T},
                     Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
                    handle(), dtype(T), 0, plan.terms, length(plan.terms), dptr(v), size(v, 1), dptr(w), size(w, 1),
                    size(v, 2), a, b, dptr(plan.ws), sizeof(plan.ws), stream_ptr()), "scimlop_tree_mul")
    end
    w
end

for Op in (:AddedOperator, :ComposedOperator, :ScaledOperator)
    @eval function mul!(w::DevMat{T}, L::$Op, v::DevMat{T}) where {T <: F}
        plan = get(PLANS, L, nothing)
        (plan === nothing || plan.k != size(v, 2)) &&
            return invoke(mul!, Tuple{AbstractVecOrMat, $Op, AbstractVecOrMat}, w, L, v)   # generic src/basic.jl path
        tree_mul!(w, L, plan, v, C_NULL, C_NULL)
    end
    @eval function mul!(w::DevMat{T}, L::$Op, v::DevMat{T}, α, β) where {T <: F}
        plan = get(PLANS, L, nothing)
        (plan === nothing || plan.k != size(v, 2)) &&
            return invoke(mul!, Tuple{AbstractVecOrMat, $Op, AbstractVecOrMat, Any, Any}, w, L, v, α, β)
        tree_mul!(w, L, plan, v, Ref(T(α)), Ref(T(β)))
    end
end

# ---- TensorSumOperator (src/tensor.jl:387-403) ----------------------------------------------------------------
function mul!(w::DevMat{T}, L::SciMLOperators.TensorSumOperator, v::DevMat{T}, α, β) where {T <: F}
    outer, inner = L.ops
    (outer isa MatrixOperator && inner isa MatrixOperator && outer.A isa CuMatrix{T} && inner.A isa CuMatrix{T}) ||
        return invoke(mul!, Tuple{AbstractVecOrMat, SciMLOperators.TensorSumOperator, AbstractVecOrMat, Any, Any}, w, L, v, α, β)
    check(ccall((:scimlop_kronsum_mul, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64,
                 Int64, Ref{T}, Ref{T}, Ptr{Cvoid}),
                handle(), dtype(T), 0, dptr(outer.A), stride(outer.A, 2), size(outer.A, 1), dptr(inner.A), stride(inner.A, 2),
                size(inner.A, 1), dptr(v), size(v, 1), dptr(w), size(w, 1), size(v, 2), T(α), T(β), stream_ptr()),
          "scimlop_kronsum_mul")
    w
end

# ---- BlockDiagonalOperator (src/block.jl:145-179): the whole operator in one call ------------------------------
function mul!(w::DevMat{T}, L::SciMLOperators.BlockDiagonalOperator, v::DevMat{T}, α, β) where {T <: F}
    blocks = CFactor[]
    for op in L.ops                       # only native leaves; anything else keeps the generic loop
        f = cfactor(op, T)
        f === nothing && return invoke(mul!, Tuple{AbstractVecOrMat, SciMLOperators.BlockDiagonalOperator, AbstractVecOrMat, Any, Any}, w, L, v, α, β)
        push!(blocks, f)
    end
    check(ccall((:scimlop_blockdiag_mul, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{CFactor}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Int64, Ref{T}, Ref{T}, Ptr{Cvoid}),
                handle(), dtype(T), 0, length(blocks), blocks, dptr(v), stride(v, 2), dptr(w), stride(w, 2), size(v, 2),
                T(α), T(β), stream_ptr()), "scimlop_blockdiag_mul")
    w
end

# ---- the solve side: factorize / ldiv! (src/matrix.jl:513-516, src/tensor.jl:227, :612-673) ---------------------
# `factorize(::MatrixOperator)` over device storage caches the explicit inverse, formed on the device by
# scimlop_invert; it is wrapped as the `F` of an InvertibleOperator, whose `ldiv!(w, L.F, v)` (src/matrix.jl:632-639)
# then dispatches to the `mul!` above.  A factorised TensorProductOperator therefore solves with scimlop_kron_mul
# on the inverse factors -- the same two tensor-core mode products as `mul!`.
struct DeviceInverse{T} <: AbstractMatrix{T}      # what `F` holds: inv(A), applied by multiplication
    inv::CuMatrix{T}
    blob::Union{Nothing, CuVector{UInt8}}         # Float64: the LU solve structures of scimlop_factorize
    cond::Float64                                 # cond_1(A) from the factorisation (NaN for Float32)
end
Base.size(F::DeviceInverse) = size(F.inv)
SciMLOperators.has_ldiv(::DeviceInverse) = true
SciMLOperators.has_ldiv!(::DeviceInverse) = true
const COND_FAST = 1.0e3       # up to here ldiv! multiplies by the inverse (residual cond * eps <= 1e-13)
stable_solve(F::DeviceInverse) = F.blob !== nothing && !(F.cond <= COND_FAST)

function kron_ldiv!(w::DevMat{Float64}, blobs, dims::Vector{Int64}, kinds::Vector{Cint}, v::DevMat{Float64})
    ws = CUDA.zeros(UInt8, sizeof(v))             # a real extension keeps this in the operator's cache
    ptrs = Ptr{Cvoid}[b === nothing ? C_NULL : dptr(b) for b in blobs]
    check(ccall((:scimlop_kron_ldiv, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Ptr{Ptr{Cvoid}}, Ptr{Int64}, Ptr{Cint}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid},
                 Csize_t, Ptr{Cvoid}),
                handle(), 0, length(dims), ptrs, dims, kinds, dptr(v), dptr(w), size(v, 2), dptr(ws), sizeof(ws),
                stream_ptr()), "scimlop_kron_ldiv")
    w
end
function LinearAlgebra.ldiv!(w::DevMat{T}, F::DeviceInverse{T}, v::DevMat{T}) where {T <: F}
    stable_solve(F) || return mul!(w, F.inv, v, true, false)
    kron_ldiv!(w, (F.blob,), Int64[size(F, 1)], Cint[DENSE], v)
end
# adjoint(::InvertibleOperator) wraps F lazily (src/matrix.jl:565): the same blob, transposed sweeps (getrs('T'))
function LinearAlgebra.ldiv!(w::DevMat{T}, Ft::Union{Adjoint{T, <:DeviceInverse{T}}, Transpose{T, <:DeviceInverse{T}}},
                             v::DevMat{T}) where {T <: F}
    F = parent(Ft)
    stable_solve(F) || return mul!(w, transpose(F.inv), v, true, false)
    kron_ldiv!(w, (F.blob,), Int64[size(F, 1)], Cint[DENSE | TRANS], v)
end

function LinearAlgebra.factorize(L::MatrixOperator{Float64, <:CuMatrix{Float64}})
    n = LinearAlgebra.checksquare(L.A)
    fb = Ref{Csize_t}(0); sb = Ref{Csize_t}(0)
    ccall((:scimlop_factorize_bytes, LIB), Cint, (Cint, Int64, Ref{Csize_t}, Ref{Csize_t}), 0, n, fb, sb)
    blob, scratch, Ainv, cond = CUDA.zeros(UInt8, fb[]), CUDA.zeros(UInt8, sb[]), similar(L.A), Ref(0.0)
    check(ccall((:scimlop_factorize, LIB), Cint,
                (Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Csize_t,
                 Ref{Float64}, Ptr{Cvoid}),
                handle(), 0, n, dptr(L.A), stride(L.A, 2), dptr(blob), fb[], dptr(Ainv), stride(Ainv, 2), dptr(scratch),
                sb[], cond, stream_ptr()), "scimlop_factorize")   # status 9 (singular) -> error, like SingularException
    SciMLOperators.InvertibleOperator(L, DeviceInverse{Float64}(Ainv, blob, cond[]))
end
function LinearAlgebra.factorize(L::MatrixOperator{Float32, <:CuMatrix{Float32}})
    n = LinearAlgebra.checksquare(L.A)
    Ainv = similar(L.A)
    nbytes = Ref{Csize_t}(0)
    ccall((:scimlop_invert_workspace_bytes, LIB), Cint, (Cint, Int64, Ref{Csize_t}), 1, n, nbytes)
    ws = CUDA.zeros(UInt8, nbytes[])
    check(ccall((:scimlop_invert, LIB), Cint,
                (Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
                handle(), 1, n, dptr(L.A), stride(L.A, 2), dptr(Ainv), stride(Ainv, 2), dptr(ws), sizeof(ws),
                stream_ptr()), "scimlop_invert")
    SciMLOperators.InvertibleOperator(L, DeviceInverse{Float32}(Ainv, nothing, NaN))
end

# like `flatten`, with every factorised leaf replaced by its cached device inverse and every λ by 1/λ
function flatten_inverse(L::TensorProductOperator, T)
    kinds = Cint[]; ptrs = Ptr{Cvoid}[]; dims = Int64[]; lds = Int64[]; λinv = one(T); blobs = Any[]
    function walk(op)
        if op isa ScaledOperator
            λinv /= convert(Number, op.λ); return walk(op.L)
        elseif op isa TensorProductOperator
            return all(walk, op.ops)
        elseif op isa IdentityOperator
            push!(kinds, IDENTITY); push!(ptrs, C_NULL); push!(lds, 0); push!(blobs, nothing)
        elseif op isa SciMLOperators.InvertibleOperator && op.F isa DeviceInverse{T}
            push!(kinds, DENSE); push!(ptrs, dptr(op.F.inv)); push!(lds, stride(op.F.inv, 2)); push!(blobs, op.F)
        else
            return false                      # not factorised on the device: the generic ldiv! runs
        end
        append!(dims, size(op)); true
    end
    walk(L) ? (kinds, ptrs, dims, lds, λinv, blobs) : nothing
end
function LinearAlgebra.ldiv!(w::DevMat{T}, L::TensorProductOperator, v::DevMat{T}) where {T <: F}
    @assert iscached(L)                                                            # src/tensor.jl:617
    fl = flatten_inverse(L, T)
    fl === nothing && return invoke(ldiv!, Tuple{AbstractVecOrMat, TensorProductOperator, AbstractVecOrMat}, w, L, v)
    kinds, ptrs, dims, lds, λinv, blobs = fl
    if T === Float64 && any(b -> b !== nothing && stable_solve(b), blobs) && all(b -> b === nothing || b.blob !== nothing, blobs)
        # an ill-conditioned factor: substitution sweeps along every mode (scimlop_kron_ldiv), then 1/λ
        src = dptr(w) == dptr(v) ? copy(v) : v
        kron_ldiv!(w, Any[b === nothing ? nothing : b.blob for b in blobs], dims[1:2:end], kinds, src)
        isone(λinv) || CUDA.@sync(w .*= λinv)
        return w
    end
    ws = L.cache[1]
    check(ccall((:scimlop_kron_mul, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Ptr{Cvoid}}, Ptr{Int64}, Ptr{Int64}, Ptr{Cint}, Ptr{Cvoid}, Int64,
                 Ptr{Cvoid}, Int64, Int64, Ptr{T}, Ptr{T}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
                handle(), dtype(T), 0, length(kinds), ptrs, dims, lds, kinds, dptr(v), size(v, 1), dptr(w),
                size(w, 1), size(v, 2), Ref(T(λinv)), C_NULL, dptr(ws), sizeof(ws), stream_ptr()), "scimlop_kron_mul")
    w
end
LinearAlgebra.ldiv!(L::TensorProductOperator, v::DevMat{T}) where {T <: F} = ldiv!(v, L, v)   # W may alias V when >= 2 stages

function LinearAlgebra.ldiv!(w::DevMat{T}, L::BatchedDiagonalOperator, v::DevMat{T}) where {T <: F}   # src/batch.jl:181-193
    check(ccall((:scimlop_diag_div, LIB), Cint,
                (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}),
                handle(), dtype(T), size(v, 1), size(v, 2), dptr(L.diag), stride(L.diag, 2), dptr(v), stride(v, 2),
                dptr(w), stride(w, 2), stream_ptr()), "scimlop_diag_div")
    w
end

# ---- solver-loop kernels (device-resident per-column scalars), for Krylov callers -----------------------------
function col_dot!(out::CuVector{T}, x::CuMatrix{T}, y::CuMatrix{T}) where {T <: F}
    check(ccall((:scimlop_col_dot, LIB), Cint,
                (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}),
                handle(), dtype(T), size(x, 1), size(x, 2), dptr(x), stride(x, 2), dptr(y), stride(y, 2), dptr(out),
                stream_ptr()), "scimlop_col_dot")
    out
end

# ---- sparse MatrixOperator (test/basic.jl:405-426, test/Downstream/alloccheck.jl:37-45) ---------------------------------
# CUDA.jl's CuSparseMatrixCSR stores 1-based Int32 rowPtr / colVal; a CuSparseMatrixCSC is the CSR form of the transpose.
using CUDA.CUSPARSE: CuSparseMatrixCSR, CuSparseMatrixCSC
function csr_mul!(w::DevMat{T}, rowptr, colval, nzval, rows, cols, trans, v::DevMat{T}, α, β) where {T <: F}
    check(ccall((:scimlop_csr_mul, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Int64, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Int64,
                 Ptr{Cvoid}, Int64, Ref{T}, Ref{T}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
                handle(), dtype(T), trans, rows, cols, size(v, 2), dptr(rowptr), dptr(colval), dptr(nzval), 1, dptr(v),
                stride(v, 2), dptr(w), stride(w, 2), T(α), T(β), C_NULL, 0, stream_ptr()), "scimlop_csr_mul")
    # (C_NULL workspace: the direct gather; `cache_operator` may keep a scimlop_csr_workspace_bytes buffer for the
    #  row-contiguous staging copy of v and pass it here)
    w
end
mul!(w::DevMat{T}, L::MatrixOperator{T, <:CuSparseMatrixCSR{T, Int32}}, v::DevMat{T}, α::Number, β::Number) where {T <: F} =
    csr_mul!(w, L.A.rowPtr, L.A.colVal, L.A.nzVal, size(L.A, 1), size(L.A, 2), 0, v, α, β)
mul!(w::DevMat{T}, L::MatrixOperator{T, <:CuSparseMatrixCSC{T, Int32}}, v::DevMat{T}, α::Number, β::Number) where {T <: F} =
    csr_mul!(w, L.A.colPtr, L.A.rowVal, L.A.nzVal, size(L.A, 2), size(L.A, 1), 1, v, α, β)
# trans = 1 is the scatter kernel (atomic adds: the last bits vary from run to run).  A CACHED operator should convert
# once instead -- `cache_operator(L::MatrixOperator{T, <:CuSparseMatrixCSC}, v)` keeping `CuSparseMatrixCSR(L.A)` (CUSPARSE
# csr2csc) next to the staging workspace -- and call the gather (trans = 0) on that: deterministic and 2-4x faster; the
# Python mirror does exactly this for lazy adjoints (operators.py, `_csr_transpose`).

# A sparse matrix as a Kronecker FACTOR: kind SCIMLOP_CSR = 6, factors[f] = pointer to a host struct of device pointers
# (include/scimlop_b200.h, scimlop_csr_factor).  `flatten` pushes this for MatrixOperator{T, <:CuSparseMatrixCSR}; keep the
# Ref alive for the duration of the ccall (GC.@preserve).
struct CsrFactor
    rowptr::Ptr{Cvoid}; colidx::Ptr{Cvoid}; values::Ptr{Cvoid}; index_base::Int32; reserved::Int32
end
const CSR = Cint(6)
csr_factor(A::CuSparseMatrixCSR) = Ref(CsrFactor(dptr(A.rowPtr), dptr(A.colVal), dptr(A.nzVal), Int32(1), Int32(0)))

# ---- ComplexF32 / ComplexF64 (test/basic.jl:405-426, alloccheck.jl:37): dtype codes 2 / 3, interleaved storage as is ----
# A complex MatrixOperator mul! is the one-factor case of scimlop_kron_mul's complex planner; the workspace
# (scimlop_kron_workspace_bytes with the complex dtype) belongs in the operator's cache.
const CF = Union{ComplexF32, ComplexF64}
cdtype(::Type{ComplexF32}) = Cint(2)
cdtype(::Type{ComplexF64}) = Cint(3)
function mul!(w::DevMat{T}, L::MatrixOperator{T, <:DenseDev{T}}, v::DevMat{T}, α::Number, β::Number) where {T <: CF}
    A, trans = _stored(L.A)
    dims = Int64[size(L, 1), size(L, 2)]; lds = Int64[stride(A, 2)]; kinds = Cint[DENSE | (trans == 1 ? TRANS : 0)]
    ptrs = Ptr{Cvoid}[dptr(A)]
    nbytes = Ref{Csize_t}(0)
    ccall((:scimlop_kron_workspace_bytes, LIB), Cint, (Cint, Cint, Ptr{Int64}, Ptr{Cint}, Int64, Ref{Csize_t}),
          cdtype(T), 1, dims, kinds, size(v, 2), nbytes)
    ws = CUDA.zeros(UInt8, nbytes[])
    check(ccall((:scimlop_kron_mul, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Ptr{Cvoid}}, Ptr{Int64}, Ptr{Int64}, Ptr{Cint}, Ptr{Cvoid}, Int64,
                 Ptr{Cvoid}, Int64, Int64, Ref{T}, Ref{T}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
                handle(), cdtype(T), 0, 1, ptrs, dims, lds, kinds, dptr(v), size(v, 1), dptr(w), size(w, 1), size(v, 2),
                T(α), T(β), dptr(ws), sizeof(ws), stream_ptr()), "scimlop_kron_mul")
    w
end

end # module
