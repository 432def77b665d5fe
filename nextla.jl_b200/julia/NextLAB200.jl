# NextLAB200.jl -- drop-in Julia side of the B200 library for NextLA.jl's recursive TRSM/TRMM path.
#
# Adds a more specific method of `NextLA.unified_rectrxm!` for CUDA.jl device matrices with the SAME positional
# signature as src/rectrxm.jl:43-51, so existing call sites (e.g. test/unified_rectrxm.jl:33) dispatch to it unchanged.
# Everything below the call is the C ABI of include/nextla_b200.h (libnextla_b200.so), reached with `ccall`.
# No KernelAbstractions, no CPU fallback: a non-zero status raises.
#
# NOTE: Julia is not installed in the build/GPU images of this project, so this file is shipped UNEXECUTED; it is a
# mechanical mirror of the ctypes binding in nextla.jl_b200/__init__.py, which is what the tests drive.  tests/test_abi.py checks, on
# CPU, that every `ccall` in this file names a symbol the library exports and passes as many arguments as the C prototype declares.
module NextLAB200

using CUDA
import NextLA

const libnextla = get(ENV, "NEXTLA_B200_LIB", joinpath(@__DIR__, "..", "libnextla_b200.so"))

const NLA_F64, NLA_F32, NLA_F16 = Cint(0), Cint(1), Cint(2)
dtype_code(::Type{Float64}) = NLA_F64
dtype_code(::Type{Float32}) = NLA_F32
dtype_code(::Type{Float16}) = NLA_F16

const B200Float = Union{Float64,Float32,Float16}

status_string(rc) = unsafe_string(ccall((:nla_status_string, libnextla), Cstring, (Cint,), rc))

function check(rc::Cint)
    rc == 0 && return nothing
    rc in (1, 2, 3, 4) && throw(ArgumentError("nextla_b200: " * status_string(rc)))
    error("nextla_b200: " * status_string(rc))
end

# One handle per (device, stream).  A handle owns device workspaces (block inverses, copies of B), so calls through one handle must be
# ordered on one stream (include/nextla_b200.h); CUDA.jl gives every Julia task its own stream, hence the cache is keyed by the stream
# handle as well: two tasks on one GPU never share workspaces.
const HANDLES = Dict{Tuple{Int,UInt},Ptr{Cvoid}}()
const HANDLES_LOCK = ReentrantLock()
function handle()
    dev = CUDA.deviceid(CUDA.device())
    key = (dev, UInt(Base.unsafe_convert(Ptr{Cvoid}, CUDA.stream().handle)))
    lock(HANDLES_LOCK) do
        get!(HANDLES, key) do
            h = Ref{Ptr{Cvoid}}(C_NULL)
            check(ccall((:nla_create, libnextla), Cint, (Ref{Ptr{Cvoid}}, Cint), h, dev))
            h[]
        end
    end
end

# Shape checks the C side cannot make (it only sees n, m and the leading dimensions): what the reference's generic kernels would
# turn into a BoundsError becomes a DimensionMismatch here instead of an out-of-bounds device access.
function check_shapes(side::Char, A, B)
    n = size(A, 1)
    size(A, 2) == n || throw(DimensionMismatch("A must be square, got $(size(A))"))
    (side == 'L' || side == 'R') || throw(ArgumentError("side must be 'L' or 'R', got '$side'"))
    nb = side == 'L' ? size(B, 1) : size(B, 2)
    nb == n || throw(DimensionMismatch("B has $(nb) $(side == 'L' ? "rows" : "columns"), A is of order $n"))
    (stride(A, 1) == 1 && stride(B, 1) == 1) || throw(ArgumentError("A and B must be column-major with unit row stride"))
    return n, (side == 'L' ? size(B, 2) : size(B, 1))
end

"""
    unified_rectrxm!(side, uplo, transpose, alpha, func, A::StridedCuMatrix{T}, B::StridedCuMatrix{T})

Same semantics and return value as the reference method (src/rectrxm.jl:43-76): in place on `B`, asynchronous on the
task-local CUDA stream (the reference does not synchronise either, :75).
"""
function NextLA.unified_rectrxm!(side::Char, uplo::Char, transpose::Char, alpha::Number, func::Char,
                                 A::StridedCuMatrix{T}, B::StridedCuMatrix{T}) where {T<:B200Float}
    n, m = check_shapes(side, A, B)
    GC.@preserve A B begin
        rc = ccall((:nla_rectrxm, libnextla), Cint,
                   (Ptr{Cvoid}, Cchar, Cchar, Cchar, Cchar, Cint, Int64, Int64, Cdouble, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, Int64, Ptr{Cvoid}),
                   handle(), side, uplo, transpose, func, dtype_code(T), n, m, Float64(alpha),
                   pointer(A), max(1, stride(A, 2)), pointer(B), max(1, stride(B, 2)), CUDA.stream().handle)
        check(rc)
    end
    return B
end

# Complex element types (SURVEY.md 8(f4)): the reference builds Adjoint(A) for transpose = 'C' (src/rectrxm.jl:57) but its recursion only
# accepts real element types (:101); this method makes the advertised complex support (README.md:20) work, with 'C' distinct from 'T'.
const B200Complex = Union{ComplexF32,ComplexF64}
dtype_code(::Type{ComplexF32}) = Cint(3)
dtype_code(::Type{ComplexF64}) = Cint(4)
function NextLA.unified_rectrxm!(side::Char, uplo::Char, transpose::Char, alpha::Number, func::Char,
                                 A::StridedCuMatrix{T}, B::StridedCuMatrix{T}) where {T<:B200Complex}
    n, m = check_shapes(side, A, B)
    a = ComplexF64(alpha)
    GC.@preserve A B check(ccall((:nla_rectrxm_complex, libnextla), Cint,
        (Ptr{Cvoid}, Cchar, Cchar, Cchar, Cchar, Cchar, Cint, Int64, Int64, Cdouble, Cdouble, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, Int64, Ptr{Cvoid}),
        handle(), side, uplo, transpose, 'N', func, dtype_code(T), n, m, real(a), imag(a),
        pointer(A), max(1, stride(A, 2)), pointer(B), max(1, stride(B, 2)), CUDA.stream().handle))
    return B
end

# Leaf launchers (src/trsm.jl:128-150, src/trmm.jl:332-389) and GEMM updates (src/matmul.jl:69-81) on device matrices.
for (fname, cfun, side, uplo) in ((:LeftLowerTRSM!, :nla_trsm_leaf, 'L', 'L'), (:LeftUpperTRSM!, :nla_trsm_leaf, 'L', 'U'),
                                  (:RightLowerTRSM!, :nla_trsm_leaf, 'R', 'L'), (:RightUpperTRSM!, :nla_trsm_leaf, 'R', 'U'),
                                  (:LeftLowerTRMM!, :nla_trmm_leaf, 'L', 'L'), (:LeftUpperTRMM!, :nla_trmm_leaf, 'L', 'U'),
                                  (:RightLowerTRMM!, :nla_trmm_leaf, 'R', 'L'), (:RightUpperTRMM!, :nla_trmm_leaf, 'R', 'U'))
    @eval function NextLA.$fname(A::StridedCuMatrix{T}, B::StridedCuMatrix{T}; kwargs...) where {T<:B200Float}
        n, m = check_shapes($side, A, B)
        GC.@preserve A B check(ccall(($(QuoteNode(cfun)), libnextla), Cint,
            (Ptr{Cvoid}, Cchar, Cchar, Cint, Int64, Int64, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, Int64, Ptr{Cvoid}),
            handle(), $side, $uplo, dtype_code(T), n, m, pointer(A), max(1, stride(A, 2)), pointer(B), max(1, stride(B, 2)), CUDA.stream().handle))
        return B
    end
end

function gemm_update!(C::StridedCuMatrix{T}, A::StridedCuMatrix{T}, B::StridedCuMatrix{T}, sign::Integer) where {T<:B200Float}
    M, N, K = size(C, 1), size(C, 2), size(A, 2)
    (size(A, 1) == M && size(B, 1) == K && size(B, 2) == N) || throw(DimensionMismatch("GEMM update: C is $(size(C)), A $(size(A)), B $(size(B))"))
    GC.@preserve A B C check(ccall((:nla_gemm_update, libnextla), Cint,
        (Ptr{Cvoid}, Cint, Cchar, Cchar, Int64, Int64, Int64, Cint, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, Int64, Ptr{Cvoid}),
        handle(), dtype_code(T), 'N', 'N', M, N, K, Cint(sign), pointer(A), max(1, stride(A, 2)), pointer(B), max(1, stride(B, 2)),
        pointer(C), max(1, stride(C, 2)), CUDA.stream().handle))
    return C
end
# GEMM_ADD!(A,B,C): C += A*B (src/matmul.jl:69-74);  GEMM_SUB!(A,B,C): A -= B*C (src/matmul.jl:76-81)
NextLA.GEMM_ADD!(A::StridedCuMatrix{T}, B::StridedCuMatrix{T}, C::StridedCuMatrix{T}; kwargs...) where {T<:B200Float} = gemm_update!(C, A, B, 1)
NextLA.GEMM_SUB!(A::StridedCuMatrix{T}, B::StridedCuMatrix{T}, C::StridedCuMatrix{T}) where {T<:B200Float} = gemm_update!(A, B, C, -1)

# trsm / trmm (src/trsm.jl:186-205, src/trmm.jl:430-448; un-exported in the reference): same argument order.  The reference ignores
# `transa` and `diag` and runs one leaf; here both are honoured and the call recurses (nla_trxm).
for (fname, func) in ((:trsm, 'S'), (:trmm, 'M'))
    @eval function NextLA.$fname(side::Char, uplo::Char, transa::Char, diag::Char, A::StridedCuMatrix{T}, B::StridedCuMatrix{T},
                                 alpha::Number = one(T)) where {T<:B200Float}
        n, m = check_shapes(side, A, B)
        GC.@preserve A B check(ccall((:nla_trxm, libnextla), Cint,
            (Ptr{Cvoid}, Cchar, Cchar, Cchar, Cchar, Cchar, Cint, Int64, Int64, Cdouble, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, Int64, Ptr{Cvoid}),
            handle(), side, uplo, transa, diag, $func, dtype_code(T), n, m, Float64(alpha),
            pointer(A), max(1, stride(A, 2)), pointer(B), max(1, stride(B, 2)), CUDA.stream().handle))
        return B
    end
end

# laswp(A, first, last, ipiv, incx) (src/lu.jl:470-530) on a device matrix; `ipiv` is a CuVector{Int64} (1-based, like the reference).
# With trsm('L','L','N','U', ...) and GEMM_SUB! above this is every O(n^3) step of one level of getrf2! (src/lu.jl:274-280) on the device.
function NextLA.laswp(A::StridedCuMatrix{T}, first::Integer, last::Integer, ipiv::CuVector{Int64}, incx::Integer) where {T<:B200Float}
    (1 <= first && last <= min(size(A, 1), length(ipiv))) || last < first || throw(DimensionMismatch("laswp: rows $first:$last out of range"))
    GC.@preserve A ipiv check(ccall((:nla_laswp, libnextla), Cint,
        (Ptr{Cvoid}, Cint, Int64, Int64, CuPtr{Cvoid}, Int64, Int64, Int64, CuPtr{Int64}, Cint, Ptr{Cvoid}),
        handle(), dtype_code(T), size(A, 1), size(A, 2), pointer(A), max(1, stride(A, 2)), Int64(first), Int64(last), pointer(ipiv), Cint(incx),
        CUDA.stream().handle))
    return A
end

"""
    unified_rectrxm_gated!(side, uplo, transpose, alpha, func, A, B, panel_cols, events)

Multi-GPU variant (no reference counterpart): `A` is still arriving in column panels of `panel_cols` columns (NCCL broadcast from its
owner on a side stream, in the order `panel_order` returns); `events[p+1]` is a `CuEvent` recorded after panel `p`.  The schedule waits
for a panel right before the first launch that reads it.
"""
function unified_rectrxm_gated!(side::Char, uplo::Char, transpose::Char, alpha::Number, func::Char, A::StridedCuMatrix{T},
                                B::StridedCuMatrix{T}, panel_cols::Integer, events::Vector{CuEvent}) where {T<:B200Float}
    n, m = check_shapes(side, A, B)
    length(events) * panel_cols >= n || throw(DimensionMismatch("$(length(events)) panels of $panel_cols columns do not cover n = $n"))
    hs = Ptr{Cvoid}[Base.unsafe_convert(Ptr{Cvoid}, e.handle) for e in events]
    GC.@preserve A B events hs check(ccall((:nla_rectrxm_gated, libnextla), Cint,
        (Ptr{Cvoid}, Cchar, Cchar, Cchar, Cchar, Cint, Int64, Int64, Cdouble, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Int64, Ptr{Ptr{Cvoid}}),
        handle(), side, uplo, transpose, func, dtype_code(T), n, m, Float64(alpha), pointer(A), max(1, stride(A, 2)),
        pointer(B), max(1, stride(B, 2)), CUDA.stream().handle, Int64(panel_cols), Int64(length(events)), hs))
    return B
end

function panel_order(side::Char, uplo::Char, transpose::Char, func::Char, n::Integer, panel_cols::Integer)
    np = cld(n, panel_cols)
    order = Vector{Int64}(undef, np)
    cnt = ccall((:nla_panel_order, libnextla), Int64, (Cchar, Cchar, Cchar, Cchar, Int64, Int64, Ptr{Int64}, Int64),
                side, uplo, transpose, func, n, panel_cols, order, np)
    cnt < 0 && check(Cint(-cnt))
    return order   # 0-based panel indices in consumption order
end

"""
    getrf2!(A::StridedCuMatrix{T}, ipiv::CuVector{Int64}, info)

The reference's recursive LU (src/lu.jl:185) on a device matrix: A = P*L*U in place, `ipiv` the 1-based row interchanges (device vector
of at least min(m, n) entries).  Returns `(A, ipiv, info)` like the reference, `info` read back from the device after the call (the one
synchronisation: the reference's return value is a host integer): 0, or the first i with U[i, i] == 0.
"""
function NextLA.getrf2!(A::StridedCuMatrix{T}, ipiv::CuVector{Int64}, info::Integer=0) where {T<:Union{Float32,Float64}}
    m, n = size(A)
    length(ipiv) >= min(m, n) || throw(DimensionMismatch("ipiv has $(length(ipiv)) entries, min(m, n) = $(min(m, n))"))
    dinfo = CUDA.zeros(Cint, 1)
    GC.@preserve A ipiv dinfo check(ccall((:nla_getrf2, libnextla), Cint,
        (Ptr{Cvoid}, Cint, Int64, Int64, CuPtr{Cvoid}, Int64, CuPtr{Int64}, CuPtr{Cint}, Ptr{Cvoid}),
        handle(), dtype_code(T), m, n, pointer(A), max(1, stride(A, 2)), pointer(ipiv), pointer(dinfo), CUDA.stream().handle))
    return A, ipiv, Int(Array(dinfo)[1])
end

"""
    lauum!(uplo, n, A::StridedCuMatrix{T}, ib)

Same signature as the reference (src/lauum.jl:52): A := U*U^H (uplo 'U') or L^H*L (uplo 'L') in the `uplo` triangle of the device matrix.
Like the reference it throws ArgumentError for a bad `uplo` or a negative `n`.
"""
function NextLA.lauum!(uplo::Char, n::Integer, A::StridedCuMatrix{T}, ib::Integer) where {T<:B200Float}
    uplo in ('U', 'L') || throw(ArgumentError("uplo must be 'U' or 'L', got '$uplo'"))
    n >= 0 || throw(ArgumentError("n must be non-negative, got $n"))
    (size(A, 1) >= n && size(A, 2) >= n) || throw(DimensionMismatch("A is $(size(A)), n = $n"))
    n == 0 && return
    GC.@preserve A check(ccall((:nla_lauum, libnextla), Cint, (Ptr{Cvoid}, Cchar, Cint, Int64, CuPtr{Cvoid}, Int64, Int64, Ptr{Cvoid}),
                               handle(), uplo, dtype_code(T), n, pointer(A), max(1, stride(A, 2)), Int64(ib), CUDA.stream().handle))
    return
end

# Workspace control (include/nextla_b200.h): by default the library grows its workspaces with the stream-ordered allocator, so calls stay
# asynchronous; `reserve!` pre-sizes them, `set_workspace!` hands the library a CuVector{UInt8} arena (kept alive by the caller).
workspace_bytes(side::Char, func::Char, ::Type{T}, n::Integer, m::Integer) where {T<:B200Float} =
    ccall((:nla_workspace_bytes, libnextla), Int64, (Ptr{Cvoid}, Cchar, Cchar, Cint, Int64, Int64), handle(), side, func, dtype_code(T), n, m)
reserve!(side::Char, func::Char, ::Type{T}, n::Integer, m::Integer) where {T<:B200Float} =
    check(ccall((:nla_reserve, libnextla), Cint, (Ptr{Cvoid}, Cchar, Cchar, Cint, Int64, Int64), handle(), side, func, dtype_code(T), n, m))
set_workspace!(buf::Union{Nothing,CuVector{UInt8}}) =
    check(ccall((:nla_set_workspace, libnextla), Cint, (Ptr{Cvoid}, CuPtr{Cvoid}, Int64), handle(),
                buf === nothing ? CU_NULL : pointer(buf), buf === nothing ? 0 : length(buf)))

# Single-process multi-GPU (nla_mg_*): the library owns the NCCL communicators, streams and replicas of A; B is sharded by the caller
# (shard i on device i: columns of B for side 'L', rows for side 'R').  Asynchronous; `mg_sync` waits.
mutable struct MultiGPU
    ptr::Ptr{Cvoid}
    function MultiGPU(devices::Vector{<:Integer} = collect(0:length(CUDA.devices())-1))
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:nla_mg_create, libnextla), Cint, (Ref{Ptr{Cvoid}}, Cint, Ptr{Cint}), h, length(devices), Cint.(devices)))
        mg = new(h[])
        finalizer(m -> ccall((:nla_mg_destroy, libnextla), Cint, (Ptr{Cvoid},), m.ptr), mg)
        return mg
    end
end
function mg_rectrxm!(mg::MultiGPU, side::Char, uplo::Char, transpose::Char, alpha::Number, func::Char, root::Integer,
                     A::StridedCuMatrix{T}, shards::Vector{<:StridedCuMatrix{T}}) where {T<:B200Float}
    n = size(A, 1)
    size(A, 2) == n || throw(DimensionMismatch("A must be square"))
    for s in shards
        (side == 'L' ? size(s, 1) : size(s, 2)) == n || throw(DimensionMismatch("shard of size $(size(s)) does not match n = $n"))
    end
    ptrs = CuPtr{Cvoid}[pointer(s) for s in shards]
    ms = Int64[side == 'L' ? size(s, 2) : size(s, 1) for s in shards]
    lds = Int64[max(1, stride(s, 2)) for s in shards]
    GC.@preserve A shards check(ccall((:nla_mg_rectrxm, libnextla), Cint,
        (Ptr{Cvoid}, Cchar, Cchar, Cchar, Cchar, Cint, Int64, Cdouble, Cint, CuPtr{Cvoid}, Int64, Ptr{CuPtr{Cvoid}}, Ptr{Int64}, Ptr{Int64}),
        mg.ptr, side, uplo, transpose, func, dtype_code(T), n, Float64(alpha), root, pointer(A), max(1, stride(A, 2)), ptrs, ms, lds))
    return shards
end
mg_sync(mg::MultiGPU) = check(ccall((:nla_mg_sync, libnextla), Cint, (Ptr{Cvoid},), mg.ptr))

end # module
