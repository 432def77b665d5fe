# NextLAB200.jl -- drop-in Julia side of the B200 library for NextLA.jl's recursive TRSM/TRMM path.
#
# Adds a more specific method of `NextLA.unified_rectrxm!` for CUDA.jl device matrices with the SAME positional
# signature as src/rectrxm.jl:43-51, so existing call sites (e.g. test/unified_rectrxm.jl:33) dispatch to it unchanged.
# Everything below the call is the C ABI of include/nextla_b200.h (libnextla_b200.so), reached with `ccall`.
# No KernelAbstractions, no CPU fallback: a non-zero status raises.
#
# NOTE: Julia is not installed in the build/GPU images of this project, so this file is shipped UNEXECUTED; it is a
# mechanical mirror of the ctypes binding in nextla.jl_b200/__init__.py, which is what the tests drive.
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

# one handle per (task-local) device
const HANDLES = Dict{Int,Ptr{Cvoid}}()
const HANDLES_LOCK = ReentrantLock()
function handle()
    dev = CUDA.deviceid(CUDA.device())
    lock(HANDLES_LOCK) do
        get!(HANDLES, dev) do
            h = Ref{Ptr{Cvoid}}(C_NULL)
            check(ccall((:nla_create, libnextla), Cint, (Ref{Ptr{Cvoid}}, Cint), h, dev))
            h[]
        end
    end
end

"""
    unified_rectrxm!(side, uplo, transpose, alpha, func, A::StridedCuMatrix{T}, B::StridedCuMatrix{T})

Same semantics and return value as the reference method (src/rectrxm.jl:43-76): in place on `B`, asynchronous on the
task-local CUDA stream (the reference does not synchronise either, :75).
"""
function NextLA.unified_rectrxm!(side::Char, uplo::Char, transpose::Char, alpha::Number, func::Char,
                                 A::StridedCuMatrix{T}, B::StridedCuMatrix{T}) where {T<:B200Float}
    n = size(A, 1)
    m = side == 'L' ? size(B, 2) : size(B, 1)
    stride(A, 1) == 1 && stride(B, 1) == 1 || throw(ArgumentError("A and B must be column-major with unit row stride"))
    GC.@preserve A B begin
        rc = ccall((:nla_rectrxm, libnextla), Cint,
                   (Ptr{Cvoid}, Cchar, Cchar, Cchar, Cchar, Cint, Int64, Int64, Cdouble, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, Int64, Ptr{Cvoid}),
                   handle(), side, uplo, transpose, func, dtype_code(T), n, m, Float64(alpha),
                   pointer(A), max(1, stride(A, 2)), pointer(B), max(1, stride(B, 2)), CUDA.stream().handle)
        check(rc)
    end
    return B
end

# Leaf launchers (src/trsm.jl:128-150, src/trmm.jl:332-389) and GEMM updates (src/matmul.jl:69-81) on device matrices.
for (fname, cfun, side, uplo) in ((:LeftLowerTRSM!, :nla_trsm_leaf, 'L', 'L'), (:LeftUpperTRSM!, :nla_trsm_leaf, 'L', 'U'),
                                  (:RightLowerTRSM!, :nla_trsm_leaf, 'R', 'L'), (:RightUpperTRSM!, :nla_trsm_leaf, 'R', 'U'),
                                  (:LeftLowerTRMM!, :nla_trmm_leaf, 'L', 'L'), (:LeftUpperTRMM!, :nla_trmm_leaf, 'L', 'U'),
                                  (:RightLowerTRMM!, :nla_trmm_leaf, 'R', 'L'), (:RightUpperTRMM!, :nla_trmm_leaf, 'R', 'U'))
    @eval function NextLA.$fname(A::StridedCuMatrix{T}, B::StridedCuMatrix{T}; kwargs...) where {T<:B200Float}
        n = size(A, 1)
        m = $side == 'L' ? size(B, 2) : size(B, 1)
        GC.@preserve A B check(ccall(($(QuoteNode(cfun)), libnextla), Cint,
            (Ptr{Cvoid}, Cchar, Cchar, Cint, Int64, Int64, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, Int64, Ptr{Cvoid}),
            handle(), $side, $uplo, dtype_code(T), n, m, pointer(A), max(1, stride(A, 2)), pointer(B), max(1, stride(B, 2)), CUDA.stream().handle))
        return B
    end
end

function gemm_update!(C::StridedCuMatrix{T}, A::StridedCuMatrix{T}, B::StridedCuMatrix{T}, sign::Integer) where {T<:B200Float}
    M, N, K = size(C, 1), size(C, 2), size(A, 2)
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
        n = size(A, 1)
        m = side == 'L' ? size(B, 2) : size(B, 1)
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
    n = size(A, 1)
    m = side == 'L' ? size(B, 2) : size(B, 1)
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

end # module
