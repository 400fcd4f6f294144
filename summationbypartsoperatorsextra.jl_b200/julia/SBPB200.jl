# SBPB200.jl -- Julia shim that keeps the reference API (`mul!(du, D, u)`, `D * u`, `D[dim] * u`, `integrate`,
# `scale_by_mass_matrix!`, `semi(du, u, p, t)`, `analyze_quantities`, `AnalysisCallback`) and routes device-resident batches
# to libsbp_b200.so through `ccall` (C ABI: include/sbp_b200.h).
#
# NOTE: this image has no Julia, so this file has not been executed here; it is the reference-side binding a maintainer adds
# (see INTEGRATION.md).  tests/test_host.py::test_julia_shim_ccall_signatures parses every `ccall` below and checks name,
# return type and argument types against the ctypes table that the GPU tests run through (`_lib.SIGNATURES`), so the
# signatures cannot drift from include/sbp_b200.h unnoticed.
#
# Usage
#   using SummationByPartsOperatorsExtra, SBPB200
#   ctx = SBPB200.Context(0)
#   D   = polynomialbases_derivative_operator(LobattoLegendre; xmin = -1.0, xmax = 1.0, N = 64)
#   U   = SBPB200.DeviceBatch(ctx, rand(64, 2^20))          # N x B, column-major: each instance contiguous
#   dU  = similar(U)
#   mul!(dU, D, U)                                          # one kernel launch for the whole batch
#   E   = integrate(abs2, U, D)                             # per-instance discrete energies (Vector{Float64})
module SBPB200

using LinearAlgebra
import LinearAlgebra: mul!
using SummationByPartsOperatorsExtra
using SummationByPartsOperatorsExtra: PolynomialBasesDerivativeOperator, SubcellOperator,
                                      MultidimensionalLinearAdvectionNonperiodicSemidiscretization,
                                      mass_matrix, mass_matrix_boundary, grid
import SummationByPartsOperatorsExtra: integrate, analyze_quantities
using SummationByPartsOperators: SummationByPartsOperators, MatrixDerivativeOperator,
                                 MultidimensionalMatrixDerivativeOperator,
                                 VariableLinearAdvectionNonperiodicSemidiscretization
import SummationByPartsOperators: scale_by_mass_matrix!, scale_by_inverse_mass_matrix!

const libsbp = get(ENV, "SBP_B200_LIB", joinpath(@__DIR__, "..", "libsbp_b200.so"))

const SBP_EINVAL = Int32(-1)
const SBP_EDIM = Int32(-2)

last_error() = unsafe_string(ccall((:sbp_last_error, libsbp), Cstring, ()))

function check(rc::Int32)
    rc == 0 && return nothing
    msg = last_error()
    rc == SBP_EDIM && throw(DimensionMismatch(msg))   # src/subcell_operators.jl:195-196
    rc == SBP_EINVAL && throw(ArgumentError(msg))     # @argcheck failures, src/polynomialbases_operators.jl:143-146
    error("libsbp_b200 error $rc: $msg")
end

# ---- context ------------------------------------------------------------------------------------------------------
# The context owns everything allocated through it: live device buffers and operator handles are registered here and
# released by `close` (also the context's finalizer), so the order in which Julia runs finalizers at exit does not matter:
# the finalizer of a batch / operator whose context is already closed is a no-op.
mutable struct Context
    handle::Ptr{Cvoid}
    buffers::Set{Ptr{Cvoid}}       # live sbp_malloc allocations
    ops::Vector{Ptr{Cvoid}}        # live sbp_op handles, in creation order (bundles after the operators they reference)
    function Context(device::Integer = 0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:sbp_ctx_create, libsbp), Int32, (Int32, Ref{Ptr{Cvoid}}), device, h))
        ctx = new(h[], Set{Ptr{Cvoid}}(), Ptr{Cvoid}[])
        finalizer(close, ctx)
        return ctx
    end
end
isopen(ctx::Context) = ctx.handle != C_NULL
function Base.close(ctx::Context)
    isopen(ctx) || return nothing
    ccall((:sbp_ctx_sync, libsbp), Int32, (Ptr{Cvoid},), ctx.handle)
    for op in reverse(ctx.ops)      # bundles first, then the operators they reference
        ccall((:sbp_op_destroy, libsbp), Int32, (Ptr{Cvoid},), op)
    end
    empty!(ctx.ops)
    for p in ctx.buffers
        ccall((:sbp_free, libsbp), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), ctx.handle, p)
    end
    empty!(ctx.buffers)
    ccall((:sbp_ctx_destroy, libsbp), Int32, (Ptr{Cvoid},), ctx.handle)
    ctx.handle = C_NULL
    return nothing
end
synchronize(ctx::Context) = check(ccall((:sbp_ctx_sync, libsbp), Int32, (Ptr{Cvoid},), ctx.handle))
Base.show(io::IO, ctx::Context) = print(io, "SBPB200.Context(", isopen(ctx) ? "open" : "closed", ", ",
                                        length(ctx.buffers), " buffers, ", length(ctx.ops), " operators)")

function device_malloc(ctx::Context, bytes::Integer)
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:sbp_malloc, libsbp), Int32, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), ctx.handle, bytes, p))
    push!(ctx.buffers, p[])
    return p[]
end
function device_free(ctx::Context, p::Ptr{Cvoid})
    (isopen(ctx) && p in ctx.buffers) || return nothing   # context already closed: it released the buffer itself
    delete!(ctx.buffers, p)
    ccall((:sbp_free, libsbp), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), ctx.handle, p)
    return nothing
end

# ---- device batch: N x B Float64, column-major (instance b at offset (b-1)*N) -----------------------------------------
# Deliberately NOT an AbstractMatrix: there is no scalar getindex on device memory, and generic fallbacks (show, broadcast,
# `D * u` of LinearAlgebra) must not be picked up silently.
mutable struct DeviceBatch
    ctx::Context
    ptr::Ptr{Float64}
    N::Int
    B::Int
    function DeviceBatch(ctx::Context, N::Integer, B::Integer)
        x = new(ctx, Ptr{Float64}(device_malloc(ctx, 8 * max(N * B, 2))), N, B)
        finalizer(b -> device_free(b.ctx, Ptr{Cvoid}(b.ptr)), x)
        return x
    end
end
function DeviceBatch(ctx::Context, host::AbstractVecOrMat{Float64})
    h = Matrix{Float64}(reshape(host, size(host, 1), :))
    x = DeviceBatch(ctx, size(h, 1), size(h, 2))
    copyto!(x, h)
    return x
end
function Base.copyto!(x::DeviceBatch, h::Matrix{Float64})
    size(h) == (x.N, x.B) || throw(DimensionMismatch("host matrix is $(size(h)), batch is ($(x.N), $(x.B))"))
    GC.@preserve h begin
        check(ccall((:sbp_memcpy_h2d, libsbp), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), x.ctx.handle, x.ptr, h,
                    sizeof(h)))
        synchronize(x.ctx)   # the copy is asynchronous: `h` must outlive it
    end
    return x
end
Base.size(x::DeviceBatch) = (x.N, x.B)
Base.size(x::DeviceBatch, d::Integer) = d == 1 ? x.N : (d == 2 ? x.B : 1)
Base.length(x::DeviceBatch) = x.N * x.B
Base.eltype(::DeviceBatch) = Float64
Base.similar(x::DeviceBatch) = DeviceBatch(x.ctx, x.N, x.B)
Base.show(io::IO, x::DeviceBatch) = print(io, "SBPB200.DeviceBatch(", x.N, " nodes x ", x.B, " instances, Float64, device)")
Base.show(io::IO, ::MIME"text/plain", x::DeviceBatch) = show(io, x)
function Base.Array(x::DeviceBatch)
    h = Matrix{Float64}(undef, x.N, x.B)
    check(ccall((:sbp_memcpy_d2h, libsbp), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), x.ctx.handle, h, x.ptr,
                sizeof(h)))   # synchronises
    return h
end

# ---- operator upload (once per operator and context; construction stays on the CPU) ---------------------------------
mutable struct DeviceOp
    ctx::Context
    handle::Ptr{Cvoid}
    function DeviceOp(ctx::Context, handle::Ptr{Cvoid})
        op = new(ctx, handle)
        push!(ctx.ops, handle)
        finalizer(destroy!, op)
        return op
    end
end
function destroy!(op::DeviceOp)
    (op.handle != C_NULL && isopen(op.ctx)) || return nothing
    i = findfirst(==(op.handle), op.ctx.ops)
    if i !== nothing
        deleteat!(op.ctx.ops, i)
        ccall((:sbp_op_destroy, libsbp), Int32, (Ptr{Cvoid},), op.handle)
    end
    op.handle = C_NULL
    return nothing
end
# (operator object, context handle) -> uploaded operator.  IdDict: the reference's operator types are immutable structs (no weak
# keys); `forget!(D)` drops the uploads of an operator that is no longer needed, `close(ctx)` releases everything of a context.
const OP_CACHE = IdDict{Any, Dict{Ptr{Cvoid}, Any}}()
function forget!(key)
    per_ctx = pop!(OP_CACHE, key, nothing)
    per_ctx === nothing && return nothing
    for v in values(per_ctx)
        v isa DeviceOp && destroy!(v)
        hasproperty(v, :op) && destroy!(v.op)
        hasproperty(v, :inner) && destroy!(v.inner)
    end
    return nothing
end
cached(make, key, ctx::Context) = get!(make, get!(() -> Dict{Ptr{Cvoid}, Any}(), OP_CACHE, key), ctx.handle)

# flags = 0: the library itself detects a centro-antisymmetric matrix (bitwise or up to rounding, as PolynomialBases' `basis.D`
# is) and serves it with the even/odd kernel (include/sbp_b200.h, SBP_DENSE_*)
function upload_dense(ctx::Context, Dmat::Matrix{Float64}, w::Vector{Float64}, scale::Float64)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    N = size(Dmat, 1)
    check(ccall((:sbp_op_dense_create, libsbp), Int32,
                (Ptr{Cvoid}, Int32, Ptr{Float64}, Int32, Ptr{Float64}, Float64, Int32, Ref{Ptr{Cvoid}}),
                ctx.handle, N, Dmat, stride(Dmat, 2), w, scale, 0, h))
    return DeviceOp(ctx, h[])
end
function upload_sparse(ctx::Context, Dmat::Matrix{Float64}, w::Vector{Float64}, scale::Float64)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    N = size(Dmat, 1)
    check(ccall((:sbp_op_sparse_from_dense, libsbp), Int32,
                (Ptr{Cvoid}, Int32, Ptr{Float64}, Int32, Ptr{Float64}, Float64, Ref{Ptr{Cvoid}}),
                ctx.handle, N, Dmat, stride(Dmat, 2), w, scale, h))
    return DeviceOp(ctx, h[])
end
# storage rule of the Python mirror (operators.py `_want_sparse`): banded / neighbourhood-sparse operators stored dense by the
# reference (ext/utils.jl:18-23) go to the block-sparse tensor-core kernel
function upload_auto(ctx::Context, Dmat::Matrix{Float64}, w::Vector{Float64}, scale::Float64)
    N = size(Dmat, 1)
    nz = count(!iszero, Dmat)
    sparse = N <= 128 ? (N >= 48 && nz < 0.25 * N^2) : nz < 0.15 * N^2
    return sparse ? upload_sparse(ctx, Dmat, w, scale) : upload_dense(ctx, Dmat, w, scale)
end

# what mul! multiplies by, per operator type (scale is folded on the device)
apply_matrix(D::PolynomialBasesDerivativeOperator) = (Matrix(D.basis.D), D.jac)            # polynomialbases_operators.jl:148
apply_matrix(D::SubcellOperator) = (Matrix(D), 1.0)                                         # subcell_operators.jl:183-186, once
apply_matrix(D::MatrixDerivativeOperator) = (Matrix(D.D), 1.0)                              # upstream
weights_of(D) = collect(diag(mass_matrix(D)))

device_op(ctx::Context, D) = cached(D, ctx) do
    Dmat, scale = apply_matrix(D)
    upload_auto(ctx, Dmat, weights_of(D), Float64(scale))
end

# ---- mul! -----------------------------------------------------------------------------------------------------------
const DeviceOperators = Union{PolynomialBasesDerivativeOperator, SubcellOperator, MatrixDerivativeOperator}

function check_sizes(N::Integer, dest::DeviceBatch, u::DeviceBatch)
    (N == u.N && N == dest.N) || throw(ArgumentError("N == length(u) must hold"))   # @argcheck of the reference
    dest.B == u.B || throw(DimensionMismatch("dest and u hold different numbers of instances"))
    return nothing
end
function apply!(dest::DeviceBatch, op::DeviceOp, u::DeviceBatch, α, β)
    check(ccall((:sbp_apply, libsbp), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Float64, Float64),
                u.ctx.handle, op.handle, dest.ptr, u.ptr, u.B, Float64(α), Float64(β)))
    return dest
end
function mul!(dest::DeviceBatch, D::DeviceOperators, u::DeviceBatch, α = true, β = false)
    check_sizes(size(D, 1), dest, u)
    return apply!(dest, device_op(u.ctx, D), u, α, β)
end
Base.:*(D::DeviceOperators, u::DeviceBatch) = mul!(similar(u), D, u)

# `D[dim] * u` for a MultidimensionalMatrixDerivativeOperator: in Julia `D[dim]` is a plain Matrix, so the direction is wrapped
struct DeviceDirection{Op <: MultidimensionalMatrixDerivativeOperator}
    parent::Op
    dim::Int
end
"""
    direction(D::MultidimensionalMatrixDerivativeOperator, dim)

`direction(D, dim) * u` / `mul!(du, direction(D, dim), u)` is `D[dim] * u` on a device batch (upstream `mul!(dest, D, u, dim)`).
"""
direction(D::MultidimensionalMatrixDerivativeOperator, dim::Integer) = DeviceDirection(D, Int(dim))
Base.size(d::DeviceDirection, i...) = size(d.parent[d.dim], i...)
device_op(ctx::Context, d::DeviceDirection) = cached((d.parent, d.dim), ctx) do
    upload_auto(ctx, Matrix(d.parent[d.dim]), collect(Float64, d.parent.weights), 1.0)
end
function mul!(dest::DeviceBatch, d::DeviceDirection, u::DeviceBatch, α = true, β = false)
    check_sizes(size(d, 1), dest, u)
    return apply!(dest, device_op(u.ctx, d), u, α, β)
end
Base.:*(d::DeviceDirection, u::DeviceBatch) = mul!(similar(u), d, u)
# upstream signature mul!(dest, D, u, dim, α, β)
mul!(dest::DeviceBatch, D::MultidimensionalMatrixDerivativeOperator, u::DeviceBatch, dim::Integer, α = true, β = false) =
    mul!(dest, direction(D, dim), u, α, β)

# host matrices: pinned-staging pipeline inside the library (H2D / compute / D2H overlapped)
function mul!(dest::Matrix{Float64}, D::DeviceOperators, u::Matrix{Float64}, ctx::Context, α = true, β = false)
    op = device_op(ctx, D)
    check(ccall((:sbp_apply_host, libsbp), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Float64, Float64, Int64),
                ctx.handle, op.handle, dest, u, size(u, 2), Float64(α), Float64(β), 0))
    return dest
end

# ---- a family of SubcellOperators (varying split point) applied to one batch ------------------------------------------------
"""
    OperatorTable(ops::Vector{<:SubcellOperator})

Instance `b` of a batch uses operator `op_index[b]` (0-based `Vector{Int32}` on the device) or `(b - 1) % G + 1` when no index is
given; `D = inv(P) (Q_left + Q_right)` (src/subcell_operators.jl:183-186) is formed once per operator instead of on every mul!.
"""
struct OperatorTable{Op}
    ops::Vector{Op}
end
device_op(ctx::Context, t::OperatorTable) = cached(t, ctx) do
    G, N = length(t.ops), size(t.ops[1], 1)
    Ds = Array{Float64}(undef, N, N, G)
    W = Array{Float64}(undef, N, G)
    for (g, D) in enumerate(t.ops)
        size(D, 1) == N || throw(DimensionMismatch("all operators of a table must have the same size"))
        Ds[:, :, g] = Matrix(D)
        W[:, g] = weights_of(D)
    end
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:sbp_op_table_create, libsbp), Int32,
                (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ref{Ptr{Cvoid}}), ctx.handle, N, G, Ds, W, h))
    DeviceOp(ctx, h[])
end
"""
    mul!(dest, table, u, α = true, β = false; energy = nothing, energy_sum = nothing, op_index = C_NULL)

`energy` (a `DeviceBatch(ctx, 1, B)`) receives `integrate(abs2, u_b, D_b)` per instance, `energy_sum` (a `DeviceBatch(ctx, 1, 1)`)
its sum over the batch, both fused into the apply pass.
"""
function mul!(dest::DeviceBatch, t::OperatorTable, u::DeviceBatch, α = true, β = false;
              energy::Union{DeviceBatch, Nothing} = nothing, energy_sum::Union{DeviceBatch, Nothing} = nothing,
              op_index::Ptr{Int32} = Ptr{Int32}(C_NULL))
    check_sizes(size(t.ops[1], 1), dest, u)
    op = device_op(u.ctx, t)
    check(ccall((:sbp_apply_table, libsbp), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int32}, Float64, Float64, Ptr{Float64},
                 Ptr{Float64}),
                u.ctx.handle, op.handle, dest.ptr, u.ptr, u.B, op_index, Float64(α), Float64(β),
                energy === nothing ? Ptr{Float64}(C_NULL) : energy.ptr,
                energy_sum === nothing ? Ptr{Float64}(C_NULL) : energy_sum.ptr))
    return dest
end

# ---- integrate(func, u, D), scale_by_(inverse_)mass_matrix! ------------------------------------------------------------------
integrate_mode(::typeof(identity)) = Int32(0)
integrate_mode(::typeof(abs2)) = Int32(1)
integrate_mode(::typeof(abs)) = Int32(3)
function integrate(func::Union{typeof(identity), typeof(abs2), typeof(abs)}, u::DeviceBatch, D::DeviceOperators)
    size(D, 2) == u.N || throw(DimensionMismatch("sizes of input vector and operator do not match"))
    op = device_op(u.ctx, D)
    out = DeviceBatch(u.ctx, 1, u.B)
    check(ccall((:sbp_integrate, libsbp), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Int32, Ptr{Float64}, Ptr{Float64}),
                u.ctx.handle, op.handle, u.ptr, C_NULL, u.B, integrate_mode(func), out.ptr, C_NULL))
    return vec(Array(out))
end
function scale_mass!(u::DeviceBatch, D::DeviceOperators, factor, inverse::Integer)
    size(D, 2) == u.N || throw(DimensionMismatch("sizes of input vector and operator do not match"))   # subcell_operators.jl:195
    op = device_op(u.ctx, D)
    check(ccall((:sbp_scale_by_mass_matrix, libsbp), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Int64, Float64, Int32),
                u.ctx.handle, op.handle, u.ptr, u.B, Float64(factor), inverse))
    return u
end
scale_by_mass_matrix!(u::DeviceBatch, D::DeviceOperators, factor = true) = scale_mass!(u, D, factor, 0)
scale_by_inverse_mass_matrix!(u::DeviceBatch, D::DeviceOperators, factor = true) = scale_mass!(u, D, factor, 1)

# ---- boundary data: one device buffer per (semidiscretization, context), refilled by an asynchronous copy ---------------------
mutable struct BoundaryBuffer
    dev::DeviceBatch            # N x 1 (2D: g per node) or 2 x 1 (1D: [left, right])
    host::Vector{Float64}       # staging: must stay alive until the copy has been consumed (stream order + `t` check below)
    t::Float64
end
function upload_boundary!(b::BoundaryBuffer, fillfn, t)
    if isnan(b.t) || b.t != t
        synchronize(b.dev.ctx)   # kernels that still read the previous contents / the previous staging copy
        fillfn(b.host, t)
        check(ccall((:sbp_memcpy_h2d, libsbp), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), b.dev.ctx.handle,
                    b.dev.ptr, b.host, sizeof(b.host)))
        b.t = t
    end
    return b.dev
end

# ---- 2D advection semidiscretization: semi(du, u, p, t) on device batches -------------------------------------------------
struct Bundle2D
    op::DeviceOp
    inner::DeviceOp
    mode::Vector{Int32}
    boundary::BoundaryBuffer
end
device_bundle(ctx::Context, semi::MultidimensionalLinearAdvectionNonperiodicSemidiscretization) = cached(semi, ctx) do
    D = semi.derivative
    A = Matrix(semi.cache.A)
    N = size(A, 1)
    w = collect(Float64, diag(mass_matrix(D)))
    inner = upload_auto(ctx, A, w, 1.0)
    b = collect(Float64, diag(semi.cache.B))
    mode = zeros(Int32, N)                      # set_bc! loop order: the last entry of a duplicated node wins (:60-71)
    tau = Float64[]
    node = Int32[]
    for (i, j) in enumerate(D.boundary_indices)
        an = dot(D.normals[i], semi.a)
        mode[j] = an < 0 ? 1 : 2
        push!(node, j - 1)
        push!(tau, SummationByPartsOperators.get_weight_boundary(D, i) * an)   # :101-103
    end
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:sbp_op_advection2d_create, libsbp), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Int32, Ptr{Int32}, Ptr{Float64},
                 Ref{Ptr{Cvoid}}),
                ctx.handle, inner.handle, b, w, mode, length(node), node, tau, h))
    Bundle2D(DeviceOp(ctx, h[]), inner, mode, BoundaryBuffer(DeviceBatch(ctx, N, 1), zeros(N), NaN))
end
function boundary_values!(g::Vector{Float64}, semi::MultidimensionalLinearAdvectionNonperiodicSemidiscretization, mode, t)
    x = grid(semi.derivative)
    for j in eachindex(x)
        g[j] = mode[j] == 1 ? semi.bc(x[j], t) : 0.0   # set_bc!: tmp1[j] = bc(node, t) on inflow nodes (:65-66)
    end
    return g
end
boundary_data(bd::Bundle2D, semi, t) = upload_boundary!(bd.boundary, (g, tt) -> boundary_values!(g, semi, bd.mode, tt), Float64(t))

function rhs2d!(du::DeviceBatch, semi, u::DeviceBatch, t, quantities::Ptr{Float64})
    check_sizes(size(semi.derivative, 1), du, u)
    bd = device_bundle(u.ctx, semi)
    g = boundary_data(bd, semi, t)
    check(ccall((:sbp_rhs_advection2d, libsbp), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}),
                u.ctx.handle, bd.op.handle, du.ptr, u.ptr, u.B, g.ptr, 0, quantities))
    return nothing
end
(semi::MultidimensionalLinearAdvectionNonperiodicSemidiscretization)(du::DeviceBatch, u::DeviceBatch, p, t) =
    rhs2d!(du, semi, u, t, Ptr{Float64}(C_NULL))

"""
    stage!(unext, semi, u, t, ca, ch[, cb, extra])

Runge-Kutta stage update riding on the RHS pass: `unext = ca*u + cb*extra + ch*semi(u, t)` in one kernel
(`sbp_rhs_advection2d_stage` / `sbp_rhs_advection1d_stage`); `extra` may be `unext` itself, `unext` must not be `u`.
"""
function stage!(unext::DeviceBatch, semi::MultidimensionalLinearAdvectionNonperiodicSemidiscretization, u::DeviceBatch, t,
                ca::Real, ch::Real, cb::Real = 0.0, extra::Union{DeviceBatch, Nothing} = nothing)
    bd = device_bundle(u.ctx, semi)
    g = boundary_data(bd, semi, t)
    check(ccall((:sbp_rhs_advection2d_stage, libsbp), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Float64, Float64,
                 Ptr{Float64}, Float64),
                u.ctx.handle, bd.op.handle, unext.ptr, u.ptr, u.B, g.ptr, 0, Float64(ca), Float64(cb),
                extra === nothing ? Ptr{Float64}(C_NULL) : extra.ptr, Float64(ch)))
    return unext
end

# analyze_quantities(semi, du, u, p, t) (:83-108): 7 x B, rows in the reference's order; re-evaluates the RHS into du like the
# callback does (analysis_callback.jl:123)
function analyze_quantities(semi::MultidimensionalLinearAdvectionNonperiodicSemidiscretization, du::DeviceBatch,
                            u::DeviceBatch, p, t)
    q = DeviceBatch(u.ctx, 7, u.B)
    rhs2d!(du, semi, u, t, q.ptr)
    return Array(q)
end

# ---- upstream 1D semidiscretization (examples/RBF_FSBP_advection.jl:37-38) ------------------------------------------------
struct Bundle1D
    op::DeviceOp
    boundary::BoundaryBuffer
end
device_bundle(ctx::Context, semi::VariableLinearAdvectionNonperiodicSemidiscretization) = cached(semi, ctx) do
    inner = device_op(ctx, semi.derivative)
    a = collect(Float64, semi.a)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:sbp_op_advection1d_create, libsbp), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ref{Ptr{Cvoid}}),
                ctx.handle, inner.handle, a, h))
    Bundle1D(DeviceOp(ctx, h[]), BoundaryBuffer(DeviceBatch(ctx, 2, 1), zeros(2), NaN))
end
boundary_data(bd::Bundle1D, semi, t) =
    upload_boundary!(bd.boundary, (b, tt) -> (b[1] = semi.left_bc(tt); b[2] = semi.right_bc(tt); b), Float64(t))

function rhs1d!(du::DeviceBatch, semi, u::DeviceBatch, t, quantities::Ptr{Float64})
    check_sizes(size(semi.derivative, 1), du, u)
    bd = device_bundle(u.ctx, semi)
    bc = boundary_data(bd, semi, t)
    check(ccall((:sbp_rhs_advection1d, libsbp), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}),
                u.ctx.handle, bd.op.handle, du.ptr, u.ptr, u.B, bc.ptr, 0, quantities))
    return nothing
end
(semi::VariableLinearAdvectionNonperiodicSemidiscretization)(du::DeviceBatch, u::DeviceBatch, p, t) =
    rhs1d!(du, semi, u, t, Ptr{Float64}(C_NULL))
function stage!(unext::DeviceBatch, semi::VariableLinearAdvectionNonperiodicSemidiscretization, u::DeviceBatch, t,
                ca::Real, ch::Real, cb::Real = 0.0, extra::Union{DeviceBatch, Nothing} = nothing)
    bd = device_bundle(u.ctx, semi)
    bc = boundary_data(bd, semi, t)
    check(ccall((:sbp_rhs_advection1d_stage, libsbp), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Float64, Float64,
                 Ptr{Float64}, Float64),
                u.ctx.handle, bd.op.handle, unext.ptr, u.ptr, u.B, bc.ptr, 0, Float64(ca), Float64(cb),
                extra === nothing ? Ptr{Float64}(C_NULL) : extra.ptr, Float64(ch)))
    return unext
end
function analyze_quantities(semi::VariableLinearAdvectionNonperiodicSemidiscretization, du::DeviceBatch, u::DeviceBatch, p, t)
    q = DeviceBatch(u.ctx, 7, u.B)
    rhs1d!(du, semi, u, t, q.ptr)
    return Array(q)
end

# ---- device-resident fixed-step SSPRK53 with the reference's AnalysisCallback ----------------------------------------------------
const SSPRK53_C = (0.0, 0.377268915331368, 0.754537830662736, 0.728985661612188, 0.699226135931670)

boundary_row!(row, bd::Bundle2D, semi, t) = boundary_values!(row, semi, bd.mode, t)
boundary_row!(row, bd::Bundle1D, semi, t) = (row[1] = semi.left_bc(t); row[2] = semi.right_bc(t); row)
bundle_row_length(bd::Bundle2D) = length(bd.mode)
bundle_row_length(::Bundle1D) = 2

"""
    record!(analysis_callback, semi, du, u, t)

What the functor of the reference's `AnalysisCallback` does with an integrator (src/conservation_laws/analysis_callback.jl:
115-131), for a device batch: `push!(tstops, t)`, re-evaluate the RHS, `push!(quantities, ...)` -- the 7 quantities of every
instance, instance after instance (for a batch of one exactly the reference's `Vector{Float64}` of length 7).
`analysis_callback` is the object returned by `AnalysisCallback(semi; interval | dt)`; `tstops(cb)` / `quantities(cb)` work as before.
"""
function record!(cb, semi, du::DeviceBatch, u::DeviceBatch, t)
    acb = hasproperty(cb.affect!, :affect!) ? cb.affect!.affect! : cb.affect!   # PeriodicCallback wraps the affect
    push!(acb.tstops, Float64(t))
    push!(acb.quantities, vec(analyze_quantities(semi, du, u, nothing, t)))
    return nothing
end

"""
    solve_ssprk53!(u, semi, tspan; dt, callback = nothing, callback_dt = nothing, callback_interval = 0)

`solve(semidiscretize(u0, semi, tspan), SSPRK53(); dt, adaptive = false, callback)` (examples/RBF_FSBP_advection.jl:44-51) for
a device batch `u` (overwritten with the final state): the time loop runs inside the library (`sbp_solve_ssprk53`, five fused
RHS + stage-update launches per step); the host evaluates the boundary functions and records `callback` (an
`AnalysisCallback(semi; dt = callback_dt)` or `AnalysisCallback(semi; interval = callback_interval)`) at the times
OrdinaryDiffEq would: every `callback_interval` accepted steps, or at `t0`, every multiple of `callback_dt` (the steps are
clipped to land on them, as the `tstops` of a PeriodicCallback do) and the final time.
"""
function solve_ssprk53!(u::DeviceBatch, semi, tspan; dt::Real, callback = nothing, callback_dt = nothing,
                        callback_interval::Integer = 0)
    t0, tend = Float64(tspan[1]), Float64(tspan[2])
    bd = device_bundle(u.ctx, semi)
    nrow = bundle_row_length(bd)
    du = callback === nothing ? nothing : similar(u)
    # step schedule (t, h) with the stop times of a PeriodicCallback
    stops = callback_dt === nothing ? Float64[] : collect(t0 + callback_dt:callback_dt:tend - 1e-12 * max(1.0, abs(tend)))
    marks = vcat(stops, tend)
    steps = Tuple{Float64, Float64}[]
    t = t0
    for m in marks
        while m - t > 1e-14 * max(1.0, abs(m))
            if t + dt >= m - 1e-12 * max(1.0, abs(m))
                push!(steps, (t, m - t)); t = m
            else
                push!(steps, (t, Float64(dt))); t += dt
            end
        end
    end
    fires(k, tn) = callback !== nothing && (callback_dt !== nothing ? (any(s -> abs(s - tn) <= 1e-12, stops) || k == length(steps)) :
                                            (callback_interval > 0 && (k % callback_interval == 0 || k == length(steps))))
    callback !== nothing && callback_dt !== nothing && record!(callback, semi, du, u, t0)   # initial_affect = true
    k = 0
    while k < length(steps)
        k1 = k
        while k1 < length(steps) && k1 - k < 64          # boundary data of up to 64 steps per upload
            k1 += 1
            fires(k1, k1 < length(steps) ? steps[k1 + 1][1] : tend) && break
        end
        seg = steps[(k + 1):k1]
        table = Matrix{Float64}(undef, nrow, 5 * length(seg))
        for (s, (ts, hs)) in enumerate(seg), (c, cc) in enumerate(SSPRK53_C)
            boundary_row!(view(table, :, 5 * (s - 1) + c), bd, semi, ts + cc * hs)
        end
        dtab = DeviceBatch(u.ctx, table)
        hsteps = Float64[h for (_, h) in seg]
        check(ccall((:sbp_solve_ssprk53, libsbp), Int32,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int64, Ptr{Float64}, Ptr{Float64}),
                    u.ctx.handle, bd.op.handle, u.ptr, u.B, dtab.ptr, nrow, length(seg), hsteps, C_NULL))
        synchronize(u.ctx)
        tn = k1 < length(steps) ? steps[k1 + 1][1] : tend
        fires(k1, tn) && record!(callback, semi, du, u, tn)
        k = k1
    end
    return u
end

# ---- low-storage 3S*+ FSAL schemes (RDPK3SpFSAL49 of examples/RBF_MFSBP_advection.jl:24) -------------------------------------
"""
    export_tableau(path, cache; order, embedded_order)

Write the coefficients of a 3S*+ FSAL low-storage scheme in the flat `SBPGOLD1` container that
`sbp_b200.LowStorage3SpTableau.load` and `sbp_solve_3sp_fsal` take.  `cache` is the constant cache of the algorithm in
OrdinaryDiffEqLowStorageRK, e.g. `OrdinaryDiffEqLowStorageRK.RDPK3SpFSAL49ConstantCache(Float64, Float64)` with
`order = 4, embedded_order = 3` (field names `γ12end, γ22end, γ32end, δ2end, β1, β2end, c2end, bhat1, bhat2end, bhatfsal` of
`LowStorageRK3SpFSALConstantCache`; the package is third party and absent from the build image, so check them against the
installed version).
"""
function export_tableau(path::AbstractString, cache; order::Integer, embedded_order::Integer)
    arrays = ["gamma1" => collect(Float64, cache.γ12end), "gamma2" => collect(Float64, cache.γ22end),
              "gamma3" => collect(Float64, cache.γ32end), "delta" => collect(Float64, cache.δ2end),
              "c" => collect(Float64, cache.c2end), "beta" => vcat(Float64(cache.β1), collect(Float64, cache.β2end)),
              "bhat" => vcat(Float64(cache.bhat1), collect(Float64, cache.bhat2end)), "bhat_fsal" => [Float64(cache.bhatfsal)],
              "order" => [Float64(order)], "embedded_order" => [Float64(embedded_order)]]
    open(path, "w") do io
        write(io, codeunits("SBPGOLD1"))
        write(io, Int64(length(arrays)))
        for (name, a) in arrays
            write(io, Int64(ncodeunits(name)))
            write(io, codeunits(name))
            write(io, Int64(1))
            write(io, Int64(length(a)))
            write(io, a)
        end
    end
    return path
end

# ---- portable operator container (replaces the `.jls` files of examples/RBF_MFSBP.jl:51-52) ---------------------------
# Flat little-endian binary file, layout documented in ../container.py (`SBPOP001`); read by `sbp_b200.load_operator`.
#   D = multidimensional_function_space_operator(...)        # built once on the CPU (Optim), as in the reference
#   SBPB200.export_operator("D_rectangle.sbpop", D)
function export_operator(path::AbstractString, D::MultidimensionalMatrixDerivativeOperator{Dim, Float64}) where {Dim}
    N = length(D.weights)
    Nb = length(D.boundary_indices)
    open(path, "w") do io
        write(io, codeunits("SBPOP001"))
        write(io, Int64[2, N, Dim, Nb, D.accuracy_order, 0])
        write(io, Float64[0.0, 0.0])
        for x in grid(D), d in 1:Dim            # nodes, node-major
            write(io, Float64(x[d]))
        end
        write(io, Vector{Float64}(D.weights))
        write(io, Vector{Int64}(D.boundary_indices))       # 1-based
        for n in D.normals, d in 1:Dim          # normals, entry-major
            write(io, Float64(n[d]))
        end
        write(io, Vector{Float64}(D.weights_boundary))
        for d in 1:Dim
            write(io, Matrix{Float64}(D[d]))    # column-major, as in memory
        end
    end
    return path
end
function export_operator(path::AbstractString, D::MatrixDerivativeOperator{Float64})
    N = length(grid(D))
    open(path, "w") do io
        write(io, codeunits("SBPOP001"))
        write(io, Int64[1, N, 1, 0, D.accuracy_order, 0])
        write(io, Float64[first(grid(D)), last(grid(D))])
        write(io, Vector{Float64}(grid(D)))
        write(io, Vector{Float64}(mass_matrix(D).diag))
        write(io, Matrix{Float64}(D))
    end
    return path
end

end # module
