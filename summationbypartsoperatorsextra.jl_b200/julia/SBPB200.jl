# SBPB200.jl -- Julia shim that keeps the reference API (`mul!(du, D, u)`, `semi(du, u, p, t)`, `integrate`,
# `analyze_quantities`) and routes device-resident batches to libsbp_b200.so through `ccall`.
#
# NOTE: this image has no Julia, so this file has not been executed here; it is the reference-side binding a
# maintainer adds (see INTEGRATION.md).  The identical C ABI is exercised by the Python `ctypes` mirror in tests/.
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
using SummationByPartsOperators: MatrixDerivativeOperator, MultidimensionalMatrixDerivativeOperator

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
mutable struct Context
    handle::Ptr{Cvoid}
    function Context(device::Integer = 0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:sbp_ctx_create, libsbp), Int32, (Int32, Ref{Ptr{Cvoid}}), device, h))
        ctx = new(h[])
        finalizer(c -> ccall((:sbp_ctx_destroy, libsbp), Int32, (Ptr{Cvoid},), c.handle), ctx)
        return ctx
    end
end
synchronize(ctx::Context) = check(ccall((:sbp_ctx_sync, libsbp), Int32, (Ptr{Cvoid},), ctx.handle))

# ---- device batch: N x B Float64, column-major (instance b at offset (b-1)*N) -----------------------------------------
mutable struct DeviceBatch <: AbstractMatrix{Float64}
    ctx::Context
    ptr::Ptr{Float64}
    N::Int
    B::Int
    function DeviceBatch(ctx::Context, N::Integer, B::Integer)
        p = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:sbp_malloc, libsbp), Int32, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), ctx.handle, 8 * N * B, p))
        x = new(ctx, Ptr{Float64}(p[]), N, B)
        finalizer(b -> ccall((:sbp_free, libsbp), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), b.ctx.handle, b.ptr), x)
        return x
    end
end
function DeviceBatch(ctx::Context, host::AbstractVecOrMat{Float64})
    h = Matrix{Float64}(reshape(host, size(host, 1), :))
    x = DeviceBatch(ctx, size(h, 1), size(h, 2))
    check(ccall((:sbp_memcpy_h2d, libsbp), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), ctx.handle, x.ptr, h,
                sizeof(h)))
    synchronize(ctx)
    return x
end
Base.size(x::DeviceBatch) = (x.N, x.B)
Base.similar(x::DeviceBatch) = DeviceBatch(x.ctx, x.N, x.B)
Base.getindex(::DeviceBatch, i...) = error("DeviceBatch lives on the GPU: use Array(x) to download")
function Base.Array(x::DeviceBatch)
    h = Matrix{Float64}(undef, x.N, x.B)
    check(ccall((:sbp_memcpy_d2h, libsbp), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), x.ctx.handle, h, x.ptr,
                sizeof(h)))
    return h
end

# ---- operator upload (once per operator and context; construction stays on the CPU) ---------------------------------
mutable struct DeviceOp
    handle::Ptr{Cvoid}
end
const OP_CACHE = IdDict{Any, Dict{Ptr{Cvoid}, DeviceOp}}()

function upload_dense(ctx::Context, Dmat::Matrix{Float64}, w::Vector{Float64}, scale::Float64)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    N = size(Dmat, 1)
    check(ccall((:sbp_op_dense_create, libsbp), Int32,
                (Ptr{Cvoid}, Int32, Ptr{Float64}, Int32, Ptr{Float64}, Float64, Int32, Ref{Ptr{Cvoid}}),
                ctx.handle, N, Dmat, stride(Dmat, 2), w, scale, 0, h))
    return DeviceOp(h[])
end
function upload_sparse(ctx::Context, Dmat::Matrix{Float64}, w::Vector{Float64}, scale::Float64)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    N = size(Dmat, 1)
    check(ccall((:sbp_op_sparse_from_dense, libsbp), Int32,
                (Ptr{Cvoid}, Int32, Ptr{Float64}, Int32, Ptr{Float64}, Float64, Ref{Ptr{Cvoid}}),
                ctx.handle, N, Dmat, stride(Dmat, 2), w, scale, h))
    return DeviceOp(h[])
end

# what mul! multiplies by, per operator type (scale is folded on the device)
apply_matrix(D::PolynomialBasesDerivativeOperator) = (Matrix(D.basis.D), D.jac)            # polynomialbases_operators.jl:148
apply_matrix(D::SubcellOperator) = (Matrix(D), 1.0)                                         # subcell_operators.jl:183-186, once
apply_matrix(D::MatrixDerivativeOperator) = (Matrix(D.D), 1.0)                              # upstream
weights_of(D) = collect(diag(mass_matrix(D)))

function device_op(ctx::Context, D)
    per_ctx = get!(() -> Dict{Ptr{Cvoid}, DeviceOp}(), OP_CACHE, D)
    return get!(per_ctx, ctx.handle) do
        Dmat, scale = apply_matrix(D)
        N = size(Dmat, 1)
        sparse = N > 128 && count(!iszero, Dmat) < 0.15 * N^2
        sparse ? upload_sparse(ctx, Dmat, weights_of(D), scale) : upload_dense(ctx, Dmat, weights_of(D), scale)
    end
end

# ---- mul! -----------------------------------------------------------------------------------------------------------
const DeviceOperators = Union{PolynomialBasesDerivativeOperator, SubcellOperator, MatrixDerivativeOperator}

function mul!(dest::DeviceBatch, D::DeviceOperators, u::DeviceBatch, α = true, β = false)
    N = size(D, 1)
    (N == u.N && N == dest.N) || throw(ArgumentError("N == length(u) must hold"))
    dest.B == u.B || throw(DimensionMismatch("dest and u hold different numbers of instances"))
    op = device_op(u.ctx, D)
    check(ccall((:sbp_apply, libsbp), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Float64, Float64),
                u.ctx.handle, op.handle, dest.ptr, u.ptr, u.B, Float64(α), Float64(β)))
    return dest
end
Base.:*(D::DeviceOperators, u::DeviceBatch) = mul!(similar(u), D, u)

# host matrices: pinned-staging pipeline inside the library (H2D / compute / D2H overlapped)
function mul!(dest::Matrix{Float64}, D::DeviceOperators, u::Matrix{Float64}, ctx::Context, α = true, β = false)
    op = device_op(ctx, D)
    check(ccall((:sbp_apply_host, libsbp), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Float64, Float64, Int64),
                ctx.handle, op.handle, dest, u, size(u, 2), Float64(α), Float64(β), 0))
    return dest
end

# ---- integrate(func, u, D) for func in (identity, abs2) ------------------------------------------------------------------
function integrate(func::Union{typeof(identity), typeof(abs2)}, u::DeviceBatch, D::DeviceOperators)
    op = device_op(u.ctx, D)
    out = DeviceBatch(u.ctx, 1, u.B)
    check(ccall((:sbp_integrate, libsbp), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Int32, Ptr{Float64}, Ptr{Float64}),
                u.ctx.handle, op.handle, u.ptr, C_NULL, u.B, func === abs2 ? 1 : 0, out.ptr, C_NULL))
    return vec(Array(out))
end

# ---- 2D advection semidiscretization: semi(du, u, p, t) on device batches -------------------------------------------------
const ADV_CACHE = IdDict{Any, Any}()

function device_bundle(ctx::Context, semi::MultidimensionalLinearAdvectionNonperiodicSemidiscretization)
    get!(ADV_CACHE, (semi, ctx.handle)) do
        D = semi.derivative
        A = Matrix(semi.cache.A)
        N = size(A, 1)
        w = collect(diag(mass_matrix(D)))
        inner = count(!iszero, A) < 0.15 * N^2 && N > 128 ? upload_sparse(ctx, A, w, 1.0) : upload_dense(ctx, A, w, 1.0)
        b = collect(diag(semi.cache.B))
        mode = zeros(Int32, N)                      # set_bc! loop order: the last entry of a duplicated node wins
        tau = Float64[]
        node = Int32[]
        for (i, j) in enumerate(D.boundary_indices)
            an = dot(D.normals[i], semi.a)
            mode[j] = an < 0 ? 1 : 2
            push!(node, j - 1)
            push!(tau, D.weights_boundary[i] * an)
        end
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:sbp_op_advection2d_create, libsbp), Int32,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Int32, Ptr{Int32}, Ptr{Float64},
                     Ref{Ptr{Cvoid}}),
                    ctx.handle, inner.handle, b, w, mode, length(node), node, tau, h))
        (DeviceOp(h[]), inner, mode)
    end
end

function boundary_data(ctx::Context, semi, mode, t)
    x = grid(semi.derivative)
    g = zeros(length(x))
    for j in eachindex(x)
        mode[j] == 1 && (g[j] = semi.bc(x[j], t))   # set_bc!: tmp1[j] = bc(node, t) on inflow nodes
    end
    return DeviceBatch(ctx, g)
end

function (semi::MultidimensionalLinearAdvectionNonperiodicSemidiscretization)(du::DeviceBatch, u::DeviceBatch, p, t)
    op, _, mode = device_bundle(u.ctx, semi)
    g = boundary_data(u.ctx, semi, mode, t)
    check(ccall((:sbp_rhs_advection2d, libsbp), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}),
                u.ctx.handle, op.handle, du.ptr, u.ptr, u.B, g.ptr, 0, C_NULL))
    return nothing
end

"""
    stage!(unext, semi, u, t, ca, ch[, cb, extra])

Runge-Kutta stage update riding on the RHS pass: `unext = ca*u + cb*extra + ch*semi(u, t)` in one kernel
(`sbp_rhs_advection2d_stage`); `extra` may be `unext` itself, `unext` must not be `u`.  A fixed-step SSPRK53 step
(examples/RBF_FSBP_advection.jl:45-51) is five such calls and never leaves the device.
"""
function stage!(unext::DeviceBatch, semi::MultidimensionalLinearAdvectionNonperiodicSemidiscretization, u::DeviceBatch, t,
                ca::Real, ch::Real, cb::Real = 0.0, extra::Union{DeviceBatch, Nothing} = nothing)
    op, _, mode = device_bundle(u.ctx, semi)
    g = boundary_data(u.ctx, semi, mode, t)
    check(ccall((:sbp_rhs_advection2d_stage, libsbp), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Float64, Float64,
                 Ptr{Float64}, Float64),
                u.ctx.handle, op.handle, unext.ptr, u.ptr, u.B, g.ptr, 0, Float64(ca), Float64(cb),
                extra === nothing ? Ptr{Float64}(C_NULL) : extra.ptr, Float64(ch)))
    return unext
end

function analyze_quantities(semi::MultidimensionalLinearAdvectionNonperiodicSemidiscretization, du::DeviceBatch,
                            u::DeviceBatch, p, t)
    op, _, mode = device_bundle(u.ctx, semi)
    g = boundary_data(u.ctx, semi, mode, t)
    q = DeviceBatch(u.ctx, 7, u.B)
    check(ccall((:sbp_rhs_advection2d, libsbp), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}),
                u.ctx.handle, op.handle, du.ptr, u.ptr, u.B, g.ptr, 0, q.ptr))
    return Array(q)   # 7 x B, rows in the reference's order
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
