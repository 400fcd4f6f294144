# julia_oracle.jl -- regenerate the OUTPUT arrays of tests/golden/hotpath_golden.npz from the REAL reference
# (SummationByPartsOperatorsExtra.jl + SummationByPartsOperators.jl), so that the numpy oracle and the CUDA path can be
# pinned to Julia's own `mul!`, `semi(du, u, p, t)` and `analyze_quantities` on identical inputs.
#
# This image has no Julia (SURVEY.md section 8c), so this script has NOT been executed here.  On a Julia machine:
#
#     julia --project=/path/to/SummationByPartsOperatorsExtra.jl baseline/julia_oracle.jl \
#           tests/golden/hotpath_inputs.bin tests/golden/hotpath_golden_julia.bin
#
# then commit `hotpath_golden_julia.bin`; `tests/test_julia_golden.py` picks it up (CPU suite: the oracle against it,
# GPU suite: the CUDA path against it) and the bench line's "parity" note can drop "unpinned".
#
# Input / output container: `SBPGOLD1` (tests/golden/golden_io.py) -- magic, Int64 count, per array: Int64 name length,
# name, Int64 ndim, Int64 dims (numpy C-order shape), Float64 data in C order.  A batch is stored as numpy (B, N), which is
# exactly a Julia Matrix(N, B) written column by column.
#
# Every block below cites the generator lines of tests/golden/make_golden.py it mirrors and the reference method it calls.
using LinearAlgebra
using SummationByPartsOperatorsExtra   # re-exports SummationByPartsOperators and PolynomialBases

# ---- SBPGOLD1 ---------------------------------------------------------------------------------------------------------
function read_gold(path)
    out = Dict{String, Array{Float64}}()
    open(path, "r") do io
        String(read(io, 8)) == "SBPGOLD1" || error("not an SBPGOLD1 file: $path")
        count = read(io, Int64)
        for _ in 1:count
            name = String(read(io, read(io, Int64)))
            nd = read(io, Int64)
            dims = Int[read(io, Int64) for _ in 1:nd]           # numpy C-order shape, e.g. (B, N)
            data = Vector{Float64}(undef, prod(dims))
            read!(io, data)
            # C order (B, N)  ==  column-major (N, B): reverse the dims
            out[name] = nd == 0 ? fill(data[1]) : reshape(data, reverse(dims)...)
        end
    end
    return out
end

function write_gold(path, arrays::Vector{Pair{String, Array{Float64}}})
    open(path, "w") do io
        write(io, codeunits("SBPGOLD1"))
        write(io, Int64(length(arrays)))
        for (name, a) in arrays
            write(io, Int64(ncodeunits(name)))
            write(io, codeunits(name))
            write(io, Int64(ndims(a)))
            for d in reverse(size(a))                            # back to the numpy shape
                write(io, Int64(d))
            end
            write(io, vec(a))                                    # column-major (N, B) == C order (B, N)
        end
    end
end

# apply `f(u) -> vector` to every column (instance) of U (N x B) and stack the results as columns
columns(f, U) = reduce(hcat, [f(U[:, b]) for b in axes(U, 2)])

function main(inpath, outpath)
    inp = read_gold(inpath)
    out = Pair{String, Array{Float64}}[]

    # (1) mul! with a PolynomialBasesDerivativeOperator (src/polynomialbases_operators.jl:139-149), alpha/beta form, energy
    #     make_golden.py: D = PolynomialBasesOperator("LobattoLegendre", -2.0, 1.4, 16)
    D = PolynomialBasesDerivativeOperator(LobattoLegendre, -2.0, 1.4, 16)
    U, Y0 = inp["lgl16_U"], inp["lgl16_Y0"]                      # 16 x 5
    push!(out, "lgl16_dU" => columns(u -> mul!(similar(u), D, u), U))
    push!(out, "lgl16_dU_axpby" => reduce(hcat, [mul!(copy(Y0[:, b]), D, U[:, b], -1.5, 0.25) for b in axes(U, 2)]))
    push!(out, "lgl16_energy" => Float64[integrate(abs2, U[:, b], D) for b in axes(U, 2)])
    push!(out, "lgl16_basis_D" => Matrix{Float64}(D.basis.D))   # the matrix mul! multiplies by (for construction parity)

    # (2) banded FD-SBP operator behind the upstream MatrixDerivativeOperator mul! (dense storage, as function_space_operator
    #     returns it: ext/function_space_operators.jl:13-14), N = 20 and odd N = 21
    for N in (20, 21)
        Dfd = derivative_operator(MattssonNordström2004(), 1, 4, 0.0, 1.0, N)
        Dm = MatrixDerivativeOperator(0.0, 1.0, collect(grid(Dfd)), diag(mass_matrix(Dfd)), Matrix(Dfd), 4,
                                      MattssonNordström2004())
        push!(out, "mn4_$(N)_dU" => columns(u -> mul!(similar(u), Dm, u), inp["mn4_$(N)_U"]))
    end

    # (3) SubcellOperator family (src/subcell_operators.jl:270-280) with the energy integrate(abs2, u, D) (:92-94)
    #     make_golden.py: N = 16, N_L = 4..11, x_M = -1 + 2 N_L / N, coupling :no for even i, :central for odd i (0-based)
    N = 16
    U = inp["subcell_U"]                                          # 16 x 8
    dU = zeros(N, 8)
    E = zeros(8)
    for (i, NL) in enumerate(4:11)
        xM = -1.0 + 2.0 * NL / N
        Dl = legendre_derivative_operator(-1.0, xM, NL)
        Dr = legendre_derivative_operator(xM, 1.0, N - NL)
        Ds = couple_subcell(Dl, Dr, xM; coupling = isodd(i) ? Val(:no) : Val(:central))   # i is 1-based here
        dU[:, i] = mul!(zeros(N), Ds, U[:, i])
        E[i] = integrate(abs2, U[:, i], Ds)
    end
    push!(out, "subcell_dU" => dU)
    push!(out, "subcell_energy" => E)

    # (4) 2D advection RHS + analysis quantities (src/conservation_laws/multidimensional_linear_advection.jl:60-108) on the
    #     tensor-product operator the reference's own tests use (test/test_unit.jl:108-118)
    D1 = derivative_operator(MattssonNordström2004(), 1, 2, -1.0, 1.0, 8)
    D2 = tensor_product_operator_2D(D1)
    a, t = (1.0, -0.5), 0.37
    g(x, tt) = exp(-30 * ((x[1] - a[1] * tt)^2 + (x[2] - a[2] * tt)^2)) + 0.1 * sin(3 * x[1] - x[2] + tt)
    semi = MultidimensionalLinearAdvectionNonperiodicSemidiscretization(D2, a, g)
    U = inp["adv2d_U"]                                            # 64 x 3
    rhs = similar(U)
    q = zeros(7, size(U, 2))
    for b in axes(U, 2)
        du = zeros(size(U, 1))
        semi(du, U[:, b], nothing, t)                             # leaves tmp1 of THIS instance in the cache
        rhs[:, b] = du
        q[:, b] = analyze_quantities(semi, du, U[:, b], nothing, t)
    end
    push!(out, "adv2d_dU" => rhs)
    push!(out, "adv2d_quantities" => q)
    # what the host mirror must reproduce when it assembles the bundle (boundary entry order, corner handling)
    push!(out, "adv2d_boundary_indices" => Float64.(D2.boundary_indices))          # 1-based
    push!(out, "adv2d_normals" => reduce(hcat, [Float64[n[1], n[2]] for n in D2.normals]))
    push!(out, "adv2d_Bdiag" => Float64.(diag(semi.cache.B)))
    push!(out, "adv2d_nodes" => reduce(hcat, [Float64[x[1], x[2]] for x in grid(D2)]))

    # (5) upstream 1D semidiscretization with Godunov SATs (examples/RBF_FSBP_advection.jl:37-38) + its quantities
    #     (src/conservation_laws/analysis_callback.jl:136-173)
    D1d = derivative_operator(MattssonNordström2004(), 1, 4, 0.0, 1.0, 21)
    semi1 = VariableLinearAdvectionNonperiodicSemidiscretization(D1d, nothing, x -> 1.0, Val(false),
                                                                 tt -> tanh(50 * (-tt - 0.1)), tt -> 0.3 * cos(tt))
    U = inp["adv1d_U"]                                            # 21 x 3
    rhs1 = similar(U)
    q1 = zeros(7, size(U, 2))
    for b in axes(U, 2)
        du = zeros(size(U, 1))
        semi1(du, U[:, b], nothing, 0.2)
        rhs1[:, b] = du
        q1[:, b] = analyze_quantities(semi1, du, U[:, b], nothing, 0.2)
    end
    push!(out, "adv1d_dU" => rhs1)
    push!(out, "adv1d_quantities" => q1)

    write_gold(outpath, out)
    println("wrote ", outpath, " (", length(out), " arrays)")
end

length(ARGS) == 2 || error("usage: julia --project=<reference> baseline/julia_oracle.jl <hotpath_inputs.bin> <hotpath_golden_julia.bin>")
main(ARGS[1], ARGS[2])
