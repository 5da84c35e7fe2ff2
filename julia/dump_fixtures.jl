# dump_fixtures.jl - writes the arrays of the reference's own run of scripts/poisson_simple.jl (config 1:
# nrefs = 3, order = 2, unit square, 5 initial modes) so that the CPU oracle of this repository (oracle/*.py) can be
# PINNED against the reference itself:
#
#     julia --project=<ExtendableASGFEM.jl checkout> julia/dump_fixtures.jl <output directory>
#     ASGFEM_JULIA_DUMP=<output directory> python -m pytest tests/test_oracle_vs_julia_dump.py
#
# No Julia is installed in the build image of this repository, so the dump is not committed; the comparator test skips
# when the directory is absent (SURVEY.md section 7.1).  Format: one raw little-endian file per array (Float64 / Int64,
# column-major as Julia stores it) plus manifest.json with name -> {dtype, shape}.
#
# Follows scripts/poisson_simple.jl:18-87 and src/modelproblems/poisson_primal.jl:37-80 line by line; only `solve!` is
# opened up so that the assembled matrices can be written before solve_primal! is called.
using ExtendableASGFEM
using ExtendableASGFEM: solve_primal!, MySystemPrimal, get_am_x, maxlength_multiindices, prepare_multi_indices!,
    add_boundary_modes, estimate
using ExtendableFEM
using ExtendableFEMBase
using ExtendableGrids
using SparseArrays
using LinearAlgebra
using Random

outdir = length(ARGS) >= 1 ? ARGS[1] : "julia_dump"
mkpath(outdir)
manifest = Dict{String, Any}()

function dump(name, A::AbstractArray{T}) where {T <: Union{Float64, Int64}}
    open(joinpath(outdir, name * ".bin"), "w") do io
        write(io, Array(A))
    end
    manifest[name] = Dict("dtype" => (T == Float64 ? "f8" : "i8"), "shape" => collect(size(A)))
    return nothing
end
dump(name, A::AbstractArray{<:Integer}) = dump(name, Int64.(A))
function dump_csc(name, S::SparseMatrixCSC)
    dump(name * "_colptr", Int64.(S.colptr))
    dump(name * "_rowval", Int64.(S.rowval))
    dump(name * "_nzval", Float64.(S.nzval))
    manifest[name] = Dict("csc" => true, "m" => size(S, 1), "n" => size(S, 2))
    return nothing
end

# ---- scripts/poisson_simple.jl:18-68 -------------------------------------------------------------------------
nrefs, order, decay, mean = 3, 2, 2.0, 1.0
initial_modes = [[0], [1, 0], [0, 1], [2, 0], [0, 0, 1]]
f! = (result, qpinfo) -> (result[1] = 1)
C = StochasticCoefficientCosinus(; τ = 0.9, decay = decay, mean = mean)
xgrid = uniform_refine(grid_unitsquare(Triangle2D), nrefs)
multi_indices = Array{Array{Int, 1}, 1}(initial_modes)
prepare_multi_indices!(multi_indices)
M = maximum(length.(multi_indices))
ansatz_deg = maximum([maximum(multi_indices[k]) for k in 1:length(multi_indices)]) + 4
TB = TensorizedBasis(LegendrePolynomials, M, ansatz_deg, 2 * ansatz_deg, 2 * ansatz_deg, multi_indices = multi_indices)
FES = FESpace{H1Pk{1, 2, order}}(xgrid)
sol = SGFEVector(FES, TB; active_modes = 1:length(multi_indices), unames = ["u"])

dump("coords", xgrid[Coordinates])
dump("cellnodes", xgrid[CellNodes])
dump("celldofs", FES[CellDofs])
dump("bfacedofs", FES[BFaceDofs])
dump("multi_indices", hcat(multi_indices...))            # M x N
dump_csc("G", SparseMatrixCSC(TB.G))                      # flushed (M*N) x N coupling matrix

# ---- src/modelproblems/poisson_primal.jl:56-71 ---------------------------------------------------------------
A0 = FEMatrix(FES, FES)
assemble!(A0, BilinearOperator(get_am_x(0, C), [grad(1)], [grad(1)]; bonus_quadorder = 2))
A = []
for m in 1:maxlength_multiindices(TB)
    Am = FEMatrix(FES, FES)
    assemble!(Am, BilinearOperator(get_am_x(m, C), [grad(1)], [grad(1)]; bonus_quadorder = 2))
    push!(A, Am)
end
b = FEVector(FES)
assemble!(b, LinearOperator(f!, [id(1)]; bonus_quadorder = 0))
flush!(A0.entries)
dump_csc("A0", SparseMatrixCSC(A0.entries.cscmatrix))
for (m, Am) in enumerate(A)
    flush!(Am.entries)
    dump_csc("A$(m)", SparseMatrixCSC(Am.entries.cscmatrix))
end
dump("b0", Float64.(b.entries))

# ---- one operator application on a seeded vector (mul!, solvers_poisson_primal.jl:86-124) ----------------------
n, N = FES.ndofs, TB.nmodes
Random.seed!(20240)
x = 2 .* rand(n * N) .- 1
dump("apply_x", x)
bdofs_probe = unique(vec(FES[BFaceDofs][:, :]))
S = MySystemPrimal(A0, A, TB.G, bdofs_probe, N)
Ax = zeros(n * N)
mul!(Ax, S, x)
dump("apply_Ax", Ax)

# ---- the solve (poisson_primal.jl:74-80 -> solve_primal!) and the estimator (estimate.jl:260-418) ----------------
bdofs = solve_primal!(sol, A0, A, b, TB.G, TB.nmodes, 1)
dump("bdofs", Int64.(bdofs))
dump("solution", Float64.(sol.entries))
multi_indices_extended = add_boundary_modes(deepcopy(multi_indices))
dump("multi_indices_extended", hcat(multi_indices_extended...))
eta4modes, eta4cell, mie = estimate(PoissonProblemPrimal, sol, C; rhs = f!, bonus_quadorder = 1)[1:3]
dump("eta4modes", Float64.(eta4modes))
dump("eta4cell", Float64.(eta4cell))

open(joinpath(outdir, "manifest.json"), "w") do io
    # minimal JSON writer (no JSON.jl dependency)
    print(io, "{")
    first = true
    for (k, v) in manifest
        first || print(io, ",")
        first = false
        print(io, "\"", k, "\":{")
        print(io, join(["\"$(kk)\":" * (vv isa AbstractString ? "\"$(vv)\"" : (vv isa Bool ? string(vv) : (vv isa AbstractVector ? "[" * join(vv, ",") * "]" : string(vv)))) for (kk, vv) in v], ","))
        print(io, "}")
    end
    print(io, "}")
end
println("wrote ", length(manifest), " arrays to ", outdir)
