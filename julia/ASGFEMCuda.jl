# ASGFEMCuda.jl - reference-side binding of libasgfem_cuda.so (include/asgfem.h).
#
# What a maintainer of ExtendableASGFEM.jl v1.0.1 would add to route the solve hot path to the GPU.
# It keeps the Julia API (SGFEVector, TensorizedBasis, solve!, estimate, scripts/poisson.jl) untouched and
# overrides the three seams of SURVEY.md §8(b):
#
#   solve_primal!(sol, A0, Am, b0, G, nmodes, bfac; atol, rtol)   src/modelproblems/solvers_poisson_primal.jl:130
#   LinearAlgebra.mul!(Ax, S::MySystemPrimal, x)                   src/modelproblems/solvers_poisson_primal.jl:86
#   LinearAlgebra.ldiv!(y, P::MyPreconditionerPrimal, b)           src/modelproblems/solvers_poisson_primal.jl:47
#   estimate(::Type{PoissonProblemPrimal}, sol, C; ...)            src/estimate.jl:260
#
# NOT executed in this repository's CI (no Julia in the image); every ccall below mirrors a prototype of
# include/asgfem.h one to one, and the Python ctypes binding (extendableasgfem.jl_b200/_lib.py) exercises the
# same entry points with the same argument conventions (1-based Int64 CSC, flat column-major entries).
module ASGFEMCuda

using ExtendableASGFEM
using ExtendableASGFEM: SGFEVector, TensorizedBasis, PoissonProblemPrimal, StochasticCoefficientCosinus,
    OrthogonalPolynomialType, LegendrePolynomials, HermitePolynomials, add_boundary_modes
using ExtendableFEMBase: BFaceDofs, CellDofs, FESpace, QuadratureRule, get_polynomialorder, num_sources
using ExtendableGrids: Coordinates, CellNodes, UniqueCellGeometries, Edge1D, num_cells
using SparseArrays: SparseMatrixCSC
using LinearAlgebra

const LIB = get(ENV, "ASGFEM_CUDA_LIB", "libasgfem_cuda.so")

struct Stats
    niter::Int64
    solved::Int32
    _pad::Int32
    rz0::Float64
    rzk::Float64
    residual::Float64
    ms_setup::Float64
    ms_iterations::Float64
    ms_apply::Float64
    ms_precond::Float64
end

mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer = 0)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:asgfem_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint), ref, device)
        rc == 0 || error("asgfem_create: " * unsafe_string(ccall((:asgfem_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
        ctx = new(ref[])
        finalizer(c -> ccall((:asgfem_destroy, LIB), Cint, (Ptr{Cvoid},), c.h), ctx)
        return ctx
    end
end

function check(ctx::Context, rc)
    rc == 0 && return
    error("libasgfem_cuda error $rc: " * unsafe_string(ccall((:asgfem_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.h)))
end

family(::Type{LegendrePolynomials}) = Cint(0)
family(::Type{HermitePolynomials}) = Cint(1)

"M x N Int64 matrix of the (already padded) multi-indices"
mimatrix(multi_indices) = Matrix{Int64}(reduce(hcat, multi_indices))

function set_multiindices!(ctx::Context, OBT, multi_indices)
    mi = mimatrix(multi_indices)
    check(ctx, ccall((:asgfem_set_multiindices, LIB), Cint, (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Int64}),
        ctx.h, family(OBT), size(mi, 2), size(mi, 1), mi))
end

"uploads A0 and Am (FEMatrix objects of poisson_primal.jl:56-63) into the shared pattern"
function set_matrices!(ctx::Context, A0, Am; extra = ())
    csc0::SparseMatrixCSC{Float64, Int64} = A0.entries.cscmatrix
    pattern = abs.(csc0)
    for A in Am            # union pattern (ExtendableSparse may have dropped exact zeros in some K_m)
        pattern += abs.(A.entries.cscmatrix)
    end
    for E in extra         # further CSC matrices that must fit the pattern (preconditioner matrix)
        pattern += abs.(E)
    end
    n = size(pattern, 1)
    check(ctx, ccall((:asgfem_set_pattern_csc, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}),
        ctx.h, n, pattern.colptr, pattern.rowval))
    check(ctx, ccall((:asgfem_set_num_stiffness, LIB), Cint, (Ptr{Cvoid}, Cint), ctx.h, length(Am)))
    for (m, A) in enumerate(vcat([A0], Am))
        c::SparseMatrixCSC{Float64, Int64} = A.entries.cscmatrix
        check(ctx, ccall((:asgfem_set_stiffness_csc, LIB), Cint,
            (Ptr{Cvoid}, Cint, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}), ctx.h, m - 1, c.colptr, c.rowval, c.nzval))
    end
end

function boundary_dofs(FES)  # solvers_poisson_primal.jl:136-142
    bfacedofs = FES[BFaceDofs]
    bdofs = Int64[]
    for bface in 1:num_sources(bfacedofs)
        append!(bdofs, view(bfacedofs, :, bface))
    end
    return unique(bdofs)
end

# ---- seam 1: the production seam -----------------------------------------------------------------------
"""
Drop-in replacement of `ExtendableASGFEM.solve_primal!`: same arguments, overwrites `sol.entries`, returns `bdofs`.
The Krylov loop (PCG with the mean-based preconditioner) runs entirely on the GPU.
"""
function solve_primal!(sol::SGFEVector, A0, Am, b0, G, nmodes, bfac; atol = 1.0e-14, rtol = 1.0e-14, device = 0)
    ctx = Context(device)
    OBT = OrthogonalPolynomialType(sol.TB.ONB)
    set_multiindices!(ctx, OBT, sol.TB.multi_indices)
    set_matrices!(ctx, A0, Am)
    bdofs = boundary_dofs(sol.FES_space[1])
    check(ctx, ccall((:asgfem_set_bdofs, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}), ctx.h, length(bdofs), bdofs))
    stats = Ref{Stats}()
    check(ctx, ccall((:asgfem_solve_primal_host, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Int64, Ref{Stats}),
        ctx.h, sol.entries, b0.entries, atol, rtol, 0, stats))
    @info "PCG on GPU: $(stats[].niter) iterations, solver residual = $(stats[].residual)"
    return bdofs
end

# ---- seam 1b: log-transformed primal problem (SURVEY.md section 8(f), row f1) ------------------------------
"""
Drop-in replacement of `ExtendableASGFEM.solve_logpoisson_primal!` (solvers_logpoisson_primal.jl:130-172): same
arguments, overwrites `sol.entries`, returns `bdofs`.  A + N0 is installed as matrix 0 and the convection matrices N_e as
matrices 1..M of the same fused operator; the preconditioner is factorised from the Laplacian A alone; BiCGStab on the GPU.
"""
function solve_logpoisson_primal!(sol::SGFEVector, A, N0, Nm, b0, G, nmodes, bfac; atol = 1.0e-14, rtol = 1.0e-14, device = 0)
    ctx = Context(device)
    OBT = OrthogonalPolynomialType(sol.TB.ONB)
    set_multiindices!(ctx, OBT, sol.TB.multi_indices)
    cA::SparseMatrixCSC{Float64, Int64} = A.entries.cscmatrix
    D = deepcopy(A)                       # diagonal block A + N0; abs.(A) keeps the entries of A in the union pattern
    D.entries.cscmatrix .+= N0.entries.cscmatrix
    set_matrices!(ctx, D, Nm; extra = (cA,))
    check(ctx, ccall((:asgfem_set_precond_matrix_csc, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
        ctx.h, cA.colptr, cA.rowval, cA.nzval))
    bdofs = boundary_dofs(sol.FES_space[1])
    check(ctx, ccall((:asgfem_set_bdofs, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}), ctx.h, length(bdofs), bdofs))
    b = zeros(Float64, length(sol.entries))          # per-mode load vectors stacked in the layout of sol.entries (:149-152)
    n = div(length(b), nmodes)
    for m in 1:nmodes
        b[((m - 1) * n + 1):(m * n)] .= b0[m][1]
    end
    stats = Ref{Stats}()
    check(ctx, ccall((:asgfem_solve_logprimal_host, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Int64, Ref{Stats}),
        ctx.h, sol.entries, b, atol, rtol, 0, stats))
    @info "BiCGStab on GPU: $(stats[].niter) iterations, preconditioned residual = $(stats[].residual)"
    return bdofs
end

# ---- seam 1c: evaluation of the SGFE solution at a batch of samples (set_sample! of sgfevector.jl:43-69) ---------
"""
`evaluate_samples(ctx, sol, samples)`: column s of the result = entries of `sol.FEV` after `set_sample!(sol, samples[:, s])`.
The univariate basis values come from the package's own `set_sample!(TB, x; normalize = true)`; the sum over the modes
runs on the GPU (the coefficient vector is uploaded once for all samples).
"""
function evaluate_samples(ctx::Context, sol::SGFEVector, samples::AbstractMatrix)
    TB = sol.TB
    M = maxlength_multiindices(TB)
    nvals = length(TB.vals[1])
    S = size(samples, 2)
    vals = zeros(Float64, nvals, M, S)                # nvals fastest = the row-major [s][m][nvals] the library expects
    for s in 1:S
        v = set_sample!(TB, view(samples, :, s); normalize = true)
        for m in 1:M
            vals[:, m, s] .= v[m]
        end
    end
    n = div(length(sol.entries), num_multiindices(sol))
    out = zeros(Float64, n, S)
    check(ctx, ccall((:asgfem_vec_alloc, LIB), Cint, (Ptr{Cvoid}, Cint), ctx.h, 1))
    check(ctx, ccall((:asgfem_vec_upload, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), ctx.h, 0, sol.entries))
    check(ctx, ccall((:asgfem_evaluate_samples, LIB), Cint,
        (Ptr{Cvoid}, Cint, Int64, Int64, Cint, Ptr{Float64}, Ptr{Float64}), ctx.h, 0, S, M, nvals, vals, out))
    return out
end

# ---- seams 2/3: single applications on host vectors (used for parity checks, not for production) ----------
struct GPUSystemPrimal
    ctx::Context
    n::Int
end
Base.size(S::GPUSystemPrimal) = (S.n, S.n)
Base.eltype(::GPUSystemPrimal) = Float64

function LinearAlgebra.mul!(Ax::Vector{Float64}, S::GPUSystemPrimal, x::Vector{Float64})
    check(S.ctx, ccall((:asgfem_apply_host, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), S.ctx.h, x, Ax))
    return Ax
end

struct GPUPreconditionerPrimal
    ctx::Context
end
function LinearAlgebra.ldiv!(y::Vector{Float64}, P::GPUPreconditionerPrimal, b::Vector{Float64})
    check(P.ctx, ccall((:asgfem_precond_apply_host, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), P.ctx.h, b, y))
    return y
end

# ---- seam 4: estimator -------------------------------------------------------------------------------------
"""
Drop-in replacement of `estimate(PoissonProblemPrimal, sol, C; rhs, bonus_quadorder, tail_extension)`.
Returns `(eta4modes, eta4cell, multi_indices_extended, 0.0)` - the 4th value is the `ζ_data` that
scripts/poisson.jl:306 destructures but the reference method does not return (SURVEY.md §3.1).
"""
function estimate(::Type{PoissonProblemPrimal}, sol::SGFEVector, C::StochasticCoefficientCosinus;
        rhs = nothing, bonus_quadorder = 1, tail_extension = [10, 2], device = 0, ctx = nothing)
    FES = sol.FES_space[1]
    xgrid = FES.xgrid
    EG = xgrid[UniqueCellGeometries][1]
    order = get_polynomialorder(eltype(FES), EG)
    ncells = num_cells(xgrid)
    OBT = OrthogonalPolynomialType(sol.TB.ONB)
    if ctx === nothing
        ctx = Context(device)
        set_multiindices!(ctx, OBT, sol.TB.multi_indices)
    end
    coords::Matrix{Float64} = xgrid[Coordinates]
    cellnodes = Matrix{Int32}(xgrid[CellNodes])
    celldofs = Matrix{Int32}(FES[CellDofs][:, :])
    check(ctx, ccall((:asgfem_set_mesh, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Ptr{Int32}),
        ctx.h, size(coords, 2), ncells, coords, cellnodes))
    check(ctx, ccall((:asgfem_set_space, LIB), Cint, (Ptr{Cvoid}, Cint, Int64, Cint, Ptr{Int32}),
        ctx.h, order, FES.ndofs, size(celldofs, 1), celldofs))
    check(ctx, ccall((:asgfem_set_coefficient_cosinus, LIB), Cint,
        (Ptr{Cvoid}, Int64, Float64, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}),
        ctx.h, length(C.decay_factors), C.mean_value, C.decay_factors, Vector{Int64}(C.b1), Vector{Int64}(C.b2)))
    # the pattern only sizes the vectors here; an identity pattern is enough if no matrices were uploaded
    multi_indices_extended = add_boundary_modes(deepcopy(sol.TB.multi_indices); tail_extension = tail_extension)
    mi_ext = mimatrix(multi_indices_extended)
    quadorder = 2 * (order - 1) + bonus_quadorder
    qf = QuadratureRule{Float64, EG}(quadorder)
    qf1 = QuadratureRule{Float64, Edge1D}(quadorder)
    xref = Matrix{Float64}(reduce(hcat, qf.xref))
    sf = Float64[x[1] for x in qf1.xref]
    # rhs at the quadrature points (the closure cannot cross the C ABI): f_at_qp[q, cell]
    rhs === nothing && error("estimate: rhs is required (the reference calls rhs(ftemp, x) unconditionally, estimate.jl:322)")
    f_at_qp = ones(Float64, length(qf.w), ncells)
    if rhs !== nothing
        x = zeros(Float64, 2); tmp = zeros(Float64, 1)
        for cell in 1:ncells, q in 1:length(qf.w)
            n1, n2, n3 = cellnodes[1, cell], cellnodes[2, cell], cellnodes[3, cell]
            x .= coords[:, n1] .+ xref[1, q] .* (coords[:, n2] .- coords[:, n1]) .+ xref[2, q] .* (coords[:, n3] .- coords[:, n1])
            rhs(tmp, x); f_at_qp[q, cell] = tmp[1]
        end
    end
    check(ctx, ccall((:asgfem_vec_alloc, LIB), Cint, (Ptr{Cvoid}, Cint), ctx.h, 1))
    check(ctx, ccall((:asgfem_vec_upload, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), ctx.h, 0, sol.entries))
    eta4cell = zeros(Float64, ncells, size(mi_ext, 2))
    eta4modes = zeros(Float64, size(mi_ext, 2))
    check(ctx, ccall((:asgfem_estimate_poisson_primal, LIB), Cint,
        (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Int64}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint,
            Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        ctx.h, 0, size(mi_ext, 2), size(mi_ext, 1), mi_ext, length(qf.w), xref, Vector{Float64}(qf.w), f_at_qp,
        length(sf), sf, Vector{Float64}(qf1.w), eta4cell, eta4modes))
    return eta4modes, eta4cell, multi_indices_extended, 0.0
end

"""
    assemble_logprimal!(ctx, M, qf) / assemble_logprimal_rhs!(ctx, qf, f_at_qp, ntrunc, slot)

Device assembly of the log-transformed primal problem (src/modelproblems/logpoisson_primal.jl:95-128): matrix 0 = the
Laplacian A, matrices 1..M = N_m = -(grad a_m . grad u, v), and the load vectors b[mu] = (lambda_mu f, v) of all modes into a
device vector slot; mesh, space and the cosinus tables must be set (as for `estimate`).  `solve_logpoisson_primal!` then
needs no matrices from the host: `asgfem_bicgstab(ctx, slot_b, slot_x, ...)`.
"""
assemble_logprimal!(ctx::Context, M::Integer, xref::Matrix{Float64}, w::Vector{Float64}) =
    check(ctx, ccall((:asgfem_assemble_logprimal, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Ptr{Float64}),
        ctx.h, M, length(w), xref, w))
assemble_logprimal_rhs!(ctx::Context, xref::Matrix{Float64}, w::Vector{Float64}, f_at_qp::Matrix{Float64}, ntrunc::Integer, slot::Integer) =
    check(ctx, ccall((:asgfem_assemble_logprimal_rhs, LIB), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Cint), ctx.h, length(w), xref, w, f_at_qp, ntrunc, slot))

"""
    estimate_logpoisson(ctx, sol, C, mi_ext, qf, qf1, f_at_qp; lam_at_qp = nothing)

Drop-in for `estimate(::Type{LogTransformedPoissonProblemPrimal}, sol, C; ...)` (src/estimate.jl:70-257).  `lam_at_qp[j, q, cell]`
= the caller's H1Pk{quadorder} interpolant of <e^-a, H_nu_j> at the cell quadrature points (what :117-131 builds with
`interpolate!`); `nothing` evaluates lambda_nu directly on the device.  Returns (eta4modes, eta4cell, zeta_data).
"""
function estimate_logpoisson(ctx::Context, sol::SGFEVector, C, mi_ext::Matrix{Int64}, xref::Matrix{Float64}, w::Vector{Float64},
        sf::Vector{Float64}, wf::Vector{Float64}, f_at_qp::Matrix{Float64}; lam_at_qp = nothing)
    ncells = size(f_at_qp, 2)
    check(ctx, ccall((:asgfem_vec_alloc, LIB), Cint, (Ptr{Cvoid}, Cint), ctx.h, 1))
    check(ctx, ccall((:asgfem_vec_upload, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), ctx.h, 0, sol.entries))
    eta4cell = zeros(Float64, ncells, size(mi_ext, 2))
    eta4modes = zeros(Float64, size(mi_ext, 2))
    zeta = zeros(Float64, 3)
    check(ctx, ccall((:asgfem_estimate_logpoisson_primal, LIB), Cint,
        (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Int64}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Cint,
            Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        ctx.h, 0, size(mi_ext, 2), size(mi_ext, 1), mi_ext, length(w), xref, w, f_at_qp,
        lam_at_qp === nothing ? C_NULL : lam_at_qp, length(C.decay_factors), length(wf), sf, wf, eta4cell, eta4modes, zeta))
    return eta4modes, eta4cell, zeta[1]
end

"""
    deterministic_sample_solutions(ctx, Samples, b) -> Matrix (ndofs x nsamples)

The deterministic reference solutions of `calculate_sampling_error` (src/sampling_error.jl:112-128) for the affine
coefficient: `ctx` holds K_0..K_M of the SAMPLING space (`FES4sampling`), uploaded with `asgfem_set_stiffness_csc` or
assembled on the device; all samples are solved at once as the columns of one block system.
"""
function deterministic_sample_solutions(ctx::Context, Samples::Matrix{Float64}, b::Vector{Float64}; atol = 1.0e-14, rtol = 1.0e-14)
    check(ctx, ccall((:asgfem_set_samples, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}),
        ctx.h, size(Samples, 2), size(Samples, 1), Samples))
    out = zeros(Float64, length(b), size(Samples, 2))
    stats = Ref{Stats}()
    check(ctx, ccall((:asgfem_solve_samples_host, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Int64, Ref{Stats}), ctx.h, out, b, atol, rtol, 0, stats))
    return out
end

"""
Row-sharded runs (one Julia task / process per GPU, NCCL inside the library): `comm_init!` with the 128-byte id of
`asgfem_comm_unique_id` (created on rank 0, distributed by the host layer), `set_halo!` with the 1-based local row lists,
then `mul!` / `solve_primal!` work on the rank's row shard.  For the estimator the rank's mesh holds its owned cells plus
the neighbouring cell layer; `set_owned_cells!` marks the owned ones and `halo_exchange!` fills the halo rows of the
solution before `estimate` (the mode totals are all-reduced inside the library).
"""
comm_init!(ctx::Context, nranks::Integer, rank::Integer, id::Vector{UInt8}) =
    check(ctx, ccall((:asgfem_comm_init, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), ctx.h, nranks, rank, id))
set_owned_cells!(ctx::Context, owned::Vector{UInt8}) =
    check(ctx, ccall((:asgfem_set_owned_cells, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{UInt8}), ctx.h, length(owned), owned))
halo_exchange!(ctx::Context, slot::Integer = 0) =
    check(ctx, ccall((:asgfem_halo_exchange, LIB), Cint, (Ptr{Cvoid}, Cint), ctx.h, slot))

"""
    ASGFEMCuda.activate!()

Overrides the reference methods so that `solve!(PoissonProblemPrimal, ...)` (src/modelproblems/poisson_primal.jl:75)
and `scripts/poisson.jl` use the GPU path without any other change.
"""
function activate!()
    @eval ExtendableASGFEM begin
        solve_primal!(sol::SGFEVector, A0, Am, b0, G, nmodes, bfac; kwargs...) =
            $(solve_primal!)(sol, A0, Am, b0, G, nmodes, bfac; kwargs...)
        estimate(T::Type{PoissonProblemPrimal}, sol::SGFEVector, C::StochasticCoefficientCosinus; kwargs...) =
            $(estimate)(T, sol, C; kwargs...)
    end
    return nothing
end

end # module
