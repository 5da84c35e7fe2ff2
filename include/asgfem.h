/* libasgfem_cuda.so - C ABI of the B200-native SGFE solve hot path.
 *
 * Drop-in boundary for ExtendableASGFEM.jl v1.0.1 (pure Julia, no FFI of its own): every entry
 * point below replaces the Julia seam cited next to it (paths relative to the reference repo).
 * The Julia-side binding a maintainer would add (ccall stubs overriding solve_primal!/mul!/ldiv!/
 * estimate) is shown in INTEGRATION.md and shipped as julia/ASGFEMCuda.jl.
 *
 * Conventions
 *  - every function returns 0 on success, a negative ASGFEM_E* code otherwise; nothing throws or
 *    aborts across the boundary; asgfem_last_error(ctx) holds the message of the last failure
 *  - host pointers are borrowed for the duration of the call only; the library owns all device memory
 *  - indices cross the boundary exactly as Julia holds them: 1-based, Int64 (Int32 where the
 *    reference stores Int32, i.e. ExtendableGrid{Float64,Int32} adjacencies)
 *  - vectors cross the boundary in the reference layout: flat length n*N, block mu contiguous
 *    (column-major n x N, src/sgfevector.jl:97-101); the device layout is private
 *  - all calls are synchronous (they return after the stream has drained); a context is not
 *    thread-safe; one context per (GPU, refinement level)
 *  - there is NO CPU fallback: every compute entry point fails with ASGFEM_ECUDA if no device is usable
 */
#ifndef ASGFEM_H
#define ASGFEM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct asgfem_ctx asgfem_ctx;

enum {
    ASGFEM_OK = 0,
    ASGFEM_EINVAL = -1,  /* bad argument / inconsistent sizes / index out of range */
    ASGFEM_ESTATE = -2,  /* call order violated (e.g. apply before set_multiindices)  */
    ASGFEM_ECUDA = -3,   /* CUDA runtime error (message in asgfem_last_error)         */
    ASGFEM_ENOMEM = -4,  /* host or device allocation failed                          */
    ASGFEM_ENUMERIC = -5, /* factorisation broke down (matrix not SPD on interior dofs) */
    ASGFEM_EINTERNAL = -6 /* unexpected C++ exception caught at the boundary (message in asgfem_last_error) */
};

enum { ASGFEM_LEGENDRE = 0, ASGFEM_HERMITE = 1 }; /* src/orthogonal_polynomials/{Legendre_uniform,Hermite_normal}.jl */

/* ---- life cycle ------------------------------------------------------------------------------- */
int asgfem_create(asgfem_ctx** out, int device);
int asgfem_destroy(asgfem_ctx* ctx);
const char* asgfem_last_error(const asgfem_ctx* ctx); /* ctx may be NULL: error of the last failed create */
const char* asgfem_version(void);

/* ---- (a2) recurrence coefficients --------------------------------------------------------------
 * g+(k) = 1/b, g-(k) = c/b of normalise_recurrence_coefficients(OBT, k), k = 0..maxdeg
 * (src/orthogonal_polynomials/orthogonal_polynomials.jl:211-221 as used at src/tensorizedbasis.jl:202-211). */
int asgfem_coupling_weights(int32_t family, int64_t maxdeg, double* gplus, double* gminus);

/* ---- (a3, a4) multi-indices, coupling matrix G, neighbour tables -------------------------------
 * mi: M x N column-major (mode j = mi[j*M .. j*M+M-1]), already padded to equal length
 * (prepare_multi_indices!, src/mopcontrol.jl:29-37).  Replaces get_tensor_multiplication_with_ym
 * (src/tensorizedbasis.jl:193-218) and get_neighbours (src/estimate.jl:1-22). */
int asgfem_set_multiindices(asgfem_ctx* ctx, int32_t family, int64_t N, int64_t M, const int64_t* mi);
int asgfem_get_coupling_nnz(asgfem_ctx* ctx, int64_t* nnz);
/* G as the flushed (M*N) x N CSC of TensorizedBasis.G: colptr[N+1], rowval[nnz], nzval[nnz], 1-based, rows sorted */
int asgfem_get_coupling_csc(asgfem_ctx* ctx, int64_t* colptr, int64_t* rowval, double* nzval);
/* PLUS/MINUS as M x N column-major Int64, 1-based mode ids, 0 = absent */
int asgfem_get_neighbours(asgfem_ctx* ctx, int64_t* plus, int64_t* minus);

/* ---- (a10, a13) host index machinery (no context, integer only) -------------------------------
 * add_boundary_modes (src/mopcontrol.jl:60-132) incl. its quirks, without the sleep(1) of :74.
 * Call with out == NULL to query N_ext / M_ext; out is M_ext x N_ext column-major. */
int asgfem_add_boundary_modes(int64_t N, int64_t M, const int64_t* mi, int64_t p_extension, int64_t tail1,
                              int64_t tail2, int64_t* N_ext, int64_t* M_ext, int64_t* out, int64_t out_capacity);
/* classify_modes (src/mopcontrol.jl:168-241): cls[j] = 0 inactive_else, 1 inactive_bnd, 2 inactive_bnd2,
 * 3 active_bnd, 4 active_int; the first N_active modes of mi_ext are the active set (scripts/poisson.jl:341) */
int asgfem_classify_modes(int64_t N_ext, int64_t M, const int64_t* mi_ext, int64_t N_active, int32_t* cls);

/* ---- (a6) stiffness matrices K_0..K_M in one shared pattern ------------------------------------
 * Host-assembled path (parity with the Julia FEMatrix objects of src/modelproblems/poisson_primal.jl:56-63):
 * pattern = Julia CSC (colptr[n+1], rowval[nnz]) 1-based, structurally symmetric; nzval on that pattern.
 * set_stiffness_csc accepts a matrix with its own (sub-)pattern, e.g. when ExtendableSparse dropped exact zeros. */
int asgfem_set_pattern_csc(asgfem_ctx* ctx, int64_t n, const int64_t* colptr, const int64_t* rowval);
int asgfem_set_num_stiffness(asgfem_ctx* ctx, int32_t M); /* allocates K_0..K_M (zero) */
int asgfem_set_stiffness(asgfem_ctx* ctx, int32_t m, const double* nzval);
int asgfem_set_stiffness_csc(asgfem_ctx* ctx, int32_t m, const int64_t* colptr, const int64_t* rowval,
                             const double* nzval);
int asgfem_get_stiffness(asgfem_ctx* ctx, int32_t m, double* nzval); /* on the shared pattern (CSC order) */
int asgfem_get_pattern_nnz(asgfem_ctx* ctx, int64_t* nnz);
int asgfem_get_pattern_csc(asgfem_ctx* ctx, int64_t* colptr, int64_t* rowval);
/* boundary dofs (1-based) as returned by solve_primal! (src/modelproblems/solvers_poisson_primal.jl:136-142) */
int asgfem_set_bdofs(asgfem_ctx* ctx, int64_t nb, const int64_t* bdofs);

/* ---- (a5, a6) device assembly from the KLE modes -----------------------------------------------
 * Mesh and space as ExtendableGrids/ExtendableFEMBase hold them: coords 2 x nnodes (Float64),
 * cellnodes 3 x ncells (Int32, 1-based), celldofs ndofs4cell x ncells (Int32, 1-based) = FES[CellDofs].
 * order = 1 (H1Pk{1,2,1}) or 2 (H1Pk{1,2,2}). */
int asgfem_set_mesh(asgfem_ctx* ctx, int64_t nnodes, int64_t ncells, const double* coords, const int32_t* cellnodes);
int asgfem_set_space(asgfem_ctx* ctx, int32_t order, int64_t ndofs, int32_t ndofs4cell, const int32_t* celldofs);
/* StochasticCoefficientCosinus fields (src/coefficients/cosinus.jl:11-17): a_0 = mean,
 * a_m(x) = decay_factors[m] cos(pi b1[m] x1) cos(pi b2[m] x2) (get_am!, cosinus.jl:58-65) */
int asgfem_set_coefficient_cosinus(asgfem_ctx* ctx, int64_t maxm, double mean, const double* decay_factors,
                                   const int64_t* b1, const int64_t* b2);
/* K_m[i,j] = sum_T |T| sum_q w[q] a_m(x_q) grad(phi_j).grad(phi_i), m = 0..M, with the caller's quadrature
 * rule (xref 2 x nq reference coordinates, w sums to 1) - the arithmetic of
 * BilinearOperator(get_am_x(m,C),[grad(1)],[grad(1)]) (src/coefficients/coefficients.jl:147-153).
 * Builds the shared pattern from celldofs if none was set. */
int asgfem_assemble_stiffness(asgfem_ctx* ctx, int32_t M, int32_t nq, const double* xref, const double* w);

/* ---- (f1) assembly of the log-transformed primal problem (src/modelproblems/logpoisson_primal.jl:95-128) -------------------
 * asgfem_assemble_logprimal: matrix 0 = the Laplacian A = (grad u, grad v) (the reference's N0 is an empty matrix), matrices
 * 1..M = N_m = -(grad a_m . grad u, v) with grad a_m from get_gradam! (cosinus.jl:67-76); one quadrature rule for all of
 * them (order 2*order - 1 + bonus_quadorder_a).  The operator, the preconditioner (factorised from matrix 0) and
 * asgfem_bicgstab then solve the system of solve_logpoisson_primal!.
 * asgfem_assemble_logprimal_rhs: the load vectors b[mu] = (lambda_mu f, phi_i) of ALL modes into a device slot, lambda_mu =
 * the PCE coefficient of exp(-a) from expa_PCE_mop (src/coefficients/coefficients.jl:236-262, factor = -1, N_truncate =
 * ntrunc, normally maxm); f_at_qp = rhs at the quadrature points (nq x ncells column-major), rule of order
 * order + bonus_quadorder_f. */
int asgfem_assemble_logprimal(asgfem_ctx* ctx, int32_t M, int32_t nq, const double* xref, const double* w);
int asgfem_assemble_logprimal_rhs(asgfem_ctx* ctx, int32_t nq, const double* xref, const double* w, const double* f_at_qp,
                                  int32_t ntrunc, int32_t slot_b);

/* ---- (a1) SGFEVector storage --------------------------------------------------------------------
 * Device-resident n x N fp64 blocks addressed by slot id (src/sgfevector.jl:18-27, entries 86-106).
 * Slots 0..nslots-1 are user slots; the PCG driver allocates its own work vectors.  asgfem_vec_alloc sets the slot COUNT:
 * slots with index >= nslots are freed, new ones are zero-filled.  The host-vector entry points (asgfem_apply_host,
 * asgfem_precond_apply_host, asgfem_solve_*_host) use slots 0 and 1 as staging and grow the table if needed: keep device
 * vectors that must survive such a call in slots >= 2. */
int asgfem_vec_alloc(asgfem_ctx* ctx, int32_t nslots);
int asgfem_vec_upload(asgfem_ctx* ctx, int32_t slot, const double* host);   /* host: n*N, reference layout */
int asgfem_vec_download(asgfem_ctx* ctx, int32_t slot, double* host);
int asgfem_vec_zero(asgfem_ctx* ctx, int32_t slot);
/* x[i + n*mu] = 2*u01(splitmix64(seed ^ (i + n*mu))) - 1, generated on the device (bench inputs) */
int asgfem_vec_fill_random(asgfem_ctx* ctx, int32_t slot, uint64_t seed);
int asgfem_vec_dot(asgfem_ctx* ctx, int32_t slot_a, int32_t slot_b, double* out);
int asgfem_vec_axpy(asgfem_ctx* ctx, double alpha, int32_t slot_x, int32_t slot_y); /* y += alpha x */
int asgfem_vec_xpay(asgfem_ctx* ctx, int32_t slot_x, double beta, int32_t slot_y); /* y = x + beta y */
int asgfem_vec_copy(asgfem_ctx* ctx, int32_t slot_src, int32_t slot_dst);

/* ---- (a7) operator ------------------------------------------------------------------------------
 * Y = sum_m (G_m (x) K_m) X with the rows of bdofs zeroed: LinearAlgebra.mul!(Ax, S::MySystemPrimal, x)
 * (src/modelproblems/solvers_poisson_primal.jl:86-124).  Columns at bdofs take part exactly as in the
 * reference (X is not masked). */
int asgfem_apply(asgfem_ctx* ctx, int32_t slot_x, int32_t slot_y);
int asgfem_apply_host(asgfem_ctx* ctx, const double* x, double* Ax); /* host vectors, reference layout */
/* selects the kernel: 0 = automatic, 1 = reference-order gather kernel, 2 = row-block tiled kernel,
 * 3 = row-resident dst-major kernel, 4 = row-resident direction-major kernel (shared-memory atomics, run-to-run
 * rounding-level differences), 5 = as 4 with warp-owned target ranges (deterministic), 6 = mode-stationary kernel
 * (operands in registers, partial products exchanged through shared memory; deterministic) */
int asgfem_set_apply_variant(asgfem_ctx* ctx, int32_t variant);
/* duration of the last asgfem_apply kernel(s) in milliseconds, from CUDA events on the library's stream */
int asgfem_last_apply_ms(asgfem_ctx* ctx, double* ms);
/* Device time (CUDA events) of the kernels of the last asgfem_estimate_poisson_primal call: cell residuals, face jumps,
 * column sums; excludes the upload of the tables and the download of eta4cell.  Measurement hook, no reference seam. */
int asgfem_last_estimate_ms(asgfem_ctx* ctx, double* ms);

/* ---- (a8) mean-based preconditioner -------------------------------------------------------------
 * setup: MyPreconditionerPrimal (solvers_poisson_primal.jl:30-44): K_0 with the boundary dofs pinned,
 * factorised once on the host (sparse Cholesky, nested dissection), uploaded as level-scheduled
 * triangular factors.  apply: ldiv!(y, P, b) (:46-78) for all N blocks at once; y may alias b.
 * Boundary rows of the result are exactly 0 (the reference leaves O(1e-60) there). */
int asgfem_precond_setup(asgfem_ctx* ctx);
int asgfem_precond_apply(asgfem_ctx* ctx, int32_t slot_r, int32_t slot_z);
int asgfem_precond_apply_host(asgfem_ctx* ctx, const double* b, double* y);
/* Host-only diagnostic of the factorisation behind asgfem_precond_setup (no context, no GPU): K (n x n CSR, 0-based,
 * symmetric, both triangles stored) is reduced by the rows/columns with is_boundary != 0, factorised exactly as the setup
 * does (nested dissection steered by coords[2 n] if not NULL, multifrontal Cholesky on the host cores) and
 * K_red x = b is solved for one vector; boundary rows of x are 0 like the preconditioner's.  *lnz = nonzeros of the strict
 * lower triangle of the factor.  Returns 0 or ASGFEM_E*; err (may be NULL) receives the message. */
int asgfem_host_factor_solve(int64_t n, const int64_t* rowptr, const int32_t* col, const double* val, const uint8_t* is_boundary,
                             const double* coords, const double* b, double* x, int64_t* lnz, char* err, int32_t errlen);

/* ---- (a9) Krylov driver -------------------------------------------------------------------------
 * solve_primal! (solvers_poisson_primal.jl:130-169) with PCG in place of Krylov.gmres:
 * rhs b[:,1] = x0[:,1] + b0 ... exactly lines 149-155 (b = deepcopy(sol); b[1] += b0; b[m][bdofs] = 0),
 * warm start from the slot content, stop when sqrt(r.z) <= atol + rtol*sqrt(r0.z0). */
typedef struct asgfem_stats {
    int64_t niter;
    int32_t solved;
    int32_t _pad;
    double rz0;          /* sqrt(r0.z0) */
    double rzk;          /* sqrt(rk.zk) at exit */
    double residual;     /* ||A x - b||_2 after the solve (the reference's "solver residual", :165-167) */
    double ms_setup;     /* rhs + initial residual */
    double ms_iterations;
    double ms_apply;     /* accumulated operator time   */
    double ms_precond;   /* accumulated preconditioner time */
} asgfem_stats;
int asgfem_pcg(asgfem_ctx* ctx, const double* b0, int32_t slot_x, double atol, double rtol, int64_t itmax,
               asgfem_stats* stats);
/* ---- log-transformed primal problem ------------------------------------------------
 * solve_logpoisson_primal!(sol, A, N0, Nm, b0, G, nmodes, bfac; atol, rtol)
 *     (src/modelproblems/solvers_logpoisson_primal.jl:130-172; operator MySystemLogPrimal.mul! :87-127,
 *      preconditioner MyPreconditionerLogPrimal = I (x) LU(A) :31-83)
 * The operator has the tensor structure of the primal one: the caller installs A + N0 as matrix 0 and the convection
 * matrices N_e as matrices 1..M (asgfem_set_stiffness*), nothing in the operator kernels assumes symmetric values.
 * The preconditioner is built from the SPD Laplacian A alone, handed over here (CSC on a sub-pattern of the shared
 * pattern, 1-based; colptr == NULL returns to the default: matrix 0).  Invalidates an existing factorisation. */
int asgfem_set_precond_matrix_csc(asgfem_ctx* ctx, const int64_t* colptr, const int64_t* rowval, const double* nzval);
/* Nonsymmetric Krylov solve on the device: BiCGStab on the left-preconditioned system P^-1 S x = P^-1 b, stopping on
 * ||P^-1 r_k|| <= atol + rtol ||P^-1 r_0|| (the quantity Krylov.gmres(...; ldiv = true, M = P) monitors, :163).  The
 * reference uses GMRES; both converge to the solution of the same nonsingular system (checked against the direct solve
 * of the assembled block system, solve_logpoisson_primal_full! :175-230).  slot_b: right-hand side (boundary rows are
 * zeroed here), slot_x: warm start / solution.  stats: rz0 / rzk = preconditioned residual norms. */
int asgfem_bicgstab(asgfem_ctx* ctx, int32_t slot_b, int32_t slot_x, double atol, double rtol, int64_t itmax,
                    asgfem_stats* stats);
/* whole seam on host vectors: sol (n*N, in: warm start, out: solution), b (n*N: the per-mode load vectors b0[m] stacked
 * in the reference layout); the right-hand side is deepcopy(sol) + b with the boundary rows zeroed (:149-156). */
int asgfem_solve_logprimal_host(asgfem_ctx* ctx, double* sol, const double* b, double atol, double rtol, int64_t itmax,
                                asgfem_stats* stats);

/* ---- evaluation at samples (the set_sample! half) ------------------------------------
 * set_sample!(SGFEV::SGFEVector, S) (src/sgfevector.jl:43-69) for a batch of samples:
 *     out[:, s] = sum_k H_k(xi_s) u_k,   H_k(xi) = prod_m vals[s][m][mu_k[m]]   (evaluate(TB, k), tensorizedbasis.jl:244-252)
 * vals holds TB.vals after set_sample!(TB, xi_s; normalize = true) (tensorizedbasis.jl:226-236) for every sample: nsamples
 * blocks of M (= length of the multi-indices) rows with nvals = maxorder + 1 entries each (row-major; the rows beyond the length of the sample are
 * [1, 0, 0, ...] as the reference sets them).  The univariate polynomial values are the caller's data, like the
 * quadrature tables.  slot_u: coefficient vector on the device; out: host, n x nsamples column-major. */
int asgfem_evaluate_samples(asgfem_ctx* ctx, int32_t slot_u, int64_t nsamples, int64_t M, int32_t nvals,
                            const double* vals, double* out);

/* whole seam on host vectors: sol (n*N, in: warm start, out: solution), b0 (n) */
int asgfem_solve_primal_host(asgfem_ctx* ctx, double* sol, const double* b0, double atol, double rtol,
                             int64_t itmax, asgfem_stats* stats);

/* ---- (a11, a12) residual error estimator -------------------------------------------------------
 * estimate(::Type{PoissonProblemPrimal}, sol, C; rhs, bonus_quadorder, tail_extension) (src/estimate.jl:260-418)
 * for the solution in slot_u.  Needs set_mesh/set_space/set_coefficient_cosinus and the ACTIVE multi-indices
 * (set_multiindices).  mi_ext: M_ext x N_ext column-major extended set (active modes first,
 * asgfem_add_boundary_modes).  Cell rule (xref 2 x nq, w) and face rule (sf nqf points on [0,1], wf) are
 * the caller's QuadratureRule tables of order 2(order-1)+bonus_quadorder (:286-287); f_at_qp is the rhs at the
 * cell quadrature points (nq x ncells column-major).  NULL means f == 1 - a convenience of THIS interface; the
 * reference calls rhs(ftemp, x) unconditionally (:322), so the Python / Julia wrappers refuse a missing rhs.  Outputs: eta4cell ncells x N_ext
 * column-major, eta4modes N_ext (the Julia return values of :417). */
int asgfem_estimate_poisson_primal(asgfem_ctx* ctx, int32_t slot_u, int64_t N_ext, int64_t M_ext,
                                   const int64_t* mi_ext, int32_t nq, const double* xref, const double* w,
                                   const double* f_at_qp, int32_t nqf, const double* sf, const double* wf,
                                   double* eta4cell, double* eta4modes);

/* (f2) estimate(::Type{LogTransformedPoissonProblemPrimal}, sol, C; rhs, bonus_quadorder, tail_extension) - src/estimate.jl:70-257.
 * Same tables as above.  lambda_nu = <e^-a, H_nu> (expa_PCE_mop, factor -1, N_truncate = ntrunc): the reference interpolates
 * it into H1Pk{quadorder} and evaluates the interpolant at the quadrature points; lam_at_qp (N_ext x nq x ncells column-major,
 * the caller's interpolated values) reproduces that, NULL evaluates lambda_nu directly at the quadrature points (the
 * commented-out line :184 of the reference).  f_at_qp is required.  zeta3 (3 doubles, may be NULL) = zeta_data,
 * zeta_data1, zeta_data2 (:156-175, :248).  eta4modes of the active modes reproduces the "+=" of :244. */
int asgfem_estimate_logpoisson_primal(asgfem_ctx* ctx, int32_t slot_u, int64_t N_ext, int64_t M_ext, const int64_t* mi_ext,
                                      int32_t nq, const double* xref, const double* w, const double* f_at_qp,
                                      const double* lam_at_qp, int32_t ntrunc, int32_t nqf, const double* sf, const double* wf,
                                      double* eta4cell, double* eta4modes, double* zeta3);

/* The same estimator with the outputs the adaptive loop consumes (scripts/poisson.jl:341-420): eta4modes and
 * cellsum[c] = sum_k eta4cell[c, sel[k]] (sel: 1-based columns, e.g. the active modes - the indicator handed to bulk_mark
 * at :402) instead of the ncells x N_ext matrix, whose transfer to the host dominates the call above. */
int asgfem_estimate_poisson_primal_marking(asgfem_ctx* ctx, int32_t slot_u, int64_t N_ext, int64_t M_ext,
                                           const int64_t* mi_ext, int32_t nq, const double* xref, const double* w,
                                           const double* f_at_qp, int32_t nqf, const double* sf, const double* wf,
                                           int64_t nsel, const int64_t* sel, double* cellsum, double* eta4modes);

/* ---- (e) row-sharded multi-GPU operation --------------------------------------------------------
 * One context per rank/GPU.  The context holds the LOCAL rows of every K_m with local column ids:
 * columns 0..n_owned-1 are the owned dofs, n_owned..n_local-1 the halo dofs (owned by neighbours).
 * Vectors have n_local rows; only owned rows are written by apply.  The host layer (torch.distributed /
 * NCCL) exchanges halo rows between the pack/unpack calls and all-reduces the partial dots. */
int asgfem_set_owned_rows(asgfem_ctx* ctx, int64_t n_owned);
/* ---- deterministic reference solutions at samples (src/sampling_error.jl:84-128: ExtendableFEM.solve per sample on host
 * threads).  For the affine coefficient the matrix of sample s is K_0 + sum_m xi[m, s] K_m; with the samples as the columns
 * of the device vectors all nsamples problems are one block system with a diagonal coupling, solved by the same operator
 * kernel / multi-RHS mean preconditioner / PCG as the SGFE system.  samples: Msamples x nsamples, column-major (Julia's
 * Samples[:, s]); Msamples <= number of matrices K_m.  Replaces the multi-index set of the context (use a context of its
 * own for the sampling space).  out: n x nsamples column-major. */
int asgfem_set_samples(asgfem_ctx* ctx, int64_t nsamples, int64_t Msamples, const double* samples);
int asgfem_solve_samples_host(asgfem_ctx* ctx, double* out, const double* b, double atol, double rtol, int64_t itmax,
                              asgfem_stats* stats);

/* Row-sharded estimator (cells and faces sharded with the rows, totals all-reduced).  The mesh of a
 * rank = its owned cells plus the layer of neighbouring cells whose dofs are its halo rows; owned[c] != 0 marks the cells it
 * owns (NULL: all cells again).  asgfem_estimate_poisson_primal then sums eta4modes over the owned cells, counts an interior
 * face with the share of its two cells that the rank owns (the neighbour adds the rest) and all-reduces the sums over the
 * communicator of asgfem_comm_init; eta4cell is meaningful in the rows of the owned cells.  The solution vector must hold
 * current halo rows (asgfem_halo_exchange). */
int asgfem_set_owned_cells(asgfem_ctx* ctx, int64_t ncells, const uint8_t* owned);
/* fills the halo rows of a vector slot from the neighbours' owned rows (the exchange asgfem_apply performs internally) */
int asgfem_halo_exchange(asgfem_ctx* ctx, int32_t slot);
/* Operator on the local rows [row0, row1) only (0-based, half open, clipped to the owned rows); the other rows of slot sy
 * are left untouched.  Lets the host layer apply the rows that reference no halo column while the halo exchange is in
 * flight, and the rows along the partition boundary afterwards.  asgfem_last_apply_ms reports this launch. */
int asgfem_apply_rows(asgfem_ctx* ctx, int32_t sx, int32_t sy, int64_t row0, int64_t row1);
/* device pointer to the private (row-major n_local x ldN) storage of a slot, and its leading dimension */
int asgfem_vec_device_ptr(asgfem_ctx* ctx, int32_t slot, void** dptr, int64_t* ld);
/* gather rows[0..nrows) (1-based local ids) of a slot into a dense device buffer (nrows x N) / scatter back */
int asgfem_pack_rows(asgfem_ctx* ctx, int32_t slot, int64_t nrows, const int64_t* rows, void* dbuf);
int asgfem_unpack_rows(asgfem_ctx* ctx, int32_t slot, int64_t nrows, const int64_t* rows, const void* dbuf);
int asgfem_vec_dot_owned(asgfem_ctx* ctx, int32_t slot_a, int32_t slot_b, double* out);


/* ---- (e) NCCL inside the library -------------------------------------------------------------------
 * One process per GPU, one context per process.  Rank 0 creates a 128-byte NCCL id (asgfem_comm_unique_id) and the host
 * layer hands it to the other ranks by any channel it has (MPI.jl bcast, a file, torch.distributed); every rank then
 * calls asgfem_comm_init.  NCCL is bound at run time (libnccl.so.2), the single-GPU library does not depend on it.
 * With a communicator and a halo plan installed, asgfem_apply / asgfem_pcg / asgfem_solve_primal_host work on the ROW
 * SHARD of this rank: asgfem_apply packs the send rows, posts grouped ncclSend/ncclRecv on a communication stream,
 * applies the operator to the owned rows [interior_row0, interior_row1) (which reference no halo column) while the
 * exchange is in flight, unpacks the halo rows and applies the remaining owned rows - one launch sequence without host
 * synchronisation; the Krylov inner products are all-reduced (ncclAllReduce).  The mean preconditioner is the rank-local
 * one (block-Jacobi over the partition, SURVEY.md section 7: same converged solution, more iterations).
 * set_halo: neighbour k has rank ranks[k], send rows send_rows[send_ptr[k] .. send_ptr[k+1]) (owned, 1-based local ids)
 * and receive rows recv_rows[recv_ptr[k] .. recv_ptr[k+1]) (halo rows, 1-based local ids); both sides of an exchange list
 * the rows in the same (global) order. */
int asgfem_comm_unique_id(void* id128);
int asgfem_comm_init(asgfem_ctx* ctx, int32_t nranks, int32_t rank, const void* id128);
int asgfem_comm_destroy(asgfem_ctx* ctx);
int asgfem_set_halo(asgfem_ctx* ctx, int32_t nneigh, const int32_t* ranks, const int64_t* send_ptr, const int64_t* send_rows,
                    const int64_t* recv_ptr, const int64_t* recv_rows, int64_t interior_row0, int64_t interior_row1);
/* EXACT mean preconditioner for the sharded solve: MyPreconditionerPrimal (solvers_poisson_primal.jl:30-78) applies
 * K_0^-1 of the GLOBAL mesh to every mode block, which no rank can do on its row shard.  The modes are independent, so
 * the library swaps from row shards to MODE shards for the preconditioner: every rank receives all rows of its share of
 * the device columns (grouped ncclSend/ncclRecv, all-to-all), sweeps them with the factor of the global K_0 and sends the
 * rows back - the iteration counts of the single-GPU solve are kept at any number of ranks.
 * The host layer hands the global K_0 (CSC, 1-based, GLOBAL numbering = the owned rows of rank 0, then rank 1, ... in
 * their local order) and the global Dirichlet dofs to every rank; row_offsets[nranks+1] are the first global rows of the
 * ranks; coords (2 x n_global, may be NULL) steer the nested dissection.  COLLECTIVE: rank 0 factorises on its host cores and
 * broadcasts the sweep tasks over NCCL, so colptr / rowval / nzval / coords are read on rank 0 only (may be NULL elsewhere). */
int asgfem_precond_setup_global(asgfem_ctx* ctx, int64_t n_global, const int64_t* colptr, const int64_t* rowval,
                                const double* nzval, int64_t nb, const int64_t* bdofs, const double* coords,
                                const int64_t* row_offsets);
/* inner product over the owned rows, summed over all ranks (= asgfem_vec_dot_owned without a communicator) */
int asgfem_vec_dot_global(asgfem_ctx* ctx, int32_t slot_a, int32_t slot_b, double* out);

#ifdef __cplusplus
}
#endif
#endif /* ASGFEM_H */
