// Internal definitions shared by the translation units of libasgfem_cuda.so.
// Product code: nothing in here may depend on oracle/.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <array>
#include <new>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/asgfem.h"

namespace asgfem {

// ---- multi-index machinery (index.cpp) ---------------------------------------------------------
struct MultiIndexSet {
    int64_t N = 0, M = 0;
    std::vector<int64_t> mi;     // N x M row-major (mode j = mi[j*M..])
    std::vector<int64_t> plus;   // M x N column-major as in the ABI: plus[m + M*j], 1-based, 0 absent
    std::vector<int64_t> minus;
    void build_neighbours();
    int64_t maxdeg() const;
};
uint64_t hash_mi(const int64_t* v, int64_t M);
void coupling_weights(int family, int64_t maxdeg, std::vector<double>& gp, std::vector<double>& gm);

// neighbour lists per mode in the summation order of mul! (nu ascending, then direction ascending)
struct Coupling {
    std::vector<int32_t> ptr;  // N+1
    std::vector<int32_t> m;    // direction 1..M (0 is the mean term, not stored)
    std::vector<int32_t> nu;   // source mode (0-based)
    std::vector<double> g;
};
void build_coupling(const MultiIndexSet& S, int family, Coupling& C);

// ---- sparse Cholesky of the Dirichlet-reduced K_0 (chol.cpp) -----------------------------------
struct BlockRec {
    int32_t start, len;     // rows [start, start+len) of the elimination order
    int32_t depth;          // depth of the owning node in the dissection tree (same depth => independent)
    int32_t chunk, nchunks;  // long separators are cut into sequentially dependent chunks
};
// std::vector whose resize() leaves new elements uninitialised: the big factor arrays are first touched by the threads
// that fill them instead of being zeroed serially
template <class T>
struct DefaultInitAlloc : std::allocator<T> {
    template <class U>
    struct rebind {
        using other = DefaultInitAlloc<U>;
    };
    template <class U, class... Args>
    void construct(U* p, Args&&... args) {
        if constexpr (sizeof...(Args) == 0)
            ::new ((void*)p) U;
        else
            ::new ((void*)p) U(std::forward<Args>(args)...);
    }
};
template <class T>
using RawVec = std::vector<T, DefaultInitAlloc<T>>;

struct CholFactor {
    int64_t n = 0;                 // reduced dimension
    std::vector<int32_t> perm;     // perm[k] = original (full) row id of the k-th eliminated unknown
    std::vector<int64_t> Lp;       // CSR of L (strictly lower part), rows in elimination order
    RawVec<int32_t> Li;
    RawVec<double> Lx;
    std::vector<double> dinv;      // 1 / L_kk
    std::vector<BlockRec> blocks;  // leaves and separators of the dissection tree, sorted by start
    // nodes of the dissection tree (= blocks with the chunks of one separator joined): columns [node_lo, node_hi) and the
    // sorted rows >= node_hi that hold an entry in one of these columns (a superset: rows added to close the tree are
    // allowed) - lets the sweep-task builder find the rows below a column range without a column copy of L
    std::vector<int32_t> node_lo, node_hi, node_rows;
    std::vector<int64_t> node_rptr;
};
// A: full n_full x n_full CSR (rowptr int64, col int32), keep[i] != 0 for interior rows. Returns 0 or ASGFEM_E*.
int cholesky_reduced(int64_t n_full, const int64_t* rowptr, const int32_t* col, const double* val,
                     const uint8_t* is_boundary, const double* coords_full, int32_t max_block, CholFactor& F,
                     std::string& err);

// sweep tasks of a factor on the host (sptrsv.cu), for tools/chol_bench.cpp: parallel builder or its serial reference
void precond_tasks_host(const CholFactor& F, bool serial_reference, std::vector<unsigned char>& blk_bytes, RawVec<unsigned char>& rec,
                        std::vector<std::array<int, 3>>& launches);

// ---- device-side plans --------------------------------------------------------------------------
struct PrecondPlan;  // sptrsv.cu
struct EstimatePlan;

}  // namespace asgfem

struct asgfem_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;

    // pattern (CSR, 0-based) and the permutation from the caller's CSC order
    int64_t n = 0, nnz = 0;
    int64_t n_owned = -1;  // rows written by apply (multi-GPU); -1 = all
    std::vector<int64_t> h_rowptr;
    std::vector<int32_t> h_col;
    std::vector<int64_t> h_csc_colptr;  // caller's CSC (0-based) for get_pattern / set_stiffness
    bool pattern_from_space = false;    // pattern derived from celldofs by asgfem_assemble_stiffness (dropped by set_mesh / set_space)
    std::vector<int32_t> h_csc_row;
    std::vector<int64_t> h_csc2csr;     // position in CSR of CSC entry p
    int64_t* d_rowptr = nullptr;
    int32_t* d_col = nullptr;
    int32_t M = -1;          // number of KLE matrices beyond the mean (K_0..K_M)
    double* d_vals = nullptr;  // (M+1) x nnz, CSR order
    std::vector<double> h_precond_vals;  // optional SPD matrix for the preconditioner (CSR order); empty: matrix 0
    std::vector<uint8_t> h_bmask;
    uint8_t* d_bmask = nullptr;
    std::vector<int64_t> h_bdofs;

    // stochastic discretisation
    int family = 0;
    asgfem::MultiIndexSet mis;
    asgfem::Coupling coup;
    asgfem::Coupling coup_col;  // the same lists in device column space: "mode" c = column c (ld entries, padding columns empty)
    std::vector<double> gp, gm;
    int32_t* d_cptr = nullptr;
    int32_t* d_cm = nullptr;
    int32_t* d_cnu = nullptr;
    double* d_cg = nullptr;
    int64_t N = 0, ld = 0;  // device vectors are row-major n x ld, ld = number of device columns (multiple of 32)
    // private column order of the device vectors, chosen for the operator (apply_mma.cu): mode mu lives in column
    // h_pos[mu]; h_inv[c] = mode of column c or -1 (padding column, kept at zero)
    std::vector<int32_t> h_pos, h_inv;
    int32_t* d_pos = nullptr;
    int32_t* d_inv = nullptr;

    // vectors
    std::vector<double*> slots;
    double* d_stage = nullptr;  // staging for layout conversion
    size_t stage_bytes = 0;
    double* d_partial = nullptr;  // reduction scratch
    size_t partial_elems = 0;

    // mesh / space / coefficient (device assembly + estimator)
    int64_t nnodes = 0, ncells = 0;
    std::vector<double> h_coords;
    std::vector<int32_t> h_cellnodes;  // 0-based, 3 x ncells
    double* d_coords = nullptr;
    int32_t* d_cellnodes = nullptr;
    int32_t order = 0, ndofs4cell = 0;
    int64_t ndofs_space = 0;
    std::vector<int32_t> h_celldofs;  // 0-based
    std::vector<uint8_t> h_cell_owned;  // row-sharded estimator: 1 = this rank owns the cell (empty: all cells)
    int32_t* d_celldofs = nullptr;
    int64_t maxm = 0;
    double mean = 0;
    std::vector<double> h_decay;
    std::vector<int64_t> h_b1, h_b2;
    double* d_decay = nullptr;
    int32_t* d_b1 = nullptr;
    int32_t* d_b2 = nullptr;

    // kernels
    int apply_variant = 0;
    bool sample_mode = false;  // columns = samples of the random vector (asgfem_set_samples): diagonal coupling, rhs in every column
    double last_apply_ms = 0;
    double last_estimate_ms = 0;
    bool apply_ready = false;  // kernel tables of the operator built for the current pattern / multi-index set
    asgfem::PrecondPlan* precond = nullptr;
    void* mmaplan = nullptr;  // asgfem::MmaPlan (apply_mma.cu)
    void* ts2plan = nullptr;  // asgfem::Ts2Plan (apply_ts2.cu)
    void* distplan = nullptr;  // asgfem::DistPlan (dist.cu): NCCL communicator + halo plan
};

namespace asgfem {

inline int fail(asgfem_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}

// Every exported function is a function-try-block closed by this macro: nothing throws or aborts across the C boundary
// (a failed host allocation becomes ASGFEM_ENOMEM, anything else ASGFEM_EINTERNAL with the exception's message).
inline int boundary_fail(asgfem_ctx* ctx, int code, const char* what) noexcept {
    try {
        if (ctx) ctx->err = what;
    } catch (...) {
    }
    return code;
}
#define ASG_BOUNDARY_CATCH(ctxp)                                                                                       \
    catch (const std::bad_alloc&) { return asgfem::boundary_fail((ctxp), ASGFEM_ENOMEM, "out of host memory"); }       \
    catch (const std::exception& e) { return asgfem::boundary_fail((ctxp), ASGFEM_EINTERNAL, e.what()); }              \
    catch (...) { return asgfem::boundary_fail((ctxp), ASGFEM_EINTERNAL, "unknown C++ exception"); }

#define ASG_CUDA(ctx, call)                                                                             \
    do {                                                                                                \
        cudaError_t _e = (call);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            return asgfem::fail((ctx), _e == cudaErrorMemoryAllocation ? ASGFEM_ENOMEM : ASGFEM_ECUDA,  \
                                std::string(#call) + ": " + cudaGetErrorString(_e));                    \
        }                                                                                               \
    } while (0)

#define ASG_CHECK(ctx, cond, code, msg)                       \
    do {                                                      \
        if (!(cond)) return asgfem::fail((ctx), (code), (msg)); \
    } while (0)

template <class T, class Alloc>
int dev_upload(asgfem_ctx* ctx, T** dptr, const std::vector<T, Alloc>& h) {
    if (*dptr) {
        cudaFree(*dptr);
        *dptr = nullptr;
    }
    size_t bytes = sizeof(T) * (h.empty() ? 1 : h.size());
    ASG_CUDA(ctx, cudaMalloc((void**)dptr, bytes));
    if (!h.empty()) ASG_CUDA(ctx, cudaMemcpyAsync(*dptr, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

// apply.cu
int apply_build_plan(asgfem_ctx* ctx);
void apply_free_plan(asgfem_ctx* ctx);
// rows [r0, r1) of Y (r1 < 0: all rows)
int apply_launch(asgfem_ctx* ctx, const double* x, double* y, int64_t r0 = 0, int64_t r1 = -1);
// apply_mma.cu
int apply_mma_layout(asgfem_ctx* ctx);   // mode-side plan + device column order (set_multiindices)
bool apply_mma_layout_ok(asgfem_ctx* ctx);
void apply_mma_free(asgfem_ctx* ctx);
// apply_blk.cu (kernel tables live inside the MmaPlan)
int apply_blk_build(asgfem_ctx* ctx);
bool apply_blk_usable(asgfem_ctx* ctx);
void apply_blk_free(asgfem_ctx* ctx);
int apply_blk_launch(asgfem_ctx* ctx, const double* x, double* y, int64_t r0, int64_t r1);
// apply_ts2.cu
int apply_ts2_build(asgfem_ctx* ctx);
void apply_ts2_free(asgfem_ctx* ctx);
int apply_ts2_launch(asgfem_ctx* ctx, const double* x, double* y, int64_t r0, int64_t r1);
bool apply_ts2_preferred(asgfem_ctx* ctx);
// dist.cu
int dist_unique_id(void* id128, std::string& err);
int dist_init(asgfem_ctx* ctx, int nranks, int rank, const void* id128);
void dist_free(asgfem_ctx* ctx);
bool dist_active(asgfem_ctx* ctx);
int dist_set_halo(asgfem_ctx* ctx, int32_t nneigh, const int32_t* ranks, const int64_t* send_ptr, const int64_t* send_rows,
                  const int64_t* recv_ptr, const int64_t* recv_rows, int64_t interior0, int64_t interior1);
int dist_halo_exchange(asgfem_ctx* ctx, double* x);                        // no-op without a communicator
int dist_apply(asgfem_ctx* ctx, const double* x, double* y);               // = apply_launch without a communicator
int dist_dot(asgfem_ctx* ctx, const double* a, const double* b, double* out);  // owned rows, summed over the ranks
int dist_max(asgfem_ctx* ctx, double* v);
int dist_allreduce_sum(asgfem_ctx* ctx, double* dbuf, size_t n);  // in place on the device, stream-ordered; no-op without a communicator
int dist_precond_setup_global(asgfem_ctx* ctx, int64_t n_global, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                              int64_t nb, const int64_t* bdofs, const double* coords, const int64_t* row_offsets);
bool dist_has_global_precond(asgfem_ctx* ctx);
int dist_precond_apply(asgfem_ctx* ctx, const double* r, double* z);  // = precond_apply without a global factor
// vecops.cu
int vec_to_device_layout(asgfem_ctx* ctx, const double* host, double* dvec);
int apply_host_pipelined(asgfem_ctx* ctx, const double* x, double* Ax, double* dX, double* dY);
int vec_to_host_layout(asgfem_ctx* ctx, const double* dvec, double* host);
int vec_dot(asgfem_ctx* ctx, const double* a, const double* b, int64_t nrows, double* out);
int vec_fill_random(asgfem_ctx* ctx, double* d, uint64_t seed);
int vec_axpy(asgfem_ctx* ctx, double alpha, const double* x, double* y);
int vec_xpay(asgfem_ctx* ctx, const double* x, double beta, double* y);  // y = x + beta*y
int vec_mask_rows(asgfem_ctx* ctx, double* x);                            // zero boundary rows
int vec_pack_rows(asgfem_ctx* ctx, const double* v, int64_t nrows, const int64_t* d_rows, double* buf);
int vec_unpack_rows(asgfem_ctx* ctx, double* v, int64_t nrows, const int64_t* d_rows, const double* buf);
// out[s*n + i] = sum_k u[i*ld + k] * R[k*Spad + s]   (R: N x Spad on the device, zero padded to a multiple of 8 samples)
int vec_eval_samples(asgfem_ctx* ctx, const double* u, const double* dR, int64_t S, int64_t Spad, double* dout);
// sptrsv.cu
int precond_setup(asgfem_ctx* ctx);
void precond_free(asgfem_ctx* ctx);
int precond_apply(asgfem_ctx* ctx, const double* r, double* z);
int precond_build(asgfem_ctx* ctx, int64_t nfull, const int64_t* rowptr, const int32_t* col, const double* k0,
                  const uint8_t* bmask, const double* xy, PrecondPlan** out);
int precond_apply_plan(asgfem_ctx* ctx, PrecondPlan* P, const double* r, double* z, int64_t nrows, int64_t ld,
                       const uint8_t* d_bmask);
void precond_free_plan(PrecondPlan* P);
void precond_plan_sizes(const PrecondPlan* P, int64_t sizes[5]);
int precond_plan_alloc(asgfem_ctx* ctx, const int64_t sizes[5], PrecondPlan** out);
void precond_plan_buffers(PrecondPlan* P, void* ptrs[3], size_t bytes[3]);
int* precond_plan_launches(PrecondPlan* P);
// pcg.cu
int pcg_solve(asgfem_ctx* ctx, const double* b0_host, double* x, double atol, double rtol, int64_t itmax,
              asgfem_stats* stats);
int bicgstab_solve(asgfem_ctx* ctx, double* b, double* x, double atol, double rtol, int64_t itmax, asgfem_stats* stats);
// assemble.cu
int assemble_stiffness(asgfem_ctx* ctx, int32_t M, int32_t nq, const double* xref, const double* w, int kind = 0);  // kind 1: A, N_1..N_M
int assemble_logprimal_rhs(asgfem_ctx* ctx, int32_t nq, const double* xref, const double* w, const double* f_at_qp, int32_t ntrunc,
                           double* dvec);
// estimate.cu
int estimate_poisson_primal(asgfem_ctx* ctx, const double* u, int64_t N_ext, int64_t M_ext, const int64_t* mi_ext,
                            int32_t nq, const double* xref, const double* w, const double* f_at_qp, int32_t nqf,
                            const double* sf, const double* wf, double* eta4cell, double* eta4modes, int64_t nsel = -1,
                            const int64_t* sel = nullptr, double* cellsum = nullptr, int kind = 0, const double* lam_at_qp = nullptr,
                            int32_t ntrunc = 0, double* zeta3 = nullptr);

}  // namespace asgfem
