// C ABI of libasgfem_cuda.so (include/asgfem.h).  Argument validation, index conversion (1-based Julia
// CSC -> 0-based CSR), slot management; all arithmetic lives in the kernels of the other translation units.
#include <algorithm>
#include <cstring>

#include "common.h"

using namespace asgfem;

static std::string g_create_error;

#define CTX_OR_FAIL(ctx) \
    if (!(ctx)) return ASGFEM_EINVAL

static int set_device(asgfem_ctx* ctx) {
    ASG_CUDA(ctx, cudaSetDevice(ctx->device));
    return 0;
}

extern "C" const char* asgfem_version(void) { return "asgfem-b200 0.1.0 (sm_100a)"; }

extern "C" const char* asgfem_last_error(const asgfem_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int asgfem_create(asgfem_ctx** out, int device) try {
    if (!out) return ASGFEM_EINVAL;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("no usable CUDA device (libasgfem_cuda has no CPU fallback): ") +
                         (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return ASGFEM_ECUDA;
    }
    if (device < 0 || device >= count) {
        g_create_error = "device index out of range";
        return ASGFEM_EINVAL;
    }
    asgfem_ctx* ctx = new asgfem_ctx();
    ctx->device = device;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreate(&ctx->stream)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->ev0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) {
        g_create_error = cudaGetErrorString(e);
        delete ctx;
        return ASGFEM_ECUDA;
    }
    *out = ctx;
    return 0;
}
ASG_BOUNDARY_CATCH(nullptr)

static void free_vec_storage(asgfem_ctx* ctx) {
    for (double* p : ctx->slots)
        if (p) cudaFree(p);
    ctx->slots.clear();
}

extern "C" int asgfem_destroy(asgfem_ctx* ctx) try {
    CTX_OR_FAIL(ctx);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    apply_free_plan(ctx);
    apply_mma_free(ctx);
    precond_free(ctx);
    dist_free(ctx);
    free_vec_storage(ctx);
    void* ptrs[] = {ctx->d_rowptr, ctx->d_col,  ctx->d_vals,      ctx->d_bmask,    ctx->d_cptr,  ctx->d_cm,
                    ctx->d_cnu,    ctx->d_cg,   ctx->d_stage,     ctx->d_partial,  ctx->d_coords, ctx->d_cellnodes,
                    ctx->d_celldofs, ctx->d_decay, ctx->d_b1,     ctx->d_b2,       ctx->d_pos,   ctx->d_inv};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

// ---- samples as columns (deterministic reference solutions of the MC error, src/sampling_error.jl:84-128) -----------
// For the affine coefficient a(x, xi) = a_0(x) + sum_m xi_m a_m(x) the deterministic problem of sample s has the matrix
// K(xi_s) = K_0 + sum_m xi_{s,m} K_m on the pattern shared by all K_m.  With the samples as the columns of the vectors,
// all nsamples systems are ONE block system with a diagonal coupling (G_m = diag(xi_{.,m})): the operator kernel, the
// multi-RHS mean preconditioner K_0^-1 and the PCG of the SGFE solve apply unchanged (the reference solves the samples
// one by one on host threads with ExtendableFEM.solve).
extern "C" int asgfem_set_samples(asgfem_ctx* ctx, int64_t nsamples, int64_t Msamples, const double* samples) try {
    CTX_OR_FAIL(ctx);
    ASG_CHECK(ctx, nsamples >= 1 && nsamples < 65536 && Msamples >= 0 && (samples || Msamples == 0), ASGFEM_EINVAL, "set_samples: bad arguments");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    free_vec_storage(ctx);
    apply_free_plan(ctx);
    apply_mma_free(ctx);
    ctx->sample_mode = true;
    ctx->mis = asgfem::MultiIndexSet();
    ctx->coup = asgfem::Coupling();
    ctx->N = nsamples;
    ctx->ld = (nsamples + 31) / 32 * 32;
    ctx->h_pos.resize((size_t)nsamples);
    ctx->h_inv.assign((size_t)ctx->ld, -1);
    for (int64_t k = 0; k < nsamples; ++k) ctx->h_pos[(size_t)k] = ctx->h_inv[(size_t)k] = (int32_t)k;
    Coupling& CC = ctx->coup_col;
    CC.ptr.assign((size_t)ctx->ld + 1, 0);
    CC.m.clear();
    CC.nu.clear();
    CC.g.clear();
    for (int64_t c = 0; c < ctx->ld; ++c) {
        if (c < nsamples)
            for (int64_t m = 0; m < Msamples; ++m) {  // directions beyond the uploaded matrices are rejected at the first apply
                CC.m.push_back((int32_t)(m + 1));
                CC.nu.push_back((int32_t)c);
                CC.g.push_back(samples[m + Msamples * c]);
            }
        CC.ptr[(size_t)c + 1] = (int32_t)CC.m.size();
    }
    ctx->coup = CC;
    ctx->coup.ptr.resize((size_t)nsamples + 1);
    int rc = 0;
    rc |= dev_upload(ctx, &ctx->d_cptr, CC.ptr);
    rc |= dev_upload(ctx, &ctx->d_cm, CC.m);
    rc |= dev_upload(ctx, &ctx->d_cnu, CC.nu);
    rc |= dev_upload(ctx, &ctx->d_cg, CC.g);
    rc |= dev_upload(ctx, &ctx->d_pos, ctx->h_pos);
    rc |= dev_upload(ctx, &ctx->d_inv, ctx->h_inv);
    if (rc) return rc;
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

// ---- multi-indices ----------------------------------------------------------------------------
extern "C" int asgfem_set_multiindices(asgfem_ctx* ctx, int32_t family, int64_t N, int64_t M, const int64_t* mi) try {
    CTX_OR_FAIL(ctx);
    ASG_CHECK(ctx, family == ASGFEM_LEGENDRE || family == ASGFEM_HERMITE, ASGFEM_EINVAL, "unknown polynomial family");
    ASG_CHECK(ctx, N >= 1 && M >= 1 && mi, ASGFEM_EINVAL, "set_multiindices: need N >= 1, M >= 1");
    ASG_CHECK(ctx, N < 65536, ASGFEM_EINVAL, "set_multiindices: at most 65535 modes are supported");
    for (int64_t k = 0; k < N * M; ++k) ASG_CHECK(ctx, mi[k] >= 0, ASGFEM_EINVAL, "negative multi-index entry");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    free_vec_storage(ctx);  // vectors are sized by the column count, and the column order changes with the set
    ctx->sample_mode = false;
    ctx->family = family;
    ctx->mis.N = N;
    ctx->mis.M = M;
    ctx->mis.mi.assign(mi, mi + N * M);
    ctx->mis.build_neighbours();
    build_coupling(ctx->mis, family, ctx->coup);
    coupling_weights(family, ctx->mis.maxdeg() + 1, ctx->gp, ctx->gm);
    ctx->N = N;
    apply_free_plan(ctx);
    // device column order: chosen by the operator's mode-side plan; identity if the plan declines the set
    int rc = apply_mma_layout(ctx);
    if (rc) return rc;
    if (!apply_mma_layout_ok(ctx)) {
        ctx->ld = (N + 31) / 32 * 32;
        ctx->h_pos.resize((size_t)N);
        ctx->h_inv.assign((size_t)ctx->ld, -1);
        for (int64_t k = 0; k < N; ++k) ctx->h_pos[(size_t)k] = ctx->h_inv[(size_t)k] = (int32_t)k;
    } else {
        ctx->ld = (int64_t)ctx->h_inv.size();
    }
    // coupling lists in column space (reference order of the entries of a mode kept)
    const Coupling& C = ctx->coup;
    Coupling& CC = ctx->coup_col;
    CC.ptr.assign((size_t)ctx->ld + 1, 0);
    CC.m.clear();
    CC.nu.clear();
    CC.g.clear();
    std::vector<int32_t>&cptr = CC.ptr, &cm = CC.m, &cnu = CC.nu;
    std::vector<double>& cg = CC.g;
    for (int64_t c = 0; c < ctx->ld; ++c) {
        const int32_t mode = ctx->h_inv[(size_t)c];
        if (mode >= 0)
            for (int32_t e = C.ptr[mode]; e < C.ptr[mode + 1]; ++e) {
                cm.push_back(C.m[e]);
                cnu.push_back(ctx->h_pos[(size_t)C.nu[e]]);
                cg.push_back(C.g[e]);
            }
        cptr[(size_t)c + 1] = (int32_t)cm.size();
    }
    rc |= dev_upload(ctx, &ctx->d_cptr, cptr);
    rc |= dev_upload(ctx, &ctx->d_cm, cm);
    rc |= dev_upload(ctx, &ctx->d_cnu, cnu);
    rc |= dev_upload(ctx, &ctx->d_cg, cg);
    rc |= dev_upload(ctx, &ctx->d_pos, ctx->h_pos);
    rc |= dev_upload(ctx, &ctx->d_inv, ctx->h_inv);
    if (rc) return rc;
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_get_coupling_nnz(asgfem_ctx* ctx, int64_t* nnz) try {
    CTX_OR_FAIL(ctx);
    ASG_CHECK(ctx, ctx->N > 0 && nnz, ASGFEM_ESTATE, "multi-indices not set");
    *nnz = (int64_t)ctx->coup.m.size();
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_get_coupling_csc(asgfem_ctx* ctx, int64_t* colptr, int64_t* rowval, double* nzval) try {
    CTX_OR_FAIL(ctx);
    ASG_CHECK(ctx, ctx->N > 0, ASGFEM_ESTATE, "multi-indices not set");
    ASG_CHECK(ctx, colptr && rowval && nzval, ASGFEM_EINVAL, "null output");
    // G[(m-1)N + j, k]: column k holds, for every mode j coupled to k in direction m, the row (m-1)N + j.
    const int64_t N = ctx->N;
    const Coupling& C = ctx->coup;
    struct Ent {
        int64_t row;
        double v;
    };
    std::vector<std::vector<Ent>> cols((size_t)N);
    for (int64_t j = 0; j < N; ++j)
        for (int32_t e = C.ptr[j]; e < C.ptr[j + 1]; ++e)
            cols[C.nu[e]].push_back({(int64_t)(C.m[e] - 1) * N + j + 1, C.g[e]});
    int64_t p = 0;
    for (int64_t k = 0; k < N; ++k) {
        colptr[k] = p + 1;
        std::sort(cols[k].begin(), cols[k].end(), [](const Ent& a, const Ent& b) { return a.row < b.row; });
        for (auto& en : cols[k]) {
            rowval[p] = en.row;
            nzval[p] = en.v;
            ++p;
        }
    }
    colptr[N] = p + 1;
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_get_neighbours(asgfem_ctx* ctx, int64_t* plus, int64_t* minus) try {
    CTX_OR_FAIL(ctx);
    ASG_CHECK(ctx, ctx->N > 0, ASGFEM_ESTATE, "multi-indices not set");
    ASG_CHECK(ctx, plus && minus, ASGFEM_EINVAL, "null output");
    std::memcpy(plus, ctx->mis.plus.data(), sizeof(int64_t) * ctx->mis.plus.size());
    std::memcpy(minus, ctx->mis.minus.data(), sizeof(int64_t) * ctx->mis.minus.size());
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

// ---- pattern and values -------------------------------------------------------------------------
static int install_pattern(asgfem_ctx* ctx, int64_t n, const std::vector<int64_t>& colptr0,
                           const std::vector<int32_t>& row0) {
    // CSC (0-based, sorted rows) -> CSR (sorted columns) + map csc position -> csr position
    int64_t nnz = (int64_t)row0.size();
    ASG_CHECK(ctx, nnz < (1ll << 31) - 1, ASGFEM_EINVAL, "pattern with >= 2^31 nonzeros not supported");
    if (n != ctx->n) free_vec_storage(ctx);
    ctx->pattern_from_space = false;  // asgfem_assemble_stiffness sets it for the pattern it derives
    ctx->n = n;
    ctx->nnz = nnz;
    ctx->h_csc_colptr = colptr0;
    ctx->h_csc_row = row0;
    ctx->h_rowptr.assign((size_t)n + 1, 0);
    for (int64_t p = 0; p < nnz; ++p) ctx->h_rowptr[row0[p] + 1]++;
    for (int64_t i = 0; i < n; ++i) ctx->h_rowptr[i + 1] += ctx->h_rowptr[i];
    ctx->h_col.assign((size_t)nnz, 0);
    ctx->h_csc2csr.assign((size_t)nnz, 0);
    std::vector<int64_t> fill(ctx->h_rowptr.begin(), ctx->h_rowptr.end() - 1);
    for (int64_t c = 0; c < n; ++c)
        for (int64_t p = colptr0[c]; p < colptr0[c + 1]; ++p) {
            int64_t q = fill[row0[p]]++;
            ctx->h_col[q] = (int32_t)c;  // columns visited ascending -> CSR rows come out sorted
            ctx->h_csc2csr[p] = q;
        }
    if (ctx->d_rowptr) cudaFree(ctx->d_rowptr), ctx->d_rowptr = nullptr;
    int rc = dev_upload(ctx, &ctx->d_rowptr, ctx->h_rowptr);
    rc |= dev_upload(ctx, &ctx->d_col, ctx->h_col);
    if (rc) return rc;
    if (ctx->h_bmask.size() != (size_t)n) {
        ctx->h_bmask.assign((size_t)n, 0);
        ctx->h_bdofs.clear();
    }
    rc = dev_upload(ctx, &ctx->d_bmask, ctx->h_bmask);
    if (rc) return rc;
    if (ctx->d_vals) cudaFree(ctx->d_vals), ctx->d_vals = nullptr;
    ctx->M = -1;
    ctx->n_owned = -1;
    apply_free_plan(ctx);
    precond_free(ctx);
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int asgfem_set_pattern_csc(asgfem_ctx* ctx, int64_t n, const int64_t* colptr, const int64_t* rowval) try {
    CTX_OR_FAIL(ctx);
    ASG_CHECK(ctx, n >= 1 && colptr && rowval, ASGFEM_EINVAL, "set_pattern_csc: bad arguments");
    ASG_CHECK(ctx, n < (1ll << 31) - 1, ASGFEM_EINVAL, "n too large");
    ASG_CHECK(ctx, colptr[0] == 1, ASGFEM_EINVAL, "set_pattern_csc: colptr must be 1-based");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    int64_t nnz = colptr[n] - 1;
    ASG_CHECK(ctx, nnz >= 0, ASGFEM_EINVAL, "set_pattern_csc: negative nnz");
    std::vector<int64_t> cp((size_t)n + 1);
    std::vector<int32_t> rv((size_t)nnz);
    for (int64_t c = 0; c <= n; ++c) {
        cp[c] = colptr[c] - 1;
        ASG_CHECK(ctx, c == 0 || cp[c] >= cp[c - 1], ASGFEM_EINVAL, "set_pattern_csc: colptr not monotone");
    }
    for (int64_t c = 0; c < n; ++c)
        for (int64_t p = cp[c]; p < cp[c + 1]; ++p) {
            int64_t r = rowval[p] - 1;
            ASG_CHECK(ctx, r >= 0 && r < n, ASGFEM_EINVAL, "set_pattern_csc: row index out of range");
            ASG_CHECK(ctx, p == cp[c] || r > rv[p - 1], ASGFEM_EINVAL, "set_pattern_csc: rows not strictly sorted in a column");
            rv[p] = (int32_t)r;
        }
    return install_pattern(ctx, n, cp, rv);
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_set_num_stiffness(asgfem_ctx* ctx, int32_t M) try {
    CTX_OR_FAIL(ctx);
    ASG_CHECK(ctx, ctx->n > 0, ASGFEM_ESTATE, "pattern not set");
    ASG_CHECK(ctx, M >= 0 && M < 4096, ASGFEM_EINVAL, "bad number of KLE terms");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    if (ctx->d_vals) cudaFree(ctx->d_vals), ctx->d_vals = nullptr;
    size_t bytes = sizeof(double) * (size_t)(M + 1) * (size_t)std::max<int64_t>(ctx->nnz, 1);
    ASG_CUDA(ctx, cudaMalloc((void**)&ctx->d_vals, bytes));
    ASG_CUDA(ctx, cudaMemsetAsync(ctx->d_vals, 0, bytes, ctx->stream));
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->M = M;
    apply_free_plan(ctx);
    precond_free(ctx);
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

static int upload_values_csr(asgfem_ctx* ctx, int32_t m, const std::vector<double>& csr) {
    ASG_CUDA(ctx, cudaMemcpyAsync(ctx->d_vals + (size_t)m * ctx->nnz, csr.data(), sizeof(double) * ctx->nnz,
                                  cudaMemcpyHostToDevice, ctx->stream));
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (m == 0) precond_free(ctx);
    return 0;
}

extern "C" int asgfem_set_stiffness(asgfem_ctx* ctx, int32_t m, const double* nzval) try {
    CTX_OR_FAIL(ctx);
    ASG_CHECK(ctx, ctx->M >= 0, ASGFEM_ESTATE, "set_num_stiffness first");
    ASG_CHECK(ctx, m >= 0 && m <= ctx->M && nzval, ASGFEM_EINVAL, "set_stiffness: m out of range");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    std::vector<double> csr((size_t)ctx->nnz);
    for (int64_t p = 0; p < ctx->nnz; ++p) csr[ctx->h_csc2csr[p]] = nzval[p];
    return upload_values_csr(ctx, m, csr);
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_set_stiffness_csc(asgfem_ctx* ctx, int32_t m, const int64_t* colptr, const int64_t* rowval,
                                        const double* nzval) try {
    CTX_OR_FAIL(ctx);
    ASG_CHECK(ctx, ctx->M >= 0, ASGFEM_ESTATE, "set_num_stiffness first");
    ASG_CHECK(ctx, m >= 0 && m <= ctx->M && colptr && rowval && nzval, ASGFEM_EINVAL, "set_stiffness_csc: bad arguments");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    std::vector<double> csr((size_t)ctx->nnz, 0.0);
    for (int64_t c = 0; c < ctx->n; ++c) {
        int64_t q = ctx->h_csc_colptr[c], q1 = ctx->h_csc_colptr[c + 1];
        for (int64_t p = colptr[c] - 1; p < colptr[c + 1] - 1; ++p) {
            int64_t r = rowval[p] - 1;
            while (q < q1 && ctx->h_csc_row[q] < r) ++q;
            ASG_CHECK(ctx, q < q1 && ctx->h_csc_row[q] == r, ASGFEM_EINVAL,
                      "set_stiffness_csc: entry outside the shared pattern");
            csr[ctx->h_csc2csr[q]] = nzval[p];
        }
    }
    return upload_values_csr(ctx, m, csr);
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_get_stiffness(asgfem_ctx* ctx, int32_t m, double* nzval) try {
    CTX_OR_FAIL(ctx);
    ASG_CHECK(ctx, ctx->M >= 0 && m >= 0 && m <= ctx->M && nzval, ASGFEM_EINVAL, "get_stiffness: m out of range");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    std::vector<double> csr((size_t)ctx->nnz);
    ASG_CUDA(ctx, cudaMemcpyAsync(csr.data(), ctx->d_vals + (size_t)m * ctx->nnz, sizeof(double) * ctx->nnz,
                                  cudaMemcpyDeviceToHost, ctx->stream));
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int64_t p = 0; p < ctx->nnz; ++p) nzval[p] = csr[ctx->h_csc2csr[p]];
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_get_pattern_nnz(asgfem_ctx* ctx, int64_t* nnz) try {
    CTX_OR_FAIL(ctx);
    ASG_CHECK(ctx, ctx->n > 0 && nnz, ASGFEM_ESTATE, "pattern not set");
    *nnz = ctx->nnz;
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_get_pattern_csc(asgfem_ctx* ctx, int64_t* colptr, int64_t* rowval) try {
    CTX_OR_FAIL(ctx);
    ASG_CHECK(ctx, ctx->n > 0 && colptr && rowval, ASGFEM_ESTATE, "pattern not set");
    for (int64_t c = 0; c <= ctx->n; ++c) colptr[c] = ctx->h_csc_colptr[c] + 1;
    for (int64_t p = 0; p < ctx->nnz; ++p) rowval[p] = ctx->h_csc_row[p] + 1;
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_set_bdofs(asgfem_ctx* ctx, int64_t nb, const int64_t* bdofs) try {
    CTX_OR_FAIL(ctx);
    ASG_CHECK(ctx, ctx->n > 0, ASGFEM_ESTATE, "pattern not set");
    ASG_CHECK(ctx, nb >= 0 && (nb == 0 || bdofs), ASGFEM_EINVAL, "set_bdofs: bad arguments");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    std::vector<uint8_t> mask((size_t)ctx->n, 0);
    for (int64_t k = 0; k < nb; ++k) {
        ASG_CHECK(ctx, bdofs[k] >= 1 && bdofs[k] <= ctx->n, ASGFEM_EINVAL, "set_bdofs: dof out of range");
        mask[bdofs[k] - 1] = 1;
    }
    ctx->h_bmask.swap(mask);
    ctx->h_bdofs.assign(bdofs, bdofs + nb);
    int rc = dev_upload(ctx, &ctx->d_bmask, ctx->h_bmask);
    if (rc) return rc;
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    precond_free(ctx);
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

// ---- mesh / space / coefficient -----------------------------------------------------------------
// A pattern derived from celldofs (asgfem_assemble_stiffness without asgfem_set_pattern_csc) belongs to the mesh / space it
// was built from: a new mesh or space drops it together with the matrices, the operator plans and the preconditioner, so
// that the next assembly rebuilds it (otherwise a space with the same ndofs but other connectivity would be assembled on
// the stale pattern).
static void drop_derived_pattern(asgfem_ctx* ctx) {
    if (!ctx->pattern_from_space) return;
    ctx->pattern_from_space = false;
    ctx->h_rowptr.clear();
    ctx->h_col.clear();
    ctx->h_csc_colptr.clear();
    ctx->h_csc_row.clear();
    ctx->h_csc2csr.clear();
    ctx->nnz = 0;
    ctx->M = -1;
    if (ctx->d_vals) {
        cudaFree(ctx->d_vals);
        ctx->d_vals = nullptr;
    }
    apply_free_plan(ctx);
    precond_free(ctx);
}

extern "C" int asgfem_set_mesh(asgfem_ctx* ctx, int64_t nnodes, int64_t ncells, const double* coords,
                               const int32_t* cellnodes) try {
    CTX_OR_FAIL(ctx);
    ASG_CHECK(ctx, nnodes >= 3 && ncells >= 1 && coords && cellnodes, ASGFEM_EINVAL, "set_mesh: bad arguments");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    drop_derived_pattern(ctx);
    ctx->h_cell_owned.clear();
    ctx->nnodes = nnodes;
    ctx->ncells = ncells;
    ctx->h_coords.assign(coords, coords + 2 * nnodes);
    ctx->h_cellnodes.resize((size_t)(3 * ncells));
    for (int64_t k = 0; k < 3 * ncells; ++k) {
        ASG_CHECK(ctx, cellnodes[k] >= 1 && cellnodes[k] <= nnodes, ASGFEM_EINVAL, "set_mesh: node id out of range");
        ctx->h_cellnodes[k] = cellnodes[k] - 1;
    }
    int rc = dev_upload(ctx, &ctx->d_coords, ctx->h_coords);
    rc |= dev_upload(ctx, &ctx->d_cellnodes, ctx->h_cellnodes);
    if (rc) return rc;
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_set_space(asgfem_ctx* ctx, int32_t order, int64_t ndofs, int32_t ndofs4cell,
                                const int32_t* celldofs) try {
    CTX_OR_FAIL(ctx);
    ASG_CHECK(ctx, ctx->ncells > 0, ASGFEM_ESTATE, "set_mesh first");
    ASG_CHECK(ctx, (order == 1 && ndofs4cell == 3) || (order == 2 && ndofs4cell == 6), ASGFEM_EINVAL,
              "set_space: only H1Pk{1,2,1} (3 dofs/cell) and H1Pk{1,2,2} (6 dofs/cell) are supported");
    ASG_CHECK(ctx, ndofs >= 3 && celldofs, ASGFEM_EINVAL, "set_space: bad arguments");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    drop_derived_pattern(ctx);
    ctx->order = order;
    ctx->ndofs4cell = ndofs4cell;
    ctx->ndofs_space = ndofs;
    ctx->h_celldofs.resize((size_t)ndofs4cell * ctx->ncells);
    for (size_t k = 0; k < ctx->h_celldofs.size(); ++k) {
        ASG_CHECK(ctx, celldofs[k] >= 1 && celldofs[k] <= ndofs, ASGFEM_EINVAL, "set_space: dof id out of range");
        ctx->h_celldofs[k] = celldofs[k] - 1;
    }
    int rc = dev_upload(ctx, &ctx->d_celldofs, ctx->h_celldofs);
    if (rc) return rc;
    if (ctx->h_rowptr.empty()) {  // no matrices yet: the space sizes the vectors (estimator-only use)
        if (ctx->n != ndofs) free_vec_storage(ctx);
        ctx->n = ndofs;
    }
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_set_coefficient_cosinus(asgfem_ctx* ctx, int64_t maxm, double mean, const double* decay_factors,
                                              const int64_t* b1, const int64_t* b2) try {
    CTX_OR_FAIL(ctx);
    ASG_CHECK(ctx, maxm >= 0 && (maxm == 0 || (decay_factors && b1 && b2)), ASGFEM_EINVAL, "set_coefficient: bad arguments");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    ctx->maxm = maxm;
    ctx->mean = mean;
    ctx->h_decay.assign(decay_factors, decay_factors + maxm);
    ctx->h_b1.assign(b1, b1 + maxm);
    ctx->h_b2.assign(b2, b2 + maxm);
    std::vector<int32_t> i1((size_t)maxm), i2((size_t)maxm);
    for (int64_t k = 0; k < maxm; ++k) {
        i1[k] = (int32_t)b1[k];
        i2[k] = (int32_t)b2[k];
    }
    int rc = dev_upload(ctx, &ctx->d_decay, ctx->h_decay);
    rc |= dev_upload(ctx, &ctx->d_b1, i1);
    rc |= dev_upload(ctx, &ctx->d_b2, i2);
    if (rc) return rc;
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_assemble_stiffness(asgfem_ctx* ctx, int32_t M, int32_t nq, const double* xref, const double* w) try {
    CTX_OR_FAIL(ctx);
    ASG_CHECK(ctx, ctx->order > 0, ASGFEM_ESTATE, "set_mesh / set_space first");
    ASG_CHECK(ctx, M >= 0 && M <= ctx->maxm, ASGFEM_EINVAL, "assemble_stiffness: M exceeds maxm of the coefficient");
    ASG_CHECK(ctx, nq >= 1 && nq <= 64 && xref && w, ASGFEM_EINVAL, "assemble_stiffness: bad quadrature rule");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    if (ctx->h_rowptr.empty()) {
        // shared pattern from celldofs: all dof pairs sharing a cell (symmetric -> CSC == CSR)
        int64_t n = ctx->ndofs_space;
        int nd = ctx->ndofs4cell;
        std::vector<int64_t> cnt((size_t)n + 1, 0);
        for (int64_t c = 0; c < ctx->ncells; ++c)
            for (int i = 0; i < nd; ++i) cnt[ctx->h_celldofs[c * nd + i] + 1] += nd;
        for (int64_t i = 0; i < n; ++i) cnt[i + 1] += cnt[i];
        std::vector<int32_t> tmp((size_t)cnt[n]);
        std::vector<int64_t> fill(cnt.begin(), cnt.end() - 1);
        for (int64_t c = 0; c < ctx->ncells; ++c)
            for (int i = 0; i < nd; ++i) {
                int32_t r = ctx->h_celldofs[c * nd + i];
                for (int j = 0; j < nd; ++j) tmp[fill[r]++] = ctx->h_celldofs[c * nd + j];
            }
        std::vector<int64_t> cp((size_t)n + 1, 0);
        std::vector<int32_t> rv;
        rv.reserve(tmp.size() / 2);
        for (int64_t i = 0; i < n; ++i) {
            std::sort(tmp.begin() + cnt[i], tmp.begin() + cnt[i + 1]);
            auto e = std::unique(tmp.begin() + cnt[i], tmp.begin() + cnt[i + 1]);
            rv.insert(rv.end(), tmp.begin() + cnt[i], e);
            cp[i + 1] = (int64_t)rv.size();
        }
        int rc = install_pattern(ctx, n, cp, rv);
        if (rc) return rc;
        ctx->pattern_from_space = true;
    }
    ASG_CHECK(ctx, ctx->n == ctx->ndofs_space, ASGFEM_EINVAL, "assemble_stiffness: pattern size differs from ndofs of the space");
    int rc = asgfem_set_num_stiffness(ctx, M);
    if (rc) return rc;
    return assemble_stiffness(ctx, M, nq, xref, w);
}
ASG_BOUNDARY_CATCH(ctx)

// Log-transformed primal problem (logpoisson_primal.jl:95-105): plane 0 = the Laplacian A (= A + N0, N0 is empty in the
// reference), planes 1..M = the convection matrices N_m, on the pattern derived from celldofs (or the caller's)
extern "C" int asgfem_assemble_logprimal(asgfem_ctx* ctx, int32_t M, int32_t nq, const double* xref, const double* w) try {
    CTX_OR_FAIL(ctx);
    ASG_CHECK(ctx, ctx->order > 0, ASGFEM_ESTATE, "set_mesh / set_space first");
    ASG_CHECK(ctx, M >= 0 && M <= ctx->maxm, ASGFEM_EINVAL, "assemble_logprimal: M exceeds maxm of the coefficient");
    ASG_CHECK(ctx, nq >= 1 && nq <= 64 && xref && w, ASGFEM_EINVAL, "assemble_logprimal: bad quadrature rule");
    // the stiffness entry point builds the pattern and sizes the value planes; its values are overwritten below
    int rc = asgfem_assemble_stiffness(ctx, 0, nq, xref, w);
    if (rc) return rc;
    if ((rc = asgfem_set_num_stiffness(ctx, M))) return rc;
    ctx->h_precond_vals.clear();  // the preconditioner is factorised from plane 0 = A
    return assemble_stiffness(ctx, M, nq, xref, w, 1);
}
ASG_BOUNDARY_CATCH(ctx)

// ---- vectors ------------------------------------------------------------------------------------
static int check_slot(asgfem_ctx* ctx, int32_t slot) {
    ASG_CHECK(ctx, slot >= 0 && slot < (int32_t)ctx->slots.size() && ctx->slots[slot], ASGFEM_EINVAL,
              "vector slot out of range (asgfem_vec_alloc first)");
    return 0;
}

extern "C" int asgfem_vec_alloc(asgfem_ctx* ctx, int32_t nslots) try {
    CTX_OR_FAIL(ctx);
    ASG_CHECK(ctx, ctx->n > 0 && ctx->N > 0, ASGFEM_ESTATE, "pattern and multi-indices must be set before vec_alloc");
    ASG_CHECK(ctx, nslots >= 0 && nslots <= 64, ASGFEM_EINVAL, "bad slot count");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    size_t bytes = sizeof(double) * (size_t)ctx->n * (size_t)ctx->ld;
    while ((int32_t)ctx->slots.size() > nslots) {
        if (ctx->slots.back()) cudaFree(ctx->slots.back());
        ctx->slots.pop_back();
    }
    while ((int32_t)ctx->slots.size() < nslots) {
        double* p = nullptr;
        ASG_CUDA(ctx, cudaMalloc((void**)&p, bytes));
        ctx->slots.push_back(p);
        ASG_CUDA(ctx, cudaMemsetAsync(p, 0, bytes, ctx->stream));
    }
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_vec_upload(asgfem_ctx* ctx, int32_t slot, const double* host) try {
    CTX_OR_FAIL(ctx);
    if (check_slot(ctx, slot)) return ASGFEM_EINVAL;
    ASG_CHECK(ctx, host, ASGFEM_EINVAL, "null host pointer");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    return vec_to_device_layout(ctx, host, ctx->slots[slot]);
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_vec_download(asgfem_ctx* ctx, int32_t slot, double* host) try {
    CTX_OR_FAIL(ctx);
    if (check_slot(ctx, slot)) return ASGFEM_EINVAL;
    ASG_CHECK(ctx, host, ASGFEM_EINVAL, "null host pointer");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    return vec_to_host_layout(ctx, ctx->slots[slot], host);
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_vec_zero(asgfem_ctx* ctx, int32_t slot) try {
    CTX_OR_FAIL(ctx);
    if (check_slot(ctx, slot)) return ASGFEM_EINVAL;
    if (set_device(ctx)) return ASGFEM_ECUDA;
    ASG_CUDA(ctx, cudaMemsetAsync(ctx->slots[slot], 0, sizeof(double) * ctx->n * ctx->ld, ctx->stream));
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_vec_fill_random(asgfem_ctx* ctx, int32_t slot, uint64_t seed) try {
    CTX_OR_FAIL(ctx);
    if (check_slot(ctx, slot)) return ASGFEM_EINVAL;
    if (set_device(ctx)) return ASGFEM_ECUDA;
    int rc = vec_fill_random(ctx, ctx->slots[slot], seed);
    if (rc) return rc;
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_vec_dot(asgfem_ctx* ctx, int32_t a, int32_t b, double* out) try {
    CTX_OR_FAIL(ctx);
    if (check_slot(ctx, a) || check_slot(ctx, b)) return ASGFEM_EINVAL;
    ASG_CHECK(ctx, out, ASGFEM_EINVAL, "null output");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    return vec_dot(ctx, ctx->slots[a], ctx->slots[b], ctx->n, out);
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_vec_dot_owned(asgfem_ctx* ctx, int32_t a, int32_t b, double* out) try {
    CTX_OR_FAIL(ctx);
    if (check_slot(ctx, a) || check_slot(ctx, b)) return ASGFEM_EINVAL;
    ASG_CHECK(ctx, out, ASGFEM_EINVAL, "null output");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    return vec_dot(ctx, ctx->slots[a], ctx->slots[b], ctx->n_owned >= 0 ? ctx->n_owned : ctx->n, out);
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_vec_axpy(asgfem_ctx* ctx, double alpha, int32_t x, int32_t y) try {
    CTX_OR_FAIL(ctx);
    if (check_slot(ctx, x) || check_slot(ctx, y)) return ASGFEM_EINVAL;
    if (set_device(ctx)) return ASGFEM_ECUDA;
    int rc = vec_axpy(ctx, alpha, ctx->slots[x], ctx->slots[y]);
    if (rc) return rc;
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_vec_xpay(asgfem_ctx* ctx, int32_t x, double beta, int32_t y) try {
    CTX_OR_FAIL(ctx);
    if (check_slot(ctx, x) || check_slot(ctx, y)) return ASGFEM_EINVAL;
    if (set_device(ctx)) return ASGFEM_ECUDA;
    int rc = vec_xpay(ctx, ctx->slots[x], beta, ctx->slots[y]);
    if (rc) return rc;
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_vec_copy(asgfem_ctx* ctx, int32_t src, int32_t dst) try {
    CTX_OR_FAIL(ctx);
    if (check_slot(ctx, src) || check_slot(ctx, dst)) return ASGFEM_EINVAL;
    if (set_device(ctx)) return ASGFEM_ECUDA;
    if (src != dst)
        ASG_CUDA(ctx, cudaMemcpyAsync(ctx->slots[dst], ctx->slots[src], sizeof(double) * ctx->n * ctx->ld,
                                      cudaMemcpyDeviceToDevice, ctx->stream));
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

// ---- operator -----------------------------------------------------------------------------------
static int ensure_ready_for_apply(asgfem_ctx* ctx) {
    ASG_CHECK(ctx, ctx->n > 0 && ctx->M >= 0, ASGFEM_ESTATE, "apply: stiffness matrices not set");
    ASG_CHECK(ctx, ctx->N > 0, ASGFEM_ESTATE, "apply: multi-indices not set");
    for (int32_t m : ctx->coup.m)
        ASG_CHECK(ctx, m <= ctx->M, ASGFEM_EINVAL,
                  "apply: multi-indices couple in a direction m > number of stiffness matrices (maxlength_multiindices > length(Am))");
    if (!ctx->apply_ready) {
        int rc = apply_build_plan(ctx);
        if (rc) return rc;
    }
    return 0;
}

extern "C" int asgfem_set_apply_variant(asgfem_ctx* ctx, int32_t variant) try {
    CTX_OR_FAIL(ctx);
    ASG_CHECK(ctx, variant == 0 || variant == 1 || variant == 7 || variant == 9, ASGFEM_EINVAL, "apply variant must be 0 (automatic), 1, 7 or 9");
    ctx->apply_variant = variant;
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_apply(asgfem_ctx* ctx, int32_t sx, int32_t sy) try {
    CTX_OR_FAIL(ctx);
    if (check_slot(ctx, sx) || check_slot(ctx, sy)) return ASGFEM_EINVAL;
    ASG_CHECK(ctx, sx != sy, ASGFEM_EINVAL, "apply: x and y must be different slots");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    int rc = ensure_ready_for_apply(ctx);
    if (rc) return rc;
    if (dist_active(ctx)) {
        // row-sharded operation: halo exchange overlapped with the rows that need no halo column (dist.cu)
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        ASG_CUDA(ctx, cudaEventCreate(&e0));
        ASG_CUDA(ctx, cudaEventCreate(&e1));
        ASG_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        rc = dist_apply(ctx, ctx->slots[sx], ctx->slots[sy]);
        if (!rc) {
            cudaEventRecord(e1, ctx->stream);
            cudaStreamSynchronize(ctx->stream);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            ctx->last_apply_ms = ms;
        }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        return rc;
    }
    rc = apply_launch(ctx, ctx->slots[sx], ctx->slots[sy]);
    if (rc) return rc;
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    ASG_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_apply_ms = ms;
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_apply_rows(asgfem_ctx* ctx, int32_t sx, int32_t sy, int64_t row0, int64_t row1) try {
    CTX_OR_FAIL(ctx);
    if (check_slot(ctx, sx) || check_slot(ctx, sy)) return ASGFEM_EINVAL;
    ASG_CHECK(ctx, sx != sy, ASGFEM_EINVAL, "apply_rows: x and y must be different slots");
    ASG_CHECK(ctx, row0 >= 0 && row1 >= row0, ASGFEM_EINVAL, "apply_rows: bad row range");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    int rc = ensure_ready_for_apply(ctx);
    if (rc) return rc;
    ctx->last_apply_ms = 0;
    const int64_t nrows_all = ctx->n_owned >= 0 ? ctx->n_owned : ctx->n;
    if (row0 >= std::min(row1, nrows_all)) return 0;
    rc = apply_launch(ctx, ctx->slots[sx], ctx->slots[sy], row0, row1);
    if (rc) return rc;
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    ASG_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_apply_ms = ms;
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_last_apply_ms(asgfem_ctx* ctx, double* ms) try {
    CTX_OR_FAIL(ctx);
    ASG_CHECK(ctx, ms, ASGFEM_EINVAL, "null output");
    *ms = ctx->last_apply_ms;
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_last_estimate_ms(asgfem_ctx* ctx, double* ms) try {
    CTX_OR_FAIL(ctx);
    ASG_CHECK(ctx, ms, ASGFEM_EINVAL, "null output");
    *ms = ctx->last_estimate_ms;
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

static int ensure_work_slots(asgfem_ctx* ctx, int need) {
    if ((int)ctx->slots.size() >= need) return 0;
    return asgfem_vec_alloc(ctx, need);
}

extern "C" int asgfem_apply_host(asgfem_ctx* ctx, const double* x, double* Ax) try {
    CTX_OR_FAIL(ctx);
    ASG_CHECK(ctx, x && Ax, ASGFEM_EINVAL, "null host pointer");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    int rc = ensure_ready_for_apply(ctx);
    if (rc) return rc;
    if ((rc = ensure_work_slots(ctx, 2))) return rc;
    // upload, operator and download overlap block by block unless the kernel cannot work on row ranges
    if (ctx->n_owned < 0 && x != Ax)
        return apply_host_pipelined(ctx, x, Ax, ctx->slots[0], ctx->slots[1]);
    if ((rc = vec_to_device_layout(ctx, x, ctx->slots[0]))) return rc;
    if ((rc = apply_launch(ctx, ctx->slots[0], ctx->slots[1]))) return rc;
    return vec_to_host_layout(ctx, ctx->slots[1], Ax);
}
ASG_BOUNDARY_CATCH(ctx)

// ---- preconditioner -----------------------------------------------------------------------------
extern "C" int asgfem_precond_setup(asgfem_ctx* ctx) try {
    CTX_OR_FAIL(ctx);
    ASG_CHECK(ctx, ctx->n > 0 && ctx->M >= 0, ASGFEM_ESTATE, "precond_setup: K_0 not set");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    return precond_setup(ctx);
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_host_factor_solve(int64_t n, const int64_t* rowptr, const int32_t* col, const double* val,
                                        const uint8_t* is_boundary, const double* coords, const double* b, double* x,
                                        int64_t* lnz, char* err, int32_t errlen) try {
    auto fail = [&](int rc, const std::string& msg) {
        if (err && errlen > 0) snprintf(err, (size_t)errlen, "%s", msg.c_str());
        return rc;
    };
    if (n < 0 || !rowptr || !col || !val || !is_boundary || !b || !x) return fail(ASGFEM_EINVAL, "host_factor_solve: null pointer or negative size");
    try {
        CholFactor F;
        std::string msg;
        int rc = cholesky_reduced(n, rowptr, col, val, is_boundary, coords, 256, F, msg);
        if (rc) return fail(rc, msg);
        std::vector<double> w((size_t)F.n);
        for (int64_t k = 0; k < F.n; ++k) w[(size_t)k] = b[F.perm[(size_t)k]];
        for (int64_t k = 0; k < F.n; ++k) {  // L w = b
            double v = w[(size_t)k];
            for (int64_t p = F.Lp[(size_t)k]; p < F.Lp[(size_t)k + 1]; ++p) v -= F.Lx[(size_t)p] * w[(size_t)F.Li[(size_t)p]];
            w[(size_t)k] = v * F.dinv[(size_t)k];
        }
        for (int64_t k = F.n - 1; k >= 0; --k) {  // L^T x = w
            const double v = w[(size_t)k] * F.dinv[(size_t)k];
            w[(size_t)k] = v;
            for (int64_t p = F.Lp[(size_t)k]; p < F.Lp[(size_t)k + 1]; ++p) w[(size_t)F.Li[(size_t)p]] -= F.Lx[(size_t)p] * v;
        }
        for (int64_t i = 0; i < n; ++i) x[i] = 0.0;
        for (int64_t k = 0; k < F.n; ++k) x[F.perm[(size_t)k]] = w[(size_t)k];
        if (lnz) *lnz = (int64_t)F.Li.size();
    } catch (const std::exception& e) {
        return fail(ASGFEM_ENOMEM, std::string("host_factor_solve: ") + e.what());
    }
    return 0;
}
ASG_BOUNDARY_CATCH(nullptr)

extern "C" int asgfem_precond_apply(asgfem_ctx* ctx, int32_t sr, int32_t sz) try {
    CTX_OR_FAIL(ctx);
    if (check_slot(ctx, sr) || check_slot(ctx, sz)) return ASGFEM_EINVAL;
    if (set_device(ctx)) return ASGFEM_ECUDA;
    if (!ctx->precond) {
        int rc = precond_setup(ctx);
        if (rc) return rc;
    }
    int rc = precond_apply(ctx, ctx->slots[sr], ctx->slots[sz]);
    if (rc) return rc;
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_precond_apply_host(asgfem_ctx* ctx, const double* b, double* y) try {
    CTX_OR_FAIL(ctx);
    ASG_CHECK(ctx, b && y, ASGFEM_EINVAL, "null host pointer");
    ASG_CHECK(ctx, ctx->N > 0, ASGFEM_ESTATE, "multi-indices not set");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    int rc;
    if (!ctx->precond && (rc = precond_setup(ctx))) return rc;
    if ((rc = ensure_work_slots(ctx, 2))) return rc;
    if ((rc = vec_to_device_layout(ctx, b, ctx->slots[0]))) return rc;
    if ((rc = precond_apply(ctx, ctx->slots[0], ctx->slots[1]))) return rc;
    return vec_to_host_layout(ctx, ctx->slots[1], y);
}
ASG_BOUNDARY_CATCH(ctx)

// ---- Krylov driver ------------------------------------------------------------------------------
extern "C" int asgfem_pcg(asgfem_ctx* ctx, const double* b0, int32_t slot_x, double atol, double rtol, int64_t itmax,
                          asgfem_stats* stats) try {
    CTX_OR_FAIL(ctx);
    if (check_slot(ctx, slot_x)) return ASGFEM_EINVAL;
    ASG_CHECK(ctx, b0, ASGFEM_EINVAL, "null b0");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    int rc = ensure_ready_for_apply(ctx);
    if (rc) return rc;
    if (!ctx->precond && !dist_has_global_precond(ctx) && (rc = precond_setup(ctx))) return rc;
    return pcg_solve(ctx, b0, ctx->slots[slot_x], atol, rtol, itmax, stats);
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_solve_primal_host(asgfem_ctx* ctx, double* sol, const double* b0, double atol, double rtol,
                                        int64_t itmax, asgfem_stats* stats) try {
    CTX_OR_FAIL(ctx);
    ASG_CHECK(ctx, sol && b0, ASGFEM_EINVAL, "null host pointer");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    int rc = ensure_ready_for_apply(ctx);
    if (rc) return rc;
    if ((rc = ensure_work_slots(ctx, 1))) return rc;
    if ((rc = vec_to_device_layout(ctx, sol, ctx->slots[0]))) return rc;
    if (!ctx->precond && (rc = precond_setup(ctx))) return rc;
    if ((rc = pcg_solve(ctx, b0, ctx->slots[0], atol, rtol, itmax, stats))) return rc;
    return vec_to_host_layout(ctx, ctx->slots[0], sol);
}
ASG_BOUNDARY_CATCH(ctx)

// out (n x nsamples, column s = solution of sample s): K(xi_s) u_s = b with u_s = 0 on the Dirichlet dofs
extern "C" int asgfem_solve_samples_host(asgfem_ctx* ctx, double* out, const double* b, double atol, double rtol, int64_t itmax,
                                         asgfem_stats* stats) try {
    CTX_OR_FAIL(ctx);
    ASG_CHECK(ctx, ctx->sample_mode, ASGFEM_ESTATE, "solve_samples: asgfem_set_samples first");
    ASG_CHECK(ctx, out && b, ASGFEM_EINVAL, "null host pointer");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    int rc = ensure_ready_for_apply(ctx);
    if (rc) return rc;
    for (int32_t m : ctx->coup_col.m) ASG_CHECK(ctx, m <= ctx->M, ASGFEM_EINVAL, "solve_samples: more sample dimensions than stiffness matrices K_m");
    if ((rc = ensure_work_slots(ctx, 1))) return rc;
    ASG_CUDA(ctx, cudaMemsetAsync(ctx->slots[0], 0, sizeof(double) * (size_t)ctx->n * (size_t)ctx->ld, ctx->stream));
    if (!ctx->precond && (rc = precond_setup(ctx))) return rc;
    if ((rc = pcg_solve(ctx, b, ctx->slots[0], atol, rtol, itmax, stats))) return rc;
    return vec_to_host_layout(ctx, ctx->slots[0], out);
}
ASG_BOUNDARY_CATCH(ctx)


// ---- log-transformed primal problem ----------------------------------------------------------------
extern "C" int asgfem_set_precond_matrix_csc(asgfem_ctx* ctx, const int64_t* colptr, const int64_t* rowval,
                                             const double* nzval) try {
    CTX_OR_FAIL(ctx);
    ASG_CHECK(ctx, ctx->n > 0 && ctx->nnz > 0, ASGFEM_ESTATE, "set_precond_matrix_csc: set_pattern_csc first");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    precond_free(ctx);
    ctx->h_precond_vals.clear();
    if (!colptr) return 0;
    ASG_CHECK(ctx, rowval && nzval, ASGFEM_EINVAL, "set_precond_matrix_csc: bad arguments");
    std::vector<double> csr((size_t)ctx->nnz, 0.0);
    for (int64_t c = 0; c < ctx->n; ++c) {
        int64_t q = ctx->h_csc_colptr[c], q1 = ctx->h_csc_colptr[c + 1];
        for (int64_t p = colptr[c] - 1; p < colptr[c + 1] - 1; ++p) {
            int64_t r = rowval[p] - 1;
            while (q < q1 && ctx->h_csc_row[q] < r) ++q;
            ASG_CHECK(ctx, q < q1 && ctx->h_csc_row[q] == r, ASGFEM_EINVAL,
                      "set_precond_matrix_csc: entry outside the shared pattern");
            csr[ctx->h_csc2csr[q]] = nzval[p];
        }
    }
    ctx->h_precond_vals.swap(csr);
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

// load vectors b[mu] = (lambda_mu f, phi_i) of the log-transformed primal problem into a device slot (logpoisson_primal.jl:108-127)
extern "C" int asgfem_assemble_logprimal_rhs(asgfem_ctx* ctx, int32_t nq, const double* xref, const double* w, const double* f_at_qp,
                                             int32_t ntrunc, int32_t slot_b) try {
    CTX_OR_FAIL(ctx);
    if (check_slot(ctx, slot_b)) return ASGFEM_EINVAL;
    ASG_CHECK(ctx, ctx->order > 0 && ctx->ncells > 0 && ctx->ndofs_space == ctx->n, ASGFEM_ESTATE, "assemble_logprimal_rhs: set_mesh / set_space first");
    ASG_CHECK(ctx, ctx->mis.N == ctx->N && ctx->N > 0 && !ctx->sample_mode, ASGFEM_ESTATE, "assemble_logprimal_rhs: multi-indices not set");
    ASG_CHECK(ctx, nq >= 1 && nq <= 64 && xref && w && f_at_qp, ASGFEM_EINVAL, "assemble_logprimal_rhs: bad quadrature rule / rhs values");
    ASG_CHECK(ctx, ntrunc >= 0 && ntrunc <= ctx->maxm && ctx->mis.M <= ctx->maxm, ASGFEM_EINVAL,
              "assemble_logprimal_rhs: N_truncate / multi-index length exceed maxm of the coefficient");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    return assemble_logprimal_rhs(ctx, nq, xref, w, f_at_qp, ntrunc, ctx->slots[slot_b]);
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_bicgstab(asgfem_ctx* ctx, int32_t slot_b, int32_t slot_x, double atol, double rtol, int64_t itmax,
                               asgfem_stats* stats) try {
    CTX_OR_FAIL(ctx);
    if (check_slot(ctx, slot_b) || check_slot(ctx, slot_x)) return ASGFEM_EINVAL;
    ASG_CHECK(ctx, slot_b != slot_x, ASGFEM_EINVAL, "bicgstab: b and x must be different slots");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    int rc = ensure_ready_for_apply(ctx);
    if (rc) return rc;
    if (!ctx->precond && (rc = precond_setup(ctx))) return rc;
    return bicgstab_solve(ctx, ctx->slots[slot_b], ctx->slots[slot_x], atol, rtol, itmax, stats);
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_solve_logprimal_host(asgfem_ctx* ctx, double* sol, const double* b, double atol, double rtol,
                                           int64_t itmax, asgfem_stats* stats) try {
    CTX_OR_FAIL(ctx);
    ASG_CHECK(ctx, sol && b, ASGFEM_EINVAL, "null host pointer");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    int rc = ensure_ready_for_apply(ctx);
    if (rc) return rc;
    if ((rc = ensure_work_slots(ctx, 2))) return rc;
    if ((rc = vec_to_device_layout(ctx, sol, ctx->slots[0]))) return rc;
    if ((rc = vec_to_device_layout(ctx, b, ctx->slots[1]))) return rc;
    if ((rc = vec_axpy(ctx, 1.0, ctx->slots[0], ctx->slots[1]))) return rc;  // b = deepcopy(sol) + b0   (:149-152)
    if (!ctx->precond && (rc = precond_setup(ctx))) return rc;
    if ((rc = bicgstab_solve(ctx, ctx->slots[1], ctx->slots[0], atol, rtol, itmax, stats))) return rc;
    return vec_to_host_layout(ctx, ctx->slots[0], sol);
}
ASG_BOUNDARY_CATCH(ctx)

// ---- evaluation at samples ------------------------------------------------------------------------
extern "C" int asgfem_evaluate_samples(asgfem_ctx* ctx, int32_t slot_u, int64_t nsamples, int64_t M_in, int32_t nvals,
                                       const double* vals, double* out) try {
    CTX_OR_FAIL(ctx);
    if (check_slot(ctx, slot_u)) return ASGFEM_EINVAL;
    ASG_CHECK(ctx, ctx->N > 0 && ctx->n > 0, ASGFEM_ESTATE, "evaluate_samples: multi-indices / pattern not set");
    ASG_CHECK(ctx, nsamples >= 1 && nvals >= 1 && vals && out, ASGFEM_EINVAL, "evaluate_samples: bad arguments");
    const int64_t N = ctx->N, M = ctx->mis.M, n = ctx->n;
    ASG_CHECK(ctx, M_in == M, ASGFEM_EINVAL, "evaluate_samples: M differs from the length of the multi-indices");
    ASG_CHECK(ctx, ctx->mis.maxdeg() < nvals, ASGFEM_EINVAL, "evaluate_samples: nvals <= maximal polynomial degree of the set");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    // H_k(xi_s) = prod_m vals[s][m][mu_k[m]]   (evaluate(TB, k), tensorizedbasis.jl:244-252: product over m ascending)
    const int64_t Spad = (nsamples + 7) / 8 * 8;
    std::vector<double> R((size_t)ctx->ld * (size_t)Spad, 0.0);  // row = device column of the mode
    for (int64_t s = 0; s < nsamples; ++s) {
        const double* v = vals + (size_t)s * (size_t)M * (size_t)nvals;
        for (int64_t k = 0; k < N; ++k) {
            double prod = 1.0;
            for (int64_t m = 0; m < M; ++m) prod *= v[m * nvals + ctx->mis.mi[k * M + m]];
            R[(size_t)ctx->h_pos[(size_t)k] * Spad + s] = prod;
        }
    }
    double *dR = nullptr, *dout = nullptr;
    cudaError_t e1 = cudaMalloc((void**)&dR, sizeof(double) * R.size());
    cudaError_t e2 = e1 == cudaSuccess ? cudaMalloc((void**)&dout, sizeof(double) * (size_t)n * (size_t)nsamples) : e1;
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
        if (dR) cudaFree(dR);
        (void)cudaGetLastError();
        return fail(ctx, ASGFEM_ENOMEM, "evaluate_samples: out of device memory for the sample block");
    }
    int rc = 0;
    cudaError_t e = cudaMemcpyAsync(dR, R.data(), sizeof(double) * R.size(), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) rc = vec_eval_samples(ctx, ctx->slots[slot_u], dR, nsamples, Spad, dout);
    if (e == cudaSuccess && !rc)
        e = cudaMemcpyAsync(out, dout, sizeof(double) * (size_t)n * (size_t)nsamples, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(dR);
    cudaFree(dout);
    if (rc) return rc;
    if (e != cudaSuccess) return fail(ctx, ASGFEM_ECUDA, std::string("evaluate_samples: ") + cudaGetErrorString(e));
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

// ---- estimator ----------------------------------------------------------------------------------
extern "C" int asgfem_estimate_poisson_primal(asgfem_ctx* ctx, int32_t slot_u, int64_t N_ext, int64_t M_ext,
                                              const int64_t* mi_ext, int32_t nq, const double* xref, const double* w,
                                              const double* f_at_qp, int32_t nqf, const double* sf, const double* wf,
                                              double* eta4cell, double* eta4modes) try {
    CTX_OR_FAIL(ctx);
    if (check_slot(ctx, slot_u)) return ASGFEM_EINVAL;
    ASG_CHECK(ctx, ctx->order > 0 && ctx->ncells > 0, ASGFEM_ESTATE, "estimate: set_mesh / set_space first");
    ASG_CHECK(ctx, ctx->ndofs_space == ctx->n, ASGFEM_ESTATE, "estimate: space and pattern sizes differ");
    ASG_CHECK(ctx, N_ext >= ctx->N && M_ext >= ctx->mis.M && mi_ext, ASGFEM_EINVAL, "estimate: bad extended multi-index set");
    ASG_CHECK(ctx, M_ext <= ctx->maxm, ASGFEM_EINVAL,
              "estimate: extended multi-indices longer than maxm of the coefficient (get_am! would be out of bounds)");
    ASG_CHECK(ctx, nq >= 1 && nq <= 64 && xref && w && nqf >= 1 && nqf <= 16 && sf && wf && eta4cell && eta4modes,
              ASGFEM_EINVAL, "estimate: bad quadrature / output arguments");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    return estimate_poisson_primal(ctx, ctx->slots[slot_u], N_ext, M_ext, mi_ext, nq, xref, w, f_at_qp, nqf, sf, wf,
                                   eta4cell, eta4modes);
}
ASG_BOUNDARY_CATCH(ctx)

// estimate(::Type{LogTransformedPoissonProblemPrimal}, ...) (src/estimate.jl:70-257)
extern "C" int asgfem_estimate_logpoisson_primal(asgfem_ctx* ctx, int32_t slot_u, int64_t N_ext, int64_t M_ext, const int64_t* mi_ext,
                                                 int32_t nq, const double* xref, const double* w, const double* f_at_qp,
                                                 const double* lam_at_qp, int32_t ntrunc, int32_t nqf, const double* sf,
                                                 const double* wf, double* eta4cell, double* eta4modes, double* zeta3) try {
    CTX_OR_FAIL(ctx);
    if (check_slot(ctx, slot_u)) return ASGFEM_EINVAL;
    ASG_CHECK(ctx, ctx->order > 0 && ctx->ncells > 0, ASGFEM_ESTATE, "estimate: set_mesh / set_space first");
    ASG_CHECK(ctx, ctx->ndofs_space == ctx->n, ASGFEM_ESTATE, "estimate: space and pattern sizes differ");
    ASG_CHECK(ctx, N_ext >= ctx->N && M_ext >= ctx->mis.M && mi_ext, ASGFEM_EINVAL, "estimate: bad extended multi-index set");
    ASG_CHECK(ctx, M_ext <= ctx->maxm && ntrunc >= 0 && ntrunc <= ctx->maxm, ASGFEM_EINVAL,
              "estimate: extended multi-indices / N_truncate exceed maxm of the coefficient");
    ASG_CHECK(ctx, nq >= 1 && nq <= 64 && xref && w && nqf >= 1 && nqf <= 16 && sf && wf && eta4cell && eta4modes && f_at_qp,
              ASGFEM_EINVAL, "estimate: bad quadrature / output arguments (the rhs values are required)");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    return estimate_poisson_primal(ctx, ctx->slots[slot_u], N_ext, M_ext, mi_ext, nq, xref, w, f_at_qp, nqf, sf, wf, eta4cell,
                                   eta4modes, -1, nullptr, nullptr, 1, lam_at_qp, ntrunc, zeta3);
}
ASG_BOUNDARY_CATCH(ctx)

// The outputs the adaptive loop consumes (scripts/poisson.jl:341-420): eta4modes and, for the spatial marking, the sum of
// eta4cell over a set of columns (the active modes) - without the D2H of the ncells x N_ext matrix
extern "C" int asgfem_estimate_poisson_primal_marking(asgfem_ctx* ctx, int32_t slot_u, int64_t N_ext, int64_t M_ext,
                                                      const int64_t* mi_ext, int32_t nq, const double* xref, const double* w,
                                                      const double* f_at_qp, int32_t nqf, const double* sf, const double* wf,
                                                      int64_t nsel, const int64_t* sel, double* cellsum, double* eta4modes) try {
    CTX_OR_FAIL(ctx);
    if (check_slot(ctx, slot_u)) return ASGFEM_EINVAL;
    ASG_CHECK(ctx, ctx->order > 0 && ctx->ncells > 0, ASGFEM_ESTATE, "estimate: set_mesh / set_space first");
    ASG_CHECK(ctx, ctx->ndofs_space == ctx->n, ASGFEM_ESTATE, "estimate: space and pattern sizes differ");
    ASG_CHECK(ctx, N_ext >= ctx->N && M_ext >= ctx->mis.M && mi_ext, ASGFEM_EINVAL, "estimate: bad extended multi-index set");
    ASG_CHECK(ctx, M_ext <= ctx->maxm, ASGFEM_EINVAL,
              "estimate: extended multi-indices longer than maxm of the coefficient (get_am! would be out of bounds)");
    ASG_CHECK(ctx, nq >= 1 && nq <= 64 && xref && w && nqf >= 1 && nqf <= 16 && sf && wf && eta4modes && cellsum && nsel >= 0 &&
                       (sel || nsel == 0),
              ASGFEM_EINVAL, "estimate: bad quadrature / output arguments");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    return estimate_poisson_primal(ctx, ctx->slots[slot_u], N_ext, M_ext, mi_ext, nq, xref, w, f_at_qp, nqf, sf, wf, nullptr,
                                   eta4modes, nsel, sel, cellsum);
}
ASG_BOUNDARY_CATCH(ctx)

// ---- multi-GPU: NCCL inside the library ------------------------------------------------------------
extern "C" int asgfem_comm_unique_id(void* id128) try {
    if (!id128) return ASGFEM_EINVAL;
    std::string err;
    int rc = dist_unique_id(id128, err);
    if (rc) g_create_error = err;
    return rc;
}
ASG_BOUNDARY_CATCH(nullptr)

extern "C" int asgfem_comm_init(asgfem_ctx* ctx, int32_t nranks, int32_t rank, const void* id128) try {
    CTX_OR_FAIL(ctx);
    ASG_CHECK(ctx, nranks >= 1 && rank >= 0 && rank < nranks && id128, ASGFEM_EINVAL, "comm_init: bad arguments");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    return dist_init(ctx, nranks, rank, id128);
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_comm_destroy(asgfem_ctx* ctx) try {
    CTX_OR_FAIL(ctx);
    if (set_device(ctx)) return ASGFEM_ECUDA;
    cudaStreamSynchronize(ctx->stream);
    dist_free(ctx);
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_set_halo(asgfem_ctx* ctx, int32_t nneigh, const int32_t* ranks, const int64_t* send_ptr,
                               const int64_t* send_rows, const int64_t* recv_ptr, const int64_t* recv_rows,
                               int64_t interior_row0, int64_t interior_row1) try {
    CTX_OR_FAIL(ctx);
    if (set_device(ctx)) return ASGFEM_ECUDA;
    return dist_set_halo(ctx, nneigh, ranks, send_ptr, send_rows, recv_ptr, recv_rows, interior_row0, interior_row1);
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_precond_setup_global(asgfem_ctx* ctx, int64_t n_global, const int64_t* colptr, const int64_t* rowval,
                                           const double* nzval, int64_t nb, const int64_t* bdofs, const double* coords,
                                           const int64_t* row_offsets) try {
    CTX_OR_FAIL(ctx);
    if (set_device(ctx)) return ASGFEM_ECUDA;
    return dist_precond_setup_global(ctx, n_global, colptr, rowval, nzval, nb, bdofs, coords, row_offsets);
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_vec_dot_global(asgfem_ctx* ctx, int32_t a, int32_t b, double* out) try {
    CTX_OR_FAIL(ctx);
    if (check_slot(ctx, a) || check_slot(ctx, b)) return ASGFEM_EINVAL;
    ASG_CHECK(ctx, out, ASGFEM_EINVAL, "null output");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    return dist_dot(ctx, ctx->slots[a], ctx->slots[b], out);
}
ASG_BOUNDARY_CATCH(ctx)

// ---- multi-GPU helpers --------------------------------------------------------------------------
extern "C" int asgfem_set_owned_rows(asgfem_ctx* ctx, int64_t n_owned) try {
    CTX_OR_FAIL(ctx);
    ASG_CHECK(ctx, ctx->n > 0 && n_owned >= 1 && n_owned <= ctx->n, ASGFEM_EINVAL, "set_owned_rows: out of range");
    ctx->n_owned = n_owned;
    apply_free_plan(ctx);
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_halo_exchange(asgfem_ctx* ctx, int32_t slot) try {
    CTX_OR_FAIL(ctx);
    if (check_slot(ctx, slot)) return ASGFEM_EINVAL;
    if (set_device(ctx)) return ASGFEM_ECUDA;
    int rc = dist_halo_exchange(ctx, ctx->slots[slot]);
    if (rc) return rc;
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_set_owned_cells(asgfem_ctx* ctx, int64_t ncells, const uint8_t* owned) try {
    CTX_OR_FAIL(ctx);
    if (!owned) {
        ctx->h_cell_owned.clear();
        return 0;
    }
    ASG_CHECK(ctx, ctx->ncells > 0 && ncells == ctx->ncells, ASGFEM_EINVAL, "set_owned_cells: one flag per cell of the mesh (asgfem_set_mesh first)");
    ctx->h_cell_owned.assign(owned, owned + ncells);
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_vec_device_ptr(asgfem_ctx* ctx, int32_t slot, void** dptr, int64_t* ld) try {
    CTX_OR_FAIL(ctx);
    if (check_slot(ctx, slot)) return ASGFEM_EINVAL;
    ASG_CHECK(ctx, dptr && ld, ASGFEM_EINVAL, "null output");
    *dptr = ctx->slots[slot];
    *ld = ctx->ld;
    return 0;
}
ASG_BOUNDARY_CATCH(ctx)

static int rows_to_device(asgfem_ctx* ctx, int64_t nrows, const int64_t* rows, int64_t** d_rows) {
    for (int64_t k = 0; k < nrows; ++k)
        ASG_CHECK(ctx, rows[k] >= 1 && rows[k] <= ctx->n, ASGFEM_EINVAL, "row id out of range");
    ASG_CUDA(ctx, cudaMalloc((void**)d_rows, sizeof(int64_t) * std::max<int64_t>(nrows, 1)));
    ASG_CUDA(ctx, cudaMemcpyAsync(*d_rows, rows, sizeof(int64_t) * nrows, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

extern "C" int asgfem_pack_rows(asgfem_ctx* ctx, int32_t slot, int64_t nrows, const int64_t* rows, void* dbuf) try {
    CTX_OR_FAIL(ctx);
    if (check_slot(ctx, slot)) return ASGFEM_EINVAL;
    ASG_CHECK(ctx, nrows >= 0 && (nrows == 0 || (rows && dbuf)), ASGFEM_EINVAL, "pack_rows: bad arguments");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    int64_t* d_rows = nullptr;
    int rc = rows_to_device(ctx, nrows, rows, &d_rows);
    if (!rc) rc = vec_pack_rows(ctx, ctx->slots[slot], nrows, d_rows, (double*)dbuf);
    cudaStreamSynchronize(ctx->stream);
    if (d_rows) cudaFree(d_rows);
    return rc;
}
ASG_BOUNDARY_CATCH(ctx)

extern "C" int asgfem_unpack_rows(asgfem_ctx* ctx, int32_t slot, int64_t nrows, const int64_t* rows, const void* dbuf) try {
    CTX_OR_FAIL(ctx);
    if (check_slot(ctx, slot)) return ASGFEM_EINVAL;
    ASG_CHECK(ctx, nrows >= 0 && (nrows == 0 || (rows && dbuf)), ASGFEM_EINVAL, "unpack_rows: bad arguments");
    if (set_device(ctx)) return ASGFEM_ECUDA;
    int64_t* d_rows = nullptr;
    int rc = rows_to_device(ctx, nrows, rows, &d_rows);
    if (!rc) rc = vec_unpack_rows(ctx, ctx->slots[slot], nrows, d_rows, (const double*)dbuf);
    cudaStreamSynchronize(ctx->stream);
    if (d_rows) cudaFree(d_rows);
    return rc;
}
ASG_BOUNDARY_CATCH(ctx)
