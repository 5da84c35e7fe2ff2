// Fused SGFE operator  Y = sum_m (G_m (x) K_m) X  with Dirichlet rows zeroed
// (LinearAlgebra.mul!(Ax, S::MySystemPrimal, x), src/modelproblems/solvers_poisson_primal.jl:86-124).
//
// Three hand-written kernels for sm_100a:
//
//  variant 7  k_apply_ts2      (apply_ts2.cu) packed mode-stationary DFMA kernel: lanes own modes, X rows of the dof row in
//             registers; the fastest kernel where its plan fits (N <= 2048 modes, M <= 63, short rows).
//  variant 9  k_apply_blk      (apply_blk.cu) block products on the fp64 MMA path (DMMA m8n8k4) with a list exchange:
//             any N (passes over the mode blocks), rows of up to 24 entries (P2), any M <= 255.
//  variant 1  k_apply_gather   thread = (dof row i, device column c); walks the neighbour list of the column's mode in the
//             reference's accumulation order (nu ascending, direction ascending) and reads X and K_m through L1/L2.
//             Simple and exact in ordering: the in-library cross-check of the other kernels and the last fallback.
//
// No kernel touches uncoupled (m, nu) pairs of the reference loop.  All kernels work in DEVICE COLUMN space: column c of
// every vector holds mode ctx->h_inv[c] (padding columns: none, kept at zero); the coupling lists uploaded by
// asgfem_set_multiindices are relabelled accordingly.  Retired kernels and their measurements: profiles/README.md.
#include <algorithm>

#include "common.h"

namespace asgfem {

__global__ void __launch_bounds__(256)
k_apply_gather(int64_t row0, int64_t nrows, int64_t ld, int64_t nnz, const int64_t* __restrict__ rowptr,
               const int32_t* __restrict__ col, const double* __restrict__ vals, const int32_t* __restrict__ cptr,
               const int32_t* __restrict__ cm, const int32_t* __restrict__ cnu, const double* __restrict__ cg,
               const uint8_t* __restrict__ bmask, const double* __restrict__ x, double* __restrict__ y) {
    int64_t i = row0 + (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    int64_t c = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    if (i >= nrows || c >= ld) return;
    double acc = 0.0;
    if (!bmask[i]) {
        int64_t p0 = rowptr[i], p1 = rowptr[i + 1];
        for (int64_t p = p0; p < p1; ++p) acc = fma(vals[p], x[(int64_t)col[p] * ld + c], acc);
        for (int32_t e = cptr[c]; e < cptr[c + 1]; ++e) {
            const double* vm = vals + (int64_t)cm[e] * nnz;
            int64_t nu = cnu[e];
            double t = 0.0;
            for (int64_t p = p0; p < p1; ++p) t = fma(vm[p], x[(int64_t)col[p] * ld + nu], t);
            acc = fma(cg[e], t, acc);
        }
    }
    y[i * ld + c] = acc;
}

void apply_free_plan(asgfem_ctx* ctx) {
    ctx->apply_ready = false;
    apply_ts2_free(ctx);
}

int apply_build_plan(asgfem_ctx* ctx) {
    int rc = apply_blk_build(ctx);
    if (rc) return rc;
    rc = apply_ts2_build(ctx);
    if (rc) return rc;
    ctx->apply_ready = true;
    return 0;
}

int apply_launch(asgfem_ctx* ctx, const double* x, double* y, int64_t r0, int64_t r1) {
    const int64_t nrows_all = ctx->n_owned >= 0 ? ctx->n_owned : ctx->n;
    if (r1 < 0) {
        r0 = 0;
        r1 = nrows_all;
    }
    r1 = std::min(r1, nrows_all);
    int variant = ctx->apply_variant;
    // automatic choice by measurement (config 4, B200): packed mode-stationary DFMA kernel 51 ms, block kernel 55 ms
    if (variant == 0) variant = apply_ts2_preferred(ctx) ? 7 : apply_blk_usable(ctx) ? 9 : 1;
    if (r1 <= r0) return 0;
    ASG_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    if (variant == 9) {
        int rc = apply_blk_launch(ctx, x, y, r0, r1);
        if (rc) return rc;
    } else if (variant == 7) {
        int rc = apply_ts2_launch(ctx, x, y, r0, r1);
        if (rc) return rc;
    } else {
        int bx = (int)std::min<int64_t>(128, ((ctx->ld + 31) / 32) * 32);
        int by = 256 / bx;
        dim3 block(bx, by);
        dim3 grid((unsigned)((r1 - r0 + by - 1) / by), (unsigned)((ctx->ld + bx - 1) / bx));
        k_apply_gather<<<grid, block, 0, ctx->stream>>>(r0, r1, ctx->ld, ctx->nnz, ctx->d_rowptr, ctx->d_col, ctx->d_vals,
                                                        ctx->d_cptr, ctx->d_cm, ctx->d_cnu, ctx->d_cg, ctx->d_bmask, x, y);
        ASG_CUDA(ctx, cudaGetLastError());
    }
    ASG_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    return 0;
}

}  // namespace asgfem
