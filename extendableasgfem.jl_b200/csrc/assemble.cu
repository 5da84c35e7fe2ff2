// Device assembly of the KLE stiffness matrices K_0..K_M into one shared CSR pattern with per-m value planes
// (BASELINE.json north_star item 2).  Replaces the M+1 separate ExtendableFEM assemblies of
// src/modelproblems/poisson_primal.jl:56-63 whose kernel is get_am_x(m, C) (src/coefficients/coefficients.jl:147-153)
// with a_m from get_am! (src/coefficients/cosinus.jl:58-65):
//     K_m[i,j] = sum_T |T| sum_q w_q a_m(x_q) grad(phi_j)(x_q) . grad(phi_i)(x_q)
//
// Deterministic gather formulation: thread = (nonzero p, mode m); it sums the contributions of the cells that
// contain both dofs in ascending cell order - the order of a serial cell loop - so results do not depend on
// scheduling (no floating-point atomics).
#include <algorithm>

#include "common.h"
#include "coeff.cuh"

namespace asgfem {

namespace {

constexpr int MAXQ = 64;
__constant__ double c_xref[2 * MAXQ];
__constant__ double c_w[MAXQ];

template <int ORDER>
__global__ void __launch_bounds__(256)
k_assemble(int64_t nnz, int nq, const int64_t* __restrict__ cptr, const int32_t* __restrict__ contrib,
           const double* __restrict__ coords, const int32_t* __restrict__ cellnodes, double mean,
           const double* __restrict__ decay, const int32_t* __restrict__ b1, const int32_t* __restrict__ b2,
           double* __restrict__ vals) {
    constexpr int ND = ORDER == 1 ? 3 : 6;
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int m = blockIdx.y;
    if (p >= nnz) return;
    double acc = 0.0;
    for (int64_t c = cptr[p]; c < cptr[p + 1]; ++c) {
        int32_t code = contrib[c];
        int32_t cell = code / (ND * ND);
        int ij = code - cell * (ND * ND);
        int i = ij / ND, j = ij - i * ND;
        const int32_t* cn = cellnodes + 3 * (int64_t)cell;
        double x1 = coords[2 * cn[0]], y1 = coords[2 * cn[0] + 1];
        double x2 = coords[2 * cn[1]], y2 = coords[2 * cn[1] + 1];
        double x3 = coords[2 * cn[2]], y3 = coords[2 * cn[2] + 1];
        double det = (x2 - x1) * (y3 - y1) - (y2 - y1) * (x3 - x1);
        double vol = 0.5 * fabs(det);
        double gl[3][2] = {{(y2 - y3) / det, (x3 - x2) / det}, {(y3 - y1) / det, (x1 - x3) / det},
                           {(y1 - y2) / det, (x2 - x1) / det}};
        double loc = 0.0;
        for (int q = 0; q < nq; ++q) {
            double xr = c_xref[2 * q], yr = c_xref[2 * q + 1];
            double px = x1 + xr * (x2 - x1) + yr * (x3 - x1);
            double py = y1 + xr * (y2 - y1) + yr * (y3 - y1);
            double s;
            if (ORDER == 1) {
                s = gl[i][0] * gl[j][0] + gl[i][1] * gl[j][1];
            } else {
                double lam[3] = {1.0 - xr - yr, xr, yr};
                double di[3], dj[3];
                p2_dphi(lam, i, di);
                p2_dphi(lam, j, dj);
                double gix = di[0] * gl[0][0] + di[1] * gl[1][0] + di[2] * gl[2][0];
                double giy = di[0] * gl[0][1] + di[1] * gl[1][1] + di[2] * gl[2][1];
                double gjx = dj[0] * gl[0][0] + dj[1] * gl[1][0] + dj[2] * gl[2][0];
                double gjy = dj[0] * gl[0][1] + dj[1] * gl[1][1] + dj[2] * gl[2][1];
                s = gix * gjx + giy * gjy;
            }
            loc += c_w[q] * eval_am(m, px, py, mean, decay, b1, b2) * s;
        }
        acc += vol * loc;
    }
    vals[(int64_t)m * nnz + p] = acc;
}

}  // namespace

int assemble_stiffness(asgfem_ctx* ctx, int32_t M, int32_t nq, const double* xref, const double* w) {
    const int nd = ctx->ndofs4cell;
    const int64_t nnz = ctx->nnz, ncells = ctx->ncells;
    ASG_CHECK(ctx, ncells * nd * nd < (1ll << 31), ASGFEM_EINVAL, "assemble_stiffness: mesh too large for 32-bit contribution codes");
    // contributions per nonzero, ascending in (cell, i, j)
    std::vector<int64_t> cptr((size_t)nnz + 1, 0);
    auto find = [&](int32_t r, int32_t c) -> int64_t {
        const int32_t* b = ctx->h_col.data() + ctx->h_rowptr[r];
        const int32_t* e = ctx->h_col.data() + ctx->h_rowptr[r + 1];
        const int32_t* it = std::lower_bound(b, e, c);
        return (it != e && *it == c) ? (int64_t)(it - ctx->h_col.data()) : -1;
    };
    std::vector<int64_t> pos((size_t)(ncells * nd * nd));
    for (int64_t cell = 0; cell < ncells; ++cell)
        for (int i = 0; i < nd; ++i)
            for (int j = 0; j < nd; ++j) {
                int64_t p = find(ctx->h_celldofs[cell * nd + i], ctx->h_celldofs[cell * nd + j]);
                ASG_CHECK(ctx, p >= 0, ASGFEM_EINVAL, "assemble_stiffness: pattern does not contain a cell coupling");
                pos[(cell * nd + i) * nd + j] = p;
                cptr[p + 1]++;
            }
    for (int64_t p = 0; p < nnz; ++p) cptr[p + 1] += cptr[p];
    std::vector<int32_t> contrib((size_t)cptr[nnz]);
    {
        std::vector<int64_t> fill(cptr.begin(), cptr.end() - 1);
        for (int64_t k = 0; k < (int64_t)pos.size(); ++k) contrib[fill[pos[k]]++] = (int32_t)k;  // k = cell*nd*nd + i*nd + j
    }
    int64_t* d_cptr = nullptr;
    int32_t* d_contrib = nullptr;
    int rc = dev_upload(ctx, &d_cptr, cptr);
    rc |= dev_upload(ctx, &d_contrib, contrib);
    if (rc) {
        if (d_cptr) cudaFree(d_cptr);
        if (d_contrib) cudaFree(d_contrib);
        return rc;
    }
    cudaMemcpyToSymbolAsync(c_xref, xref, sizeof(double) * 2 * nq, 0, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyToSymbolAsync(c_w, w, sizeof(double) * nq, 0, cudaMemcpyHostToDevice, ctx->stream);
    dim3 grid((unsigned)((nnz + 255) / 256), (unsigned)(M + 1));
    if (ctx->order == 1)
        k_assemble<1><<<grid, 256, 0, ctx->stream>>>(nnz, nq, d_cptr, d_contrib, ctx->d_coords, ctx->d_cellnodes, ctx->mean,
                                                     ctx->d_decay, ctx->d_b1, ctx->d_b2, ctx->d_vals);
    else
        k_assemble<2><<<grid, 256, 0, ctx->stream>>>(nnz, nq, d_cptr, d_contrib, ctx->d_coords, ctx->d_cellnodes, ctx->mean,
                                                     ctx->d_decay, ctx->d_b1, ctx->d_b2, ctx->d_vals);
    cudaError_t e = cudaGetLastError();
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_cptr);
    cudaFree(d_contrib);
    if (e != cudaSuccess || e2 != cudaSuccess)
        return fail(ctx, ASGFEM_ECUDA, std::string("assemble_stiffness: ") + cudaGetErrorString(e != cudaSuccess ? e : e2));
    precond_free(ctx);
    return 0;
}

}  // namespace asgfem
