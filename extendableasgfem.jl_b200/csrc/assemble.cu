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

// KIND 0: K_m = (a_m grad phi_j, grad phi_i).  KIND 1 (log-transformed primal problem, logpoisson_primal.jl:95-105): plane 0 =
// the Laplacian A, plane m >= 1 = N_m[i,j] = - int (grad a_m . grad phi_j) phi_i (kernel get_gradam_x_sigma, factor -1).
template <int ORDER, int KIND>
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
            double lam[3] = {1.0 - xr - yr, xr, yr};
            double gix, giy, gjx, gjy;
            if (ORDER == 1) {
                gix = gl[i][0], giy = gl[i][1], gjx = gl[j][0], gjy = gl[j][1];
            } else {
                double di[3], dj[3];
                p2_dphi(lam, i, di);
                p2_dphi(lam, j, dj);
                gix = di[0] * gl[0][0] + di[1] * gl[1][0] + di[2] * gl[2][0];
                giy = di[0] * gl[0][1] + di[1] * gl[1][1] + di[2] * gl[2][1];
                gjx = dj[0] * gl[0][0] + dj[1] * gl[1][0] + dj[2] * gl[2][0];
                gjy = dj[0] * gl[0][1] + dj[1] * gl[1][1] + dj[2] * gl[2][1];
            }
            if (KIND == 0) {
                loc += c_w[q] * eval_am(m, px, py, mean, decay, b1, b2) * (gix * gjx + giy * gjy);
            } else if (m == 0) {
                loc += c_w[q] * (gix * gjx + giy * gjy);
            } else {
                double ax, ay;
                eval_gradam(m, px, py, decay, b1, b2, ax, ay);
                loc -= c_w[q] * basis_value<ORDER>(lam, i) * (ax * gjx + ay * gjy);
            }
        }
        acc += vol * loc;
    }
    vals[(int64_t)m * nnz + p] = acc;
}

}  // namespace

int assemble_stiffness(asgfem_ctx* ctx, int32_t M, int32_t nq, const double* xref, const double* w, int kind) {
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
#define ASG_LAUNCH_ASSEMBLE(O, K)                                                                                              \
    k_assemble<O, K><<<grid, 256, 0, ctx->stream>>>(nnz, nq, d_cptr, d_contrib, ctx->d_coords, ctx->d_cellnodes, ctx->mean, \
                                                    ctx->d_decay, ctx->d_b1, ctx->d_b2, ctx->d_vals)
    if (ctx->order == 1 && kind == 0) ASG_LAUNCH_ASSEMBLE(1, 0);
    else if (ctx->order == 1) ASG_LAUNCH_ASSEMBLE(1, 1);
    else if (kind == 0) ASG_LAUNCH_ASSEMBLE(2, 0);
    else ASG_LAUNCH_ASSEMBLE(2, 1);
#undef ASG_LAUNCH_ASSEMBLE
    cudaError_t e = cudaGetLastError();
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_cptr);
    cudaFree(d_contrib);
    if (e != cudaSuccess || e2 != cudaSuccess)
        return fail(ctx, ASGFEM_ECUDA, std::string("assemble_stiffness: ") + cudaGetErrorString(e != cudaSuccess ? e : e2));
    precond_free(ctx);
    return 0;
}

// ---- load vectors of the log-transformed primal problem: b[mu] = (lambda_mu f, phi_i)   (logpoisson_primal.jl:108-127) ------
namespace {
// tab[(cell * nq + q) * (Mmi + 1) + 0] = |T| w_q f(x_q) exp(1/2 sum_{m <= ntrunc} a_m(x_q)^2) exp(mean * factor),
//                                + d] = a_d(x_q), d = 1..Mmi   (expa_PCE_mop, src/coefficients/coefficients.jl:236-262)
__global__ void k_lograd_table(int64_t ncells, int nq, int Mmi, int ntrunc, double factor, const double* __restrict__ coords,
                               const int32_t* __restrict__ cellnodes, double mean, const double* __restrict__ decay,
                               const int32_t* __restrict__ b1, const int32_t* __restrict__ b2, const double* __restrict__ fq,
                               double* __restrict__ tab) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncells * nq) return;
    const int64_t cell = t / nq;
    const int q = (int)(t - cell * nq);
    const int32_t* cn = cellnodes + 3 * cell;
    const double x1 = coords[2 * cn[0]], y1 = coords[2 * cn[0] + 1], x2 = coords[2 * cn[1]], y2 = coords[2 * cn[1] + 1],
                 x3 = coords[2 * cn[2]], y3 = coords[2 * cn[2] + 1];
    const double vol = 0.5 * fabs((x2 - x1) * (y3 - y1) - (y2 - y1) * (x3 - x1));
    const double xr = c_xref[2 * q], yr = c_xref[2 * q + 1];
    const double px = x1 + xr * (x2 - x1) + yr * (x3 - x1), py = y1 + xr * (y2 - y1) + yr * (y3 - y1);
    double s = 0.0;
    for (int m = 1; m <= ntrunc; ++m) {
        const double a = eval_am(m, px, py, mean, decay, b1, b2);
        s += a * a;
        if (m <= Mmi) tab[t * (Mmi + 1) + m] = a;
    }
    tab[t * (Mmi + 1)] = vol * c_w[q] * fq[t] * (exp(s / 2) * exp(mean * factor));  // fq: nq x ncells column-major = index t
}

// thread = (dof i, mode k): cells of the dof in ascending order, quadrature points in order (deterministic)
template <int ORDER>
__global__ void k_lograd_rhs(int64_t n, int64_t N, int64_t ld, int nq, int Mmi, const int64_t* __restrict__ dptr,
                             const int32_t* __restrict__ dcontrib, const double* __restrict__ tab, const int32_t* __restrict__ mi,
                             const double* __restrict__ den, const int32_t* __restrict__ pos, double* __restrict__ b) {
    constexpr int ND = ORDER == 1 ? 3 : 6;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * N) return;
    const int64_t i = t / N;
    const int k = (int)(t - i * N);
    double acc = 0.0;
    for (int64_t c = dptr[i]; c < dptr[i + 1]; ++c) {
        const int32_t code = dcontrib[c];
        const int64_t cell = code / ND;
        const int li = code - (int32_t)cell * ND;
        for (int q = 0; q < nq; ++q) {
            const double xr = c_xref[2 * q], yr = c_xref[2 * q + 1];
            const double lam[3] = {1.0 - xr - yr, xr, yr};
            const double* row = tab + (cell * nq + q) * (Mmi + 1);
            double amu = 1.0;
            for (int d = 0; d < Mmi; ++d) {
                const int e = mi[(int64_t)k * Mmi + d];
                for (int r = 0; r < e; ++r) amu *= row[d + 1];
            }
            acc += basis_value<ORDER>(lam, li) * (amu / den[k] * row[0]);
        }
    }
    b[i * ld + pos[k]] = acc;
}
}  // namespace

int assemble_logprimal_rhs(asgfem_ctx* ctx, int32_t nq, const double* xref, const double* w, const double* f_at_qp, int32_t ntrunc,
                           double* dvec) {
    const int nd = ctx->ndofs4cell;
    const int64_t n = ctx->n, N = ctx->N, ncells = ctx->ncells;
    const int Mmi = (int)ctx->mis.M;
    ASG_CHECK(ctx, ncells * nd < (1ll << 31), ASGFEM_EINVAL, "assemble_logprimal_rhs: mesh too large");
    // cells of every dof, ascending in (cell, local index)
    std::vector<int64_t> dptr((size_t)n + 1, 0);
    for (int64_t k = 0; k < ncells * nd; ++k) dptr[(size_t)ctx->h_celldofs[(size_t)k] + 1]++;
    for (int64_t i = 0; i < n; ++i) dptr[(size_t)i + 1] += dptr[(size_t)i];
    std::vector<int32_t> dcontrib((size_t)dptr[(size_t)n]);
    {
        std::vector<int64_t> fill(dptr.begin(), dptr.end() - 1);
        for (int64_t k = 0; k < ncells * nd; ++k) dcontrib[(size_t)fill[(size_t)ctx->h_celldofs[(size_t)k]]++] = (int32_t)k;
    }
    // per mode: sqrt(prod mu_d!) * factor^|mu| with factor = -1   (coefficients.jl:258)
    std::vector<double> den((size_t)N);
    std::vector<int32_t> mi32((size_t)(N * Mmi));
    for (int64_t k = 0; k < N; ++k) {
        double fac = 1.0;
        int64_t deg = 0;
        for (int d = 0; d < Mmi; ++d) {
            const int64_t e = ctx->mis.mi[(size_t)(k * Mmi + d)];
            mi32[(size_t)(k * Mmi + d)] = (int32_t)e;
            for (int64_t r = 2; r <= e; ++r) fac *= (double)r;
            deg += e;
        }
        den[(size_t)k] = std::sqrt(fac) * ((deg & 1) ? -1.0 : 1.0);
    }
    int64_t* d_dptr = nullptr;
    int32_t *d_dc = nullptr, *d_mi = nullptr;
    double *d_den = nullptr, *d_fq = nullptr, *d_tab = nullptr;
    auto cleanup = [&]() {
        for (void* p : {(void*)d_dptr, (void*)d_dc, (void*)d_mi, (void*)d_den, (void*)d_fq, (void*)d_tab})
            if (p) cudaFree(p);
    };
    int rc = dev_upload(ctx, &d_dptr, dptr);
    rc |= dev_upload(ctx, &d_dc, dcontrib);
    rc |= dev_upload(ctx, &d_mi, mi32);
    rc |= dev_upload(ctx, &d_den, den);
    if (!rc) {
        std::vector<double> fq(f_at_qp, f_at_qp + (size_t)nq * ncells);
        rc = dev_upload(ctx, &d_fq, fq);
        if (!rc) cudaStreamSynchronize(ctx->stream);  // fq goes out of scope
    }
    if (rc) {
        cleanup();
        return rc;
    }
    cudaError_t e = cudaMalloc((void**)&d_tab, sizeof(double) * (size_t)ncells * nq * (Mmi + 1));
    if (e != cudaSuccess) {
        cleanup();
        return fail(ctx, ASGFEM_ENOMEM, "assemble_logprimal_rhs: table allocation failed");
    }
    cudaMemcpyToSymbolAsync(c_xref, xref, sizeof(double) * 2 * nq, 0, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyToSymbolAsync(c_w, w, sizeof(double) * nq, 0, cudaMemcpyHostToDevice, ctx->stream);
    k_lograd_table<<<(unsigned)((ncells * nq + 255) / 256), 256, 0, ctx->stream>>>(ncells, nq, Mmi, ntrunc, -1.0, ctx->d_coords,
                                                                                 ctx->d_cellnodes, ctx->mean, ctx->d_decay, ctx->d_b1,
                                                                                 ctx->d_b2, d_fq, d_tab);
    cudaMemsetAsync(dvec, 0, sizeof(double) * (size_t)n * (size_t)ctx->ld, ctx->stream);
    const unsigned blocks = (unsigned)((n * N + 255) / 256);
    if (ctx->order == 1)
        k_lograd_rhs<1><<<blocks, 256, 0, ctx->stream>>>(n, N, ctx->ld, nq, Mmi, d_dptr, d_dc, d_tab, d_mi, d_den, ctx->d_pos, dvec);
    else
        k_lograd_rhs<2><<<blocks, 256, 0, ctx->stream>>>(n, N, ctx->ld, nq, Mmi, d_dptr, d_dc, d_tab, d_mi, d_den, ctx->d_pos, dvec);
    e = cudaGetLastError();
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    cleanup();
    if (e != cudaSuccess || e2 != cudaSuccess)
        return fail(ctx, ASGFEM_ECUDA, std::string("assemble_logprimal_rhs: ") + cudaGetErrorString(e != cudaSuccess ? e : e2));
    return 0;
}

}  // namespace asgfem
