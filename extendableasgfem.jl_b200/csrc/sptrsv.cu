// Mean-based preconditioner  Z[:,mu] = K_0^{-1} R[:,mu]  for all modes at once
// (LinearAlgebra.ldiv!(y, P::MyPreconditionerPrimal, b), src/modelproblems/solvers_poisson_primal.jl:46-78;
//  the reference runs N sequential UMFPACK solves, one per mode block).
//
// Factor P K_0 P^T = L L^T from chol.cpp.  The N right-hand sides are independent, so the triangular solves
// need no inter-CTA synchronisation at all: every CTA owns a tile of MT modes and sweeps the level-scheduled
// rows of L (forward) and L^T (backward) by itself, with one __syncthreads per level.  Rows of a level are
// spread over the RS row slots of the CTA; when a level has fewer rows than slots (the dense separator chains
// at the top of the elimination tree) the nonzeros of a row are split over several slots and reduced in shared
// memory.  Work vectors live in elimination order (W[k,:] <-> dof perm[k]), which keeps subtrees contiguous.
#include <algorithm>

#include "common.h"

namespace asgfem {

constexpr int MT = 16;                 // modes per CTA (half a warp wide)
constexpr int RS = 16;                 // row slots per CTA
constexpr int TRSV_THREADS = MT * RS;  // 256

struct PrecondPlan {
    int64_t nred = 0;
    int32_t* d_perm = nullptr;
    double* d_dinv = nullptr;
    // forward: rows of L; backward: rows of L^T (= columns of L)
    int64_t *d_fptr = nullptr, *d_bptr = nullptr;
    int32_t *d_fidx = nullptr, *d_bidx = nullptr;
    double *d_fval = nullptr, *d_bval = nullptr;
    int32_t nflev = 0, nblev = 0;
    int32_t *d_flevptr = nullptr, *d_flevrows = nullptr, *d_blevptr = nullptr, *d_blevrows = nullptr;
    double* d_work = nullptr;  // nred x ld
    int64_t lnz = 0;
};

void precond_free(asgfem_ctx* ctx) {
    PrecondPlan* P = ctx->precond;
    if (!P) return;
    void* ptrs[] = {P->d_perm, P->d_dinv, P->d_fptr, P->d_bptr, P->d_fidx, P->d_bidx, P->d_fval, P->d_bval,
                    P->d_flevptr, P->d_flevrows, P->d_blevptr, P->d_blevrows, P->d_work};
    for (void* q : ptrs)
        if (q) cudaFree(q);
    delete P;
    ctx->precond = nullptr;
}

namespace {

void level_schedule(int64_t n, const std::vector<int64_t>& ptr, const std::vector<int32_t>& idx, bool reverse,
                    std::vector<int32_t>& levptr, std::vector<int32_t>& levrows) {
    std::vector<int32_t> level((size_t)n, 0);
    int32_t nlev = 0;
    if (!reverse) {
        for (int64_t k = 0; k < n; ++k) {
            int32_t l = 0;
            for (int64_t p = ptr[k]; p < ptr[k + 1]; ++p) l = std::max(l, level[idx[p]] + 1);
            level[k] = l;
            nlev = std::max(nlev, l + 1);
        }
    } else {
        for (int64_t k = n - 1; k >= 0; --k) {
            int32_t l = 0;
            for (int64_t p = ptr[k]; p < ptr[k + 1]; ++p) l = std::max(l, level[idx[p]] + 1);
            level[k] = l;
            nlev = std::max(nlev, l + 1);
        }
    }
    levptr.assign((size_t)nlev + 1, 0);
    for (int64_t k = 0; k < n; ++k) levptr[level[k] + 1]++;
    for (int32_t l = 0; l < nlev; ++l) levptr[l + 1] += levptr[l];
    levrows.resize((size_t)n);
    std::vector<int32_t> fill(levptr.begin(), levptr.end() - 1);
    for (int64_t k = 0; k < n; ++k) levrows[fill[level[k]]++] = (int32_t)k;
}

// W[k, :] = R[perm[k], :]   (gather into elimination order)
__global__ void k_gather_perm(const double* __restrict__ r, double* __restrict__ w, const int32_t* __restrict__ perm,
                              int64_t nred, int64_t ld) {
    int64_t total = nred * (ld / 2);
    const double2* r2 = reinterpret_cast<const double2*>(r);
    double2* w2 = reinterpret_cast<double2*>(w);
    int64_t h = ld / 2;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t k = t / h, c = t - k * h;
        w2[t] = r2[(int64_t)perm[k] * h + c];
    }
}

// Z[perm[k], :] = W[k, :]; boundary rows of Z are zeroed beforehand
__global__ void k_scatter_perm(const double* __restrict__ w, double* __restrict__ z, const int32_t* __restrict__ perm,
                               int64_t nred, int64_t ld) {
    int64_t total = nred * (ld / 2);
    const double2* w2 = reinterpret_cast<const double2*>(w);
    double2* z2 = reinterpret_cast<double2*>(z);
    int64_t h = ld / 2;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t k = t / h, c = t - k * h;
        z2[(int64_t)perm[k] * h + c] = w2[t];
    }
}

__global__ void k_zero_masked_rows(double* __restrict__ z, const uint8_t* __restrict__ bmask, int64_t n, int64_t ld) {
    for (int64_t i = blockIdx.x; i < n; i += gridDim.x) {
        if (!bmask[i]) continue;
        for (int64_t k = threadIdx.x; k < ld; k += blockDim.x) z[i * ld + k] = 0.0;
    }
}

// One triangular sweep over all levels for the mode tile of this CTA.
//   w[k,:] <- (w[k,:] - sum_p val[p] * w[idx[p],:]) * dinv[k]       rows k in level order
__global__ void __launch_bounds__(TRSV_THREADS)
k_trsv_sweep(double* __restrict__ w, int64_t ld, const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
             const double* __restrict__ val, const double* __restrict__ dinv, int32_t nlev,
             const int32_t* __restrict__ levptr, const int32_t* __restrict__ levrows) {
    __shared__ double red[RS][MT + 1];
    const int lane = threadIdx.x % MT;   // mode within the tile
    const int slot = threadIdx.x / MT;   // row slot
    const int64_t mode = (int64_t)blockIdx.x * MT + lane;
    if ((int64_t)blockIdx.x * MT >= ld) return;
    for (int32_t l = 0; l < nlev; ++l) {
        const int32_t r0 = levptr[l], r1 = levptr[l + 1];
        const int32_t nr = r1 - r0;
        if (nr > RS / 2) {
            // many rows: one slot per row
            for (int32_t q = r0 + slot; q < r1; q += RS) {
                const int32_t k = levrows[q];
                double acc = w[(int64_t)k * ld + mode];
                const int64_t p1 = ptr[k + 1];
#pragma unroll 4
                for (int64_t p = ptr[k]; p < p1; ++p) acc = fma(-val[p], w[(int64_t)idx[p] * ld + mode], acc);
                w[(int64_t)k * ld + mode] = acc * dinv[k];
            }
        } else {
            // few rows: split every row over `per` slots and reduce
            int per = RS;
            while (per > 1 && per * nr > RS) per >>= 1;
            per = max(per, 1);
            const int rloc = slot / per, part = slot % per;
            for (int32_t base = r0; base < r1; base += RS / per) {
                const int32_t q = base + rloc;
                double acc = 0.0;
                int32_t k = -1;
                if (q < r1 && rloc < RS / per) {
                    k = levrows[q];
                    const int64_t p1 = ptr[k + 1];
#pragma unroll 4
                    for (int64_t p = ptr[k] + part; p < p1; p += per) acc = fma(-val[p], w[(int64_t)idx[p] * ld + mode], acc);
                }
                red[slot][lane] = acc;
                __syncthreads();
                if (k >= 0 && part == 0) {
                    double s = w[(int64_t)k * ld + mode];
                    for (int j = 0; j < per; ++j) s += red[rloc * per + j][lane];
                    w[(int64_t)k * ld + mode] = s * dinv[k];
                }
                __syncthreads();
            }
        }
        __syncthreads();
    }
}

}  // namespace

int precond_setup(asgfem_ctx* ctx) {
    precond_free(ctx);
    ASG_CHECK(ctx, ctx->N > 0, ASGFEM_ESTATE, "precond_setup: multi-indices not set");
    // download K_0 (device holds the authoritative copy, e.g. after device assembly)
    std::vector<double> k0((size_t)ctx->nnz);
    ASG_CUDA(ctx, cudaMemcpyAsync(k0.data(), ctx->d_vals, sizeof(double) * ctx->nnz, cudaMemcpyDeviceToHost, ctx->stream));
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    CholFactor F;
    std::string err;
    int rc = cholesky_reduced(ctx->n, ctx->h_rowptr.data(), ctx->h_col.data(), k0.data(), ctx->h_bmask.data(), F, err);
    if (rc) return fail(ctx, rc, "precond_setup: " + err);
    PrecondPlan* P = new PrecondPlan();
    ctx->precond = P;
    P->nred = F.n;
    P->lnz = (int64_t)F.Li.size();
    // backward structure: rows of L^T = columns of L
    std::vector<int64_t> bptr((size_t)F.n + 1, 0);
    for (int32_t j : F.Li) bptr[j + 1]++;
    for (int64_t k = 0; k < F.n; ++k) bptr[k + 1] += bptr[k];
    std::vector<int32_t> bidx(F.Li.size());
    std::vector<double> bval(F.Li.size());
    {
        std::vector<int64_t> fill(bptr.begin(), bptr.end() - 1);
        for (int64_t k = 0; k < F.n; ++k)
            for (int64_t p = F.Lp[k]; p < F.Lp[k + 1]; ++p) {
                int64_t at = fill[F.Li[p]]++;
                bidx[at] = (int32_t)k;
                bval[at] = F.Lx[p];
            }
    }
    std::vector<int32_t> flp, flr, blp, blr;
    level_schedule(F.n, F.Lp, F.Li, false, flp, flr);
    level_schedule(F.n, bptr, bidx, true, blp, blr);
    P->nflev = (int32_t)flp.size() - 1;
    P->nblev = (int32_t)blp.size() - 1;
    rc = 0;
    rc |= dev_upload(ctx, &P->d_perm, F.perm);
    rc |= dev_upload(ctx, &P->d_dinv, F.dinv);
    rc |= dev_upload(ctx, &P->d_fptr, F.Lp);
    rc |= dev_upload(ctx, &P->d_fidx, F.Li);
    rc |= dev_upload(ctx, &P->d_fval, F.Lx);
    rc |= dev_upload(ctx, &P->d_bptr, bptr);
    rc |= dev_upload(ctx, &P->d_bidx, bidx);
    rc |= dev_upload(ctx, &P->d_bval, bval);
    rc |= dev_upload(ctx, &P->d_flevptr, flp);
    rc |= dev_upload(ctx, &P->d_flevrows, flr);
    rc |= dev_upload(ctx, &P->d_blevptr, blp);
    rc |= dev_upload(ctx, &P->d_blevrows, blr);
    if (rc) return rc;
    ASG_CUDA(ctx, cudaMalloc((void**)&P->d_work, sizeof(double) * (size_t)std::max<int64_t>(F.n, 1) * ctx->ld));
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int precond_apply(asgfem_ctx* ctx, const double* r, double* z) {
    PrecondPlan* P = ctx->precond;
    ASG_CHECK(ctx, P, ASGFEM_ESTATE, "precond_apply: setup missing");
    const int64_t ld = ctx->ld;
    int blocks = (int)std::min<int64_t>(148 * 8, std::max<int64_t>(1, (P->nred * (ld / 2) + 255) / 256));
    if (P->nred > 0) {
        k_gather_perm<<<blocks, 256, 0, ctx->stream>>>(r, P->d_work, P->d_perm, P->nred, ld);
        int tiles = (int)((ctx->N + MT - 1) / MT);
        k_trsv_sweep<<<tiles, TRSV_THREADS, 0, ctx->stream>>>(P->d_work, ld, P->d_fptr, P->d_fidx, P->d_fval, P->d_dinv,
                                                              P->nflev, P->d_flevptr, P->d_flevrows);
        k_trsv_sweep<<<tiles, TRSV_THREADS, 0, ctx->stream>>>(P->d_work, ld, P->d_bptr, P->d_bidx, P->d_bval, P->d_dinv,
                                                              P->nblev, P->d_blevptr, P->d_blevrows);
    }
    // z may alias r: boundary rows are zeroed first, interior rows are overwritten from the work vector
    k_zero_masked_rows<<<(unsigned)std::min<int64_t>(ctx->n, 148 * 8), 128, 0, ctx->stream>>>(z, ctx->d_bmask, ctx->n, ld);
    if (P->nred > 0) k_scatter_perm<<<blocks, 256, 0, ctx->stream>>>(P->d_work, z, P->d_perm, P->nred, ld);
    ASG_CUDA(ctx, cudaGetLastError());
    return 0;
}

}  // namespace asgfem
