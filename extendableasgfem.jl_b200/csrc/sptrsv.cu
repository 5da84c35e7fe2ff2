// Mean-based preconditioner  Z[:,mu] = K_0^{-1} R[:,mu]  for all modes at once
// (LinearAlgebra.ldiv!(y, P::MyPreconditionerPrimal, b), src/modelproblems/solvers_poisson_primal.jl:46-78;
//  the reference runs N sequential UMFPACK solves, one per mode block).
//
// Factor P K_0 P^T = L L^T from chol.cpp.  The N right-hand sides are independent, so the triangular solves need
// no inter-CTA synchronisation at all: every CTA owns a tile of MT modes and performs the whole sweep by itself.
//
// Blocked right-looking sweep (k_trsv_blocked).  The unknowns are cut into column blocks of BW consecutive rows of
// the elimination order (nested dissection keeps subtrees and separators contiguous).  For block J the CTA
//   1. stages W[J, tile] in shared memory,
//   2. solves the diagonal block L[J,J] there, level by level (rows of a level spread over the half-warps; the
//      sequential rows of dense separator blocks are split over all half-warps and reduced),
//   3. writes the finished rows back, and
//   4. pushes W[r, tile] -= L[r, J] * Z_J to every later row r with entries in J: one half-warp per (row, block)
//      segment of the CSR row, the sources Z_J come from shared memory (each is reused by all rows below), the
//      target row is read and written once per segment instead of once per nonzero.
// The backward sweep L^T z = y is the same algorithm on the reversed numbering k -> n-1-k (rows of L^T reversed).
// Work vectors live in elimination order (W[k,:] <-> dof perm[k]).
#include <algorithm>

#include "common.h"

namespace asgfem {

constexpr int MT = 16;                 // modes per CTA (half a warp wide)
constexpr int RS = 64;                 // half-warps (row slots) per CTA
constexpr int TRSV_THREADS = MT * RS;  // 1024
constexpr int BW = 256;                // max rows per column block
constexpr int DE_MAX = 6144;           // diagonal-block entries staged in shared memory (else streamed from L2)
constexpr int PANEL = 16;              // rows per panel in dense (separator) blocks

struct TriDev {  // one triangular system in its own (forward) numbering
    // off-diagonal-block part: CSR of the strict lower triangle; only the entries outside a row's own block are used
    int32_t* idx = nullptr;
    double* val = nullptr;
    int32_t* segptr = nullptr;     // per block: range of push segments
    int32_t* seg = nullptr;        // per segment: target row, first nonzero, length
    // diagonal-block part, stored compactly block after block (block-local 16-bit column ids)
    double* dinv = nullptr;        // per row
    int32_t* drow = nullptr;       // per row: offset of its diagonal-block entries (n+1 entries)
    int32_t* dsplit = nullptr;     // per row: number of those entries left of the row's 16-row panel
    uint16_t* didx = nullptr;
    double* dval = nullptr;
    int32_t* blk_start = nullptr;  // nblocks+1 block boundaries (aligned with the dissection tree)
    int32_t* blk_info = nullptr;   // per block: offset into levptr, number of levels, dense flag, (pad)
    int32_t* levptr = nullptr;     // per level: offset into levrows
    uint16_t* levrows = nullptr;   // rows (block-local ids) sorted by level
    int nblocks = 0;
};

struct PrecondPlan {
    int64_t nred = 0;
    int32_t* d_perm = nullptr;
    TriDev fwd, bwd;
    double* d_work = nullptr;  // nred x ld
    int64_t lnz = 0;
};

static void free_tri(TriDev& T) {
    void* ptrs[] = {T.idx, T.val, T.segptr, T.seg, T.dinv, T.drow, T.dsplit, T.didx, T.dval, T.blk_start, T.blk_info, T.levptr, T.levrows};
    for (void* q : ptrs)
        if (q) cudaFree(q);
    T = TriDev();
}

void precond_free(asgfem_ctx* ctx) {
    PrecondPlan* P = ctx->precond;
    if (!P) return;
    free_tri(P->fwd);
    free_tri(P->bwd);
    if (P->d_perm) cudaFree(P->d_perm);
    if (P->d_work) cudaFree(P->d_work);
    delete P;
    ctx->precond = nullptr;
}

namespace {

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

// sum_e val[e] * Zs[idx[e]][lane] over e = e0, e0+step, ... < e1 with four independent accumulators (ILP)
template <class VT, class IT>
__device__ __forceinline__ double dot_rows(const VT* __restrict__ val, const IT* __restrict__ idx, int e0, int e1, int step,
                                           const double (*Zs)[MT], int lane) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int e = e0;
    for (; e + 3 * step < e1; e += 4 * step) {
        const int i0 = idx[e], i1 = idx[e + step], i2 = idx[e + 2 * step], i3 = idx[e + 3 * step];
        const double v0 = val[e], v1 = val[e + step], v2 = val[e + 2 * step], v3 = val[e + 3 * step];
        a0 = fma(v0, Zs[i0][lane], a0);
        a1 = fma(v1, Zs[i1][lane], a1);
        a2 = fma(v2, Zs[i2][lane], a2);
        a3 = fma(v3, Zs[i3][lane], a3);
    }
    for (; e < e1; e += step) a0 = fma(val[e], Zs[idx[e]][lane], a0);
    return (a0 + a1) + (a2 + a3);
}

// Blocked right-looking sweep for the mode tile of this CTA; `rev` maps the system's numbering to memory rows.
__global__ void __launch_bounds__(TRSV_THREADS, 1)
k_trsv_blocked(double* __restrict__ w, int64_t ld, int64_t n, int rev, TriDev T) {
    extern __shared__ __align__(16) double trsv_smem[];
    double(*Zs)[MT] = reinterpret_cast<double(*)[MT]>(trsv_smem);                     // [BW][MT] rows of the block
    double(*red)[MT + 1] = reinterpret_cast<double(*)[MT + 1]>(trsv_smem + BW * MT);  // [RS][MT+1] partial sums
    double* s_dinv = trsv_smem + BW * MT + RS * (MT + 1);                             // [BW]
    double* s_dval = s_dinv + BW;                                                     // [DE_MAX]
    int32_t* s_drow = reinterpret_cast<int32_t*>(s_dval + DE_MAX);                    // [BW+1] block-relative offsets
    int32_t* s_dsplit = s_drow + BW + 1;                                              // [BW]
    int32_t* s_levptr = s_dsplit + BW;                                                // [BW+1]
    uint16_t* s_levrows = reinterpret_cast<uint16_t*>(s_levptr + BW + 1);             // [BW]
    uint16_t* s_didx = s_levrows + BW;                                                // [DE_MAX]
    const int tid = threadIdx.x;
    const int lane = tid % MT;  // mode within the tile
    const int hw = tid / MT;    // half-warp = row slot
    const int64_t mode = (int64_t)blockIdx.x * MT + lane;
    auto phys = [&](int64_t k) { return rev ? (n - 1 - k) : k; };
    for (int b = 0; b < T.nblocks; ++b) {
        const int64_t j0 = T.blk_start[b];
        const int width = (int)(T.blk_start[b + 1] - j0);
        const int lev0 = T.blk_info[4 * b], nlev = T.blk_info[4 * b + 1], dense = T.blk_info[4 * b + 2];
        const int d0 = T.drow[j0];
        const int nde = T.drow[j0 + width] - d0;
        const bool staged = nde <= DE_MAX;
        // ---- stage rows, per-row data and (if they fit) the diagonal-block entries -------------------------
        for (int slot = hw; slot < width; slot += RS) Zs[slot][lane] = w[phys(j0 + slot) * ld + mode];
        for (int k = tid; k < width; k += TRSV_THREADS) {
            s_dinv[k] = T.dinv[j0 + k];
            s_dsplit[k] = T.dsplit[j0 + k];
            s_levrows[k] = T.levrows[j0 + k];
        }
        for (int k = tid; k <= width; k += TRSV_THREADS) s_drow[k] = T.drow[j0 + k] - d0;
        for (int k = tid; k <= nlev; k += TRSV_THREADS) s_levptr[k] = T.levptr[lev0 + k];
        if (staged)
            for (int k = tid; k < nde; k += TRSV_THREADS) {
                s_dval[k] = T.dval[d0 + k];
                s_didx[k] = T.didx[d0 + k];
            }
        __syncthreads();
        const double* gv = T.dval + d0;
        const uint16_t* gi = T.didx + d0;
        if (!dense) {
            // ---- sparse diagonal block (subtree): level by level, one half-warp per row ---------------------
            for (int l = 0; l < nlev; ++l) {
                const int r0 = s_levptr[l], r1 = s_levptr[l + 1];
                for (int q = r0 + hw; q < r1; q += RS) {
                    const int rl = s_levrows[q];
                    const int e0 = s_drow[rl], e1 = s_drow[rl + 1];
                    const double dot = staged ? dot_rows(s_dval, s_didx, e0, e1, 1, Zs, lane) : dot_rows(gv, gi, e0, e1, 1, Zs, lane);
                    Zs[rl][lane] = (Zs[rl][lane] - dot) * s_dinv[rl];
                }
                __syncthreads();
            }
        } else {
            // ---- dense diagonal block (separator chain): panels of PANEL rows --------------------------------
            // (a) all half-warps: the part of every panel row that only needs rows left of the panel
            // (b) half-warp 0: the PANEL x PANEL triangle, sequentially
            constexpr int PER = RS / PANEL;  // half-warps per panel row
            for (int p0 = 0; p0 < width; p0 += PANEL) {
                const int rl = p0 + hw / PER, part = hw % PER;
                double acc = 0.0;
                if (rl < width) {
                    const int e0 = s_drow[rl], e1 = e0 + s_dsplit[rl];
                    acc = staged ? dot_rows(s_dval, s_didx, e0 + part, e1, PER, Zs, lane)
                                 : dot_rows(gv, gi, e0 + part, e1, PER, Zs, lane);
                }
                red[hw][lane] = acc;
                __syncthreads();
                if (hw == 0) {
                    const int pend = min(p0 + PANEL, width);
                    for (int r = p0; r < pend; ++r) {
                        double sum = Zs[r][lane];
#pragma unroll
                        for (int j = 0; j < PER; ++j) sum -= red[(r - p0) * PER + j][lane];
                        const int e0 = s_drow[r] + s_dsplit[r], e1 = s_drow[r + 1];
                        sum -= staged ? dot_rows(s_dval, s_didx, e0, e1, 1, Zs, lane) : dot_rows(gv, gi, e0, e1, 1, Zs, lane);
                        Zs[r][lane] = sum * s_dinv[r];  // every lane only ever touches its own mode column
                    }
                }
                __syncthreads();
            }
        }
        // ---- finished rows back to memory, then push to all later rows --------------------------------------
        for (int slot = hw; slot < width; slot += RS) w[phys(j0 + slot) * ld + mode] = Zs[slot][lane];
        const int s0 = T.segptr[b], s1 = T.segptr[b + 1];
        for (int sg = s0 + hw; sg < s1; sg += RS) {
            const int r = T.seg[3 * sg], p0 = T.seg[3 * sg + 1], len = T.seg[3 * sg + 2];
            const double* v = T.val + p0;
            const int32_t* ix = T.idx + p0;
            double* wr = w + phys(r) * ld + mode;
            const double old = *wr;  // issued early: the latency hides behind the dot product
            double a[4] = {0.0, 0.0, 0.0, 0.0};
            int p = 0;
            for (; p + 8 <= len; p += 8) {
                double vv[8];
                int ii[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    vv[u] = __ldg(v + p + u);
                    ii[u] = __ldg(ix + p + u) - (int)j0;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) a[u & 3] = fma(vv[u], Zs[ii[u]][lane], a[u & 3]);
            }
            for (; p < len; ++p) a[0] = fma(__ldg(v + p), Zs[__ldg(ix + p) - (int)j0][lane], a[0]);
            *wr = old - ((a[0] + a[1]) + (a[2] + a[3]));
        }
        __syncthreads();
    }
}

constexpr size_t TRSV_SMEM = sizeof(double) * ((size_t)BW * MT + (size_t)RS * (MT + 1) + BW + DE_MAX) +
                             sizeof(int32_t) * (3 * BW + 2) + sizeof(uint16_t) * (BW + DE_MAX) + 16;

struct TriHost {
    std::vector<int64_t> ptr;
    std::vector<int32_t> idx;
    std::vector<double> val, dinv;
};

int upload_tri(asgfem_ctx* ctx, const TriHost& H, int64_t n, const std::vector<int32_t>& starts, TriDev& D) {
    const int nblocks = (int)starts.size() - 1;
    std::vector<int32_t> blk_of((size_t)n);
    for (int b = 0; b < nblocks; ++b)
        for (int32_t k = starts[b]; k < starts[b + 1]; ++k) blk_of[k] = b;
    std::vector<int32_t> drow((size_t)n + 1, 0), dsplit((size_t)n, 0), blk_info((size_t)4 * nblocks, 0), levptr, segptr((size_t)nblocks + 1, 0), seg;
    std::vector<uint16_t> didx, levrows((size_t)n);
    std::vector<double> dval;
    std::vector<int32_t> lev((size_t)BW + 1), dcount((size_t)n, 0);
    for (int b = 0; b < nblocks; ++b) {
        const int64_t j0 = starts[b], j1 = starts[b + 1];
        int nlev = 0;
        for (int64_t k = j0; k < j1; ++k) {
            // diagonal-block part = trailing entries of the sorted row with column >= j0
            int64_t p = H.ptr[k + 1];
            while (p > H.ptr[k] && H.idx[p - 1] >= j0) --p;
            dcount[k] = (int32_t)(H.ptr[k + 1] - p);
            const int64_t panel0 = j0 + ((k - j0) / PANEL) * PANEL;
            int l = 0, split = 0;
            for (int64_t q = p; q < H.ptr[k + 1]; ++q) {
                didx.push_back((uint16_t)(H.idx[q] - j0));
                dval.push_back(H.val[q]);
                l = std::max(l, lev[H.idx[q] - j0] + 1);
                if (H.idx[q] < panel0) ++split;
            }
            drow[k + 1] = (int32_t)didx.size();
            dsplit[k] = split;
            lev[k - j0] = l;
            nlev = std::max(nlev, l + 1);
        }
        const int width = (int)(j1 - j0);
        blk_info[4 * b] = (int32_t)levptr.size();
        blk_info[4 * b + 1] = nlev;
        blk_info[4 * b + 2] = (2 * nlev > width && width > PANEL) ? 1 : 0;  // chain-like block: panel algorithm
        std::vector<int32_t> cnt((size_t)nlev + 1, 0);
        for (int64_t k = j0; k < j1; ++k) cnt[lev[k - j0] + 1]++;
        for (int l = 0; l < nlev; ++l) cnt[l + 1] += cnt[l];
        for (int l = 0; l <= nlev; ++l) levptr.push_back(cnt[l]);
        std::vector<int32_t> fill(cnt.begin(), cnt.end() - 1);
        for (int64_t k = j0; k < j1; ++k) levrows[j0 + fill[lev[k - j0]]++] = (uint16_t)(k - j0);
    }
    // push segments: maximal runs of a row's entries inside one earlier block
    for (int pass = 0; pass < 2; ++pass) {
        std::vector<int32_t> fill;
        if (pass == 1) {
            for (int b = 0; b < nblocks; ++b) segptr[b + 1] += segptr[b];
            seg.resize((size_t)3 * segptr[nblocks]);
            fill.assign(segptr.begin(), segptr.end() - 1);
        }
        for (int64_t k = 0; k < n; ++k) {
            const int64_t pend = H.ptr[k + 1] - dcount[k];
            int64_t p = H.ptr[k];
            while (p < pend) {
                const int b = blk_of[H.idx[p]];
                int64_t q = p + 1;
                while (q < pend && blk_of[H.idx[q]] == b) ++q;
                if (pass == 0) {
                    segptr[b + 1]++;
                } else {
                    int32_t at = fill[b]++;
                    seg[3 * (size_t)at] = (int32_t)k;
                    seg[3 * (size_t)at + 1] = (int32_t)p;
                    seg[3 * (size_t)at + 2] = (int32_t)(q - p);
                }
                p = q;
            }
        }
    }
    D.nblocks = nblocks;
    int rc = 0;
    rc |= dev_upload(ctx, &D.idx, H.idx);
    rc |= dev_upload(ctx, &D.val, H.val);
    rc |= dev_upload(ctx, &D.segptr, segptr);
    rc |= dev_upload(ctx, &D.seg, seg);
    rc |= dev_upload(ctx, &D.dinv, H.dinv);
    rc |= dev_upload(ctx, &D.drow, drow);
    rc |= dev_upload(ctx, &D.dsplit, dsplit);
    rc |= dev_upload(ctx, &D.didx, didx);
    rc |= dev_upload(ctx, &D.dval, dval);
    rc |= dev_upload(ctx, &D.blk_start, starts);
    rc |= dev_upload(ctx, &D.blk_info, blk_info);
    rc |= dev_upload(ctx, &D.levptr, levptr);
    rc |= dev_upload(ctx, &D.levrows, levrows);
    return rc;
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
    // dof coordinates (if mesh and space are known) steer the nested dissection towards straight separators
    std::vector<double> xy;
    if (ctx->order > 0 && ctx->ndofs_space == ctx->n && !ctx->h_coords.empty()) {
        xy.assign((size_t)2 * ctx->n, 0.0);
        const int nd = ctx->ndofs4cell;
        for (int64_t c = 0; c < ctx->ncells; ++c) {
            const int32_t* cn = ctx->h_cellnodes.data() + 3 * c;
            const int32_t* cd = ctx->h_celldofs.data() + (int64_t)nd * c;
            for (int i = 0; i < 3; ++i) {
                xy[2 * (int64_t)cd[i]] = ctx->h_coords[2 * (int64_t)cn[i]];
                xy[2 * (int64_t)cd[i] + 1] = ctx->h_coords[2 * (int64_t)cn[i] + 1];
            }
            for (int f = 0; f < nd - 3; ++f) {
                int a = cn[f], b = cn[(f + 1) % 3];
                xy[2 * (int64_t)cd[3 + f]] = 0.5 * (ctx->h_coords[2 * (int64_t)a] + ctx->h_coords[2 * (int64_t)b]);
                xy[2 * (int64_t)cd[3 + f] + 1] = 0.5 * (ctx->h_coords[2 * (int64_t)a + 1] + ctx->h_coords[2 * (int64_t)b + 1]);
            }
        }
    }
    int rc = cholesky_reduced(ctx->n, ctx->h_rowptr.data(), ctx->h_col.data(), k0.data(), ctx->h_bmask.data(),
                              xy.empty() ? nullptr : xy.data(), BW, F, err);
    if (rc) return fail(ctx, rc, "precond_setup: " + err);
    ASG_CHECK(ctx, (int64_t)F.Li.size() < (1ll << 31), ASGFEM_EINVAL, "precond_setup: factor with >= 2^31 nonzeros not supported");
    PrecondPlan* P = new PrecondPlan();
    ctx->precond = P;
    P->nred = F.n;
    P->lnz = (int64_t)F.Li.size();
    const int64_t n = F.n;
    TriHost fw, bw;
    fw.ptr = F.Lp;
    fw.idx = F.Li;
    fw.val = F.Lx;
    fw.dinv = F.dinv;
    // backward system L^T in the reversed numbering k' = n-1-k: row k' holds the column k of L, rows i > k mapped to
    // i' = n-1-i < k' in ascending order
    {
        std::vector<int64_t> cptr((size_t)n + 1, 0);
        for (int32_t j : F.Li) cptr[j + 1]++;
        for (int64_t k = 0; k < n; ++k) cptr[k + 1] += cptr[k];
        std::vector<int32_t> cidx(F.Li.size());
        std::vector<double> cval(F.Li.size());
        std::vector<int64_t> fill(cptr.begin(), cptr.end() - 1);
        for (int64_t k = 0; k < n; ++k)
            for (int64_t p = F.Lp[k]; p < F.Lp[k + 1]; ++p) {
                int64_t at = fill[F.Li[p]]++;
                cidx[at] = (int32_t)k;  // rows ascending inside every column
                cval[at] = F.Lx[p];
            }
        bw.ptr.assign((size_t)n + 1, 0);
        bw.idx.resize(F.Li.size());
        bw.val.resize(F.Li.size());
        bw.dinv.resize((size_t)n);
        int64_t at = 0;
        for (int64_t kp = 0; kp < n; ++kp) {
            const int64_t k = n - 1 - kp;
            for (int64_t p = cptr[k + 1] - 1; p >= cptr[k]; --p) {  // descending i -> ascending i'
                bw.idx[at] = (int32_t)(n - 1 - cidx[p]);
                bw.val[at] = cval[p];
                ++at;
            }
            bw.ptr[kp + 1] = at;
            bw.dinv[kp] = F.dinv[k];
        }
    }
    rc = dev_upload(ctx, &P->d_perm, F.perm);
    std::vector<int32_t> rstarts;
    for (size_t k = F.block_start.size(); k-- > 0;) rstarts.push_back((int32_t)(n - F.block_start[k]));
    rc |= upload_tri(ctx, fw, n, F.block_start, P->fwd);
    rc |= upload_tri(ctx, bw, n, rstarts, P->bwd);
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
        const size_t smem = TRSV_SMEM;
        ASG_CUDA(ctx, cudaFuncSetAttribute(k_trsv_blocked, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_trsv_blocked<<<tiles, TRSV_THREADS, smem, ctx->stream>>>(P->d_work, ld, P->nred, 0, P->fwd);
        k_trsv_blocked<<<tiles, TRSV_THREADS, smem, ctx->stream>>>(P->d_work, ld, P->nred, 1, P->bwd);
    }
    // z may alias r: boundary rows are zeroed first, interior rows are overwritten from the work vector
    k_zero_masked_rows<<<(unsigned)std::min<int64_t>(ctx->n, 148 * 8), 128, 0, ctx->stream>>>(z, ctx->d_bmask, ctx->n, ld);
    if (P->nred > 0) k_scatter_perm<<<blocks, 256, 0, ctx->stream>>>(P->d_work, z, P->d_perm, P->nred, ld);
    ASG_CUDA(ctx, cudaGetLastError());
    return 0;
}

}  // namespace asgfem
