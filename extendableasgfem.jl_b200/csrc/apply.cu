// Fused SGFE operator  Y = sum_m (G_m (x) K_m) X  with Dirichlet rows zeroed
// (LinearAlgebra.mul!(Ax, S::MySystemPrimal, x), src/modelproblems/solvers_poisson_primal.jl:86-124).
//
// Two hand-written kernels for sm_100a:
//
//  variant 1  k_apply_gather   thread = (dof row i, mode mu); walks the neighbour list of mu in the
//             reference's accumulation order (nu ascending, direction ascending) and reads X and K_m
//             through L1/L2.  Simple and exact in ordering; used for tiny problems, as the in-library
//             cross-check of variant 2 and for patterns the tiled plan cannot hold.
//
//  variant 2  k_apply_tiled    one CTA per row block (<= 128 dofs clustered by BFS on the matrix graph),
//             looping over mode tiles.  Per tile the block of X restricted to (columns of the row block) x
//             (own modes + coupled halo modes) is staged once in shared memory, every K_m row is held in
//             registers (lanes = dof rows) and reused over all (mu <- nu) couplings of direction m inside
//             the tile; partial sums live in shared memory and are written to Y once, coalesced.
//             K_m is streamed once per (row block, tile, direction), X once per (row block, tile).
//
// Both kernels never touch uncoupled (m, nu) pairs: flops = 2 nnz (N + 2E).
#include <algorithm>
#include <cstring>
#include <numeric>
#include <queue>

#include "common.h"

namespace asgfem {

// ------------------------------------------------------------------------------------------------
// variant 1: gather kernel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_apply_gather(int64_t row0, int64_t nrows, int64_t N, int64_t ld, int64_t nnz, const int64_t* __restrict__ rowptr,
               const int32_t* __restrict__ col, const double* __restrict__ vals, const int32_t* __restrict__ cptr,
               const int32_t* __restrict__ cm, const int32_t* __restrict__ cnu, const double* __restrict__ cg,
               const uint8_t* __restrict__ bmask, const double* __restrict__ x, double* __restrict__ y) {
    int64_t i = row0 + (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    int64_t mu = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    if (i >= nrows || mu >= ld) return;
    double acc = 0.0;
    if (mu < N && !bmask[i]) {
        int64_t p0 = rowptr[i], p1 = rowptr[i + 1];
        for (int64_t p = p0; p < p1; ++p) acc = fma(vals[p], x[(int64_t)col[p] * ld + mu], acc);
        for (int32_t e = cptr[mu]; e < cptr[mu + 1]; ++e) {
            const double* vm = vals + (int64_t)cm[e] * nnz;
            int64_t nu = cnu[e];
            double t = 0.0;
            for (int64_t p = p0; p < p1; ++p) t = fma(vm[p], x[(int64_t)col[p] * ld + nu], t);
            acc = fma(cg[e], t, acc);
        }
    }
    y[i * ld + mu] = acc;
}

// ------------------------------------------------------------------------------------------------
// variant 2: tiled kernel - host plan
// ------------------------------------------------------------------------------------------------
constexpr int TILED_WARPS = 4;
constexpr int TILED_ROWS = 32 * TILED_WARPS;  // dof rows per CTA
constexpr int SMEM_BUDGET = 225 * 1024;
constexpr int EU = 4;  // entries processed together per thread in the tiled kernel

struct ApplyPlan {
    // row blocks
    int nblocks = 0, KW = 0, Cmax = 0, Cpad = 0, Rpad = TILED_ROWS + 1;
    int32_t* d_blk_rows = nullptr;    // nblocks x TILED_ROWS global row id (-1 = padding)
    int32_t* d_blk_colptr = nullptr;  // nblocks+1
    int32_t* d_blk_cols = nullptr;    // global column ids of each block
    int32_t* d_ell_lcol = nullptr;    // nblocks x KW x TILED_ROWS local column slot
    int32_t* d_ell_pos = nullptr;     // nblocks x KW x TILED_ROWS position in CSR values (-1 = padding)
    // mode tiles
    int ntiles = 0, Smax = 0, Tmax = 0;
    int32_t* d_tile_t0 = nullptr;      // first own mode
    int32_t* d_tile_t = nullptr;       // number of own modes
    int32_t* d_tile_sptr = nullptr;    // ntiles+1 -> staged mode list (own first, then halo)
    int32_t* d_tile_smodes = nullptr;
    int32_t* d_tile_dptr = nullptr;    // ntiles+1 -> directions present in the tile
    int32_t* d_dir_m = nullptr;        // direction (0 = mean term)
    int32_t* d_dir_eptr = nullptr;     // ndirs+1 -> entries
    uint32_t* d_ent_ds = nullptr;      // (dst << 16) | src, dst = own slot, src = staged slot
    double* d_ent_g = nullptr;
    size_t smem_bytes = 0;
    bool usable = false;
};

namespace {

template <class T>
int up(asgfem_ctx* ctx, T** d, const std::vector<T>& h) {
    return dev_upload(ctx, d, h);
}

void free_plan_arrays(ApplyPlan* p) {
    void* ptrs[] = {p->d_blk_rows, p->d_blk_colptr, p->d_blk_cols, p->d_ell_lcol, p->d_ell_pos,  p->d_tile_t0,
                    p->d_tile_t,   p->d_tile_sptr,  p->d_tile_smodes, p->d_tile_dptr, p->d_dir_m, p->d_dir_eptr,
                    p->d_ent_ds,   p->d_ent_g};
    for (void* q : ptrs)
        if (q) cudaFree(q);
}

// Greedy BFS clustering of the output rows into blocks of TILED_ROWS rows with few distinct columns.
// `order` holds whole blocks of TILED_ROWS entries; short blocks are padded with -1.
void cluster_rows(int64_t nrows, const std::vector<int64_t>& rowptr, const std::vector<int32_t>& col,
                  std::vector<int32_t>& order) {
    order.clear();
    order.reserve((size_t)nrows + TILED_ROWS);
    std::vector<uint8_t> seen((size_t)nrows, 0);  // 1 = assigned or queued in the current block
    std::vector<int32_t> frontier;                // carried between blocks so that consecutive blocks are neighbours
    size_t fhead = 0;
    int64_t next_seed = 0, assigned = 0;
    std::vector<int32_t> q;
    while (assigned < nrows) {
        q.clear();
        size_t head = 0;
        int taken = 0;
        while (taken < TILED_ROWS && assigned < nrows) {
            if (head == q.size()) {  // (re)seed: frontier of earlier blocks first, then the lowest free row
                int32_t seed = -1;
                while (fhead < frontier.size()) {
                    int32_t c = frontier[fhead++];
                    if (!seen[c]) {
                        seed = c;
                        break;
                    }
                }
                if (seed < 0) {
                    while (seen[next_seed]) ++next_seed;
                    seed = (int32_t)next_seed;
                }
                seen[seed] = 1;
                q.push_back(seed);
            }
            int32_t r = q[head++];
            order.push_back(r);
            ++taken;
            ++assigned;
            for (int64_t p = rowptr[r]; p < rowptr[r + 1]; ++p) {
                int32_t c = col[p];
                if (c < nrows && !seen[c]) {
                    seen[c] = 1;
                    q.push_back(c);
                }
            }
        }
        for (size_t k = head; k < q.size(); ++k) {  // queued but not taken: release, remember as future seeds
            seen[q[k]] = 0;
            frontier.push_back(q[k]);
        }
        if (fhead > (1u << 22)) {
            frontier.erase(frontier.begin(), frontier.begin() + (long)fhead);
            fhead = 0;
        }
        for (int k = taken; k < TILED_ROWS; ++k) order.push_back(-1);
    }
}

}  // namespace

void apply_free_plan(asgfem_ctx* ctx) {
    apply_rows_free(ctx);
    apply_dir_free(ctx);
    apply_ts_free(ctx);
    apply_ts2_free(ctx);
    if (!ctx->plan) return;
    free_plan_arrays(ctx->plan);
    delete ctx->plan;
    ctx->plan = nullptr;
}

int apply_build_plan(asgfem_ctx* ctx) {
    apply_free_plan(ctx);
    ApplyPlan* P = new ApplyPlan();
    ctx->plan = P;
    const int64_t n = ctx->n, N = ctx->N;
    const int64_t nrows = ctx->n_owned >= 0 ? ctx->n_owned : n;
    if (n == 0 || N == 0 || ctx->M < 0) return 0;

    // ---- row blocks ---------------------------------------------------------------------------
    std::vector<int32_t> order;
    cluster_rows(nrows, ctx->h_rowptr, ctx->h_col, order);
    // rows of a block sorted by global id: with mesh-ordered numberings lane r then reads column slot ~ r + const,
    // which keeps the shared-memory gathers of X (nearly) bank-conflict free
    for (size_t b0 = 0; b0 + TILED_ROWS <= order.size(); b0 += TILED_ROWS) {
        auto first = order.begin() + (long)b0, last = first + TILED_ROWS;
        auto mid = std::partition(first, last, [](int32_t v) { return v >= 0; });
        std::sort(first, mid);
    }
    P->nblocks = (int)(order.size() / TILED_ROWS);
    int KW = 0;
    for (int64_t i = 0; i < nrows; ++i) KW = std::max<int>(KW, (int)(ctx->h_rowptr[i + 1] - ctx->h_rowptr[i]));
    P->KW = KW;
    if (KW > 32) return 0;  // plan not usable -> gather kernel
    int KWt = KW <= 8 ? 8 : (KW <= 16 ? 16 : (KW <= 24 ? 24 : 32));
    P->KW = KWt;
    std::vector<int32_t> blk_colptr(1, 0), blk_cols, ell_lcol((size_t)P->nblocks * KWt * TILED_ROWS, 0),
        ell_pos((size_t)P->nblocks * KWt * TILED_ROWS, -1);
    std::vector<int32_t> cols;
    std::unordered_map<int32_t, int32_t> slot;
    int Cmax = 0;
    for (int b = 0; b < P->nblocks; ++b) {
        cols.clear();
        for (int r = 0; r < TILED_ROWS; ++r) {
            int32_t row = order[(size_t)b * TILED_ROWS + r];
            if (row < 0) continue;
            for (int64_t p = ctx->h_rowptr[row]; p < ctx->h_rowptr[row + 1]; ++p) cols.push_back(ctx->h_col[p]);
        }
        std::sort(cols.begin(), cols.end());
        cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
        slot.clear();
        for (size_t k = 0; k < cols.size(); ++k) slot[cols[k]] = (int32_t)k;
        for (int r = 0; r < TILED_ROWS; ++r) {
            int32_t row = order[(size_t)b * TILED_ROWS + r];
            if (row < 0) continue;
            int k = 0;
            for (int64_t p = ctx->h_rowptr[row]; p < ctx->h_rowptr[row + 1]; ++p, ++k) {
                size_t at = ((size_t)b * KWt + k) * TILED_ROWS + r;
                ell_lcol[at] = slot[ctx->h_col[p]];
                ell_pos[at] = (int32_t)p;
            }
        }
        blk_cols.insert(blk_cols.end(), cols.begin(), cols.end());
        blk_colptr.push_back((int32_t)blk_cols.size());
        Cmax = std::max<int>(Cmax, (int)cols.size());
    }
    P->Cmax = Cmax;
    P->Cpad = Cmax | 1;  // odd leading dimension: conflict-free transposed staging

    // ---- mode tiles ---------------------------------------------------------------------------
    // shared memory: X_s[S][Cpad] + Y_s[T][Rpad]; T = S/2 is a good split for +-e_m couplings
    int64_t avail = SMEM_BUDGET - 1024;
    int Smax = 0, Tmax = 0;
    for (int S = 16; S <= 1024; S += 8) {
        int T = std::max(8, S / 2);
        int64_t need = (int64_t)S * P->Cpad * 8 + (int64_t)T * P->Rpad * 8;
        if (need <= avail) {
            Smax = S;
            Tmax = T;
        }
    }
    if (Smax < 16) return 0;
    Tmax = (int)std::min<int64_t>(Tmax, N);
    P->Smax = Smax;
    P->Tmax = Tmax;
    const Coupling& C = ctx->coup;
    std::vector<int32_t> t0s, ts, sptr(1, 0), smodes, dptr(1, 0), dir_m, dir_eptr(1, 0);
    std::vector<uint32_t> ent_ds;
    std::vector<double> ent_g;
    std::vector<int32_t> mark((size_t)N, -1);
    int64_t mu = 0;
    int tile_id = 0;
    while (mu < N) {
        // grow the own range while own + halo fits
        std::vector<int32_t> halo;
        int64_t t = 0;
        while (mu + t < N && t < Tmax) {
            int64_t cand = mu + t;
            size_t before = halo.size();
            for (int32_t e = C.ptr[cand]; e < C.ptr[cand + 1]; ++e) {
                int32_t nu = C.nu[e];
                if (nu >= mu && nu <= cand) continue;  // inside own range (so far)
                if (mark[nu] != tile_id) {
                    mark[nu] = tile_id;
                    halo.push_back(nu);
                }
            }
            // halo modes that the growing own range swallows are dropped later; conservative count here
            if ((int64_t)halo.size() + (t + 1) > Smax && t > 0) {
                for (size_t k = before; k < halo.size(); ++k) mark[halo[k]] = -1;
                halo.resize(before);
                break;
            }
            ++t;
        }
        if ((int64_t)halo.size() + t > Smax) {  // a single mode with more neighbours than fit
            free_plan_arrays(P);
            *P = ApplyPlan();
            return 0;
        }
        std::vector<int32_t> staged;
        for (int64_t k = 0; k < t; ++k) staged.push_back((int32_t)(mu + k));
        std::sort(halo.begin(), halo.end());
        for (int32_t h : halo)
            if (h < mu || h >= mu + t) staged.push_back(h);
        std::unordered_map<int32_t, int32_t> spos;
        for (size_t k = 0; k < staged.size(); ++k) spos[staged[k]] = (int32_t)k;
        // entries grouped by direction; direction 0 = mean term (dst = src = own slot)
        std::vector<std::vector<std::pair<uint32_t, double>>> per_dir((size_t)ctx->M + 1);
        for (int64_t k = 0; k < t; ++k) {
            per_dir[0].push_back({(uint32_t)((k << 16) | k), 1.0});
            for (int32_t e = C.ptr[mu + k]; e < C.ptr[mu + k + 1]; ++e) {
                int m = C.m[e];
                if (m > ctx->M) continue;  // checked at set_multiindices/apply time
                per_dir[m].push_back({(uint32_t)((k << 16) | (uint32_t)spos[C.nu[e]]), C.g[e]});
            }
        }
        for (int m = 0; m <= ctx->M; ++m) {
            if (per_dir[m].empty()) continue;
            dir_m.push_back(m);
            for (auto& pr : per_dir[m]) {
                ent_ds.push_back(pr.first);
                ent_g.push_back(pr.second);
            }
            dir_eptr.push_back((int32_t)ent_ds.size());
        }
        dptr.push_back((int32_t)dir_m.size());
        t0s.push_back((int32_t)mu);
        ts.push_back((int32_t)t);
        smodes.insert(smodes.end(), staged.begin(), staged.end());
        sptr.push_back((int32_t)smodes.size());
        mu += t;
        ++tile_id;
    }
    P->ntiles = tile_id;
    P->smem_bytes = (size_t)Smax * P->Cpad * 8 + (size_t)Tmax * P->Rpad * 8;

    int rc = 0;
    rc |= up(ctx, &P->d_blk_rows, order);
    rc |= up(ctx, &P->d_blk_colptr, blk_colptr);
    rc |= up(ctx, &P->d_blk_cols, blk_cols);
    rc |= up(ctx, &P->d_ell_lcol, ell_lcol);
    rc |= up(ctx, &P->d_ell_pos, ell_pos);
    rc |= up(ctx, &P->d_tile_t0, t0s);
    rc |= up(ctx, &P->d_tile_t, ts);
    rc |= up(ctx, &P->d_tile_sptr, sptr);
    rc |= up(ctx, &P->d_tile_smodes, smodes);
    rc |= up(ctx, &P->d_tile_dptr, dptr);
    rc |= up(ctx, &P->d_dir_m, dir_m);
    rc |= up(ctx, &P->d_dir_eptr, dir_eptr);
    rc |= up(ctx, &P->d_ent_ds, ent_ds);
    rc |= up(ctx, &P->d_ent_g, ent_g);
    if (rc) return ASGFEM_ECUDA;
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    P->usable = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// variant 2: tiled kernel - device code
// ------------------------------------------------------------------------------------------------
struct TiledArgs {
    int64_t ld, nnz;
    int nblocks, ntiles, Cpad, Rpad, Smax;
    const int32_t *blk_rows, *blk_colptr, *blk_cols, *ell_lcol, *ell_pos;
    const int32_t *tile_t0, *tile_t, *tile_sptr, *tile_smodes, *tile_dptr, *dir_m, *dir_eptr;
    const uint32_t* ent_ds;
    const double* ent_g;
    const double* vals;
    const uint8_t* bmask;
    const double* x;
    double* y;
};

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
    unsigned saddr = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(saddr), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

template <int KW>
__global__ void __launch_bounds__(TILED_ROWS, 1) k_apply_tiled(TiledArgs a) {
    extern __shared__ __align__(16) double smem[];
    double* Xs = smem;                                  // [Smax][Cpad]  staged X, column (dof) index fastest
    double* Ys = smem + (size_t)a.Smax * a.Cpad;        // [Tmax][Rpad]  partial sums, row (dof) index fastest
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int b = blockIdx.x; b < a.nblocks; b += gridDim.x) {
        const int c0 = a.blk_colptr[b], C = a.blk_colptr[b + 1] - c0;
        int xoff[KW];
        int kpos[KW];
#pragma unroll
        for (int k = 0; k < KW; ++k) {
            size_t at = ((size_t)b * KW + k) * TILED_ROWS + tid;
            xoff[k] = a.ell_lcol[at];
            kpos[k] = a.ell_pos[at];
        }

        for (int tile = 0; tile < a.ntiles; ++tile) {
            const int s0 = a.tile_sptr[tile], S = a.tile_sptr[tile + 1] - s0;
            const int t0 = a.tile_t0[tile], T = a.tile_t[tile];
            const int d0 = a.tile_dptr[tile], d1 = a.tile_dptr[tile + 1];
            __syncthreads();  // previous tile fully consumed
            // stage X asynchronously (LDGSTS): warp per column, lanes over staged modes - the own modes of the tile
            // are contiguous in global memory, the halo modes are gathered
            for (int c = warp; c < C; c += TILED_WARPS) {
                const double* xr = a.x + (int64_t)a.blk_cols[c0 + c] * a.ld;
                for (int s = lane; s < S; s += 32) cp_async8(Xs + s * a.Cpad + c, xr + a.tile_smodes[s0 + s]);
            }
            for (int k = tid; k < T * a.Rpad; k += TILED_ROWS) Ys[k] = 0.0;
            // K_m row of the first direction travels while the copies land
            double kr[KW], kn[KW];
            {
                const double* vm = a.vals + (int64_t)a.dir_m[d0] * a.nnz;
#pragma unroll
                for (int k = 0; k < KW; ++k) kn[k] = kpos[k] >= 0 ? __ldg(vm + kpos[k]) : 0.0;
            }
            cp_async_wait_all();
            __syncthreads();

            for (int d = d0; d < d1; ++d) {
#pragma unroll
                for (int k = 0; k < KW; ++k) kr[k] = kn[k];
                if (d + 1 < d1) {  // prefetch the K_m row of the next direction
                    const double* vm = a.vals + (int64_t)a.dir_m[d + 1] * a.nnz;
#pragma unroll
                    for (int k = 0; k < KW; ++k) kn[k] = kpos[k] >= 0 ? __ldg(vm + kpos[k]) : 0.0;
                }
                const int e0 = a.dir_eptr[d], e1 = a.dir_eptr[d + 1];
                for (int e = e0; e < e1; e += EU) {  // EU independent dot-product chains for ILP
                    uint32_t ds[EU];
                    double g[EU], acc[EU];
                    const double* xp[EU];
#pragma unroll
                    for (int u = 0; u < EU; ++u) {
                        int ee = min(e + u, e1 - 1);
                        ds[u] = __ldg(a.ent_ds + ee);
                        g[u] = (e + u < e1) ? __ldg(a.ent_g + ee) : 0.0;
                        xp[u] = Xs + (ds[u] & 0xffffu) * a.Cpad;
                        acc[u] = 0.0;
                    }
#pragma unroll
                    for (int k = 0; k < KW; ++k) {
#pragma unroll
                        for (int u = 0; u < EU; ++u) acc[u] = fma(kr[k], xp[u][xoff[k]], acc[u]);
                    }
#pragma unroll
                    for (int u = 0; u < EU; ++u) {  // in order: consecutive entries may share dst
                        double* yp = Ys + (ds[u] >> 16) * a.Rpad + tid;
                        *yp = fma(g[u], acc[u], *yp);
                    }
                }
            }
            __syncthreads();
            // write Y: warp per row, lanes over own modes (coalesced)
            for (int r = warp; r < TILED_ROWS; r += TILED_WARPS) {
                int32_t grow = a.blk_rows[(size_t)b * TILED_ROWS + r];
                if (grow < 0) continue;
                bool m = a.bmask[grow];
                double* yr = a.y + (int64_t)grow * a.ld + t0;
                for (int k = lane; k < T; k += 32) yr[k] = m ? 0.0 : Ys[k * a.Rpad + r];
            }
        }
    }
}

template <int KW>
static int launch_tiled(asgfem_ctx* ctx, const TiledArgs& a, size_t smem) {
    static bool configured = false;
    if (!configured) {
        ASG_CUDA(ctx, cudaFuncSetAttribute(k_apply_tiled<KW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
        configured = true;
    }
    int grid = std::min(a.nblocks, 148);
    k_apply_tiled<KW><<<grid, TILED_ROWS, smem, ctx->stream>>>(a);
    ASG_CUDA(ctx, cudaGetLastError());
    return 0;
}

int apply_launch(asgfem_ctx* ctx, const double* x, double* y, int64_t r0, int64_t r1) {
    const int64_t nrows_all = ctx->n_owned >= 0 ? ctx->n_owned : ctx->n;
    const bool ranged = r1 >= 0;
    if (!ranged) {
        r0 = 0;
        r1 = nrows_all;
    }
    r1 = std::min(r1, nrows_all);
    const int64_t nrows = r1;
    ApplyPlan* P = ctx->plan;
    int variant = ctx->apply_variant;
    // automatic choice (profiles/r01_*, 1M dofs x 2000 modes on B200): direction-major row-resident kernel 76 ms,
    // dst-major row-resident kernel 125 ms (both with DRAM traffic at the algorithmic minimum), gather kernel 129 ms
    // with 2.6x the traffic, row-block tiled kernel 350 ms
    if (variant == 0) {
        variant = 1;
        if (ctx->n * ctx->N >= (1 << 16)) {
            if (apply_ts2_preferred(ctx))
                variant = 7;
            else if (apply_ts_preferred(ctx))
                variant = 6;
            else if (apply_dir_preferred(ctx))
                variant = 4;
            else if (apply_rows_preferred(ctx))
                variant = 3;
        }
    }
    if (variant == 2 && !(P && P->usable))
        return fail(ctx, ASGFEM_ESTATE, "tiled operator plan not available for this pattern / multi-index set");
    if (variant == 2 && ranged) return fail(ctx, ASGFEM_ESTATE, "row ranges are not available for the tiled operator");
    if (r1 <= r0) return 0;
    ASG_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    if (variant == 7) {
        int rc = apply_ts2_launch(ctx, x, y, r0, r1);
        if (rc) return rc;
    } else if (variant == 6) {
        int rc = apply_ts_launch(ctx, x, y, r0, r1);
        if (rc) return rc;
    } else if (variant == 4 || variant == 5) {
        int rc = apply_dir_launch(ctx, x, y, variant == 5, r0, r1);
        if (rc) return rc;
    } else if (variant == 3) {
        int rc = apply_rows_launch(ctx, x, y, r0, r1);
        if (rc) return rc;
    } else if (variant == 1) {
        int bx = (int)std::min<int64_t>(128, ((ctx->ld + 31) / 32) * 32);
        int by = 256 / bx;
        dim3 block(bx, by);
        dim3 grid((unsigned)((r1 - r0 + by - 1) / by), (unsigned)((ctx->ld + bx - 1) / bx));
        k_apply_gather<<<grid, block, 0, ctx->stream>>>(r0, nrows, ctx->N, ctx->ld, ctx->nnz, ctx->d_rowptr, ctx->d_col,
                                                        ctx->d_vals, ctx->d_cptr, ctx->d_cm, ctx->d_cnu, ctx->d_cg,
                                                        ctx->d_bmask, x, y);
        ASG_CUDA(ctx, cudaGetLastError());
    } else {
        TiledArgs a;
        a.ld = ctx->ld;
        a.nnz = ctx->nnz;
        a.nblocks = P->nblocks;
        a.ntiles = P->ntiles;
        a.Cpad = P->Cpad;
        a.Rpad = P->Rpad;
        a.Smax = P->Smax;
        a.blk_rows = P->d_blk_rows;
        a.blk_colptr = P->d_blk_colptr;
        a.blk_cols = P->d_blk_cols;
        a.ell_lcol = P->d_ell_lcol;
        a.ell_pos = P->d_ell_pos;
        a.tile_t0 = P->d_tile_t0;
        a.tile_t = P->d_tile_t;
        a.tile_sptr = P->d_tile_sptr;
        a.tile_smodes = P->d_tile_smodes;
        a.tile_dptr = P->d_tile_dptr;
        a.dir_m = P->d_dir_m;
        a.dir_eptr = P->d_dir_eptr;
        a.ent_ds = P->d_ent_ds;
        a.ent_g = P->d_ent_g;
        a.vals = ctx->d_vals;
        a.bmask = ctx->d_bmask;
        a.x = x;
        a.y = y;
        int rc;
        switch (P->KW) {
            case 8: rc = launch_tiled<8>(ctx, a, P->smem_bytes); break;
            case 16: rc = launch_tiled<16>(ctx, a, P->smem_bytes); break;
            case 24: rc = launch_tiled<24>(ctx, a, P->smem_bytes); break;
            default: rc = launch_tiled<32>(ctx, a, P->smem_bytes); break;
        }
        if (rc) return rc;
    }
    ASG_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    return 0;
}

}  // namespace asgfem
