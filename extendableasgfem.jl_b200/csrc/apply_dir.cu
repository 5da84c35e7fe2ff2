// variant 4 of the fused SGFE operator: row-resident staging (as variant 3) + direction-major contraction.
//
// For one dof row i the rows X[j_k, :] (k < row length) are staged once in shared memory.  The couplings are then
// processed direction by direction: a warp takes a unit = (direction m, chunk range of its couplings mu <- nu), loads
// the KW values K_m[i, j_k] ONCE into registers (warp-uniform operand), and every lane evaluates
//     t = sum_k K_m[i,j_k] * X[j_k, nu]          (one 8-byte shared load per FMA instead of two)
// for its coupling and adds g*t into the shared accumulator Y[mu].  Couplings of one direction and sign have distinct
// targets; different warps may hit the same target from different directions, hence the shared-memory atomic add
// (CAS loop on sm_100a; contention is rare).  The mean term is direction 0 with nu = mu, g = 1.
// The summation order over directions therefore varies from run to run at rounding level (1e-16 relative); use
// variant 3 or 1 when bit-reproducibility between runs matters.
//
// Coupling words (32 bit, resident in shared memory for the whole kernel): dst | src << 13 | gidx << 26.
#include <algorithm>

#include "common.h"

namespace asgfem {

constexpr int DIR_THREADS = 512;
constexpr int DIR_WARPS = DIR_THREADS / 32;
constexpr int UNIT_CHUNKS = 8;
constexpr int DEU = 4;  // couplings per lane processed together (ILP)  // couplings per unit = 32 * UNIT_CHUNKS

struct DirPlan {
    int nslot = 0, KW = 0, Mp = 0, Np = 0, nunits = 0, nunit_words = 0;
    size_t nwords = 0, smem_bytes = 0;
    bool usable = false, owned = false;
    uint32_t* d_words = nullptr;   // coupling words, unit after unit
    int32_t* d_units = nullptr;    // per unit: m, first word, #words
    int32_t* d_wptr = nullptr;     // DIR_WARPS+1: units of warp w = [wptr[w], wptr[w+1])
    double* d_gtab = nullptr;
};

static DirPlan* dp_of(asgfem_ctx* ctx) { return reinterpret_cast<DirPlan*>(ctx->dirplan); }

void apply_dir_free(asgfem_ctx* ctx) {
    DirPlan* P = dp_of(ctx);
    if (!P) return;
    void* ptrs[] = {P->d_words, P->d_units, P->d_wptr, P->d_gtab};
    for (void* q : ptrs)
        if (q) cudaFree(q);
    delete P;
    ctx->dirplan = nullptr;
}

int apply_dir_build(asgfem_ctx* ctx, bool owned) {
    apply_dir_free(ctx);
    DirPlan* P = new DirPlan();
    ctx->dirplan = P;
    P->owned = owned;
    const int64_t N = ctx->N, nrows = ctx->n_owned >= 0 ? ctx->n_owned : ctx->n;
    const Coupling& C = ctx->coup;
    if (N > 8192 || ctx->M + 1 > 255 || ctx->mis.maxdeg() > 30) return 0;
    int nslot = 1;
    for (int64_t i = 0; i < nrows; ++i) nslot = std::max<int>(nslot, (int)(ctx->h_rowptr[i + 1] - ctx->h_rowptr[i]));
    if (nslot > 32) return 0;
    P->nslot = nslot;
    {
        const int avail[] = {4, 6, 7, 8, 9, 10, 12, 14, 16, 20, 24, 32};
        P->KW = 32;
        for (int v : avail)
            if (v >= nslot) {
                P->KW = v;
                break;
            }
    }
    nslot = P->KW;  // shared memory is sized for the instantiated slot count
    P->nslot = nslot;
    P->Mp = ctx->M + 1;
    P->Np = (int)((N + 1) / 2 * 2);
    std::vector<double> gtab(64, 0.0);
    gtab[0] = 1.0;
    for (size_t d = 0; d < ctx->gp.size() && d < 31; ++d) {
        gtab[1 + d] = ctx->gp[d];
        gtab[32 + d] = ctx->gm[d];
    }
    // couplings grouped by (direction, sign): targets are distinct inside a group
    std::vector<std::vector<uint32_t>> groups((size_t)(2 * ctx->M + 1));
    for (int64_t mu = 0; mu < N; ++mu) groups[0].push_back((uint32_t)mu | ((uint32_t)mu << 13));
    for (int64_t mu = 0; mu < N; ++mu)
        for (int32_t e = C.ptr[mu]; e < C.ptr[mu + 1]; ++e) {
            int m = C.m[e];
            int64_t deg = ctx->mis.mi[mu * ctx->mis.M + (m - 1)];
            bool plus = ctx->mis.plus[(m - 1) + ctx->mis.M * mu] == C.nu[e] + 1;
            uint32_t gidx = plus ? (uint32_t)(1 + deg) : (uint32_t)(32 + deg);
            if (gtab[gidx] != C.g[e]) return 0;
            groups[(size_t)(2 * m - (plus ? 1 : 0))].push_back((uint32_t)mu | ((uint32_t)C.nu[e] << 13) | (gidx << 26));
        }
    struct Unit {
        int m, first, count;
    };
    std::vector<Unit> units;
    std::vector<uint32_t> words;
    std::vector<std::vector<int>> per_warp(DIR_WARPS);
    if (!owned) {
        // atomic mode: units = slices of one (direction, sign) group, balanced over the warps (largest first)
        for (size_t gi = 0; gi < groups.size(); ++gi) {
            int m = gi == 0 ? 0 : (int)((gi + 1) / 2);
            const auto& g = groups[gi];
            for (size_t at = 0; at < g.size(); at += 32 * UNIT_CHUNKS) {
                size_t cnt = std::min<size_t>(32 * UNIT_CHUNKS, g.size() - at);
                units.push_back({m, (int)words.size(), (int)cnt});
                words.insert(words.end(), g.begin() + (long)at, g.begin() + (long)(at + cnt));
            }
        }
        std::vector<int> order(units.size());
        for (size_t k = 0; k < order.size(); ++k) order[k] = (int)k;
        std::sort(order.begin(), order.end(), [&](int a, int b) { return units[a].count > units[b].count; });
        std::vector<int64_t> load(DIR_WARPS, 0);
        for (int u : order) {
            int w = (int)(std::min_element(load.begin(), load.end()) - load.begin());
            per_warp[w].push_back(u);
            load[w] += (units[u].count + 31) / 32 * 32 + 48;
        }
        for (int w = 0; w < DIR_WARPS; ++w) std::sort(per_warp[w].begin(), per_warp[w].end());
    } else {
        // owned mode: every warp owns a contiguous range of target modes (balanced by number of couplings) and
        // processes, direction by direction, exactly the couplings that end in its range -> plain read-modify-write
        // of the shared accumulators, fixed summation order, no atomics
        std::vector<int64_t> cnt((size_t)N, 1);
        for (int64_t mu = 0; mu < N; ++mu) cnt[mu] += C.ptr[mu + 1] - C.ptr[mu];
        int64_t total = 0;
        for (int64_t v : cnt) total += v;
        std::vector<int> owner((size_t)N, 0);
        {
            int64_t acc = 0;
            int w = 0;
            for (int64_t mu = 0; mu < N; ++mu) {
                while (w + 1 < DIR_WARPS && acc >= (total * (w + 1)) / DIR_WARPS) ++w;
                owner[mu] = w;
                acc += cnt[mu];
            }
        }
        for (int w = 0; w < DIR_WARPS; ++w)
            for (size_t gi = 0; gi < groups.size(); ++gi) {
                int m = gi == 0 ? 0 : (int)((gi + 1) / 2);
                int first = (int)words.size();
                for (uint32_t wd : groups[gi])
                    if (owner[wd & 0x1fffu] == w) words.push_back(wd);
                int count = (int)words.size() - first;
                if (count > 0) {
                    per_warp[w].push_back((int)units.size());
                    units.push_back({m, first, count});
                }
            }
    }
    std::vector<int32_t> wptr(1, 0), ulist;
    for (int w = 0; w < DIR_WARPS; ++w) {
        for (int u : per_warp[w]) {
            ulist.push_back(units[u].m);
            ulist.push_back(units[u].first);
            ulist.push_back(units[u].count);
        }
        wptr.push_back((int32_t)(ulist.size() / 3));
    }
    P->nunits = (int)units.size();
    P->nunit_words = (int)ulist.size();
    P->nwords = words.size();
    P->smem_bytes = (size_t)nslot * P->Np * 8 + (size_t)P->Np * 8 + (size_t)nslot * P->Mp * 8 + 64 * 8 + words.size() * 4 +
                    ulist.size() * 4 + 16;
    if (P->smem_bytes > 227 * 1024 - 1024) return 0;
    int rc = 0;
    rc |= dev_upload(ctx, &P->d_words, words);
    rc |= dev_upload(ctx, &P->d_units, ulist);
    rc |= dev_upload(ctx, &P->d_wptr, wptr);
    rc |= dev_upload(ctx, &P->d_gtab, gtab);
    if (rc) return rc;
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    P->usable = true;
    return 0;
}

bool apply_dir_preferred(asgfem_ctx* ctx) {
    if (!dp_of(ctx) && apply_dir_build(ctx, false)) return false;
    DirPlan* P = dp_of(ctx);
    return P && P->usable;
}

struct DirArgs {
    int64_t row0, nrows, ld, nnz;  // rows [row0, nrows)
    int N, Np, M, Mp, nslot, nunit_words;
    size_t nwords;
    const int64_t* rowptr;
    const int32_t* col;
    const double* vals;
    const uint8_t* bmask;
    const uint32_t* words;
    const int32_t *units, *wptr;
    const double* gtab;
    const double* x;
    double* y;
};

__device__ __forceinline__ void dcp16(void* smem_dst, const void* gsrc) {
    unsigned saddr = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(saddr), "l"(gsrc));
}
__device__ __forceinline__ void dcp_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory"); }

__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned lds_u32(unsigned addr) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// NS = number of column slots = longest row of the pattern (rows shorter than NS see zero K entries and stale,
// finite X slots).  All shared-memory operands are addressed with 32-bit shared-window addresses.
template <int NS, bool ATOMIC>
__global__ void __launch_bounds__(DIR_THREADS, 1) k_apply_dir(DirArgs a) {
    extern __shared__ __align__(16) double smem[];
    double* Xs = smem;                              // [NS][Np]
    double* Ys = Xs + (size_t)NS * a.Np;            // [Np]
    double* Ks = Ys + a.Np;                         // [NS][Mp]
    double* gt = Ks + (size_t)NS * a.Mp;            // [64]
    uint32_t* ws = reinterpret_cast<uint32_t*>(gt + 64);      // coupling words
    int32_t* us = reinterpret_cast<int32_t*>(ws + a.nwords);  // unit table (m, first, count) in warp order
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 64) gt[tid] = a.gtab[tid];
    for (int k = tid; k < (int)a.nwords; k += DIR_THREADS) ws[k] = a.words[k];
    for (int k = tid; k < a.nunit_words; k += DIR_THREADS) us[k] = a.units[k];
    for (int k = tid; k < NS * a.Np; k += DIR_THREADS) Xs[k] = 0.0;  // stale slots must stay finite
    const int u0 = a.wptr[warp], u1 = a.wptr[warp + 1];
    const int half = a.Np >> 1;
    const unsigned xs_base = (unsigned)__cvta_generic_to_shared(Xs);
    const unsigned ws_base = (unsigned)__cvta_generic_to_shared(ws);
    const unsigned gt_base = (unsigned)__cvta_generic_to_shared(gt);
    const unsigned xstride = (unsigned)a.Np * 8u;

    for (int64_t row = a.row0 + blockIdx.x; row < a.nrows; row += gridDim.x) {
        const int64_t rp = a.rowptr[row];
        const int len = (int)(a.rowptr[row + 1] - rp);
        const bool masked = a.bmask[row] != 0;
        __syncthreads();  // previous row fully written out
        if (!masked) {
            for (int idx = tid; idx < NS * (a.M + 1); idx += DIR_THREADS) {
                int m = idx / NS, k = idx - m * NS;
                Ks[k * a.Mp + m] = k < len ? __ldg(a.vals + (int64_t)m * a.nnz + rp + k) : 0.0;
            }
            for (int idx = tid; idx < len * half; idx += DIR_THREADS) {
                int k = idx / half, s2 = idx - k * half;
                dcp16(Xs + k * a.Np + 2 * s2, a.x + (int64_t)a.col[rp + k] * a.ld + 2 * s2);
            }
        }
        for (int k = tid; k < a.Np; k += DIR_THREADS) Ys[k] = 0.0;
        dcp_wait_all();
        __syncthreads();
        if (!masked) {
            int mprev = -1;
            double kr[NS];
            for (int u = u0; u < u1; ++u) {
                const int m = us[3 * u], first = us[3 * u + 1], count = us[3 * u + 2];
                if (m != mprev) {
#pragma unroll
                    for (int k = 0; k < NS; ++k) kr[k] = Ks[k * a.Mp + m];
                    mprev = m;
                }
                const unsigned wa = ws_base + 4u * (unsigned)(first + lane);
                // DEU couplings per lane and iteration: their loads and FMA chains are independent (ILP), the
                // accumulator updates follow in order
                for (int e = lane; e < count; e += 32 * DEU) {
                    unsigned w[DEU], xa[DEU];
                    double t[DEU];
#pragma unroll
                    for (int q = 0; q < DEU; ++q) {
                        const bool on = e + 32 * q < count;
                        w[q] = on ? lds_u32(wa + 4u * (unsigned)(e - lane + 32 * q)) : 0u;
                        xa[q] = xs_base + ((w[q] >> 10) & 0xfff8u);  // src * 8 (inactive lanes read slot 0)
                        t[q] = 0.0;
                    }
#pragma unroll
                    for (int k = 0; k < NS; ++k) {
#pragma unroll
                        for (int q = 0; q < DEU; ++q) {
                            t[q] = fma(kr[k], lds_f64(xa[q]), t[q]);
                            xa[q] += xstride;
                        }
                    }
#pragma unroll
                    for (int q = 0; q < DEU; ++q) {
                        if (e + 32 * q < count) {
                            const double g = lds_f64(gt_base + ((w[q] >> 23) & 0x1f8u));
                            if (ATOMIC) {
                                atomicAdd(Ys + (w[q] & 0x1fffu), g * t[q]);
                            } else {
                                double* yp = Ys + (w[q] & 0x1fffu);  // target owned by this warp, distinct inside the unit
                                *yp = fma(g, t[q], *yp);
                            }
                        }
                    }
                }
                if (!ATOMIC) __syncwarp();  // the next unit may revisit a target from another lane
            }
        }
        __syncthreads();
        double* yr = a.y + row * a.ld;
        for (int k = tid; k < a.N; k += DIR_THREADS) yr[k] = Ys[k];
    }
}

int apply_dir_launch(asgfem_ctx* ctx, const double* x, double* y, bool owned, int64_t r0, int64_t r1) {
    DirPlan* P = dp_of(ctx);
    if (!P || P->owned != owned) {
        int rc = apply_dir_build(ctx, owned);
        if (rc) return rc;
        P = dp_of(ctx);
    }
    if (!P->usable) return fail(ctx, ASGFEM_ESTATE, "direction-major operator plan not available (too many modes / row too long)");
    DirArgs a;
    a.row0 = r0;
    a.nrows = r1;
    a.ld = ctx->ld;
    a.nnz = ctx->nnz;
    a.N = (int)ctx->N;
    a.Np = P->Np;
    a.M = ctx->M;
    a.Mp = P->Mp;
    a.nslot = P->nslot;
    a.nwords = P->nwords;
    a.rowptr = ctx->d_rowptr;
    a.col = ctx->d_col;
    a.vals = ctx->d_vals;
    a.bmask = ctx->d_bmask;
    a.words = P->d_words;
    a.units = P->d_units;
    a.wptr = P->d_wptr;
    a.gtab = P->d_gtab;
    a.x = x;
    a.y = y;
    if (r1 <= r0) return 0;
    int grid = (int)std::min<int64_t>(r1 - r0, 148);
#define LAUNCH_DIR(KWV)                                                                                              \
    do {                                                                                                             \
        if (owned) {                                                                                                 \
            ASG_CUDA(ctx, cudaFuncSetAttribute(k_apply_dir<KWV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
            k_apply_dir<KWV, false><<<grid, DIR_THREADS, P->smem_bytes, ctx->stream>>>(a);                           \
        } else {                                                                                                     \
            ASG_CUDA(ctx, cudaFuncSetAttribute(k_apply_dir<KWV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
            k_apply_dir<KWV, true><<<grid, DIR_THREADS, P->smem_bytes, ctx->stream>>>(a);                            \
        }                                                                                                            \
    } while (0)
    a.nunit_words = P->nunit_words;
    switch (P->KW) {
        case 4: LAUNCH_DIR(4); break;
        case 6: LAUNCH_DIR(6); break;
        case 7: LAUNCH_DIR(7); break;
        case 8: LAUNCH_DIR(8); break;
        case 9: LAUNCH_DIR(9); break;
        case 10: LAUNCH_DIR(10); break;
        case 12: LAUNCH_DIR(12); break;
        case 14: LAUNCH_DIR(14); break;
        case 16: LAUNCH_DIR(16); break;
        case 20: LAUNCH_DIR(20); break;
        case 24: LAUNCH_DIR(24); break;
        default: LAUNCH_DIR(32); break;
    }
#undef LAUNCH_DIR
    ASG_CUDA(ctx, cudaGetLastError());
    return 0;
}

}  // namespace asgfem
