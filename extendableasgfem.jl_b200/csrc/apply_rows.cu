// variant 3 of the fused SGFE operator: one dof row per CTA iteration, (almost) all modes resident in shared memory.
//
//   Y[i, mu] = sum_k K_0[i,j_k] X[j_k,mu] + sum_{(m,nu) ~ mu} g sum_k K_m[i,j_k] X[j_k,nu]      (mul!, :101-117)
//
// For row i the <= KW rows X[j_k, :] it touches are staged ONCE into shared memory with 16-byte asynchronous copies
// (LDGSTS; the rows are contiguous 8*N-byte segments of the mode-fastest device layout), together with the (M+1) x KW
// values K_m[i, j_k].  Every thread then owns output modes mu (dst-major: no partial sums in memory, no atomics, the
// reference's summation order nu ascending / m ascending is kept), walks the coupling list of mu - kept in shared
// memory for the whole kernel in a rank-major (ELL) layout so that consecutive lanes read consecutive words - and
// gathers X[j_k, nu] and K_m[i, j_k] from shared memory.  In graded-lex / adaptively grown multi-index sets
// consecutive mu have consecutive neighbours nu, which keeps those gathers nearly bank-conflict free.
// If KW * N doubles do not fit, the modes are tiled (own range + halo columns), with the lists read from L2.
//
// Traffic: X is read KW/overlap times from L2 (neighbouring rows run concurrently on other SMs), once from HBM;
// K_m exactly once; Y written once, coalesced.
#include <algorithm>

#include "common.h"

namespace asgfem {

constexpr int ROWS_THREADS = 512;
constexpr uint32_t META_NONE = 0xffffffffu;

struct RowPlan {
    int KW = 0, Mp = 0, Sp = 0, ntiles = 0;
    bool meta_smem = false, usable = false;
    size_t smem_bytes = 0, meta_words = 0;
    int32_t *d_t0 = nullptr, *d_T = nullptr, *d_S = nullptr, *d_soff = nullptr, *d_smodes = nullptr;
    // per tile: [0] offset of the slice table, [1] #slices, [2] offset of the long-list table, [3] #long lists
    int32_t* d_tinfo = nullptr;
    // slice table: per slice (offset of its ELL block, nrank); long table: per long dst (dst, offset, length)
    // all offsets index the word array d_meta
    uint32_t* d_meta = nullptr;
    double* d_gtab = nullptr;
};
constexpr int LONG_LIST = 12;  // coupling lists longer than this are walked by a whole warp

static RowPlan* rp_of(asgfem_ctx* ctx) { return reinterpret_cast<RowPlan*>(ctx->rowplan); }

void apply_rows_free(asgfem_ctx* ctx) {
    RowPlan* P = rp_of(ctx);
    if (!P) return;
    void* ptrs[] = {P->d_t0, P->d_T, P->d_S, P->d_soff, P->d_smodes, P->d_tinfo, P->d_meta, P->d_gtab};
    for (void* q : ptrs)
        if (q) cudaFree(q);
    delete P;
    ctx->rowplan = nullptr;
}

int apply_rows_build(asgfem_ctx* ctx) {
    apply_rows_free(ctx);
    RowPlan* P = new RowPlan();
    ctx->rowplan = P;
    const int64_t N = ctx->N, nrows = ctx->n_owned >= 0 ? ctx->n_owned : ctx->n;
    const Coupling& C = ctx->coup;
    int KW = 1;
    for (int64_t i = 0; i < nrows; ++i) KW = std::max<int>(KW, (int)(ctx->h_rowptr[i + 1] - ctx->h_rowptr[i]));
    P->KW = KW;
    P->Mp = ctx->M + 1;
    if (ctx->M + 1 > 127 || ctx->mis.maxdeg() > 30) return 0;  // packing limits of the list words
    // coupling weight table: 0 -> 1 (mean term), 1 + deg -> g+(deg), 32 + deg -> g-(deg)
    std::vector<double> gtab(64, 0.0);
    gtab[0] = 1.0;
    for (size_t d = 0; d < ctx->gp.size() && d < 31; ++d) {
        gtab[1 + d] = ctx->gp[d];
        gtab[32 + d] = ctx->gm[d];
    }
    const int64_t budget = 227 * 1024 - 1536;
    const int64_t fixed = (int64_t)KW * P->Mp * 8 + 64 * 8;
    int64_t Smax = (budget - fixed) / (8 * (int64_t)KW);
    Smax = Smax / 2 * 2;
    if (Smax < 32) return 0;
    Smax = std::min<int64_t>(Smax, 8190);

    std::vector<int32_t> t0s, Ts, Ss, soff, smodes, tinfo;
    std::vector<uint32_t> meta;
    bool gtab_ok = true;
    std::vector<int32_t> mark((size_t)N, -1);
    int64_t mu = 0;
    int tile_id = 0;
    while (mu < N) {
        std::vector<int32_t> halo;
        int64_t t = 0;
        const int64_t Npad = (N + 1) / 2 * 2;
        if (Npad <= Smax) {
            t = N;  // everything resident: no halo
        } else {
            while (mu + t < N) {
                int64_t cand = mu + t;
                size_t before = halo.size();
                for (int32_t e = C.ptr[cand]; e < C.ptr[cand + 1]; ++e) {
                    int32_t nu = C.nu[e];
                    if (nu >= mu && nu <= cand) continue;
                    if (mark[nu] != tile_id) {
                        mark[nu] = tile_id;
                        halo.push_back(nu);
                    }
                }
                if ((int64_t)halo.size() + (t + 2) > Smax && t > 0) {
                    for (size_t k = before; k < halo.size(); ++k) mark[halo[k]] = -1;
                    halo.resize(before);
                    break;
                }
                ++t;
            }
            if (mu + t < N) t = std::max<int64_t>(2, t / 2 * 2);  // even tile starts keep the 16-byte copies aligned
            if ((int64_t)halo.size() + t > Smax) {
                *P = RowPlan();
                return 0;
            }
        }
        std::vector<int32_t> staged;
        const int64_t Tp = (t + 1) / 2 * 2;  // own part is copied in 16-byte units (padding column is zero / next mode)
        std::sort(halo.begin(), halo.end());
        std::vector<int32_t> halo2;
        for (int32_t h : halo)
            if (h < mu || h >= mu + t) halo2.push_back(h);
        std::unordered_map<int32_t, int32_t> spos;
        for (int64_t k = 0; k < t; ++k) spos[(int32_t)(mu + k)] = (int32_t)k;
        for (size_t k = 0; k < halo2.size(); ++k) spos[halo2[k]] = (int32_t)(Tp + k);
        // packed list words of every own mode: rank 0 = mean term, then the couplings in the order of mul!
        std::vector<std::vector<uint32_t>> lists((size_t)t);
        for (int64_t k = 0; k < t; ++k) {
            lists[k].push_back((uint32_t)k);  // src = own slot, m = 0, gidx = 0 (weight 1)
            for (int32_t e = C.ptr[mu + k]; e < C.ptr[mu + k + 1]; ++e) {
                int64_t deg = ctx->mis.mi[(mu + k) * ctx->mis.M + (C.m[e] - 1)];
                bool plus = ctx->mis.plus[(C.m[e] - 1) + ctx->mis.M * (mu + k)] == C.nu[e] + 1;
                uint32_t gidx = plus ? (uint32_t)(1 + deg) : (uint32_t)(32 + deg);
                if (gtab[gidx] != C.g[e]) gtab_ok = false;
                lists[k].push_back((uint32_t)spos[C.nu[e]] | ((uint32_t)C.m[e] << 13) | (gidx << 20));
            }
        }
        const int nslices = (int)((t + 31) / 32);
        std::vector<uint32_t> slicetab((size_t)nslices * 2), longtab, body;
        // word layout of a tile: [slice table][long table][ELL blocks ...][long lists ...]; offsets are tile-relative
        for (int sidx = 0; sidx < nslices; ++sidx) {
            int nr = 0;
            for (int l = 0; l < 32; ++l) {
                int64_t k = (int64_t)sidx * 32 + l;
                if (k < t && (int)lists[k].size() <= LONG_LIST) nr = std::max(nr, (int)lists[k].size());
            }
            slicetab[2 * sidx] = (uint32_t)body.size();
            slicetab[2 * sidx + 1] = (uint32_t)nr;
            size_t at = body.size();
            body.resize(at + (size_t)nr * 32, META_NONE);
            for (int l = 0; l < 32; ++l) {
                int64_t k = (int64_t)sidx * 32 + l;
                if (k >= t || (int)lists[k].size() > LONG_LIST) continue;
                for (size_t r = 0; r < lists[k].size(); ++r) body[at + r * 32 + l] = lists[k][r];
            }
        }
        for (int64_t k = 0; k < t; ++k)
            if ((int)lists[k].size() > LONG_LIST) {
                longtab.push_back((uint32_t)k);
                longtab.push_back((uint32_t)body.size());
                longtab.push_back((uint32_t)lists[k].size());
                body.insert(body.end(), lists[k].begin(), lists[k].end());
            }
        const uint32_t tile_base = (uint32_t)meta.size();
        const uint32_t body_base = tile_base + (uint32_t)slicetab.size() + (uint32_t)longtab.size();
        for (int sidx = 0; sidx < nslices; ++sidx) slicetab[2 * sidx] += body_base;
        for (size_t q = 0; q < longtab.size(); q += 3) longtab[q + 1] += body_base;
        tinfo.push_back((int32_t)tile_base);
        tinfo.push_back(nslices);
        tinfo.push_back((int32_t)(tile_base + slicetab.size()));
        tinfo.push_back((int32_t)(longtab.size() / 3));
        meta.insert(meta.end(), slicetab.begin(), slicetab.end());
        meta.insert(meta.end(), longtab.begin(), longtab.end());
        meta.insert(meta.end(), body.begin(), body.end());
        t0s.push_back((int32_t)mu);
        Ts.push_back((int32_t)t);
        Ss.push_back((int32_t)(Tp + halo2.size()));
        soff.push_back((int32_t)smodes.size());
        smodes.insert(smodes.end(), halo2.begin(), halo2.end());  // only the halo part needs a list
        mu += t;
        ++tile_id;
    }
    P->ntiles = tile_id;
    if (!gtab_ok) {
        *P = RowPlan();
        return 0;
    }
    int Sp = 0;
    for (int32_t s : Ss) Sp = std::max(Sp, (int)s);
    P->Sp = (Sp + 1) / 2 * 2;
    P->meta_words = meta.size();
    size_t base_bytes = (size_t)KW * P->Sp * 8 + (size_t)fixed;
    P->meta_smem = P->ntiles == 1 && base_bytes + meta.size() * 4 <= (size_t)budget;
    P->smem_bytes = base_bytes + (P->meta_smem ? meta.size() * 4 : 0);
    int rc = 0;
    rc |= dev_upload(ctx, &P->d_t0, t0s);
    rc |= dev_upload(ctx, &P->d_T, Ts);
    rc |= dev_upload(ctx, &P->d_S, Ss);
    rc |= dev_upload(ctx, &P->d_soff, soff);
    rc |= dev_upload(ctx, &P->d_smodes, smodes);
    rc |= dev_upload(ctx, &P->d_tinfo, tinfo);
    rc |= dev_upload(ctx, &P->d_meta, meta);
    rc |= dev_upload(ctx, &P->d_gtab, gtab);
    if (rc) return rc;
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    P->usable = true;
    return 0;
}

// true if the row-resident kernel can keep all modes and the coupling lists in shared memory (its fast path)
bool apply_rows_preferred(asgfem_ctx* ctx) {
    if (!rp_of(ctx) && apply_rows_build(ctx)) return false;
    RowPlan* P = rp_of(ctx);
    return P && P->usable && P->ntiles == 1 && P->meta_smem;
}

struct RowArgs {
    int64_t row0, nrows, ld, nnz;  // rows [row0, nrows)
    int M, Mp, Sp, KW, ntiles;
    const int64_t* rowptr;
    const int32_t* col;
    const double* vals;
    const uint8_t* bmask;
    const int32_t *t0, *T, *S, *soff, *smodes, *tinfo;
    const uint32_t* meta;
    size_t meta_words;
    const double* gtab;
    const double* x;
    double* y;
};

__device__ __forceinline__ void cpa16(void* smem_dst, const void* gsrc) {
    unsigned saddr = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(saddr), "l"(gsrc));
}
__device__ __forceinline__ void cpa8(void* smem_dst, const void* gsrc) {
    unsigned saddr = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(saddr), "l"(gsrc));
}
__device__ __forceinline__ void cpa_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory"); }

template <bool META_SMEM>
__global__ void __launch_bounds__(ROWS_THREADS, 1) k_apply_rows(RowArgs a) {
    extern __shared__ __align__(16) double smem[];
    double* Xs = smem;                          // [KW][Sp]
    double* Ks = Xs + (size_t)a.KW * a.Sp;      // [KW][Mp]
    double* gt = Ks + (size_t)a.KW * a.Mp;      // [64]
    uint32_t* ms = reinterpret_cast<uint32_t*>(gt + 64);
    const int tid = threadIdx.x;
    if (tid < 64) gt[tid] = a.gtab[tid];
    if (META_SMEM)
        for (size_t k = tid; k < a.meta_words; k += ROWS_THREADS) ms[k] = a.meta[k];

    for (int64_t row = a.row0 + blockIdx.x; row < a.nrows; row += gridDim.x) {
        const int64_t rp = a.rowptr[row];
        const int len = (int)(a.rowptr[row + 1] - rp);
        const bool masked = a.bmask[row] != 0;
        for (int tile = 0; tile < a.ntiles; ++tile) {
            const int t0 = a.t0[tile], T = a.T[tile], S = a.S[tile];
            const int Tp = (T + 1) & ~1;
            __syncthreads();  // previous row / tile fully consumed
            if (!masked) {
                if (tile == 0)
                    for (int idx = tid; idx < len * (a.M + 1); idx += ROWS_THREADS) {
                        int m = idx / len, k = idx - m * len;
                        Ks[k * a.Mp + m] = __ldg(a.vals + (int64_t)m * a.nnz + rp + k);
                    }
                const int half = Tp >> 1;
                for (int idx = tid; idx < len * half; idx += ROWS_THREADS) {
                    int k = idx / half, s2 = idx - k * half;
                    const double* src = a.x + (int64_t)a.col[rp + k] * a.ld + t0 + 2 * s2;
                    cpa16(Xs + (size_t)k * a.Sp + 2 * s2, src);
                }
                const int nh = S - Tp;
                const int32_t* hm = a.smodes + a.soff[tile];
                for (int idx = tid; idx < len * nh; idx += ROWS_THREADS) {
                    int k = idx / nh, h = idx - k * nh;
                    cpa8(Xs + (size_t)k * a.Sp + Tp + h, a.x + (int64_t)a.col[rp + k] * a.ld + hm[h]);
                }
                cpa_wait_all();
            }
            __syncthreads();
            const uint32_t* mt = META_SMEM ? ms : a.meta;  // list words: shared memory if resident, else L2
            const int lane = tid & 31, warp = tid >> 5;
            const int* ti = a.tinfo + 4 * tile;
            const int slice_tab = ti[0], nslices = ti[1], long_tab = ti[2], nlong = ti[3];
            double* yr = a.y + row * a.ld + t0;
            // (A) short lists: lane = output mode, rank-major (ELL) words of the 32-mode slice
            for (int sl = warp; sl < nslices; sl += ROWS_THREADS / 32) {
                const int d = sl * 32 + lane;
                const uint32_t off = mt[slice_tab + 2 * sl], nr = mt[slice_tab + 2 * sl + 1];
                double acc = 0.0;
                bool mine = false;
                if (!masked) {
                    for (uint32_t r = 0; r < nr; ++r) {
                        const uint32_t w = mt[off + r * 32 + lane];
                        if (w == META_NONE) break;
                        mine = true;
                        const double* xs = Xs + (w & 0x1fffu);
                        const double* ks = Ks + ((w >> 13) & 0x7fu);
                        double t = 0.0;
#pragma unroll 4
                        for (int k = 0; k < len; ++k) t = fma(ks[k * a.Mp], xs[(size_t)k * a.Sp], t);
                        acc = fma(gt[w >> 20], t, acc);
                    }
                } else {
                    mine = nr > 0 && mt[off + lane] != META_NONE;
                }
                if (d < T && mine) yr[d] = acc;
            }
            // (B) long lists (hub modes): one warp per output mode, lanes over its couplings, shuffle reduction
            for (int li = warp; li < nlong; li += ROWS_THREADS / 32) {
                const uint32_t d = mt[long_tab + 3 * li], off = mt[long_tab + 3 * li + 1], n = mt[long_tab + 3 * li + 2];
                double acc = 0.0;
                if (!masked) {
                    for (uint32_t e = lane; e < n; e += 32) {
                        const uint32_t w = mt[off + e];
                        const double* xs = Xs + (w & 0x1fffu);
                        const double* ks = Ks + ((w >> 13) & 0x7fu);
                        double t = 0.0;
#pragma unroll 4
                        for (int k = 0; k < len; ++k) t = fma(ks[k * a.Mp], xs[(size_t)k * a.Sp], t);
                        acc = fma(gt[w >> 20], t, acc);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                }
                if (lane == 0) yr[d] = acc;
            }
        }
    }
}

int apply_rows_launch(asgfem_ctx* ctx, const double* x, double* y, int64_t r0, int64_t r1) {
    RowPlan* P = rp_of(ctx);
    if (!P) {
        int rc = apply_rows_build(ctx);
        if (rc) return rc;
        P = rp_of(ctx);
    }
    if (!P->usable) return fail(ctx, ASGFEM_ESTATE, "row-resident operator plan not available (row too long or too many KLE terms)");
    RowArgs a;
    a.row0 = r0;
    a.nrows = r1;
    a.ld = ctx->ld;
    a.nnz = ctx->nnz;
    a.M = ctx->M;
    a.Mp = P->Mp;
    a.Sp = P->Sp;
    a.KW = P->KW;
    a.ntiles = P->ntiles;
    a.rowptr = ctx->d_rowptr;
    a.col = ctx->d_col;
    a.vals = ctx->d_vals;
    a.bmask = ctx->d_bmask;
    a.t0 = P->d_t0;
    a.T = P->d_T;
    a.S = P->d_S;
    a.soff = P->d_soff;
    a.smodes = P->d_smodes;
    a.tinfo = P->d_tinfo;
    a.meta = P->d_meta;
    a.meta_words = P->meta_words;
    a.gtab = P->d_gtab;
    a.x = x;
    a.y = y;
    if (r1 <= r0) return 0;
    int grid = (int)std::min<int64_t>(r1 - r0, 148);
    if (P->meta_smem) {
        ASG_CUDA(ctx, cudaFuncSetAttribute(k_apply_rows<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        k_apply_rows<true><<<grid, ROWS_THREADS, P->smem_bytes, ctx->stream>>>(a);
    } else {
        ASG_CUDA(ctx, cudaFuncSetAttribute(k_apply_rows<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        k_apply_rows<false><<<grid, ROWS_THREADS, P->smem_bytes, ctx->stream>>>(a);
    }
    ASG_CUDA(ctx, cudaGetLastError());
    return 0;
}

}  // namespace asgfem
