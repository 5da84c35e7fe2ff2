// variant 9 of the fused SGFE operator (default): block products on the fp64 MMA path with a LIST exchange.
//
//   Y[i, mu] = sum_k K_0[i,j_k] X[j_k,mu] + sum_{(m,nu) ~ mu} g sum_k K_m[i,j_k] X[j_k,nu]      (mul!, :101-117)
//
// Same mode-side plan as variant 8 (apply_mma.cu: home blocks of 8 modes, D-sets of 8 directions, pairs of blocks =
// 16 consecutive device columns), other data movement - chosen from the ncu counters of variants 7 and 8
// (profiles/README.md: both are bound by shared-memory wavefronts and instruction issue, not by DRAM or the fp64 pipe):
//
//   * PRODUCE.  A step = (pair of home blocks, k-th D-set of each): 2 KS DMMA m8n8k4 (A = K rows of the 8 directions,
//     B = X rows of the 8 modes) give T[dir, mode] for 2 x 8 x 8 pairs.  All steps of a pair run in ONE warp, so the X
//     fragments of the pair are loaded (coalesced 16-byte loads, 4 lines per instruction) and shuffled into fragment
//     order once per dof row.  The K fragments of a step come with one 16-byte shared-memory load per block from the
//     staged K rows (layout [direction][k mod 4][k / 4]); the byte offsets of the 8 directions of a step are one 4-byte
//     table load.  The outputs are stored in FRAGMENT ORDER: one conflict-free 16-byte store per block and lane, no
//     address words, no predicates.  Unneeded entries of a block (about half) are simply never read.
//   * CONSUME.  Thread = device column (4 groups of 32 consecutive columns per warp at N = 2000).  The needed products of
//     a column are a LIST of 16-bit indices into the T buffer of the pass, rows of 32 lanes, ordered by weight so that
//     the weight of a list row is warp-uniform (g takes few values): Y += w_row * T[idx].  Products with two consumers
//     are simply read twice.  The order of a lane's entries inside a weight class is chosen on the host so that the lanes
//     of a half-warp read different banks where possible.  Fixed summation order, no atomics.
//   * PIPELINE.  The dof row is processed in P passes over the pairs (two T buffers fit the 227 KB); stage t = (row,
//     pass): a warp produces its pairs of stage t into buffer t & 1, loads the X fragments of stage t + 1 into the
//     registers it has just used, and sums its lists of stage t - 1 from the other buffer.  One block barrier per stage.
//     Row records (CSR position, length, columns) and the K rows of the next row arrive by cp.async one / two rows ahead.
#include <algorithm>
#include <cstring>
#include <map>
#include <numeric>

#include "apply_mma_plan.h"
#include "common.h"

namespace asgfem {

namespace {

struct BlkPlan {
    bool usable = false;
    int KS = 0, P = 1, NPW = 0, NG = 0, W = 16;
    bool lists_global = false;
    bool single = false;  // one pass with a single T buffer (two barriers per dof row)
    uint32_t nwords = 0, off_pair = 0, off_crec = 0, off_gcol = 0, off_pbase = 0, off_srec = 0, off_list = 0, off_roww = 0, tbuf_doubles = 0;
    size_t smem_bytes = 0;
    uint32_t* d_blob = nullptr;
    double* d_zero = nullptr;
    int32_t* d_rowmeta = nullptr;
    int grid = 148;
    double dmma_per_row = 0, list_rows = 0, list_ideal = 0;
};

BlkPlan* bp_of(asgfem_ctx* ctx) {
    MmaPlan* P = mp_of(ctx);
    return P ? reinterpret_cast<BlkPlan*>(P->blk) : nullptr;
}

void blk_free(BlkPlan* B) {
    if (!B) return;
    if (B->d_blob) cudaFree(B->d_blob);
    if (B->d_zero) cudaFree(B->d_zero);
    if (B->d_rowmeta) cudaFree(B->d_rowmeta);
    delete B;
}

inline uint64_t wbits(double w) {
    uint64_t b;
    std::memcpy(&b, &w, 8);
    return b;
}

}  // namespace

void apply_blk_free(asgfem_ctx* ctx) {
    MmaPlan* P = mp_of(ctx);
    if (!P || !P->blk) return;
    blk_free(reinterpret_cast<BlkPlan*>(P->blk));
    P->blk = nullptr;
}

bool apply_blk_usable(asgfem_ctx* ctx) {
    BlkPlan* B = bp_of(ctx);
    return B && B->usable;
}

// ---------------------------------------------------------------------------------------------------------------------
// kernel tables
// ---------------------------------------------------------------------------------------------------------------------
int apply_blk_build(asgfem_ctx* ctx) {
    apply_blk_free(ctx);
    MmaPlan* P = mp_of(ctx);
    if (!P || !P->layout_ok) return 0;
    const bool verbose = getenv("ASGFEM_BLK_VERBOSE") != nullptr;
    const int64_t nrows = ctx->n_owned >= 0 ? ctx->n_owned : ctx->n;
    const int M = ctx->M, Mp = M + 1;
    if (P->M > M || nrows <= 0) return 0;
    int maxlen = 1;
    for (int64_t i = 0; i < nrows; ++i) maxlen = std::max<int>(maxlen, (int)(ctx->h_rowptr[i + 1] - ctx->h_rowptr[i]));
    const int KS = maxlen <= 8 ? 2 : maxlen <= 16 ? 4 : maxlen <= 24 ? 6 : 0;
    if (!KS) return 0;
    const uint32_t krow_bytes = 32u * (uint32_t)KS;  // one direction: [k mod 4][k / 4]
    if ((uint64_t)(Mp + 1) * krow_bytes > 65535u) return 0;
    const int W = 16;  // 32 warps with half the registers measured slower (74 vs 55 ms at config 4)
    const int NPWcap = std::max(1, 16 / KS);  // X fragments of a stage: NPW * KS double2 per lane
    const int npairs = (int)P->pairs.size(), ngroups = npairs / 2, ncols = P->ncols;
    int NG = (ngroups + W - 1) / W;
    if (NG > 16) return 0;
    NG = NG <= 2 ? 2 : NG <= 4 ? 4 : NG <= 8 ? 8 : 16;

    // ---- D-sets with the directions of a set ordered so that rows (2j, 2j+1) have different parity where possible: the
    //      quarter-warp of a 16-byte K fragment load then covers two different 64-byte halves of a 128-byte line (KS = 2)
    std::vector<std::array<int, 8>> dsets = P->dsets;
    for (auto& d : dsets) {
        std::vector<int> ev, od;
        for (int r = 0; r < 8; ++r) {
            const int dir = d[(size_t)r] < 0 ? Mp : d[(size_t)r];
            (dir & 1 ? od : ev).push_back(d[(size_t)r]);
        }
        std::array<int, 8> o;
        size_t ie = 0, io = 0;
        for (int j = 0; j < 4; ++j) {
            // one even and one odd direction while both kinds are left
            if (ie < ev.size() && io < od.size()) {
                o[(size_t)(2 * j)] = ev[ie++];
                o[(size_t)(2 * j + 1)] = od[io++];
            } else if (ie < ev.size()) {
                o[(size_t)(2 * j)] = ev[ie++];
                o[(size_t)(2 * j + 1)] = ev[ie++];
            } else {
                o[(size_t)(2 * j)] = od[io++];
                o[(size_t)(2 * j + 1)] = od[io++];
            }
        }
        d = o;
    }

    int total_steps = 0;
    for (auto& ps : P->pairs) total_steps += ps.nsteps;
    if (total_steps == 0) return 0;

    auto align4 = [](uint32_t v) { return (v + 3u) & ~3u; };
    bool try_single = true;
    if (const char* e = getenv("ASGFEM_BLK_SINGLE")) try_single = atoi(e) != 0;
    // trial = (lists in shared memory, passes 1..8), then (lists in global memory, passes 1..64): the lists of many modes
    // (config 5: 37 k items) do not fit next to two T buffers
    for (int trial = 0; trial < 72; ++trial) {
        const bool lists_global = trial >= 8;
        const int npass = lists_global ? trial - 7 : trial + 1;
        // ---- pairs -> passes (contiguous ranges of the sorted pair list with about equal step counts), then -> warps ----
        std::vector<int> pass_of((size_t)npairs, 0);
        {
            long run = 0;
            for (int q = 0; q < npairs; ++q) {
                pass_of[(size_t)q] = (int)std::min<long>(npass - 1, run * npass / total_steps);
                run += P->pairs[(size_t)q].nsteps;
            }
        }
        std::vector<std::vector<int>> bins((size_t)npass * W);  // [pass * W + warp]: pairs
        std::vector<int> binload((size_t)npass * W, 0);
        bool ok = true;
        for (int pass = 0; pass < npass && ok; ++pass) {
            std::vector<int> mine;
            for (int q = 0; q < npairs; ++q)
                if (pass_of[(size_t)q] == pass && P->pairs[(size_t)q].nsteps > 0) mine.push_back(q);
            std::stable_sort(mine.begin(), mine.end(), [&](int a, int b) { return P->pairs[(size_t)a].nsteps > P->pairs[(size_t)b].nsteps; });
            for (int q : mine) {
                int best = -1;
                for (int w = 0; w < W; ++w) {
                    const size_t b = (size_t)pass * W + w;
                    if ((int)bins[b].size() >= NPWcap) continue;
                    if (best < 0 || binload[b] < binload[(size_t)pass * W + best]) best = w;
                }
                if (best < 0) {
                    ok = false;
                    break;
                }
                bins[(size_t)pass * W + best].push_back(q);
                binload[(size_t)pass * W + best] += P->pairs[(size_t)q].nsteps;
            }
        }
        if (!ok) continue;
        int NPW = 1;
        for (auto& b : bins) NPW = std::max(NPW, (int)b.size());
        NPW = NPW <= 2 ? 2 : NPW <= 4 ? 4 : 8;
        if (NPW > NPWcap && NPWcap >= 2) continue;

        // ---- T layout of a pass: step slots in (warp, pair, step) order, 128 doubles each (E block, O block) ---------------
        // pair word: pair index | steps << 12 | first step slot of the pair inside its pass << 16
        std::vector<uint32_t> pword((size_t)npass * W * NPW, 0u);
        std::vector<uint32_t> srec;  // 8 words per step: K row byte offsets of direction q, E | O << 16
        std::vector<uint32_t> passbase((size_t)npass, 0u);  // first global step of a pass
        struct Entry {
            uint32_t idx;
            double w;
        };
        std::vector<std::vector<std::vector<Entry>>> inbox((size_t)npass, std::vector<std::vector<Entry>>((size_t)ncols));
        uint32_t tmax = 0;
        bool fits = true;
        for (int pass = 0; pass < npass; ++pass) {
            uint32_t slot = 0;
            passbase[(size_t)pass] = (uint32_t)srec.size() / 8u;
            for (int w = 0; w < W; ++w) {
                const auto& bn = bins[(size_t)pass * W + w];
                for (size_t p = 0; p < bn.size(); ++p) {
                    const PairSteps& ps = P->pairs[(size_t)bn[p]];
                    if (bn[p] >= 4096 || ps.nsteps >= 16 || slot >= 65536u) fits = false;
                    pword[((size_t)pass * W + w) * NPW + p] = (uint32_t)bn[p] | (uint32_t)ps.nsteps << 12 | slot << 16;
                    for (int k = 0; k < ps.nsteps; ++k, ++slot) {
                        int dE = 0, dO = 0;
                        if (ps.blockE >= 0 && k < (int)P->block_dsets[(size_t)ps.blockE].size()) dE = P->block_dsets[(size_t)ps.blockE][(size_t)k];
                        if (ps.blockO >= 0 && k < (int)P->block_dsets[(size_t)ps.blockO].size()) dO = P->block_dsets[(size_t)ps.blockO][(size_t)k];
                        for (int q = 0; q < 8; ++q) {
                            const int de = dsets[(size_t)dE][(size_t)q], dv = dsets[(size_t)dO][(size_t)q];
                            srec.push_back((uint32_t)(de < 0 ? Mp : de) * krow_bytes | ((uint32_t)(dv < 0 ? Mp : dv) * krow_bytes) << 16);
                        }
                        // outputs: lane l = 4 q + kk holds T[direction q, modes 2 kk, 2 kk + 1] of both blocks
                        for (int lane = 0; lane < 32; ++lane) {
                            const int q = lane >> 2, kk = lane & 3;
                            for (int half = 0; half < 2; ++half) {
                                const int blk = half ? ps.blockO : ps.blockE;
                                const int ds = half ? dO : dE;
                                if (blk < 0 || ds == 0) continue;
                                const int dir = dsets[(size_t)ds][(size_t)q];
                                if (dir < 0) continue;
                                for (int j = 0; j < 2; ++j) {
                                    const int mode = P->blocks[(size_t)blk][(size_t)(2 * kk + j)];
                                    if (mode < 0) continue;
                                    const uint32_t idx = slot * 128u + (uint32_t)half * 64u + (uint32_t)lane * 2u + (uint32_t)j;
                                    for (auto& it : P->prod[(size_t)mode])
                                        if (it.dir == dir) inbox[(size_t)pass][(size_t)ctx->h_pos[(size_t)it.consumer]].push_back({idx, it.w});
                                }
                            }
                        }
                    }
                }
            }
            tmax = std::max(tmax, slot * 128u);
        }
        if (!fits) return 0;
        const uint32_t zero_idx = tmax;  // one entry per buffer that stays zero: padding of the lists
        const uint32_t tbuf = tmax + 2u;
        if (tbuf > 65535u) continue;

        // ---- lists: per (pass, group) rows of 32 indices, ordered by weight class; bank-aware order inside a class ---------
        struct GroupList {
            std::vector<uint32_t> idx;  // rows * 32
            std::vector<double> w;      // per row
        };
        std::vector<GroupList> glist((size_t)npass * ngroups);
        double rows_total = 0, rows_ideal = 0;
        long conflicts_before = 0, conflicts_after = 0;
        for (int pass = 0; pass < npass; ++pass)
            for (int g = 0; g < ngroups; ++g) {
                GroupList& L = glist[(size_t)pass * ngroups + g];
                std::vector<double> cls;
                std::vector<uint32_t> cnt;
                size_t items = 0;
                for (int c = 0; c < 32; ++c) {
                    std::vector<uint32_t> mine(cls.size(), 0u);
                    for (auto& o : inbox[(size_t)pass][(size_t)g * 32 + c]) {
                        size_t k = 0;
                        while (k < cls.size() && wbits(cls[k]) != wbits(o.w)) ++k;
                        if (k == cls.size()) cls.push_back(o.w), cnt.push_back(0u), mine.push_back(0u);
                        ++mine[k];
                        ++items;
                    }
                    for (size_t k = 0; k < mine.size(); ++k) cnt[k] = std::max(cnt[k], mine[k]);
                }
                rows_ideal += (double)(items + 31) / 32;
                std::vector<size_t> order(cls.size());
                std::iota(order.begin(), order.end(), 0);
                std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return cnt[a] > cnt[b]; });
                for (size_t oi = 0; oi < order.size(); ++oi) {
                    const size_t k = order[oi];
                    const uint32_t R = cnt[k];
                    // entries of the class per lane
                    std::vector<std::vector<uint32_t>> lane_items(32);
                    for (int c = 0; c < 32; ++c)
                        for (auto& o : inbox[(size_t)pass][(size_t)g * 32 + c])
                            if (wbits(o.w) == wbits(cls[k])) lane_items[(size_t)c].push_back(o.idx);
                    // rows: greedy, half-warp by half-warp (an 8-byte access is served per half-warp, bank = idx % 16): every
                    // lane picks an unused entry in a bank no earlier lane of its half-warp uses in this row; the lanes
                    // with the most entries left choose first
                    std::vector<uint32_t> rowsidx((size_t)R * 32, zero_idx);
                    auto row_cost = [&](const uint32_t* row) {
                        long cost = 0;
                        for (int h = 0; h < 2; ++h) {
                            int bank[16] = {0}, mx = 0;
                            for (int c = 0; c < 16; ++c) {
                                const uint32_t v = row[h * 16 + c];
                                if (v == zero_idx) continue;
                                mx = std::max(mx, ++bank[v & 15u]);
                            }
                            cost += std::max(mx, 1);
                        }
                        return cost;
                    };
                    {
                        std::vector<uint32_t> naive((size_t)R * 32, zero_idx);
                        for (int c = 0; c < 32; ++c)
                            for (size_t j = 0; j < lane_items[(size_t)c].size(); ++j) naive[j * 32 + (size_t)c] = lane_items[(size_t)c][j];
                        for (uint32_t r = 0; r < R; ++r) conflicts_before += row_cost(&naive[(size_t)r * 32]);
                    }
                    std::vector<std::vector<uint32_t>> left = lane_items;
                    for (uint32_t r = 0; r < R; ++r) {
                        for (int h = 0; h < 2; ++h) {
                            int bank[16] = {0};
                            std::vector<int> lanes;
                            for (int c = 0; c < 16; ++c)
                                if (!left[(size_t)(h * 16 + c)].empty()) lanes.push_back(h * 16 + c);
                            // a lane must place an entry in this row if it has as many entries left as rows left
                            std::stable_sort(lanes.begin(), lanes.end(), [&](int a2, int b2) { return left[(size_t)a2].size() > left[(size_t)b2].size(); });
                            for (int c : lanes) {
                                auto& lf = left[(size_t)c];
                                size_t best = 0;
                                int bc = 1 << 30;
                                for (size_t j = 0; j < lf.size(); ++j) {
                                    const int cst = bank[lf[j] & 15u];
                                    if (cst < bc) bc = cst, best = j;
                                }
                                const bool must = lf.size() >= (size_t)(R - r);
                                if (!must && bc > 0) continue;  // wait for a later row without a conflict
                                rowsidx[(size_t)r * 32 + (size_t)c] = lf[best];
                                ++bank[lf[best] & 15u];
                                lf.erase(lf.begin() + (long)best);
                            }
                        }
                        conflicts_after += row_cost(&rowsidx[(size_t)r * 32]);
                    }
                    for (uint32_t r = 0; r < R; ++r) {
                        L.w.push_back(cls[k]);
                        for (int c = 0; c < 32; ++c) L.idx.push_back(rowsidx[(size_t)r * 32 + (size_t)c]);
                    }
                }
                if (L.w.size() & 1) {
                    L.w.push_back(0.0);
                    for (int c = 0; c < 32; ++c) L.idx.push_back(zero_idx);
                }
                rows_total += (double)L.w.size();
            }

        // ---- consumer groups -> warps (longest list first) ------------------------------------------------------------------
        std::vector<int> gorder((size_t)ngroups), gload((size_t)ngroups, 0);
        std::iota(gorder.begin(), gorder.end(), 0);
        for (int g = 0; g < ngroups; ++g)
            for (int pass = 0; pass < npass; ++pass) gload[(size_t)g] += (int)glist[(size_t)pass * ngroups + g].w.size() + 2;
        std::stable_sort(gorder.begin(), gorder.end(), [&](int a2, int b2) { return gload[(size_t)a2] > gload[(size_t)b2]; });
        std::vector<std::vector<int>> wg((size_t)W);
        std::vector<int> wload((size_t)W, 0);
        for (int g : gorder) {
            int best = -1;
            for (int w = 0; w < W; ++w)
                if ((int)wg[(size_t)w].size() < NG && (best < 0 || wload[(size_t)w] < wload[(size_t)best])) best = w;
            wg[(size_t)best].push_back(g);
            wload[(size_t)best] += gload[(size_t)g];
        }

        // ---- blob -----------------------------------------------------------------------------------------------------------
        // consumer record: byte offset of the list (relative to the first list) | row pairs << 20; the weights of list row
        // pair j of the blob lie at off_roww + 16 * (global row pair)
        uint32_t at = 0;
        const uint32_t off_pair = at;
        at = align4(at + (uint32_t)npass * W * NPW);
        const uint32_t off_crec = at;
        at = align4(at + (uint32_t)npass * W * NG);
        const uint32_t off_gcol = at;
        at = align4(at + (uint32_t)W * NG);
        const uint32_t off_pbase = at;
        at = align4(at + (uint32_t)npass);
        const uint32_t off_srec = at;
        at = align4(at + (uint32_t)srec.size());
        const uint32_t off_list = at;
        std::vector<uint32_t> list_off((size_t)npass * ngroups);
        for (size_t k = 0; k < glist.size(); ++k) {
            list_off[k] = at;
            at += (uint32_t)glist[k].idx.size() / 2u;
        }
        at = align4(at);
        const uint32_t off_roww = at;
        at = align4(at + (at - off_list) / 8u);  // 2 doubles per 32-word list row pair
        const uint32_t nwords = at;
        if ((off_roww - off_list) * 4u >= (1u << 20)) continue;
        const uint32_t nwords_smem = lists_global ? off_list : nwords;  // words copied into shared memory
        // one pass with the lists in shared memory: a single T buffer (two barriers per dof row) if that is what fits
        const bool single = npass == 1 && !lists_global && try_single;
        const size_t smem = (size_t)nwords_smem * 4 + 8ull * (4 + 4 * KS) * 4ull + 3ull * (size_t)(Mp + 1) * krow_bytes +
                            (single ? 1ull : 2ull) * tbuf * 8ull + 16;
        if (verbose)
            fprintf(stderr,
                    "[blk] KS=%d W=%d passes=%d%s lists in %s NPW=%d NG=%d steps=%d T=%u doubles (x2) list rows %.0f (ideal %.0f) bank cost %ld -> %ld "
                    "tables=%u B smem=%zu B%s\n",
                    KS, W, npass, single ? " (single T buffer)" : "", lists_global ? "global memory" : "shared memory", NPW, NG, total_steps, tbuf, rows_total, rows_ideal,
                    conflicts_before, conflicts_after, nwords * 4, smem,
                    smem > (size_t)SMEM_LIMIT ? " (too large)" : "");
        if (smem > (size_t)SMEM_LIMIT) continue;

        std::vector<uint32_t> blob((size_t)nwords, 0u);
        std::memcpy(&blob[off_pair], pword.data(), pword.size() * 4);
        std::memcpy(&blob[off_pbase], passbase.data(), passbase.size() * 4);
        std::memcpy(&blob[off_srec], srec.data(), srec.size() * 4);
        for (int w = 0; w < W; ++w)
            for (int k = 0; k < NG; ++k) {
                const bool used = k < (int)wg[(size_t)w].size();
                const int g = used ? wg[(size_t)w][(size_t)k] : -1;
                blob[off_gcol + (size_t)w * NG + k] = used ? (uint32_t)g * 32u * 8u : 0xFFFFFFFFu;
                for (int pass = 0; pass < npass; ++pass) {
                    if (!used) continue;
                    const size_t gi = (size_t)pass * ngroups + g;
                    if (glist[gi].w.size() / 2u >= 4096u) return 0;
                    blob[off_crec + ((size_t)pass * W + w) * NG + k] = (list_off[gi] - off_list) * 4u | (uint32_t)(glist[gi].w.size() / 2u) << 20;
                }
            }
        for (size_t k = 0; k < glist.size(); ++k) {
            const GroupList& L = glist[k];
            for (size_t r = 0; r + 1 < L.w.size(); r += 2)
                for (int c = 0; c < 32; ++c) blob[list_off[k] + (r / 2) * 32 + (size_t)c] = L.idx[r * 32 + (size_t)c] | L.idx[(r + 1) * 32 + (size_t)c] << 16;
            if (!L.w.empty()) std::memcpy(&blob[off_roww + (list_off[k] - off_list) / 8u], L.w.data(), L.w.size() * 8);
        }

        BlkPlan* B = new BlkPlan();
        P->blk = B;
        B->KS = KS;
        B->P = npass;
        B->NPW = NPW;
        B->NG = NG;
        B->W = W;
        B->nwords = nwords_smem;
        B->lists_global = lists_global;
        B->single = single;
        B->off_pair = off_pair;
        B->off_crec = off_crec;
        B->off_gcol = off_gcol;
        B->off_pbase = off_pbase;
        B->off_srec = off_srec;
        B->off_list = off_list;
        B->off_roww = off_roww;
        B->tbuf_doubles = tbuf;
        B->smem_bytes = smem;
        B->dmma_per_row = 2.0 * KS * total_steps;
        B->list_rows = rows_total;
        B->list_ideal = rows_ideal;
        ASG_CUDA(ctx, cudaMalloc((void**)&B->d_blob, blob.size() * 4));
        ASG_CUDA(ctx, cudaMemcpyAsync(B->d_blob, blob.data(), blob.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        {
            // row records: what the kernel needs of the CSR structure of a row, in one contiguous piece
            const int ME = 4 + 4 * KS;
            std::vector<int32_t> meta((size_t)nrows * ME, 0);
            for (int64_t i = 0; i < nrows; ++i) {
                int32_t* m = &meta[(size_t)i * ME];
                const int64_t p0 = ctx->h_rowptr[i];
                std::memcpy(m, &p0, 8);
                m[2] = (int32_t)(ctx->h_rowptr[i + 1] - p0);
                m[3] = ctx->h_bmask.empty() ? 0 : ctx->h_bmask[(size_t)i];
                for (int k = 0; k < m[2]; ++k) m[4 + k] = ctx->h_col[(size_t)(p0 + k)];
            }
            ASG_CUDA(ctx, cudaMalloc((void**)&B->d_rowmeta, meta.size() * 4));
            ASG_CUDA(ctx, cudaMemcpyAsync(B->d_rowmeta, meta.data(), meta.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
            ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
        ASG_CUDA(ctx, cudaMalloc((void**)&B->d_zero, sizeof(double) * (size_t)ctx->ld));
        ASG_CUDA(ctx, cudaMemsetAsync(B->d_zero, 0, sizeof(double) * (size_t)ctx->ld, ctx->stream));
        ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        B->grid = 148;
        if (const char* e = getenv("ASGFEM_BLK_GRID")) {
            const int v = atoi(e);
            if (v >= 1 && v <= 1024) B->grid = v;
        }
        B->usable = true;
        return 0;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------------------------------
namespace {

struct BlkArgs {
    const double* x;
    double* y;
    const double* vals;
    const uint8_t* bmask;
    const uint32_t* blob;
    const double* zero_row;
    const int32_t* rowmeta;
    int64_t nnz, ld, r0, r1;
    int Mp, P, single;
    uint32_t nwords, off_pair, off_crec, off_gcol, off_pbase, off_srec, off_list, off_roww, tbuf_doubles;
};

__device__ __forceinline__ void b_cp_async8(unsigned s, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void b_cp_async16(unsigned s, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void b_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void b_cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void b_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void b_dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// shared memory through 32-bit addresses (generic pointers cost 64-bit address arithmetic per access)
__device__ __forceinline__ unsigned b_lds_u32(unsigned addr) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ int b_lds_s32(unsigned addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 b_lds_u32x2(unsigned addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 b_lds_u32x4(unsigned addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ double b_lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ double2 b_lds_f64x2(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void b_sts_f64x2(unsigned addr, double a, double b) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void b_sts_f64(unsigned addr, double a) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(a) : "memory"); }
__device__ __forceinline__ double2 b_ldg_f64x2(const char* p) {
    double2 v;
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

__device__ __forceinline__ unsigned b_ldg_u32(const unsigned char* p) {
    unsigned v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// GL: the lists (indices and row weights) stay in global memory (read through L1 / L2) instead of shared memory
template <int KS, int NPW, int NG, bool GL>
__global__ void __launch_bounds__(512, 1) k_apply_blk(const BlkArgs a) {
    constexpr int WARPS = 16;
    extern __shared__ __align__(16) unsigned char sm[];
    constexpr int THREADS = WARPS * 32;
    constexpr int ME = 4 + 4 * KS;  // ints per row record
    constexpr int RING = 8;         // row records in flight
    constexpr uint32_t KROW = 32u * KS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = lane >> 2, kk = lane & 3;
    const int LA = a.P >= 2 ? 1 : 2;  // rows of lookahead: data issued in stage t is complete at the end of stage t + 1
    const int KB = LA + 1;            // K buffers in use
    const uint32_t kbuf_bytes = (uint32_t)(a.Mp + 1) * KROW;  // directions 0..M and the null row
    const unsigned sm0 = (unsigned)__cvta_generic_to_shared(sm);
    const unsigned meta_s = sm0 + a.nwords * 4u, ks_s = meta_s + RING * ME * 4u, tb_s = ks_s + 3u * kbuf_bytes;
    const uint32_t tb_bytes = a.tbuf_doubles * 8u;

    const int64_t rstep = gridDim.x;
    const int64_t rb = a.r0 + (int64_t)blockIdx.x;
    if (rb >= a.r1) return;
    const int nri = (int)((a.r1 - rb + rstep - 1) / rstep);  // rows of this CTA

    {
        uint32_t* blob = reinterpret_cast<uint32_t*>(sm);
        for (uint32_t i = tid; i < a.nwords; i += THREADS) blob[i] = a.blob[i];
        double* z = reinterpret_cast<double*>(sm + a.nwords * 4u);
        for (uint32_t i = tid; i < (RING * ME * 4u + 3u * kbuf_bytes + (a.single ? 1u : 2u) * tb_bytes) / 8u; i += THREADS) z[i] = 0.0;
    }
    __syncthreads();

    const unsigned pair_s = sm0 + a.off_pair * 4u + warp * (NPW * 4);  // + pass * WARPS * NPW * 4
    const unsigned crec_s = sm0 + a.off_crec * 4u + warp * (NG * 4);   // + pass * WARPS * NG * 4
    const unsigned srec_s = sm0 + a.off_srec * 4u + q * 4;             // + global step * 32
    const unsigned list_s = sm0 + a.off_list * 4u + lane * 4, roww_s = sm0 + a.off_roww * 4u;
    const unsigned char* list_g = reinterpret_cast<const unsigned char*>(a.blob + a.off_list) + lane * 4;
    const unsigned char* roww_g = reinterpret_cast<const unsigned char*>(a.blob + a.off_roww);
    auto ld_idx = [&](uint32_t off) { return GL ? b_ldg_u32(list_g + off) : b_lds_u32(list_s + off); };
    auto ld_w = [&](uint32_t off) { return GL ? b_ldg_f64x2(reinterpret_cast<const char*>(roww_g + off)) : b_lds_f64x2(roww_s + off); };

    uint32_t ycol[NG];  // byte offset of this thread's column of group g inside a row of Y (0xFFFFFFFF: unused slot)
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        const uint32_t c = b_lds_u32(sm0 + a.off_gcol * 4u + (warp * NG + g) * 4);
        ycol[g] = c == 0xFFFFFFFFu ? c : c + (uint32_t)lane * 8u;
    }

    auto meta_fetch = [&](int ri) {
        if (tid < ME / 4) b_cp_async16(meta_s + (unsigned)((ri & (RING - 1)) * ME + tid * 4) * 4u, a.rowmeta + (rb + ri * rstep) * ME + tid * 4);
    };
    // K rows of row ri -> Ks[ri % KB][m][k & 3][k >> 2] (k >= length of the row: 0); one element per thread and trip
    const int sk_m = tid / (4 * KS), sk_k = tid - sk_m * (4 * KS);
    const unsigned sk_dst = (unsigned)(sk_m * (4 * KS) + (sk_k & 3) * KS + (sk_k >> 2)) * 8u;
    auto stage_k = [&](int ri, int kslot) {  // kslot = ri % KB
        const unsigned m = meta_s + (unsigned)((ri & (RING - 1)) * ME) * 4u;
        const uint2 pp = b_lds_u32x2(m);
        const int64_t p0 = (int64_t)((uint64_t)pp.x | (uint64_t)pp.y << 32);
        const int len = b_lds_s32(m + 8);
        unsigned dst = ks_s + (unsigned)kslot * kbuf_bytes + sk_dst;
        const double* src = a.vals + (int64_t)sk_m * a.nnz + p0 + sk_k;
        for (int mm = sk_m; mm < a.Mp; mm += THREADS / (4 * KS)) {
            if (sk_k < len)
                b_cp_async8(dst, src);
            else
                b_sts_f64(dst, 0.0);
            dst += (THREADS / (4 * KS)) * KROW;
            src += (int64_t)(THREADS / (4 * KS)) * a.nnz;
        }
    };
    // X rows this lane LOADS for a dof row: lane l reads the 16 bytes (l & 7) of the row of slot 4 s + (l >> 3), so that
    // eight consecutive lanes read one 128-byte line; the fragments are permuted with shuffles when they are used
    // (lane 4 q + kk <- lane 8 kk + q).  Slots beyond the row read a row of zeros.
    const int lrow = lane >> 3, lchunk = lane & 7;
    const int frag_src = (kk << 3) | q;
    auto row_ptrs = [&](int ri, const char* (&xr)[KS]) {
        const unsigned m = meta_s + (unsigned)((ri & (RING - 1)) * ME) * 4u;
        const int len = b_lds_s32(m + 8);
#pragma unroll
        for (int s = 0; s < KS; ++s) {
            const int slot = 4 * s + lrow;
            const int col = b_lds_s32(m + 16 + slot * 4);
            xr[s] = reinterpret_cast<const char*>((slot < len ? a.x + (int64_t)col * a.ld : a.zero_row) + 2 * lchunk);
        }
    };
    auto load_x = [&](int pass, const char* const (&xr)[KS], double2 (&X)[NPW][KS]) {
        uint32_t pw[NPW];
        if constexpr (NPW == 2) {
            const uint2 v = b_lds_u32x2(pair_s + pass * (WARPS * NPW * 4));
            pw[0] = v.x, pw[1] = v.y;
        } else {
#pragma unroll
            for (int p4 = 0; p4 < NPW / 4; ++p4) {
                const uint4 v = b_lds_u32x4(pair_s + pass * (WARPS * NPW * 4) + p4 * 16);
                pw[4 * p4] = v.x, pw[4 * p4 + 1] = v.y, pw[4 * p4 + 2] = v.z, pw[4 * p4 + 3] = v.w;
            }
        }
#pragma unroll
        for (int p = 0; p < NPW; ++p) {
            if (pw[p] & 0xF000u) {
                const uint32_t cb = (pw[p] & 0xFFFu) * 128u;
#pragma unroll
                for (int s = 0; s < KS; ++s) X[p][s] = b_ldg_f64x2(xr[s] + cb);
            }
        }
    };

    double2 X[NPW][KS];
    double acc[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) acc[g] = 0.0;
#pragma unroll
    for (int p = 0; p < NPW; ++p)
#pragma unroll
        for (int s = 0; s < KS; ++s) X[p][s] = make_double2(0.0, 0.0);

    // prologue: row records of the first 2 LA rows, K rows of the first LA rows, X fragments of the first stage
    for (int j = 0; j < 2 * LA && j < nri; ++j) meta_fetch(j);
    b_cp_async_wait_all();
    __syncthreads();
    for (int j = 0; j < LA && j < nri; ++j) stage_k(j, j);
    const char* xr[KS];  // X rows of the dof row of the NEXT stage
    row_ptrs(0, xr);
    load_x(0, xr, X);
    b_cp_async_wait_all();
    __syncthreads();

    int kcur = 0, knew = LA;  // K buffers of row ri and of row ri + LA (LA < KB)
    // products of (row ri, pass) into T buffer par; X fragments of the next stage are loaded at the end
    auto do_produce = [&](int ri, int pass, uint32_t par) {
        if (pass == 0) {
            if (ri + 2 * LA < nri) meta_fetch(ri + 2 * LA);
            if (ri + LA < nri) stage_k(ri + LA, knew);
        }
        int npass = pass + 1;
        int nxt = ri;
        if (npass == a.P) npass = 0, ++nxt;
        const unsigned tb = tb_s + par * tb_bytes + lane * 16;
        const unsigned ksrc = ks_s + (unsigned)kcur * kbuf_bytes + kk * (8 * KS);
        const unsigned sbase = srec_s + b_lds_u32(sm0 + a.off_pbase * 4u + pass * 4) * 32u;
        uint32_t pw[NPW];
        if constexpr (NPW == 2) {
            const uint2 v = b_lds_u32x2(pair_s + pass * (WARPS * NPW * 4));
            pw[0] = v.x, pw[1] = v.y;
        } else {
#pragma unroll
            for (int p4 = 0; p4 < NPW / 4; ++p4) {
                const uint4 v = b_lds_u32x4(pair_s + pass * (WARPS * NPW * 4) + p4 * 16);
                pw[4 * p4] = v.x, pw[4 * p4 + 1] = v.y, pw[4 * p4 + 2] = v.z, pw[4 * p4 + 3] = v.w;
            }
        }
#pragma unroll
        for (int p = 0; p < NPW; ++p) {
            const uint32_t nst = (pw[p] >> 12) & 0xFu;
            if (nst != 0) {  // warp-uniform
                double xe[KS], xo[KS];
#pragma unroll
                for (int s = 0; s < KS; ++s) {
                    xe[s] = __shfl_sync(0xffffffffu, X[p][s].x, frag_src);
                    xo[s] = __shfl_sync(0xffffffffu, X[p][s].y, frag_src);
                }
                const uint32_t slot = pw[p] >> 16;
                unsigned sr = sbase + slot * 32u;
                unsigned td = tb + slot * 1024u;
                for (uint32_t st = 0; st < nst; ++st, sr += 32u, td += 1024u) {
                    const uint32_t w = b_lds_u32(sr);
                    const unsigned ke = ksrc + (w & 0xFFFFu), ko = ksrc + (w >> 16);
                    double ae[KS], ao[KS];
#pragma unroll
                    for (int s = 0; s < KS; s += 2) {
                        const double2 ve = b_lds_f64x2(ke + s * 8), vo = b_lds_f64x2(ko + s * 8);
                        ae[s] = ve.x, ae[s + 1] = ve.y;
                        ao[s] = vo.x, ao[s + 1] = vo.y;
                    }
                    double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
#pragma unroll
                    for (int s = 0; s < KS; ++s) {
                        b_dmma(c0, c1, ae[s], xe[s]);
                        b_dmma(c2, c3, ao[s], xo[s]);
                    }
                    b_sts_f64x2(td, c0, c1);
                    b_sts_f64x2(td + 512u, c2, c3);
                }
            }
        }
        // X fragments of the next stage: one batch of loads, in flight during the list sums / the barrier
        if (nxt < nri) {
            if (npass == 0) row_ptrs(nxt, xr);
            load_x(npass, xr, X);
        }
    };
    // weighted list sums of (row cri, pass cpass) from T buffer par into the accumulators; the last pass writes the Y row
    auto do_consume = [&](int cri, int cpass, uint32_t par) {
        // ---- consume the previous stage from the other buffer: weighted list sums of the groups of this warp ---------
        const int64_t crow = rb + (int64_t)cri * rstep;
        const bool last = cpass == a.P - 1;
        uint8_t bm = 0;
        if (last) bm = a.bmask[crow];  // in flight during the sums
        const unsigned tbr = tb_s + (par ^ 1u) * tb_bytes;
        uint32_t cw[NG];
        if constexpr (NG == 2) {
            const uint2 v = b_lds_u32x2(crec_s + cpass * (WARPS * NG * 4));
            cw[0] = v.x, cw[1] = v.y;
        } else {
#pragma unroll
            for (int g4 = 0; g4 < NG / 4; ++g4) {
                const uint4 v = b_lds_u32x4(crec_s + cpass * (WARPS * NG * 4) + g4 * 16);
                cw[4 * g4] = v.x, cw[4 * g4 + 1] = v.y, cw[4 * g4 + 2] = v.z, cw[4 * g4 + 3] = v.w;
            }
        }
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            uint32_t lo = cw[g] & 0xFFFFFu, wo = lo >> 3;
            double t0 = acc[g], t1 = 0.0;
            uint32_t r = cw[g] >> 20;
            while (r >= 2) {  // 4 list rows per trip: independent loads first
                const uint32_t i0 = ld_idx(lo), i1 = ld_idx(lo + 128);
                const double2 w0 = ld_w(wo), w1 = ld_w(wo + 16);
                const double v0 = b_lds_f64(tbr + (i0 & 0xFFFFu) * 8u), v1 = b_lds_f64(tbr + (i0 >> 16) * 8u);
                const double v2 = b_lds_f64(tbr + (i1 & 0xFFFFu) * 8u), v3 = b_lds_f64(tbr + (i1 >> 16) * 8u);
                t0 = fma(w0.x, v0, t0);
                t1 = fma(w0.y, v1, t1);
                t0 = fma(w1.x, v2, t0);
                t1 = fma(w1.y, v3, t1);
                lo += 256, wo += 32, r -= 2;
            }
            if (r) {
                const uint32_t i0 = ld_idx(lo);
                const double2 w0 = ld_w(wo);
                const double v0 = b_lds_f64(tbr + (i0 & 0xFFFFu) * 8u), v1 = b_lds_f64(tbr + (i0 >> 16) * 8u);
                t0 = fma(w0.x, v0, t0);
                t1 = fma(w0.y, v1, t1);
            }
            acc[g] = t0 + t1;
        }
        if (last) {
            unsigned char* yr = reinterpret_cast<unsigned char*>(a.y + crow * a.ld);
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                if (ycol[g] != 0xFFFFFFFFu) *reinterpret_cast<double*>(yr + ycol[g]) = bm ? 0.0 : acc[g];
                acc[g] = 0.0;
            }
        }
    };

    if (a.single) {
        // one pass, ONE T buffer: produce | barrier | sum the lists | barrier.  Fewer list rows (no split of a column's list
        // over passes) and half the per-stage overhead; the two phases no longer overlap.
        for (int ri = 0; ri < nri; ++ri) {
            do_produce(ri, 0, 0u);
            __syncthreads();
            do_consume(ri, 0, 1u);  // reads buffer (1 ^ 1) = 0
            b_cp_async_commit();
            b_cp_async_wait_1();
            __syncthreads();
            kcur = kcur + 1 == KB ? 0 : kcur + 1;
            knew = knew + 1 == KB ? 0 : knew + 1;
        }
        return;
    }
    int pass = 0;
    int ri = 0;
    uint32_t par = 0;  // parity of the stage = T buffer
    bool have_prev = false;
    int cpass = 0;
    int cri = 0;
    while (ri < nri || have_prev) {
        const bool produce = ri < nri;
        // odd warps consume first: the shared-memory reads of one half of the warps overlap the fp64 products of the other
        for (int phase = 0; phase < 2; ++phase) {
            if (phase == (warp & 1)) {
                if (produce) do_produce(ri, pass, par);
            } else if (have_prev) {
                do_consume(cri, cpass, par);
            }
        }
        b_cp_async_commit();
        b_cp_async_wait_1();  // everything but the copies issued in this stage has landed
        __syncthreads();
        have_prev = produce;
        cpass = pass;
        cri = ri;
        par ^= 1u;
        if (produce && ++pass == a.P) {
            pass = 0, ++ri;
            kcur = kcur + 1 == KB ? 0 : kcur + 1;
            knew = knew + 1 == KB ? 0 : knew + 1;
        }
    }
}

template <int KS, int NPW, int NG, bool GL>
int launch_blk(asgfem_ctx* ctx, BlkPlan* B, const BlkArgs& a) {
    static bool configured = false;
    if (!configured) {
        ASG_CUDA(ctx, cudaFuncSetAttribute(k_apply_blk<KS, NPW, NG, GL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        configured = true;
    }
    const int64_t nr = a.r1 - a.r0;
    const int grid = (int)std::min<int64_t>(B->grid, nr);
    k_apply_blk<KS, NPW, NG, GL><<<grid, 512, B->smem_bytes, ctx->stream>>>(a);
    ASG_CUDA(ctx, cudaGetLastError());
    return 0;
}

template <int KS, int NPW, bool GL>
int launch_blk_ng(asgfem_ctx* ctx, BlkPlan* B, const BlkArgs& a) {
    switch (B->NG) {
        case 2: return launch_blk<KS, NPW, 2, GL>(ctx, B, a);
        case 4: return launch_blk<KS, NPW, 4, GL>(ctx, B, a);
        case 8: return launch_blk<KS, NPW, 8, GL>(ctx, B, a);
        default: return launch_blk<KS, NPW, 16, GL>(ctx, B, a);
    }
}

template <int KS, bool GL>
int launch_blk_np(asgfem_ctx* ctx, BlkPlan* B, const BlkArgs& a) {
    constexpr int CAP = 16 / KS;
    if (B->NPW <= 2) return launch_blk_ng<KS, 2, GL>(ctx, B, a);
    if constexpr (CAP >= 4) {
        if (B->NPW <= 4) return launch_blk_ng<KS, 4, GL>(ctx, B, a);
    }
    if constexpr (CAP >= 8) {
        if (B->NPW <= 8) return launch_blk_ng<KS, 8, GL>(ctx, B, a);
    }
    return fail(ctx, ASGFEM_ESTATE, "block operator: no kernel instance for this shape");
}

}  // namespace

int apply_blk_launch(asgfem_ctx* ctx, const double* x, double* y, int64_t r0, int64_t r1) {
    BlkPlan* B = bp_of(ctx);
    if (!B || !B->usable) return fail(ctx, ASGFEM_ESTATE, "block operator plan not available for this pattern / multi-index set");
    BlkArgs a;
    a.x = x;
    a.y = y;
    a.vals = ctx->d_vals;
    a.bmask = ctx->d_bmask;
    a.blob = B->d_blob;
    a.zero_row = B->d_zero;
    a.rowmeta = B->d_rowmeta;
    a.nnz = ctx->nnz;
    a.ld = ctx->ld;
    a.r0 = r0;
    a.r1 = r1;
    a.Mp = ctx->M + 1;
    a.P = B->P;
    a.single = B->single ? 1 : 0;
    a.nwords = B->nwords;
    a.off_pair = B->off_pair;
    a.off_crec = B->off_crec;
    a.off_gcol = B->off_gcol;
    a.off_pbase = B->off_pbase;
    a.off_srec = B->off_srec;
    a.off_list = B->off_list;
    a.off_roww = B->off_roww;
    a.tbuf_doubles = B->tbuf_doubles;
    switch (B->KS * 2 + (B->lists_global ? 1 : 0)) {
        case 4: return launch_blk_np<2, false>(ctx, B, a);
        case 5: return launch_blk_np<2, true>(ctx, B, a);
        case 8: return launch_blk_np<4, false>(ctx, B, a);
        case 9: return launch_blk_np<4, true>(ctx, B, a);
        case 12: return launch_blk_np<6, false>(ctx, B, a);
        case 13: return launch_blk_np<6, true>(ctx, B, a);
        default: return fail(ctx, ASGFEM_ESTATE, "block operator: no kernel instance for this row length");
    }
}

}  // namespace asgfem
