// Fused SGFE operator, default kernel: block products on the fp64 MMA path (DMMA m8n8k4), mailbox exchange.
//
//   Y[i, mu] = sum_k K_0[i,j_k] X[j_k,mu] + sum_{(m,nu) ~ mu} g sum_k K_m[i,j_k] X[j_k,nu]      (mul!, :101-117)
//
// For one dof row i the partial products T[m, nu] = sum_k K_m[i, j_k] X[j_k, nu] form the matrix product
// (K-rows of the row: (M+1) x len) x (X rows of the neighbours: len x N), of which only the pairs (m, nu) with a coupling
// are needed (30 % on the benchmark set).  Measured on B200 (tools/ubench_fp64.cu, profiles/r02_ubench_fp64.txt): DFMA and
// DMMA share one pipe (58 vs 64 FMA/clk/SM, no gain when mixed), so the MMA form gives no extra flops - what it gives is
// operand delivery: one DMMA = 256 FMAs for one issue slot with 1 + 1 + 2 register operands per lane, where the DFMA
// formulations of round 1 (apply_ts2.cu, retired) needed one shared-memory operand per FMA instruction and ran at 11 %
// of the fp64 pipe.  The price is evaluating 8 x 8 blocks of (direction, mode) pairs of which 50-87 % are needed.
//
// Host plan (mode side, at set_multiindices; independent of the mesh):
//   * every needed product T[m, nu] is an ITEM (producer mode nu, direction m) with one consumer mode mu and weight g (the
//     mean term is the item (mu, 0) with consumer mu and weight 1); the few products with two consumers (nu - e_m and
//     nu + e_m both in the set) keep one primary consumer, the other one fetches the value through a short extra list;
//   * modes are clustered greedily into HOME BLOCKS of 8 modes whose key sets overlap, so that the union of the keys of a
//     block needs few D-SETS of 8 keys (benchmark set: 250 blocks, 319 D-set uses = 638 DMMAs per row; lower bound 630);
//   * two blocks with equally many D-sets form a PAIR = 16 consecutive columns of the device layout (even columns = first
//     block): one 16-byte load per lane fetches the B fragments of both.  The device column order of ALL vectors is this
//     order (ctx->h_pos; the layout is private, converted at upload / download), consumers are grouped by 32 columns;
//   * items are delivered through a MAILBOX in shared memory: consumer group g owns S rows of 32 doubles per pass, the
//     producer lane stores its output to (row, consumer lane) - one 16-bit address per output in the step's store
//     words - and the consumer lane just sums its column of the mailbox.  The rows of a group are ordered in RUNS of equal
//     weight (the planner sorts the items of a consumer by weight; g takes few values), so the consumer multiplies once per
//     run with a warp-uniform weight: no index lists, no per-item weights, fixed summation order.
// Kernel: persistent CTAs (one per SM, 16 warps) walk contiguous row ranges.  The row is processed in P passes over the
// pairs (P chosen so that two mailbox buffers fit); stage t = (row, pass): every warp PRODUCES its pairs of stage t into
// mailbox t&1 (B fragments in registers, loaded one stage ahead; A fragments from the staged K rows of the row) and then
// CONSUMES stage t-1 from the other mailbox into its Y accumulators - one block barrier per stage, and the fp64 pipe of
// the producers overlaps the shared-memory reads of the consumers of other warps.
#include <algorithm>
#include <cstring>
#include <map>
#include <numeric>

#include "common.h"

namespace asgfem {

namespace {
constexpr int MMA_WARPS = 16;
constexpr int MMA_THREADS = MMA_WARPS * 32;
constexpr uint32_t NOSTORE = 0xFFFFu;
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int MMA_MAXP = 3, MMA_MAXROWW = 1280, MMA_MAXW = 64;  // capacity of the kernel parameter block

struct PairSteps {
    int blockE = -1, blockO = -1;
    int nsteps = 0;
};
}  // namespace

struct MmaPlan {
    bool layout_ok = false;  // mode-side clustering done (set_multiindices)
    bool usable = false;     // kernel tables built (first apply)
    int64_t N = 0;
    int M = 0;
    int ncols = 0;
    // mode side
    struct Item {
        int dir, consumer;
        double w;
        bool primary;  // the first consumer of (producer, direction) owns the mailbox slot
    };
    std::vector<std::vector<Item>> prod;  // per producer mode
    std::vector<std::array<int, 8>> blocks;              // modes of a home block (-1 = empty)
    std::vector<std::vector<int>> block_dsets;           // D-set ids per block
    std::vector<std::array<int, 8>> dsets;               // directions (-1 = null row)
    std::vector<PairSteps> pairs;
    // kernel side
    int KS = 0, P = 1, NS = 0, NG = 0, nsteps = 0;
    uint32_t off_step = 0, off_sw = 0, off_dtab = 0, off_crec = 0, off_tailw = 0, off_extra = 0, off_wtab = 0, off_gcol = 0, nwords = 0;
    uint32_t mb_doubles = 0;
    size_t smem_bytes = 0;
    uint32_t* d_blob = nullptr;
    double* d_zero = nullptr;  // one row of zeros: X row of the unused neighbour slots
    int32_t* d_rowmeta = nullptr;  // per row 4 + 4 KS ints: first CSR position (int64), length, Dirichlet flag, columns (ELL)
    std::vector<uint32_t> h_step, h_crec, h_gcol;  // warp-uniform tables: passed as kernel parameters (constant bank)
    std::vector<double> h_roww, h_wtab;
    int grid = 148;
    double dmma_per_row = 0;
};

static MmaPlan* mp_of(asgfem_ctx* ctx) { return reinterpret_cast<MmaPlan*>(ctx->mmaplan); }

void apply_mma_free(asgfem_ctx* ctx) {
    MmaPlan* P = mp_of(ctx);
    if (!P) return;
    if (P->d_blob) cudaFree(P->d_blob);
    if (P->d_zero) cudaFree(P->d_zero);
    if (P->d_rowmeta) cudaFree(P->d_rowmeta);
    delete P;
    ctx->mmaplan = nullptr;
}

// ---------------------------------------------------------------------------------------------------------------------
// mode-side plan: items, keys, home blocks, pairs, device column order
// ---------------------------------------------------------------------------------------------------------------------
namespace {
struct Mask {
    std::vector<uint64_t> w;
    explicit Mask(int words = 0) : w((size_t)words, 0ull) {}
    void set(int b) { w[(size_t)b >> 6] |= 1ull << (b & 63); }
    int count() const {
        int c = 0;
        for (uint64_t x : w) c += __builtin_popcountll(x);
        return c;
    }
};
inline int union_count(const Mask& a, const Mask& b) {
    int c = 0;
    for (size_t k = 0; k < a.w.size(); ++k) c += __builtin_popcountll(a.w[k] | b.w[k]);
    return c;
}
inline int inter_count(const Mask& a, const Mask& b) {
    int c = 0;
    for (size_t k = 0; k < a.w.size(); ++k) c += __builtin_popcountll(a.w[k] & b.w[k]);
    return c;
}
}  // namespace

// Builds the mode-side plan and the device column order.  Called from asgfem_set_multiindices; on failure (too many keys)
// the layout falls back to the identity and only the gather kernel is available.
int apply_mma_layout(asgfem_ctx* ctx) {
    apply_mma_free(ctx);
    MmaPlan* P = new MmaPlan();
    ctx->mmaplan = P;
    const int64_t N = ctx->N;
    const Coupling& C = ctx->coup;
    P->N = N;
    int M = 0;
    for (int32_t m : C.m) M = std::max(M, (int)m);
    P->M = M;

    // ---- items: producer nu, direction m -> consumer mu with weight g ------------------------------------------------
    P->prod.assign((size_t)N, {});
    for (int64_t mu = 0; mu < N; ++mu) P->prod[(size_t)mu].push_back({0, (int)mu, 1.0, true});
    for (int64_t mu = 0; mu < N; ++mu)
        for (int32_t e = C.ptr[mu]; e < C.ptr[mu + 1]; ++e) {
            const int nu = C.nu[e];
            bool primary = true;
            for (auto& it : P->prod[(size_t)nu])
                if (it.dir == C.m[e]) primary = false;
            P->prod[(size_t)nu].push_back({(int)C.m[e], (int)mu, C.g[e], primary});
        }
    const int nkeys = M + 1;
    const int words = (nkeys + 63) / 64;
    std::vector<Mask> need((size_t)N, Mask(words));
    for (int64_t nu = 0; nu < N; ++nu)
        for (auto& it : P->prod[(size_t)nu]) need[(size_t)nu].set(it.dir);

    // ---- greedy clustering into home blocks of 8 modes --------------------------------------------------------------
    std::vector<int> cnt((size_t)N);
    for (int64_t nu = 0; nu < N; ++nu) cnt[(size_t)nu] = need[(size_t)nu].count();
    std::vector<int> order((size_t)N);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cnt[(size_t)a] > cnt[(size_t)b]; });
    std::vector<char> used((size_t)N, 0);
    std::vector<int> unassigned((size_t)N);  // sorted by index, compacted lazily
    std::iota(unassigned.begin(), unassigned.end(), 0);
    const int WINDOW = 4096;
    std::vector<Mask> block_union;
    for (int seed : order) {
        if (used[(size_t)seed]) continue;
        used[(size_t)seed] = 1;
        std::array<int, 8> blk;
        blk.fill(-1);
        blk[0] = seed;
        Mask u = need[(size_t)seed];
        int ucnt = cnt[(size_t)seed];
        // candidate window around the seed in the list of unassigned modes
        unassigned.erase(std::remove_if(unassigned.begin(), unassigned.end(), [&](int j) { return used[(size_t)j] != 0; }),
                         unassigned.end());
        size_t lo = 0, hi = unassigned.size();
        if (unassigned.size() > (size_t)WINDOW) {
            size_t at = std::lower_bound(unassigned.begin(), unassigned.end(), seed) - unassigned.begin();
            lo = at > (size_t)WINDOW / 2 ? at - WINDOW / 2 : 0;
            hi = std::min(unassigned.size(), lo + WINDOW);
        }
        for (int slot = 1; slot < 8; ++slot) {
            int best = -1;
            long bc0 = 0, bc1 = 0, bc2 = 0, bc3 = 0;
            for (size_t q = lo; q < hi; ++q) {
                const int j = unassigned[q];
                if (used[(size_t)j]) continue;
                const int uc = union_count(u, need[(size_t)j]);
                const long c0 = (uc + 7) / 8, c1 = uc, c2 = -inter_count(u, need[(size_t)j]), c3 = std::abs(j - seed);
                if (best < 0 || std::tie(c0, c1, c2, c3) < std::tie(bc0, bc1, bc2, bc3)) {
                    best = j;
                    bc0 = c0, bc1 = c1, bc2 = c2, bc3 = c3;
                }
            }
            if (best < 0) break;
            used[(size_t)best] = 1;
            blk[(size_t)slot] = best;
            for (int k = 0; k < words; ++k) u.w[(size_t)k] |= need[(size_t)best].w[(size_t)k];
            ucnt = (int)bc1;
        }
        P->blocks.push_back(blk);
        block_union.push_back(u);
    }

    // ---- D-sets per block (keys ascending, chunks of 8), global table of distinct D-sets ----------------------------
    std::map<std::array<int, 8>, int> dset_id;
    auto get_dset = [&](const std::array<int, 8>& d) {
        auto it = dset_id.find(d);
        if (it != dset_id.end()) return it->second;
        int id = (int)P->dsets.size();
        dset_id[d] = id;
        P->dsets.push_back(d);
        return id;
    };
    {
        std::array<int, 8> nul;
        nul.fill(-1);
        get_dset(nul);  // D-set 0 = all null rows
    }
    P->block_dsets.assign(P->blocks.size(), {});
    for (size_t b = 0; b < P->blocks.size(); ++b) {
        std::vector<int> keys;
        for (int k = 0; k < nkeys; ++k)
            if (block_union[b].w[(size_t)k >> 6] >> (k & 63) & 1ull) keys.push_back(k);
        for (size_t c = 0; c < keys.size(); c += 8) {
            std::array<int, 8> d;
            d.fill(-1);
            for (size_t r = 0; r < 8 && c + r < keys.size(); ++r) d[r] = keys[c + r];
            P->block_dsets[b].push_back(get_dset(d));
        }
    }

    // ---- pairs: blocks sorted by (number of D-sets descending, D-set ids), consecutive blocks paired -----------------
    std::vector<int> border(P->blocks.size());
    std::iota(border.begin(), border.end(), 0);
    std::stable_sort(border.begin(), border.end(), [&](int a, int b) {
        const auto &da = P->block_dsets[(size_t)a], &db = P->block_dsets[(size_t)b];
        if (da.size() != db.size()) return da.size() > db.size();
        return da < db;
    });
    for (size_t q = 0; q < border.size(); q += 2) {
        PairSteps ps;
        ps.blockE = border[q];
        ps.blockO = q + 1 < border.size() ? border[q + 1] : -1;
        ps.nsteps = (int)P->block_dsets[(size_t)ps.blockE].size();
        if (ps.blockO >= 0) ps.nsteps = std::max(ps.nsteps, (int)P->block_dsets[(size_t)ps.blockO].size());
        P->pairs.push_back(ps);
    }
    // 32 columns per consumer group: an even number of pairs
    if (P->pairs.size() % 2) P->pairs.push_back(PairSteps());
    P->ncols = (int)P->pairs.size() * 16;

    // ---- bank-aware order of the modes inside their blocks ---------------------------------------------------------------
    // The mailbox entry of consumer column c lies in 8-byte bank c % 16 = 2 * (slot in its block) + (odd block of its pair),
    // whatever group / row it is in.  A store instruction = (block, D-set, column parity): its 32 lanes (8 directions x 4
    // modes) write to the consumers of their items, and costs as many wavefronts as the most loaded bank holds entries.
    // Local search: swap two modes of a block, or the two blocks of a pair, while the summed cost of the stores goes down.
    {
        const size_t nb = P->blocks.size();
        std::vector<int> block_of((size_t)N, -1), slot_of((size_t)N, -1), half_of(nb, 0), pair_of(nb, -1);
        for (size_t b = 0; b < nb; ++b)
            for (int sl = 0; sl < 8; ++sl)
                if (P->blocks[b][(size_t)sl] >= 0) block_of[(size_t)P->blocks[b][(size_t)sl]] = (int)b, slot_of[(size_t)P->blocks[b][(size_t)sl]] = sl;
        for (size_t q = 0; q < P->pairs.size(); ++q) {
            if (P->pairs[q].blockE >= 0) half_of[(size_t)P->pairs[q].blockE] = 0, pair_of[(size_t)P->pairs[q].blockE] = (int)q;
            if (P->pairs[q].blockO >= 0) half_of[(size_t)P->pairs[q].blockO] = 1, pair_of[(size_t)P->pairs[q].blockO] = (int)q;
        }
        // primary consumer of (mode, direction), and the producers (mode, direction) feeding a mode
        std::vector<std::vector<std::pair<int, int>>> feeds((size_t)N);  // consumer -> (producer mode, direction)
        for (int64_t nu = 0; nu < N; ++nu)
            for (auto& it : P->prod[(size_t)nu])
                if (it.primary) feeds[(size_t)it.consumer].push_back({(int)nu, it.dir});
        auto consumer_of = [&](int mode, int dir) {
            for (auto& it : P->prod[(size_t)mode])
                if (it.primary && it.dir == dir) return it.consumer;
            return -1;
        };
        auto store_cost = [&](int b, int k, int par) {
            int cnt[16] = {0}, best = 0;
            const auto& ds = P->dsets[(size_t)P->block_dsets[(size_t)b][(size_t)k]];
            for (int c = 0; c < 4; ++c) {
                const int mode = P->blocks[(size_t)b][(size_t)(2 * c + par)];
                if (mode < 0) continue;
                for (int r = 0; r < 8; ++r) {
                    if (ds[(size_t)r] < 0) continue;
                    const int cons = consumer_of(mode, ds[(size_t)r]);
                    if (cons < 0) continue;
                    best = std::max(best, ++cnt[2 * slot_of[(size_t)cons] + half_of[(size_t)block_of[(size_t)cons]]]);
                }
            }
            return best;
        };
        auto stores_into = [&](int mode, std::vector<std::array<int, 3>>& out) {  // stores that deliver to mode
            for (auto& f : feeds[(size_t)mode]) {
                const int b = block_of[(size_t)f.first];
                for (size_t k = 0; k < P->block_dsets[(size_t)b].size(); ++k) {
                    const auto& ds = P->dsets[(size_t)P->block_dsets[(size_t)b][k]];
                    if (std::find(ds.begin(), ds.end(), f.second) != ds.end()) out.push_back({b, (int)k, slot_of[(size_t)f.first] & 1});
                }
            }
        };
        auto total_of = [&](std::vector<std::array<int, 3>>& st) {
            std::sort(st.begin(), st.end());
            st.erase(std::unique(st.begin(), st.end()), st.end());
            long t = 0;
            for (auto& x : st) t += store_cost(x[0], x[1], x[2]);
            return t;
        };
        long before_all = 0;
        for (size_t b = 0; b < nb; ++b)
            for (size_t k = 0; k < P->block_dsets[b].size(); ++k) before_all += store_cost((int)b, (int)k, 0) + store_cost((int)b, (int)k, 1);
        for (int sweep = 0; sweep < 4; ++sweep) {
            long gain = 0;
            for (size_t b = 0; b < nb; ++b) {
                for (int s1 = 0; s1 < 8; ++s1)
                    for (int s2 = s1 + 1; s2 < 8; ++s2) {
                        const int m1 = P->blocks[b][(size_t)s1], m2 = P->blocks[b][(size_t)s2];
                        if (m1 < 0 && m2 < 0) continue;
                        auto affected = [&](std::vector<std::array<int, 3>>& st) {
                            st.clear();
                            for (size_t k = 0; k < P->block_dsets[b].size(); ++k) st.push_back({(int)b, (int)k, 0}), st.push_back({(int)b, (int)k, 1});
                            if (m1 >= 0) stores_into(m1, st);
                            if (m2 >= 0) stores_into(m2, st);
                        };
                        std::vector<std::array<int, 3>> st;
                        affected(st);
                        const long c0 = total_of(st);
                        std::swap(P->blocks[b][(size_t)s1], P->blocks[b][(size_t)s2]);
                        if (m1 >= 0) slot_of[(size_t)m1] = s2;
                        if (m2 >= 0) slot_of[(size_t)m2] = s1;
                        affected(st);  // the parity of the producers' stores may have changed
                        const long c1 = total_of(st);
                        if (c1 < c0) {
                            gain += c0 - c1;
                        } else {
                            std::swap(P->blocks[b][(size_t)s1], P->blocks[b][(size_t)s2]);
                            if (m1 >= 0) slot_of[(size_t)m1] = s1;
                            if (m2 >= 0) slot_of[(size_t)m2] = s2;
                        }
                    }
            }
            for (size_t q = 0; q < P->pairs.size(); ++q) {
                PairSteps& ps = P->pairs[q];
                if (ps.blockE < 0 || ps.blockO < 0) continue;
                std::vector<std::array<int, 3>> st;
                for (int half = 0; half < 2; ++half)
                    for (int mode : P->blocks[(size_t)(half ? ps.blockO : ps.blockE)])
                        if (mode >= 0) stores_into(mode, st);
                const long c0 = total_of(st);
                std::swap(ps.blockE, ps.blockO);
                half_of[(size_t)ps.blockE] = 0, half_of[(size_t)ps.blockO] = 1;
                const long c1 = total_of(st);
                if (c1 < c0) {
                    gain += c0 - c1;
                } else {
                    std::swap(ps.blockE, ps.blockO);
                    half_of[(size_t)ps.blockE] = 0, half_of[(size_t)ps.blockO] = 1;
                }
            }
            if (gain == 0) break;
        }
        if (getenv("ASGFEM_MMA_VERBOSE")) {
            long after_all = 0, nst = 0;
            for (size_t b = 0; b < nb; ++b)
                for (size_t k = 0; k < P->block_dsets[b].size(); ++k) after_all += store_cost((int)b, (int)k, 0) + store_cost((int)b, (int)k, 1), nst += 2;
            fprintf(stderr, "[mma] mailbox stores: %ld instructions, bank cost %ld -> %ld wavefronts per row\n", nst, before_all, after_all);
        }
    }

    // ---- device column order ---------------------------------------------------------------------------------------
    ctx->h_pos.assign((size_t)N, -1);
    ctx->h_inv.assign((size_t)P->ncols, -1);
    for (size_t q = 0; q < P->pairs.size(); ++q)
        for (int half = 0; half < 2; ++half) {
            const int b = half ? P->pairs[q].blockO : P->pairs[q].blockE;
            if (b < 0) continue;
            for (int s = 0; s < 8; ++s) {
                const int mode = P->blocks[(size_t)b][(size_t)s];
                if (mode < 0) continue;
                const int col = (int)q * 16 + 2 * s + half;
                ctx->h_pos[(size_t)mode] = col;
                ctx->h_inv[(size_t)col] = mode;
            }
        }
    for (int64_t mu = 0; mu < N; ++mu)
        if (ctx->h_pos[(size_t)mu] < 0) return fail(ctx, ASGFEM_ESTATE, "internal: mode without a column");
    P->layout_ok = true;
    double steps = 0;
    for (auto& ps : P->pairs) steps += ps.nsteps;
    P->dmma_per_row = steps * 2;  // per k-step
    if (getenv("ASGFEM_MMA_VERBOSE")) {
        fprintf(stderr, "[mma] N=%lld M=%d keys=%d blocks=%zu dsets=%zu pairs=%zu steps=%.0f columns=%d\n", (long long)N, M, nkeys,
                P->blocks.size(), P->dsets.size(), P->pairs.size(), steps, P->ncols);
        std::map<int, int> hist;
        for (auto& d : P->block_dsets) hist[(int)d.size()]++;
        for (auto& h : hist) fprintf(stderr, "[mma]   blocks with %d D-sets: %d\n", h.first, h.second);
    }
    return 0;
}

bool apply_mma_layout_ok(asgfem_ctx* ctx) {
    MmaPlan* P = mp_of(ctx);
    return P && P->layout_ok;
}

// ---------------------------------------------------------------------------------------------------------------------
// kernel-side tables
// ---------------------------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------------------------
// kernel-side tables
// ---------------------------------------------------------------------------------------------------------------------
namespace {
struct StepRef {
    int pair, k;  // D-set index k of the pair's blocks
};
struct ConsRec {     // consumer record of one (pass, warp, group slot): 16 bytes
    uint32_t base;   // byte offset of the first mailbox entry of the group inside a buffer
    uint32_t roww;   // byte offset (inside the row-weight table) of the weights of the rows, and number of row PAIRS << 20
    uint32_t extra0; // byte offset of the first extra word row
    uint32_t nextra; // extra rows
};
struct GroupPlan {  // per (pass, group) while planning
    uint32_t base = 0, n0 = 0, ntail = 0, tail0 = 0, extra0 = 0, nextra = 0;
    double w0 = 0;
};
}  // namespace

int apply_mma_build(asgfem_ctx* ctx) {
    MmaPlan* P = mp_of(ctx);
    if (!P || !P->layout_ok) return 0;
    if (P->d_blob) cudaFree(P->d_blob);
    if (P->d_zero) cudaFree(P->d_zero);
    if (P->d_rowmeta) cudaFree(P->d_rowmeta);
    P->d_blob = nullptr;
    P->d_zero = nullptr;
    P->d_rowmeta = nullptr;
    P->usable = false;
    const int64_t nrows = ctx->n_owned >= 0 ? ctx->n_owned : ctx->n;
    const int M = ctx->M, Mp = M + 1;
    if (P->M > M) return 0;
    int maxlen = 1;
    for (int64_t i = 0; i < nrows; ++i) maxlen = std::max<int>(maxlen, (int)(ctx->h_rowptr[i + 1] - ctx->h_rowptr[i]));
    const int KS = maxlen <= 8 ? 2 : maxlen <= 16 ? 4 : maxlen <= 24 ? 6 : 0;
    if (!KS) return 0;
    P->KS = KS;
    const int KSTR = 4 * KS + 4;
    const int NScap = KS == 2 ? 8 : KS == 4 ? 4 : 2;  // B fragments of a stage: NS * KS double2 per lane (<= 64 registers)
    const int W = MMA_WARPS;
    const int npairs = (int)P->pairs.size(), ngroups = npairs / 2, ncols = P->ncols;
    int NG = (ngroups + W - 1) / W;
    if (NG > 8) return 0;
    NG = NG <= 2 ? 2 : NG <= 4 ? 4 : 8;  // the kernel instances
    P->NG = NG;
    // K buffers: 2 (two or more passes) or 3 (one pass); ring of 8 row records
    auto ks_bytes_of = [&](int npass) { return (size_t)(npass >= 2 ? 2 : 3) * (size_t)(Mp + 1) * (size_t)KSTR * 8ull + 8ull * (4 + 4 * KS) * 4ull; };
    auto wbits = [](double w) {
        uint64_t b;
        std::memcpy(&b, &w, 8);
        return b;
    };

    std::vector<size_t> pair_items((size_t)npairs, 0);
    size_t total_items = 0;
    int total_steps = 0;
    for (int q = 0; q < npairs; ++q) {
        for (int half = 0; half < 2; ++half) {
            const int b = half ? P->pairs[(size_t)q].blockO : P->pairs[(size_t)q].blockE;
            if (b < 0) continue;
            for (int mode : P->blocks[(size_t)b])
                if (mode >= 0) pair_items[(size_t)q] += P->prod[(size_t)mode].size();
        }
        total_items += pair_items[(size_t)q];
        total_steps += P->pairs[(size_t)q].nsteps;
    }

    // The passes are contiguous ranges of the sorted pair list; the cut positions balance alpha * items + (1 - alpha) * steps:
    // equal steps give the shortest stages, equal items the smallest mailboxes.  First fit wins.
    for (int trial = 0; trial < 64 * 5; ++trial) {
        const int npass = std::max(1, (total_steps + W * NScap - 1) / (W * NScap)) + trial / 5;
        const double alpha = 0.25 * (trial % 5);
        if (npass == 1 && trial % 5) continue;
        // ---- steps -> (pass, warp).  A pass is a contiguous range of the sorted pair list (blocks with many D-sets first)
        // holding about 1/npass of the items: producers of one kind share a pass, so the consumers of a group receive
        // similar numbers of items per pass (little mailbox padding).  Within a pass the steps go round robin over the
        // warps in (D-set) order.
        std::vector<int> pass_of((size_t)npairs, 0);
        {
            double run = 0;
            const double total = alpha * (double)total_items / std::max<size_t>(total_items, 1) + (1 - alpha);
            for (int q = 0; q < npairs; ++q) {
                pass_of[(size_t)q] = std::min(npass - 1, (int)(run * npass / total));
                run += alpha * (double)pair_items[(size_t)q] / std::max<size_t>(total_items, 1) +
                       (1 - alpha) * (double)P->pairs[(size_t)q].nsteps / std::max(total_steps, 1);
            }
        }
        std::vector<std::vector<StepRef>> bins((size_t)W * npass);  // [warp * npass + pass]
        {
            std::vector<int> next((size_t)npass, 0);
            for (int q = 0; q < npairs; ++q)
                for (int k = 0; k < P->pairs[(size_t)q].nsteps; ++k) {
                    const int pass = pass_of[(size_t)q];
                    bins[(size_t)(next[(size_t)pass]++ % W) * npass + pass].push_back({q, k});
                }
        }
        int NS = 0;
        for (auto& bn : bins) NS = std::max(NS, (int)bn.size());
        NS = (NS + 1) / 2 * 2;  // steps are processed two at a time
        if (NS > NScap) continue;

        // ---- step tables, inbox of every consumer column per pass -------------------------------------------------------
        struct Out {
            int slot, lane, which;  // slot = ((pass * W + warp) * NS + s)
            double w;
            bool primary;
        };
        std::vector<std::vector<std::vector<Out>>> inbox((size_t)npass, std::vector<std::vector<Out>>((size_t)ncols));
        const int nslots = npass * W * NS;
        std::vector<uint32_t> stepdesc((size_t)nslots * 4, 0u);  // colbase in bytes | D-sets (0 = unused slot) | store words | -
        std::vector<int> sw_of_slot((size_t)nslots, -1);
        int nreal = 0;
        for (int pass = 0; pass < npass; ++pass)
            for (int w = 0; w < W; ++w) {
                const auto& bn = bins[(size_t)w * npass + pass];
                for (size_t s = 0; s < bn.size(); ++s) {
                    const int slot = (pass * W + w) * NS + (int)s;
                    const PairSteps& ps = P->pairs[(size_t)bn[s].pair];
                    const int k = bn[s].k;
                    int dE = 0, dO = 0;
                    if (ps.blockE >= 0 && k < (int)P->block_dsets[(size_t)ps.blockE].size()) dE = P->block_dsets[(size_t)ps.blockE][(size_t)k];
                    if (ps.blockO >= 0 && k < (int)P->block_dsets[(size_t)ps.blockO].size()) dO = P->block_dsets[(size_t)ps.blockO][(size_t)k];
                    if (dE >= 65536 || dO >= 65536) return 0;
                    stepdesc[(size_t)slot * 4] = (uint32_t)(bn[s].pair * 16 * 8);
                    stepdesc[(size_t)slot * 4 + 1] = (uint32_t)dE | ((uint32_t)dO << 16);
                    stepdesc[(size_t)slot * 4 + 2] = (uint32_t)nreal * 256u;  // byte offset of the step's store words
                    sw_of_slot[(size_t)slot] = nreal++;
                    // outputs: lane l holds rows r = l/4, columns 2c, 2c+1 (c = l%4) of both halves
                    for (int lane = 0; lane < 32; ++lane) {
                        const int r = lane >> 2, c = lane & 3;
                        for (int which = 0; which < 4; ++which) {
                            const int half = which >> 1, cc = 2 * c + (which & 1);
                            const int blk = half ? ps.blockO : ps.blockE;
                            const int ds = half ? dO : dE;
                            if (blk < 0 || ds == 0) continue;
                            const int dir = P->dsets[(size_t)ds][(size_t)r];
                            const int mode = P->blocks[(size_t)blk][(size_t)cc];
                            if (dir < 0 || mode < 0) continue;
                            for (auto& it : P->prod[(size_t)mode])
                                if (it.dir == dir) inbox[(size_t)pass][(size_t)ctx->h_pos[(size_t)it.consumer]].push_back({slot, lane, which, it.w, it.primary});
                        }
                    }
                }
            }
        // ---- mailbox geometry: per (pass, group) the rows of the main weight class first, then one-weight tail rows -------
        std::vector<uint32_t> slot_addr((size_t)nslots * 128, NOSTORE);
        std::vector<GroupPlan> grec((size_t)npass * ngroups);
        std::vector<double> tailw;
        uint32_t mb_max = 0;
        bool ok = true;
        for (int pass = 0; pass < npass; ++pass) {
            uint32_t at = 0;
            for (int g = 0; g < ngroups; ++g) {
                GroupPlan& R = grec[(size_t)pass * ngroups + g];
                // weight classes of the primary items of the group with their row counts (max over the 32 columns)
                std::vector<double> cls;
                std::vector<uint32_t> cnt;
                for (int c = 0; c < 32; ++c) {
                    std::vector<uint32_t> mine(cls.size(), 0u);
                    for (auto& o : inbox[(size_t)pass][(size_t)g * 32 + c]) {
                        if (!o.primary) continue;
                        size_t k = 0;
                        while (k < cls.size() && wbits(cls[k]) != wbits(o.w)) ++k;
                        if (k == cls.size()) cls.push_back(o.w), cnt.push_back(0u), mine.push_back(0u);
                        ++mine[k];
                    }
                    for (size_t k = 0; k < mine.size(); ++k) cnt[k] = std::max(cnt[k], mine[k]);
                }
                // main class = most rows; the others become tail rows
                std::vector<size_t> order(cls.size());
                std::iota(order.begin(), order.end(), 0);
                std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return cnt[a] > cnt[b]; });
                R.base = at;
                R.tail0 = (uint32_t)tailw.size();
                uint32_t rows = 0;
                for (size_t oi = 0; oi < order.size(); ++oi) {
                    const size_t k = order[oi];
                    if (oi == 0) {
                        R.n0 = cnt[k];
                        R.w0 = cls[k];
                    } else {
                        R.ntail += cnt[k];
                        for (uint32_t j = 0; j < cnt[k]; ++j) tailw.push_back(cls[k]);
                    }
                    for (int c = 0; c < 32; ++c) {
                        uint32_t j = 0;
                        for (auto& o : inbox[(size_t)pass][(size_t)g * 32 + c])
                            if (o.primary && wbits(o.w) == wbits(cls[k])) {
                                slot_addr[((size_t)o.slot * 32 + o.lane) * 4 + o.which] = at + (rows + j) * 32u + (uint32_t)c;
                                ++j;
                            }
                    }
                    rows += cnt[k];
                }
                at += rows * 32u;
            }
            mb_max = std::max(mb_max, at);
        }
        // tail of every buffer: 32 dummy entries (stores of unneeded outputs, one per lane) and one zero entry (empty extras)
        const uint32_t zero_slot = mb_max + 32;  // behind 32 entries that the padded last row pair of the last group may read
        mb_max += 34;
        if (mb_max >= NOSTORE) ok = false;
        if (!ok) continue;
        // store words: 4 x 16 bit mailbox entries per lane and step
        std::vector<uint32_t> sw((size_t)(nreal + 1) * 64, 0xFFFFFFFFu);  // last block: no stores (unused step slots)
        int same_dsets = 0;
        for (int slot = 0; slot < nslots; ++slot) {
            if (sw_of_slot[(size_t)slot] < 0) stepdesc[(size_t)slot * 4 + 2] = (uint32_t)nreal * 256u;
            else if ((stepdesc[(size_t)slot * 4 + 1] & 0xFFFFu) == (stepdesc[(size_t)slot * 4 + 1] >> 16)) ++same_dsets;
        }
        for (int slot = 0; slot < nslots; ++slot)
            for (int lane = 0; lane < 32 && sw_of_slot[(size_t)slot] >= 0; ++lane)
                for (int which = 0; which < 4; ++which) {
                    uint32_t addr = slot_addr[((size_t)slot * 32 + lane) * 4 + which];
                    if (addr == NOSTORE) continue;
                    uint32_t& word = sw[((size_t)sw_of_slot[(size_t)slot] * 32 + lane) * 2 + (which >> 1)];
                    if (which & 1)
                        word = (word & 0x0000FFFFu) | (addr << 16);
                    else
                        word = (word & 0xFFFF0000u) | addr;
                }
        // extras: secondary consumers read the entry of the primary one; per (pass, group) rows of 32 words (entry | widx << 16)
        std::vector<double> wtab(1, 0.0);
        auto widx = [&](double w) {
            for (size_t k = 0; k < wtab.size(); ++k)
                if (wbits(wtab[k]) == wbits(w)) return (uint32_t)k;
            wtab.push_back(w);
            return (uint32_t)(wtab.size() - 1);
        };
        std::vector<uint32_t> extra;
        for (int pass = 0; pass < npass; ++pass)
            for (int g = 0; g < ngroups; ++g) {
                GroupPlan& R = grec[(size_t)pass * ngroups + g];
                uint32_t E = 0;
                for (int c = 0; c < 32; ++c) {
                    uint32_t k = 0;
                    for (auto& o : inbox[(size_t)pass][(size_t)g * 32 + c]) k += !o.primary;
                    E = std::max(E, k);
                }
                R.extra0 = (uint32_t)extra.size() / 32u;
                R.nextra = E;
                extra.resize(extra.size() + (size_t)E * 32, zero_slot);
                for (int c = 0; c < 32; ++c) {
                    uint32_t k = 0;
                    for (auto& o : inbox[(size_t)pass][(size_t)g * 32 + c])
                        if (!o.primary) {
                            const uint32_t addr = slot_addr[((size_t)o.slot * 32 + o.lane) * 4 + o.which];
                            if (addr == NOSTORE) return fail(ctx, ASGFEM_ESTATE, "internal: secondary item without a primary slot");
                            if (wtab.size() >= 65535) return 0;
                            extra[((size_t)R.extra0 + k) * 32 + c] = addr | (widx(o.w) << 16);
                            ++k;
                        }
                }
            }
        // consumer groups -> warps (longest processing time first)
        std::vector<int> gorder((size_t)ngroups);
        std::iota(gorder.begin(), gorder.end(), 0);
        std::vector<int> gload((size_t)ngroups, 0);
        uint32_t total_rows = 0;
        for (int g = 0; g < ngroups; ++g)
            for (int pass = 0; pass < npass; ++pass) {
                const GroupPlan& R = grec[(size_t)pass * ngroups + g];
                gload[(size_t)g] += (int)(3 * (R.n0 + R.ntail) + 8 * R.nextra + 8);
                total_rows += R.n0 + R.ntail;
            }
        std::stable_sort(gorder.begin(), gorder.end(), [&](int a, int b) { return gload[(size_t)a] > gload[(size_t)b]; });
        std::vector<std::vector<int>> wg((size_t)W);
        std::vector<int> wload((size_t)W, 0);
        for (int g : gorder) {
            int best = -1;
            for (int w = 0; w < W; ++w)
                if ((int)wg[(size_t)w].size() < NG && (best < 0 || wload[(size_t)w] < wload[(size_t)best])) best = w;
            wg[(size_t)best].push_back(g);
            wload[(size_t)best] += gload[(size_t)g];
        }
        // consumer records [pass][warp][NG] and the weights of the mailbox rows (padded to row pairs with weight 0)
        std::vector<ConsRec> crec((size_t)npass * W * NG, ConsRec{0u, 0u, 0u, 0u});
        std::vector<double> roww;
        std::vector<uint32_t> gcol((size_t)W * NG, 0xFFFFFFFFu);
        for (int w = 0; w < W; ++w)
            for (size_t k = 0; k < wg[(size_t)w].size(); ++k) {
                const int g = wg[(size_t)w][k];
                gcol[(size_t)w * NG + k] = (uint32_t)g * 32u * 8u;  // byte offset of the group's columns in a row of Y
                for (int pass = 0; pass < npass; ++pass) {
                    const GroupPlan& R = grec[(size_t)pass * ngroups + g];
                    ConsRec& c = crec[((size_t)pass * W + w) * NG + k];
                    const uint32_t rows = R.n0 + R.ntail, pairs2 = (rows + 1) / 2;
                    if (roww.size() * 8 >= (1u << 20) || pairs2 >= 4096) return 0;
                    c.base = R.base * 8u;
                    c.roww = (uint32_t)(roww.size() * 8) | (pairs2 << 20);
                    for (uint32_t r = 0; r < R.n0; ++r) roww.push_back(R.w0);
                    for (uint32_t r = 0; r < R.ntail; ++r) roww.push_back(tailw[R.tail0 + r]);
                    if (rows & 1) roww.push_back(0.0);
                    c.extra0 = R.extra0 * 128u;
                    c.nextra = R.nextra;
                }
            }
        // D-set table: byte offset of the K row of (D-set, row r) inside a K buffer
        std::vector<uint32_t> dtab(P->dsets.size() * 8);
        for (size_t d = 0; d < P->dsets.size(); ++d)
            for (int r = 0; r < 8; ++r) {
                const int dir = P->dsets[d][(size_t)r];
                dtab[d * 8 + r] = (uint32_t)((dir < 0 ? Mp : dir) * KSTR * 8);  // null row: the zero row behind K_M
            }
        // blob
        auto align4 = [](uint32_t v) { return (v + 3u) & ~3u; };
        uint32_t at = 0;
        P->off_sw = at;
        at = align4(at + (uint32_t)sw.size());
        P->off_dtab = at;
        at = align4(at + (uint32_t)dtab.size());
        P->off_extra = at;
        at = align4(at + (uint32_t)extra.size() + 4u);
        P->nwords = at;
        const size_t smem = (size_t)at * 4 + ks_bytes_of(npass) + 2ull * mb_max * 8ull + 16;
        if (getenv("ASGFEM_MMA_VERBOSE")) {
            size_t items = 0;
            for (auto& pr : P->prod) items += pr.size();
            fprintf(stderr,
                    "[mma] KS=%d passes=%d alpha=%.2f NS=%d NG=%d steps=%d items=%zu mailbox=%u doubles (x2), rows %u (ideal %zu) tail rows=%zu "
                    "extras=%zu steps with one D-set for both halves=%d tables=%u B smem=%zu B%s\n",
                    KS, npass, alpha, NS, NG, total_steps, items, mb_max, total_rows, (items + 31) / 32, tailw.size(), extra.size() / 32, same_dsets,
                    at * 4, smem, smem > (size_t)SMEM_LIMIT ? " (too large)" : "");
        }
        if (smem > (size_t)SMEM_LIMIT) continue;  // more passes: smaller mailboxes
        std::vector<uint32_t> blob((size_t)at, 0u);
        std::memcpy(&blob[P->off_sw], sw.data(), sw.size() * 4);
        std::memcpy(&blob[P->off_dtab], dtab.data(), dtab.size() * 4);
        if (!extra.empty()) std::memcpy(&blob[P->off_extra], extra.data(), extra.size() * 4);
        // warp-uniform tables with the fixed strides of the kernel parameter block
        if (npass > MMA_MAXP || roww.size() > (size_t)MMA_MAXROWW || wtab.size() > (size_t)MMA_MAXW) continue;
        P->h_step.assign((size_t)MMA_MAXP * W * 8 * 4, 0u);
        P->h_crec.assign((size_t)MMA_MAXP * W * 8 * 4, 0u);
        P->h_gcol.assign((size_t)W * 8, 0xFFFFFFFFu);
        for (int pass = 0; pass < npass; ++pass)
            for (int w = 0; w < W; ++w) {
                for (int st = 0; st < NS; ++st)
                    std::memcpy(&P->h_step[(((size_t)pass * W + w) * 8 + st) * 4], &stepdesc[(((size_t)pass * W + w) * NS + st) * 4], 16);
                for (int g = 0; g < NG; ++g)
                    std::memcpy(&P->h_crec[(((size_t)pass * W + w) * 8 + g) * 4], &crec[((size_t)pass * W + w) * NG + g], 16);
            }
        for (int w = 0; w < W; ++w)
            for (int g = 0; g < NG; ++g) P->h_gcol[(size_t)w * 8 + g] = gcol[(size_t)w * NG + g];
        P->h_roww = roww;
        P->h_wtab = wtab;
        ASG_CUDA(ctx, cudaMalloc((void**)&P->d_blob, blob.size() * 4));
        ASG_CUDA(ctx, cudaMemcpyAsync(P->d_blob, blob.data(), blob.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        {
            // row records: what the kernel needs of the CSR structure of a row, in one contiguous piece (fetched two rows ahead)
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
            ASG_CUDA(ctx, cudaMalloc((void**)&P->d_rowmeta, meta.size() * 4));
            ASG_CUDA(ctx, cudaMemcpyAsync(P->d_rowmeta, meta.data(), meta.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
            ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
        ASG_CUDA(ctx, cudaMalloc((void**)&P->d_zero, sizeof(double) * (size_t)ctx->ld));
        ASG_CUDA(ctx, cudaMemsetAsync(P->d_zero, 0, sizeof(double) * (size_t)ctx->ld, ctx->stream));
        ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        P->P = npass;
        P->NS = NS;
        P->nsteps = total_steps;
        P->mb_doubles = mb_max;
        P->smem_bytes = smem;
        P->usable = true;
        P->grid = 148;
        if (const char* e = getenv("ASGFEM_MMA_GRID")) {
            int v = atoi(e);
            if (v >= 1 && v <= 148) P->grid = v;
        }
        return 0;
    }
    return 0;
}

bool apply_mma_usable(asgfem_ctx* ctx) {
    MmaPlan* P = mp_of(ctx);
    return P && P->usable;
}

// ---------------------------------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------------------------------
namespace {

struct MmaArgs {
    const double* x;
    double* y;
    const double* vals;
    const int64_t* rowptr;
    const int32_t* col;
    const uint8_t* bmask;
    const uint32_t* blob;
    const double* zero_row;
    const int32_t* rowmeta;
    int64_t nnz, ld, r0, r1;
    int Mp, P, zero_after_read, debug_skip;  // debug_skip: 1 = no products, 2 = no consumer sums (timing experiments only)
    uint32_t nwords, off_sw, off_dtab, off_extra, mb_doubles;
};

// warp-uniform tables: kernel parameter (constant bank), so that their loads stay off the shared-memory pipe
struct MmaTables {
    uint4 step[MMA_MAXP * MMA_WARPS * 8];  // [pass][warp][8]: colbase bytes | D-sets | store word offset | -
    uint4 crec[MMA_MAXP * MMA_WARPS * 8];  // [pass][warp][8]: ConsRec
    double roww[MMA_MAXROWW];
    double wtab[MMA_MAXW];
    uint32_t gcol[MMA_WARPS * 8];
};

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
struct ConsRecD {  // device view of ConsRec
    uint32_t base, roww, extra0, nextra;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

template <int KS, int NS, int NG>
__global__ void __launch_bounds__(MMA_THREADS, 1) k_apply_mma(const MmaArgs a, const __grid_constant__ MmaTables tab) {
    extern __shared__ __align__(16) unsigned char sm[];
    constexpr int KSTR = 4 * KS + 4;
    constexpr int ME = 4 + 4 * KS;  // ints per row record
    constexpr int RING = 8;         // row records in flight
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = lane >> 2, kk = lane & 3;
    const int LA = a.P >= 2 ? 1 : 2;  // rows of lookahead: data issued in stage t is complete at the end of stage t + 1
    const int KB = LA + 1;            // K buffers
    const uint32_t kbuf_bytes = (uint32_t)(a.Mp + 1) * KSTR * 8u;  // one K buffer (directions 0..M and the null row)
    const uint32_t meta_off = a.nwords * 4u, ks_off = meta_off + RING * ME * 4u, mb_off = ks_off + KB * kbuf_bytes,
                   mb_bytes = a.mb_doubles * 8u;

    // rows of this CTA: r0 + blockIdx.x + k * gridDim.x.  The CTAs walk the mesh side by side, so the X rows of the
    // neighbouring mesh lines (read again a few hundred rows later) are still in L2.
    const int64_t rstep = gridDim.x;
    const int64_t rb = a.r0 + (int64_t)blockIdx.x;
    if (rb >= a.r1) return;
    const int64_t nri = (a.r1 - rb + rstep - 1) / rstep;  // rows of this CTA

    {
        uint32_t* blob = reinterpret_cast<uint32_t*>(sm);
        for (uint32_t i = tid; i < a.nwords; i += MMA_THREADS) blob[i] = a.blob[i];
        double* z = reinterpret_cast<double*>(sm + meta_off);
        for (uint32_t i = tid; i < (RING * ME * 4u + KB * kbuf_bytes + 2u * mb_bytes) / 8u; i += MMA_THREADS) z[i] = 0.0;
    }
    __syncthreads();

    const unsigned char* dtab_l = sm + a.off_dtab * 4u + q * 4;   // + D-set * 32: byte offset of this lane's K row
    const unsigned char* sw_l = sm + a.off_sw * 4u + lane * 8;     // + step * 256: store words of this lane
    const unsigned char* ks_l = sm + ks_off + kk * 8;              // + buffer + row offset + 32 s: A fragment entries
    const uint4* stepd = tab.step + warp * 8;
    const uint4* crec = tab.crec + warp * 8;
    const unsigned char* extra_l = sm + a.off_extra * 4u + lane * 4;
    unsigned char* mb0 = sm + mb_off;
    int32_t* meta = reinterpret_cast<int32_t*>(sm + meta_off);

    uint32_t ycol[NG];  // byte offset of this thread's column of group g inside a row of Y (0xFFFFFFFF: unused slot)
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        const uint32_t c = tab.gcol[warp * 8 + g];
        ycol[g] = c == 0xFFFFFFFFu ? c : c + (uint32_t)lane * 8u;
    }

    // row record ri (first CSR position, length, columns) -> ring slot ri % RING, asynchronously
    auto meta_fetch = [&](int64_t ri) {
        if (tid < ME / 4) cp_async16(meta + (ri % RING) * ME + tid * 4, a.rowmeta + (rb + ri * rstep) * ME + tid * 4);
    };
    // K rows of row ri -> Ks[ri % KB][m][k] (k >= length of the row: 0); needs the row record; one element per thread and trip
    const int sk_m = tid / (4 * KS), sk_k = tid - sk_m * (4 * KS);
    auto stage_k = [&](int64_t ri) {
        const int32_t* m = meta + (ri % RING) * ME;
        const int64_t p0 = *reinterpret_cast<const int64_t*>(m);
        const int len = m[2];
        double* dst = reinterpret_cast<double*>(sm + ks_off + (uint32_t)(ri % KB) * kbuf_bytes);
        for (int mm = sk_m; mm < a.Mp; mm += MMA_THREADS / (4 * KS)) {
            if (sk_k < len)
                cp_async8(dst + mm * KSTR + sk_k, a.vals + (int64_t)mm * a.nnz + p0 + sk_k);
            else
                dst[mm * KSTR + sk_k] = 0.0;
        }
    };
    // X rows this lane LOADS for a dof row.  The B fragment wants lane 4 q + kk to hold columns (2q, 2q+1) of the X row of
    // slot 4s + kk, i.e. the four lanes of a quad read four different rows - the L1 data stage then spends one wavefront per
    // 32-byte sector (measured: 16 per load instruction).  So lane l loads the 16 bytes (l & 7) of the row of slot
    // 4s + (l >> 3) (eight consecutive lanes = one 128-byte line, 4 wavefronts) and the fragments are permuted with
    // shuffles when they are used (lane 4q + kk <- lane 8 kk + q).  Slots beyond the row read a row of zeros.
    const int lrow = lane >> 3, lchunk = lane & 7;
    const int frag_src = (kk << 3) | q;
    auto row_ptrs = [&](int64_t ri, const char* (&xr)[KS]) {
        const int32_t* m = meta + (ri % RING) * ME;
        const int len = m[2];
#pragma unroll
        for (int s = 0; s < KS; ++s) {
            const int slot = 4 * s + lrow;
            xr[s] = reinterpret_cast<const char*>((slot < len ? a.x + (int64_t)m[4 + slot] * a.ld : a.zero_row) + 2 * lchunk);
        }
    };

    double2 B[NS][KS];
    double acc[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) acc[g] = 0.0;

    auto load_b = [&](int pass, const char* const (&xr)[KS]) {
        const uint4* sd = stepd + pass * (MMA_WARPS * 8);
#pragma unroll
        for (int st = 0; st < NS; ++st) {
            const uint4 d = sd[st];
            if (d.y != 0) {
#pragma unroll
                for (int s = 0; s < KS; ++s) B[st][s] = *reinterpret_cast<const double2*>(xr[s] + d.x);
            }
        }
    };

    // prologue: row records of the first 2 LA rows, K rows of the first LA rows, B fragments of the first stage
    for (int64_t j = 0; j < 2 * LA && j < nri; ++j) meta_fetch(j);
    cp_async_wait_all();
    __syncthreads();
    for (int64_t j = 0; j < LA && j < nri; ++j) stage_k(j);
    {
        const char* xr[KS];
        row_ptrs(0, xr);
        load_b(0, xr);
    }
    cp_async_wait_all();
    __syncthreads();

    int pass = 0;
    int64_t ri = 0;
    uint32_t par = 0;  // parity of the stage = mailbox buffer
    bool have_prev = false;
    int cpass = 0;
    int64_t cri = 0;
    while (ri < nri || have_prev) {
        const bool produce = ri < nri;
        // odd warps consume first: the shared-memory reads of one half of the warps overlap the fp64 products of the other
        for (int phase = 0; phase < 2; ++phase) {
            if (phase == (warp & 1)) {
                if (produce) {
                    if (pass == 0) {
                        if (ri + 2 * LA < nri) meta_fetch(ri + 2 * LA);
                        if (ri + LA < nri) stage_k(ri + LA);
                    }
                    // pointers of the next stage's row
                    int npass = pass + 1;
                    int64_t nxt = ri;
                    if (npass == a.P) npass = 0, ++nxt;
                    const char* xr[KS];
                    if (nxt < nri) row_ptrs(nxt, xr);

                    // ---- produce: NS steps = (pair of home blocks, D-set of each); outputs go to the mailbox of this stage ---
                    unsigned char* mb = mb0 + par * mb_bytes;
                    const unsigned char* ksrc = ks_l + (uint32_t)(ri % KB) * kbuf_bytes;
                    const uint4* sd = stepd + pass * (MMA_WARPS * 8);
#pragma unroll
                    for (int st = 0; st < NS; st += 2) {
                        const uint4 d0 = sd[st], d1 = sd[st + 1];
                        if (d0.y != 0 && !(a.debug_skip & 1)) {  // warp-uniform; an unused partner slot computes zeros, stores nothing
                            const uint4 dd[2] = {d0, d1};
                            double aE[2][KS], aO[2][KS];
                            uint2 w[2];
#pragma unroll
                            for (int u = 0; u < 2; ++u) {
                                const uint32_t oE = *reinterpret_cast<const uint32_t*>(dtab_l + (dd[u].y & 0xFFFFu) * 32u);
                                const uint32_t oO = *reinterpret_cast<const uint32_t*>(dtab_l + (dd[u].y >> 16) * 32u);
#pragma unroll
                                for (int s = 0; s < KS; ++s) {
                                    aE[u][s] = *reinterpret_cast<const double*>(ksrc + oE + 32 * s);
                                    aO[u][s] = *reinterpret_cast<const double*>(ksrc + oO + 32 * s);
                                }
                                w[u] = *reinterpret_cast<const uint2*>(sw_l + dd[u].z);
                            }
                            double c[2][4];
#pragma unroll
                            for (int u = 0; u < 2; ++u) c[u][0] = c[u][1] = c[u][2] = c[u][3] = 0.0;
#pragma unroll
                            for (int s = 0; s < KS; ++s)
#pragma unroll
                                for (int u = 0; u < 2; ++u) {
                                    const double bx = __shfl_sync(0xffffffffu, B[st + u][s].x, frag_src);
                                    const double by = __shfl_sync(0xffffffffu, B[st + u][s].y, frag_src);
                                    dmma(c[u][0], c[u][1], aE[u][s], bx);
                                    dmma(c[u][2], c[u][3], aO[u][s], by);
                                }
#pragma unroll
                            for (int u = 0; u < 2; ++u) {
                                const uint32_t a0 = w[u].x & 0xFFFFu, a1 = w[u].x >> 16, a2 = w[u].y & 0xFFFFu, a3 = w[u].y >> 16;
                                if (a0 != NOSTORE) *reinterpret_cast<double*>(mb + a0 * 8u) = c[u][0];
                                if (a1 != NOSTORE) *reinterpret_cast<double*>(mb + a1 * 8u) = c[u][1];
                                if (a2 != NOSTORE) *reinterpret_cast<double*>(mb + a2 * 8u) = c[u][2];
                                if (a3 != NOSTORE) *reinterpret_cast<double*>(mb + a3 * 8u) = c[u][3];
                            }
                        }
                    }
                    // ---- B fragments of the next stage ------------------------------------------------------------------
                    if (nxt < nri && !(a.debug_skip & 4)) load_b(npass, xr);
                }
            } else if (have_prev) {
                // ---- consume the previous stage from the other mailbox: weighted column sums of the groups of this warp ------
                const int64_t crow = rb + cri * rstep;
                const bool last = cpass == a.P - 1;
                uint8_t bm = 0;
                if (last) bm = a.bmask[crow];  // in flight during the sums
                unsigned char* mb = mb0 + (par ^ 1u) * mb_bytes;
                const uint4* cr = crec + cpass * (MMA_WARPS * 8);
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const uint4 hh = cr[g];
                    const ConsRecD h = {hh.x, hh.y, hh.z, hh.w};
                    const unsigned char* src = mb + h.base + lane * 8;
                    const unsigned char* wr = reinterpret_cast<const unsigned char*>(tab.roww) + (h.roww & 0xFFFFFu);
                    const uint32_t npair = (a.debug_skip & 2) ? 0u : h.roww >> 20;
                    double t0 = acc[g], t1 = 0.0;
                    uint32_t r = npair;
                    while (r >= 4) {  // 8 mailbox rows per trip: independent loads first
                        double v[8];
                        double2 w2[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            w2[u] = *reinterpret_cast<const double2*>(wr + u * 16);
                            v[2 * u] = *reinterpret_cast<const double*>(src + u * 512);
                            v[2 * u + 1] = *reinterpret_cast<const double*>(src + u * 512 + 256);
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            t0 = fma(w2[u].x, v[2 * u], t0);
                            t1 = fma(w2[u].y, v[2 * u + 1], t1);
                        }
                        wr += 64, src += 2048, r -= 4;
                    }
                    if (r & 2) {
                        const double2 wa = *reinterpret_cast<const double2*>(wr), wb = *reinterpret_cast<const double2*>(wr + 16);
                        const double v0 = *reinterpret_cast<const double*>(src), v1 = *reinterpret_cast<const double*>(src + 256),
                                     v2 = *reinterpret_cast<const double*>(src + 512), v3 = *reinterpret_cast<const double*>(src + 768);
                        t0 = fma(wa.x, v0, t0);
                        t1 = fma(wa.y, v1, t1);
                        t0 = fma(wb.x, v2, t0);
                        t1 = fma(wb.y, v3, t1);
                        wr += 32, src += 1024;
                    }
                    if (r & 1) {
                        const double2 wa = *reinterpret_cast<const double2*>(wr);
                        t0 = fma(wa.x, *reinterpret_cast<const double*>(src), t0);
                        t1 = fma(wa.y, *reinterpret_cast<const double*>(src + 256), t1);
                    }
                    const unsigned char* ex = extra_l + h.extra0;
                    for (uint32_t e = 0; e < h.nextra; ++e) {
                        const uint32_t word = *reinterpret_cast<const uint32_t*>(ex + e * 128);
                        t1 = fma(tab.wtab[word >> 16], *reinterpret_cast<const double*>(mb + (word & 0xFFFFu) * 8u), t1);
                    }
                    acc[g] = t0 + t1;
                }
                if (a.zero_after_read) {
                    // more than two passes: a buffer serves passes with different layouts, so the entries read here are zeroed
                    // again (after all consumers, incl. the extra lists of other warps, are through)
                    __syncthreads();
#pragma unroll
                    for (int g = 0; g < NG; ++g) {
                        const uint4 hh = cr[g];
                        unsigned char* dst = mb + hh.x + lane * 8;
                        for (uint32_t r = 0; r < 2 * (hh.y >> 20); ++r) *reinterpret_cast<double*>(dst + r * 256) = 0.0;
                    }
                }
                if (last) {
                    unsigned char* yr = reinterpret_cast<unsigned char*>(a.y + crow * a.ld);
#pragma unroll
                    for (int g = 0; g < NG; ++g) {
                        if (ycol[g] != 0xFFFFFFFFu) *reinterpret_cast<double*>(yr + ycol[g]) = bm ? 0.0 : acc[g];
                        acc[g] = 0.0;
                    }
                }
            }
        }
        cp_async_commit();
        cp_async_wait_1();  // everything but the copies issued in this stage has landed
        __syncthreads();
        have_prev = produce;
        cpass = pass;
        cri = ri;
        par ^= 1u;
        if (produce && ++pass == a.P) pass = 0, ++ri;
    }
}

template <int KS, int NS, int NG>
int launch_mma(asgfem_ctx* ctx, MmaPlan* P, const MmaArgs& a, const MmaTables& tab) {
    static bool configured = false;
    if (!configured) {
        ASG_CUDA(ctx, cudaFuncSetAttribute(k_apply_mma<KS, NS, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        configured = true;
    }
    const int64_t nr = a.r1 - a.r0;
    const int grid = (int)std::min<int64_t>(P->grid, nr);
    k_apply_mma<KS, NS, NG><<<grid, MMA_THREADS, P->smem_bytes, ctx->stream>>>(a, tab);
    ASG_CUDA(ctx, cudaGetLastError());
    return 0;
}

template <int KS, int NS>
int launch_mma_ng(asgfem_ctx* ctx, MmaPlan* P, const MmaArgs& a, const MmaTables& tab) {
    if (P->NG <= 2) return launch_mma<KS, NS, 2>(ctx, P, a, tab);
    if (P->NG <= 4) return launch_mma<KS, NS, 4>(ctx, P, a, tab);
    return launch_mma<KS, NS, 8>(ctx, P, a, tab);
}

}  // namespace

int apply_mma_launch(asgfem_ctx* ctx, const double* x, double* y, int64_t r0, int64_t r1) {
    MmaPlan* P = mp_of(ctx);
    if (!P || !P->usable) return fail(ctx, ASGFEM_ESTATE, "MMA operator plan not available for this pattern / multi-index set");
    MmaArgs a;
    a.x = x;
    a.y = y;
    a.vals = ctx->d_vals;
    a.rowptr = ctx->d_rowptr;
    a.col = ctx->d_col;
    a.bmask = ctx->d_bmask;
    a.blob = P->d_blob;
    a.zero_row = P->d_zero;
    a.rowmeta = P->d_rowmeta;
    a.nnz = ctx->nnz;
    a.ld = ctx->ld;
    a.r0 = r0;
    a.r1 = r1;
    a.Mp = ctx->M + 1;
    a.P = P->P;
    a.debug_skip = getenv("ASGFEM_MMA_SKIP") ? atoi(getenv("ASGFEM_MMA_SKIP")) : 0;
    a.zero_after_read = P->P > 2 ? 1 : 0;  // a mailbox buffer serves passes with different layouts: padding must stay zero
    a.nwords = P->nwords;
    a.off_sw = P->off_sw;
    a.off_dtab = P->off_dtab;
    a.off_extra = P->off_extra;
    a.mb_doubles = P->mb_doubles;
    static MmaTables tab;  // 28 KB: filled per launch from the plan (host copies only)
    std::memcpy(tab.step, P->h_step.data(), sizeof(tab.step));
    std::memcpy(tab.crec, P->h_crec.data(), sizeof(tab.crec));
    std::memset(tab.roww, 0, sizeof(tab.roww));
    std::memcpy(tab.roww, P->h_roww.data(), P->h_roww.size() * 8);
    std::memset(tab.wtab, 0, sizeof(tab.wtab));
    std::memcpy(tab.wtab, P->h_wtab.data(), P->h_wtab.size() * 8);
    std::memcpy(tab.gcol, P->h_gcol.data(), sizeof(tab.gcol));
    switch (P->KS * 16 + P->NS) {
        case 2 * 16 + 2: return launch_mma_ng<2, 2>(ctx, P, a, tab);
        case 2 * 16 + 4: return launch_mma_ng<2, 4>(ctx, P, a, tab);
        case 2 * 16 + 6: return launch_mma_ng<2, 6>(ctx, P, a, tab);
        case 2 * 16 + 8: return launch_mma_ng<2, 8>(ctx, P, a, tab);
        case 4 * 16 + 2: return launch_mma_ng<4, 2>(ctx, P, a, tab);
        case 4 * 16 + 4: return launch_mma_ng<4, 4>(ctx, P, a, tab);
        case 6 * 16 + 2: return launch_mma_ng<6, 2>(ctx, P, a, tab);
        default: return fail(ctx, ASGFEM_ESTATE, "MMA operator: no kernel instance for this shape");
    }
}

}  // namespace asgfem
