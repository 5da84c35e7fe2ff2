// Mode-side plan of the block-product operator kernel (apply_blk.cu) and the DEVICE COLUMN ORDER of all vectors.
//
//   Y[i, mu] = sum_k K_0[i,j_k] X[j_k,mu] + sum_{(m,nu) ~ mu} g sum_k K_m[i,j_k] X[j_k,nu]      (mul!, :101-117)
//
// For one dof row i the partial products T[m, nu] = sum_k K_m[i, j_k] X[j_k, nu] form the matrix product
// (K-rows of the row: (M+1) x len) x (X rows of the neighbours: len x N), of which only the pairs (m, nu) with a coupling
// are needed (30 % on the benchmark set).  Built at asgfem_set_multiindices, independent of the mesh:
//   * every needed product T[m, nu] is an ITEM (producer mode nu, direction m) with a consumer mode mu and weight g (the
//     mean term is the item (mu, 0) with consumer mu and weight 1; a product may have two consumers, nu - e_m and nu + e_m);
//   * modes are clustered greedily into HOME BLOCKS of 8 modes whose direction sets overlap, so that the union of the
//     directions of a block needs few D-SETS of 8 directions (benchmark set: 250 blocks, 319 D-set uses = 638 DMMA m8n8k4
//     per dof row; lower bound 630);
//   * two blocks with equally many D-sets form a PAIR = 16 consecutive columns of the device layout (even columns = first
//     block): one 16-byte load per lane fetches the X fragments of both.  The device column order of ALL vectors is this
//     order (ctx->h_pos; the layout is private, converted at upload / download).
// The mailbox kernel that first used this plan (variant 8 of round 2, 72 ms at config 4) is retired; its measurements are
// in profiles/README.md.
#include <algorithm>
#include <cstring>
#include <map>
#include <numeric>

#include "common.h"
#include "apply_mma_plan.h"

namespace asgfem {


void apply_mma_free(asgfem_ctx* ctx) {
    MmaPlan* P = mp_of(ctx);
    if (!P) return;
    apply_blk_free(ctx);
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


}  // namespace asgfem
