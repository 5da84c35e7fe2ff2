// variant 7 of the fused SGFE operator: mode-stationary lanes with PACKED direction units (default kernel).
//
//   Y[i, mu] = sum_k K_0[i,j_k] X[j_k,mu] + sum_{(m,nu) ~ mu} g sum_k K_m[i,j_k] X[j_k,nu]      (mul!, :101-117)
//
// Same two-phase scheme as variant 6 (apply_ts.cu): every lane OWNS modes (32 consecutive modes = one group, four groups
// = slots per warp, 16 warps), keeps X[j_k, nu] of its modes in registers, forms T_m[nu] = sum_k K_m[i,j_k] X[j_k,nu] for
// the pairs (m, nu) that have a coupling, exchanges them through shared memory and gathers g * T_m[nu] into Y[i, mu].
// Variant 6 evaluates a direction for all four groups of a warp as soon as one lane needs it: on the benchmark set
// (total degree <= 3 in 20 dimensions + part of degree 4: 252 modes with all 20 directions, 1748 modes with <= 3) only
// 29 % of the evaluated lanes are needed.  Here a group is one of two kinds, decided on the host:
//
//   dense   one unit per direction that any lane of the group needs; K_m[i, .] is the same for all lanes (broadcast
//           16-byte loads), all 32 lanes store T into a full block of the exchange buffer (no predicate).  The block of
//           the k-th unit is rotated by k lanes, so that the entries of ONE mode for consecutive directions - gathered by
//           consecutive lanes of a sparse group - lie in different banks;
//   sparse  every lane of the group needs at most Q directions: Q units, in unit q lane l handles ITS q-th direction.
//           The K row is addressed per lane (8-byte loads from a second copy of the K rows with an odd stride, so that
//           rows of different directions start in different banks); the per-lane constants (K row offset, T index) live
//           in Q registers per slot, the T blocks are first-fit packed; lanes with fewer directions compute on K_0 and
//           store to a dummy entry.
//
// A unit is 7 (8) FMAs on register operands + the K loads + 1 store; the benchmark set needs 388 units per row instead of
// the 1276 that variant 6 evaluates.  Units are processed several at a time (independent FMA chains).
// Phase 2 gathers with per-lane lists whose ORDER is chosen on the host so that the lanes of a half-warp read 16
// different banks.  Phase 2 of a dense group (a long list) does not need the X registers of the group, so it runs in a
// warp with little phase-1 work as that warp's gather-only slot; the owner exports the mean term of the group through
// the exchange buffer (one more list entry with weight 1).  One block barrier per row (double-buffered exchange), no
// atomics, fixed summation order.  The CSR metadata rings (warp 0, cp.async two / three rows ahead, issued at the top of
// the row), the cp.async staging of the next row's K values and the X loads of the next row issued before phase 2 hide
// the global-memory latency behind the arithmetic of the current row.
// Measured history and the rejected alternatives: DESIGN.md section 4, profiles/README.md.
#include <algorithm>
#include <numeric>

#include "common.h"

namespace asgfem {

namespace {
constexpr int TS2_DUMMY = 50;  // exchange-buffer tail: entries D..D+15 = 0.0 (padding of the gather lists, one bank per
                               // lane of a half-warp), D+16..D+47 dummy store block
}  // namespace

struct Ts2Plan {
    bool usable = false;
    int warps = 0, slots = 4, NS = 0, nbuf = 1, kstr = 0, kstr2 = 0, nchunk_max = 1, Q = 3;
    int D = 0;
    int nwords = 0, ndl = 0;
    int units = 0;  // evaluated units per row (statistics)
    int grid_sms = 148;  // CTAs of the persistent grid (ASGFEM_TS2_GRID: fewer when collectives need SMs of their own)
    size_t smem_bytes = 0;
    int32_t* d_slotinfo = nullptr;  // [warps*(slots+1)*8] group (-1 unused), kind (1 dense, 2 sparse, 3 dense with exported
                                    // phase 2), #dense units, list base, words base, jmax, byte offset of the mean-term block
    uint32_t* d_lane = nullptr;     // [warps*4*4*32] sparse lanes: K row byte offset | T index << 15 per unit
    uint32_t* d_dl = nullptr;       // dense unit lists (same packing, T index of the block start)
    uint32_t* d_words = nullptr;    // phase-2 words: T index << 12 | weight index << 3
    double* d_gtab = nullptr;       // [64]
};

static Ts2Plan* tp2_of(asgfem_ctx* ctx) { return reinterpret_cast<Ts2Plan*>(ctx->ts2plan); }

void apply_ts2_free(asgfem_ctx* ctx) {
    Ts2Plan* P = tp2_of(ctx);
    if (!P) return;
    void* ptrs[] = {P->d_slotinfo, P->d_lane, P->d_dl, P->d_words, P->d_gtab};
    for (void* q : ptrs)
        if (q) cudaFree(q);
    delete P;
    ctx->ts2plan = nullptr;
}

struct Ts2Layout {
    uint32_t ts, ks, ks2, words, dl, sinfo, total, dpad;
};
static __host__ __device__ inline Ts2Layout ts2_layout(int D, int nbuf, int Mp, int kstr, int kstr2, int nwords, int ndl,
                                                       int warps, int slots) {
    Ts2Layout L;
    uint32_t at = 0;
    L.dpad = (uint32_t)((D + TS2_DUMMY + 1) / 2 * 2);
    L.ts = at;
    at += (uint32_t)nbuf * L.dpad * 8u;
    L.ks = at;
    at += 2u * (uint32_t)Mp * (uint32_t)kstr * 8u;
    L.ks2 = at;  // second copy of the K rows with an odd stride (per-lane 8-byte loads of the sparse units)
    at += 2u * (uint32_t)Mp * (uint32_t)kstr2 * 8u;
    L.words = at;
    at += (uint32_t)nwords * 4u;
    at = (at + 15u) & ~15u;
    L.dl = at;
    at += (uint32_t)ndl * 4u;
    at = (at + 15u) & ~15u;
    L.sinfo = at;
    at += (uint32_t)warps * (uint32_t)(slots + 1) * 32u;  // + the gather-only slot of every warp
    L.total = at + 16u;
    return L;
}

int apply_ts2_build(asgfem_ctx* ctx) {
    apply_ts2_free(ctx);
    Ts2Plan* P = new Ts2Plan();
    ctx->ts2plan = P;
    // works in device column space: "mode" c = column c of the vectors (ctx->coup_col), all ld columns are written
    const int64_t N = ctx->ld, nrows = ctx->n_owned >= 0 ? ctx->n_owned : ctx->n;
    const int M = ctx->M, Mp = M + 1;
    const Coupling& C = ctx->coup_col;
    if (N <= 0 || M < 0 || M > 63 || ctx->sample_mode) return 0;  // per-sample weights do not fit the 64-entry weight table
    int maxlen = 1;
    for (int64_t i = 0; i < nrows; ++i) maxlen = std::max<int>(maxlen, (int)(ctx->h_rowptr[i + 1] - ctx->h_rowptr[i]));
    P->NS = maxlen <= 7 ? 7 : 8;
    P->nchunk_max = (maxlen + P->NS - 1) / P->NS;
    P->kstr = (P->nchunk_max * P->NS + 1) / 2 * 2;
    if ((P->kstr / 2) % 2 == 0) P->kstr += 2;  // odd number of 16-byte granules per K row
    P->kstr2 = P->nchunk_max * P->NS;
    if (P->kstr2 % 2 == 0) ++P->kstr2;  // odd stride: rows of different directions start in different banks
    if (P->kstr > 66 || Mp * P->kstr * 8 > 32768 || Mp * P->kstr2 * 8 > 32768) return 0;
    const int G = (int)((N + 31) / 32);
    if (G > 64) return 0;
    // 16 warps x 4 slots; 8 warps x 8 slots (255 registers per thread, more independent FMA chains per warp) and
    // 32 warps x 2 slots (64 registers, spills) were measured slower at N = 2000: 67 / 72 ms against 55.6 ms
    P->slots = 4;
    if (const char* e = getenv("ASGFEM_TS2_SLOTS")) {
        int v = atoi(e);
        if (v == 4 || v == 8) P->slots = v;
    }
    const int TS2_SLOTS = P->slots;
    const int maxW = TS2_SLOTS == 8 ? 8 : 16;
    const int S_used = (G + maxW - 1) / maxW, W = (G + S_used - 1) / S_used;
    P->warps = W;

    // ---- directions per mode ---------------------------------------------------------------------------
    std::vector<unsigned long long> actmode((size_t)N, 0ull);
    for (int64_t mu = 0; mu < N; ++mu)
        for (int32_t e = C.ptr[mu]; e < C.ptr[mu + 1]; ++e) {
            if (C.m[e] < 1 || C.m[e] > M) return 0;
            actmode[(size_t)C.nu[e]] |= 1ull << C.m[e];
        }
    std::vector<unsigned long long> gunion((size_t)G, 0ull);
    std::vector<int> gmaxd((size_t)G, 0);
    for (int64_t nu = 0; nu < N; ++nu) {
        gunion[(size_t)(nu / 32)] |= actmode[(size_t)nu];
        gmaxd[(size_t)(nu / 32)] = std::max(gmaxd[(size_t)(nu / 32)], __builtin_popcountll(actmode[(size_t)nu]));
    }
    auto nd4 = [&](int g) { return (__builtin_popcountll(gunion[(size_t)g]) + 3) / 4 * 4; };
    // Q = 3 or 4: whichever evaluates fewer units (a group is sparse if no lane needs more than Q directions and the
    // Q per-lane units are not more than the dense units of the group)
    auto total_units = [&](int Q) {
        int64_t t = 0;
        for (int g = 0; g < G; ++g) t += (gmaxd[(size_t)g] <= Q && Q < nd4(g)) ? Q : nd4(g);
        return t;
    };
    P->Q = total_units(4) < total_units(3) ? 4 : 3;
    if (const char* e = getenv("ASGFEM_TS2_Q")) {
        int v = atoi(e);
        if (v == 3 || v == 4) P->Q = v;
    }
    const int Q = P->Q;
    std::vector<uint8_t> sparse((size_t)G, 0);
    for (int g = 0; g < G; ++g) sparse[(size_t)g] = gmaxd[(size_t)g] <= Q && Q < nd4(g);
    P->units = (int)total_units(Q) + G;

    // ---- units -> exchange buffer (first fit on the 32-bit lane masks, fullest first) -------------------
    struct Unit {
        int g, key;  // key = direction (dense) or rank q (sparse)
        uint32_t mask;
        int32_t disp;
    };
    std::vector<Unit> units;
    std::vector<int32_t> uidx((size_t)G * 64, -1);  // (g, key) -> unit
    for (int g = 0; g < G; ++g) {
        if (!sparse[(size_t)g]) {
            for (int m = 1; m <= M; ++m)
                if (gunion[(size_t)g] >> m & 1ull) {
                    uidx[(size_t)g * 64 + m] = (int32_t)units.size();
                    units.push_back({g, m, 0xffffffffu, 0});
                }
        } else {
            for (int q = 0; q < Q; ++q) {
                uint32_t mk = 0;
                for (int l = 0; l < 32; ++l) {
                    const int64_t nu = 32ll * g + l;
                    if (nu < N && __builtin_popcountll(actmode[(size_t)nu]) > q) mk |= 1u << l;
                }
                if (!mk) continue;
                uidx[(size_t)g * 64 + q] = (int32_t)units.size();
                units.push_back({g, q, mk, 0});
            }
        }
    }
    int D = 0;
    {
        std::vector<int> ord(units.size());
        std::iota(ord.begin(), ord.end(), 0);
        std::stable_sort(ord.begin(), ord.end(),
                         [&](int a, int b) { return __builtin_popcount(units[a].mask) > __builtin_popcount(units[b].mask); });
        std::vector<uint8_t> occ;
        size_t low = 0;  // everything below is occupied
        for (int k : ord) {
            const uint32_t mk = units[k].mask;
            const int first = __builtin_ctz(mk);
            size_t d = low > (size_t)first ? low - first : 0;
            for (;; ++d) {
                if (occ.size() < d + 32) occ.resize(d + 32 + 1024, 0);
                bool clash = false;
                for (int l = first; l < 32 && !clash; ++l) clash = (mk >> l & 1u) && occ[d + l];
                if (!clash) break;
            }
            for (int l = first; l < 32; ++l)
                if (mk >> l & 1u) {
                    occ[d + l] = 1;
                    D = std::max<int>(D, (int)(d + l + 1));
                }
            while (low < occ.size() && occ[low]) ++low;
            units[k].disp = (int32_t)d;
        }
    }
    if (D + TS2_DUMMY >= (1 << 17)) return 0;
    P->D = D;
    auto rank_of = [&](int64_t nu, int m) {  // position of m among the directions of nu
        return __builtin_popcountll(actmode[(size_t)nu] & ((1ull << m) - 1ull));
    };
    auto pairidx = [&](int64_t nu, int m) -> int32_t {
        const int g = (int)(nu / 32);
        const int32_t u = uidx[(size_t)g * 64 + (sparse[(size_t)g] ? rank_of(nu, m) : m)];
        if (sparse[(size_t)g]) return units[(size_t)u].disp + (int32_t)(nu % 32);
        // dense blocks are rotated by the position of the unit in the list of the group: the entries of one mode for
        // consecutive directions (gathered by consecutive lanes of a sparse group) then lie in different banks
        const int k = __builtin_popcountll(gunion[(size_t)g] & ((1ull << m) - 1ull));
        return units[(size_t)u].disp + (int32_t)((nu % 32 + k) & 31);
    };

    // ---- weight table: distinct coupling coefficients -------------------------------------------------
    std::vector<double> gtab(64, 0.0);
    int ng = 1;  // entry 0 = 0.0 (dummy)
    auto gindex = [&](double g) {
        for (int k = 1; k < ng; ++k)
            if (gtab[k] == g) return k;
        if (ng >= 64) return -1;
        gtab[ng] = g;
        return ng++;
    };

    // ---- groups -> (warp, slot): longest processing time first, then warps dealt to the four schedulers ---------
    // Phase 2 of a listed group (a long gather list) does not need the X registers of the group: it may run in another
    // warp.  The owner then exports the mean term through the exchange buffer (one more entry of the gather list with
    // weight 1) and a warp with little phase-1 work takes the list as its extra, gather-only slot.
    std::vector<int> jlen((size_t)G, 0);
    for (int g = 0; g < G; ++g)
        for (int l = 0; l < 32; ++l) {
            int64_t mu = 32ll * g + l;
            if (mu < N) jlen[(size_t)g] = std::max(jlen[(size_t)g], (int)(C.ptr[mu + 1] - C.ptr[mu]));
        }
    auto even = [](int v) { return v; };  // list lengths need no padding: the gather loop has tails of 2 and 1
    // relative cost of a unit and of a gather slot for the load balance of the warps
    int64_t CU = 16, CG = 9;  // measured best of (16,9), (23,17), (20,20), (30,12): 51.2 / 51.6 / 53.3 / 51.5 ms
    if (const char* e = getenv("ASGFEM_TS2_COST")) sscanf(e, "%lld,%lld", (long long*)&CU, (long long*)&CG);
    auto cost1 = [&](int g) -> int64_t {  // phase 1 (+ phase 2 for sparse groups, which stay with their owner)
        return sparse[(size_t)g] ? CU * Q + CG * even(jlen[(size_t)g]) + 20 : CU * nd4(g) + 20;
    };
    auto cost2 = [&](int g) -> int64_t { return CG * even(jlen[(size_t)g] + 1) + 10; };
    std::vector<int> order((size_t)G);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost1(a) > cost1(b); });
    std::vector<int64_t> load((size_t)W, 0);
    std::vector<std::vector<int>> wg((size_t)W);
    std::vector<int> owner((size_t)G, 0);
    for (int g : order) {
        int best = -1;
        for (int w = 0; w < W; ++w)
            if ((int)wg[(size_t)w].size() < S_used && (best < 0 || load[(size_t)w] < load[(size_t)best])) best = w;
        wg[(size_t)best].push_back(g);
        owner[(size_t)g] = best;
        load[(size_t)best] += cost1(g);
    }
    std::vector<int> extra((size_t)W, -1);     // logical warp -> listed group whose phase 2 it runs
    std::vector<int> exported((size_t)G, -1);  // listed group -> index of its mean-term block, -1: phase 2 stays with the owner
    int nexp = 0;
    {
        std::vector<int> lst;
        for (int g = 0; g < G; ++g)
            if (!sparse[(size_t)g]) lst.push_back(g);
        std::stable_sort(lst.begin(), lst.end(), [&](int a, int b) { return cost2(a) > cost2(b); });
        const bool allow = getenv("ASGFEM_TS2_NOEXPORT") == nullptr;
        for (int g : lst) {
            int best = -1;
            for (int w = 0; w < W; ++w)
                if (extra[(size_t)w] < 0 && (best < 0 || load[(size_t)w] < load[(size_t)best])) best = w;
            if (!allow || best < 0 || best == owner[(size_t)g] ||
                load[(size_t)best] + cost2(g) >= load[(size_t)owner[(size_t)g]] + cost2(g) - CU) {
                load[(size_t)owner[(size_t)g]] += cost2(g) - CG;  // no export: the list is one entry shorter
                continue;
            }
            extra[(size_t)best] = g;
            exported[(size_t)g] = nexp++;
            load[(size_t)best] += cost2(g);
            load[(size_t)owner[(size_t)g]] += 2;
        }
    }
    const int D0 = D;  // mean-term blocks of the exported groups behind the unit blocks
    D = D0 + 32 * nexp;
    if (D + TS2_DUMMY >= (1 << 17)) return 0;
    P->D = D;
    std::vector<int> wrank((size_t)W);
    std::iota(wrank.begin(), wrank.end(), 0);
    std::stable_sort(wrank.begin(), wrank.end(), [&](int a, int b) { return load[(size_t)a] > load[(size_t)b]; });
    std::vector<int> phys((size_t)W, -1);  // logical warp (by rank) -> physical warp id: snake over warp_id % 4
    {
        std::vector<int> freeids;
        for (int r = 0; r < W; ++r) {
            const int j = r / 4, c = r % 4;
            const int smsp = (j & 1) ? 3 - c : c;
            int id = 4 * j + smsp;
            if (id >= W) id = -1;
            phys[(size_t)r] = id;
        }
        std::vector<uint8_t> taken((size_t)W, 0);
        for (int r = 0; r < W; ++r)
            if (phys[(size_t)r] >= 0) taken[(size_t)phys[(size_t)r]] = 1;
        for (int id = 0; id < W; ++id)
            if (!taken[(size_t)id]) freeids.push_back(id);
        for (int r = 0; r < W; ++r)
            if (phys[(size_t)r] < 0) {
                phys[(size_t)r] = freeids.back();
                freeids.pop_back();
            }
    }

    // ---- phase-2 lists per group: words[wbase + 32*j + lane] ------------------------------------------
    std::vector<int32_t> jmax((size_t)G, 0), wbase((size_t)G, 0);
    std::vector<uint32_t> words;
    int64_t confl_before = 0, confl_after = 0;
    for (int g = 0; g < G; ++g) {
        const bool exp = exported[(size_t)g] >= 0;
        const int jm = jlen[(size_t)g] + (exp ? 1 : 0);  // the gather loop is unrolled by 4 with tails of 2 and 1
        jmax[(size_t)g] = jm;
        wbase[(size_t)g] = (int32_t)words.size();
        words.resize(words.size() + (size_t)jm * 32, 0u);
        uint32_t* wl = words.data() + wbase[(size_t)g];
        for (int l = 0; l < 32; ++l) {
            // padding: weight index 0 -> 0.0 and a zero entry of the exchange buffer in the lane's own bank
            for (int j = 0; j < jm; ++j) wl[(size_t)j * 32 + l] = (uint32_t)(D + (l & 15)) << 12;
            int64_t mu = 32ll * g + l;
            if (mu >= N) continue;
            int j = 0;
            for (int32_t e = C.ptr[mu]; e < C.ptr[mu + 1]; ++e, ++j) {
                int gi = gindex(C.g[e]);
                if (gi < 0) return 0;
                wl[(size_t)j * 32 + l] = ((uint32_t)pairidx(C.nu[e], C.m[e]) << 12) | ((uint32_t)gi << 3);
            }
            if (exp) {  // mean term of the mode, written by the owner of the group
                int gi = gindex(1.0);
                if (gi < 0) return 0;
                wl[(size_t)j * 32 + l] = ((uint32_t)(D0 + 32 * exported[(size_t)g] + l) << 12) | ((uint32_t)gi << 3);
            }
        }
        // The order of a lane's couplings is free: permute every lane's list so that the 16 lanes of a half-warp read
        // 16 different banks (8-byte granules mod 16) in as many slots as possible (local search on pairwise swaps).
        for (int h = 0; h < 2; ++h) {
            std::vector<int> cnt((size_t)jm * 16, 0);
            auto bank = [&](int l, int j) { return (int)((wl[(size_t)j * 32 + l] >> 12) & 15u); };
            for (int j = 0; j < jm; ++j)
                for (int l = 16 * h; l < 16 * h + 16; ++l) ++cnt[(size_t)j * 16 + bank(l, j)];
            auto excess = [&]() {
                int64_t x = 0;
                for (int c : cnt) x += c > 1 ? c - 1 : 0;
                return x;
            };
            confl_before += excess();
            for (int pass = 0; pass < 40; ++pass) {
                bool improved = false;
                for (int l = 16 * h; l < 16 * h + 16; ++l)
                    for (int j1 = 0; j1 < jm; ++j1) {
                        const int b1 = bank(l, j1);
                        if (cnt[(size_t)j1 * 16 + b1] < 2) continue;  // not in conflict
                        int best = -1, bestgain = 0;
                        for (int j2 = 0; j2 < jm; ++j2) {
                            if (j2 == j1) continue;
                            const int b2 = bank(l, j2);
                            if (b1 == b2) continue;
                            // excess change: slot j1 loses b1 gains b2, slot j2 loses b2 gains b1
                            int gain = 0;
                            gain += cnt[(size_t)j1 * 16 + b1] > 1 ? 1 : 0;
                            gain -= cnt[(size_t)j1 * 16 + b2] > 0 ? 1 : 0;
                            gain += cnt[(size_t)j2 * 16 + b2] > 1 ? 1 : 0;
                            gain -= cnt[(size_t)j2 * 16 + b1] > 0 ? 1 : 0;
                            if (gain > bestgain) bestgain = gain, best = j2;
                        }
                        if (best >= 0) {
                            const int b2 = bank(l, best);
                            --cnt[(size_t)j1 * 16 + b1], ++cnt[(size_t)j1 * 16 + b2];
                            --cnt[(size_t)best * 16 + b2], ++cnt[(size_t)best * 16 + b1];
                            std::swap(wl[(size_t)j1 * 32 + l], wl[(size_t)best * 32 + l]);
                            improved = true;
                        }
                    }
                if (!improved) break;
            }
            confl_after += excess();
        }
    }
    P->nwords = (int)words.size();

    const int SI = TS2_SLOTS + 1;  // slot records per warp: the X slots + one gather-only slot
    std::vector<int32_t> slotinfo((size_t)W * SI * 8, 0);
    for (size_t k = 0; k < slotinfo.size(); k += 8) slotinfo[k] = -1;
    std::vector<uint32_t> lane((size_t)W * TS2_SLOTS * 4 * 32, 0u);
    std::vector<uint32_t> dl;
    const uint32_t kbytes = (uint32_t)P->kstr * 8u, kbytes2 = (uint32_t)P->kstr2 * 8u;
    const uint32_t dummy_lane = (uint32_t)(D + 16) << 15;  // K_0 row, dummy T block
    for (int r = 0; r < W; ++r) {
        const int lw = wrank[(size_t)r], w = phys[(size_t)r];
        for (int s = 0; s < (int)wg[(size_t)lw].size(); ++s) {
            const int g = wg[(size_t)lw][(size_t)s];
            int32_t* si = &slotinfo[((size_t)w * SI + s) * 8];
            si[0] = g;
            si[4] = wbase[(size_t)g];
            si[5] = jmax[(size_t)g];
            if (!sparse[(size_t)g]) {
                si[1] = exported[(size_t)g] >= 0 ? 3 : 1;  // 3: phase 2 runs elsewhere, the mean term is exported
                si[6] = exported[(size_t)g] >= 0 ? 8 * (D0 + 32 * exported[(size_t)g]) : 0;
                si[3] = (int32_t)dl.size();
                int cnt = 0;
                for (int m = 1; m <= M; ++m) {
                    const int32_t u = uidx[(size_t)g * 64 + m];
                    if (u < 0) continue;
                    dl.push_back((uint32_t)m * kbytes | (uint32_t)units[(size_t)u].disp << 15);
                    ++cnt;
                }
                for (; cnt % 4; ++cnt) dl.push_back(dummy_lane);  // lanes add 8 * lane: the dummy block has 32 entries
                si[2] = cnt;
            } else {
                si[1] = 2;
                for (int l = 0; l < 32; ++l) {
                    const int64_t nu = 32ll * g + l;
                    int q = 0;
                    if (nu < N)
                        for (int m = 1; m <= M; ++m)
                            if (actmode[(size_t)nu] >> m & 1ull) {
                                const int32_t u = uidx[(size_t)g * 64 + q];
                                lane[(((size_t)w * TS2_SLOTS + s) * 4 + q) * 32 + l] =
                                    (uint32_t)m * kbytes2 | (uint32_t)(units[(size_t)u].disp + l) << 15;
                                ++q;
                            }
                    // padding units repeat the lane's first unit (same K row, same T entry: the same value is stored again by
                    // the same thread); a lane without any unit - or any lane when long rows accumulate T in chunks -
                    // computes on K_0 and stores to its own dummy entry
                    const uint32_t fill = q > 0 && P->nchunk_max == 1 ? lane[(((size_t)w * TS2_SLOTS + s) * 4 + 0) * 32 + l] : (uint32_t)(D + 16 + l) << 15;
                    for (; q < 4; ++q) lane[(((size_t)w * TS2_SLOTS + s) * 4 + q) * 32 + l] = fill;
                }
            }
        }
    }
    for (int r = 0; r < W; ++r) {  // gather-only slots
        const int lw = wrank[(size_t)r], w = phys[(size_t)r];
        int32_t* si = &slotinfo[((size_t)w * SI + TS2_SLOTS) * 8];
        const int g = extra[(size_t)lw];
        si[0] = g;
        if (g >= 0) si[4] = wbase[(size_t)g], si[5] = jmax[(size_t)g];
    }
    P->ndl = (int)dl.size();

    const size_t limit = 226 * 1024;  // 512 bytes of static shared memory (weight table) + the metadata rings come on top
    P->nbuf = 2;
    P->smem_bytes = ts2_layout(P->D, 2, Mp, P->kstr, P->kstr2, P->nwords, P->ndl, W, TS2_SLOTS).total;
    if (P->smem_bytes > limit) {
        P->nbuf = 1;
        P->smem_bytes = ts2_layout(P->D, 1, Mp, P->kstr, P->kstr2, P->nwords, P->ndl, W, TS2_SLOTS).total;
        if (P->smem_bytes > limit) return 0;
    }
    int rc = 0;
    rc |= dev_upload(ctx, &P->d_slotinfo, slotinfo);
    rc |= dev_upload(ctx, &P->d_lane, lane);
    rc |= dev_upload(ctx, &P->d_dl, dl);
    rc |= dev_upload(ctx, &P->d_words, words);
    rc |= dev_upload(ctx, &P->d_gtab, gtab);
    if (rc) return rc;
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (const char* e = getenv("ASGFEM_TS2_GRID")) P->grid_sms = std::max(1, std::min(148, atoi(e)));
    P->usable = true;
    if (getenv("ASGFEM_TS2_VERBOSE")) {
        int64_t lmin = 1 << 30, lmax = 0;
        for (int64_t v : load) lmin = std::min(lmin, v), lmax = std::max(lmax, v);
        fprintf(stderr,
                "[ts2] N=%lld G=%d warps=%d slots=%d Q=%d units/row=%d D=%d words=%d smem=%zu nbuf=%d NS=%d chunks=%d "
                "gather bank excess %lld -> %lld, exported phase-2 lists %d, warp cost %lld..%lld\n",
                (long long)N, G, W, TS2_SLOTS, Q, P->units, P->D, P->nwords, P->smem_bytes, P->nbuf, P->NS, P->nchunk_max,
                (long long)confl_before, (long long)confl_after, nexp, (long long)lmin, (long long)lmax);
    }
    return 0;
}

bool apply_ts2_preferred(asgfem_ctx* ctx) {
    if (!tp2_of(ctx) && apply_ts2_build(ctx)) return false;
    Ts2Plan* P = tp2_of(ctx);
    return P && P->usable && P->nchunk_max == 1;
}

struct Ts2Args {
    int64_t row0, nrows, ld, nnz;  // rows [row0, nrows)
    int N, M, Mp, D, nbuf, kstr, kstr2, nwords, ndl, warps;
    const int64_t* rowptr;
    const int32_t* col;
    const double* vals;
    const uint8_t* bmask;
    const int32_t* slotinfo;
    const uint32_t* lane;
    const uint32_t* dl;
    const uint32_t* words;
    const double* gtab;
    const double* x;
    double* y;
};

namespace {
__device__ __forceinline__ double t2_ldg_f64(const double* p) {
    double v;
    // (L1::no_allocate was measured slower: 57.9 ms against 55.0 ms)
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double t2_lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ double2 t2_lds_f64x2(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned t2_lds_u32(unsigned addr) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 t2_lds_u32x4(unsigned addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void t2_sts_f64(unsigned addr, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}

// NA + NB units at once: T = sum_k K[k] x[k] with the K row at kbase + (w & 0x7fff), stored at tbase + 8 * (w >> 15);
// the first NA units belong to the slot with operands xa, the others to xb.  NA + NB independent FMA chains, all K loads
// of the batch are issued before the first FMA needs them; K is read in 16-byte pairs (the pad entry of a row is zero).
template <int NS, int NA, int NB, bool MULTI, bool WIDE>
__device__ __forceinline__ void t2_units(const unsigned* wa, const double (&xa)[NS], unsigned tba, const unsigned* wb,
                                         const double (&xb)[NS], unsigned tbb, unsigned kbase, bool first) {
    constexpr int NK2 = (NS + 1) / 2, NU = NA + NB;
    unsigned ka[NU], w[NU];
    double t[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        w[u] = u < NA ? wa[u] : wb[u - NA];
        ka[u] = kbase + (w[u] & 0x7fffu);
    }
    if (WIDE) {
#pragma unroll
        for (int kk = 0; kk < NK2; ++kk) {
#pragma unroll
            for (int u = 0; u < NU; ++u) {
                const double2 v = t2_lds_f64x2(ka[u] + 16u * (unsigned)kk);
                const double x0 = u < NA ? xa[2 * kk] : xb[2 * kk];
                t[u] = kk == 0 ? v.x * x0 : fma(v.x, x0, t[u]);
                if (2 * kk + 1 < NS) t[u] = fma(v.y, u < NA ? xa[2 * kk + 1] : xb[2 * kk + 1], t[u]);
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < NS; ++k) {
#pragma unroll
            for (int u = 0; u < NU; ++u) {
                const double v = t2_lds_f64(ka[u] + 8u * (unsigned)k);
                const double xk = u < NA ? xa[k] : xb[k];
                t[u] = k == 0 ? v * xk : fma(v, xk, t[u]);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        // dense units (WIDE): tba = start of the exchange buffer, tbb = lane + position of the first unit in the list of
        // the group; the 32-entry block of unit u is rotated by its position
        const unsigned addr = WIDE ? tba + 8u * ((tbb + (unsigned)u) & 31u) + ((w[u] >> 12) & 0xfffffff8u)
                                   : (u < NA ? tba : tbb) + ((w[u] >> 12) & 0xfffffff8u);
        if (!MULTI || first)
            t2_sts_f64(addr, t[u]);
        else
            t2_sts_f64(addr, t2_lds_f64(addr) + t[u]);
    }
}
}  // namespace

// S slots per warp; MULTI: rows longer than NS columns exist (processed in chunks, T accumulates in the exchange buffer)
template <int NS, int Q, int S, bool MULTI>
__global__ void __launch_bounds__(S == 8 ? 256 : 512, 1) k_apply_ts2(Ts2Args a) {
    constexpr int SLOTS = S;
    extern __shared__ __align__(16) unsigned char ts2_raw[];
    __shared__ __align__(512) double gt[64];
    // CSR metadata of the CTA's upcoming rows, fetched by warp 0 with cp.async two / three rows ahead (see apply_ts.cu)
    __shared__ __align__(16) long long info_rp[4][2];  // [row j % 4]: rowptr[row], rowptr[row + 1]
    __shared__ int info_msk[4];                         // Dirichlet flag (1 also for rows past the end)
    __shared__ int cols_ring[2][NS];                    // [row j % 2]: first NS column indices
    const unsigned raw32 = (unsigned)__cvta_generic_to_shared(ts2_raw);
    const Ts2Layout L = ts2_layout(a.D, a.nbuf, a.Mp, a.kstr, a.kstr2, a.nwords, a.ndl, a.warps, SLOTS);
    const int Dpad = (int)L.dpad;
    double* Ts = reinterpret_cast<double*>(ts2_raw + L.ts);            // [nbuf][Dpad]
    double* Ks = reinterpret_cast<double*>(ts2_raw + L.ks);            // [2][Mp][kstr]
    double* Ks2 = reinterpret_cast<double*>(ts2_raw + L.ks2);          // [2][Mp][kstr2]
    uint32_t* words = reinterpret_cast<uint32_t*>(ts2_raw + L.words);  // [nwords]
    uint32_t* dls = reinterpret_cast<uint32_t*>(ts2_raw + L.dl);       // [ndl]
    int32_t* sinfo = reinterpret_cast<int32_t*>(ts2_raw + L.sinfo);    // [warps][SLOTS + 1][8]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
    const unsigned gt32 = (unsigned)__cvta_generic_to_shared(gt), ts32 = raw32 + L.ts, words32 = raw32 + L.words,
                   ks32 = raw32 + L.ks, ks2_32 = raw32 + L.ks2, dl32 = raw32 + L.dl;

    for (int k = tid; k < 64; k += nthr) gt[k] = a.gtab[k];
    for (int k = tid; k < a.nwords; k += nthr) words[k] = a.words[k];
    for (int k = tid; k < a.ndl; k += nthr) dls[k] = a.dl[k];
    for (int k = tid; k < a.warps * (SLOTS + 1) * 8; k += nthr) sinfo[k] = a.slotinfo[k];
    for (int k = tid; k < a.nbuf * Dpad; k += nthr) Ts[k] = 0.0;  // includes the zero entry D of the padded gather lists
    unsigned li[SLOTS][Q];   // sparse slots: per-lane unit constants
    int moff[SLOTS];         // own mode of the slot (clamped to a valid mode for idle lanes / unused slots: loads stay in
    unsigned validmask = 0;  // bounds and unpredicated, nothing is stored for them)
    unsigned kindmask = 0;   // 2 bits per slot: 0 unused, 1 dense, 2 sparse, 3 dense with exported phase 2
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        const int32_t* si = a.slotinfo + ((size_t)warp * (SLOTS + 1) + s) * 8;
        const int g = si[0];
        const bool valid = g >= 0 && 32 * g + lane < a.N;
        moff[s] = valid ? 32 * g + lane : (g >= 0 ? a.N - 1 : 0);
        validmask |= valid ? 1u << s : 0u;
        kindmask |= (g >= 0 ? (unsigned)si[1] : 0u) << (2 * s);
#pragma unroll
        for (int q = 0; q < Q; ++q) li[s][q] = a.lane[(((size_t)warp * SLOTS + s) * 4 + q) * 32 + lane];
    }
    const int32_t* myinfo = sinfo + warp * (SLOTS + 1) * 8;
    const int xgroup = a.slotinfo[((size_t)warp * (SLOTS + 1) + SLOTS) * 8];  // gather-only slot: group or -1

    double x[SLOTS][NS];
    auto stage_k = [&](int64_t rp, int len, int buf) {
        double* dst = Ks + (size_t)buf * a.Mp * a.kstr;
        double* dst2 = Ks2 + (size_t)buf * a.Mp * a.kstr2;
        for (int idx = tid; idx < a.Mp * a.kstr; idx += nthr) {
            const int m = idx / a.kstr, k = idx - m * a.kstr;
            const double v = k < len ? __ldg(a.vals + (int64_t)m * a.nnz + rp + k) : 0.0;
            dst[idx] = v;
            if (k < a.kstr2) dst2[m * a.kstr2 + k] = v;
        }
    };
    const unsigned long long xbase = (unsigned long long)a.x;
    const unsigned ldb = (unsigned)a.ld * 8u;  // row pitch in bytes (< 4 GB)
    auto load_x = [&](int64_t rp, int len, int c0, const int* cols_s) {
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            if (c0 + k < len) {  // block-uniform
                const unsigned cj = cols_s ? (unsigned)cols_s[k] : (unsigned)__ldg(a.col + rp + c0 + k);
                const unsigned long long rowp = xbase + (unsigned long long)cj * ldb;
#pragma unroll
                for (int s = 0; s < SLOTS; ++s)
                    x[s][k] = t2_ldg_f64(reinterpret_cast<const double*>(rowp + (unsigned)(8 * moff[s])));
            } else {
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) x[s][k] = 0.0;
            }
        }
    };

    int64_t row = a.row0 + blockIdx.x;
    int64_t rp = 0;
    int len = 0;
    bool masked = true;
    const int64_t G = gridDim.x;
    // metadata loader (warp 0): threads 0..NS-1 fetch column ids, thread NS the row pointers, thread NS+1 the Dirichlet flags
    auto fetch_info = [&](int64_t j) {  // row j of this CTA -> info ring (cp.async, completes at the next wait)
        const int64_t r = a.row0 + blockIdx.x + j * G;
        if (tid == NS) {
            if (r < a.nrows) {
                const unsigned dst = (unsigned)__cvta_generic_to_shared(&info_rp[j & 3][0]);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(a.rowptr + r));
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8u), "l"(a.rowptr + r + 1));
            } else {
                info_rp[j & 3][0] = 0;
                info_rp[j & 3][1] = 0;
            }
        }
    };
    auto fetch_cols = [&](int64_t j) {  // needs info of row j complete
        if (tid < NS) {
            const long long p0 = info_rp[j & 3][0], p1 = info_rp[j & 3][1];
            if (tid < (int)(p1 - p0)) {
                const unsigned dst = (unsigned)__cvta_generic_to_shared(&cols_ring[j & 1][tid]);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(a.col + p0 + tid));
            }
        }
    };
    auto load_msk = [&](int64_t j) -> unsigned {  // Dirichlet flag of row j of this CTA (1 also for rows past the end);
        const int64_t r = a.row0 + blockIdx.x + j * G;  // volatile: issued here, consumed an iteration later
        unsigned v = 1u;
        if (tid == NS + 1 && r < a.nrows) asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(a.bmask + r));
        return v;
    };
    auto cp_wait_all = [&]() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); };
    if (row < a.nrows) {
        rp = a.rowptr[row];
        len = (int)(a.rowptr[row + 1] - rp);
        masked = a.bmask[row] != 0 || len == 0;
        if (!masked) {
            stage_k(rp, len, 0);
            load_x(rp, len, 0, nullptr);
        }
    }
    if (warp == 0) {
        fetch_info(1);
        fetch_info(2);
        const unsigned m1 = load_msk(1), m2 = load_msk(2);
        if (tid == NS + 1) info_msk[1] = (int)m1, info_msk[2] = (int)m2;
    }
    cp_wait_all();
    __syncthreads();
    if (warp == 0) fetch_cols(1);
    cp_wait_all();
    __syncthreads();
    int buf = 0;
    for (int64_t it = 0; row < a.nrows; row += gridDim.x, buf ^= 1, ++it) {
        // loader (warp 0): column ids of row it+2 (its row pointers arrived one iteration ago), metadata of row it+3;
        // issued first, so that the copies complete behind phase 1
        unsigned msk3 = 0u;
        if (warp == 0) {
            fetch_cols(it + 2);
            fetch_info(it + 3);
            msk3 = load_msk(it + 3);
        }
        // K values of the next row: loaded now (HBM latency hidden behind phase 1), stored to shared memory after it
        const int64_t nrow = row + gridDim.x;
        int64_t nrp = 0;
        int nlen = 0;
        bool nmasked = true;
        if (nrow < a.nrows) {
            const int js = (int)((it + 1) & 3);
            nrp = info_rp[js][0];
            nlen = (int)(info_rp[js][1] - nrp);
            nmasked = info_msk[js] != 0 || nlen == 0;
        }
        // K values of the next row: asynchronous copies into the other buffer (both layouts) while phase 1 runs on this
        // one; columns past the end of the row are zero-filled (source size 0)
        const int kper = (a.Mp * a.kstr + nthr - 1) / nthr;  // K values per thread (1 for P1: 210 values)
        if (!nmasked && kper == 1 && tid < a.Mp * a.kstr) {
            const int m = tid / a.kstr, k = tid - m * a.kstr;
            const double* src = a.vals + (int64_t)m * a.nnz + nrp + (k < nlen ? k : 0);
            const unsigned sz = k < nlen ? 8u : 0u;
            const unsigned d1 = ks32 + (unsigned)((buf ^ 1) * a.Mp * a.kstr + tid) * 8u;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d1), "l"(src), "r"(sz));
            if (k < a.kstr2) {
                const unsigned d2 = ks2_32 + (unsigned)((buf ^ 1) * a.Mp * a.kstr2 + m * a.kstr2 + k) * 8u;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d2), "l"(src), "r"(sz));
            }
        }
        const int tb = a.nbuf == 2 ? buf : 0;
        const unsigned T32 = ts32 + (unsigned)tb * (unsigned)Dpad * 8u;
        const unsigned Tl32 = T32 + 8u * (unsigned)lane;
        const unsigned K32 = ks32 + (unsigned)buf * (unsigned)(a.Mp * a.kstr) * 8u;
        const unsigned K32s = ks2_32 + (unsigned)buf * (unsigned)(a.Mp * a.kstr2) * 8u;
        double acc[SLOTS];
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) acc[s] = 0.0;
        // ---------------- phase 1: T for the units of the own slots, all FMA operands in registers -----------
        if (!masked) {
            for (int c0 = 0; c0 < (MULTI ? len : 1); c0 += NS) {
                if (MULTI && c0 > 0) load_x(rp, len, c0, nullptr);
                const unsigned Kc = K32 + 8u * (unsigned)c0, Kcs = K32s + 8u * (unsigned)c0;
                const bool first = c0 == 0;
                {  // mean term: K_0 row broadcast, one chain per slot
                    constexpr int NK2 = (NS + 1) / 2;
#pragma unroll
                    for (int kk = 0; kk < NK2; ++kk) {
                        const double2 v = t2_lds_f64x2(Kc + 16u * (unsigned)kk);
#pragma unroll
                        for (int s = 0; s < SLOTS; ++s) {
                            acc[s] = fma(v.x, x[s][2 * kk], acc[s]);
                            if (2 * kk + 1 < NS) acc[s] = fma(v.y, x[s][2 * kk + 1], acc[s]);
                        }
                    }
                }
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    const unsigned kind = (kindmask >> (2 * s)) & 3u;
                    if (kind & 1u) {  // dense: unit list in shared memory, four units per step
                        const int nd = myinfo[8 * s + 2];
                        unsigned p = dl32 + 4u * (unsigned)myinfo[8 * s + 3];
                        for (int u = 0; u < nd; u += 4, p += 16u) {
                            const uint4 ww = t2_lds_u32x4(p);
                            const unsigned w4[4] = {ww.x, ww.y, ww.z, ww.w};
                            t2_units<NS, 4, 0, MULTI, true>(w4, x[s], T32, w4, x[s], (unsigned)(lane + u), Kc, first);
                        }
                    }
                }
                // sparse: Q per-lane units per slot, two slots (2 Q chains) at a time
#pragma unroll
                for (int s = 0; s < SLOTS; s += 2) {
                    const unsigned k0 = (kindmask >> (2 * s)) & 3u, k1 = (kindmask >> (2 * s + 2)) & 3u;
                    if (k0 == 2u && k1 == 2u) {
                        t2_units<NS, Q, Q, MULTI, false>(li[s], x[s], T32, li[s + 1], x[s + 1], T32, Kcs, first);
                    } else {
                        if (k0 == 2u) t2_units<NS, Q, 0, MULTI, false>(li[s], x[s], T32, li[s], x[s], T32, Kcs, first);
                        if (k1 == 2u) t2_units<NS, Q, 0, MULTI, false>(li[s + 1], x[s + 1], T32, li[s + 1], x[s + 1], T32, Kcs, first);
                    }
                }
            }
#pragma unroll
            for (int s = 0; s < SLOTS; ++s)  // mean term of the groups whose phase 2 runs in another warp
                if (((kindmask >> (2 * s)) & 3u) == 3u) t2_sts_f64(Tl32 + (unsigned)myinfo[8 * s + 6], acc[s]);
        }
        // ---------------- next row: K values into the other buffer, X loads in flight during phase 2 --------
        if (!nmasked) {
            if (kper != 1) stage_k(nrp, nlen, buf ^ 1);
            load_x(nrp, nlen, 0, cols_ring[(it + 1) & 1]);
        }
        if (tid == NS + 1) info_msk[(it + 3) & 3] = (int)msk3;
        cp_wait_all();
        __syncthreads();
        // ---------------- phase 2: gather the couplings that end in the own modes ---------------------------
        double* yr = a.y + row * a.ld;
        auto gather = [&](int jm, unsigned wp, double r0) -> double {
            double r1 = 0.0;
            for (int j = 0; j + 4 <= jm; j += 4, wp += 512u) {
                const unsigned w0 = t2_lds_u32(wp), w1 = t2_lds_u32(wp + 128u), w2 = t2_lds_u32(wp + 256u),
                               w3 = t2_lds_u32(wp + 384u);
                const double g0 = t2_lds_f64(gt32 + (w0 & 0x1f8u)), t0 = t2_lds_f64(T32 + (w0 >> 9));
                const double g1 = t2_lds_f64(gt32 + (w1 & 0x1f8u)), t1 = t2_lds_f64(T32 + (w1 >> 9));
                const double g2 = t2_lds_f64(gt32 + (w2 & 0x1f8u)), t2 = t2_lds_f64(T32 + (w2 >> 9));
                const double g3 = t2_lds_f64(gt32 + (w3 & 0x1f8u)), t3 = t2_lds_f64(T32 + (w3 >> 9));
                r0 = fma(g0, t0, r0);
                r1 = fma(g1, t1, r1);
                r0 = fma(g2, t2, r0);
                r1 = fma(g3, t3, r1);
            }
            if (jm & 2) {
                const unsigned w0 = t2_lds_u32(wp), w1 = t2_lds_u32(wp + 128u);
                const double g0 = t2_lds_f64(gt32 + (w0 & 0x1f8u)), t0 = t2_lds_f64(T32 + (w0 >> 9));
                const double g1 = t2_lds_f64(gt32 + (w1 & 0x1f8u)), t1 = t2_lds_f64(T32 + (w1 >> 9));
                r0 = fma(g0, t0, r0);
                r1 = fma(g1, t1, r1);
                wp += 256u;
            }
            if (jm & 1) {
                const unsigned w0 = t2_lds_u32(wp);
                r0 = fma(t2_lds_f64(gt32 + (w0 & 0x1f8u)), t2_lds_f64(T32 + (w0 >> 9)), r0);
            }
            return r0 + r1;
        };
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            if (((kindmask >> (2 * s)) & 3u) == 3u) continue;  // phase 2 of this group runs in another warp
            if (!(validmask >> s & 1u)) continue;
            double v = 0.0;
            if (!masked) v = gather(myinfo[8 * s + 5], words32 + 4u * (unsigned)(myinfo[8 * s + 4] + lane), acc[s]);
            yr[moff[s]] = v;
        }
        if (xgroup >= 0) {  // gather-only slot: a listed group owned by another warp (its mean term is in the list)
            double v = 0.0;
            if (!masked)
                v = gather(myinfo[8 * SLOTS + 5], words32 + 4u * (unsigned)(myinfo[8 * SLOTS + 4] + lane), 0.0);
            if (32 * xgroup + lane < a.N) yr[32 * xgroup + lane] = v;
        }
        if (a.nbuf == 1) __syncthreads();  // single exchange buffer: phase 1 of the next row overwrites it
        rp = nrp;
        len = nlen;
        masked = nmasked;
    }
}

int apply_ts2_launch(asgfem_ctx* ctx, const double* x, double* y, int64_t r0, int64_t r1) {
    Ts2Plan* P = tp2_of(ctx);
    if (!P) {
        int rc = apply_ts2_build(ctx);
        if (rc) return rc;
        P = tp2_of(ctx);
    }
    if (!P->usable)
        return fail(ctx, ASGFEM_ESTATE, "packed mode-stationary operator plan not available (too many modes / directions / long rows)");
    if (r1 <= r0) return 0;
    Ts2Args a;
    a.row0 = r0;
    a.nrows = r1;
    a.ld = ctx->ld;
    a.nnz = ctx->nnz;
    a.N = (int)ctx->ld;
    a.M = ctx->M;
    a.Mp = ctx->M + 1;
    a.D = P->D;
    a.nbuf = P->nbuf;
    a.kstr = P->kstr;
    a.kstr2 = P->kstr2;
    a.nwords = P->nwords;
    a.ndl = P->ndl;
    a.warps = P->warps;
    a.rowptr = ctx->d_rowptr;
    a.col = ctx->d_col;
    a.vals = ctx->d_vals;
    a.bmask = ctx->d_bmask;
    a.slotinfo = P->d_slotinfo;
    a.lane = P->d_lane;
    a.dl = P->d_dl;
    a.words = P->d_words;
    a.gtab = P->d_gtab;
    a.x = x;
    a.y = y;
    const int threads = 32 * P->warps;
    const size_t smem = P->smem_bytes;
#define LAUNCH_TS2(NSV, QV, SV, MV)                                                                                   \
    do {                                                                                                              \
        auto kern = k_apply_ts2<NSV, QV, SV, MV>;                                                                     \
        ASG_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));           \
        int per_sm = 1;                                                                                               \
        ASG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));                   \
        per_sm = std::max(1, std::min(per_sm, 4));                                                                    \
        int grid = (int)std::min<int64_t>(r1 - r0, (int64_t)P->grid_sms * per_sm);                                    \
        kern<<<grid, threads, smem, ctx->stream>>>(a);                                                                \
    } while (0)
#define LAUNCH_TS2_Q(NSV, SV, MV)         \
    do {                                  \
        if (P->Q == 3)                    \
            LAUNCH_TS2(NSV, 3, SV, MV);   \
        else                              \
            LAUNCH_TS2(NSV, 4, SV, MV);   \
    } while (0)
#define LAUNCH_TS2_S(NSV, MV)             \
    do {                                  \
        if (P->slots == 8)                \
            LAUNCH_TS2_Q(NSV, 8, MV);     \
        else                              \
            LAUNCH_TS2_Q(NSV, 4, MV);     \
    } while (0)
    if (P->nchunk_max > 1)
        LAUNCH_TS2_S(8, true);  // long rows: chunks of 8 columns
    else if (P->NS == 7)
        LAUNCH_TS2_S(7, false);
    else
        LAUNCH_TS2_S(8, false);
#undef LAUNCH_TS2_S
#undef LAUNCH_TS2_Q
#undef LAUNCH_TS2
    ASG_CUDA(ctx, cudaGetLastError());
    return 0;
}

}  // namespace asgfem
