// variant 6 of the fused SGFE operator: mode-stationary lanes, operands in registers, partial products exchanged
// through shared memory.
//
//   Y[i, mu] = sum_k K_0[i,j_k] X[j_k,mu] + sum_{(m,nu) ~ mu} g sum_k K_m[i,j_k] X[j_k,nu]      (mul!, :101-117)
//
// Variants 3-5 stage the rows X[j_k, :] in shared memory and pay one 8-byte shared-memory operand per FMA, which
// saturates the shared-memory pipe at a few percent of the fp64 rate (profiles/r01_apply_rows*, r01_apply_dir*).
// Here every lane OWNS modes (32 consecutive modes = one group, SLOTS groups per warp) and keeps X[j_k, nu] for
// its modes in registers, loaded straight from global memory (coalesced 256-byte segments, no staging):
//
//   phase 1  for every direction m (warp-uniform; K_m[i, j_k] broadcast from shared memory into registers)
//                T_m[nu] = sum_k K_m[i,j_k] X[j_k,nu]            all operands in registers
//            is formed by the lanes whose mode nu has a neighbour in direction m and written to the exchange buffer
//            at  disp[group, m] + lane.  The displacements are chosen on the host by first-fit packing of the 32-bit
//            activity masks (row-displacement compression: > 99 % dense for total-degree sets), so the store needs
//            no rank computation; m = 0 starts the accumulator of Y[i,nu].
//   phase 2  every lane gathers g * T_m[nu] for the couplings that END in its modes (16-bit Ts index + weight index
//            per coupling, lists resident in shared memory) into its register accumulator and writes Y[i, mu],
//            coalesced.  No atomics, fixed summation order.
//
// Shared-memory traffic per (row, pair (m,nu)): one 8-byte store + about one 8-byte gather, instead of 7 gathers and a
// read-modify-write.  Ts is double-buffered when it fits (one block barrier per row); the loads of the next row's
// X values are issued before phase 2 of the current row.  Rows longer than NS columns are processed in chunks of NS
// (Ts accumulates).
#include <algorithm>
#include <numeric>
#include <type_traits>

#include "common.h"

namespace asgfem {

struct TsPlan {
    bool usable = false, wide = false;
    int warps = 0, slots = 0, NS = 0, nbuf = 1, kstr = 0, nchunk_max = 1;
    int D = 0;  // T entries (index D is a dummy holding 0)
    int nwords = 0;
    size_t smem_bytes = 0;
    unsigned long long* d_act = nullptr;  // [warps*slots*32] bit m: lane's mode has a neighbour in direction m
    int32_t* d_slotinfo = nullptr;        // [warps*slots*4] group (-1 unused), -, words base, jmax
    uint32_t* d_rec = nullptr;            // [warps*(M+1)] records {slot mask, -, -, - | byte displacement per slot}
    uint32_t* d_words = nullptr;          // phase-2 words: Ts index << 12 | weight index << 3
    double* d_gtab = nullptr;             // [64]
};

static TsPlan* tp_of(asgfem_ctx* ctx) { return reinterpret_cast<TsPlan*>(ctx->tsplan); }

void apply_ts_free(asgfem_ctx* ctx) {
    TsPlan* P = tp_of(ctx);
    if (!P) return;
    void* ptrs[] = {P->d_act, P->d_slotinfo, P->d_rec, P->d_words, P->d_gtab};
    for (void* q : ptrs)
        if (q) cudaFree(q);
    delete P;
    ctx->tsplan = nullptr;
}

// carve-up of the dynamic shared memory (byte offsets); the 64-entry weight table is static shared memory, so its
// address is an immediate of the gather loads
struct TsLayout {
    uint32_t ts, ks, words, rec, sinfo, total;
};
// record of one (warp, direction): word 0 = mask of the slots with an active lane, then (from word 4 on; word 2 for two
// slots) the byte displacement of every slot in the exchange buffer
static __host__ __device__ inline int ts_rec_words(int slots) { return slots == 2 ? 4 : 4 + slots; }
static __host__ __device__ inline TsLayout ts_layout(int D, int nbuf, int Mp, int kstr, int nwords, int warps, int slots) {
    TsLayout L;
    uint32_t at = 0;
    L.ts = at;
    at += (uint32_t)nbuf * (uint32_t)((D + 2) / 2 * 2) * 8u;
    L.ks = at;
    at += 2u * (uint32_t)Mp * (uint32_t)kstr * 8u;
    L.words = at;
    at += (uint32_t)nwords * 4u;
    at = (at + 15u) & ~15u;
    L.rec = at;
    at += (uint32_t)warps * (uint32_t)Mp * (uint32_t)ts_rec_words(slots) * 4u;
    L.sinfo = at;
    at += (uint32_t)warps * (uint32_t)slots * 16u;
    L.total = at + 16u;
    return L;
}

int apply_ts_build(asgfem_ctx* ctx) {
    apply_ts_free(ctx);
    TsPlan* P = new TsPlan();
    ctx->tsplan = P;
    const int64_t N = ctx->N, nrows = ctx->n_owned >= 0 ? ctx->n_owned : ctx->n;
    const int M = ctx->M, Mp = M + 1;
    const Coupling& C = ctx->coup;
    if (N <= 0 || M > 63) return 0;
    P->wide = M > 31;
    int maxlen = 1;
    for (int64_t i = 0; i < nrows; ++i) maxlen = std::max<int>(maxlen, (int)(ctx->h_rowptr[i + 1] - ctx->h_rowptr[i]));
    P->NS = maxlen <= 7 ? 7 : 8;
    P->nchunk_max = (maxlen + P->NS - 1) / P->NS;
    P->kstr = (P->nchunk_max * P->NS + 1) / 2 * 2;
    if (P->kstr > 64) return 0;
    const int G = (int)((N + 31) / 32);
    P->slots = G > 64 ? 8 : (G > 16 ? 4 : 2);  // 16 warps x 4 slots measured faster than 8 x 8 at N = 2000
    if (const char* e = getenv("ASGFEM_TS_SLOTS")) {  // tuning knob
        int v = atoi(e);
        if (v == 2 || v == 4 || v == 8) P->slots = v;
    }
    P->warps = (G + P->slots - 1) / P->slots;
    if (P->warps > (P->slots == 8 ? 8 : 16)) return 0;  // register file: 8 slots need ~240 registers per thread

    // ---- pairs (m, nu): activity masks per (group, direction), packed into the exchange buffer by first fit ----
    std::vector<unsigned long long> actmode((size_t)N, 0ull);
    for (int64_t mu = 0; mu < N; ++mu)
        for (int32_t e = C.ptr[mu]; e < C.ptr[mu + 1]; ++e) {
            if (C.m[e] < 1 || C.m[e] > M) return 0;
            actmode[(size_t)C.nu[e]] |= 1ull << C.m[e];
        }
    std::vector<uint32_t> gmask((size_t)G * Mp, 0u);
    for (int64_t nu = 0; nu < N; ++nu)
        for (int m = 1; m <= M; ++m)
            if (actmode[(size_t)nu] >> m & 1ull) gmask[(size_t)(nu / 32) * Mp + m] |= 1u << (nu % 32);
    std::vector<int32_t> gdisp((size_t)G * Mp, 0);
    int D = 0;
    {
        std::vector<int> ord;
        for (int k = 0; k < G * Mp; ++k)
            if (gmask[k]) ord.push_back(k);
        std::stable_sort(ord.begin(), ord.end(),
                         [&](int a, int b) { return __builtin_popcount(gmask[a]) > __builtin_popcount(gmask[b]); });
        std::vector<uint8_t> occ;
        size_t low = 0;  // everything below is occupied
        for (int k : ord) {
            const uint32_t mk = gmask[k];
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
            gdisp[k] = (int32_t)d;
        }
    }
    if (D + 1 > 65535) return 0;
    P->D = D;
    auto pairidx = [&](int64_t nu, int m) { return gdisp[(size_t)(nu / 32) * Mp + m] + (int32_t)(nu % 32); };

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

    // ---- phase-2 lists per group: words[wbase + 32*j + lane] ------------------------------------------
    std::vector<int32_t> jmax((size_t)G, 0), wbase((size_t)G, 0), ndirs((size_t)G, 0);
    std::vector<uint32_t> words;
    const uint32_t dummy = (uint32_t)D << 12;  // weight index 0 -> 0.0, Ts[D] = 0.0
    for (int g = 0; g < G; ++g) {
        int jm = 0;
        for (int l = 0; l < 32; ++l) {
            int64_t mu = 32ll * g + l;
            if (mu < N) jm = std::max(jm, (int)(C.ptr[mu + 1] - C.ptr[mu]));
        }
        jm = (jm + 1) / 2 * 2;  // the gather loop is unrolled by 4 with a tail of 2
        jmax[g] = jm;
        wbase[g] = (int32_t)words.size();
        words.resize(words.size() + (size_t)jm * 32, dummy);
        for (int l = 0; l < 32; ++l) {
            int64_t mu = 32ll * g + l;
            if (mu >= N) continue;
            int j = 0;
            for (int32_t e = C.ptr[mu]; e < C.ptr[mu + 1]; ++e, ++j) {
                int gi = gindex(C.g[e]);
                if (gi < 0) return 0;
                words[(size_t)wbase[g] + (size_t)j * 32 + l] = ((uint32_t)pairidx(C.nu[e], C.m[e]) << 12) | ((uint32_t)gi << 3);
            }
        }
        for (int m = 1; m <= M; ++m) ndirs[g] += gmask[(size_t)g * Mp + m] != 0;
    }
    P->nwords = (int)words.size();

    // ---- groups -> (warp, slot): longest processing time first ------------------------------------------
    const int W = P->warps, S = P->slots;
    std::vector<int> order((size_t)G);
    std::iota(order.begin(), order.end(), 0);
    auto cost = [&](int g) { return (int64_t)(P->NS + 4) * (ndirs[g] + 1) + 6ll * jmax[g]; };
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost(a) > cost(b); });
    std::vector<int64_t> load((size_t)W, 0);
    std::vector<int> used((size_t)W, 0);
    std::vector<int32_t> slotinfo((size_t)W * S * 4, 0);
    for (int w = 0; w < W; ++w)
        for (int s = 0; s < S; ++s) slotinfo[((size_t)w * S + s) * 4] = -1;
    const int RW = ts_rec_words(S), RD = S == 2 ? 2 : 4;
    std::vector<uint32_t> rec((size_t)W * Mp * RW, 0u);
    std::vector<unsigned long long> act((size_t)W * S * 32, 0ull);
    // strided assignment: warp w owns groups w, w + W, w + 2W, ... (the heavy low-order groups are spread over the warps)
    (void)order;
    (void)load;
    (void)used;
    for (int g = 0; g < G; ++g) {
        const int w = g % W, s = g / W;
        int32_t* si = &slotinfo[((size_t)w * S + s) * 4];
        si[0] = g;
        si[1] = 0;
        si[2] = wbase[g];
        si[3] = jmax[g];
        for (int m = 1; m <= M; ++m) {
            uint32_t* r = &rec[((size_t)w * Mp + m) * RW];
            r[RD + s] = (uint32_t)gdisp[(size_t)g * Mp + m] * 8u;
            if (gmask[(size_t)g * Mp + m]) r[0] |= 1u << s;
        }
        for (int l = 0; l < 32; ++l) {
            int64_t nu = 32ll * g + l;
            act[((size_t)w * S + s) * 32 + l] = nu < N ? actmode[(size_t)nu] : 0ull;
        }
    }

    const size_t limit = 226 * 1024;  // 512 bytes of static shared memory (weight table) come on top
    P->nbuf = 2;
    P->smem_bytes = ts_layout(P->D, 2, Mp, P->kstr, P->nwords, W, S).total;
    if (P->smem_bytes > limit) {
        P->nbuf = 1;
        P->smem_bytes = ts_layout(P->D, 1, Mp, P->kstr, P->nwords, W, S).total;
        if (P->smem_bytes > limit) return 0;
    }
    int rc = 0;
    rc |= dev_upload(ctx, &P->d_act, act);
    rc |= dev_upload(ctx, &P->d_slotinfo, slotinfo);
    rc |= dev_upload(ctx, &P->d_rec, rec);
    rc |= dev_upload(ctx, &P->d_words, words);
    rc |= dev_upload(ctx, &P->d_gtab, gtab);
    if (rc) return rc;
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    P->usable = true;
    return 0;
}

bool apply_ts_preferred(asgfem_ctx* ctx) {
    if (!tp_of(ctx) && apply_ts_build(ctx)) return false;
    TsPlan* P = tp_of(ctx);
    return P && P->usable && P->nchunk_max == 1;
}

struct TsArgs {
    int64_t row0, nrows, ld, nnz;  // rows [row0, nrows)
    int N, M, Mp, D, nbuf, kstr, nwords, warps;
    const int64_t* rowptr;
    const int32_t* col;
    const double* vals;
    const uint8_t* bmask;
    const unsigned long long* act;
    const int32_t* slotinfo;
    const uint32_t* rec;
    const uint32_t* words;
    const double* gtab;
    const double* x;
    double* y;
};

__device__ __forceinline__ double ldg_f64_volatile(const double* p) {
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ts_lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned ts_lds_u32(unsigned addr) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// predicated shared-memory store: no branch around the store, 32-bit address
__device__ __forceinline__ void ts_sts_f64_if(unsigned addr, double v, unsigned on) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.f64 [%0], %1;\n\t}" ::"r"(addr), "d"(v), "r"(on));
}
// store if (bits & mask) != 0: the test is one LOP3 with predicate output
__device__ __forceinline__ void ts_sts_f64_ifbit(unsigned addr, double v, unsigned bits, unsigned mask) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tand.b32 t, %2, %3;\n\tsetp.ne.u32 p, t, 0;\n\t@p st.shared.f64 [%0], %1;\n\t}" ::"r"(addr),
                 "d"(v), "r"(bits), "r"(mask));
}
__device__ __forceinline__ void ts_sts_f64_ifbit(unsigned addr, double v, unsigned long long bits, unsigned long long mask) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 t;\n\tand.b64 t, %2, %3;\n\tsetp.ne.u64 p, t, 0;\n\t@p st.shared.f64 [%0], %1;\n\t}" ::"r"(addr),
                 "d"(v), "l"(bits), "l"(mask));
}

// T_m for the slots BASE + {bits of MK}: independent FMA chains, then the stores of the active lanes
template <int NS, int SLOTS, int BASE, unsigned MK, bool FIRST, typename ACT>
__device__ __forceinline__ void ts_dir_block(const double (&kr)[NS + (NS & 1)], const double (&x)[SLOTS][NS],
                                             const ACT (&act)[SLOTS], const unsigned* d, int m, unsigned Tl32) {
    constexpr int NB = SLOTS - BASE;  // all slots of the warp: NB independent FMA chains
    double t[NB];
#pragma unroll
    for (int q = 0; q < NB; ++q)
        if (MK >> q & 1u) t[q] = kr[0] * x[BASE + q][0];
#pragma unroll
    for (int k = 1; k < NS; ++k) {
#pragma unroll
        for (int q = 0; q < NB; ++q)
            if (MK >> q & 1u) t[q] = fma(kr[k], x[BASE + q][k], t[q]);
    }
#pragma unroll
    for (int q = 0; q < NB; ++q)
        if (MK >> q & 1u) {
            const ACT mbit = (ACT)1 << m;
            const unsigned addr = Tl32 + d[q];
            if (FIRST) {
                ts_sts_f64_ifbit(addr, t[q], act[BASE + q], mbit);
            } else if (act[BASE + q] & mbit) {
                ts_sts_f64_if(addr, ts_lds_f64(addr) + t[q], 1u);
            }
        }
}

template <int NS, int SLOTS, int BASE, bool FIRST, typename ACT>
__device__ __forceinline__ void ts_dir_switch(unsigned mk, const double (&kr)[NS + (NS & 1)], const double (&x)[SLOTS][NS],
                                              const ACT (&act)[SLOTS], const unsigned* d, int m, unsigned Tl32) {
#define TS_CASE(V) \
    case V: ts_dir_block<NS, SLOTS, BASE, V, FIRST, ACT>(kr, x, act, d, m, Tl32); break;
    // all slots of the block at once: the FMA chains of unused slots are wasted, but 4 independent chains per warp keep
    // the fp64 pipe busy (a per-mask switch with only the needed chains was measured 15 % slower: latency-bound)
    if (mk) ts_dir_block<NS, SLOTS, BASE, (1u << (SLOTS - BASE)) - 1u, FIRST, ACT>(kr, x, act, d, m, Tl32);
#undef TS_CASE
}

template <int NS, int SLOTS, typename ACT>
__global__ void __launch_bounds__(SLOTS == 8 ? 256 : 512, 1) k_apply_ts(TsArgs a) {
    extern __shared__ __align__(16) unsigned char ts_raw[];
    __shared__ __align__(512) double gt[64];
    // CSR metadata of the CTA's upcoming rows, fetched with cp.async two / three rows ahead so that the address chain
    // rowptr -> col -> X is off the critical path when the X loads of the next row are issued
    __shared__ __align__(16) long long info_rp[4][2];  // [row j % 4]: rowptr[row], rowptr[row + 1]
    __shared__ int info_msk[4];                         // Dirichlet flag (1 also for rows past the end)
    __shared__ int cols_ring[2][NS];                    // [row j % 2]: first NS column indices
    const unsigned raw32 = (unsigned)__cvta_generic_to_shared(ts_raw);
    const TsLayout L = ts_layout(a.D, a.nbuf, a.Mp, a.kstr, a.nwords, a.warps, SLOTS);
    const int Dpad = (a.D + 2) / 2 * 2;
    double* Ts = reinterpret_cast<double*>(ts_raw + L.ts);               // [nbuf][Dpad]
    double* Ks = reinterpret_cast<double*>(ts_raw + L.ks);               // [2][Mp][kstr]
    uint32_t* words = reinterpret_cast<uint32_t*>(ts_raw + L.words);     // [nwords]
    uint32_t* rec = reinterpret_cast<uint32_t*>(ts_raw + L.rec);         // [warps][Mp][RW]
    int32_t* sinfo = reinterpret_cast<int32_t*>(ts_raw + L.sinfo);       // [warps][SLOTS][4]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
    const unsigned gt32 = (unsigned)__cvta_generic_to_shared(gt), ts32 = raw32 + L.ts, words32 = raw32 + L.words;

    for (int k = tid; k < 64; k += nthr) gt[k] = a.gtab[k];
    for (int k = tid; k < a.nwords; k += nthr) words[k] = a.words[k];
    constexpr int RW = SLOTS == 2 ? 4 : 4 + SLOTS;
    for (int k = tid; k < a.warps * a.Mp * RW; k += nthr) rec[k] = a.rec[k];
    for (int k = tid; k < a.warps * SLOTS * 4; k += nthr) sinfo[k] = a.slotinfo[k];
    if (tid < a.nbuf) Ts[(size_t)tid * Dpad + a.D] = 0.0;  // dummy entry of the padded gather lists
    ACT act[SLOTS];
    int moff[SLOTS];        // own mode of the slot (clamped to a valid mode for idle lanes / unused slots: loads stay in
    unsigned validmask = 0; // bounds and unpredicated, nothing is stored for them)
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        act[s] = (ACT)a.act[((size_t)warp * SLOTS + s) * 32 + lane];
        const int g = a.slotinfo[((size_t)warp * SLOTS + s) * 4];
        const bool valid = g >= 0 && 32 * g + lane < a.N;
        moff[s] = valid ? 32 * g + lane : (g >= 0 ? a.N - 1 : 0);
        validmask |= valid ? 1u << s : 0u;
    }
    const uint4* myrec = reinterpret_cast<const uint4*>(rec + (size_t)warp * a.Mp * RW);
    const int32_t* myinfo = sinfo + warp * SLOTS * 4;

    double x[SLOTS][NS];
    // stage K values of `row` into Ks[buf]; issue the X loads of its first column chunk
    auto stage_k = [&](int64_t rp, int len, int buf) {
        double* dst = Ks + (size_t)buf * a.Mp * a.kstr;
        for (int idx = tid; idx < a.Mp * a.kstr; idx += nthr) {
            const int m = idx / a.kstr, k = idx - m * a.kstr;
            dst[idx] = k < len ? __ldg(a.vals + (int64_t)m * a.nnz + rp + k) : 0.0;
        }
    };
    const unsigned long long xbase = (unsigned long long)a.x, ldb = (unsigned long long)a.ld * 8ull;
    auto load_x = [&](int64_t rp, int len, int c0, const int* cols_s) {
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            if (c0 + k < len) {  // block-uniform
                const unsigned cj = cols_s ? (unsigned)cols_s[k] : (unsigned)__ldg(a.col + rp + c0 + k);
                const unsigned long long rowp = xbase + (unsigned long long)cj * ldb;
#pragma unroll
                for (int s = 0; s < SLOTS; ++s)
                    x[s][k] = ldg_f64_volatile(reinterpret_cast<const double*>(rowp + (unsigned)(8 * moff[s])));
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
    // metadata loader: threads 0..NS-1 fetch column ids, thread NS the row pointers, thread NS+1 the Dirichlet flags
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
    auto fetch_msk = [&](int64_t j) -> int {
        const int64_t r = a.row0 + blockIdx.x + j * G;
        return r < a.nrows ? (int)a.bmask[r] : 1;
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
    fetch_info(1);
    fetch_info(2);
    if (tid == NS + 1) {
        info_msk[1] = fetch_msk(1);
        info_msk[2] = fetch_msk(2);
    }
    cp_wait_all();
    __syncthreads();
    fetch_cols(1);
    cp_wait_all();
    __syncthreads();
    int buf = 0;
    for (int64_t it = 0; row < a.nrows; row += gridDim.x, buf ^= 1, ++it) {
        // loader: column ids of row it+2 (its row pointers arrived one iteration ago), row pointers of row it+3
        fetch_cols(it + 2);
        fetch_info(it + 3);
        const int msk3 = tid == NS + 1 ? fetch_msk(it + 3) : 0;
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
        const int kper = (a.Mp * a.kstr + nthr - 1) / nthr;  // K values per thread (1 for P1: 168 values, 512 threads)
        double kval0 = 0.0;
        if (!nmasked && kper == 1 && tid < a.Mp * a.kstr) {
            const int m = tid / a.kstr, k = tid - m * a.kstr;
            if (k < nlen) kval0 = __ldg(a.vals + (int64_t)m * a.nnz + nrp + k);
        }
        const int tb = a.nbuf == 2 ? buf : 0;
        const unsigned T32 = ts32 + (unsigned)tb * (unsigned)Dpad * 8u;
        unsigned Tl32 = T32 + 8u * (unsigned)lane;
        asm volatile("mov.u32 %0, %0;" : "+r"(Tl32));  // opaque: kept in a register instead of being rematerialised per store
        const double* K = Ks + (size_t)buf * a.Mp * a.kstr;
        double acc[SLOTS];
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) acc[s] = 0.0;
        // ---------------- phase 1: T_m[nu] for the own modes, all operands in registers ---------------------
        if (!masked) {
            for (int c0 = 0; c0 < len; c0 += NS) {
                if (c0 > 0) load_x(rp, len, c0, nullptr);
                constexpr int NK = NS + (NS & 1);  // K rows are read as 16-byte pairs (the pad entry is zero)
                double kr[NK];
                {
                    const double2* K2 = reinterpret_cast<const double2*>(K + c0);
#pragma unroll
                    for (int k = 0; k < NK / 2; ++k) {
                        const double2 v = K2[k];
                        kr[2 * k] = v.x, kr[2 * k + 1] = v.y;
                    }
                }
#pragma unroll
                for (int k = 0; k < NS; ++k) {
#pragma unroll
                    for (int s = 0; s < SLOTS; ++s) acc[s] = fma(kr[k], x[s][k], acc[s]);
                }
                for (int m = 1; m <= a.M; ++m) {
                    const uint4* r = myrec + m * (RW / 4);
                    const uint4 r0 = r[0];
                    if (r0.x == 0u) continue;
                    const double2* K2 = reinterpret_cast<const double2*>(K + m * a.kstr + c0);
#pragma unroll
                    for (int k = 0; k < NK / 2; ++k) {
                        const double2 v = K2[k];
                        kr[2 * k] = v.x, kr[2 * k + 1] = v.y;
                    }
                    auto dirs = [&](auto first_tag) {
                        constexpr bool FIRST = decltype(first_tag)::value;
                        if constexpr (SLOTS == 2) {
                            const unsigned d[2] = {r0.z, r0.w};
                            ts_dir_switch<NS, SLOTS, 0, FIRST, ACT>(r0.x, kr, x, act, d, m, Tl32);
                        } else {
                            const uint4 r1 = r[1];
                            if constexpr (SLOTS == 8) {
                                const uint4 r2 = r[2];
                                const unsigned d[8] = {r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
                                ts_dir_switch<NS, SLOTS, 0, FIRST, ACT>(r0.x, kr, x, act, d, m, Tl32);
                            } else {
                                const unsigned d[4] = {r1.x, r1.y, r1.z, r1.w};
                                ts_dir_switch<NS, SLOTS, 0, FIRST, ACT>(r0.x, kr, x, act, d, m, Tl32);
                            }
                        }
                    };
                    if (c0 == 0)
                        dirs(std::true_type{});
                    else
                        dirs(std::false_type{});
                }
            }
        }
        // ---------------- next row: K values into the other buffer, X loads in flight during phase 2 --------
        if (!nmasked) {
            if (kper == 1) {
                if (tid < a.Mp * a.kstr) (Ks + (size_t)(buf ^ 1) * a.Mp * a.kstr)[tid] = kval0;
            } else {
                stage_k(nrp, nlen, buf ^ 1);
            }
            load_x(nrp, nlen, 0, cols_ring[(it + 1) & 1]);
        }
        if (tid == NS + 1) info_msk[(it + 3) & 3] = msk3;
        cp_wait_all();
        __syncthreads();
        // ---------------- phase 2: gather the couplings that end in the own modes ---------------------------
        double* yr = a.y + row * a.ld;
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            if (!(validmask >> s & 1u)) continue;
            double r0 = acc[s], r1 = 0.0;
            if (!masked) {
                const int jm = myinfo[4 * s + 3];
                unsigned wp = words32 + 4u * (unsigned)(myinfo[4 * s + 2] + lane);
                for (int j = 0; j + 4 <= jm; j += 4, wp += 512u) {
                    const unsigned w0 = ts_lds_u32(wp), w1 = ts_lds_u32(wp + 128u), w2 = ts_lds_u32(wp + 256u),
                                   w3 = ts_lds_u32(wp + 384u);
                    const double g0 = ts_lds_f64(gt32 + (w0 & 0x1f8u)), t0 = ts_lds_f64(T32 + (w0 >> 9));
                    const double g1 = ts_lds_f64(gt32 + (w1 & 0x1f8u)), t1 = ts_lds_f64(T32 + (w1 >> 9));
                    const double g2 = ts_lds_f64(gt32 + (w2 & 0x1f8u)), t2 = ts_lds_f64(T32 + (w2 >> 9));
                    const double g3 = ts_lds_f64(gt32 + (w3 & 0x1f8u)), t3 = ts_lds_f64(T32 + (w3 >> 9));
                    r0 = fma(g0, t0, r0);
                    r1 = fma(g1, t1, r1);
                    r0 = fma(g2, t2, r0);
                    r1 = fma(g3, t3, r1);
                }
                if (jm & 2) {
                    const unsigned w0 = ts_lds_u32(wp), w1 = ts_lds_u32(wp + 128u);
                    const double g0 = ts_lds_f64(gt32 + (w0 & 0x1f8u)), t0 = ts_lds_f64(T32 + (w0 >> 9));
                    const double g1 = ts_lds_f64(gt32 + (w1 & 0x1f8u)), t1 = ts_lds_f64(T32 + (w1 >> 9));
                    r0 = fma(g0, t0, r0);
                    r1 = fma(g1, t1, r1);
                }
            }
            yr[moff[s]] = masked ? 0.0 : r0 + r1;
        }
        if (a.nbuf == 1) __syncthreads();  // single exchange buffer: phase 1 of the next row overwrites it
        rp = nrp;
        len = nlen;
        masked = nmasked;
    }
}

int apply_ts_launch(asgfem_ctx* ctx, const double* x, double* y, int64_t r0, int64_t r1) {
    TsPlan* P = tp_of(ctx);
    if (!P) {
        int rc = apply_ts_build(ctx);
        if (rc) return rc;
        P = tp_of(ctx);
    }
    if (!P->usable)
        return fail(ctx, ASGFEM_ESTATE, "mode-stationary operator plan not available (too many modes / directions / long rows)");
    if (r1 <= r0) return 0;
    TsArgs a;
    a.row0 = r0;
    a.nrows = r1;
    a.ld = ctx->ld;
    a.nnz = ctx->nnz;
    a.N = (int)ctx->N;
    a.M = ctx->M;
    a.Mp = ctx->M + 1;
    a.D = P->D;
    a.nbuf = P->nbuf;
    a.kstr = P->kstr;
    a.nwords = P->nwords;
    a.warps = P->warps;
    a.rowptr = ctx->d_rowptr;
    a.col = ctx->d_col;
    a.vals = ctx->d_vals;
    a.bmask = ctx->d_bmask;
    a.act = P->d_act;
    a.slotinfo = P->d_slotinfo;
    a.rec = P->d_rec;
    a.words = P->d_words;
    a.gtab = P->d_gtab;
    a.x = x;
    a.y = y;
    const int threads = 32 * P->warps;
    const size_t smem = P->smem_bytes;
#define LAUNCH_TS(NSV, SV, ACTT)                                                                                      \
    do {                                                                                                              \
        auto kern = k_apply_ts<NSV, SV, ACTT>;                                                                        \
        ASG_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));           \
        int per_sm = 1;                                                                                               \
        ASG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));                   \
        per_sm = std::max(1, std::min(per_sm, 4));                                                                    \
        int grid = (int)std::min<int64_t>(r1 - r0, 148ll * per_sm);                                                   \
        kern<<<grid, threads, smem, ctx->stream>>>(a);                                                                \
    } while (0)
#define LAUNCH_TS_A(NSV, SV)                              \
    do {                                                  \
        if (P->wide)                                      \
            LAUNCH_TS(NSV, SV, unsigned long long);       \
        else                                              \
            LAUNCH_TS(NSV, SV, unsigned);                 \
    } while (0)
#define LAUNCH_TS_NS(SV)                                 \
    do {                                                 \
        if (P->NS == 7)                                  \
            LAUNCH_TS_A(7, SV);                          \
        else                                             \
            LAUNCH_TS_A(8, SV);                          \
    } while (0)
    if (P->slots == 8)
        LAUNCH_TS_NS(8);
    else if (P->slots == 4)
        LAUNCH_TS_NS(4);
    else
        LAUNCH_TS_NS(2);
#undef LAUNCH_TS_NS
#undef LAUNCH_TS_A
#undef LAUNCH_TS
    ASG_CUDA(ctx, cudaGetLastError());
    return 0;
}

}  // namespace asgfem
