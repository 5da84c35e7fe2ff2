// Mean-based preconditioner  Z[:,mu] = K_0^{-1} R[:,mu]  for all modes at once
// (LinearAlgebra.ldiv!(y, P::MyPreconditionerPrimal, b), src/modelproblems/solvers_poisson_primal.jl:46-78;
//  the reference runs N sequential UMFPACK solves, one per mode block).
//
// Factor P K_0 P^T = L L^T from chol.cpp (nested dissection).  The N right-hand sides are independent, so the
// triangular solves need no inter-CTA synchronisation: every CTA owns a tile of MT modes and performs the whole
// sweep by itself.  Inside the CTA the sweep is TREE-PARALLEL and right-looking:
//   * every leaf and every separator (chunk) of the dissection tree is a task; tasks of the same tree depth are
//     independent, so one block barrier per tree depth replaces one barrier per row level;
//   * small tasks (<= SMALL rows) are solved by single warps, 32 of them concurrently: rows and the task's diagonal
//     block are staged in the warp's shared-memory slice, rows are eliminated in order (the two half-warps split a
//     row's entries), then the task pushes  W[r] -= L[r, task] * Z_task  to the rows r above it - one (row, task)
//     segment of the CSR row at a time, sources from shared memory, one fp64 RED per segment and mode (sibling
//     tasks may hit the same ancestor row concurrently);
//   * big tasks (separator chunks up to BW rows) are handled by the whole CTA: panels of PANEL rows, the part left
//     of the panel for all panel rows in parallel (4 half-warps per row), the PANEL x PANEL triangle by one
//     half-warp from shared memory, pushes by all 64 half-warps.
// The backward sweep L^T z = y is the same algorithm on the reversed numbering k -> n-1-k with the tree walked from
// the root down.  Work vectors live in elimination order (W[k,:] <-> dof perm[k]).
#include <algorithm>
#include <array>

#include "common.h"

namespace asgfem {

constexpr int MT = 16;                 // modes per CTA (half a warp wide)
constexpr int NWARP = 32;
constexpr int TRSV_THREADS = 32 * NWARP;  // 1024
constexpr int RS = TRSV_THREADS / MT;  // half-warps per CTA
constexpr int BW = 256;                // max rows of a big task
constexpr int SMALL = 24;              // max rows of a warp task (= leaf size of the dissection)
constexpr int SMALL_DE = SMALL * (SMALL - 1) / 2;  // max entries of its diagonal block
constexpr int SMALL_SEG = 56;           // push segments of a warp task staged in shared memory (more are read from L2)
constexpr int PANEL = 16;
constexpr int PER = RS / PANEL;        // half-warps per panel row

struct TriDev {  // one triangular system in its own (forward) numbering
    uint16_t* idx = nullptr;       // push entries, stored source block after source block, segment after segment:
    double* val = nullptr;         // block-local source row and value (a task streams one contiguous range)
    int32_t* segptr = nullptr;     // per block: range of push segments
    int32_t* seg = nullptr;        // per segment: target row, first nonzero, length
    double* dinv = nullptr;        // per row
    int32_t* drow = nullptr;       // per row: offset of its diagonal-block entries (n+1)
    int32_t* dsplit = nullptr;     // per row: number of those entries left of the row's panel
    uint16_t* didx = nullptr;      // diagonal-block entries, block-local column ids
    double* dval = nullptr;
    int32_t* blk_start = nullptr;  // nblocks+1
    int32_t* step_ptr = nullptr;   // nsteps+1 -> ranges of `tasks`
    int32_t* step_nsmall = nullptr;  // per step: the first so many tasks are small
    int32_t* tasks = nullptr;      // block ids
    int nblocks = 0, nsteps = 0;
};

struct PrecondPlan {
    int64_t nred = 0;
    int32_t* d_perm = nullptr;
    TriDev fwd, bwd;
    double* d_work = nullptr;  // nred x ld
    int64_t lnz = 0;
};

static void free_tri(TriDev& T) {
    void* ptrs[] = {T.idx, T.val, T.segptr, T.seg, T.dinv, T.drow, T.dsplit, T.didx, T.dval, T.blk_start, T.step_ptr,
                    T.step_nsmall, T.tasks};
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

// shared-memory slice of one warp during the small-task phase
struct WarpSlice {
    double z[SMALL][MT];
    double dval[SMALL_DE];
    double dinv[SMALL];
    int32_t drow[SMALL + 1];
    int32_t seg[3 * SMALL_SEG];
    uint16_t didx[SMALL_DE + 2];
};
// shared memory of the big-task phase (aliases the warp slices; the phases are separated by block barriers)
struct BigSlice {
    double z[BW][MT];
    double red[RS][MT + 1];
    double dinv[BW];
    double ptri[PANEL * PANEL];
    int32_t drow[BW + 1];
    int32_t dsplit[BW];
};
constexpr size_t TRSV_SMEM = sizeof(WarpSlice) * NWARP > sizeof(BigSlice) ? sizeof(WarpSlice) * NWARP : sizeof(BigSlice);

// Push segments of one half-warp.  The 16 lanes fetch 16 (value, source row) entries at once (coalesced, the entries of
// a task are one contiguous stream) and hand them round with shuffles.  Software pipeline across chunks AND segments:
// while one chunk is consumed the next one - of the same segment or the first of the half-warp's next segment - is in
// flight, so the L2 latency is paid once per half-warp instead of once per segment.
struct SegDesc {
    int r, p0, len;
};
// Each half-warp walks its own segments; the first chunk of the next segment is prefetched during the last chunk of
// the current one.  (Walking the two halves of a warp in lockstep to keep the warp converged was measured: no gain.)
template <class GetDesc, class Emit>
__device__ __forceinline__ void push_segments(int first, int end, int stride, GetDesc get_desc, Emit emit,
                                              const double* __restrict__ val, const uint16_t* __restrict__ idx,
                                              const double (*Z)[MT], int m, unsigned hmask, int hbase) {
    int nmine = first < end ? (end - first + stride - 1) / stride : 0;
    const int nit = nmine;
    if (nit == 0) return;
    SegDesc cur{0, 0, 0};
    if (nmine > 0) cur = get_desc(first);
    double nv = 0.0;
    int ni = 0;
    if (m < cur.len) {
        nv = __ldg(val + cur.p0 + m);
        ni = __ldg(idx + cur.p0 + m);
    }
    for (int it = 0; it < nit; ++it) {
        const int sgn = first + (it + 1) * stride;
        const bool has_next = sgn < end;
        SegDesc nxt{0, 0, 0};
        if (has_next) nxt = get_desc(sgn);
        const int lenmax = cur.len;
        double a0 = 0.0, a1 = 0.0;
        for (int p = 0; p < lenmax; p += MT) {
            const double myv = nv;
            const int myi = ni;
            nv = 0.0;
            ni = 0;
            if (p + MT < cur.len) {
                if (p + MT + m < cur.len) {
                    nv = __ldg(val + cur.p0 + p + MT + m);
                    ni = __ldg(idx + cur.p0 + p + MT + m);
                }
            } else if (p + MT >= lenmax && m < nxt.len) {  // last chunk of the pair: first chunk of the next segment
                nv = __ldg(val + nxt.p0 + m);
                ni = __ldg(idx + nxt.p0 + m);
            }
            const int cnt = min(MT, lenmax - p);
            for (int u = 0; u < cnt; u += 2) {  // entries past the own length carry v = 0, i = 0
                const double v0 = __shfl_sync(hmask, myv, hbase + u), v1 = __shfl_sync(hmask, myv, hbase + ((u + 1) & 15));
                const int i0 = __shfl_sync(hmask, myi, hbase + u), i1 = __shfl_sync(hmask, myi, hbase + ((u + 1) & 15));
                a0 = fma(v0, Z[i0][m], a0);
                a1 = fma(v1, Z[i1][m], a1);
            }
        }
        if (it < nmine) emit(cur.r, a0 + a1);
        cur = nxt;
    }
}

__global__ void __launch_bounds__(TRSV_THREADS, 1)
k_trsv_tree(double* __restrict__ w, int64_t ld, int64_t n, int rev, TriDev T) {
    extern __shared__ __align__(16) unsigned char trsv_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m = lane & (MT - 1), half = lane >> 4;  // mode within the tile, half-warp within the warp
    const int hw = tid / MT;                          // half-warp within the CTA
    const int64_t mode = (int64_t)blockIdx.x * MT + m;
    const unsigned hmask = half ? 0xffff0000u : 0x0000ffffu;
    const int hbase = half * MT;
    auto phys = [&](int64_t k) { return rev ? (n - 1 - k) : k; };

    for (int st = 0; st < T.nsteps; ++st) {
        const int t0 = T.step_ptr[st], t1 = T.step_ptr[st + 1], ns = T.step_nsmall[st];
        // ================= small tasks: one warp each ======================================================
        if (ns > 0) {
            WarpSlice& S = reinterpret_cast<WarpSlice*>(trsv_raw)[warp];
            for (int t = t0 + warp; t < t0 + ns; t += NWARP) {
                const int b = T.tasks[t];
                const int j0 = T.blk_start[b], len = T.blk_start[b + 1] - j0;
                const int d0 = T.drow[j0], nde = T.drow[j0 + len] - d0;
                for (int r = half; r < len; r += 2) S.z[r][m] = __ldcg(w + phys(j0 + r) * ld + mode);
                if (lane <= len) S.drow[lane] = T.drow[j0 + lane] - d0;
                if (lane < len) S.dinv[lane] = T.dinv[j0 + lane];
                for (int e = lane; e < nde; e += 32) {
                    S.dval[e] = T.dval[d0 + e];
                    S.didx[e] = T.didx[d0 + e];
                }
                const int s0 = T.segptr[b], s1 = T.segptr[b + 1];
                for (int e = lane; e < 3 * min(s1 - s0, SMALL_SEG); e += 32) S.seg[e] = T.seg[3 * s0 + e];
                __syncwarp();
                for (int r = 0; r < len; ++r) {
                    const int e0 = S.drow[r], e1 = S.drow[r + 1];
                    double dota = 0.0, dotb = 0.0;
                    int e = e0 + half;
                    for (; e + 2 < e1; e += 4) {
                        dota = fma(S.dval[e], S.z[S.didx[e]][m], dota);
                        dotb = fma(S.dval[e + 2], S.z[S.didx[e + 2]][m], dotb);
                    }
                    if (e < e1) dota = fma(S.dval[e], S.z[S.didx[e]][m], dota);
                    double dot = dota + dotb;
                    dot += __shfl_xor_sync(0xffffffffu, dot, 16);
                    const double z = (S.z[r][m] - dot) * S.dinv[r];
                    __syncwarp();
                    if (half == 0) S.z[r][m] = z;
                    __syncwarp();
                }
                for (int r = half; r < len; r += 2) w[phys(j0 + r) * ld + mode] = S.z[r][m];
                push_segments(
                    s0 + half, s1, 2,
                    [&](int sg) {
                        const int q = sg - s0;
                        return q < SMALL_SEG ? SegDesc{S.seg[3 * q], S.seg[3 * q + 1], S.seg[3 * q + 2]}
                                             : SegDesc{T.seg[3 * sg], T.seg[3 * sg + 1], T.seg[3 * sg + 2]};
                    },
                    [&](int r, double dot) { atomicAdd(w + phys(r) * ld + mode, -dot); }, T.val, T.idx, S.z, m, hmask, hbase);
                __syncwarp();
            }
            __threadfence();
            __syncthreads();
        }
        // ================= big tasks: whole CTA, one after the other =======================================
        if (t1 > t0 + ns) {
            BigSlice& B = *reinterpret_cast<BigSlice*>(trsv_raw);
            for (int t = t0 + ns; t < t1; ++t) {
                const int b = T.tasks[t];
                const int j0 = T.blk_start[b], width = T.blk_start[b + 1] - j0;
                const int d0 = T.drow[j0];
                for (int slot = hw; slot < width; slot += RS) B.z[slot][m] = __ldcg(w + phys(j0 + slot) * ld + mode);
                for (int k = tid; k < width; k += TRSV_THREADS) {
                    B.dinv[k] = T.dinv[j0 + k];
                    B.dsplit[k] = T.dsplit[j0 + k];
                }
                for (int k = tid; k <= width; k += TRSV_THREADS) B.drow[k] = T.drow[j0 + k] - d0;
                __syncthreads();
                const double* gv = T.dval + d0;
                const uint16_t* gi = T.didx + d0;
                for (int p0 = 0; p0 < width; p0 += PANEL) {
                    // (a) part of every panel row left of the panel: PER half-warps per row, 16 entries per fetch
                    const int rl = p0 + hw / PER, part = hw % PER;
                    double acc = 0.0;
                    if (rl < width) {
                        const int e0 = B.drow[rl], nleft = B.dsplit[rl];
                        // contiguous quarter of the row's left entries
                        const int q0 = (int)(((int64_t)nleft * part) / PER), q1 = (int)(((int64_t)nleft * (part + 1)) / PER);
                        double a0 = 0.0, a1 = 0.0;
                        for (int p = q0; p < q1; p += MT) {
                            const int cnt = min(MT, q1 - p);
                            double myv = 0.0;
                            int myi = 0;
                            if (m < cnt) {
                                myv = __ldg(gv + e0 + p + m);
                                myi = __ldg(gi + e0 + p + m);
                            }
#pragma unroll
                            for (int u = 0; u < MT; u += 2) {
                                const double v0 = __shfl_sync(hmask, myv, hbase + u), v1 = __shfl_sync(hmask, myv, hbase + u + 1);
                                const int i0 = __shfl_sync(hmask, myi, hbase + u), i1 = __shfl_sync(hmask, myi, hbase + u + 1);
                                a0 = fma(v0, B.z[i0][m], a0);
                                a1 = fma(v1, B.z[i1][m], a1);
                            }
                        }
                        acc = a0 + a1;
                    }
                    B.red[hw][m] = acc;
                    // the panel triangle travels to shared memory meanwhile
                    {
                        // dense PANEL x PANEL table ptri[row][col] = L[p0+row, p0+col] (zero where absent)
                        if (tid < PANEL * PANEL) B.ptri[tid] = 0.0;
                        __syncwarp();
                        const int r = p0 + tid / PANEL, c = tid % PANEL;  // threads 0..255 = warps 0..7
                        if (tid < PANEL * PANEL && r < width) {
                            const int e = B.drow[r] + B.dsplit[r] + c;
                            if (e < B.drow[r + 1]) B.ptri[(tid / PANEL) * PANEL + (__ldg(gi + e) - p0)] = __ldg(gv + e);
                        }
                    }
                    __syncthreads();
                    // (b) the PANEL x PANEL triangle, right-looking: half-warp h owns panel row p0+h and keeps its running
                    //     sum in a register; row after row the owner finalises z_r, everybody below subtracts L[r',r] z_r
                    {
                        const int pend = min(p0 + PANEL, width);
                        const int myr = p0 + hw;  // only half-warps 0..PANEL-1 own a row
                        double sum = 0.0;
                        if (hw < PANEL && myr < pend) {
                            sum = B.z[myr][m];
#pragma unroll
                            for (int j = 0; j < PER; ++j) sum -= B.red[hw * PER + j][m];
                        }
                        for (int r = p0; r < pend; ++r) {
                            if (myr == r && hw < PANEL) B.z[r][m] = sum * B.dinv[r];
                            __syncthreads();
                            if (hw < PANEL && myr > r && myr < pend) sum = fma(-B.ptri[hw * PANEL + (r - p0)], B.z[r][m], sum);
                        }
                    }
                    __syncthreads();
                }
                for (int slot = hw; slot < width; slot += RS) w[phys(j0 + slot) * ld + mode] = B.z[slot][m];
                const int s0 = T.segptr[b], s1 = T.segptr[b + 1];
                push_segments(
                    s0 + hw, s1, RS, [&](int sg) { return SegDesc{T.seg[3 * sg], T.seg[3 * sg + 1], T.seg[3 * sg + 2]}; },
                    [&](int r, double dot) { atomicAdd(w + phys(r) * ld + mode, -dot); }, T.val, T.idx, B.z, m, hmask, hbase);
                __threadfence();
                __syncthreads();
            }
        }
    }
}

struct TriHost {
    std::vector<int64_t> ptr;
    std::vector<int32_t> idx;
    std::vector<double> val, dinv;
};

struct BlockDesc {
    int32_t start, len;
    int64_t step;  // tasks with equal step are independent; steps are processed in ascending order
};

int upload_tri(asgfem_ctx* ctx, const TriHost& H, int64_t n, std::vector<BlockDesc> blocks, TriDev& D) {
    std::sort(blocks.begin(), blocks.end(), [](const BlockDesc& a, const BlockDesc& b) { return a.start < b.start; });
    const int nblocks = (int)blocks.size();
    std::vector<int32_t> starts((size_t)nblocks + 1), blk_of((size_t)n);
    for (int b = 0; b < nblocks; ++b) {
        starts[b] = blocks[b].start;
        for (int32_t k = blocks[b].start; k < blocks[b].start + blocks[b].len; ++k) blk_of[k] = b;
    }
    starts[nblocks] = (int32_t)n;
    std::vector<int32_t> drow((size_t)n + 1, 0), dsplit((size_t)n, 0), dcount((size_t)n, 0), segptr((size_t)nblocks + 1, 0), seg;
    std::vector<uint16_t> didx;
    std::vector<double> dval;
    for (int b = 0; b < nblocks; ++b) {
        const int64_t j0 = starts[b], j1 = starts[b + 1];
        for (int64_t k = j0; k < j1; ++k) {
            int64_t p = H.ptr[k + 1];
            while (p > H.ptr[k] && H.idx[p - 1] >= j0) --p;  // trailing entries inside the own block
            dcount[k] = (int32_t)(H.ptr[k + 1] - p);
            const int64_t panel0 = j0 + ((k - j0) / PANEL) * PANEL;
            int split = 0;
            for (int64_t q = p; q < H.ptr[k + 1]; ++q) {
                didx.push_back((uint16_t)(H.idx[q] - j0));
                dval.push_back(H.val[q]);
                if (H.idx[q] < panel0) ++split;
            }
            drow[k + 1] = (int32_t)didx.size();
            dsplit[k] = split;
        }
    }
    std::vector<int32_t> segent((size_t)nblocks + 1, 0), entfill;
    std::vector<uint16_t> pidx;
    std::vector<double> pval;
    for (int pass = 0; pass < 2; ++pass) {  // push segments: maximal runs of a row's entries inside one earlier block
        std::vector<int32_t> fill;
        if (pass == 1) {
            for (int b = 0; b < nblocks; ++b) {
                segptr[b + 1] += segptr[b];
                segent[b + 1] += segent[b];
            }
            seg.resize((size_t)3 * segptr[nblocks]);
            fill.assign(segptr.begin(), segptr.end() - 1);
            entfill.assign(segent.begin(), segent.end() - 1);
            pidx.resize((size_t)segent[nblocks]);
            pval.resize((size_t)segent[nblocks]);
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
                    segent[b + 1] += (int32_t)(q - p);
                } else {
                    int32_t at = fill[b]++;
                    int32_t eo = entfill[b];
                    entfill[b] += (int32_t)(q - p);
                    seg[3 * (size_t)at] = (int32_t)k;
                    seg[3 * (size_t)at + 1] = eo;  // offset into the block-major entry arrays
                    seg[3 * (size_t)at + 2] = (int32_t)(q - p);
                    for (int64_t e = p; e < q; ++e) {
                        pidx[(size_t)eo + (size_t)(e - p)] = (uint16_t)(H.idx[e] - starts[b]);
                        pval[(size_t)eo + (size_t)(e - p)] = H.val[e];
                    }
                }
                p = q;
            }
        }
    }
    for (int b = 0; b < nblocks; ++b) {  // partners in a warp get segments of similar length
        std::vector<std::array<int32_t, 3>> tmp;
        for (int32_t q = segptr[b]; q < segptr[b + 1]; ++q) tmp.push_back({seg[3 * (size_t)q], seg[3 * (size_t)q + 1], seg[3 * (size_t)q + 2]});
        std::stable_sort(tmp.begin(), tmp.end(), [](const std::array<int32_t, 3>& x, const std::array<int32_t, 3>& y) { return x[2] > y[2]; });
        for (size_t k = 0; k < tmp.size(); ++k)
            for (int c = 0; c < 3; ++c) seg[3 * ((size_t)segptr[b] + k) + c] = tmp[k][c];
    }
    // schedule: steps in ascending order; inside a step the small tasks first
    std::vector<int32_t> order((size_t)nblocks);
    for (int b = 0; b < nblocks; ++b) order[b] = b;
    auto small = [&](int b) { return blocks[b].len <= SMALL; };
    std::sort(order.begin(), order.end(), [&](int a, int b) {
        if (blocks[a].step != blocks[b].step) return blocks[a].step < blocks[b].step;
        if (small(a) != small(b)) return small(a);
        return a < b;
    });
    std::vector<int32_t> step_ptr(1, 0), step_nsmall;
    for (int k = 0; k < nblocks;) {
        int k2 = k, nsm = 0;
        while (k2 < nblocks && blocks[order[k2]].step == blocks[order[k]].step) {
            nsm += small(order[k2]);
            ++k2;
        }
        step_nsmall.push_back(nsm);
        step_ptr.push_back(k2);
        k = k2;
    }
    D.nblocks = nblocks;
    D.nsteps = (int)step_nsmall.size();
    int rc = 0;
    rc |= dev_upload(ctx, &D.idx, pidx);
    rc |= dev_upload(ctx, &D.val, pval);
    rc |= dev_upload(ctx, &D.segptr, segptr);
    rc |= dev_upload(ctx, &D.seg, seg);
    rc |= dev_upload(ctx, &D.dinv, H.dinv);
    rc |= dev_upload(ctx, &D.drow, drow);
    rc |= dev_upload(ctx, &D.dsplit, dsplit);
    rc |= dev_upload(ctx, &D.didx, didx);
    rc |= dev_upload(ctx, &D.dval, dval);
    rc |= dev_upload(ctx, &D.blk_start, starts);
    rc |= dev_upload(ctx, &D.step_ptr, step_ptr);
    rc |= dev_upload(ctx, &D.step_nsmall, step_nsmall);
    rc |= dev_upload(ctx, &D.tasks, order);
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
    // forward: deepest tree level first, chunks of a separator in order; backward (reversed numbering): root first
    std::vector<BlockDesc> fb, bb;
    for (const BlockRec& b : F.blocks) {
        fb.push_back({b.start, b.len, -(int64_t)b.depth * 4096 + b.chunk});
        bb.push_back({(int32_t)(n - b.start - b.len), b.len, (int64_t)b.depth * 4096 + (b.nchunks - 1 - b.chunk)});
    }
    rc |= upload_tri(ctx, fw, n, fb, P->fwd);
    rc |= upload_tri(ctx, bw, n, bb, P->bwd);
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
        ASG_CUDA(ctx, cudaFuncSetAttribute(k_trsv_tree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_trsv_tree<<<tiles, TRSV_THREADS, smem, ctx->stream>>>(P->d_work, ld, P->nred, 0, P->fwd);
        k_trsv_tree<<<tiles, TRSV_THREADS, smem, ctx->stream>>>(P->d_work, ld, P->nred, 1, P->bwd);
    }
    // z may alias r: boundary rows are zeroed first, interior rows are overwritten from the work vector
    k_zero_masked_rows<<<(unsigned)std::min<int64_t>(ctx->n, 148 * 8), 128, 0, ctx->stream>>>(z, ctx->d_bmask, ctx->n, ld);
    if (P->nred > 0) k_scatter_perm<<<blocks, 256, 0, ctx->stream>>>(P->d_work, z, P->d_perm, P->nred, ld);
    ASG_CUDA(ctx, cudaGetLastError());
    return 0;
}

}  // namespace asgfem
