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
//
// BOTTOM FOREST (k_small_sweep): the subtrees of the dissection tree that consist of blocks of <= 32 rows only
// (leaves and the small separators right above them: 98 % of the blocks, a third of the flops) are handled by a
// dense supernodal kernel instead.  The columns of one block share their row structure, so block T owns a DENSE panel
// P = L[A(T), T] (A(T): the rows below the block that it touches) and the explicit inverse of its diagonal triangle:
//   forward   z_T = inv(L_TT) w_T,  W[A(T)] -= P z_T      (one fp64 RED per target row and mode)
//   backward  y_T = inv(L_TT)^T (z_T - P^T y_A(T))         (pull: no atomics)
// A warp takes a block; its two half-warps hold the 16 modes of the tile, the block's <= 32 values per mode live in
// registers, L is read with warp-uniform 16-byte loads (no column indices, no shuffles, no shared memory).  The
// forest runs before (forward) / after (backward) the tree kernel above, which keeps the remaining top of the tree.
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

// One block of the bottom forest (32 bytes, read with two uniform 16-byte loads).  Its factor data is ONE contiguous
// record (16-byte aligned) so that a single bulk copy (TMA) moves a piece of it into shared memory:
//   [ inverse diagonal triangle, rows packed in pairs of equal even length ]
//   [ chunk 0 ][ chunk 1 ] ...     chunk = [ int32 target rows, padded to arB bytes ][ panel rows, even(w) doubles each ]
// every chunk holds R = small_chunk_rows(w, buffer size) panel rows (the last one fewer).
struct SmallBlk {
    int32_t j0, w, nA, pad0;
    int64_t rec_off, pad1;  // byte offset of the record
};
struct SmallDev {
    SmallBlk* blk = nullptr;
    int32_t* tasks = nullptr;     // block ids, launch after launch (forward order: deepest tree level first)
    unsigned char* rec = nullptr;  // records
    int nblocks = 0, nsteps = 0;
};

struct PrecondPlan {
    int64_t nred = 0;
    int32_t* d_perm = nullptr;
    TriDev fwd, bwd;
    SmallDev small;
    std::vector<std::array<int, 3>> small_launches;  // (t0, t1, width class), forward order
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
    {
        SmallDev& S = P->small;
        void* ptrs[] = {S.blk, S.tasks, S.rec};
        for (void* q : ptrs)
            if (q) cudaFree(q);
    }
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


// =====================================================================================================================
// bottom forest: dense supernodal warp tasks
// =====================================================================================================================
constexpr int SMALL_W = 32;        // widest block of the bottom forest
constexpr int SWEEP_THREADS = 512;
constexpr int SWEEP_SMEM = 224 * 1024;  // two staging buffers per warp

// offset of row r in the pair-packed inverse triangle: rows 2i and 2i+1 both hold 2i+2 entries
__host__ __device__ __forceinline__ int inv_row_off(int r) {
    const int i = r >> 1;
    return 2 * i * (i + 1) + (r & 1) * (2 * i + 2);
}
// record geometry for a block of width w staged through buffers of `buf` bytes
struct SmallGeom {
    int wp, invB, R, arB, chunkB;  // even(w), bytes of the inverse, rows per chunk, bytes of a chunk's row list, full chunk
};
__host__ __device__ __forceinline__ SmallGeom small_geom(int w, int buf) {
    SmallGeom g;
    g.wp = (w + 1) & ~1;
    g.invB = inv_row_off(g.wp) * 8;
    g.R = ((buf - 16) / (g.wp * 8 + 4)) & ~1;  // even, so that the pairing of rows over the half-warps stays aligned
    g.arB = (4 * g.R + 15) & ~15;
    g.chunkB = g.arB + g.R * g.wp * 8;
    return g;
}
__host__ __device__ __forceinline__ int small_buf_bytes(int wclass) { return wclass <= 24 ? SWEEP_SMEM / (2 * 16) : SWEEP_SMEM / (2 * 8); }

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ SmallBlk load_blk(const SmallDev& S, int t) {
    const int4* bp = reinterpret_cast<const int4*>(S.blk + S.tasks[t]);
    const int4 q0 = __ldg(bp), q1 = __ldg(bp + 1);
    SmallBlk b;
    b.j0 = q0.x, b.w = q0.y, b.nA = q0.z, b.pad0 = 0;
    b.rec_off = ((int64_t)(uint32_t)q1.y << 32) | (uint32_t)q1.x;
    b.pad1 = 0;
    return b;
}

// Per-warp two-stage pipeline of bulk copies (cp.async.bulk + mbarrier): while the warp works on one piece of a record,
// the next piece - of the same task or of the warp's next task - is already on its way into the other buffer.  The factor
// data is streamed from HBM exactly once per CTA with no reuse; without staging every warp-uniform load paid the full
// memory latency in a dependent chain (profiles/r01_small_*).
template <bool BWD>
struct SmallPipe {
    const SmallDev& S;
    unsigned char* buf;   // this warp's two buffers
    unsigned bar32;       // shared address of this warp's two mbarriers
    int bufB, t1, stride, lane;
    // producer position: piece `pc` of task `pt` (forward: -1 = inverse, then chunks 0..; backward: chunks, then -1)
    int pt, pc, pnch;
    SmallBlk pb;
    SmallGeom pg;
    unsigned k = 0, phase = 0;  // pieces acquired so far, parity bits of the two barriers

    __device__ __forceinline__ void set_task(int t) {
        pt = t;
        if (t < t1) {
            pb = load_blk(S, t);
            pg = small_geom(pb.w, bufB);
            pnch = (pb.nA + pg.R - 1) / pg.R;
            pc = BWD ? (pnch > 0 ? 0 : -1) : -1;
        }
    }
    __device__ __forceinline__ void advance() {
        if (BWD) {
            if (pc == -1)
                set_task(pt + stride);
            else
                pc = pc + 1 < pnch ? pc + 1 : -1;
        } else {
            if (pc + 1 < pnch)
                ++pc;
            else
                set_task(pt + stride);
        }
    }
    __device__ __forceinline__ void issue(unsigned stage) {  // current producer piece -> buffer `stage`
        if (pt >= t1) return;
        const unsigned char* src = S.rec + pb.rec_off;
        int bytes;
        if (pc < 0) {
            bytes = pg.invB;
        } else {
            src += pg.invB + (int64_t)pc * pg.chunkB;
            const int r = min(pg.R, pb.nA - pc * pg.R);
            bytes = pg.arB + r * pg.wp * 8;
        }
        if (lane == 0) {
            const unsigned bar = bar32 + 8u * stage;
            const unsigned dst = (unsigned)__cvta_generic_to_shared(buf + (size_t)stage * bufB);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                         "l"(src), "r"(bytes), "r"(bar)
                         : "memory");
        }
        advance();
    }
    __device__ __forceinline__ void start(int t0w) {
        set_task(t0w);
        issue(0);
    }
    // waits for the next piece in sequence and returns its buffer; the piece after it is put in flight first
    __device__ __forceinline__ const unsigned char* acquire() {
        const unsigned stage = k & 1u;
        __syncwarp();  // every lane is done with the other buffer
        issue(stage ^ 1u);
        const unsigned bar = bar32 + 8u * stage, par = (phase >> stage) & 1u;
        asm volatile(
            "{\n\t.reg .pred P1;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar),
            "r"(par)
            : "memory");
        phase ^= 1u << stage;
        ++k;
        return buf + (size_t)stage * bufB;
    }
};

__device__ __forceinline__ double2 lds_d2(const double* p) { return *reinterpret_cast<const double2*>(p); }

// forward task: z_T = inv(L_TT) w_T, then W[A(T)] -= P z_T
template <int WMAX>
__device__ __forceinline__ void small_fwd(const SmallBlk b, SmallPipe<false>& pipe, double* __restrict__ W, int64_t ld, int64_t mode,
                                          int half) {
    const int w = b.w;
    double* wt = W + (int64_t)b.j0 * ld + mode;
    double v[WMAX];
#pragma unroll
    for (int c = 0; c < WMAX; ++c) v[c] = c < w ? __ldcg(wt + (int64_t)c * ld) : 0.0;
    const double* inv = reinterpret_cast<const double*>(pipe.acquire());
    // rows from the bottom up: row r needs v[c <= r] only, so z overwrites v in place (one register array)
#pragma unroll
    for (int i = WMAX / 2 - 1; i >= 0; --i) {
        if (2 * i < w) {  // warp-uniform
            const double* row = inv + 2 * i * (i + 1) + half * (2 * i + 2);  // half h computes row 2i + h
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int q = 0; q <= i; ++q) {
                const double2 l = lds_d2(row + 2 * q);
                a0 = fma(l.x, v[2 * q], a0);
                a1 = fma(l.y, v[2 * q + 1], a1);
            }
            const double mine = a0 + a1;
            const double other = __shfl_xor_sync(0xffffffffu, mine, 16);
            v[2 * i] = half ? other : mine;
            v[2 * i + 1] = half ? mine : other;
            if (2 * i + half < w) wt[(int64_t)(2 * i + half) * ld] = mine;
        }
    }
    double(&z)[WMAX] = v;
    const SmallGeom g = small_geom(w, pipe.bufB);
    for (int a0r = 0; a0r < b.nA; a0r += g.R) {
        const unsigned char* ch = pipe.acquire();
        const int32_t* ar = reinterpret_cast<const int32_t*>(ch);
        const double* P = reinterpret_cast<const double*>(ch + g.arB);
        const int nr = min(g.R, b.nA - a0r);
        for (int p = 0; p < nr; p += 2) {
            const bool ok = p + half < nr;
            const int a = ok ? p + half : nr - 1;
            const int64_t r = ar[a];
            const double* prow = P + a * g.wp;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
            for (int q = 0; q < WMAX / 2; q += 2) {
                if (2 * q < w) {
                    const double2 l = lds_d2(prow + 2 * q);
                    s0 = fma(l.x, z[2 * q], s0);
                    s1 = fma(l.y, z[2 * q + 1], s1);
                }
                if (2 * q + 2 < w) {
                    const double2 l = lds_d2(prow + 2 * q + 2);
                    s2 = fma(l.x, z[2 * q + 2], s2);
                    s3 = fma(l.y, z[2 * q + 3], s3);
                }
            }
            if (ok) atomicAdd(W + r * ld + mode, -((s0 + s1) + (s2 + s3)));  // sibling blocks share ancestor rows
        }
    }
}

// backward task: y_T = inv(L_TT)^T (z_T - P^T y_A(T))
template <int WMAX>
__device__ __forceinline__ void small_bwd(const SmallBlk b, SmallPipe<true>& pipe, double* __restrict__ W, int64_t ld, int64_t mode,
                                          int half) {
    const int w = b.w;
    double* wt = W + (int64_t)b.j0 * ld + mode;
    double acc[WMAX];
#pragma unroll
    for (int c = 0; c < WMAX; ++c) acc[c] = 0.0;
    const SmallGeom g = small_geom(w, pipe.bufB);
    for (int a0r = 0; a0r < b.nA; a0r += g.R) {
        const unsigned char* ch = pipe.acquire();
        const int32_t* ar = reinterpret_cast<const int32_t*>(ch);
        const double* P = reinterpret_cast<const double*>(ch + g.arB);
        const int nr = min(g.R, b.nA - a0r);
        for (int a = pipe.lane; a < nr; a += 32) prefetch_l2(W + (int64_t)ar[a] * ld + (mode & ~(int64_t)(MT - 1)));
        for (int p = 0; p < nr; p += 2) {
            const bool ok = p + half < nr;
            const int a = ok ? p + half : nr - 1;
            const int64_t r = ar[a];
            const double y = ok ? __ldcg(W + r * ld + mode) : 0.0;
            const double* prow = P + a * g.wp;
#pragma unroll
            for (int q = 0; q < WMAX / 2; ++q) {
                if (2 * q < w) {
                    const double2 l = lds_d2(prow + 2 * q);
                    acc[2 * q] = fma(l.x, y, acc[2 * q]);
                    acc[2 * q + 1] = fma(l.y, y, acc[2 * q + 1]);
                }
            }
        }
    }
    double(&t)[WMAX] = acc;  // t = z_T - P^T y_A, in place
#pragma unroll
    for (int c = 0; c < WMAX; ++c) {
        const double own = c < w ? __ldcg(wt + (int64_t)c * ld) : 0.0;
        t[c] = own - (acc[c] + __shfl_xor_sync(0xffffffffu, acc[c], 16));
    }
    const double* inv = reinterpret_cast<const double*>(pipe.acquire());
#pragma unroll
    for (int i = 0; i < WMAX / 2; ++i) {
        if (2 * i < w) {
            // y_r = sum_{c >= r} inv[c][r] t[c], r = 2i + half; inv[2i][2i+1] is a stored zero
            const double* col = inv + 2 * i + half;
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int j = i; j < WMAX / 2; ++j) {
                if (2 * j < w) {
                    a0 = fma(col[2 * j * (j + 1)], t[2 * j], a0);
                    a1 = fma(col[2 * j * (j + 1) + 2 * j + 2], t[2 * j + 1], a1);
                }
            }
            if (2 * i + half < w) wt[(int64_t)(2 * i + half) * ld] = a0 + a1;
        }
    }
}

// one launch per (tree level, width class): tasks [t0, t1) are independent; the launch boundary orders the levels
template <int WMAX, bool BWD>
__global__ void __launch_bounds__(WMAX <= 24 ? SWEEP_THREADS : SWEEP_THREADS / 2, 1)
k_small_step(double* __restrict__ W, int64_t ld, SmallDev S, int t0, int t1) {
    extern __shared__ __align__(128) unsigned char sweep_smem[];
    __shared__ __align__(8) unsigned long long sweep_bars[2 * (SWEEP_THREADS / 32)];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int half = lane >> 4;
    const int64_t mode0 = (int64_t)blockIdx.x * MT, mode = mode0 + (lane & (MT - 1));
    const int bufB = small_buf_bytes(WMAX);
    if (lane == 0) {
        const unsigned bar = (unsigned)__cvta_generic_to_shared(sweep_bars + 2 * warp);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar + 8u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    SmallPipe<BWD> pipe{S, sweep_smem + (size_t)warp * 2 * bufB, (unsigned)__cvta_generic_to_shared(sweep_bars + 2 * warp),
                        bufB,  t1, nwarps, lane};
    pipe.start(t0 + warp);
    for (int t = t0 + warp; t < t1; t += nwarps) {
        const SmallBlk b = load_blk(S, t);
        if (t + nwarps < t1) {  // rows of W of the warp's next task -> L2
            const SmallBlk nb = load_blk(S, t + nwarps);
            if (lane < nb.w) prefetch_l2(W + (int64_t)(nb.j0 + lane) * ld + mode0);
        }
        if constexpr (BWD)
            small_bwd<WMAX>(b, pipe, W, ld, mode, half);
        else
            small_fwd<WMAX>(b, pipe, W, ld, mode, half);
    }
}

struct SmallLaunch {
    int t0, t1, wmax;
};

template <int WMAX, bool BWD>
static void launch_small_k(const SmallLaunch& L, int tiles, cudaStream_t st, double* W, int64_t ld, const SmallDev& S) {
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k_small_step<WMAX, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SWEEP_SMEM);
        configured = true;
    }
    k_small_step<WMAX, BWD><<<tiles, WMAX <= 24 ? SWEEP_THREADS : SWEEP_THREADS / 2, SWEEP_SMEM, st>>>(W, ld, S, L.t0, L.t1);
}

template <bool BWD>
static void launch_small(const SmallLaunch& L, int tiles, cudaStream_t st, double* W, int64_t ld, const SmallDev& S) {
    switch (L.wmax) {
        case 8: launch_small_k<8, BWD>(L, tiles, st, W, ld, S); break;
        case 16: launch_small_k<16, BWD>(L, tiles, st, W, ld, S); break;
        case 24: launch_small_k<24, BWD>(L, tiles, st, W, ld, S); break;
        default: launch_small_k<32, BWD>(L, tiles, st, W, ld, S); break;
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

// `top[k] != 0`: row k (numbering of this system) belongs to the part of the tree that this kernel handles; blocks and
// push segments that involve other rows are left out (they belong to the bottom forest)
int upload_tri(asgfem_ctx* ctx, const TriHost& H, int64_t n, std::vector<BlockDesc> blocks, const std::vector<uint8_t>& top,
               TriDev& D) {
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
                if (!top[(size_t)k] || !top[(size_t)starts[b]]) {
                    p = q;
                    continue;
                }
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
    std::vector<int32_t> order;
    for (int b = 0; b < nblocks; ++b)
        if (top[(size_t)starts[b]]) order.push_back(b);
    const int ntasks = (int)order.size();
    auto small = [&](int b) { return blocks[b].len <= SMALL; };
    std::sort(order.begin(), order.end(), [&](int a, int b) {
        if (blocks[a].step != blocks[b].step) return blocks[a].step < blocks[b].step;
        if (small(a) != small(b)) return small(a);
        return a < b;
    });
    std::vector<int32_t> step_ptr(1, 0), step_nsmall;
    for (int k = 0; k < ntasks;) {
        int k2 = k, nsm = 0;
        while (k2 < ntasks && blocks[order[k2]].step == blocks[order[k]].step) {
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


// Splits the dissection tree into the bottom forest (subtrees made of blocks of <= SMALL_W rows) and the top part, and
// builds the dense panels / inverse triangles of the forest.  Lp/Li/Lx: rows of L; cptr/cidx/cval: columns of L with
// ascending rows.  top[k] = 1 for rows that stay with the tree kernel.
int build_small(asgfem_ctx* ctx, const CholFactor& F, const std::vector<int64_t>& cptr, const std::vector<int32_t>& cidx,
                const std::vector<double>& cval, std::vector<uint8_t>& top, SmallDev& D,
                std::vector<std::array<int, 3>>& launches) {
    const int64_t n = F.n;
    const int nb = (int)F.blocks.size();
    std::vector<int32_t> blk_of((size_t)n);
    for (int b = 0; b < nb; ++b)
        for (int32_t k = 0; k < F.blocks[b].len; ++k) blk_of[(size_t)F.blocks[b].start + k] = b;
    // a block stays on top if it is wide or if a top block pushes into it (its sources must be done before it)
    std::vector<uint8_t> btop((size_t)nb, 0);
    for (int b = 0; b < nb; ++b) {  // blocks are sorted by start; targets always come later
        if (F.blocks[b].len > SMALL_W) btop[b] = 1;
        if (!btop[b]) continue;
        const int64_t j0 = F.blocks[b].start, j1 = j0 + F.blocks[b].len;
        for (int64_t c = j0; c < j1; ++c)
            for (int64_t p = cptr[c]; p < cptr[c + 1]; ++p)
                if (cidx[p] >= j1) btop[blk_of[(size_t)cidx[p]]] = 1;
    }
    top.assign((size_t)n, 0);
    for (int64_t k = 0; k < n; ++k) top[(size_t)k] = btop[blk_of[(size_t)k]];

    std::vector<SmallBlk> blk;
    std::vector<int32_t> depth_of;
    std::vector<unsigned char> rec;
    std::vector<int32_t> mark((size_t)n, -1), list;
    std::vector<double> Ld, X;
    for (int b = 0; b < nb; ++b) {
        if (btop[b]) continue;
        const int32_t j0 = F.blocks[b].start, w = F.blocks[b].len;
        const int wcl = w <= 8 ? 8 : (w <= 16 ? 16 : (w <= 24 ? 24 : 32));
        const SmallGeom g = small_geom(w, small_buf_bytes(wcl));
        const int wp = g.wp;
        SmallBlk sb;
        sb.j0 = j0;
        sb.w = w;
        sb.pad0 = 0;
        sb.pad1 = 0;
        // target rows A(T)
        list.clear();
        for (int32_t c = j0; c < j0 + w; ++c)
            for (int64_t p = cptr[c]; p < cptr[c + 1]; ++p) {
                const int32_t i = cidx[p];
                if (i >= j0 + w && mark[(size_t)i] != b) {
                    mark[(size_t)i] = b;
                    list.push_back(i);
                }
            }
        std::sort(list.begin(), list.end());
        const int nA = (int)list.size(), nch = (nA + g.R - 1) / g.R;
        sb.nA = nA;
        sb.rec_off = (int64_t)rec.size();
        size_t bytes = (size_t)g.invB;
        for (int c = 0; c < nch; ++c) bytes += (size_t)g.arB + (size_t)std::min(g.R, nA - c * g.R) * wp * 8;
        // intermediate chunks are full, so chunk c starts at invB + c * chunkB
        rec.resize(rec.size() + bytes, 0);
        unsigned char* base = rec.data() + sb.rec_off;
        double* inv = reinterpret_cast<double*>(base);
        for (size_t a = 0; a < list.size(); ++a) mark[(size_t)list[a]] = (int32_t)a;  // position in the panel
        for (int a = 0; a < nA; ++a) {
            unsigned char* ch = base + g.invB + (size_t)(a / g.R) * g.chunkB;
            reinterpret_cast<int32_t*>(ch)[a % g.R] = list[(size_t)a];
        }
        for (int32_t c = j0; c < j0 + w; ++c)
            for (int64_t p = cptr[c]; p < cptr[c + 1]; ++p) {
                const int32_t i = cidx[p];
                if (i < j0 + w) continue;
                const int a = mark[(size_t)i];
                unsigned char* ch = base + g.invB + (size_t)(a / g.R) * g.chunkB;
                reinterpret_cast<double*>(ch + g.arB)[(size_t)(a % g.R) * wp + (size_t)(c - j0)] = cval[p];
            }
        for (int32_t i : list) mark[(size_t)i] = -1;
        // inverse of the diagonal triangle
        Ld.assign((size_t)w * w, 0.0);
        for (int32_t r = 0; r < w; ++r) {
            Ld[(size_t)r * w + r] = 1.0 / F.dinv[(size_t)j0 + r];
            for (int64_t p = F.Lp[(size_t)j0 + r]; p < F.Lp[(size_t)j0 + r + 1]; ++p)
                if (F.Li[p] >= j0) Ld[(size_t)r * w + (F.Li[p] - j0)] = F.Lx[p];
        }
        X.assign((size_t)w * w, 0.0);
        for (int32_t j = 0; j < w; ++j) {
            X[(size_t)j * w + j] = 1.0 / Ld[(size_t)j * w + j];
            for (int32_t r = j + 1; r < w; ++r) {
                double sum = 0.0;
                for (int32_t k = j; k < r; ++k) sum += Ld[(size_t)r * w + k] * X[(size_t)k * w + j];
                X[(size_t)r * w + j] = -sum / Ld[(size_t)r * w + r];
            }
        }
        for (int32_t r = 0; r < w; ++r)
            for (int32_t c = 0; c <= r; ++c) inv[(size_t)inv_row_off(r) + c] = X[(size_t)r * w + c];
        blk.push_back(sb);
        depth_of.push_back(F.blocks[b].depth);
    }
    D.nblocks = (int)blk.size();
    // launches: deepest level first, inside a level one launch per width class
    auto wclass = [&](int k) { return blk[k].w <= 8 ? 8 : (blk[k].w <= 16 ? 16 : (blk[k].w <= 24 ? 24 : 32)); };
    std::vector<int32_t> order((size_t)D.nblocks);
    for (int k = 0; k < D.nblocks; ++k) order[k] = k;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        return depth_of[a] != depth_of[b] ? depth_of[a] > depth_of[b] : wclass(a) < wclass(b);
    });
    launches.clear();
    for (int k = 0, k0 = 0; k < D.nblocks; ++k)
        if (k + 1 == D.nblocks || depth_of[order[k + 1]] != depth_of[order[k]] || wclass(order[k + 1]) != wclass(order[k])) {
            launches.push_back({k0, k + 1, wclass(order[k])});
            k0 = k + 1;
        }
    D.nsteps = (int)launches.size();
    int rc = 0;
    rc |= dev_upload(ctx, &D.blk, blk);
    rc |= dev_upload(ctx, &D.tasks, order);
    rc |= dev_upload(ctx, &D.rec, rec);
    if (rc) return rc;
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the host vectors go out of scope
    return 0;
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
    // columns of L (rows ascending inside every column)
    std::vector<int64_t> cptr((size_t)n + 1, 0);
    for (int32_t j : F.Li) cptr[j + 1]++;
    for (int64_t k = 0; k < n; ++k) cptr[k + 1] += cptr[k];
    std::vector<int32_t> cidx(F.Li.size());
    std::vector<double> cval(F.Li.size());
    {
        std::vector<int64_t> fill(cptr.begin(), cptr.end() - 1);
        for (int64_t k = 0; k < n; ++k)
            for (int64_t p = F.Lp[k]; p < F.Lp[k + 1]; ++p) {
                int64_t at = fill[F.Li[p]]++;
                cidx[at] = (int32_t)k;
                cval[at] = F.Lx[p];
            }
    }
    // bottom forest -> dense supernodal kernel; the rest of the tree -> tree kernel
    std::vector<uint8_t> top_f, top_b;
    rc = build_small(ctx, F, cptr, cidx, cval, top_f, P->small, P->small_launches);
    if (rc) return rc;
    top_b.assign(top_f.rbegin(), top_f.rend());
    // backward system L^T in the reversed numbering k' = n-1-k: row k' holds the column k of L, rows i > k mapped to
    // i' = n-1-i < k' in ascending order
    {
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
    rc |= upload_tri(ctx, fw, n, fb, top_f, P->fwd);
    rc |= upload_tri(ctx, bw, n, bb, top_b, P->bwd);
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
        const bool has_small = P->small.nblocks > 0, has_top = P->fwd.nsteps > 0;
        if (has_small)
            for (size_t k = 0; k < P->small_launches.size(); ++k) {
                const auto& L = P->small_launches[k];
                launch_small<false>({L[0], L[1], L[2]}, tiles, ctx->stream, P->d_work, ld, P->small);
            }
        if (has_top) {
            k_trsv_tree<<<tiles, TRSV_THREADS, smem, ctx->stream>>>(P->d_work, ld, P->nred, 0, P->fwd);
            k_trsv_tree<<<tiles, TRSV_THREADS, smem, ctx->stream>>>(P->d_work, ld, P->nred, 1, P->bwd);
        }
        if (has_small)
            for (size_t k = P->small_launches.size(); k-- > 0;) {
                const auto& L = P->small_launches[k];
                launch_small<true>({L[0], L[1], L[2]}, tiles, ctx->stream, P->d_work, ld, P->small);
            }
    }
    // z may alias r: boundary rows are zeroed first, interior rows are overwritten from the work vector
    k_zero_masked_rows<<<(unsigned)std::min<int64_t>(ctx->n, 148 * 8), 128, 0, ctx->stream>>>(z, ctx->d_bmask, ctx->n, ld);
    if (P->nred > 0) k_scatter_perm<<<blocks, 256, 0, ctx->stream>>>(P->d_work, z, P->d_perm, P->nred, ld);
    ASG_CUDA(ctx, cudaGetLastError());
    return 0;
}

}  // namespace asgfem
