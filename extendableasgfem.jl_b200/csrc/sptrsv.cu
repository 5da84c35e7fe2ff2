// Mean-based preconditioner  Z[:,mu] = K_0^{-1} R[:,mu]  for all modes at once
// (LinearAlgebra.ldiv!(y, P::MyPreconditionerPrimal, b), src/modelproblems/solvers_poisson_primal.jl:46-78;
//  the reference runs N sequential UMFPACK solves, one per mode block).
//
// Factor P K_0 P^T = L L^T from chol.cpp (nested dissection: leaves and separators of the dissection tree are the
// blocks).  The N right-hand sides are independent, so the triangular sweeps need no synchronisation between CTAs:
// every CTA owns a tile of MT = 16 modes and walks the whole tree.  Work vectors live in elimination order
// (W[k,:] <-> dof perm[k]).
//
// DENSE SUPERNODAL WARP TASKS.  The columns of one block share their row structure, so every block is cut into
// sub-blocks T of <= 32 columns, each with
//   * the explicit inverse of its diagonal triangle inv(L_TT) (computed on the host), and
//   * a DENSE panel P = L[A(T), T], A(T) = the rows below T that it touches (rest of its block + ancestor separators;
//     measured fill-in of the dense storage: 1.00 for separators, 1.76 for the 24-node leaves).
//   forward   z_T = inv(L_TT) w_T,  W[A(T)] -= P z_T        (one fp64 RED per target row and mode: sibling subtrees
//                                                            update the same ancestor rows concurrently)
//   backward  y_T = inv(L_TT)^T (z_T - P^T y_A(T))           (pull: no atomics)
// A warp takes a task; its two half-warps hold the 16 modes of the tile (lane = mode), the <= 32 values of the
// sub-block per mode live in registers, and the factor data - read exactly once per CTA, no reuse - is streamed
// through a per-warp two-stage pipeline of bulk copies (cp.async.bulk + mbarrier) into shared memory, from where it is
// read with half-warp-uniform 16-byte loads: no column indices, no shuffles in the inner loops.
// Sub-blocks with many panel rows (the separators near the root) are split into a solve-only task and push-only tasks
// of <= 64 rows, so that the few large separators at the top of the tree keep all warps of the CTA busy.
// Launches are ordered by (tree depth descending, separator chunk, sub-block, phase): everything inside one launch is
// independent, the launch boundary is the only synchronisation.  The backward sweep runs the launches in reverse.
#include <algorithm>
#include <array>

#include "common.h"

namespace asgfem {

constexpr int MT = 16;           // modes per CTA (half a warp wide)
constexpr int SMALL_W = 32;      // columns of a sub-block
constexpr int SPLIT_ROWS = 64;   // panel rows of a push-only task; panels up to this size stay fused with their solve
constexpr int SWEEP_THREADS = 512;
constexpr int SWEEP_SMEM = 224 * 1024;  // two staging buffers per warp

enum : int32_t { TASK_FUSED = 0, TASK_SOLVE_ONLY = 1, TASK_PUSH_ONLY = 2 };

// One task (32 bytes, read with two uniform 16-byte loads).  Its factor data is ONE contiguous record (16-byte aligned)
// so that a single bulk copy moves a piece of it into shared memory:
//   [ inverse diagonal triangle, rows packed in pairs of equal even length ]      (absent for push-only tasks)
//   [ chunk 0 ][ chunk 1 ] ...     chunk = [ int32 target rows, padded to arB bytes ][ panel rows, even(w) doubles each ]
// every chunk holds R = small_geom(w, buffer size).R panel rows (the last one fewer).
struct SmallBlk {
    int32_t j0, w, nA, kind;
    int64_t rec_off, pad1;  // byte offset of the record
};
struct SmallDev {
    SmallBlk* blk = nullptr;
    unsigned char* rec = nullptr;
    int nblocks = 0;
};

struct PrecondPlan {
    int64_t nred = 0;
    int32_t* d_perm = nullptr;
    SmallDev small;
    std::vector<std::array<int, 3>> launches;  // (t0, t1, width class), forward order
    double* d_work = nullptr;  // nred x work_ld, (re)allocated by precond_apply when the column count of the vectors changes
    int64_t work_ld = 0;
    int64_t lnz = 0;
};

void precond_free_plan(PrecondPlan* P) {
    if (!P) return;
    void* ptrs[] = {P->small.blk, P->small.rec, P->d_perm, P->d_work};
    for (void* q : ptrs)
        if (q) cudaFree(q);
    delete P;
}

void precond_free(asgfem_ctx* ctx) {
    precond_free_plan(ctx->precond);
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

// offset of row r in the pair-packed inverse triangle: rows 2i and 2i+1 both hold 2i+2 entries
__host__ __device__ __forceinline__ int inv_row_off(int r) {
    const int i = r >> 1;
    return 2 * i * (i + 1) + (r & 1) * (2 * i + 2);
}
// record geometry for a sub-block of width w staged through buffers of `buf` bytes
struct SmallGeom {
    int wp, invB, R, arB, chunkB;  // even(w), bytes of the inverse, rows per chunk, bytes of a chunk's row list, full chunk
};
__host__ __device__ __forceinline__ SmallGeom small_geom(int w, int buf, int kind) {
    SmallGeom g;
    g.wp = (w + 1) & ~1;
    g.invB = kind == TASK_PUSH_ONLY ? 0 : inv_row_off(g.wp) * 8;
    g.R = ((buf - 16) / (g.wp * 8 + 4)) & ~1;
    g.arB = (4 * g.R + 15) & ~15;
    g.chunkB = g.arB + g.R * g.wp * 8;
    return g;
}
__host__ __device__ __forceinline__ int small_buf_bytes(int wclass) { return wclass <= 24 ? SWEEP_SMEM / (2 * 16) : SWEEP_SMEM / (2 * 8); }

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ SmallBlk load_blk(const SmallDev& S, int t) {
    const int4* bp = reinterpret_cast<const int4*>(S.blk + t);
    const int4 q0 = __ldg(bp), q1 = __ldg(bp + 1);
    SmallBlk b;
    b.j0 = q0.x, b.w = q0.y, b.nA = q0.z, b.kind = q0.w;
    b.rec_off = ((int64_t)(uint32_t)q1.y << 32) | (uint32_t)q1.x;
    b.pad1 = 0;
    return b;
}

// Per-warp two-stage pipeline of bulk copies: while the warp works on one piece of a record, the next piece - of the
// same task or of the warp's next task - is already on its way into the other buffer.  Without staging every
// warp-uniform load paid the full memory latency in a dependent chain (profiles/README.md, triangular solves).
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
            pg = small_geom(pb.w, bufB, pb.kind);
            pnch = (pb.nA + pg.R - 1) / pg.R;
            const bool hdr = pb.kind != TASK_PUSH_ONLY;
            pc = BWD ? (pnch > 0 ? 0 : -1) : (hdr ? -1 : 0);
        }
    }
    __device__ __forceinline__ void advance() {
        const bool hdr = pb.kind != TASK_PUSH_ONLY;
        if (BWD) {
            if (pc == -1)
                set_task(pt + stride);
            else if (pc + 1 < pnch)
                ++pc;
            else if (hdr)
                pc = -1;
            else
                set_task(pt + stride);
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

// forward task: z_T = inv(L_TT) w_T (unless push-only: W already holds z_T), then W[A(T)] -= P z_T
template <int WMAX>
__device__ __forceinline__ void small_fwd(const SmallBlk b, SmallPipe<false>& pipe, double* __restrict__ W, int64_t ld, int64_t mode,
                                          int half) {
    const int w = b.w;
    double* wt = W + (int64_t)b.j0 * ld + mode;
    double v[WMAX];
#pragma unroll
    for (int c = 0; c < WMAX; ++c) v[c] = c < w ? __ldcg(wt + (int64_t)c * ld) : 0.0;
    if (b.kind != TASK_PUSH_ONLY) {
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
    }
    double(&z)[WMAX] = v;
    const SmallGeom g = small_geom(w, pipe.bufB, b.kind);
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
            if (ok) atomicAdd(W + r * ld + mode, -((s0 + s1) + (s2 + s3)));  // other tasks share these target rows
        }
    }
}

// backward task: y_T = inv(L_TT)^T (z_T - P^T y_A(T)); a pull-only task (part of the panel rows of a split sub-block)
// subtracts its share of P^T y_A from z_T in memory and leaves the solve to the solve-only task of the next launch
template <int WMAX>
__device__ __forceinline__ void small_bwd(const SmallBlk b, SmallPipe<true>& pipe, double* __restrict__ W, int64_t ld, int64_t mode,
                                          int half) {
    const int w = b.w;
    double* wt = W + (int64_t)b.j0 * ld + mode;
    double acc[WMAX];
#pragma unroll
    for (int c = 0; c < WMAX; ++c) acc[c] = 0.0;
    const SmallGeom g = small_geom(w, pipe.bufB, b.kind);
    for (int a0r = 0; a0r < b.nA; a0r += g.R) {
        const unsigned char* ch = pipe.acquire();
        const int32_t* ar = reinterpret_cast<const int32_t*>(ch);
        const double* P = reinterpret_cast<const double*>(ch + g.arB);
        const int nr = min(g.R, b.nA - a0r);
        // y values two pairs ahead of their use: the gathers from W are the only global loads of this loop
        auto load_y = [&](int p) {
            const int a = p + half;
            return a < nr ? __ldcg(W + (int64_t)ar[a] * ld + mode) : 0.0;
        };
        double y0 = load_y(0), y1 = load_y(2);
        for (int p = 0; p < nr; p += 2) {
            const double y = y0;
            y0 = y1;
            y1 = load_y(p + 4);
            const int a = min(p + half, nr - 1);
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
    if (b.kind == TASK_PUSH_ONLY) {
#pragma unroll
        for (int i = 0; i < WMAX / 2; ++i) {
            if (2 * i < w) {
                const double s0 = acc[2 * i] + __shfl_xor_sync(0xffffffffu, acc[2 * i], 16);
                const double s1 = acc[2 * i + 1] + __shfl_xor_sync(0xffffffffu, acc[2 * i + 1], 16);
                if (2 * i + half < w) atomicAdd(wt + (int64_t)(2 * i + half) * ld, -(half ? s1 : s0));
            }
        }
        return;
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

// one launch per (tree level, chunk, sub-block, phase, width class): tasks [t0, t1) are independent
template <int WMAX, bool BWD>
__global__ void __launch_bounds__(WMAX <= 24 ? SWEEP_THREADS : SWEEP_THREADS / 2, 1)
k_small_step(double* __restrict__ W, int64_t ld, SmallDev S, int t0, int t1) {
    extern __shared__ __align__(128) unsigned char sweep_smem[];
    __shared__ __align__(8) unsigned long long sweep_bars[2 * (SWEEP_THREADS / 32)];
    // the tasks of a launch are independent: gridDim.y CTAs share them (mode shards of a multi-GPU solve have few mode
    // tiles; without the split only ld / 16 SMs would work)
    const int lane = threadIdx.x & 31, warp = (threadIdx.x >> 5) + blockIdx.y * (blockDim.x >> 5), nwarps = (blockDim.x >> 5) * gridDim.y;
    const int lwarp = threadIdx.x >> 5;
    if (t0 + warp >= t1) return;
    const int half = lane >> 4;
    const int64_t mode0 = (int64_t)blockIdx.x * MT, mode = mode0 + (lane & (MT - 1));
    const int bufB = small_buf_bytes(WMAX);
    if (lane == 0) {
        const unsigned bar = (unsigned)__cvta_generic_to_shared(sweep_bars + 2 * lwarp);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar + 8u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    SmallPipe<BWD> pipe{S, sweep_smem + (size_t)lwarp * 2 * bufB, (unsigned)__cvta_generic_to_shared(sweep_bars + 2 * lwarp),
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

template <int WMAX, bool BWD>
void launch_small_k(int t0, int t1, int tiles, int split, cudaStream_t st, double* W, int64_t ld, const SmallDev& S) {
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k_small_step<WMAX, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SWEEP_SMEM);
        configured = true;
    }
    constexpr int threads = WMAX <= 24 ? SWEEP_THREADS : SWEEP_THREADS / 2;
    const int ntask = t1 - t0, per = threads / 32;
    const int gy = std::max(1, std::min(split, (ntask + per - 1) / per));  // no CTAs without a task
    k_small_step<WMAX, BWD><<<dim3((unsigned)tiles, (unsigned)gy), threads, SWEEP_SMEM, st>>>(W, ld, S, t0, t1);
}

template <bool BWD>
void launch_small(const std::array<int, 3>& L, int tiles, int split, cudaStream_t st, double* W, int64_t ld, const SmallDev& S) {
    switch (L[2]) {
        case 8: launch_small_k<8, BWD>(L[0], L[1], tiles, split, st, W, ld, S); break;
        case 16: launch_small_k<16, BWD>(L[0], L[1], tiles, split, st, W, ld, S); break;
        case 24: launch_small_k<24, BWD>(L[0], L[1], tiles, split, st, W, ld, S); break;
        default: launch_small_k<32, BWD>(L[0], L[1], tiles, split, st, W, ld, S); break;
    }
}

// Cuts the blocks of the dissection tree into sub-blocks of <= SMALL_W columns and builds their tasks (records with the
// inverse triangle and the dense panel) in launch order.  Lp/Li/Lx: rows of L; cptr/cidx/cval: columns of L with
// ascending rows.
int build_tasks(asgfem_ctx* ctx, const CholFactor& F, const std::vector<int64_t>& cptr, const std::vector<int32_t>& cidx,
                const std::vector<double>& cval, SmallDev& D, std::vector<std::array<int, 3>>& launches) {
    const int64_t n = F.n;
    struct Task {
        SmallBlk blk;
        int64_t key;  // launch order
        int wclass;
    };
    std::vector<Task> tasks;
    std::vector<unsigned char> rec;
    std::vector<int32_t> mark((size_t)n, -1), list;
    std::vector<double> Ld, X;
    auto wclass_of = [](int w) { return w <= 8 ? 8 : (w <= 16 ? 16 : (w <= 24 ? 24 : 32)); };
    int maxdepth = 0;
    for (const BlockRec& b : F.blocks) maxdepth = std::max(maxdepth, b.depth);
    for (const BlockRec& B : F.blocks) {
        const int nsub = (B.len + SMALL_W - 1) / SMALL_W;
        for (int sub = 0; sub < nsub; ++sub) {
            const int32_t j0 = B.start + sub * SMALL_W, w = std::min(SMALL_W, B.start + B.len - j0);
            const int wcl = wclass_of(w), buf = small_buf_bytes(wcl);
            // target rows: everything below the sub-block that its columns touch (mark: -1 unseen, -2 seen)
            list.clear();
            for (int32_t c = j0; c < j0 + w; ++c)
                for (int64_t p = cptr[c]; p < cptr[c + 1]; ++p) {
                    const int32_t i = cidx[p];
                    if (i >= j0 + w && mark[(size_t)i] == -1) {
                        mark[(size_t)i] = -2;
                        list.push_back(i);
                    }
                }
            std::sort(list.begin(), list.end());
            const int nA = (int)list.size();
            for (int a = 0; a < nA; ++a) mark[(size_t)list[(size_t)a]] = a;  // position in the panel
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
            // launch key: deepest level first; chunks of a separator and sub-blocks of a chunk in order; solve before push
            const int64_t base_key = ((((int64_t)(maxdepth - B.depth) * 4096 + B.chunk) * 16 + sub) * 2);
            auto emit = [&](int kind, int a_lo, int a_hi, int phase) {
                const SmallGeom g = small_geom(w, buf, kind);
                const int cnt = a_hi - a_lo, nch = (cnt + g.R - 1) / g.R;
                Task T;
                T.blk.j0 = j0;
                T.blk.w = w;
                T.blk.nA = cnt;
                T.blk.kind = kind;
                T.blk.rec_off = (int64_t)rec.size();
                T.blk.pad1 = 0;
                T.key = base_key + phase;
                T.wclass = wcl;
                size_t bytes = (size_t)g.invB;
                for (int c = 0; c < nch; ++c) bytes += (size_t)g.arB + (size_t)std::min(g.R, cnt - c * g.R) * g.wp * 8;
                rec.resize(rec.size() + bytes, 0);
                unsigned char* base = rec.data() + T.blk.rec_off;
                if (g.invB) {
                    double* inv = reinterpret_cast<double*>(base);
                    for (int32_t r = 0; r < w; ++r)
                        for (int32_t c = 0; c <= r; ++c) inv[(size_t)inv_row_off(r) + c] = X[(size_t)r * w + c];
                }
                for (int a = a_lo; a < a_hi; ++a) {
                    unsigned char* ch = base + g.invB + (size_t)((a - a_lo) / g.R) * g.chunkB;
                    reinterpret_cast<int32_t*>(ch)[(a - a_lo) % g.R] = list[(size_t)a];
                }
                if (cnt > 0)
                    for (int32_t c = j0; c < j0 + w; ++c)
                        for (int64_t p = cptr[c]; p < cptr[c + 1]; ++p) {
                            const int32_t i = cidx[p];
                            if (i < j0 + w) continue;
                            const int a = mark[(size_t)i];
                            if (a < a_lo || a >= a_hi) continue;
                            unsigned char* ch = base + g.invB + (size_t)((a - a_lo) / g.R) * g.chunkB;
                            reinterpret_cast<double*>(ch + g.arB)[(size_t)((a - a_lo) % g.R) * g.wp + (size_t)(c - j0)] = cval[p];
                        }
                tasks.push_back(T);
            };
            if (nA <= SPLIT_ROWS) {
                emit(TASK_FUSED, 0, nA, 0);
            } else {
                emit(TASK_SOLVE_ONLY, 0, 0, 0);
                const int nparts = (nA + SPLIT_ROWS - 1) / SPLIT_ROWS;
                for (int q = 0; q < nparts; ++q)
                    emit(TASK_PUSH_ONLY, (int)((int64_t)nA * q / nparts), (int)((int64_t)nA * (q + 1) / nparts), 1);
            }
            for (int32_t i : list) mark[(size_t)i] = -1;
        }
    }
    std::vector<int32_t> order(tasks.size());
    for (size_t k = 0; k < order.size(); ++k) order[k] = (int32_t)k;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        return tasks[a].key != tasks[b].key ? tasks[a].key < tasks[b].key : tasks[a].wclass < tasks[b].wclass;
    });
    std::vector<SmallBlk> blk(tasks.size());
    launches.clear();
    for (size_t k = 0, k0 = 0; k < order.size(); ++k) {
        blk[k] = tasks[(size_t)order[k]].blk;
        const Task& cur = tasks[(size_t)order[k]];
        if (k + 1 == order.size() || tasks[(size_t)order[k + 1]].key != cur.key || tasks[(size_t)order[k + 1]].wclass != cur.wclass) {
            launches.push_back({(int)k0, (int)k + 1, cur.wclass});
            k0 = k + 1;
        }
    }
    D.nblocks = (int)blk.size();
    int rc = 0;
    rc |= dev_upload(ctx, &D.blk, blk);
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
    if (!ctx->h_precond_vals.empty()) {  // asgfem_set_precond_matrix_csc: e.g. the Laplacian of the log-transformed problem
        k0 = ctx->h_precond_vals;
    } else {
        ASG_CUDA(ctx, cudaMemcpyAsync(k0.data(), ctx->d_vals, sizeof(double) * ctx->nnz, cudaMemcpyDeviceToHost, ctx->stream));
        ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
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
    PrecondPlan* P = nullptr;
    int rc = precond_build(ctx, ctx->n, ctx->h_rowptr.data(), ctx->h_col.data(), k0.data(), ctx->h_bmask.data(),
                           xy.empty() ? nullptr : xy.data(), &P);
    if (rc) return rc;
    ctx->precond = P;
    return 0;
}

// Factorises the Dirichlet-reduced matrix (rows / columns with bmask != 0 eliminated) on the host and uploads the sweep tasks.
// Used for the matrix of this context (precond_setup) and for the GLOBAL mean matrix of a row-sharded run (dist.cu).
int precond_build(asgfem_ctx* ctx, int64_t nfull, const int64_t* rowptr, const int32_t* col, const double* k0,
                  const uint8_t* bmask, const double* xy, PrecondPlan** out) {
    *out = nullptr;
    CholFactor F;
    std::string err;
    int rc = cholesky_reduced(nfull, rowptr, col, k0, bmask, xy, 256, F, err);
    if (rc) return fail(ctx, rc, "precond_setup: " + err);
    ASG_CHECK(ctx, (int64_t)F.Li.size() < (1ll << 31), ASGFEM_EINVAL, "precond_setup: factor with >= 2^31 nonzeros not supported");
    PrecondPlan* P = new PrecondPlan();
    P->nred = F.n;
    P->lnz = (int64_t)F.Li.size();
    const int64_t n = F.n;
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
    rc = build_tasks(ctx, F, cptr, cidx, cval, P->small, P->launches);
    if (!rc) rc = dev_upload(ctx, &P->d_perm, F.perm);
    if (rc) {
        precond_free_plan(P);
        return rc;
    }
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = P;
    return 0;
}

int precond_apply(asgfem_ctx* ctx, const double* r, double* z) {
    ASG_CHECK(ctx, ctx->precond, ASGFEM_ESTATE, "precond_apply: setup missing");
    return precond_apply_plan(ctx, ctx->precond, r, z, ctx->n, ctx->ld, ctx->d_bmask);
}

// z = K^-1 r for all columns of row-major nrows x ld blocks (ld a multiple of 16); rows with d_bmask != 0 give 0
int precond_apply_plan(asgfem_ctx* ctx, PrecondPlan* P, const double* r, double* z, int64_t nrows, int64_t ld,
                       const uint8_t* d_bmask) {
    ASG_CHECK(ctx, ld > 0 && ld % MT == 0, ASGFEM_ESTATE, "precond_apply: column count must be a positive multiple of 16");
    if (P->work_ld != ld) {  // the factor survives a change of the multi-index set, the work vector does not
        if (P->d_work) cudaFree(P->d_work);
        P->d_work = nullptr;
        P->work_ld = 0;
        ASG_CUDA(ctx, cudaMalloc((void**)&P->d_work, sizeof(double) * (size_t)std::max<int64_t>(P->nred, 1) * ld));
        P->work_ld = ld;
    }
    int blocks = (int)std::min<int64_t>(148 * 8, std::max<int64_t>(1, (P->nred * (ld / 2) + 255) / 256));
    if (P->nred > 0) {
        k_gather_perm<<<blocks, 256, 0, ctx->stream>>>(r, P->d_work, P->d_perm, P->nred, ld);
        const int tiles = (int)(ld / MT);  // all device columns (the column order is private, padding columns hold zeros)
        int split = std::max(1, 148 / tiles);  // CTAs per mode tile: fill the SMs when there are few tiles
        if (const char* e = getenv("ASGFEM_SWEEP_SPLIT")) split = std::max(1, atoi(e));
        for (size_t k = 0; k < P->launches.size(); ++k)
            launch_small<false>(P->launches[k], tiles, split, ctx->stream, P->d_work, ld, P->small);
        for (size_t k = P->launches.size(); k-- > 0;)
            launch_small<true>(P->launches[k], tiles, split, ctx->stream, P->d_work, ld, P->small);
    }
    // z may alias r: boundary rows are zeroed first, interior rows are overwritten from the work vector
    k_zero_masked_rows<<<(unsigned)std::min<int64_t>(nrows, 148 * 8), 128, 0, ctx->stream>>>(z, d_bmask, nrows, ld);
    if (P->nred > 0) k_scatter_perm<<<blocks, 256, 0, ctx->stream>>>(P->d_work, z, P->d_perm, P->nred, ld);
    ASG_CUDA(ctx, cudaGetLastError());
    return 0;
}

}  // namespace asgfem
