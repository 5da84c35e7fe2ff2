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
// A warp takes a task and 16 modes (two MMA column tiles of 8).  All products are DMMA m8n8k4: the factor data (read exactly
// once per CTA, no reuse; streamed through a per-warp two-stage pipeline of bulk copies, cp.async.bulk + mbarrier, into
// shared memory) gives the A fragments - 8 rows x 4 columns, row stride = 4 mod 8 doubles so that a half-warp touches 16
// different banks, the backward sweep reads the same records transposed -, the 32 x 16 values of the sub-block are B
// fragments / accumulators in registers (fragment orders converted with shuffles).  Round 1 used one DFMA per lane and
// 16-byte operand load (lane = mode): 127 ms per application at config 4, issue / latency bound.
// Sub-blocks with many panel rows (the separators near the root) are split into a solve-only task and push-only tasks
// of <= 64 rows, so that the few large separators at the top of the tree keep all warps of the CTA busy.
// Launches are ordered by (tree depth descending, separator chunk, sub-block, phase): everything inside one launch is
// independent, the launch boundary is the only synchronisation.  The backward sweep runs the launches in reverse.
#include <algorithm>
#include <array>

#include <cstdlib>
#include <cstring>
#include <thread>

#include "common.h"
#include "dense_chol.h"

namespace asgfem {

constexpr int MT = 16;           // modes per CTA (half a warp wide)
constexpr int SMALL_W = 32;      // columns of a sub-block
constexpr int SPLIT_ROWS = 64;   // panel rows of a push-only task; panels up to this size stay fused with their solve
constexpr int SWEEP_THREADS = 512;
constexpr int TOP_DEPTH = 5;     // dissection levels 0..5 (<= 63 separators): launches merged into one kernel per sweep
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
    int64_t rec_bytes = 0;  // size of small.rec
    int* d_launches = nullptr;  // device copy of the launch list (merged launches), made by the first application
};

void precond_free_plan(PrecondPlan* P) {
    if (!P) return;
    void* ptrs[] = {P->small.blk, P->small.rec, P->d_perm, P->d_work, P->d_launches};
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

// Record geometry for a sub-block of width w staged through buffers of `buf` bytes.  Rows of the inverse and of the panel
// have the stride wp doubles with wp % 16 in {4, 12}: the MMA fragment loads (8 rows x 4 columns, or 4 rows x 8 columns
// for the transposed use of the backward sweep) then touch 16 different 8-byte banks per half-warp.
struct SmallGeom {
    int wp, wr, invB, R, arB, chunkB;  // row stride, rows of the inverse (w rounded up to 8), bytes of the inverse, panel rows
                                       // per chunk (multiple of 8), bytes of a chunk's row list, bytes of a full chunk
};
__host__ __device__ __forceinline__ SmallGeom small_geom(int w, int buf, int kind) {
    SmallGeom g;
    g.wr = (w + 7) & ~7;
    g.wp = g.wr + 4;  // 12, 20, 28, 36: a fragment never reads past the end of a row
    g.invB = kind == TASK_PUSH_ONLY ? 0 : g.wr * g.wp * 8;
    g.R = ((buf - 16) / (g.wp * 8 + 4)) & ~7;
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
            const int r = (min(pg.R, pb.nA - pc * pg.R) + 7) & ~7;  // the record pads the last chunk with zero rows
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
    __device__ __forceinline__ void start(int t0w) {  // first piece of the warp's first task -> the buffer the next acquire uses
        set_task(t0w);
        issue(k & 1u);
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

__device__ __forceinline__ void sw_dmma(double2& c, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c.x), "+d"(c.y) : "d"(a), "d"(b));
}
__device__ __forceinline__ void red_add_f64(double* p, double v) { asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }

// C fragments (rows 8 i + g, modes 8 j + 2 t, 2 t + 1) -> B fragments (rows 4 s + t, mode 8 j + g) of the same 32 x 16 block:
// the value of B lane (g, t) sits in C lane (g' = 4 (s & 1) + t, t' = g >> 1), component g & 1 of m-tile s >> 1
template <int NQ>
__device__ __forceinline__ void c_to_b(const double2 (&c)[(NQ + 1) / 2][2], double (&bq)[NQ][2], int g, int t) {
#pragma unroll
    for (int s = 0; s < NQ; ++s) {
        const int src = 4 * (4 * (s & 1) + t) + (g >> 1);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const double x = __shfl_sync(0xffffffffu, c[s >> 1][j].x, src), y = __shfl_sync(0xffffffffu, c[s >> 1][j].y, src);
            bq[s][j] = (g & 1) ? y : x;
        }
    }
}

// The sweeps work on 16 modes per warp (two MMA column tiles of 8); lane = (g, t) = (lane / 4, lane % 4).
// forward task: z_T = inv(L_TT) w_T (unless push-only: W already holds z_T), then W[A(T)] -= P z_T.  All products are
// DMMA m8n8k4: A fragments (inverse / panel, 8 rows x 4 columns) from the staged record, B fragments (4 rows x 8 modes)
// in registers, accumulators in fragment order; the updates of the target rows are fp64 REDs (other tasks share them).
template <int WMAX>
__device__ __forceinline__ void small_fwd(const SmallBlk b, SmallPipe<false>& pipe, double* __restrict__ W, int64_t ld, int64_t mode0,
                                          int lane) {
    constexpr int NQ = WMAX / 4, NM = WMAX / 8;
    const int g = lane >> 2, t = lane & 3;
    const int w = b.w;
    const SmallGeom geo = small_geom(w, pipe.bufB, b.kind);
    double* wt = W + (int64_t)b.j0 * ld + mode0;
    double zb[NQ][2];  // B fragments of w_T, later of z_T
#pragma unroll
    for (int s = 0; s < NQ; ++s)
#pragma unroll
        for (int j = 0; j < 2; ++j) zb[s][j] = 4 * s + t < w ? __ldcg(wt + (int64_t)(4 * s + t) * ld + 8 * j + g) : 0.0;
    if (b.kind != TASK_PUSH_ONLY) {
        const double* inv = reinterpret_cast<const double*>(pipe.acquire()) + g * geo.wp + t;
        double2 zc[NM][2];
#pragma unroll
        for (int i = 0; i < NM; ++i) {
            zc[i][0] = zc[i][1] = make_double2(0.0, 0.0);
            if (8 * i < w) {  // warp-uniform
#pragma unroll
                for (int s = 0; s <= 2 * i + 1; ++s) {
                    const double a = inv[8 * i * geo.wp + 4 * s];
                    sw_dmma(zc[i][0], a, zb[s][0]);
                    sw_dmma(zc[i][1], a, zb[s][1]);
                }
                if (8 * i + g < w) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) *reinterpret_cast<double2*>(wt + (int64_t)(8 * i + g) * ld + 8 * j + 2 * t) = zc[i][j];
                }
            }
        }
        c_to_b<NQ>(zc, zb, g, t);
    }
    for (int a0r = 0; a0r < b.nA; a0r += geo.R) {
        const unsigned char* ch = pipe.acquire();
        const int32_t* ar = reinterpret_cast<const int32_t*>(ch);
        const double* P = reinterpret_cast<const double*>(ch + geo.arB) + g * geo.wp + t;
        const int nr = min(geo.R, b.nA - a0r);
        for (int p = 0; p < nr; p += 8) {
            const int64_t r = p + g < nr ? ar[p + g] : -1;
            double2 c0 = make_double2(0.0, 0.0), c1 = c0;
            const double* pr = P + p * geo.wp;
#pragma unroll
            for (int s = 0; s < NQ; ++s) {
                if (4 * s < w) {
                    const double a = pr[4 * s];
                    sw_dmma(c0, a, zb[s][0]);
                    sw_dmma(c1, a, zb[s][1]);
                }
            }
            if (r >= 0) {
                double* dst = W + r * ld + mode0 + 2 * t;
                red_add_f64(dst, -c0.x);
                red_add_f64(dst + 1, -c0.y);
                red_add_f64(dst + 8, -c1.x);
                red_add_f64(dst + 9, -c1.y);
            }
        }
    }
}

// backward task: y_T = inv(L_TT)^T (z_T - P^T y_A(T)); a pull-only task (part of the panel rows of a split sub-block)
// subtracts its share of P^T y_A from z_T in memory and leaves the solve to the solve-only task of the next launch.
// The transposed products read the same records: A fragment element (column 8 i + g, row 4 s + t) of P^T or inv^T.
template <int WMAX>
__device__ __forceinline__ void small_bwd(const SmallBlk b, SmallPipe<true>& pipe, double* __restrict__ W, int64_t ld, int64_t mode0,
                                          int lane) {
    constexpr int NQ = WMAX / 4, NM = WMAX / 8;
    const int g = lane >> 2, t = lane & 3;
    const int w = b.w;
    const SmallGeom geo = small_geom(w, pipe.bufB, b.kind);
    double* wt = W + (int64_t)b.j0 * ld + mode0;
    double2 acc[NM][2];
#pragma unroll
    for (int i = 0; i < NM; ++i) acc[i][0] = acc[i][1] = make_double2(0.0, 0.0);
    for (int a0r = 0; a0r < b.nA; a0r += geo.R) {
        const unsigned char* ch = pipe.acquire();
        const int32_t* ar = reinterpret_cast<const int32_t*>(ch);
        const double* P = reinterpret_cast<const double*>(ch + geo.arB) + t * geo.wp + g;
        const int nr = min(geo.R, b.nA - a0r);
        // y values (B fragments: row 4 s + t of the chunk, mode 8 j + g) one step ahead of their use
        auto load_y = [&](int p, double (&y)[2]) {
            const bool ok = p + t < nr;
            const double* src = W + (int64_t)(ok ? ar[p + t] : 0) * ld + mode0 + g;
            y[0] = ok ? __ldcg(src) : 0.0;
            y[1] = ok ? __ldcg(src + 8) : 0.0;
        };
        double y0[2], y1[2];
        load_y(0, y0);
        for (int p = 0; p < nr; p += 4) {
            load_y(p + 4, y1);
            const double* pr = P + p * geo.wp;
#pragma unroll
            for (int i = 0; i < NM; ++i) {
                if (8 * i < w) {
                    const double a = pr[8 * i];
                    sw_dmma(acc[i][0], a, y0[0]);
                    sw_dmma(acc[i][1], a, y0[1]);
                }
            }
            y0[0] = y1[0], y0[1] = y1[1];
        }
    }
    if (b.kind == TASK_PUSH_ONLY) {
#pragma unroll
        for (int i = 0; i < NM; ++i) {
            if (8 * i + g < w) {
                double* dst = wt + (int64_t)(8 * i + g) * ld + 2 * t;
                red_add_f64(dst, -acc[i][0].x);
                red_add_f64(dst + 1, -acc[i][0].y);
                red_add_f64(dst + 8, -acc[i][1].x);
                red_add_f64(dst + 9, -acc[i][1].y);
            }
        }
        return;
    }
    // t = z_T - P^T y_A in fragment order, then as B fragments
#pragma unroll
    for (int i = 0; i < NM; ++i) {
        const bool ok = 8 * i + g < w;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const double2 z = ok ? __ldcg(reinterpret_cast<const double2*>(wt + (int64_t)(8 * i + g) * ld + 8 * j + 2 * t)) : make_double2(0.0, 0.0);
            acc[i][j].x = z.x - acc[i][j].x;
            acc[i][j].y = z.y - acc[i][j].y;
        }
    }
    double tb[NQ][2];
    c_to_b<NQ>(acc, tb, g, t);
    const double* inv = reinterpret_cast<const double*>(pipe.acquire()) + t * geo.wp + g;
#pragma unroll
    for (int i = 0; i < NM; ++i) {
        if (8 * i < w) {
            double2 c0 = make_double2(0.0, 0.0), c1 = c0;
#pragma unroll
            for (int s = 2 * i; s < NQ; ++s) {
                if (4 * s < w) {
                    const double a = inv[4 * s * geo.wp + 8 * i];  // inv[column 4 s + t][row 8 i + g]
                    sw_dmma(c0, a, tb[s][0]);
                    sw_dmma(c1, a, tb[s][1]);
                }
            }
            if (8 * i + g < w) {
                *reinterpret_cast<double2*>(wt + (int64_t)(8 * i + g) * ld + 2 * t) = c0;
                *reinterpret_cast<double2*>(wt + (int64_t)(8 * i + g) * ld + 8 + 2 * t) = c1;
            }
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
    const int64_t mode0 = (int64_t)blockIdx.x * MT;
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
            small_bwd<WMAX>(b, pipe, W, ld, mode0, lane);
        else
            small_fwd<WMAX>(b, pipe, W, ld, mode0, lane);
    }
}

// Several consecutive launches of the list in ONE kernel: the mode tiles are independent, so the only synchronisation a
// launch boundary provides - all tasks of launch l are done before a task of launch l + 1 starts - is needed inside a CTA
// only, and a block barrier gives it.  The top of the dissection tree is a long chain of small dependent launches (370 per
// sweep at config 4, 30 us each with a few warps busy); merged they cost a barrier each.
template <int WMAX, bool BWD>
__global__ void __launch_bounds__(WMAX <= 24 ? SWEEP_THREADS : SWEEP_THREADS / 2, 1)
k_small_multi(double* __restrict__ W, int64_t ld, SmallDev S, const int* __restrict__ L, int l0, int l1) {
    extern __shared__ __align__(128) unsigned char sweep_smem[];
    __shared__ __align__(8) unsigned long long sweep_bars[2 * (SWEEP_THREADS / 32)];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int64_t mode0 = (int64_t)blockIdx.x * MT;
    const int bufB = small_buf_bytes(WMAX);
    if (lane == 0) {
        const unsigned bar = (unsigned)__cvta_generic_to_shared(sweep_bars + 2 * warp);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar + 8u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    SmallPipe<BWD> pipe{S, sweep_smem + (size_t)warp * 2 * bufB, (unsigned)__cvta_generic_to_shared(sweep_bars + 2 * warp),
                        bufB,  0, nwarps, lane};
    for (int li = BWD ? l1 - 1 : l0; BWD ? li >= l0 : li < l1; li += BWD ? -1 : 1) {
        const int t0 = L[3 * li], t1 = L[3 * li + 1];
        if (t0 + warp < t1) {
            pipe.t1 = t1;
            pipe.start(t0 + warp);
            for (int t = t0 + warp; t < t1; t += nwarps) {
                const SmallBlk b = load_blk(S, t);
                if (t + nwarps < t1) {
                    const SmallBlk nb = load_blk(S, t + nwarps);
                    if (lane < nb.w) prefetch_l2(W + (int64_t)(nb.j0 + lane) * ld + mode0);
                }
                if constexpr (BWD)
                    small_bwd<WMAX>(b, pipe, W, ld, mode0, lane);
                else
                    small_fwd<WMAX>(b, pipe, W, ld, mode0, lane);
            }
        }
        __threadfence();  // stores and reductions of this launch before the loads of the next one (other warps of the CTA)
        __syncthreads();
    }
}

template <int WMAX, bool BWD>
void launch_small_multi(int l0, int l1, int tiles, cudaStream_t st, double* W, int64_t ld, const SmallDev& S, const int* dL) {
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k_small_multi<WMAX, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SWEEP_SMEM);
        configured = true;
    }
    constexpr int threads = WMAX <= 24 ? SWEEP_THREADS : SWEEP_THREADS / 2;
    k_small_multi<WMAX, BWD><<<tiles, threads, SWEEP_SMEM, st>>>(W, ld, S, dL, l0, l1);
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
// inverse triangle and the dense panel) in launch order, on the host cores: pass A finds the rows below every sub-block
// that its columns touch (candidates = the later rows of its tree node and the node's outside rows, F.node_rows; a
// candidate counts if its sorted row of L has an entry in the column range), a serial sweep lays the records out, pass B
// fills them.  Both passes run over the sub-blocks in parallel; no column copy of L is needed.
void build_tasks_host(const CholFactor& F, std::vector<SmallBlk>& blk, RawVec<unsigned char>& rec,
                      std::vector<std::array<int, 3>>& launches) {
    struct Sub {
        int32_t j0, w, node, wcl;
        int64_t base_key;
        int32_t first_task, ntask;
    };
    struct Task {
        SmallBlk blk;
        int64_t key;  // launch order
        int wclass;
        int32_t a_lo;  // first panel row of this task in the row list of its sub-block
    };
    auto wclass_of = [](int w) { return w <= 8 ? 8 : (w <= 16 ? 16 : (w <= 24 ? 24 : 32)); };
    int maxdepth = 0;
    for (const BlockRec& b : F.blocks) maxdepth = std::max(maxdepth, b.depth);
    std::vector<Sub> subs;
    {
        int32_t node = 0;
        for (const BlockRec& B : F.blocks) {
            while (node + 1 < (int32_t)F.node_lo.size() && B.start >= F.node_hi[(size_t)node]) ++node;
            const int nsub = (B.len + SMALL_W - 1) / SMALL_W;
            for (int sub = 0; sub < nsub; ++sub) {
                Sub s;
                s.j0 = B.start + sub * SMALL_W;
                s.w = std::min(SMALL_W, B.start + B.len - s.j0);
                s.node = node;
                // the top of the tree (few, long separators: a chain of small dependent launches) uses the 32-column kernel
                // for all widths, so that its launches can be merged into one kernel; below, a sub-block keeps its width class
                s.wcl = B.depth <= TOP_DEPTH ? 32 : wclass_of(s.w);
                // launch key: deepest level first; chunks of a separator and sub-blocks of a chunk in order; solve before push
                s.base_key = ((((int64_t)(maxdepth - B.depth) * 4096 + B.chunk) * 16 + sub) * 2);
                s.first_task = s.ntask = 0;
                subs.push_back(s);
            }
        }
    }
    const int32_t nsubs = (int32_t)subs.size();
    int nthreads = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u);
    if (const char* e = std::getenv("ASGFEM_CHOL_THREADS")) nthreads = std::max(1, atoi(e));
    if (F.n < 20000) nthreads = 1;
    DensePool* pool = dense_pool_create(nthreads);
    struct PoolGuard {
        DensePool* p;
        ~PoolGuard() { dense_pool_destroy(p); }
    } pool_guard{pool};
    const int32_t grain = 16, njobs = (nsubs + grain - 1) / grain;
    // ---- pass A: rows below every sub-block
    std::vector<std::vector<int32_t>> lists((size_t)nsubs);
    dense_pool_run(pool, njobs, [&](int, int job) {
        for (int32_t q = job * grain; q < std::min(nsubs, (job + 1) * grain); ++q) {
            const Sub& s = subs[(size_t)q];
            std::vector<int32_t>& list = lists[(size_t)q];
            const int32_t c0 = s.j0, c1 = s.j0 + s.w;
            auto touches = [&](int32_t i) {
                const int32_t* b = F.Li.data() + F.Lp[(size_t)i];
                const int32_t* e = F.Li.data() + F.Lp[(size_t)i + 1];
                const int32_t* at = std::lower_bound(b, e, c0);
                return at < e && *at < c1;
            };
            for (int32_t i = c1; i < F.node_hi[(size_t)s.node]; ++i)
                if (touches(i)) list.push_back(i);
            for (int64_t p = F.node_rptr[(size_t)s.node]; p < F.node_rptr[(size_t)s.node + 1]; ++p)
                if (touches(F.node_rows[(size_t)p])) list.push_back(F.node_rows[(size_t)p]);
        }
    });
    // ---- layout of the records, in sub-block order
    std::vector<Task> tasks;
    int64_t rec_size = 0;
    auto record_bytes = [](int w, int wcl, int kind, int cnt) {
        const SmallGeom g = small_geom(w, small_buf_bytes(wcl), kind);
        const int nch = (cnt + g.R - 1) / g.R;
        size_t bytes = (size_t)g.invB;
        for (int c = 0; c < nch; ++c) bytes += (size_t)g.arB + (size_t)((std::min(g.R, cnt - c * g.R) + 7) & ~7) * g.wp * 8;
        return bytes;
    };
    for (int32_t q = 0; q < nsubs; ++q) {
        Sub& s = subs[(size_t)q];
        const int nA = (int)lists[(size_t)q].size();
        s.first_task = (int32_t)tasks.size();
        auto emit = [&](int kind, int a_lo, int a_hi, int phase) {
            Task T;
            T.blk.j0 = s.j0;
            T.blk.w = s.w;
            T.blk.nA = a_hi - a_lo;
            T.blk.kind = kind;
            T.blk.rec_off = rec_size;
            T.blk.pad1 = 0;
            T.key = s.base_key + phase;
            T.wclass = s.wcl;
            T.a_lo = a_lo;
            rec_size += (int64_t)record_bytes(s.w, s.wcl, kind, a_hi - a_lo);
            tasks.push_back(T);
        };
        if (nA <= SPLIT_ROWS) {
            emit(TASK_FUSED, 0, nA, 0);
        } else {
            emit(TASK_SOLVE_ONLY, 0, 0, 0);
            const int nparts = (nA + SPLIT_ROWS - 1) / SPLIT_ROWS;
            for (int p = 0; p < nparts; ++p)
                emit(TASK_PUSH_ONLY, (int)((int64_t)nA * p / nparts), (int)((int64_t)nA * (p + 1) / nparts), 1);
        }
        s.ntask = (int32_t)tasks.size() - s.first_task;
    }
    rec.resize((size_t)rec_size);  // not value-initialised: zeroed record by record in pass B
    // ---- pass B: inverse triangles and panels
    struct Scratch {
        std::vector<double> Ld, X;
    };
    std::vector<Scratch> scratch((size_t)nthreads);
    dense_pool_run(pool, njobs, [&](int th, int job) {
        Scratch& sc = scratch[(size_t)th];
        for (int32_t q = job * grain; q < std::min(nsubs, (job + 1) * grain); ++q) {
            const Sub& s = subs[(size_t)q];
            const std::vector<int32_t>& list = lists[(size_t)q];
            const int32_t j0 = s.j0, w = s.w;
            // inverse of the diagonal triangle
            sc.Ld.assign((size_t)w * w, 0.0);
            sc.X.assign((size_t)w * w, 0.0);
            double* Ld = sc.Ld.data();
            double* X = sc.X.data();
            for (int32_t r = 0; r < w; ++r) {
                Ld[(size_t)r * w + r] = 1.0 / F.dinv[(size_t)j0 + r];
                for (int64_t p = F.Lp[(size_t)j0 + r]; p < F.Lp[(size_t)j0 + r + 1]; ++p)
                    if (F.Li[p] >= j0) Ld[(size_t)r * w + (F.Li[p] - j0)] = F.Lx[p];
            }
            for (int32_t j = 0; j < w; ++j) {
                X[(size_t)j * w + j] = 1.0 / Ld[(size_t)j * w + j];
                for (int32_t r = j + 1; r < w; ++r) {
                    double sum = 0.0;
                    for (int32_t k = j; k < r; ++k) sum += Ld[(size_t)r * w + k] * X[(size_t)k * w + j];
                    X[(size_t)r * w + j] = -sum / Ld[(size_t)r * w + r];
                }
            }
            for (int32_t t = s.first_task; t < s.first_task + s.ntask; ++t) {
                const Task& T = tasks[(size_t)t];
                const SmallGeom g = small_geom(w, small_buf_bytes(s.wcl), T.blk.kind);
                const int cnt = T.blk.nA;
                unsigned char* base = rec.data() + T.blk.rec_off;
                std::memset(base, 0, record_bytes(w, s.wcl, T.blk.kind, cnt));
                if (g.invB) {
                    double* inv = reinterpret_cast<double*>(base);
                    for (int32_t r = 0; r < w; ++r)
                        for (int32_t c = 0; c <= r; ++c) inv[(size_t)r * g.wp + c] = X[(size_t)r * w + c];
                }
                for (int a = 0; a < cnt; ++a) {
                    unsigned char* ch = base + g.invB + (size_t)(a / g.R) * g.chunkB;
                    const int32_t i = list[(size_t)(T.a_lo + a)];
                    reinterpret_cast<int32_t*>(ch)[a % g.R] = i;
                    double* prow = reinterpret_cast<double*>(ch + g.arB) + (size_t)(a % g.R) * g.wp;
                    const int32_t* b = F.Li.data() + F.Lp[(size_t)i];
                    const int32_t* e = F.Li.data() + F.Lp[(size_t)i + 1];
                    for (const int32_t* at = std::lower_bound(b, e, j0); at < e && *at < j0 + w; ++at)
                        prow[*at - j0] = F.Lx[(size_t)(at - F.Li.data())];
                }
            }
        }
    });
    std::vector<int32_t> order(tasks.size());
    for (size_t k = 0; k < order.size(); ++k) order[k] = (int32_t)k;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        return tasks[a].key != tasks[b].key ? tasks[a].key < tasks[b].key : tasks[a].wclass < tasks[b].wclass;
    });
    blk.resize(tasks.size());
    launches.clear();
    for (size_t k = 0, k0 = 0; k < order.size(); ++k) {
        blk[k] = tasks[(size_t)order[k]].blk;
        const Task& cur = tasks[(size_t)order[k]];
        if (k + 1 == order.size() || tasks[(size_t)order[k + 1]].key != cur.key || tasks[(size_t)order[k + 1]].wclass != cur.wclass) {
            launches.push_back({(int)k0, (int)k + 1, cur.wclass});
            k0 = k + 1;
        }
    }
}

// The serial builder the parallel one above replaced, working on a column copy of L: kept as its cross-check
// (ASGFEM_TASKS_SERIAL=1, tools/chol_bench.cpp compares the two byte by byte).
void build_tasks_serial(const CholFactor& F, std::vector<SmallBlk>& blk, RawVec<unsigned char>& rec,
                        std::vector<std::array<int, 3>>& launches) {
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
    struct Task {
        SmallBlk blk;
        int64_t key;  // launch order
        int wclass;
    };
    std::vector<Task> tasks;
    rec.clear();
    std::vector<int32_t> mark((size_t)n, -1), list;
    std::vector<double> Ld, X;
    auto wclass_of = [](int w) { return w <= 8 ? 8 : (w <= 16 ? 16 : (w <= 24 ? 24 : 32)); };
    int maxdepth = 0;
    for (const BlockRec& b : F.blocks) maxdepth = std::max(maxdepth, b.depth);
    for (const BlockRec& B : F.blocks) {
        const int nsub = (B.len + SMALL_W - 1) / SMALL_W;
        for (int sub = 0; sub < nsub; ++sub) {
            const int32_t j0 = B.start + sub * SMALL_W, w = std::min(SMALL_W, B.start + B.len - j0);
            // the top of the tree (few, long separators: a chain of small dependent launches) uses the 32-column kernel for
            // all widths, so that its launches can be merged into one kernel; below, a sub-block keeps its width class
            const int wcl = B.depth <= TOP_DEPTH ? 32 : wclass_of(w), buf = small_buf_bytes(wcl);
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
                for (int c = 0; c < nch; ++c) bytes += (size_t)g.arB + (size_t)((std::min(g.R, cnt - c * g.R) + 7) & ~7) * g.wp * 8;
                {
                    const size_t old_size = rec.size();
                    rec.resize(old_size + bytes);
                    std::memset(rec.data() + old_size, 0, bytes);
                }
                unsigned char* base = rec.data() + T.blk.rec_off;
                if (g.invB) {
                    double* inv = reinterpret_cast<double*>(base);
                    for (int32_t r = 0; r < w; ++r)
                        for (int32_t c = 0; c <= r; ++c) inv[(size_t)r * g.wp + c] = X[(size_t)r * w + c];
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
    blk.resize(tasks.size());
    launches.clear();
    for (size_t k = 0, k0 = 0; k < order.size(); ++k) {
        blk[k] = tasks[(size_t)order[k]].blk;
        const Task& cur = tasks[(size_t)order[k]];
        if (k + 1 == order.size() || tasks[(size_t)order[k + 1]].key != cur.key || tasks[(size_t)order[k + 1]].wclass != cur.wclass) {
            launches.push_back({(int)k0, (int)k + 1, cur.wclass});
            k0 = k + 1;
        }
    }
}

int build_tasks(asgfem_ctx* ctx, const CholFactor& F, SmallDev& D, std::vector<std::array<int, 3>>& launches, int64_t* rec_bytes) {
    std::vector<SmallBlk> blk;
    RawVec<unsigned char> rec;
    if (std::getenv("ASGFEM_TASKS_SERIAL"))
        build_tasks_serial(F, blk, rec, launches);
    else
        build_tasks_host(F, blk, rec, launches);
    D.nblocks = (int)blk.size();
    int rc = 0;
    rc |= dev_upload(ctx, &D.blk, blk);
    rc |= dev_upload(ctx, &D.rec, rec);
    *rec_bytes = (int64_t)rec.size();
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

// host-only access to the task builders for tools/chol_bench.cpp (task descriptors as raw bytes)
void precond_tasks_host(const CholFactor& F, bool serial_reference, std::vector<unsigned char>& blk_bytes, RawVec<unsigned char>& rec,
                        std::vector<std::array<int, 3>>& launches) {
    std::vector<SmallBlk> blk;
    if (serial_reference)
        build_tasks_serial(F, blk, rec, launches);
    else
        build_tasks_host(F, blk, rec, launches);
    blk_bytes.assign(reinterpret_cast<const unsigned char*>(blk.data()), reinterpret_cast<const unsigned char*>(blk.data() + blk.size()));
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
    rc = build_tasks(ctx, F, P->small, P->launches, &P->rec_bytes);
    if (!rc) rc = dev_upload(ctx, &P->d_perm, F.perm);
    if (rc) {
        precond_free_plan(P);
        return rc;
    }
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = P;
    return 0;
}

// ---- hand-over of a built plan to other ranks (dist.cu: rank 0 factorises the global mean matrix, the others receive the
//      sweep tasks over NCCL).  sizes = {nred, lnz, task count, record bytes, launch count}; buffers = {perm (int32 x nred),
//      task descriptors (32 bytes each), records}; the launch list (3 ints per launch) travels as host data.
void precond_plan_sizes(const PrecondPlan* P, int64_t sizes[5]) {
    sizes[0] = P->nred, sizes[1] = P->lnz, sizes[2] = P->small.nblocks, sizes[3] = P->rec_bytes, sizes[4] = (int64_t)P->launches.size();
}
int precond_plan_alloc(asgfem_ctx* ctx, const int64_t sizes[5], PrecondPlan** out) {
    PrecondPlan* P = new PrecondPlan();
    P->nred = sizes[0], P->lnz = sizes[1], P->small.nblocks = (int)sizes[2], P->rec_bytes = sizes[3];
    P->launches.assign((size_t)sizes[4], {0, 0, 0});
    *out = P;
    ASG_CUDA(ctx, cudaMalloc((void**)&P->d_perm, sizeof(int32_t) * (size_t)std::max<int64_t>(sizes[0], 1)));
    ASG_CUDA(ctx, cudaMalloc((void**)&P->small.blk, sizeof(SmallBlk) * (size_t)std::max<int64_t>(sizes[2], 1)));
    ASG_CUDA(ctx, cudaMalloc((void**)&P->small.rec, (size_t)std::max<int64_t>(sizes[3], 16)));
    return 0;
}
void precond_plan_buffers(PrecondPlan* P, void* ptrs[3], size_t bytes[3]) {
    ptrs[0] = P->d_perm, bytes[0] = sizeof(int32_t) * (size_t)P->nred;
    ptrs[1] = P->small.blk, bytes[1] = sizeof(SmallBlk) * (size_t)P->small.nblocks;
    ptrs[2] = P->small.rec, bytes[2] = (size_t)P->rec_bytes;
}
int* precond_plan_launches(PrecondPlan* P) { return P->launches.empty() ? nullptr : P->launches[0].data(); }

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
        // runs of consecutive 32-column launches go into one kernel each (needs all tasks of a mode tile in one CTA)
        const bool merge = split == 1 && !getenv("ASGFEM_SWEEP_NOMERGE");
        if (merge && !P->d_launches && !P->launches.empty()) {
            ASG_CUDA(ctx, cudaMalloc((void**)&P->d_launches, sizeof(int) * 3 * P->launches.size()));
            ASG_CUDA(ctx, cudaMemcpyAsync(P->d_launches, P->launches[0].data(), sizeof(int) * 3 * P->launches.size(), cudaMemcpyHostToDevice, ctx->stream));
        }
        std::vector<std::array<int, 2>> runs;  // [first launch, one past the last), in forward order
        for (size_t k = 0; k < P->launches.size();) {
            size_t e = k + 1;
            if (merge && P->launches[k][2] == 32)
                while (e < P->launches.size() && P->launches[e][2] == 32) ++e;
            runs.push_back({(int)k, (int)e});
            k = e;
        }
        for (size_t q = 0; q < runs.size(); ++q) {
            if (runs[q][1] - runs[q][0] > 1)
                launch_small_multi<32, false>(runs[q][0], runs[q][1], tiles, ctx->stream, P->d_work, ld, P->small, P->d_launches);
            else
                launch_small<false>(P->launches[(size_t)runs[q][0]], tiles, split, ctx->stream, P->d_work, ld, P->small);
        }
        for (size_t q = runs.size(); q-- > 0;) {
            if (runs[q][1] - runs[q][0] > 1)
                launch_small_multi<32, true>(runs[q][0], runs[q][1], tiles, ctx->stream, P->d_work, ld, P->small, P->d_launches);
            else
                launch_small<true>(P->launches[(size_t)runs[q][0]], tiles, split, ctx->stream, P->d_work, ld, P->small);
        }
    }
    // z may alias r: boundary rows are zeroed first, interior rows are overwritten from the work vector
    k_zero_masked_rows<<<(unsigned)std::min<int64_t>(nrows, 148 * 8), 128, 0, ctx->stream>>>(z, d_bmask, nrows, ld);
    if (P->nred > 0) k_scatter_perm<<<blocks, 256, 0, ctx->stream>>>(P->d_work, z, P->d_perm, P->nred, ld);
    ASG_CUDA(ctx, cudaGetLastError());
    return 0;
}

}  // namespace asgfem
