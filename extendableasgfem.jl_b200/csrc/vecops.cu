// SGFEVector storage and the PCG vector operations.
//
// Device layout (private): row-major n x ld, element (dof i, mode mu) at i*ld + pos[mu] with the column order chosen by
// the operator's plan (ctx->h_pos, apply_mma.cu), ld = number of columns (multiple of 32); padding columns are kept at zero.  The boundary layout is the reference's flat entries vector
// (src/sgfevector.jl:97-101: block mu contiguous = column-major n x N); conversion happens here, on the
// device, chunk by chunk through a staging buffer.
#include <algorithm>
#include <cstdlib>

#include "common.h"

namespace asgfem {

namespace {

constexpr int TP = 32;

// stage: mc x n (mode-major chunk of the reference layout) -> d[i*ld + pos[mu0 + k]]
__global__ void k_chunk_to_device(const double* __restrict__ stage, double* __restrict__ d, int64_t n, int64_t ld,
                                  int64_t mu0, int mc, const int32_t* __restrict__ pos) {
    __shared__ double tile[TP][TP + 1];
    int64_t i0 = (int64_t)blockIdx.x * TP;
    int k0 = blockIdx.y * TP;
    for (int r = threadIdx.y; r < TP; r += blockDim.y) {
        int k = k0 + r;
        int64_t i = i0 + threadIdx.x;
        tile[r][threadIdx.x] = (k < mc && i < n) ? stage[(int64_t)k * n + i] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < TP; r += blockDim.y) {
        int64_t i = i0 + r;
        int k = k0 + threadIdx.x;
        if (i < n && k < mc) d[i * ld + pos[mu0 + k]] = tile[threadIdx.x][r];
    }
}

__global__ void k_chunk_to_host(const double* __restrict__ d, double* __restrict__ stage, int64_t n, int64_t ld,
                                int64_t mu0, int mc, const int32_t* __restrict__ pos) {
    __shared__ double tile[TP][TP + 1];
    int64_t i0 = (int64_t)blockIdx.x * TP;
    int k0 = blockIdx.y * TP;
    for (int r = threadIdx.y; r < TP; r += blockDim.y) {
        int64_t i = i0 + r;
        int k = k0 + threadIdx.x;
        tile[r][threadIdx.x] = (i < n && k < mc) ? d[i * ld + pos[mu0 + k]] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < TP; r += blockDim.y) {
        int k = k0 + r;
        int64_t i = i0 + threadIdx.x;
        if (k < mc && i < n) stage[(int64_t)k * n + i] = tile[threadIdx.x][r];
    }
}

__device__ __forceinline__ double splitmix_pm1(uint64_t idx, uint64_t seed) {
    uint64_t z = (idx ^ seed) + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return 2.0 * ((double)(z >> 11) * (1.0 / 9007199254740992.0)) - 1.0;
}

__global__ void k_fill_random(double* __restrict__ d, int64_t n, int64_t ld, uint64_t seed, const int32_t* __restrict__ inv) {
    int64_t total = n * ld;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = t / ld, c = t - i * ld;
        const int64_t mu = inv[c];
        d[t] = mu >= 0 ? splitmix_pm1((uint64_t)(i + n * mu), seed) : 0.0;
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double block_sum(double v) {
    __shared__ double sm[32];
    v = warp_sum(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sm[w] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? sm[threadIdx.x] : 0.0;
    if (w == 0) v = warp_sum(v);
    return v;
}

// deterministic two-stage dot: fixed grid, fixed per-thread order, partials reduced by one block
__global__ void k_dot_partial(const double* __restrict__ a, const double* __restrict__ b, int64_t total,
                              double* __restrict__ partial) {
    double acc = 0.0;
    const double2* a2 = reinterpret_cast<const double2*>(a);
    const double2* b2 = reinterpret_cast<const double2*>(b);
    int64_t half = total >> 1;  // total is a multiple of 8
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < half; t += (int64_t)gridDim.x * blockDim.x) {
        double2 x = a2[t], y = b2[t];
        acc = fma(x.x, y.x, acc);
        acc = fma(x.y, y.y, acc);
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

__global__ void k_reduce_final(const double* __restrict__ partial, int nparts, double* __restrict__ out) {
    double acc = 0.0;
    for (int t = threadIdx.x; t < nparts; t += blockDim.x) acc += partial[t];
    acc = block_sum(acc);
    if (threadIdx.x == 0) out[0] = acc;
}

__global__ void k_axpy(double alpha, const double* __restrict__ x, double* __restrict__ y, int64_t total) {
    const double2* x2 = reinterpret_cast<const double2*>(x);
    double2* y2 = reinterpret_cast<double2*>(y);
    int64_t half = total >> 1;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < half; t += (int64_t)gridDim.x * blockDim.x) {
        double2 a = x2[t], b = y2[t];
        b.x = fma(alpha, a.x, b.x);
        b.y = fma(alpha, a.y, b.y);
        y2[t] = b;
    }
}

__global__ void k_xpay(const double* __restrict__ x, double beta, double* __restrict__ y, int64_t total) {
    const double2* x2 = reinterpret_cast<const double2*>(x);
    double2* y2 = reinterpret_cast<double2*>(y);
    int64_t half = total >> 1;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < half; t += (int64_t)gridDim.x * blockDim.x) {
        double2 a = x2[t], b = y2[t];
        b.x = fma(beta, b.x, a.x);
        b.y = fma(beta, b.y, a.y);
        y2[t] = b;
    }
}

__global__ void k_mask_rows(double* __restrict__ x, const uint8_t* __restrict__ bmask, int64_t n, int64_t ld) {
    int64_t total = n * ld;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        if (bmask[t / ld]) x[t] = 0.0;
    }
}

// the exchange buffer is nrows x N in MODE order (both ends of an exchange may not share the private column order)
__global__ void k_pack_rows(const double* __restrict__ v, int64_t ld, int64_t N, int64_t nrows,
                            const int64_t* __restrict__ rows, double* __restrict__ buf, int pack,
                            const int32_t* __restrict__ pos) {
    for (int64_t r = blockIdx.x; r < nrows; r += gridDim.x) {
        int64_t i = rows[r] - 1;
        for (int64_t k = threadIdx.x; k < N; k += blockDim.x) {
            if (pack)
                buf[r * N + k] = v[i * ld + pos[k]];
            else
                const_cast<double*>(v)[i * ld + pos[k]] = buf[r * N + k];
        }
    }
}

int ensure_stage(asgfem_ctx* ctx, size_t bytes) {
    if (ctx->stage_bytes >= bytes) return 0;
    if (ctx->d_stage) cudaFree(ctx->d_stage);
    ctx->d_stage = nullptr;
    ctx->stage_bytes = 0;
    ASG_CUDA(ctx, cudaMalloc((void**)&ctx->d_stage, bytes));
    ctx->stage_bytes = bytes;
    return 0;
}

int64_t chunk_modes(const asgfem_ctx* ctx) {
    const int64_t budget = 256ll << 20;
    int64_t mc = budget / (8 * std::max<int64_t>(ctx->n, 1));
    return std::max<int64_t>(1, std::min<int64_t>(mc, ctx->N));
}

inline int grid_for(int64_t work, int threads) {
    int64_t b = (work + threads - 1) / threads;
    return (int)std::max<int64_t>(1, std::min<int64_t>(b, 148 * 16));
}

}  // namespace

int vec_to_device_layout(asgfem_ctx* ctx, const double* host, double* dvec) {
    int64_t n = ctx->n, N = ctx->N, ld = ctx->ld;
    int64_t mc = chunk_modes(ctx);
    int rc = ensure_stage(ctx, (size_t)(8 * n * mc));
    if (rc) return rc;
    if (ld != N) ASG_CUDA(ctx, cudaMemsetAsync(dvec, 0, sizeof(double) * n * ld, ctx->stream));
    for (int64_t mu0 = 0; mu0 < N; mu0 += mc) {
        int c = (int)std::min<int64_t>(mc, N - mu0);
        ASG_CUDA(ctx, cudaMemcpyAsync(ctx->d_stage, host + n * mu0, sizeof(double) * n * c, cudaMemcpyHostToDevice,
                                      ctx->stream));
        dim3 grid((unsigned)((n + TP - 1) / TP), (unsigned)((c + TP - 1) / TP));
        k_chunk_to_device<<<grid, dim3(TP, 8), 0, ctx->stream>>>(ctx->d_stage, dvec, n, ld, mu0, c, ctx->d_pos);
    }
    ASG_CUDA(ctx, cudaGetLastError());
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int vec_to_host_layout(asgfem_ctx* ctx, const double* dvec, double* host) {
    int64_t n = ctx->n, N = ctx->N, ld = ctx->ld;
    int64_t mc = chunk_modes(ctx);
    int rc = ensure_stage(ctx, (size_t)(8 * n * mc));
    if (rc) return rc;
    for (int64_t mu0 = 0; mu0 < N; mu0 += mc) {
        int c = (int)std::min<int64_t>(mc, N - mu0);
        dim3 grid((unsigned)((n + TP - 1) / TP), (unsigned)((c + TP - 1) / TP));
        k_chunk_to_host<<<grid, dim3(TP, 8), 0, ctx->stream>>>(dvec, ctx->d_stage, n, ld, mu0, c, ctx->d_pos);
        ASG_CUDA(ctx, cudaMemcpyAsync(host + n * mu0, ctx->d_stage, sizeof(double) * n * c, cudaMemcpyDeviceToHost,
                                      ctx->stream));
    }
    ASG_CUDA(ctx, cudaGetLastError());
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// Y = A X with X and Y on the host (the mul! seam): row blocks are pipelined over three streams so that the
// upload of block b+1, the operator on block b and the download of block b-1 overlap (PCIe is full duplex).  A row
// block may be applied as soon as every X row its columns reference has arrived, i.e. once the block holding its largest
// column index is on the device (blocks are uploaded in ascending order).
int apply_host_pipelined(asgfem_ctx* ctx, const double* x, double* Ax, double* dX, double* dY) {
    const int64_t n = ctx->n, N = ctx->N, ld = ctx->ld;
    int64_t RB = (256ll << 20) / (8 * std::max<int64_t>(N, 1));
    RB = std::max<int64_t>(1024, RB / 1024 * 1024);
    if (const char* e = std::getenv("ASGFEM_HOST_BLOCK_ROWS")) {  // test hook: force many small blocks
        int64_t v = std::atoll(e);
        if (v >= 32) RB = v / 32 * 32;
    }
    const int64_t nb = (n + RB - 1) / RB;
    std::vector<int64_t> dep((size_t)nb, 0);
    for (int64_t b = 0; b < nb; ++b) {
        const int64_t i0 = b * RB, i1 = std::min(n, i0 + RB);
        int32_t maxc = (int32_t)i0;
        for (int64_t p = ctx->h_rowptr[i0]; p < ctx->h_rowptr[i1]; ++p) maxc = std::max(maxc, ctx->h_col[p]);
        dep[b] = maxc / RB;
    }
    struct Pipe {
        cudaStream_t sH = nullptr, sD = nullptr;
        cudaEvent_t copied[2] = {}, in_free[2] = {}, yready[2] = {}, out_free[2] = {};
        double *in[2] = {}, *out[2] = {};
        ~Pipe() {
            for (int k = 0; k < 2; ++k) {
                if (copied[k]) cudaEventDestroy(copied[k]);
                if (in_free[k]) cudaEventDestroy(in_free[k]);
                if (yready[k]) cudaEventDestroy(yready[k]);
                if (out_free[k]) cudaEventDestroy(out_free[k]);
                if (in[k]) cudaFree(in[k]);
                if (out[k]) cudaFree(out[k]);
            }
            if (sH) cudaStreamDestroy(sH);
            if (sD) cudaStreamDestroy(sD);
        }
    } P;
    ASG_CUDA(ctx, cudaStreamCreateWithFlags(&P.sH, cudaStreamNonBlocking));
    ASG_CUDA(ctx, cudaStreamCreateWithFlags(&P.sD, cudaStreamNonBlocking));
    const size_t stage_bytes = sizeof(double) * (size_t)std::min(RB, n) * (size_t)N;
    for (int k = 0; k < 2; ++k) {
        ASG_CUDA(ctx, cudaEventCreateWithFlags(&P.copied[k], cudaEventDisableTiming));
        ASG_CUDA(ctx, cudaEventCreateWithFlags(&P.in_free[k], cudaEventDisableTiming));
        ASG_CUDA(ctx, cudaEventCreateWithFlags(&P.yready[k], cudaEventDisableTiming));
        ASG_CUDA(ctx, cudaEventCreateWithFlags(&P.out_free[k], cudaEventDisableTiming));
        ASG_CUDA(ctx, cudaMalloc((void**)&P.in[k], stage_bytes));
        ASG_CUDA(ctx, cudaMalloc((void**)&P.out[k], stage_bytes));
    }
    if (ld != N) ASG_CUDA(ctx, cudaMemsetAsync(dX, 0, sizeof(double) * n * ld, ctx->stream));
    auto block_grid = [&](int64_t rows) { return dim3((unsigned)((rows + TP - 1) / TP), (unsigned)((N + TP - 1) / TP)); };
    int64_t next_apply = 0;
    for (int64_t b = 0; b < nb; ++b) {
        const int s = (int)(b & 1);
        const int64_t i0 = b * RB, rows = std::min(n, i0 + RB) - i0;
        if (b >= 2) ASG_CUDA(ctx, cudaStreamWaitEvent(P.sH, P.in_free[s], 0));
        ASG_CUDA(ctx, cudaMemcpy2DAsync(P.in[s], sizeof(double) * rows, x + i0, sizeof(double) * n, sizeof(double) * rows,
                                        (size_t)N, cudaMemcpyHostToDevice, P.sH));
        ASG_CUDA(ctx, cudaEventRecord(P.copied[s], P.sH));
        ASG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, P.copied[s], 0));
        k_chunk_to_device<<<block_grid(rows), dim3(TP, 8), 0, ctx->stream>>>(P.in[s], dX + i0 * ld, rows, ld, 0, (int)N, ctx->d_pos);
        ASG_CUDA(ctx, cudaEventRecord(P.in_free[s], ctx->stream));
        while (next_apply < nb && dep[next_apply] <= b) {
            const int64_t a = next_apply++;
            const int t = (int)(a & 1);
            const int64_t a0 = a * RB, arows = std::min(n, a0 + RB) - a0;
            int rc = apply_launch(ctx, dX, dY, a0, a0 + arows);
            if (rc) return rc;
            if (a >= 2) ASG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, P.out_free[t], 0));
            k_chunk_to_host<<<block_grid(arows), dim3(TP, 8), 0, ctx->stream>>>(dY + a0 * ld, P.out[t], arows, ld, 0, (int)N, ctx->d_pos);
            ASG_CUDA(ctx, cudaEventRecord(P.yready[t], ctx->stream));
            ASG_CUDA(ctx, cudaStreamWaitEvent(P.sD, P.yready[t], 0));
            ASG_CUDA(ctx, cudaMemcpy2DAsync(Ax + a0, sizeof(double) * n, P.out[t], sizeof(double) * arows,
                                            sizeof(double) * arows, (size_t)N, cudaMemcpyDeviceToHost, P.sD));
            ASG_CUDA(ctx, cudaEventRecord(P.out_free[t], P.sD));
        }
    }
    ASG_CUDA(ctx, cudaGetLastError());
    ASG_CUDA(ctx, cudaStreamSynchronize(P.sH));
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ASG_CUDA(ctx, cudaStreamSynchronize(P.sD));
    return 0;
}

int vec_dot(asgfem_ctx* ctx, const double* a, const double* b, int64_t nrows, double* out) {
    const int threads = 256, blocks = 148 * 8;
    if (ctx->partial_elems < (size_t)blocks + 8) {
        if (ctx->d_partial) cudaFree(ctx->d_partial);
        ctx->d_partial = nullptr;
        ASG_CUDA(ctx, cudaMalloc((void**)&ctx->d_partial, sizeof(double) * (blocks + 8)));
        ctx->partial_elems = blocks + 8;
    }
    int64_t total = nrows * ctx->ld;
    k_dot_partial<<<blocks, threads, 0, ctx->stream>>>(a, b, total, ctx->d_partial);
    k_reduce_final<<<1, 1024, 0, ctx->stream>>>(ctx->d_partial, blocks, ctx->d_partial + blocks);
    ASG_CUDA(ctx, cudaGetLastError());
    ASG_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_partial + blocks, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int vec_fill_random(asgfem_ctx* ctx, double* d, uint64_t seed) {
    int64_t total = ctx->n * ctx->ld;
    k_fill_random<<<grid_for(total, 256), 256, 0, ctx->stream>>>(d, ctx->n, ctx->ld, seed, ctx->d_inv);
    ASG_CUDA(ctx, cudaGetLastError());
    return 0;
}

int vec_axpy(asgfem_ctx* ctx, double alpha, const double* x, double* y) {
    int64_t total = ctx->n * ctx->ld;
    k_axpy<<<grid_for(total / 2, 256), 256, 0, ctx->stream>>>(alpha, x, y, total);
    ASG_CUDA(ctx, cudaGetLastError());
    return 0;
}

int vec_xpay(asgfem_ctx* ctx, const double* x, double beta, double* y) {
    int64_t total = ctx->n * ctx->ld;
    k_xpay<<<grid_for(total / 2, 256), 256, 0, ctx->stream>>>(x, beta, y, total);
    ASG_CUDA(ctx, cudaGetLastError());
    return 0;
}

int vec_mask_rows(asgfem_ctx* ctx, double* x) {
    int64_t total = ctx->n * ctx->ld;
    k_mask_rows<<<grid_for(total, 256), 256, 0, ctx->stream>>>(x, ctx->d_bmask, ctx->n, ctx->ld);
    ASG_CUDA(ctx, cudaGetLastError());
    return 0;
}

int vec_pack_rows(asgfem_ctx* ctx, const double* v, int64_t nrows, const int64_t* d_rows, double* buf) {
    if (nrows == 0) return 0;
    k_pack_rows<<<(unsigned)std::min<int64_t>(nrows, 148 * 8), 256, 0, ctx->stream>>>(v, ctx->ld, ctx->N, nrows, d_rows,
                                                                                     buf, 1, ctx->d_pos);
    ASG_CUDA(ctx, cudaGetLastError());
    return 0;
}

int vec_unpack_rows(asgfem_ctx* ctx, double* v, int64_t nrows, const int64_t* d_rows, const double* buf) {
    if (nrows == 0) return 0;
    k_pack_rows<<<(unsigned)std::min<int64_t>(nrows, 148 * 8), 256, 0, ctx->stream>>>(
        v, ctx->ld, ctx->N, nrows, d_rows, const_cast<double*>(buf), 0, ctx->d_pos);
    ASG_CUDA(ctx, cudaGetLastError());
    return 0;
}

namespace {
// warp = (dof row, tile of 8 samples): lanes stride over the modes, 8 accumulators, butterfly reduction
__global__ void k_eval_samples(const double* __restrict__ u, int64_t n, int64_t ld, int N, const double* __restrict__ R,
                               int64_t Spad, int64_t S, double* __restrict__ out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int64_t tile = blockIdx.y;
    for (int64_t row = (int64_t)blockIdx.x * nw + warp; row < n; row += (int64_t)gridDim.x * nw) {
        double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const double* ur = u + row * ld;
        for (int k = lane; k < N; k += 32) {
            const double uk = ur[k];
            const double2* r = reinterpret_cast<const double2*>(R + (int64_t)k * Spad + tile * 8);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double2 v = r[j];
                acc[2 * j] = fma(uk, v.x, acc[2 * j]);
                acc[2 * j + 1] = fma(uk, v.y, acc[2 * j + 1]);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (lane == j && tile * 8 + j < S) out[(tile * 8 + j) * n + row] = acc[j];
    }
}
}  // namespace

int vec_eval_samples(asgfem_ctx* ctx, const double* u, const double* dR, int64_t S, int64_t Spad, double* dout) {
    if (S <= 0 || ctx->n <= 0) return 0;
    dim3 grid((unsigned)std::min<int64_t>((ctx->n + 7) / 8, 148 * 8), (unsigned)(Spad / 8));
    k_eval_samples<<<grid, 256, 0, ctx->stream>>>(u, ctx->n, ctx->ld, (int)ctx->ld, dR, Spad, S, dout);
    ASG_CUDA(ctx, cudaGetLastError());
    return 0;
}

}  // namespace asgfem
