// Mean-preconditioned conjugate gradients on the Dirichlet-reduced SPD system: the device counterpart of
// solve_primal! (src/modelproblems/solvers_poisson_primal.jl:130-169).  The reference calls Krylov.gmres;
// for PoissonProblemPrimal the reduced operator and I (x) K_0^{-1} are SPD, so PCG converges to the same
// solution.  Everything stays on the device; per iteration the host reads back two
// scalars (p.Ap and r.z) for the step sizes and the stopping test.
#include <chrono>
#include <cmath>

#include "common.h"

namespace asgfem {

namespace {

// b[:, 0] += b0 ; rows at bdofs zeroed   (solvers_poisson_primal.jl:149-155)
__global__ void k_make_rhs(double* __restrict__ b, const double* __restrict__ b0, const uint8_t* __restrict__ bmask,
                           int64_t n, int64_t ld, int64_t col0) {
    int64_t total = n * ld;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = t / ld, mu = t - i * ld;
        double v = b[t];
        if (mu == col0 || (col0 < 0 && mu < -col0 - 1)) v += b0[i];  // column of the mean mode, or the first -col0 - 1 columns
        b[t] = bmask[i] ? 0.0 : v;
    }
}

// x += alpha p ; r -= alpha q   (one pass over four vectors)
__global__ void k_update_xr(double alpha, const double* __restrict__ p, const double* __restrict__ q,
                            double* __restrict__ x, double* __restrict__ r, int64_t total) {
    const double2* p2 = reinterpret_cast<const double2*>(p);
    const double2* q2 = reinterpret_cast<const double2*>(q);
    double2* x2 = reinterpret_cast<double2*>(x);
    double2* r2 = reinterpret_cast<double2*>(r);
    int64_t half = total >> 1;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < half; t += (int64_t)gridDim.x * blockDim.x) {
        double2 a = p2[t], b = q2[t], xv = x2[t], rv = r2[t];
        xv.x = fma(alpha, a.x, xv.x);
        xv.y = fma(alpha, a.y, xv.y);
        rv.x = fma(-alpha, b.x, rv.x);
        rv.y = fma(-alpha, b.y, rv.y);
        x2[t] = xv;
        r2[t] = rv;
    }
}

struct Timer {
    cudaEvent_t a, b;
    cudaStream_t s;
    explicit Timer(cudaStream_t st) : s(st) {
        cudaEventCreate(&a);
        cudaEventCreate(&b);
    }
    ~Timer() {
        cudaEventDestroy(a);
        cudaEventDestroy(b);
    }
    void start() { cudaEventRecord(a, s); }
    double stop() {
        cudaEventRecord(b, s);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        return ms;
    }
};

}  // namespace

int pcg_solve(asgfem_ctx* ctx, const double* b0_host, double* x, double atol, double rtol, int64_t itmax,
              asgfem_stats* stats) {
    const int64_t n = ctx->n, ld = ctx->ld, total = n * ld;
    const size_t bytes = sizeof(double) * (size_t)total;
    if (itmax <= 0) itmax = 2 * n * ctx->N;  // Krylov's itmax = 0 convention
    double *r = nullptr, *z = nullptr, *p = nullptr, *q = nullptr, *b0 = nullptr;
    auto cleanup = [&]() {
        for (double* v : {r, z, p, q, b0})
            if (v) cudaFree(v);
    };
#define PCG_CUDA(call)                                                   \
    do {                                                                 \
        cudaError_t _e = (call);                                         \
        if (_e != cudaSuccess) {                                         \
            cleanup();                                                   \
            return fail(ctx, _e == cudaErrorMemoryAllocation ? ASGFEM_ENOMEM : ASGFEM_ECUDA, \
                        std::string("pcg: ") + #call + ": " + cudaGetErrorString(_e)); \
        }                                                                \
    } while (0)
#define PCG_RC(expr)      \
    do {                  \
        int _rc = (expr); \
        if (_rc) {        \
            cleanup();    \
            return _rc;   \
        }                 \
    } while (0)
    PCG_CUDA(cudaMalloc((void**)&r, bytes));
    PCG_CUDA(cudaMalloc((void**)&z, bytes));
    PCG_CUDA(cudaMalloc((void**)&p, bytes));
    PCG_CUDA(cudaMalloc((void**)&q, bytes));
    PCG_CUDA(cudaMalloc((void**)&b0, sizeof(double) * n));
    PCG_CUDA(cudaMemcpyAsync(b0, b0_host, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((total / 2 + 255) / 256, 148 * 16));
    Timer tall(ctx->stream), tpart(ctx->stream);
    double ms_apply = 0, ms_prec = 0;

    tall.start();
    // b = deepcopy(sol); b[1] += b0; b[m][bdofs] = 0   -> stored in p for now
    PCG_CUDA(cudaMemcpyAsync(p, x, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    k_make_rhs<<<grid, 256, 0, ctx->stream>>>(p, b0, ctx->d_bmask, n, ld, ctx->sample_mode ? -(ctx->N + 1) : (int64_t)ctx->h_pos[0]);
    // r = b - A x
    PCG_RC(dist_apply(ctx, x, q));
    PCG_CUDA(cudaMemcpyAsync(r, p, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    PCG_RC(vec_axpy(ctx, -1.0, q, r));
    // the reduced system keeps x[bdofs] fixed: residual rows at bdofs are zero by construction
    PCG_RC(dist_precond_apply(ctx, r, z));
    PCG_CUDA(cudaMemcpyAsync(p, z, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    double rz = 0;
    PCG_RC(dist_dot(ctx, r, z, &rz));
    const double rz0 = rz;
    const double eps = atol + rtol * std::sqrt(std::max(rz0, 0.0));
    double ms_setup = tall.stop();

    int64_t k = 0;
    tall.start();
    while (k < itmax && std::sqrt(std::max(rz, 0.0)) > eps) {
        tpart.start();
        PCG_RC(dist_apply(ctx, p, q));
        ms_apply += tpart.stop();
        double pq = 0;
        PCG_RC(dist_dot(ctx, p, q, &pq));
        if (!(pq > 0.0) || !std::isfinite(pq)) {
            cleanup();
            return fail(ctx, ASGFEM_ENUMERIC, "pcg: p.Ap <= 0 - operator not positive definite on the search direction");
        }
        const double alpha = rz / pq;
        k_update_xr<<<grid, 256, 0, ctx->stream>>>(alpha, p, q, x, r, total);
        tpart.start();
        PCG_RC(dist_precond_apply(ctx, r, z));
        ms_prec += tpart.stop();
        double rz_new = 0;
        PCG_RC(dist_dot(ctx, r, z, &rz_new));
        const double beta = rz_new / rz;
        PCG_RC(vec_xpay(ctx, z, beta, p));  // p = z + beta p
        rz = rz_new;
        ++k;
    }
    double ms_iter = tall.stop();

    if (stats) {
        // the reference prints ||A x - b|| (solvers_poisson_primal.jl:165-167); b depends on the warm start that x
        // has overwritten, so the norm of the recursively updated residual r_k = b - A x_k is reported instead
        double rr = 0;
        PCG_RC(dist_dot(ctx, r, r, &rr));
        stats->niter = k;
        stats->solved = std::sqrt(std::max(rz, 0.0)) <= eps ? 1 : 0;
        stats->_pad = 0;
        stats->rz0 = std::sqrt(std::max(rz0, 0.0));
        stats->rzk = std::sqrt(std::max(rz, 0.0));
        stats->residual = std::sqrt(std::max(rr, 0.0));
        stats->ms_setup = ms_setup;
        stats->ms_iterations = ms_iter;
        stats->ms_apply = ms_apply;
        stats->ms_precond = ms_prec;
    }
    PCG_CUDA(cudaStreamSynchronize(ctx->stream));
    cleanup();
    return 0;
}

// BiCGStab on the left-preconditioned system P^-1 S x = P^-1 b (device counterpart of the Krylov.gmres call of
// solve_logpoisson_primal!, src/modelproblems/solvers_logpoisson_primal.jl:163): S = the tensorized operator with
// nonsymmetric blocks, P = I (x) chol(A).  b is overwritten (boundary rows zeroed, then used as work space).
int bicgstab_solve(asgfem_ctx* ctx, double* b, double* x, double atol, double rtol, int64_t itmax, asgfem_stats* stats) {
    const int64_t n = ctx->n, ld = ctx->ld, total = n * ld;
    const size_t bytes = sizeof(double) * (size_t)total;
    if (itmax <= 0) itmax = 2 * n * ctx->N;
    double *r = nullptr, *rh = nullptr, *p = nullptr, *v = nullptr, *t = nullptr, *w = nullptr;
    auto cleanup = [&]() {
        for (double* q : {r, rh, p, v, t, w})
            if (q) cudaFree(q);
    };
    PCG_CUDA(cudaMalloc((void**)&r, bytes));
    PCG_CUDA(cudaMalloc((void**)&rh, bytes));
    PCG_CUDA(cudaMalloc((void**)&p, bytes));
    PCG_CUDA(cudaMalloc((void**)&v, bytes));
    PCG_CUDA(cudaMalloc((void**)&t, bytes));
    PCG_CUDA(cudaMalloc((void**)&w, bytes));
    Timer tall(ctx->stream), tpart(ctx->stream);
    double ms_apply = 0, ms_prec = 0;
    auto op = [&](const double* in, double* out) -> int {  // out = P^-1 S in
        tpart.start();
        int rc = apply_launch(ctx, in, w);
        ms_apply += tpart.stop();
        if (rc) return rc;
        tpart.start();
        rc = precond_apply(ctx, w, out);
        ms_prec += tpart.stop();
        return rc;
    };
    tall.start();
    PCG_RC(vec_mask_rows(ctx, b));
    // r = P^-1 (b - S x)
    PCG_RC(apply_launch(ctx, x, w));
    PCG_RC(vec_axpy(ctx, -1.0, w, b));
    PCG_RC(precond_apply(ctx, b, r));
    PCG_CUDA(cudaMemcpyAsync(rh, r, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    PCG_CUDA(cudaMemcpyAsync(p, r, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    double rr = 0;
    PCG_RC(vec_dot(ctx, r, r, n, &rr));
    const double r0 = std::sqrt(std::max(rr, 0.0));
    const double eps = atol + rtol * r0;
    double rho = rr, alpha = 1, omega = 1;
    double ms_setup = tall.stop();
    int64_t k = 0;
    int restarts = 0;
    double rn = r0;
    tall.start();
    while (k < itmax && rn > eps) {
        PCG_RC(op(p, v));
        double rhv = 0;
        PCG_RC(vec_dot(ctx, rh, v, n, &rhv));
        if (rhv == 0.0 || !std::isfinite(rhv)) {
            cleanup();
            return fail(ctx, ASGFEM_ENUMERIC, "bicgstab: breakdown (shadow residual orthogonal to P^-1 S p)");
        }
        alpha = rho / rhv;
        PCG_RC(vec_axpy(ctx, alpha, p, x));    // x += alpha p
        PCG_RC(vec_axpy(ctx, -alpha, v, r));   // s = r - alpha v   (kept in r)
        double ss = 0;
        PCG_RC(vec_dot(ctx, r, r, n, &ss));
        ++k;
        rn = std::sqrt(std::max(ss, 0.0));
        if (rn <= eps) break;
        PCG_RC(op(r, t));
        double ts = 0, tt = 0;
        PCG_RC(vec_dot(ctx, t, r, n, &ts));
        PCG_RC(vec_dot(ctx, t, t, n, &tt));
        if (!(tt > 0.0) || !std::isfinite(tt)) {
            cleanup();
            return fail(ctx, ASGFEM_ENUMERIC, "bicgstab: breakdown (P^-1 S s = 0)");
        }
        omega = ts / tt;
        PCG_RC(vec_axpy(ctx, omega, r, x));    // x += omega s
        PCG_RC(vec_axpy(ctx, -omega, t, r));   // r = s - omega t
        PCG_RC(vec_dot(ctx, r, r, n, &ss));
        rn = std::sqrt(std::max(ss, 0.0));
        double rho_new = 0;
        PCG_RC(vec_dot(ctx, rh, r, n, &rho_new));
        if (rn <= eps) break;
        // (near-)breakdown: the shadow residual has become orthogonal to r, or the stabilisation step stagnated:
        // restart from the current iterate with a new shadow residual
        if (!(std::fabs(rho_new) > 1e-14 * rn * r0) || omega == 0.0 || !std::isfinite(rho_new)) {
            if (++restarts > 50) {
                cleanup();
                return fail(ctx, ASGFEM_ENUMERIC, "bicgstab: repeated breakdown (rho = 0 or omega = 0)");
            }
            PCG_CUDA(cudaMemcpyAsync(rh, r, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
            PCG_CUDA(cudaMemcpyAsync(p, r, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
            rho = ss;
            continue;
        }
        const double beta = (rho_new / rho) * (alpha / omega);
        PCG_RC(vec_axpy(ctx, -omega, v, p));   // p = r + beta (p - omega v)
        PCG_RC(vec_xpay(ctx, r, beta, p));
        rho = rho_new;
    }
    double ms_iter = tall.stop();
    if (stats) {
        stats->niter = k;
        stats->solved = rn <= eps ? 1 : 0;
        stats->_pad = 0;
        stats->rz0 = r0;
        stats->rzk = rn;
        stats->residual = rn;  // norm of the (recursively updated) preconditioned residual
        stats->ms_setup = ms_setup;
        stats->ms_iterations = ms_iter;
        stats->ms_apply = ms_apply;
        stats->ms_precond = ms_prec;
    }
    PCG_CUDA(cudaStreamSynchronize(ctx->stream));
    cleanup();
    return 0;
}

}  // namespace asgfem
