// CPU-only timing and cross-check of the host Cholesky (csrc/chol.cpp) on the 7-point pattern of a structured P1 mesh
// (nx x nx nodes): the multifrontal factor against the scalar up-looking one (same patterns, values to rounding) and the
// residual of L L^T x = b; then the sweep-task builder of sptrsv.cu (parallel, from the rows of L) against its serial
// reference (column copy of L), byte by byte.  Links libasgfem_cuda.so; no GPU needed.  Build/run: tools/chol_bench.sh [nx] [check]
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <array>
#include <vector>

#include "common.h"
using namespace asgfem;

int main(int argc, char** argv) {
    const int nx = argc > 1 ? atoi(argv[1]) : 512;
    const bool check = argc > 2;
    const int64_t n = (int64_t)nx * nx;
    std::vector<int64_t> rp((size_t)n + 1, 0);
    std::vector<int32_t> col;
    std::vector<double> val;
    std::vector<uint8_t> bnd((size_t)n, 0);
    std::vector<double> xy((size_t)2 * n);
    for (int y = 0; y < nx; ++y)
        for (int x = 0; x < nx; ++x) {
            const int64_t i = (int64_t)y * nx + x;
            xy[2 * i] = x / (double)(nx - 1);
            xy[2 * i + 1] = y / (double)(nx - 1);
            if (x == 0 || y == 0 || x == nx - 1 || y == nx - 1) bnd[i] = 1;
            const int dx[7] = {-1, 0, -1, 0, 1, 0, 1}, dy[7] = {-1, -1, 0, 0, 0, 1, 1};
            for (int k = 0; k < 7; ++k) {
                const int xx = x + dx[k], yy = y + dy[k];
                if (xx < 0 || yy < 0 || xx >= nx || yy >= nx) continue;
                col.push_back((int32_t)(yy * nx + xx));
                // a variable-coefficient 5-point stencil plus weak diagonal couplings, symmetric by construction
                const double w = 1.0 + 0.3 * std::sin(0.37 * (x + xx)) * std::cos(0.23 * (y + yy));
                val.push_back(k == 3 ? 0.0 : (k == 0 || k == 6 ? -0.05 * w : -w));
            }
            rp[i + 1] = (int64_t)col.size();
        }
    for (int64_t i = 0; i < n; ++i) {  // diagonal = 1e-3 - sum of the off-diagonal entries
        double sum = 0.0;
        int64_t dpos = -1;
        for (int64_t p = rp[i]; p < rp[i + 1]; ++p) {
            if (col[p] == i)
                dpos = p;
            else
                sum += val[p];
        }
        val[dpos] = 1.0e-3 - sum;
    }
    CholFactor F;
    std::string err;
    auto t0 = std::chrono::steady_clock::now();
    int rc = cholesky_reduced(n, rp.data(), col.data(), val.data(), bnd.data(), xy.data(), 256, F, err);
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("nx=%d rc=%d lnz=%zu time %.2f s %s\n", nx, rc, F.Li.size(), s, err.c_str());
    if (rc != 0) return rc;
    if (!check) {
        std::vector<unsigned char> b1;
        RawVec<unsigned char> r1;
        std::vector<std::array<int, 3>> l1;
        t0 = std::chrono::steady_clock::now();
        precond_tasks_host(F, false, b1, r1, l1);
        s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("sweep tasks: %zu tasks, %zu launches, %.1f MB of records, %.2f s\n", b1.size() / 32, l1.size(), r1.size() / 1.0e6, s);
        return 0;
    }
    // residual of the solve with the factor: P A P^T = L L^T
    const int64_t nr = F.n;
    std::vector<double> b((size_t)nr), x((size_t)nr);
    for (int64_t k = 0; k < nr; ++k) b[k] = std::sin(0.001 * k) + 0.5;
    x = b;
    for (int64_t k = 0; k < nr; ++k) {
        double v = x[k];
        for (int64_t p = F.Lp[k]; p < F.Lp[k + 1]; ++p) v -= F.Lx[p] * x[F.Li[p]];
        x[k] = v * F.dinv[k];
    }
    for (int64_t k = nr - 1; k >= 0; --k) {
        const double v = x[k] * F.dinv[k];
        x[k] = v;
        for (int64_t p = F.Lp[k]; p < F.Lp[k + 1]; ++p) x[F.Li[p]] -= F.Lx[p] * v;
    }
    std::vector<double> xf((size_t)n, 0.0);
    for (int64_t k = 0; k < nr; ++k) xf[F.perm[k]] = x[k];
    double rmax = 0.0, bmax = 0.0;
    for (int64_t k = 0; k < nr; ++k) {
        const int64_t i = F.perm[k];
        double v = 0.0;
        for (int64_t p = rp[i]; p < rp[i + 1]; ++p) v += val[p] * xf[col[p]];
        rmax = std::max(rmax, std::fabs(v - b[k]));
        bmax = std::max(bmax, std::fabs(b[k]));
    }
    printf("solve residual |A x - b|_inf / |b|_inf = %.3e\n", rmax / bmax);
    // the other numeric phase on the same ordering
    const bool was_up = getenv("ASGFEM_CHOL_UPLOOKING") != nullptr;
    if (was_up)
        unsetenv("ASGFEM_CHOL_UPLOOKING");
    else
        setenv("ASGFEM_CHOL_UPLOOKING", "1", 1);
    CholFactor G;
    t0 = std::chrono::steady_clock::now();
    rc = cholesky_reduced(n, rp.data(), col.data(), val.data(), bnd.data(), xy.data(), 256, G, err);
    s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("%s: rc=%d lnz=%zu time %.2f s\n", was_up ? "multifrontal" : "up-looking", rc, G.Li.size(), s);
    bool same = F.Lp == G.Lp && F.Li == G.Li && F.perm == G.perm;
    double dmax = 0.0;
    if (same) {
        for (size_t p = 0; p < F.Lx.size(); ++p) dmax = std::max(dmax, std::fabs(F.Lx[p] - G.Lx[p]) / (1.0 + std::fabs(G.Lx[p])));
        for (size_t k = 0; k < F.dinv.size(); ++k) dmax = std::max(dmax, std::fabs(F.dinv[k] - G.dinv[k]) / std::fabs(G.dinv[k]));
    }
    printf("patterns identical: %s, max value difference %.3e\n", same ? "yes" : "NO", dmax);
    // sweep tasks: parallel builder against the serial reference, for both factors
    bool tasks_same = true;
    for (const CholFactor* Q : {&F, &G}) {
        std::vector<unsigned char> b1, b2;
        RawVec<unsigned char> r1, r2;
        std::vector<std::array<int, 3>> l1, l2;
        t0 = std::chrono::steady_clock::now();
        precond_tasks_host(*Q, false, b1, r1, l1);
        const double s1 = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        t0 = std::chrono::steady_clock::now();
        precond_tasks_host(*Q, true, b2, r2, l2);
        const double s2 = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const bool eq = b1 == b2 && r1 == r2 && l1 == l2;
        printf("sweep tasks: %zu tasks, %zu launches, %.1f MB of records; parallel %.2f s, serial reference %.2f s, identical: %s\n",
               b1.size() / 32, l1.size(), r1.size() / 1.0e6, s1, s2, eq ? "yes" : "NO");
        tasks_same = tasks_same && eq;
    }
    return same && tasks_same && dmax < 1e-9 && rmax / bmax < 1e-8 ? 0 : 1;
}
