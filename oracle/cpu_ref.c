/* Oracle (test infrastructure / timed CPU baseline only - never linked into the product).
 *
 * C restatement of LinearAlgebra.mul!(Ax, S::MySystemPrimal, x)
 * (reference: src/modelproblems/solvers_poisson_primal.jl:86-124): for every mode mu one CSC sweep with A0
 * (:107, addblock_matmul!), then one CSC sweep with A_e per nonzero G[(e-1)N+mu, nu] (:110-117), then the
 * boundary rows are zeroed (:119-121).  N + nnz(G) sweeps per application, each re-reading a whole matrix -
 * exactly the reference's data flow.  The reference's additional N^2*M sparse getindex probes (:111) are NOT
 * charged (only the nonzero couplings are visited), which favours this baseline.
 * The reference is single-threaded; `nthreads` > 1 parallelises over mu with OpenMP ("fair" CPU number).
 *
 * Layout: x, Ax flat n*N, block mu contiguous (src/sgfevector.jl:97-101).  Matrices: shared CSC pattern
 * (colptr, rowval, 0-based) with (M+1) value planes.  Couplings as CSR over mu: (cptr, cm (1..M), cnu, cg) in
 * the visiting order nu ascending, e ascending.
 */
#include <omp.h>
#include <stdint.h>
#include <string.h>

static void addblock_matmul(double* a, const int64_t* colptr, const int32_t* rowval, const double* val, const double* b,
                            int64_t n, double factor) {
    for (int64_t col = 0; col < n; ++col) {
        double bc = b[col];
        for (int64_t p = colptr[col]; p < colptr[col + 1]; ++p) a[rowval[p]] += bc * val[p] * factor;
    }
}

void cpu_ref_mul(int64_t n, int64_t N, int64_t nnz, const int64_t* colptr, const int32_t* rowval, const double* vals,
                 const int32_t* cptr, const int32_t* cm, const int32_t* cnu, const double* cg, int64_t nb,
                 const int64_t* bdofs, const double* x, double* Ax, int nthreads) {
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int64_t mu = 0; mu < N; ++mu) {
        double* a = Ax + mu * n;
        memset(a, 0, sizeof(double) * n);
        addblock_matmul(a, colptr, rowval, vals, x + mu * n, n, 1.0);
        for (int32_t e = cptr[mu]; e < cptr[mu + 1]; ++e)
            addblock_matmul(a, colptr, rowval, vals + (int64_t)cm[e] * nnz, x + (int64_t)cnu[e] * n, n, cg[e]);
        for (int64_t k = 0; k < nb; ++k) a[bdofs[k]] = 0.0;
    }
}

int cpu_ref_max_threads(void) { return omp_get_max_threads(); }
