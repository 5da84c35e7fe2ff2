// Dense building blocks of the multifrontal host Cholesky (chol.cpp): a persistent worker pool and the blocked partial
// factorisation of one frontal matrix.  Compiled by the host compiler alone (dense_chol.cpp, no CUDA headers) so that the
// AVX2/AVX-512 micro-kernels can use target attributes; the instruction set is picked at run time.
#pragma once
#include <cstdint>
#include <functional>

namespace asgfem {

class DensePool;
DensePool* dense_pool_create(int nthreads);
void dense_pool_destroy(DensePool* p);
int dense_pool_threads(const DensePool* p);
// body(thread, task) for task = 0..ntask-1, dynamically scheduled; returns when all tasks are done.  Not reentrant.
void dense_pool_run(DensePool* p, int ntask, const std::function<void(int, int)>& body);

// rows/columns of slack a front needs beyond its dimension m (trailing tiles are not aligned to the front)
int dense_front_slack();

// Partial Cholesky of the leading s columns of the m x m frontal matrix `a` (column-major, leading dimension ld >=
// m + dense_front_slack(), (m + slack) columns allocated, everything outside the lower triangle of the m x m part zero
// or ignorable): on return columns 0..s-1 hold L (diagonal included) and the trailing (m-s) x (m-s) lower triangle
// holds the Schur complement.  diag0[c] is the magnitude a pivot is compared with (|A_cc| of the original matrix):
// a pivot <= 1e-12 * diag0[c] or non-finite is replaced by 1 and its index (first one) is returned, else -1.
// pool == nullptr runs on the calling thread.  pack is scratch of at least dense_pack_size(m, s) doubles.
int64_t dense_pack_size(int m, int s);
int dense_partial_cholesky(double* a, int m, int s, int64_t ld, const double* diag0, DensePool* pool, double* pack);

const char* dense_kernel_name();

}  // namespace asgfem
