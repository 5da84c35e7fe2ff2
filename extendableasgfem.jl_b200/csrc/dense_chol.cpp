// Dense kernels of the multifrontal host Cholesky (see dense_chol.h and chol.cpp).
//
// A frontal matrix is factorised right-looking in panels of NB columns: (1) the NB x NB diagonal block on the calling
// thread, (2) the rows below it (triangular solve) and their copy into a packed, row-block-major panel in parallel over
// row chunks, (3) the rank-NB update of the trailing lower triangle in parallel over strips of rows, by an MR x 4
// register tile (AVX-512: 16 x 4, AVX2+FMA: 8 x 4, portable C otherwise) that streams the packed panel.
#include "dense_chol.h"

#include <immintrin.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdlib>
#include <exception>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace asgfem {

// ---- worker pool ---------------------------------------------------------------------------------------------------
// Workers spin briefly between jobs (the top of the elimination tree issues thousands of short jobs back to back) and
// sleep on a condition variable otherwise.  Every worker acknowledges every job, so the job record is never rewritten
// while somebody still reads it.  An exception inside a task (e.g. std::bad_alloc) is carried to the thread that called run().
class DensePool {
   public:
    explicit DensePool(int nthreads) : nthreads_(std::max(1, nthreads)) {
        for (int t = 1; t < nthreads_; ++t) workers_.emplace_back([this, t]() { worker(t); });
    }
    ~DensePool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            gen_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        for (std::thread& w : workers_) w.join();
    }
    int threads() const { return nthreads_; }
    void run(int ntask, const std::function<void(int, int)>& body) {
        if (ntask <= 0) return;
        if (nthreads_ == 1 || ntask == 1) {
            for (int t = 0; t < ntask; ++t) body(0, t);
            return;
        }
        body_ = &body;
        ntask_ = ntask;
        next_.store(0, std::memory_order_relaxed);
        acks_.store(0, std::memory_order_relaxed);
        gen_.fetch_add(1);  // seq_cst with the sleepers_ / gen_ pair below: a worker about to sleep is either seen here or sees the job
        if (sleepers_.load() > 0) {
            std::lock_guard<std::mutex> lk(mu_);
            cv_.notify_all();
        }
        work(0);
        while (acks_.load(std::memory_order_acquire) != nthreads_ - 1) _mm_pause();
        if (failed_.load(std::memory_order_acquire)) {  // first exception of a task, rethrown on the calling thread
            std::exception_ptr e = error_;
            error_ = nullptr;
            failed_.store(false, std::memory_order_release);
            std::rethrow_exception(e);
        }
    }

   private:
    void work(int th) {
        for (int t = next_.fetch_add(1, std::memory_order_relaxed); t < ntask_; t = next_.fetch_add(1, std::memory_order_relaxed)) {
            if (failed_.load(std::memory_order_relaxed)) continue;  // drain the remaining tasks after a failure
            try {
                (*body_)(th, t);
            } catch (...) {
                std::lock_guard<std::mutex> lk(mu_);
                if (!failed_.load(std::memory_order_relaxed)) {
                    error_ = std::current_exception();
                    failed_.store(true, std::memory_order_release);
                }
            }
        }
    }
    void worker(int th) {
        uint64_t seen = 0;
        for (;;) {
            int spins = 0;
            while (gen_.load(std::memory_order_acquire) == seen) {
                if (++spins < 20000) {
                    _mm_pause();
                    continue;
                }
                std::unique_lock<std::mutex> lk(mu_);
                sleepers_.fetch_add(1);
                cv_.wait(lk, [&]() { return gen_.load() != seen; });
                sleepers_.fetch_sub(1);
            }
            seen = gen_.load(std::memory_order_acquire);
            if (stop_) return;
            work(th);
            acks_.fetch_add(1, std::memory_order_release);
        }
    }
    int nthreads_;
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::atomic<uint64_t> gen_{0};
    std::atomic<int> sleepers_{0}, next_{0}, acks_{0};
    const std::function<void(int, int)>* body_ = nullptr;
    int ntask_ = 0;
    bool stop_ = false;
    std::atomic<bool> failed_{false};
    std::exception_ptr error_;
};

DensePool* dense_pool_create(int nthreads) { return new DensePool(nthreads); }
void dense_pool_destroy(DensePool* p) { delete p; }
int dense_pool_threads(const DensePool* p) { return p ? p->threads() : 1; }
void dense_pool_run(DensePool* p, int ntask, const std::function<void(int, int)>& body) {
    if (p)
        p->run(ntask, body);
    else
        for (int t = 0; t < ntask; ++t) body(0, t);
}

// ---- register tiles ------------------------------------------------------------------------------------------------
// C[0..MR) x [0..4) -= sum_k A[k * MR + r] * B[k * MR + c]   (A, B inside the packed panel; C column-major, ld)
namespace {

typedef void (*TileFn)(const double* A, const double* B, int kb, double* C, int64_t ld);

void tile_c8(const double* A, const double* B, int kb, double* C, int64_t ld) {
    double acc[4][8] = {};
    for (int k = 0; k < kb; ++k)
        for (int c = 0; c < 4; ++c)
            for (int r = 0; r < 8; ++r) acc[c][r] += A[k * 8 + r] * B[k * 8 + c];
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 8; ++r) C[r + c * ld] -= acc[c][r];
}

__attribute__((target("avx2,fma"))) void tile_avx2(const double* A, const double* B, int kb, double* C, int64_t ld) {
    __m256d c00 = _mm256_setzero_pd(), c10 = c00, c01 = c00, c11 = c00, c02 = c00, c12 = c00, c03 = c00, c13 = c00;
    for (int k = 0; k < kb; ++k) {
        const __m256d a0 = _mm256_loadu_pd(A + 8 * k), a1 = _mm256_loadu_pd(A + 8 * k + 4);
        __m256d b = _mm256_broadcast_sd(B + 8 * k);
        c00 = _mm256_fmadd_pd(a0, b, c00);
        c10 = _mm256_fmadd_pd(a1, b, c10);
        b = _mm256_broadcast_sd(B + 8 * k + 1);
        c01 = _mm256_fmadd_pd(a0, b, c01);
        c11 = _mm256_fmadd_pd(a1, b, c11);
        b = _mm256_broadcast_sd(B + 8 * k + 2);
        c02 = _mm256_fmadd_pd(a0, b, c02);
        c12 = _mm256_fmadd_pd(a1, b, c12);
        b = _mm256_broadcast_sd(B + 8 * k + 3);
        c03 = _mm256_fmadd_pd(a0, b, c03);
        c13 = _mm256_fmadd_pd(a1, b, c13);
    }
    _mm256_storeu_pd(C, _mm256_sub_pd(_mm256_loadu_pd(C), c00));
    _mm256_storeu_pd(C + 4, _mm256_sub_pd(_mm256_loadu_pd(C + 4), c10));
    _mm256_storeu_pd(C + ld, _mm256_sub_pd(_mm256_loadu_pd(C + ld), c01));
    _mm256_storeu_pd(C + ld + 4, _mm256_sub_pd(_mm256_loadu_pd(C + ld + 4), c11));
    _mm256_storeu_pd(C + 2 * ld, _mm256_sub_pd(_mm256_loadu_pd(C + 2 * ld), c02));
    _mm256_storeu_pd(C + 2 * ld + 4, _mm256_sub_pd(_mm256_loadu_pd(C + 2 * ld + 4), c12));
    _mm256_storeu_pd(C + 3 * ld, _mm256_sub_pd(_mm256_loadu_pd(C + 3 * ld), c03));
    _mm256_storeu_pd(C + 3 * ld + 4, _mm256_sub_pd(_mm256_loadu_pd(C + 3 * ld + 4), c13));
}

__attribute__((target("avx512f"))) void tile_avx512(const double* A, const double* B, int kb, double* C, int64_t ld) {
    __m512d c00 = _mm512_setzero_pd(), c10 = c00, c01 = c00, c11 = c00, c02 = c00, c12 = c00, c03 = c00, c13 = c00;
    for (int k = 0; k < kb; ++k) {
        const __m512d a0 = _mm512_loadu_pd(A + 16 * k), a1 = _mm512_loadu_pd(A + 16 * k + 8);
        __m512d b = _mm512_set1_pd(B[16 * k]);
        c00 = _mm512_fmadd_pd(a0, b, c00);
        c10 = _mm512_fmadd_pd(a1, b, c10);
        b = _mm512_set1_pd(B[16 * k + 1]);
        c01 = _mm512_fmadd_pd(a0, b, c01);
        c11 = _mm512_fmadd_pd(a1, b, c11);
        b = _mm512_set1_pd(B[16 * k + 2]);
        c02 = _mm512_fmadd_pd(a0, b, c02);
        c12 = _mm512_fmadd_pd(a1, b, c12);
        b = _mm512_set1_pd(B[16 * k + 3]);
        c03 = _mm512_fmadd_pd(a0, b, c03);
        c13 = _mm512_fmadd_pd(a1, b, c13);
    }
    _mm512_storeu_pd(C, _mm512_sub_pd(_mm512_loadu_pd(C), c00));
    _mm512_storeu_pd(C + 8, _mm512_sub_pd(_mm512_loadu_pd(C + 8), c10));
    _mm512_storeu_pd(C + ld, _mm512_sub_pd(_mm512_loadu_pd(C + ld), c01));
    _mm512_storeu_pd(C + ld + 8, _mm512_sub_pd(_mm512_loadu_pd(C + ld + 8), c11));
    _mm512_storeu_pd(C + 2 * ld, _mm512_sub_pd(_mm512_loadu_pd(C + 2 * ld), c02));
    _mm512_storeu_pd(C + 2 * ld + 8, _mm512_sub_pd(_mm512_loadu_pd(C + 2 * ld + 8), c12));
    _mm512_storeu_pd(C + 3 * ld, _mm512_sub_pd(_mm512_loadu_pd(C + 3 * ld), c03));
    _mm512_storeu_pd(C + 3 * ld + 8, _mm512_sub_pd(_mm512_loadu_pd(C + 3 * ld + 8), c13));
}

struct Kernel {
    int mr;
    TileFn tile;
    const char* name;
};

const Kernel& kernel() {
    static const Kernel k = []() -> Kernel {
        const char* e = std::getenv("ASGFEM_CHOL_ISA");
        const std::string want = e ? e : "";
        __builtin_cpu_init();
        if (want != "avx2" && want != "c" && __builtin_cpu_supports("avx512f")) return {16, tile_avx512, "avx512 16x4"};
        if (want != "c" && __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma")) return {8, tile_avx2, "avx2 8x4"};
        return {8, tile_c8, "portable 8x4"};
    }();
    return k;
}

constexpr int NB = 64;      // panel width
constexpr int STRIP = 64;   // rows of the trailing matrix per update task
constexpr int CHUNK = 128;  // rows per triangular-solve task

// unblocked Cholesky of the nb x nb lower triangle at d (column-major, ld); returns the first bad pivot or -1
__attribute__((target_clones("default", "avx2,fma", "avx512f"))) int diag_block(double* d, int nb, int64_t ld, const double* diag0) {
    int bad = -1;
    for (int j = 0; j < nb; ++j) {
        double* cj = d + (int64_t)j * ld;
        double piv = cj[j];
        if (!(piv > 1.0e-12 * diag0[j]) || !std::isfinite(piv)) {
            if (bad < 0) bad = j;
            piv = 1.0;
        }
        const double l = std::sqrt(piv), inv = 1.0 / l;
        cj[j] = l;
        for (int i = j + 1; i < nb; ++i) cj[i] *= inv;
        for (int c = j + 1; c < nb; ++c) {
            double* cc = d + (int64_t)c * ld;
            const double f = cj[c];
            for (int i = c; i < nb; ++i) cc[i] -= cj[i] * f;
        }
    }
    return bad;
}

// rows [r0, r1) of the panel below the diagonal block: X <- X L^{-T}, then the copy into the packed panel
// (x = first panel column at row 0 of the front, d = diagonal block, t0 = first trailing row)
__attribute__((target_clones("default", "avx2,fma", "avx512f"))) void solve_and_pack(double* x, const double* d, int nb, int64_t ld, int m,
                                                                                       int t0, int r0, int r1, int mr, double* pack) {
    const int re = std::min(r1, m);
    for (int j = 0; j < nb; ++j) {
        double* xj = x + (int64_t)j * ld;
        for (int c = 0; c < j; ++c) {
            const double* xc = x + (int64_t)c * ld;
            const double f = d[j + (int64_t)c * ld];
            for (int i = r0; i < re; ++i) xj[i] -= xc[i] * f;
        }
        const double inv = 1.0 / d[j + (int64_t)j * ld];
        for (int i = r0; i < re; ++i) xj[i] *= inv;
    }
    // packed panel: block b = (row - t0) / mr holds nb x mr values, [k][r]
    for (int b0 = r0; b0 < r1; b0 += mr) {
        double* p = pack + (int64_t)((b0 - t0) / mr) * mr * nb;
        const int rows = std::max(0, std::min(mr, m - b0));
        for (int k = 0; k < nb; ++k) {
            const double* xk = x + (int64_t)k * ld + b0;
            for (int r = 0; r < rows; ++r) p[k * mr + r] = xk[r];
            for (int r = rows; r < mr; ++r) p[k * mr + r] = 0.0;
        }
    }
}

// trailing update of the row strip [s0, s1) (indices relative to t0, multiples of mr): C[i, j] -= sum_k P[i, k] P[j, k], j <= i
void update_strip(double* c, int64_t ld, const double* pack, int nb, int s0, int s1, int mr, TileFn tile) {
    const int jend = s1;  // columns 0 .. s1-1 touch the lower triangle of this strip
    for (int jc = 0; jc < jend; jc += 64) {
        const int je = std::min(jc + 64, jend);
        for (int i = std::max(s0, jc / mr * mr); i < s1; i += mr) {
            const double* A = pack + (int64_t)(i / mr) * mr * nb;
            const int jlim = std::min(je, i + mr);  // tiles with j > i + mr - 1 lie above the diagonal
            for (int j = jc; j < jlim; j += 4) {
                const double* B = pack + (int64_t)(j / mr) * mr * nb + (j % mr);
                tile(A, B, nb, c + i + (int64_t)j * ld, ld);
            }
        }
    }
}

}  // namespace

int dense_front_slack() { return kernel().mr; }
const char* dense_kernel_name() { return kernel().name; }

int64_t dense_pack_size(int m, int s) {
    const int mr = kernel().mr;
    (void)s;
    return (int64_t)((m + mr - 1) / mr + 1) * mr * NB;
}

int dense_partial_cholesky(double* a, int m, int s, int64_t ld, const double* diag0, DensePool* pool, double* pack) {
    const Kernel& K = kernel();
    const int mr = K.mr;
    int bad = -1;
    for (int p0 = 0; p0 < s; p0 += NB) {
        const int nb = std::min(NB, s - p0);
        double* d = a + p0 + (int64_t)p0 * ld;
        const int b = diag_block(d, nb, ld, diag0 + p0);
        if (b >= 0 && bad < 0) bad = p0 + b;
        const int t0 = p0 + nb, mt = m - t0;
        if (mt <= 0) break;
        double* x = a + (int64_t)p0 * ld;  // panel columns, row 0 of the front
        const int mtp = (mt + mr - 1) / mr * mr;
        const int chunk = std::max(mr, CHUNK / mr * mr);
        const int nchunk = (mtp + chunk - 1) / chunk;
        auto solve = [&](int, int t) {
            const int r0 = t0 + t * chunk, r1 = std::min(t0 + mtp, r0 + chunk);
            solve_and_pack(x, d, nb, ld, m, t0, r0, r1, mr, pack);
        };
        const int strip = std::max(mr, STRIP / mr * mr);
        const int nstrip = (mtp + strip - 1) / strip;
        double* c = a + t0 + (int64_t)t0 * ld;
        auto update = [&](int, int t) {
            const int u = nstrip - 1 - t;  // longest strips first
            update_strip(c, ld, pack, nb, u * strip, std::min(mtp, (u + 1) * strip), mr, K.tile);
        };
        if (pool && mt >= 256) {
            dense_pool_run(pool, nchunk, solve);
            dense_pool_run(pool, nstrip, update);
        } else {
            for (int t = 0; t < nchunk; ++t) solve(0, t);
            for (int t = 0; t < nstrip; ++t) update(0, t);
        }
    }
    return bad;
}

}  // namespace asgfem
