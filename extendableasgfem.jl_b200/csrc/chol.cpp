// Host-side sparse Cholesky of the Dirichlet-reduced mean stiffness matrix K_0.
//
// Replaces `lu(A0.cscmatrix)` of MyPreconditionerPrimal (src/modelproblems/solvers_poisson_primal.jl:30-44):
// the reference pins boundary dofs with a 1e60 diagonal and hands the matrix to UMFPACK; here the boundary
// rows/columns are removed (the result on those rows is defined as exactly 0) and the remaining SPD matrix is
// factorised as P K P^T = L L^T once per refinement level:
//   1. fill-reducing ordering: graph nested dissection with BFS level-set separators (no METIS in the image),
//   2. elimination tree + column counts by row-subtree traversal,
//   3. numeric factorisation: multifrontal over the dissection tree (every leaf / separator is one relaxed supernode
//      whose frontal matrix is factorised by the blocked dense kernels of dense_chol.cpp; subtrees in parallel at the
//      bottom of the tree, threaded dense kernels at the top).  ASGFEM_CHOL_UPLOOKING=1 selects the scalar
//      up-looking sweep it replaced (kept as the cross-check of tools/chol_bench.cpp).
// The factor is returned row-wise (strictly lower part + inverse diagonal) for the device triangular solves.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <numeric>
#include <thread>

#include "common.h"
#include "dense_chol.h"

namespace asgfem {

namespace {

struct Graph {
    int32_t n;
    std::vector<int64_t> ptr;
    std::vector<int32_t> adj;
};

// BFS restricted to nodes with tag[node] == id; returns the level structure in `out` (concatenated), level starts in lv
int32_t bfs(const Graph& g, int32_t start, const std::vector<int32_t>& tag, int32_t id, std::vector<int32_t>& dist,
            std::vector<int32_t>& out, std::vector<int32_t>& lv) {
    out.clear();
    lv.clear();
    out.push_back(start);
    dist[start] = 0;
    lv.push_back(0);
    size_t head = 0;
    int32_t cur = 0;
    while (head < out.size()) {
        int32_t u = out[head];
        if (dist[u] != cur) {
            cur = dist[u];
            lv.push_back((int32_t)head);
        }
        ++head;
        for (int64_t p = g.ptr[u]; p < g.ptr[u + 1]; ++p) {
            int32_t v = g.adj[p];
            if (tag[v] == id && dist[v] < 0) {
                dist[v] = cur + 1;
                out.push_back(v);
            }
        }
    }
    lv.push_back((int32_t)out.size());
    return (int32_t)out.size();
}

// nested dissection ordering; returns perm (elimination order -> node).  The tree is built level by level with the
// subdomains of a level in parallel (they own disjoint node sets and disjoint ranges of perm).
void nested_dissection(const Graph& g, std::vector<int32_t>& perm, const double* xy, std::vector<BlockRec>& blocks,
                       int32_t max_block, DensePool* pool) {
    const int32_t n = g.n;
    // dissection stops at subdomains of <= LEAF unknowns (dense leaf blocks of the triangular sweeps); measured at config 4
    // (sptrsv.cu, DMMA sweeps): see DESIGN.md section 5
    int32_t LEAF = 24;
    if (const char* e = std::getenv("ASGFEM_CHOL_LEAF")) LEAF = std::max(4, atoi(e));
    perm.assign((size_t)n, -1);
    std::vector<int32_t> tag((size_t)n, 0), dist((size_t)n, -1);
    struct Task {
        std::vector<int32_t> nodes;
        int32_t lo;     // this set occupies perm[lo, lo + nodes.size())
        int32_t depth;  // depth in the dissection tree
    };
    // blocks for the tree-parallel triangular solves: every leaf and every separator of the dissection tree is a
    // block (separators longer than max_block are cut into sequentially dependent chunks)
    blocks.clear();
    struct Scratch {
        std::vector<int32_t> bfsout, lv, tmp;
        std::vector<BlockRec> blocks;
        std::vector<Task> out;
    };
    std::vector<Scratch> scratch((size_t)dense_pool_threads(pool));
    std::vector<Task> frontier;
    {
        Task t;
        t.nodes.resize((size_t)n);
        std::iota(t.nodes.begin(), t.nodes.end(), 0);
        t.lo = 0;
        t.depth = 0;
        frontier.push_back(std::move(t));
    }
    std::atomic<int32_t> next_id{1};
    auto process = [&](Task& t, Scratch& sc) {
        std::vector<int32_t>&bfsout = sc.bfsout, &lv = sc.lv, &tmp = sc.tmp;
        std::vector<Task>& stack = sc.out;
        auto emit_range = [&](int32_t lo, int32_t len, int32_t depth) {
            if (len <= 0) return;
            int32_t nch = (len + max_block - 1) / max_block;
            for (int32_t c = 0; c < nch; ++c)
                sc.blocks.push_back({lo + c * max_block, std::min(max_block, len - c * max_block), depth, c, nch});
        };
        const int32_t sz = (int32_t)t.nodes.size();
        if (sz == 0) return;
        if (sz <= LEAF) {
            for (int32_t k = 0; k < sz; ++k) perm[t.lo + k] = t.nodes[k];
            emit_range(t.lo, sz, t.depth);
            return;
        }
        const int32_t id = next_id.fetch_add(1);
        for (int32_t v : t.nodes) {
            tag[v] = id;
            dist[v] = -1;
        }
        if (xy) {
            // geometric separator: cut the bounding box at the median of its longer axis; the separator is the layer
            // of nodes on the upper side that touch the lower side (straight lines on mesh-like point sets)
            double lo[2] = {1e300, 1e300}, hi[2] = {-1e300, -1e300};
            for (int32_t v : t.nodes)
                for (int d = 0; d < 2; ++d) {
                    lo[d] = std::min(lo[d], xy[2 * v + d]);
                    hi[d] = std::max(hi[d], xy[2 * v + d]);
                }
            const int ax = (hi[0] - lo[0]) >= (hi[1] - lo[1]) ? 0 : 1;
            if (hi[ax] > lo[ax]) {
                tmp.assign(t.nodes.begin(), t.nodes.end());
                auto midit = tmp.begin() + sz / 2;
                std::nth_element(tmp.begin(), midit, tmp.end(), [&](int32_t a, int32_t b) {
                    return xy[2 * a + ax] != xy[2 * b + ax] ? xy[2 * a + ax] < xy[2 * b + ax] : a < b;
                });
                const double cut = xy[2 * (*midit) + ax];
                // dist doubles as side marker: 0 = lower side (coordinate < cut), 1 = upper side
                int32_t nlow = 0;
                for (int32_t v : t.nodes) {
                    dist[v] = xy[2 * v + ax] < cut ? 0 : 1;
                    nlow += dist[v] == 0;
                }
                if (nlow > 0 && nlow < sz) {
                    Task a, b;
                    std::vector<int32_t> sep;
                    for (int32_t v : t.nodes) {
                        if (dist[v] == 0) {
                            a.nodes.push_back(v);
                            continue;
                        }
                        bool touches = false;
                        for (int64_t p = g.ptr[v]; p < g.ptr[v + 1] && !touches; ++p) {
                            int32_t u = g.adj[p];
                            touches = tag[u] == id && dist[u] == 0;
                        }
                        (touches ? sep : b.nodes).push_back(v);
                    }
                    if (!sep.empty() || a.nodes.empty() || b.nodes.empty()) {
                        a.lo = t.lo;
                        b.lo = t.lo + (int32_t)a.nodes.size();
                        int32_t seplo = b.lo + (int32_t)b.nodes.size();
                        for (size_t k = 0; k < sep.size(); ++k) perm[seplo + (int32_t)k] = sep[k];
                        for (int32_t v : t.nodes) dist[v] = -1;
                        a.depth = b.depth = t.depth + 1;
                        emit_range(seplo, (int32_t)sep.size(), t.depth);
                        stack.push_back(std::move(a));
                        stack.push_back(std::move(b));
                        return;
                    }
                    // the two sides are not connected: fall through to the component split below
                }
            }
            for (int32_t v : t.nodes) dist[v] = -1;
        }
        // connected component of the first node
        int32_t cnt = bfs(g, t.nodes[0], tag, id, dist, bfsout, lv);
        if (cnt < sz) {  // disconnected: split off the component, no separator needed
            Task a, b;
            a.nodes = bfsout;
            for (int32_t v : a.nodes) tag[v] = -id;
            for (int32_t v : t.nodes)
                if (tag[v] == id) b.nodes.push_back(v);
            a.lo = t.lo;
            b.lo = t.lo + (int32_t)a.nodes.size();
            a.depth = b.depth = t.depth + 1;
            stack.push_back(std::move(a));
            stack.push_back(std::move(b));
            return;
        }
        // pseudo-peripheral start: two more sweeps from the farthest node
        for (int sweep = 0; sweep < 2; ++sweep) {
            int32_t far = bfsout.back();
            for (int32_t v : t.nodes) dist[v] = -1;
            bfs(g, far, tag, id, dist, bfsout, lv);
        }
        int32_t nlev = (int32_t)lv.size() - 1;
        if (nlev < 3) {  // (nearly) complete graph: no useful separator
            for (int32_t k = 0; k < sz; ++k) perm[t.lo + k] = t.nodes[k];
            emit_range(t.lo, sz, t.depth);
            return;
        }
        // separator = the level whose removal balances the two sides best
        int32_t best = 1;
        int64_t bestcost = INT64_MAX;
        for (int32_t l = 1; l < nlev - 1; ++l) {
            int64_t a = lv[l], s = lv[l + 1] - lv[l], b = sz - lv[l + 1];
            int64_t cost = std::llabs(a - b) + 4 * s;  // balance + separator size
            if (cost < bestcost) {
                bestcost = cost;
                best = l;
            }
        }
        Task a, b;
        a.nodes.assign(bfsout.begin(), bfsout.begin() + lv[best]);
        b.nodes.assign(bfsout.begin() + lv[best + 1], bfsout.end());
        int32_t nsep = lv[best + 1] - lv[best];
        a.lo = t.lo;
        b.lo = t.lo + (int32_t)a.nodes.size();
        int32_t seplo = b.lo + (int32_t)b.nodes.size();
        for (int32_t k = 0; k < nsep; ++k) perm[seplo + k] = bfsout[lv[best] + k];
        a.depth = b.depth = t.depth + 1;
        emit_range(seplo, nsep, t.depth);
        stack.push_back(std::move(a));
        stack.push_back(std::move(b));
    };
    while (!frontier.empty()) {
        dense_pool_run(pool, (int)frontier.size(), [&](int th, int i) { process(frontier[(size_t)i], scratch[(size_t)th]); });
        frontier.clear();
        for (Scratch& sc : scratch) {
            for (Task& t : sc.out) frontier.push_back(std::move(t));
            sc.out.clear();
        }
        // big subdomains first: the pool hands tasks out in order
        std::sort(frontier.begin(), frontier.end(), [](const Task& x, const Task& y) { return x.nodes.size() > y.nodes.size(); });
    }
    for (Scratch& sc : scratch) blocks.insert(blocks.end(), sc.blocks.begin(), sc.blocks.end());
    std::sort(blocks.begin(), blocks.end(), [](const BlockRec& x, const BlockRec& y) { return x.start < y.start; });
}

}  // namespace

namespace {
// phase timing on stderr with ASGFEM_CHOL_VERBOSE=1
struct CholTick {
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    bool on = std::getenv("ASGFEM_CHOL_VERBOSE") != nullptr;
    void operator()(const char* what) {
        auto now = std::chrono::steady_clock::now();
        if (on) fprintf(stderr, "[chol] %-28s %.3f s\n", what, std::chrono::duration<double>(now - t).count());
        t = now;
    }
};
}  // namespace

namespace {

// One node of the dissection tree = one relaxed supernode: columns [lo, hi) of the elimination order, treated as dense.
struct Front {
    int32_t lo, hi;
    int32_t parent = -1;            // front that holds the smallest row of `rows`
    std::vector<int32_t> rows;      // rows k >= hi with L[k, j] != 0 for some column j of the front (sorted)
    std::vector<int32_t> children;
    std::vector<double> update;     // Schur complement on `rows` (rows.size()^2, column-major lower triangle)
    double work = 0.0, subtree = 0.0;
};

// Multifrontal numeric phase.  C = P A P^T by columns of its upper triangle (Cp, Ci, Cx); F.Lp / F.Li hold the exact,
// column-sorted row patterns of L and F.Lx / F.dinv receive the values.  Every front assembles the entries of A in its
// columns plus the update matrices of its children (extend-add through the global-row -> front-row map), factorises its
// columns with dense_partial_cholesky and leaves its own update matrix for the parent.  Entries of the dense front that
// are structurally zero in L stay exactly zero and are simply not copied out.
int32_t multifrontal_numeric(int32_t n, const std::vector<int64_t>& Cp, const std::vector<int32_t>& Ci, const std::vector<double>& Cx,
                             std::vector<Front>& fronts, CholFactor& F, DensePool* pool) {
    const int nthreads = dense_pool_threads(pool);
    const int slack = dense_front_slack();
    // strict lower triangle by columns (= transpose of the strict upper triangle held in C) and the diagonal
    std::vector<int64_t> Tp((size_t)n + 1, 0);
    std::vector<double> diag((size_t)n, 0.0);
    for (int32_t k = 0; k < n; ++k)
        for (int64_t p = Cp[k]; p < Cp[k + 1]; ++p)
            if (Ci[p] < k) Tp[(size_t)Ci[p] + 1]++;
    for (int32_t k = 0; k < n; ++k) Tp[(size_t)k + 1] += Tp[(size_t)k];
    std::vector<int32_t> Ti((size_t)Tp[(size_t)n]);
    std::vector<double> Tx((size_t)Tp[(size_t)n]);
    {
        std::vector<int64_t> at(Tp.begin(), Tp.end() - 1);
        for (int32_t k = 0; k < n; ++k)
            for (int64_t p = Cp[k]; p < Cp[k + 1]; ++p) {
                if (Ci[p] == k) {
                    diag[(size_t)k] += Cx[p];
                } else if (Ci[p] < k) {
                    const int64_t q = at[(size_t)Ci[p]]++;
                    Ti[(size_t)q] = k;
                    Tx[(size_t)q] = Cx[p];
                }
            }
    }
    struct Scratch {
        std::vector<double> front, pack, diag0;
        std::vector<int32_t> pos, idx;
    };
    std::vector<Scratch> scratch((size_t)nthreads);
    for (Scratch& sc : scratch) sc.pos.assign((size_t)n, 0);
    std::atomic<int32_t> bad_pivot{-1};

    auto process = [&](int32_t t, int th, DensePool* pl) {
        Front& fr = fronts[(size_t)t];
        Scratch& sc = scratch[(size_t)th];
        const int32_t s = fr.hi - fr.lo, r = (int32_t)fr.rows.size(), m = s + r;
        int64_t ld = ((int64_t)m + slack + 7) / 8 * 8;
        if (ld % 512 == 0) ld += 8;  // keep the columns of big fronts off the same cache sets
        const size_t need = (size_t)ld * (size_t)(m + slack);
        if (sc.front.size() < need) sc.front.resize(need);
        const int64_t npack = dense_pack_size(m, s);
        if ((int64_t)sc.pack.size() < npack) sc.pack.resize((size_t)npack);
        double* a = sc.front.data();
        int32_t* pos = sc.pos.data();
        for (int32_t i = 0; i < s; ++i) pos[fr.lo + i] = i;
        for (int32_t i = 0; i < r; ++i) pos[fr.rows[(size_t)i]] = s + i;
        // zero + entries of A, by columns
        const int32_t cchunk = 64, ncchunk = (m + cchunk - 1) / cchunk;
        auto init = [&](int, int c) {
            const int32_t c0 = c * cchunk, c1 = std::min(m, c0 + cchunk);
            std::memset(a + (int64_t)c0 * ld, 0, sizeof(double) * (size_t)ld * (size_t)(c1 - c0));
            for (int32_t j = c0; j < std::min(c1, s); ++j) {
                double* col = a + (int64_t)j * ld;
                col[j] = diag[(size_t)(fr.lo + j)];
                for (int64_t p = Tp[(size_t)(fr.lo + j)]; p < Tp[(size_t)(fr.lo + j) + 1]; ++p) col[pos[Ti[(size_t)p]]] += Tx[(size_t)p];
            }
        };
        if (pl && m >= 512)
            dense_pool_run(pl, ncchunk, init);
        else
            for (int32_t c = 0; c < ncchunk; ++c) init(0, c);
        // extend-add of the children
        for (int32_t ch : fr.children) {
            Front& cf = fronts[(size_t)ch];
            const int32_t rc = (int32_t)cf.rows.size();
            if (rc == 0) continue;
            if ((int32_t)sc.idx.size() < rc) sc.idx.resize((size_t)rc);
            int32_t* idx = sc.idx.data();
            for (int32_t i = 0; i < rc; ++i) idx[i] = pos[cf.rows[(size_t)i]];
            const double* u = cf.update.data();
            auto add = [&](int, int c) {
                const int32_t b0 = c * cchunk, b1 = std::min(rc, b0 + cchunk);
                for (int32_t b = b0; b < b1; ++b) {
                    double* col = a + (int64_t)idx[b] * ld;
                    const double* ub = u + (int64_t)b * rc;
                    for (int32_t i = b; i < rc; ++i) col[idx[i]] += ub[i];
                }
            };
            const int32_t nadd = (rc + cchunk - 1) / cchunk;
            if (pl && rc >= 512)
                dense_pool_run(pl, nadd, add);
            else
                for (int32_t c = 0; c < nadd; ++c) add(0, c);
            std::vector<double>().swap(cf.update);
        }
        if ((int32_t)sc.diag0.size() < s) sc.diag0.resize((size_t)s);
        for (int32_t i = 0; i < s; ++i) sc.diag0[(size_t)i] = std::fabs(diag[(size_t)(fr.lo + i)]);
        const int bad = dense_partial_cholesky(a, m, s, ld, sc.diag0.data(), pl, sc.pack.data());
        if (bad >= 0) {
            int32_t expect = -1;
            bad_pivot.compare_exchange_strong(expect, fr.lo + bad);
        }
        // values of L in the (exact) row patterns; the update matrix for the parent
        if (r > 0) fr.update.resize((size_t)r * (size_t)r);
        double* u = fr.update.data();
        const int32_t rchunk = 64, nrchunk = (m + rchunk - 1) / rchunk;
        auto extract = [&](int, int c) {
            const int32_t i0 = c * rchunk, i1 = std::min(m, i0 + rchunk);
            for (int32_t i = i0; i < i1; ++i) {
                if (i < s) {
                    const int32_t k = fr.lo + i;
                    F.dinv[(size_t)k] = 1.0 / a[i + (int64_t)i * ld];
                    for (int64_t p = F.Lp[(size_t)k + 1] - 1; p >= F.Lp[(size_t)k] && F.Li[(size_t)p] >= fr.lo; --p)
                        F.Lx[(size_t)p] = a[i + (int64_t)(F.Li[(size_t)p] - fr.lo) * ld];
                } else {
                    const int32_t k = fr.rows[(size_t)(i - s)];
                    const int32_t* b = F.Li.data() + F.Lp[(size_t)k];
                    const int32_t* e = F.Li.data() + F.Lp[(size_t)k + 1];
                    for (const int32_t* q = std::lower_bound(b, e, fr.lo); q < e && *q < fr.hi; ++q)
                        F.Lx[(size_t)(q - F.Li.data())] = a[i + (int64_t)(*q - fr.lo) * ld];
                    const int32_t bcol = i - s;  // column bcol of the update matrix: rows bcol..r-1
                    std::memcpy(u + (int64_t)bcol * r + bcol, a + i + (int64_t)i * ld, sizeof(double) * (size_t)(r - bcol));
                }
            }
        };
        if (pl && m >= 512)
            dense_pool_run(pl, nrchunk, extract);
        else
            for (int32_t c = 0; c < nrchunk; ++c) extract(0, c);
    };

    // subtrees below the work threshold go to one thread each; what remains on top runs front by front, threaded inside
    const int32_t nf = (int32_t)fronts.size();
    double total = 0.0;
    for (int32_t t = 0; t < nf; ++t) {
        Front& fr = fronts[(size_t)t];
        const double s = fr.hi - fr.lo, m = s + (double)fr.rows.size();
        fr.work = s * m * m - s * s * m + s * s * s / 3.0 + 50.0 * m;
        fr.subtree += fr.work;
        if (fr.parent >= 0) fronts[(size_t)fr.parent].subtree += fr.subtree;
        total += fr.work;
    }
    double div = 8.0;
    if (const char* e = std::getenv("ASGFEM_CHOL_SPLIT")) div = std::max(0.25, atof(e));
    const double thresh = nthreads > 1 ? total / (div * nthreads) : 2.0 * total;
    std::vector<int32_t> group((size_t)nf, -1), roots;
    for (int32_t t = nf - 1; t >= 0; --t) {
        const Front& fr = fronts[(size_t)t];
        if (fr.subtree > thresh) continue;  // top
        if (fr.parent < 0 || group[(size_t)fr.parent] < 0) {
            group[(size_t)t] = (int32_t)roots.size();
            roots.push_back(t);
        } else {
            group[(size_t)t] = group[(size_t)fr.parent];
        }
    }
    std::vector<std::vector<int32_t>> members(roots.size());
    for (int32_t t = 0; t < nf; ++t)
        if (group[(size_t)t] >= 0) members[(size_t)group[(size_t)t]].push_back(t);
    std::vector<int32_t> order(roots.size());
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return fronts[(size_t)roots[(size_t)x]].subtree > fronts[(size_t)roots[(size_t)y]].subtree; });
    const auto t_sub = std::chrono::steady_clock::now();
    dense_pool_run(pool, (int)order.size(), [&](int th, int g) {
        for (int32_t t : members[(size_t)order[(size_t)g]]) process(t, th, nullptr);
    });
    const auto t_top = std::chrono::steady_clock::now();
    int32_t ntop = 0;
    for (int32_t t = 0; t < nf; ++t)
        if (group[(size_t)t] < 0) {
            process(t, 0, pool);
            ++ntop;
        }
    if (std::getenv("ASGFEM_CHOL_VERBOSE")) {
        const double s_sub = std::chrono::duration<double>(t_top - t_sub).count();
        const double s_top = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_top).count();
        fprintf(stderr, "[chol] multifrontal: %d fronts, %.1f GFLOP (%s, %d threads): %zu subtrees %.3f s, %d top fronts %.3f s\n", nf,
                total * 1.0e-9, dense_kernel_name(), nthreads, roots.size(), s_sub, ntop, s_top);
    }
    return bad_pivot.load();
}

}  // namespace

int cholesky_reduced(int64_t n_full, const int64_t* rowptr, const int32_t* col, const double* val,
                     const uint8_t* is_boundary, const double* coords_full, int32_t max_block, CholFactor& F,
                     std::string& err) {
    CholTick chol_tick;
    // ---- reduced numbering -----------------------------------------------------------------------
    std::vector<int32_t> red((size_t)n_full, -1), full;
    for (int64_t i = 0; i < n_full; ++i)
        if (!is_boundary[i]) {
            red[i] = (int32_t)full.size();
            full.push_back((int32_t)i);
        }
    const int32_t n = (int32_t)full.size();
    F.n = n;
    F.perm.clear();
    F.Lp.assign(1, 0);
    F.Li.clear();
    F.Lx.clear();
    F.dinv.clear();
    F.node_lo.clear(), F.node_hi.clear(), F.node_rows.clear();
    F.node_rptr.assign(1, 0);
    if (n == 0) return 0;

    int nthreads = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u);
    if (const char* e = std::getenv("ASGFEM_CHOL_THREADS")) nthreads = std::max(1, atoi(e));
    if (n < 20000) nthreads = 1;
    DensePool* pool = dense_pool_create(nthreads);
    struct PoolGuard {
        DensePool* p;
        ~PoolGuard() { dense_pool_destroy(p); }
    } pool_guard{pool};

    // loops over the rows of a matrix, in parallel over chunks of rows
    auto for_rows = [&](int32_t nrows, const std::function<void(int32_t, int32_t)>& body) {
        const int32_t nchunk = (int32_t)std::min<int64_t>(std::max<int32_t>(nrows, 1), 8 * (int64_t)nthreads);
        dense_pool_run(pool, nchunk, [&](int, int c) {
            body((int32_t)((int64_t)nrows * c / nchunk), (int32_t)((int64_t)nrows * (c + 1) / nchunk));
        });
    };
    Graph g;
    g.n = n;
    g.ptr.assign((size_t)n + 1, 0);
    for_rows(n, [&](int32_t r0, int32_t r1) {
        for (int32_t r = r0; r < r1; ++r) {
            int64_t i = full[r], cnt = 0;
            for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p) cnt += red[col[p]] >= 0 && col[p] != i;
            g.ptr[r + 1] = cnt;
        }
    });
    for (int32_t r = 0; r < n; ++r) g.ptr[r + 1] += g.ptr[r];
    g.adj.resize((size_t)g.ptr[n]);
    for_rows(n, [&](int32_t r0, int32_t r1) {
        for (int32_t r = r0; r < r1; ++r) {
            int64_t i = full[r], q = g.ptr[r];
            for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p)
                if (red[col[p]] >= 0 && col[p] != i) g.adj[q++] = red[col[p]];
        }
    });

    std::vector<int32_t> perm;
    std::vector<double> xy;
    if (coords_full) {
        xy.resize((size_t)2 * n);
        for (int32_t r = 0; r < n; ++r) {
            xy[2 * r] = coords_full[2 * (int64_t)full[r]];
            xy[2 * r + 1] = coords_full[2 * (int64_t)full[r] + 1];
        }
    }
    chol_tick("graph");
    nested_dissection(g, perm, coords_full ? xy.data() : nullptr, F.blocks, max_block, pool);
    chol_tick("nested dissection");
    {
        int32_t at = 0;
        for (const BlockRec& b : F.blocks) {
            if (b.start != at) {
                err = "internal error: dissection blocks do not tile the elimination order";
                return ASGFEM_ENUMERIC;
            }
            at += b.len;
        }
        if (at != n) {
            err = "internal error: dissection blocks do not cover all unknowns";
            return ASGFEM_ENUMERIC;
        }
    }
    std::vector<int32_t> iperm((size_t)n);
    for (int32_t k = 0; k < n; ++k) {
        if (perm[k] < 0) {
            err = "internal error: incomplete ordering";
            return ASGFEM_ENUMERIC;
        }
        iperm[perm[k]] = k;
    }

    // ---- permuted upper triangle by columns: C = P A P^T, column k holds rows i <= k ----------------
    std::vector<int64_t> Cp((size_t)n + 1, 0);
    for_rows(n, [&](int32_t k0, int32_t k1) {
        for (int32_t k = k0; k < k1; ++k) {
            int64_t i = full[perm[k]], cnt = 0;
            for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p) {
                int32_t rj = red[col[p]];
                cnt += rj >= 0 && iperm[rj] <= k;
            }
            Cp[k + 1] = cnt;
        }
    });
    for (int32_t k = 0; k < n; ++k) Cp[k + 1] += Cp[k];
    std::vector<int32_t> Ci((size_t)Cp[n]);
    std::vector<double> Cx((size_t)Cp[n]);
    for_rows(n, [&](int32_t k0, int32_t k1) {
        for (int32_t k = k0; k < k1; ++k) {
            int64_t i = full[perm[k]], q = Cp[k];
            for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p) {
                int32_t rj = red[col[p]];
                if (rj >= 0 && iperm[rj] <= k) {
                    Ci[q] = iperm[rj];
                    Cx[q] = val[p];  // A symmetric: A[i, j] used as C[j', k]
                    ++q;
                }
            }
        }
    });

    chol_tick("permuted matrix");
    // ---- elimination tree (Liu) ---------------------------------------------------------------------
    std::vector<int32_t> parent((size_t)n, -1), anc((size_t)n, -1);
    for (int32_t k = 0; k < n; ++k)
        for (int64_t p = Cp[k]; p < Cp[k + 1]; ++p) {
            int32_t i = Ci[p];
            while (i != -1 && i < k) {
                int32_t nxt = anc[i];
                anc[i] = k;
                if (nxt == -1) parent[i] = k;
                i = nxt;
            }
        }

    chol_tick("elimination tree");
    std::vector<int32_t> rowlen((size_t)n, 0);
    // ---- symbolic + numeric factorisation: multifrontal (default) ---------------------------------------------------------
    const bool uplooking = std::getenv("ASGFEM_CHOL_UPLOOKING") != nullptr;
    if (!uplooking) {
        // fronts of the multifrontal phase: the nodes of the dissection tree (chunks of one separator joined again)
        std::vector<Front> fronts;
        for (const BlockRec& b : F.blocks) {
            if (!fronts.empty() && b.chunk > 0 && fronts.back().hi == b.start)
                fronts.back().hi = b.start + b.len;
            else {
                fronts.emplace_back();
                fronts.back().lo = b.start;
                fronts.back().hi = b.start + b.len;
            }
        }
        std::vector<int32_t> frontof((size_t)n);
        for (int32_t t = 0; t < (int32_t)fronts.size(); ++t)
            for (int32_t k = fronts[(size_t)t].lo; k < fronts[(size_t)t].hi; ++k) frontof[(size_t)k] = t;
        chol_tick("fronts");
        // Row patterns by row-subtree traversal, in parallel over chunks of rows.  The pattern of row k is the union of
        // the tree paths that start at the entries of row k of A; every path ascends, so the pattern is sorted by merging
        // these runs.  Pass 1 counts and records which fronts a row belongs to, pass 2 writes the sorted patterns.
        struct Sym {
            std::vector<int32_t> flag, stack, tmp, seen;
            std::vector<int32_t> runs;
        };
        std::vector<Sym> sym((size_t)nthreads);
        const int32_t nchunk = (int32_t)std::min<int64_t>(n, 8 * (int64_t)nthreads);
        std::vector<std::vector<std::pair<int32_t, int32_t>>> member((size_t)nchunk);  // (front, row) per chunk, rows ascending
        auto row_pattern = [&](Sym& w, int32_t k, int32_t tag) -> int32_t {  // unsorted pattern in w.stack, run starts in w.runs
            int32_t* fl = w.flag.data();
            int32_t* st = w.stack.data();
            int32_t len = 0;
            w.runs.clear();
            fl[k] = tag;
            for (int64_t p = Cp[k]; p < Cp[k + 1]; ++p) {
                const int32_t before = len;
                for (int32_t i = Ci[p]; i < k && fl[i] != tag; i = parent[i]) {
                    st[len++] = i;
                    fl[i] = tag;
                }
                if (len > before) w.runs.push_back(before);
            }
            return len;
        };
        auto sym_init = [&](Sym& w) {
            if (w.flag.empty()) {
                w.flag.assign((size_t)n, -1);
                w.stack.resize((size_t)n);
                w.tmp.resize((size_t)n);
                w.seen.assign(fronts.size(), -1);
            }
        };
        dense_pool_run(pool, nthreads, [&](int th, int) { sym_init(sym[(size_t)th]); });
        chol_tick("symbolic (scratch)");
        dense_pool_run(pool, nchunk, [&](int th, int c) {
            Sym& w = sym[(size_t)th];
            sym_init(w);
            const int32_t k0 = (int32_t)((int64_t)n * c / nchunk), k1 = (int32_t)((int64_t)n * (c + 1) / nchunk);
            for (int32_t k = k0; k < k1; ++k) {
                const int32_t len = row_pattern(w, k, k);
                rowlen[k] = len;
                const int32_t own = frontof[(size_t)k];
                int32_t cur_lo = fronts[(size_t)own].lo, cur_hi = fronts[(size_t)own].hi;  // columns known to need no lookup
                for (int32_t q = 0; q < len; ++q) {
                    const int32_t i = w.stack[(size_t)q];
                    if (i >= cur_lo && i < cur_hi) continue;  // tree paths ascend: mostly the front of the previous column
                    const int32_t t = frontof[(size_t)i];
                    cur_lo = fronts[(size_t)t].lo, cur_hi = fronts[(size_t)t].hi;
                    if (t != own && w.seen[(size_t)t] != k) {
                        w.seen[(size_t)t] = k;
                        member[(size_t)c].push_back({t, k});
                    }
                }
            }
        });
        chol_tick("symbolic (row counts, parallel part)");
        for (int32_t c = 0; c < nchunk; ++c) {
            for (const std::pair<int32_t, int32_t>& tk : member[(size_t)c]) fronts[(size_t)tk.first].rows.push_back(tk.second);
            std::vector<std::pair<int32_t, int32_t>>().swap(member[(size_t)c]);
        }
        chol_tick("symbolic (row counts)");
        int64_t lnz = 0;
        try {
            F.Lp.assign((size_t)n + 1, 0);
            for (int32_t k = 0; k < n; ++k) F.Lp[k + 1] = F.Lp[k] + rowlen[k];
            lnz = F.Lp[(size_t)n];
            F.Li.resize((size_t)lnz);  // not value-initialised: first touched by the threads that fill them
            F.Lx.resize((size_t)lnz);
            F.dinv.resize((size_t)n);
        } catch (const std::bad_alloc&) {
            err = "out of host memory for the Cholesky factor";
            return ASGFEM_ENOMEM;
        }
        // parent = the front of the smallest outside row; the rows of a front must be covered by its parent (columns
        // or rows), which holds for dissection orderings and is enforced here for anything else (explicit zeros)
        std::vector<int32_t> merged;
        for (int32_t t = 0; t < (int32_t)fronts.size(); ++t) {
            Front& fr = fronts[(size_t)t];
            if (fr.rows.empty()) continue;
            fr.parent = frontof[(size_t)fr.rows[0]];
            Front& pf = fronts[(size_t)fr.parent];
            pf.children.push_back(t);
            auto first_out = std::lower_bound(fr.rows.begin(), fr.rows.end(), pf.hi);
            if (!std::includes(pf.rows.begin(), pf.rows.end(), first_out, fr.rows.end())) {
                merged.clear();
                std::set_union(pf.rows.begin(), pf.rows.end(), first_out, fr.rows.end(), std::back_inserter(merged));
                pf.rows = merged;
            }
        }
        F.node_lo.clear(), F.node_hi.clear(), F.node_rows.clear();
        F.node_rptr.assign(1, 0);
        for (const Front& fr : fronts) {
            F.node_lo.push_back(fr.lo);
            F.node_hi.push_back(fr.hi);
            F.node_rows.insert(F.node_rows.end(), fr.rows.begin(), fr.rows.end());
            F.node_rptr.push_back((int64_t)F.node_rows.size());
        }
        chol_tick("front tree");
        dense_pool_run(pool, nchunk, [&](int th, int c) {
            Sym& w = sym[(size_t)th];
            sym_init(w);
            const int32_t k0 = (int32_t)((int64_t)n * c / nchunk), k1 = (int32_t)((int64_t)n * (c + 1) / nchunk);
            for (int32_t k = k0; k < k1; ++k) {
                const int32_t len = row_pattern(w, k, n + k);  // tags distinct from those of pass 1
                // merge the ascending runs pairwise until one is left
                int32_t* src = w.stack.data();
                int32_t* dst = w.tmp.data();
                std::vector<int32_t>& runs = w.runs;
                runs.push_back(len);
                while (runs.size() > 2) {
                    size_t nr = 0;
                    for (size_t q = 0; q + 1 < runs.size(); q += 2) {
                        const int32_t a0 = runs[q], a1 = runs[q + 1], b1 = q + 2 < runs.size() ? runs[q + 2] : a1;
                        std::merge(src + a0, src + a1, src + a1, src + b1, dst + a0);
                        runs[nr++] = a0;
                    }
                    runs[nr++] = len;
                    runs.resize(nr);
                    std::swap(src, dst);
                }
                std::copy(src, src + len, F.Li.begin() + F.Lp[k]);
            }
        });
        chol_tick("symbolic (row patterns)");
        const int32_t bad = multifrontal_numeric(n, Cp, Ci, Cx, fronts, F, pool);
        chol_tick("numeric (multifrontal)");
        if (bad >= 0) {
            err = "K_0 restricted to the interior dofs is not positive definite (pivot " + std::to_string(bad) + ")";
            return ASGFEM_ENUMERIC;
        }
        F.perm.resize((size_t)n);
        for (int32_t k = 0; k < n; ++k) F.perm[k] = full[perm[k]];
        return 0;
    }
    // ---- cross-check path (ASGFEM_CHOL_UPLOOKING=1): scalar up-looking factorisation -----------------------------------
    // row patterns by row-subtree traversal (ereach); pass 1 counts, pass 2 factorises
    std::vector<int32_t> flag((size_t)n, -1), stack((size_t)n);
    std::vector<int64_t> colcount((size_t)n, 1);  // diagonal
    auto ereach = [&](int32_t k, int32_t& top) {
        top = n;
        flag[k] = k;
        for (int64_t p = Cp[k]; p < Cp[k + 1]; ++p) {
            int32_t i = Ci[p];
            if (i > k) continue;
            int32_t len = 0;
            for (; flag[i] != k; i = parent[i]) {
                stack[len++] = i;
                flag[i] = k;
            }
            while (len > 0) stack[--top] = stack[--len];
        }
    };
    int64_t lnz = 0;
    {
        F.node_lo.clear(), F.node_hi.clear();
        for (const BlockRec& b : F.blocks) {
            if (!F.node_lo.empty() && b.chunk > 0 && F.node_hi.back() == b.start)
                F.node_hi.back() = b.start + b.len;
            else {
                F.node_lo.push_back(b.start);
                F.node_hi.push_back(b.start + b.len);
            }
        }
        const size_t nn = F.node_lo.size();
        std::vector<int32_t> nodeof((size_t)n), seen(nn, -1);
        for (size_t t = 0; t < nn; ++t)
            for (int32_t k = F.node_lo[t]; k < F.node_hi[t]; ++k) nodeof[(size_t)k] = (int32_t)t;
        std::vector<std::vector<int32_t>> rows(nn);
        for (int32_t k = 0; k < n; ++k) {
            int32_t top;
            ereach(k, top);
            rowlen[k] = n - top;
            lnz += n - top;
            for (int32_t q = top; q < n; ++q) {
                colcount[stack[q]]++;
                const int32_t t = nodeof[(size_t)stack[q]];
                if (t != nodeof[(size_t)k] && seen[(size_t)t] != k) {
                    seen[(size_t)t] = k;
                    rows[(size_t)t].push_back(k);
                }
            }
        }
        F.node_rows.clear();
        F.node_rptr.assign(1, 0);
        for (size_t t = 0; t < nn; ++t) {
            F.node_rows.insert(F.node_rows.end(), rows[t].begin(), rows[t].end());
            F.node_rptr.push_back((int64_t)F.node_rows.size());
        }
    }
    chol_tick("symbolic (row counts)");
    // column storage for the numeric phase (diagonal first), row storage for the output
    std::vector<int64_t> Lcp((size_t)n + 1, 0);
    for (int32_t j = 0; j < n; ++j) Lcp[j + 1] = Lcp[j] + colcount[j];
    std::vector<int32_t> Lci;
    std::vector<double> Lcx;
    try {
        Lci.resize((size_t)Lcp[n]);
        Lcx.resize((size_t)Lcp[n]);
        F.Lp.assign((size_t)n + 1, 0);
        for (int32_t k = 0; k < n; ++k) F.Lp[k + 1] = F.Lp[k] + rowlen[k];
        F.Li.resize((size_t)lnz);
        F.Lx.resize((size_t)lnz);
        F.dinv.resize((size_t)n);
    } catch (const std::bad_alloc&) {
        err = "out of host memory for the Cholesky factor";
        return ASGFEM_ENOMEM;
    }
    std::vector<int64_t> cnext(Lcp.begin(), Lcp.end() - 1);
    // Rows of different nodes of the dissection tree at the same depth reach disjoint sets of columns (their subtrees),
    // so the up-looking sweep runs level by level from the deepest nodes to the root with the nodes of a level in
    // parallel (threads with private work vectors); the chunks of one separator stay in order inside one task.
    struct Task {
        int32_t lo, hi, depth;
    };
    std::vector<Task> tasks;
    int32_t maxdepth = 0;
    for (const BlockRec& b : F.blocks) {
        if (!tasks.empty() && b.chunk > 0 && tasks.back().hi == b.start && tasks.back().depth == b.depth)
            tasks.back().hi = b.start + b.len;
        else
            tasks.push_back({b.start, b.start + b.len, b.depth});
        maxdepth = std::max(maxdepth, b.depth);
    }
    std::vector<std::vector<int32_t>> level((size_t)maxdepth + 1);
    for (int32_t t = 0; t < (int32_t)tasks.size(); ++t) level[(size_t)tasks[(size_t)t].depth].push_back(t);
    struct Work {
        std::vector<double> x;
        std::vector<int32_t> flag, stack;
    };
    std::vector<Work> work((size_t)nthreads);
    for (Work& w : work) {
        w.x.assign((size_t)n, 0.0);
        w.flag.assign((size_t)n, -1);
        w.stack.resize((size_t)n);
    }
    std::atomic<int32_t> bad_pivot{-1};
    auto factor_rows = [&](Work& w, int32_t lo, int32_t hi) {
        double* x = w.x.data();
        int32_t* flag = w.flag.data();
        int32_t* stack = w.stack.data();
        for (int32_t k = lo; k < hi; ++k) {
            // row pattern by row-subtree traversal (as in the counting pass)
            int32_t top = n;
            flag[k] = k;
            for (int64_t p = Cp[k]; p < Cp[k + 1]; ++p) {
                int32_t i = Ci[p];
                if (i > k) continue;
                int32_t len = 0;
                for (; flag[i] != k; i = parent[i]) {
                    stack[len++] = i;
                    flag[i] = k;
                }
                while (len > 0) stack[--top] = stack[--len];
            }
            double d = 0.0;
            for (int64_t p = Cp[k]; p < Cp[k + 1]; ++p) {
                if (Ci[p] == k)
                    d += Cx[p];
                else
                    x[Ci[p]] += Cx[p];
            }
            const double d_orig = d;
            int64_t rp = F.Lp[k];
            for (int32_t q = top; q < n; ++q) {
                int32_t j = stack[q];
                double lkj = x[j] / Lcx[Lcp[j]];
                x[j] = 0.0;
                for (int64_t p = Lcp[j] + 1; p < cnext[j]; ++p) x[Lci[p]] -= Lcx[p] * lkj;
                d -= lkj * lkj;
                int64_t at = cnext[j]++;
                Lci[at] = k;
                Lcx[at] = lkj;
                F.Li[rp] = j;
                F.Lx[rp] = lkj;
                ++rp;
            }
            // a pivot that cancelled to rounding level means a (numerically) singular matrix, e.g. no Dirichlet dofs
            if (!(d > 1.0e-12 * std::fabs(d_orig)) || !std::isfinite(d)) {
                int32_t expect = -1;
                bad_pivot.compare_exchange_strong(expect, k);
                d = 1.0;  // keep going with a harmless value; the factorisation is rejected below
            }
            double lkk = std::sqrt(d);
            int64_t at = cnext[k]++;
            Lci[at] = k;
            Lcx[at] = lkk;
            F.dinv[k] = 1.0 / lkk;
        }
    };
    auto run_parallel = [&](int32_t ntask, const std::function<void(int, int32_t)>& body) {
        const int nt = std::min<int>(nthreads, std::max<int32_t>(ntask, 1));
        if (nt <= 1) {
            for (int32_t t = 0; t < ntask; ++t) body(0, t);
            return;
        }
        std::atomic<int32_t> next{0};
        std::vector<std::thread> pool;
        for (int th = 0; th < nt; ++th)
            pool.emplace_back([&, th]() {
                for (int32_t t = next.fetch_add(1); t < ntask; t = next.fetch_add(1)) body(th, t);
            });
        for (std::thread& th : pool) th.join();
    };
    if (nthreads == 1) {  // elimination order (best locality)
        factor_rows(work[0], 0, n);
        maxdepth = -1;
    }
    // Near the root a level has fewer nodes than threads and a node is one long separator S = [lo, hi).  Its rows are
    // then swept in three phases: (A) every row against the columns of the descendants (j < lo; these columns hold
    // only descendant rows at that time, so the rows of S are independent), (B) the Schur contributions of those
    // columns to the S x S block (needs the L[S, j] of phase A, read-only), both in parallel over the rows, and (C) the
    // up-looking sweep restricted to the columns of S, in order.
    std::vector<int64_t> snap;
    auto factor_separator = [&](int32_t lo, int32_t hi) {
        const int32_t ns = hi - lo;
        std::vector<double> Cd((size_t)ns * (size_t)ns, 0.0);  // row k - lo: entries of the columns lo..k (diagonal last)
        std::vector<int64_t> rowD((size_t)ns, 0);
        std::vector<std::vector<int32_t>> sreach((size_t)ns);
        snap.assign(cnext.begin(), cnext.begin() + lo);        // end of the descendant part of the columns j < lo
        run_parallel(ns, [&](int th, int32_t t) {               // ---- phase A
            Work& w = work[(size_t)th];
            double* x = w.x.data();
            int32_t* flag = w.flag.data();
            int32_t* stack = w.stack.data();
            const int32_t k = lo + t;
            int32_t top = n;
            flag[k] = k;
            for (int64_t p = Cp[k]; p < Cp[k + 1]; ++p) {
                int32_t i = Ci[p];
                if (i > k) continue;
                int32_t len = 0;
                for (; flag[i] != k; i = parent[i]) {
                    stack[len++] = i;
                    flag[i] = k;
                }
                while (len > 0) stack[--top] = stack[--len];
            }
            double d = 0.0;
            for (int64_t p = Cp[k]; p < Cp[k + 1]; ++p) {
                if (Ci[p] == k)
                    d += Cx[p];
                else
                    x[Ci[p]] += Cx[p];
            }
            Cd[(size_t)t * ns + t] = d;  // A[k,k]; the products are subtracted in phase C order below
            double dsub = 0.0;
            int64_t rp = F.Lp[k];
            for (int32_t q = top; q < n; ++q) {
                const int32_t j = stack[q];
                if (j >= lo) {
                    sreach[(size_t)t].push_back(j);
                    continue;
                }
                const double lkj = x[j] / Lcx[Lcp[j]];
                x[j] = 0.0;
                for (int64_t p = Lcp[j] + 1; p < snap[(size_t)j]; ++p) x[Lci[p]] -= Lcx[p] * lkj;
                dsub += lkj * lkj;
                F.Li[rp] = j;
                F.Lx[rp] = lkj;
                ++rp;
            }
            rowD[(size_t)t] = rp - F.Lp[k];
            for (int32_t j : sreach[(size_t)t]) {
                Cd[(size_t)t * ns + (j - lo)] = x[j];
                x[j] = 0.0;
            }
            Cd[(size_t)t * ns + t] -= dsub;
        });
        for (int32_t t = 0; t < ns; ++t) {                      // L[S, j] into the columns, in row order
            const int32_t k = lo + t;
            for (int64_t rp = F.Lp[k]; rp < F.Lp[k] + rowD[(size_t)t]; ++rp) {
                const int64_t at = cnext[F.Li[rp]]++;
                Lci[at] = k;
                Lcx[at] = F.Lx[rp];
            }
        }
        run_parallel(ns, [&](int, int32_t t) {                  // ---- phase B
            const int32_t k = lo + t;
            double* c = &Cd[(size_t)t * ns];
            for (int64_t rp = F.Lp[k]; rp < F.Lp[k] + rowD[(size_t)t]; ++rp) {
                const int32_t j = F.Li[rp];
                const double lkj = F.Lx[rp];
                for (int64_t p = snap[(size_t)j]; p < cnext[j] && Lci[p] < k; ++p) c[Lci[p] - lo] -= Lcx[p] * lkj;
            }
        });
        double* x = work[0].x.data();
        for (int32_t t = 0; t < ns; ++t) {                      // ---- phase C
            const int32_t k = lo + t;
            for (int32_t j : sreach[(size_t)t]) x[j] = Cd[(size_t)t * ns + (j - lo)];
            double d = Cd[(size_t)t * ns + t];
            const double d_orig = std::fabs(d) + 1.0;
            int64_t rp = F.Lp[k] + rowD[(size_t)t];
            for (int32_t j : sreach[(size_t)t]) {
                const double lkj = x[j] / Lcx[Lcp[j]];
                x[j] = 0.0;
                for (int64_t p = Lcp[j] + 1; p < cnext[j]; ++p) x[Lci[p]] -= Lcx[p] * lkj;
                d -= lkj * lkj;
                const int64_t at = cnext[j]++;
                Lci[at] = k;
                Lcx[at] = lkj;
                F.Li[rp] = j;
                F.Lx[rp] = lkj;
                ++rp;
            }
            if (!(d > 1.0e-12 * (d_orig - 1.0)) || !std::isfinite(d)) {
                int32_t expect = -1;
                bad_pivot.compare_exchange_strong(expect, k);
                d = 1.0;
            }
            const double lkk = std::sqrt(d);
            const int64_t at = cnext[k]++;
            Lci[at] = k;
            Lcx[at] = lkk;
            F.dinv[k] = 1.0 / lkk;
        }
    };
    for (int32_t dpt = maxdepth; dpt >= 0; --dpt) {
        const std::vector<int32_t>& lv = level[(size_t)dpt];
        if (nthreads >= 4 && 2 * (int)lv.size() <= nthreads && getenv("ASGFEM_CHOL_NOSPLIT") == nullptr) {
            for (int32_t t : lv) {
                const Task& tk = tasks[(size_t)t];
                if (tk.hi - tk.lo >= 96 && tk.hi - tk.lo <= 8192)
                    factor_separator(tk.lo, tk.hi);
                else
                    factor_rows(work[0], tk.lo, tk.hi);
            }
            continue;
        }
        run_parallel((int32_t)lv.size(), [&](int th, int32_t t) {
            const Task& tk = tasks[(size_t)lv[(size_t)t]];
            factor_rows(work[(size_t)th], tk.lo, tk.hi);
        });
    }
    if (bad_pivot.load() >= 0) {
        err = "K_0 restricted to the interior dofs is not positive definite (pivot " + std::to_string(bad_pivot.load()) + ")";
        return ASGFEM_ENUMERIC;
    }
    // rows of L in the output are sorted by column for coalesced/monotone access (ereach order is topological, not sorted)
    {
        const int32_t nchunk = (int32_t)std::min<int64_t>(n, 4 * (int64_t)nthreads);
        run_parallel(nchunk, [&](int, int32_t c) {
            const int32_t k0 = (int32_t)((int64_t)n * c / nchunk), k1 = (int32_t)((int64_t)n * (c + 1) / nchunk);
            std::vector<std::pair<int32_t, double>> tmp;
            for (int32_t k = k0; k < k1; ++k) {
                const int64_t a = F.Lp[k], b = F.Lp[k + 1];
                tmp.clear();
                for (int64_t p = a; p < b; ++p) tmp.push_back({F.Li[p], F.Lx[p]});
                std::sort(tmp.begin(), tmp.end(),
                          [](const std::pair<int32_t, double>& u, const std::pair<int32_t, double>& v) { return u.first < v.first; });
                for (int64_t p = a; p < b; ++p) {
                    F.Li[p] = tmp[(size_t)(p - a)].first;
                    F.Lx[p] = tmp[(size_t)(p - a)].second;
                }
            }
        });
    }
    F.perm.resize((size_t)n);
    for (int32_t k = 0; k < n; ++k) F.perm[k] = full[perm[k]];
    chol_tick("numeric + row sort");
    return 0;
}

}  // namespace asgfem
