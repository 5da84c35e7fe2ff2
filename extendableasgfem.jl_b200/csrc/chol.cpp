// Host-side sparse Cholesky of the Dirichlet-reduced mean stiffness matrix K_0.
//
// Replaces `lu(A0.cscmatrix)` of MyPreconditionerPrimal (src/modelproblems/solvers_poisson_primal.jl:30-44):
// the reference pins boundary dofs with a 1e60 diagonal and hands the matrix to UMFPACK; here the boundary
// rows/columns are removed (the result on those rows is defined as exactly 0) and the remaining SPD matrix is
// factorised as P K P^T = L L^T once per refinement level:
//   1. fill-reducing ordering: graph nested dissection with BFS level-set separators (no METIS in the image),
//   2. elimination tree + column counts by row-subtree traversal,
//   3. up-looking numeric factorisation.
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

// nested dissection ordering; returns perm (elimination order -> node)
void nested_dissection(const Graph& g, std::vector<int32_t>& perm, const double* xy, std::vector<BlockRec>& blocks,
                       int32_t max_block) {
    const int32_t n = g.n;
    // dissection stops at subdomains of <= LEAF unknowns (dense leaf blocks of the triangular sweeps); measured at config 4
    // (sptrsv.cu, DMMA sweeps): see DESIGN.md section 5
    int32_t LEAF = 24;
    if (const char* e = std::getenv("ASGFEM_CHOL_LEAF")) LEAF = std::max(4, atoi(e));
    perm.assign((size_t)n, -1);
    std::vector<int32_t> tag((size_t)n, 0), dist((size_t)n, -1), bfsout, lv, tmp;
    struct Task {
        std::vector<int32_t> nodes;
        int32_t lo;     // this set occupies perm[lo, lo + nodes.size())
        int32_t depth;  // depth in the dissection tree
    };
    // blocks for the tree-parallel triangular solves: every leaf and every separator of the dissection tree is a
    // block (separators longer than max_block are cut into sequentially dependent chunks)
    blocks.clear();
    auto emit_range = [&](int32_t lo, int32_t len, int32_t depth) {
        if (len <= 0) return;
        int32_t nch = (len + max_block - 1) / max_block;
        for (int32_t c = 0; c < nch; ++c)
            blocks.push_back({lo + c * max_block, std::min(max_block, len - c * max_block), depth, c, nch});
    };
    std::vector<Task> stack;
    {
        Task t;
        t.nodes.resize((size_t)n);
        std::iota(t.nodes.begin(), t.nodes.end(), 0);
        t.lo = 0;
        t.depth = 0;
        stack.push_back(std::move(t));
    }
    int32_t next_id = 1;
    while (!stack.empty()) {
        Task t = std::move(stack.back());
        stack.pop_back();
        const int32_t sz = (int32_t)t.nodes.size();
        if (sz == 0) continue;
        if (sz <= LEAF) {
            for (int32_t k = 0; k < sz; ++k) perm[t.lo + k] = t.nodes[k];
            emit_range(t.lo, sz, t.depth);
            continue;
        }
        const int32_t id = next_id++;
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
                        continue;
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
            continue;
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
            continue;
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
    }
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
    if (n == 0) return 0;

    Graph g;
    g.n = n;
    g.ptr.assign((size_t)n + 1, 0);
    for (int32_t r = 0; r < n; ++r) {
        int64_t i = full[r];
        for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p)
            if (red[col[p]] >= 0 && col[p] != i) g.ptr[r + 1]++;
    }
    for (int32_t r = 0; r < n; ++r) g.ptr[r + 1] += g.ptr[r];
    g.adj.resize((size_t)g.ptr[n]);
    for (int32_t r = 0; r < n; ++r) {
        int64_t i = full[r], q = g.ptr[r];
        for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p)
            if (red[col[p]] >= 0 && col[p] != i) g.adj[q++] = red[col[p]];
    }

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
    nested_dissection(g, perm, coords_full ? xy.data() : nullptr, F.blocks, max_block);
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
    for (int32_t k = 0; k < n; ++k) {
        int64_t i = full[perm[k]];
        for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p) {
            int32_t rj = red[col[p]];
            if (rj >= 0 && iperm[rj] <= k) Cp[k + 1]++;
        }
    }
    for (int32_t k = 0; k < n; ++k) Cp[k + 1] += Cp[k];
    std::vector<int32_t> Ci((size_t)Cp[n]);
    std::vector<double> Cx((size_t)Cp[n]);
    for (int32_t k = 0; k < n; ++k) {
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
    // ---- row patterns by row-subtree traversal (ereach); pass 1 counts, pass 2 factorises ------------
    std::vector<int32_t> flag((size_t)n, -1), stack((size_t)n), rowlen((size_t)n, 0);
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
    for (int32_t k = 0; k < n; ++k) {
        int32_t top;
        ereach(k, top);
        rowlen[k] = n - top;
        lnz += n - top;
        for (int32_t q = top; q < n; ++q) colcount[stack[q]]++;
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
    int nthreads = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u);
    if (const char* e = std::getenv("ASGFEM_CHOL_THREADS")) nthreads = std::max(1, atoi(e));
    if (n < 20000) nthreads = 1;
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
