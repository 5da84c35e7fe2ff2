// CPU-side check of the operator's host planner (no GPU needed): prints the plan statistics for a graded-lex set.
// Build: nvcc -std=c++17 -O2 -I../extendableasgfem.jl_b200/csrc plan_test.cu ../extendableasgfem.jl_b200/csrc/{apply_mma,index}.o -o plan_test.bin
#include <cstdlib>
#include <functional>
#include "common.h"
using namespace asgfem;
int main(int argc, char** argv) {
    int M = argc > 1 ? atoi(argv[1]) : 20, N = argc > 2 ? atoi(argv[2]) : 2000, rowlen = argc > 3 ? atoi(argv[3]) : 7;
    std::vector<std::vector<int64_t>> out;
    std::function<void(int, int, std::vector<int64_t>&)> rec = [&](int m, int d, std::vector<int64_t>& cur) {
        if ((int)out.size() >= N) return;
        if (m == 1) { cur.push_back(d); out.push_back(cur); cur.pop_back(); return; }
        for (int first = d; first >= 0; --first) { cur.push_back(first); rec(m - 1, d - first, cur); cur.pop_back(); if ((int)out.size() >= N) return; }
    };
    for (int d = 0; d <= 8 && (int)out.size() < N; ++d) { std::vector<int64_t> cur; rec(M, d, cur); }
    asgfem_ctx ctx;
    ctx.N = N;
    ctx.mis.N = N; ctx.mis.M = M;
    for (auto& r : out) for (auto v : r) ctx.mis.mi.push_back(v);
    ctx.mis.build_neighbours();
    build_coupling(ctx.mis, 0, ctx.coup);
    setenv("ASGFEM_MMA_VERBOSE", "1", 1);
    int rc = apply_mma_layout(&ctx);
    printf("layout rc=%d ok=%d ld=%zu\n", rc, (int)apply_mma_layout_ok(&ctx), ctx.h_inv.size());
    ctx.n = 64; ctx.M = M; ctx.ld = ctx.h_inv.size();
    ctx.h_rowptr.resize(ctx.n + 1);
    for (int i = 0; i <= ctx.n; ++i) ctx.h_rowptr[i] = (int64_t)i * rowlen;
    rc = apply_mma_build(&ctx);
    printf("build rc=%d (%s)\n", rc, ctx.err.c_str());
    return 0;
}
