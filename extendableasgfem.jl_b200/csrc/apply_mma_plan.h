// Mode-side plan of the block-product operator kernels (apply_mma.cu builds it at asgfem_set_multiindices; apply_mma.cu
// and apply_blk.cu derive their kernel tables from it).
#pragma once
#include <array>
#include <vector>

#include "common.h"

namespace asgfem {

constexpr int SMEM_LIMIT = 227 * 1024;

struct PairSteps {
    int blockE = -1, blockO = -1;
    int nsteps = 0;
};

struct MmaPlan {
    bool layout_ok = false;  // mode-side clustering done (set_multiindices)
    void* blk = nullptr;     // BlkPlan (apply_blk.cu): kernel tables of the block kernel with list exchange
    int64_t N = 0;
    int M = 0;
    int ncols = 0;
    // mode side
    struct Item {
        int dir, consumer;
        double w;
        bool primary;  // the first consumer of (producer, direction)
    };
    std::vector<std::vector<Item>> prod;  // per producer mode
    std::vector<std::array<int, 8>> blocks;              // modes of a home block (-1 = empty)
    std::vector<std::vector<int>> block_dsets;           // D-set ids per block
    std::vector<std::array<int, 8>> dsets;               // directions (-1 = null row)
    std::vector<PairSteps> pairs;
    double dmma_per_row = 0;
};

static inline MmaPlan* mp_of(asgfem_ctx* ctx) { return reinterpret_cast<MmaPlan*>(ctx->mmaplan); }

}  // namespace asgfem
