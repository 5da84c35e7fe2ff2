// Row-sharded multi-GPU operation inside the library: NCCL communicator per context, halo exchange
// of X rows overlapped with the operator on the rows that need no halo column, all-reduced inner products.
// The reference has no distributed path (dead `using Distributed`, src/ExtendableASGFEM.jl:3); the seams are those of the
// single-GPU path - mul! (solvers_poisson_primal.jl:86-124) and the Krylov driver (:130-169) - on the rows of one rank.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, or the copy a host process such as PyTorch has already loaded): the
// single-GPU library has no link-time dependency on it.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>

#include "common.h"

namespace asgfem {

namespace {
// the few NCCL entry points used, with the ABI of nccl.h 2.x (ncclComm_t is an opaque pointer, ncclUniqueId 128 bytes)
typedef void* nccl_comm_t;
struct nccl_uid {
    char internal[128];
};
enum { NCCL_FLOAT64 = 8, NCCL_SUM = 0 };
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(nccl_uid*) = nullptr;
    int (*CommInitRank)(nccl_comm_t*, int, nccl_uid, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
NcclApi g_nccl;

bool load_nccl(std::string& err) {
    if (g_nccl.ok) return true;
    void* h = nullptr;
    // a copy already loaded by the host process (PyTorch bundles its own) is found through the global scope first
    if (dlsym(RTLD_DEFAULT, "ncclCommInitRank")) h = RTLD_DEFAULT;
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        err = std::string("cannot load NCCL: ") + dlerror();
        return false;
    }
    g_nccl.handle = h;
#define BIND(field, sym)                                                     \
    do {                                                                     \
        *(void**)(&g_nccl.field) = dlsym(h, sym);                            \
        if (!g_nccl.field) {                                                 \
            err = std::string("NCCL symbol missing: ") + sym;                \
            return false;                                                    \
        }                                                                    \
    } while (0)
    BIND(GetUniqueId, "ncclGetUniqueId");
    BIND(CommInitRank, "ncclCommInitRank");
    BIND(CommDestroy, "ncclCommDestroy");
    BIND(GroupStart, "ncclGroupStart");
    BIND(GroupEnd, "ncclGroupEnd");
    BIND(Send, "ncclSend");
    BIND(Recv, "ncclRecv");
    BIND(AllReduce, "ncclAllReduce");
    BIND(Broadcast, "ncclBroadcast");
    BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
    g_nccl.ok = true;
    return true;
}
}  // namespace

struct DistPlan {
    nccl_comm_t comm = nullptr;
    int nranks = 1, rank = 0;
    cudaStream_t cs = nullptr;  // communication stream
    cudaEvent_t packed = nullptr, received = nullptr;
    // halo plan (local numbering, 0-based on the device): per neighbour a send and a receive row list
    std::vector<int> nb_rank;
    std::vector<int64_t> send_ptr, recv_ptr;
    int64_t* d_send_rows = nullptr;  // 1-based (k_pack_rows convention)
    int64_t* d_recv_rows = nullptr;
    double* d_sendbuf = nullptr;
    double* d_recvbuf = nullptr;
    int64_t buf_N = 0;               // N the buffers were sized for
    int64_t interior0 = 0, interior1 = 0;  // owned rows [interior0, interior1) reference no halo column
    double* d_scalar = nullptr;      // all-reduce scratch (2 doubles)
    bool halo_set = false;
    // global mean preconditioner (exact K_0^-1 for every mode): the ranks swap from row shards to MODE shards, every rank
    // applies the factor of the global K_0 to its columns, and swap back
    PrecondPlan* gprec = nullptr;
    int64_t n_global = 0;
    std::vector<int64_t> row_off, col_off;  // nranks+1: global row offsets (rank-major order) / device column chunks (x16)
    uint8_t* d_gbmask = nullptr;
    double* d_T = nullptr;       // n_global x (columns of this rank)
    double* d_stage = nullptr;   // n_owned x ld, blocks [n_owned x columns of rank q] one after the other
    int64_t T_cols = 0, stage_ld = 0;
};

static DistPlan* dp_of(asgfem_ctx* ctx) { return reinterpret_cast<DistPlan*>(ctx->distplan); }

void dist_free(asgfem_ctx* ctx) {
    DistPlan* D = dp_of(ctx);
    if (!D) return;
    if (D->comm && g_nccl.ok) g_nccl.CommDestroy(D->comm);
    precond_free_plan(D->gprec);
    void* ptrs[] = {D->d_send_rows, D->d_recv_rows, D->d_sendbuf, D->d_recvbuf, D->d_scalar, D->d_gbmask, D->d_T, D->d_stage};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    if (D->packed) cudaEventDestroy(D->packed);
    if (D->received) cudaEventDestroy(D->received);
    if (D->cs) cudaStreamDestroy(D->cs);
    delete D;
    ctx->distplan = nullptr;
}

bool dist_active(asgfem_ctx* ctx) {
    DistPlan* D = dp_of(ctx);
    return D && D->comm && D->nranks > 1;
}

#define NCCL_CHECK(ctx, call)                                                                                         \
    do {                                                                                                              \
        int _r = (call);                                                                                              \
        if (_r != 0) return fail((ctx), ASGFEM_ECUDA, std::string(#call) + ": " + g_nccl.GetErrorString(_r));         \
    } while (0)

int dist_unique_id(void* id128, std::string& err) {
    if (!load_nccl(err)) return ASGFEM_ESTATE;
    nccl_uid id;
    int r = g_nccl.GetUniqueId(&id);
    if (r != 0) {
        err = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r);
        return ASGFEM_ECUDA;
    }
    std::memcpy(id128, &id, 128);
    return 0;
}

int dist_init(asgfem_ctx* ctx, int nranks, int rank, const void* id128) {
    std::string err;
    if (!load_nccl(err)) return fail(ctx, ASGFEM_ESTATE, err);
    dist_free(ctx);
    DistPlan* D = new DistPlan();
    ctx->distplan = D;
    D->nranks = nranks;
    D->rank = rank;
    nccl_uid id;
    std::memcpy(&id, id128, 128);
    NCCL_CHECK(ctx, g_nccl.CommInitRank(&D->comm, nranks, id, rank));
    ASG_CUDA(ctx, cudaStreamCreateWithFlags(&D->cs, cudaStreamNonBlocking));
    ASG_CUDA(ctx, cudaEventCreateWithFlags(&D->packed, cudaEventDisableTiming));
    ASG_CUDA(ctx, cudaEventCreateWithFlags(&D->received, cudaEventDisableTiming));
    ASG_CUDA(ctx, cudaMalloc((void**)&D->d_scalar, 4 * sizeof(double)));
    return 0;
}

int dist_set_halo(asgfem_ctx* ctx, int32_t nneigh, const int32_t* ranks, const int64_t* send_ptr, const int64_t* send_rows,
                  const int64_t* recv_ptr, const int64_t* recv_rows, int64_t interior0, int64_t interior1) {
    DistPlan* D = dp_of(ctx);
    ASG_CHECK(ctx, D && D->comm, ASGFEM_ESTATE, "set_halo: asgfem_comm_init first");
    ASG_CHECK(ctx, ctx->n > 0 && ctx->n_owned > 0, ASGFEM_ESTATE, "set_halo: pattern and owned rows first");
    ASG_CHECK(ctx, nneigh >= 0 && (nneigh == 0 || (ranks && send_ptr && recv_ptr)), ASGFEM_EINVAL, "set_halo: bad arguments");
    ASG_CHECK(ctx, 0 <= interior0 && interior0 <= interior1 && interior1 <= ctx->n_owned, ASGFEM_EINVAL, "set_halo: bad interior range");
    D->nb_rank.assign(ranks, ranks + nneigh);
    D->send_ptr.assign(send_ptr, send_ptr + nneigh + 1);
    D->recv_ptr.assign(recv_ptr, recv_ptr + nneigh + 1);
    for (int k = 0; k < nneigh; ++k)
        ASG_CHECK(ctx, ranks[k] >= 0 && ranks[k] < D->nranks && ranks[k] != D->rank, ASGFEM_EINVAL, "set_halo: bad neighbour rank");
    const int64_t ns = D->send_ptr[nneigh], nr = D->recv_ptr[nneigh];
    for (int64_t k = 0; k < ns; ++k) ASG_CHECK(ctx, send_rows[k] >= 1 && send_rows[k] <= ctx->n_owned, ASGFEM_EINVAL, "set_halo: send row not owned");
    for (int64_t k = 0; k < nr; ++k)
        ASG_CHECK(ctx, recv_rows[k] > ctx->n_owned && recv_rows[k] <= ctx->n, ASGFEM_EINVAL, "set_halo: receive row is not a halo row");
    for (void** p : {(void**)&D->d_send_rows, (void**)&D->d_recv_rows, (void**)&D->d_sendbuf, (void**)&D->d_recvbuf})
        if (*p) {
            cudaFree(*p);
            *p = nullptr;
        }
    ASG_CUDA(ctx, cudaMalloc((void**)&D->d_send_rows, sizeof(int64_t) * std::max<int64_t>(ns, 1)));
    ASG_CUDA(ctx, cudaMalloc((void**)&D->d_recv_rows, sizeof(int64_t) * std::max<int64_t>(nr, 1)));
    ASG_CUDA(ctx, cudaMemcpyAsync(D->d_send_rows, send_rows, sizeof(int64_t) * ns, cudaMemcpyHostToDevice, ctx->stream));
    ASG_CUDA(ctx, cudaMemcpyAsync(D->d_recv_rows, recv_rows, sizeof(int64_t) * nr, cudaMemcpyHostToDevice, ctx->stream));
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    D->buf_N = 0;
    D->interior0 = interior0;
    D->interior1 = interior1;
    D->halo_set = true;
    return 0;
}

static int ensure_buffers(asgfem_ctx* ctx, DistPlan* D) {
    if (D->buf_N == ctx->N && D->d_sendbuf) return 0;
    for (void** p : {(void**)&D->d_sendbuf, (void**)&D->d_recvbuf})
        if (*p) {
            cudaFree(*p);
            *p = nullptr;
        }
    const int64_t ns = D->send_ptr.empty() ? 0 : D->send_ptr.back(), nr = D->recv_ptr.empty() ? 0 : D->recv_ptr.back();
    ASG_CUDA(ctx, cudaMalloc((void**)&D->d_sendbuf, sizeof(double) * std::max<int64_t>(ns * ctx->N, 1)));
    ASG_CUDA(ctx, cudaMalloc((void**)&D->d_recvbuf, sizeof(double) * std::max<int64_t>(nr * ctx->N, 1)));
    D->buf_N = ctx->N;
    return 0;
}

// Y = A X on the owned rows of this rank.  One launch sequence, no host synchronisation:
//   stream:       pack send rows | operator on [interior0, interior1) | wait | unpack halo rows | operator on the other owned rows
//   comm stream:                 | grouped ncclSend / ncclRecv         |
// halo rows of x <- owned rows of the neighbours (pack, ncclSend / ncclRecv, unpack on the context's stream)
int dist_halo_exchange(asgfem_ctx* ctx, double* x) {
    DistPlan* D = dp_of(ctx);
    if (!dist_active(ctx)) return 0;
    ASG_CHECK(ctx, D->halo_set, ASGFEM_ESTATE, "halo exchange: asgfem_set_halo first");
    int rc = ensure_buffers(ctx, D);
    if (rc) return rc;
    const int nn = (int)D->nb_rank.size();
    const int64_t N = ctx->N, ns = D->send_ptr[nn], nr = D->recv_ptr[nn];
    if ((rc = vec_pack_rows(ctx, x, ns, D->d_send_rows, D->d_sendbuf))) return rc;
    NCCL_CHECK(ctx, g_nccl.GroupStart());
    for (int k = 0; k < nn; ++k) {
        const int64_t s0 = D->send_ptr[k], s1 = D->send_ptr[k + 1], r0 = D->recv_ptr[k], r1 = D->recv_ptr[k + 1];
        if (s1 > s0) NCCL_CHECK(ctx, g_nccl.Send(D->d_sendbuf + s0 * N, (size_t)((s1 - s0) * N), NCCL_FLOAT64, D->nb_rank[k], D->comm, ctx->stream));
        if (r1 > r0) NCCL_CHECK(ctx, g_nccl.Recv(D->d_recvbuf + r0 * N, (size_t)((r1 - r0) * N), NCCL_FLOAT64, D->nb_rank[k], D->comm, ctx->stream));
    }
    NCCL_CHECK(ctx, g_nccl.GroupEnd());
    return vec_unpack_rows(ctx, x, nr, D->d_recv_rows, D->d_recvbuf);
}

// Exchange first, then ONE operator launch over all owned rows.  The halo rows are 2 x 16 MB per neighbour (about 0.1 ms
// over NVLink); overlapping them with the interior rows (round 1 / first version of this function: NCCL on a second
// stream, three operator launches) cost more than it hid: the persistent operator CTAs and the NCCL CTAs compete for the
// SMs, a rank whose NCCL kernel starts late stalls its neighbours' kernels, and every extra launch pays the 222 KB table
// prologue - 73.8 ms per step at 4 GPUs against 51.5 ms for the launch alone.
int dist_apply(asgfem_ctx* ctx, const double* x, double* y) {
    if (!dist_active(ctx)) return apply_launch(ctx, x, y);
    int rc = dist_halo_exchange(ctx, const_cast<double*>(x));
    if (rc) return rc;
    return apply_launch(ctx, x, y, 0, ctx->n_owned);
}

// sum over the ranks of the inner product on the owned rows (deterministic local part + ncclAllReduce)
int dist_dot(asgfem_ctx* ctx, const double* a, const double* b, double* out) {
    DistPlan* D = dp_of(ctx);
    const int64_t rows = ctx->n_owned >= 0 ? ctx->n_owned : ctx->n;
    if (!dist_active(ctx)) return vec_dot(ctx, a, b, rows, out);
    double local = 0;
    int rc = vec_dot(ctx, a, b, rows, &local);  // leaves the value in d_partial as well; the host copy is what we reduce
    if (rc) return rc;
    ASG_CUDA(ctx, cudaMemcpyAsync(D->d_scalar, &local, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    NCCL_CHECK(ctx, g_nccl.AllReduce(D->d_scalar, D->d_scalar + 1, 1, NCCL_FLOAT64, NCCL_SUM, D->comm, ctx->stream));
    ASG_CUDA(ctx, cudaMemcpyAsync(out, D->d_scalar + 1, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---- global mean preconditioner ---------------------------------------------------------------------------------------
namespace {
struct ColChunks {
    int n;
    int64_t off[33];
};
// stage block q = rows x (columns of rank q), row-major, blocks concatenated; pack: v -> stage, unpack: stage -> v
__global__ void k_swap_cols(double* __restrict__ v, double* __restrict__ stage, int64_t nrows, int64_t ld, ColChunks cc, int pack) {
    const int64_t total = nrows * (ld / 2);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / (ld / 2), c = 2 * (t - i * (ld / 2));
        int q = 0;
        while (q + 1 < cc.n && c >= cc.off[q + 1]) ++q;
        const int64_t w = cc.off[q + 1] - cc.off[q];
        double2* sp = reinterpret_cast<double2*>(stage + nrows * cc.off[q] + i * w + (c - cc.off[q]));
        double2* vp = reinterpret_cast<double2*>(v + i * ld + c);
        if (pack)
            *sp = *vp;
        else
            *vp = *sp;
    }
}
}  // namespace

int dist_precond_setup_global(asgfem_ctx* ctx, int64_t n_global, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                              int64_t nb, const int64_t* bdofs, const double* coords, const int64_t* row_offsets) {
    DistPlan* D = dp_of(ctx);
    ASG_CHECK(ctx, D && D->comm, ASGFEM_ESTATE, "precond_setup_global: asgfem_comm_init first");
    ASG_CHECK(ctx, ctx->ld > 0 && ctx->n_owned > 0, ASGFEM_ESTATE, "precond_setup_global: multi-indices and owned rows first");
    ASG_CHECK(ctx, n_global > 0 && row_offsets && (D->rank != 0 || (colptr && rowval && nzval)), ASGFEM_EINVAL,
              "precond_setup_global: bad arguments (the matrix is needed on rank 0 only)");
    ASG_CHECK(ctx, D->nranks <= 32, ASGFEM_EINVAL, "precond_setup_global: at most 32 ranks");
    D->row_off.assign(row_offsets, row_offsets + D->nranks + 1);
    ASG_CHECK(ctx, D->row_off[0] == 0 && D->row_off[D->nranks] == n_global &&
                       D->row_off[D->rank + 1] - D->row_off[D->rank] == ctx->n_owned,
              ASGFEM_EINVAL, "precond_setup_global: row offsets do not match the owned rows of this rank");
    std::vector<uint8_t> bm((size_t)n_global, 0);
    for (int64_t k = 0; k < nb; ++k) {
        ASG_CHECK(ctx, bdofs[k] >= 1 && bdofs[k] <= n_global, ASGFEM_EINVAL, "precond_setup_global: boundary dof out of range");
        bm[(size_t)(bdofs[k] - 1)] = 1;
    }
    precond_free_plan(D->gprec);
    D->gprec = nullptr;
    // Rank 0 factorises (all host cores to itself; with every rank factorising the same matrix the host Cholesky of the 4 M
    // dof mean matrix took 47 s at 4 GPUs) and hands the sweep tasks to the others over NCCL.
    int rc = 0;
    int64_t hdr[6] = {0, 0, 0, 0, 0, 0};  // sizes + status of the root
    if (D->rank == 0) {
        // CSC (1-based) -> CSR (0-based); K_0 is symmetric, the transposition keeps the routine general
        const int64_t nnz = colptr[n_global] - 1;
        std::vector<int64_t> rp((size_t)n_global + 1, 0);
        bool ok = true;
        for (int64_t p = 0; p < nnz; ++p) {
            if (rowval[p] < 1 || rowval[p] > n_global) {
                ok = false;
                break;
            }
            rp[(size_t)rowval[p]]++;
        }
        if (!ok) {
            rc = fail(ctx, ASGFEM_EINVAL, "precond_setup_global: row index out of range");
        } else {
            for (int64_t i = 0; i < n_global; ++i) rp[(size_t)i + 1] += rp[(size_t)i];
            std::vector<int32_t> ci((size_t)nnz);
            std::vector<double> cv((size_t)nnz);
            std::vector<int64_t> fill(rp.begin(), rp.end() - 1);
            for (int64_t c = 0; c < n_global; ++c)
                for (int64_t p = colptr[c] - 1; p < colptr[c + 1] - 1; ++p) {
                    const int64_t at = fill[(size_t)(rowval[p] - 1)]++;
                    ci[(size_t)at] = (int32_t)c;
                    cv[(size_t)at] = nzval[p];
                }
            rc = precond_build(ctx, n_global, rp.data(), ci.data(), cv.data(), bm.data(), coords, &D->gprec);
        }
        if (!rc) precond_plan_sizes(D->gprec, hdr);
        hdr[5] = rc;
    }
    {
        // header and launch list through a small device buffer
        int64_t* d_hdr = nullptr;
        ASG_CUDA(ctx, cudaMalloc((void**)&d_hdr, sizeof(hdr)));
        ASG_CUDA(ctx, cudaMemcpyAsync(d_hdr, hdr, sizeof(hdr), cudaMemcpyHostToDevice, ctx->stream));
        NCCL_CHECK(ctx, g_nccl.Broadcast(d_hdr, d_hdr, sizeof(hdr), 0 /* ncclInt8 */, 0, D->comm, ctx->stream));
        ASG_CUDA(ctx, cudaMemcpyAsync(hdr, d_hdr, sizeof(hdr), cudaMemcpyDeviceToHost, ctx->stream));
        ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(d_hdr);
        if (hdr[5] != 0) return D->rank == 0 ? rc : fail(ctx, (int)hdr[5], "precond_setup_global: the factorisation on rank 0 failed");
        if (D->rank != 0 && (rc = precond_plan_alloc(ctx, hdr, &D->gprec))) return rc;
        void* ptrs[3];
        size_t bytes[3];
        precond_plan_buffers(D->gprec, ptrs, bytes);
        for (int k = 0; k < 3; ++k)
            if (bytes[k] > 0) NCCL_CHECK(ctx, g_nccl.Broadcast(ptrs[k], ptrs[k], bytes[k], 0 /* ncclInt8 */, 0, D->comm, ctx->stream));
        const size_t lbytes = sizeof(int) * 3 * (size_t)hdr[4];
        if (lbytes > 0) {
            int* d_l = nullptr;
            ASG_CUDA(ctx, cudaMalloc((void**)&d_l, lbytes));
            if (D->rank == 0) ASG_CUDA(ctx, cudaMemcpyAsync(d_l, precond_plan_launches(D->gprec), lbytes, cudaMemcpyHostToDevice, ctx->stream));
            NCCL_CHECK(ctx, g_nccl.Broadcast(d_l, d_l, lbytes, 0 /* ncclInt8 */, 0, D->comm, ctx->stream));
            if (D->rank != 0) ASG_CUDA(ctx, cudaMemcpyAsync(precond_plan_launches(D->gprec), d_l, lbytes, cudaMemcpyDeviceToHost, ctx->stream));
            ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            cudaFree(d_l);
        }
        ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    D->n_global = n_global;
    if ((rc = dev_upload(ctx, &D->d_gbmask, bm))) return rc;
    // column chunks: tiles of 16 device columns dealt out evenly
    const int64_t tiles = ctx->ld / 16;
    D->col_off.assign((size_t)D->nranks + 1, 0);
    for (int q = 0; q <= D->nranks; ++q) D->col_off[(size_t)q] = 16 * (tiles * q / D->nranks);
    for (void** p : {(void**)&D->d_T, (void**)&D->d_stage})
        if (*p) {
            cudaFree(*p);
            *p = nullptr;
        }
    D->T_cols = D->col_off[(size_t)D->rank + 1] - D->col_off[(size_t)D->rank];
    D->stage_ld = ctx->ld;
    ASG_CUDA(ctx, cudaMalloc((void**)&D->d_T, sizeof(double) * (size_t)std::max<int64_t>(n_global * D->T_cols, 1)));
    ASG_CUDA(ctx, cudaMalloc((void**)&D->d_stage, sizeof(double) * (size_t)ctx->n_owned * (size_t)ctx->ld));
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

bool dist_has_global_precond(asgfem_ctx* ctx) {
    DistPlan* D = dp_of(ctx);
    return dist_active(ctx) && D->gprec;
}

// z = (I (x) K_0^-1) r with the GLOBAL K_0: row shards -> mode shards (all-to-all), sweeps, and back
int dist_precond_apply(asgfem_ctx* ctx, const double* r, double* z) {
    DistPlan* D = dp_of(ctx);
    if (!dist_has_global_precond(ctx)) return precond_apply(ctx, r, z);
    ASG_CHECK(ctx, D->stage_ld == ctx->ld, ASGFEM_ESTATE, "precond_apply: multi-index set changed after precond_setup_global");
    const int64_t no = ctx->n_owned, ld = ctx->ld, wp = D->T_cols;
    ColChunks cc;
    cc.n = D->nranks;
    for (int q = 0; q <= D->nranks; ++q) cc.off[q] = D->col_off[(size_t)q];
    const int grid = (int)std::min<int64_t>(148 * 16, (no * (ld / 2) + 255) / 256);
    k_swap_cols<<<grid, 256, 0, ctx->stream>>>(const_cast<double*>(r), D->d_stage, no, ld, cc, 1);
    ASG_CUDA(ctx, cudaGetLastError());
    auto exchange = [&](bool forward) -> int {
        NCCL_CHECK(ctx, g_nccl.GroupStart());
        for (int q = 0; q < D->nranks; ++q) {
            const int64_t wq = D->col_off[(size_t)q + 1] - D->col_off[(size_t)q], nq = D->row_off[(size_t)q + 1] - D->row_off[(size_t)q];
            double* mine = D->d_stage + no * D->col_off[(size_t)q];  // my rows, columns of rank q
            double* theirs = D->d_T + D->row_off[(size_t)q] * wp;    // rows of rank q, my columns
            if (q == D->rank) {
                if (wp > 0)
                    ASG_CUDA(ctx, cudaMemcpyAsync(forward ? theirs : mine, forward ? mine : theirs, sizeof(double) * (size_t)(no * wp),
                                                  cudaMemcpyDeviceToDevice, ctx->stream));
                continue;
            }
            if (forward) {
                if (no * wq > 0) NCCL_CHECK(ctx, g_nccl.Send(mine, (size_t)(no * wq), NCCL_FLOAT64, q, D->comm, ctx->stream));
                if (nq * wp > 0) NCCL_CHECK(ctx, g_nccl.Recv(theirs, (size_t)(nq * wp), NCCL_FLOAT64, q, D->comm, ctx->stream));
            } else {
                if (nq * wp > 0) NCCL_CHECK(ctx, g_nccl.Send(theirs, (size_t)(nq * wp), NCCL_FLOAT64, q, D->comm, ctx->stream));
                if (no * wq > 0) NCCL_CHECK(ctx, g_nccl.Recv(mine, (size_t)(no * wq), NCCL_FLOAT64, q, D->comm, ctx->stream));
            }
        }
        NCCL_CHECK(ctx, g_nccl.GroupEnd());
        return 0;
    };
    int rc = exchange(true);
    if (rc) return rc;
    if (wp > 0 && (rc = precond_apply_plan(ctx, D->gprec, D->d_T, D->d_T, D->n_global, wp, D->d_gbmask))) return rc;
    if ((rc = exchange(false))) return rc;
    k_swap_cols<<<grid, 256, 0, ctx->stream>>>(z, D->d_stage, no, ld, cc, 0);
    ASG_CUDA(ctx, cudaGetLastError());
    // rows beyond the owned ones (halo) carry no part of z
    if (ctx->n > no) ASG_CUDA(ctx, cudaMemsetAsync(z + no * ld, 0, sizeof(double) * (size_t)((ctx->n - no) * ld), ctx->stream));
    return 0;
}

// sum over the ranks of a device array (estimator totals), in place, on the context's stream
int dist_allreduce_sum(asgfem_ctx* ctx, double* dbuf, size_t n) {
    DistPlan* D = dp_of(ctx);
    if (!dist_active(ctx) || n == 0) return 0;
    NCCL_CHECK(ctx, g_nccl.AllReduce(dbuf, dbuf, n, NCCL_FLOAT64, NCCL_SUM, D->comm, ctx->stream));
    return 0;
}

// max over the ranks of a host scalar (device timings)
int dist_max(asgfem_ctx* ctx, double* v) {
    DistPlan* D = dp_of(ctx);
    if (!dist_active(ctx)) return 0;
    ASG_CUDA(ctx, cudaMemcpyAsync(D->d_scalar + 2, v, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    NCCL_CHECK(ctx, g_nccl.AllReduce(D->d_scalar + 2, D->d_scalar + 3, 1, NCCL_FLOAT64, 2 /* ncclMax */, D->comm, ctx->stream));
    ASG_CUDA(ctx, cudaMemcpyAsync(v, D->d_scalar + 3, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    ASG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

}  // namespace asgfem
