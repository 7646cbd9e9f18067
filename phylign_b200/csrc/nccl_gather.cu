// nccl_gather.cu -- multi-GPU exchange of the match stage.
//
// Batches (indexes) are independent, so the gather/count path has no collective at all:
// each GPU holds a shard of the indexes and sees every query.  The only real exchange is
// the one filter_queries.py performs across batch files
// (/root/reference/scripts/filter_queries.py:178-185): the per-GPU top-N + ties lists are
// exchanged over NVLink and merged once more with the same kernel.  Correct because global
// top-N + ties is a subset of the union of per-GPU top-N + ties.
//
// libnccl is resolved with dlopen at phy_nccl_init time, so single-GPU users never need it
// and the process shares whichever libnccl.so.2 is already loaded (e.g. PyTorch's).
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>

#include "phy_internal.cuh"

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommFinalize)(ncclComm_t) = nullptr;
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int load_nccl(phy_ctx* ctx) {
    if (g_nccl.handle) return PHY_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        phy_set_error(ctx, "cannot load libnccl.so.2: %s", dlerror());
        return PHY_ERR_NCCL;
    }
#define LOAD(field, sym)                                         \
    g_nccl.field = (decltype(g_nccl.field))dlsym(h, sym);        \
    if (!g_nccl.field) {                                         \
        phy_set_error(ctx, "libnccl lacks symbol %s", sym);      \
        return PHY_ERR_NCCL;                                     \
    }
    LOAD(GetUniqueId, "ncclGetUniqueId")
    LOAD(CommInitRank, "ncclCommInitRank")
    LOAD(CommDestroy, "ncclCommDestroy")
    LOAD(CommFinalize, "ncclCommFinalize")
    LOAD(CommAbort, "ncclCommAbort")
    LOAD(AllGather, "ncclAllGather")
    LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
    g_nccl.handle = h;
    return PHY_OK;
}

#define PHY_NCCL(ctx, call)                                                                   \
    do {                                                                                      \
        ncclResult_t r_ = (call);                                                             \
        if (r_ != ncclSuccess) {                                                              \
            phy_set_error(ctx, "%s failed: %s", #call, g_nccl.GetErrorString(r_));            \
            return PHY_ERR_NCCL;                                                              \
        }                                                                                     \
    } while (0)

// ---- query-sharded merge: rank s finalises queries [qb[s], qb[s+1]) -----------------------------
// B[r][s] = foffs_all[r][qb[s]]: where rank r's candidates for rank s's queries start
__global__ void shard_bounds_kernel(const uint64_t* __restrict__ foffs_all, uint32_t n_ranks, uint32_t nq,
                                    uint64_t* __restrict__ bounds) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_ranks * (n_ranks + 1)) return;
    const uint32_t r = t / (n_ranks + 1), s = t % (n_ranks + 1);
    const uint32_t qb = (uint32_t)((uint64_t)nq * s / n_ranks);
    bounds[t] = foffs_all[(uint64_t)r * (nq + 1) + qb];
}
__global__ void __launch_bounds__(256) rank_totals_range_kernel(const uint64_t* __restrict__ foffs_all, uint32_t n_ranks,
                                                                uint32_t nq, uint32_t q_lo, uint32_t q_hi,
                                                                uint32_t* __restrict__ totals) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    uint64_t t = 0;
    if (q >= q_lo && q < q_hi)
        for (uint32_t r = 0; r < n_ranks; r++) {
            const uint64_t* f = foffs_all + (uint64_t)r * (nq + 1);
            t += f[q + 1] - f[q];
        }
    totals[q] = (uint32_t)t;
}
// recv holds, rank after rank, every rank's candidates for queries [q_lo, q_hi)
__global__ void __launch_bounds__(128) regroup_range_kernel(const uint64_t* __restrict__ foffs_all,
                                                            const uint64_t* __restrict__ rank_base, uint32_t n_ranks,
                                                            uint32_t nq, uint32_t q_lo, uint32_t q_hi,
                                                            const phy_cand* __restrict__ recv,
                                                            const uint64_t* __restrict__ qoffs_c,
                                                            uint64_t* __restrict__ ckey, uint32_t* __restrict__ cval) {
    for (uint32_t q = q_lo + blockIdx.x; q < q_hi; q += gridDim.x) {
        uint64_t dst = qoffs_c[q];
        for (uint32_t r = 0; r < n_ranks; r++) {
            const uint64_t* f = foffs_all + (uint64_t)r * (nq + 1);
            const uint64_t src = rank_base[r] + (f[q] - f[q_lo]);
            const uint32_t n = (uint32_t)(f[q + 1] - f[q]);
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
                phy_cand c = recv[src + i];
                ckey[dst + i] = ((uint64_t)(~c.score) << 32) | ((uint64_t)c.batch_rank << 20) | c.ref_rank;
                cval[dst + i] = c.doc;
            }
            dst += n;
        }
    }
}

}  // namespace

// queries' bases: every rank uploads 1/R of them over its own PCIe link, the slices meet over NVLink
int phy_nccl_allgather_inplace(phy_ctx* ctx, void* buf, size_t slice_bytes) {
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    PHY_NCCL(ctx, g_nccl.AllGather((const uint8_t*)buf + (size_t)ctx->rank * slice_bytes, buf, slice_bytes, ncclChar,
                                   comm, ctx->stream));
    return PHY_OK;
}

extern "C" int phy_nccl_unique_id(void* id_out) {
    if (!id_out) return PHY_ERR_ARG;
    PHY_TRY(load_nccl(nullptr));
    static_assert(sizeof(ncclUniqueId) == PHY_NCCL_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    PHY_NCCL(nullptr, g_nccl.GetUniqueId(&id));
    memcpy(id_out, &id, sizeof id);
    return PHY_OK;
}

extern "C" int phy_nccl_init(phy_ctx* ctx, const void* id, int rank, int n_ranks) {
    if (!ctx || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return PHY_ERR_ARG;
    PHY_TRY(load_nccl(ctx));
    PHY_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof uid);
    ncclComm_t comm;
    PHY_NCCL(ctx, g_nccl.CommInitRank(&comm, n_ranks, uid, rank));
    ctx->nccl_comm = comm;
    ctx->rank = rank;
    ctx->n_ranks = n_ranks;
    // first collective now: NCCL sets up its peer connections lazily (hundreds of ms), and callers
    // run the init beside their index load -- not inside the first query upload or merge
    void* d = nullptr;
    PHY_TRY(phy_ws_alloc(ctx, &d, 256 * (size_t)n_ranks));
    PHY_NCCL(ctx, g_nccl.AllGather((const uint8_t*)d + 256 * (size_t)rank, d, 256, ncclChar, comm, ctx->stream));
    PHY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    phy_ws_free(ctx, d);
    return PHY_OK;
}

// Collective, orderly end of the communicator: every rank calls it at the same point of the job
// (ncclCommFinalize flushes outstanding work, ncclCommDestroy then only frees local resources).
extern "C" int phy_nccl_finalize(phy_ctx* ctx) {
    if (!ctx) return PHY_ERR_ARG;
    if (!ctx->nccl_comm) return PHY_OK;
    PHY_CUDA(ctx, cudaSetDevice(ctx->device));
    PHY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    ctx->nccl_comm = nullptr;
    ctx->n_ranks = 1;
    ctx->rank = 0;
    PHY_NCCL(ctx, g_nccl.CommFinalize(comm));
    PHY_NCCL(ctx, g_nccl.CommDestroy(comm));
    return PHY_OK;
}

// Context teardown without phy_nccl_finalize (a rank leaving on an error path, ranks ending at
// different times): ncclCommAbort releases the communicator locally and never waits for peers --
// ncclCommDestroy alone blocked for tens of seconds in that situation (measured in round 1).
void phy_nccl_shutdown(phy_ctx* ctx) {
    if (ctx->nccl_comm && g_nccl.CommAbort) g_nccl.CommAbort((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    ctx->n_ranks = 1;
    ctx->rank = 0;
}

// After the local merge: the per-GPU top-N + ties lists meet over NVLink and are merged once more with
// the same kernel.  "merge_mode" 1 (query-sharded): rank s finalises ITS slice of the queries and the
// merged lists stay distributed over the ranks (phy_merged_range tells which queries a rank holds), so
// the merge work and the final download shrink with the number of GPUs.  "merge_mode" 0: rank 0
// finalises every query.
int phy_merge_segments_bounded(phy_ctx* ctx, uint32_t top_n, uint64_t max_total);

int phy_nccl_merge(phy_ctx* ctx, uint32_t top_n) {
    const bool sharded = ctx->merge_sharded;
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    const uint32_t nq = ctx->nq, R = (uint32_t)ctx->n_ranks, me = (uint32_t)ctx->rank;
    auto qb = [&](uint32_t s) { return (uint32_t)((uint64_t)nq * s / R); };
    DevBuf<uint64_t>& all = ctx->d_foffs_all;
    PHY_TRY(phy_ensure(ctx, all, (size_t)R * (nq + 1) + (size_t)R * (R + 1)));
    uint64_t* d_bounds = all.p + (size_t)R * (nq + 1);
    PHY_NCCL(ctx, g_nccl.AllGather(ctx->d_foffs.p, all.p, nq + 1, ncclUint64, comm, ctx->stream));
    shard_bounds_kernel<<<(R * (R + 1) + 127) / 128, 128, 0, ctx->stream>>>(all.p, R, nq, d_bounds);
    ctx->launches++;
    std::vector<uint64_t> B((size_t)R * (R + 1));
    PHY_CUDA(ctx, cudaMemcpyAsync(B.data(), d_bounds, B.size() * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    PHY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    auto bound = [&](uint32_t r, uint32_t s) { return B[(size_t)r * (R + 1) + s]; };
    // Exchange: ONE padded all-gather of every rank's list (a few tens of MB in all) instead of R x R
    // send/recv pairs -- it runs on the collective connections set up at init, whereas the first
    // ncclSend/ncclRecv to each peer costs seconds of lazy point-to-point set-up (measured: 7.7 s for the
    // first match pass of a fresh 8-GPU job).  Each rank then reads its query slice out of every list.
    uint64_t maxn = 1;
    for (uint32_t r = 0; r < R; r++) maxn = std::max<uint64_t>(maxn, bound(r, R));
    std::vector<uint64_t> base(R + 1, 0);
    const uint32_t q_lo = sharded ? qb(me) : 0, q_hi = sharded ? qb(me + 1) : (me == 0 ? nq : 0);
    for (uint32_t r = 0; r < R; r++)       // rank r's first candidate for my slice of the queries
        base[r] = (uint64_t)r * maxn + (sharded ? bound(r, me) : 0);
    PHY_TRY(phy_ensure(ctx, ctx->d_recv, (size_t)R * maxn + 1));
    static_assert(sizeof(phy_cand) == 16, "phy_cand is exchanged as 4 x uint32");
    if (bound(me, R))
        PHY_CUDA(ctx, cudaMemcpyAsync(ctx->d_recv.p + (size_t)me * maxn, ctx->d_final.p, bound(me, R) * sizeof(phy_cand),
                                      cudaMemcpyDeviceToDevice, ctx->stream));
    PHY_NCCL(ctx, g_nccl.AllGather(ctx->d_recv.p + (size_t)me * maxn, ctx->d_recv.p, maxn * 4, ncclUint32, comm, ctx->stream));
    PHY_TRY(phy_ensure(ctx, ctx->d_rank_base, R + 1));
    PHY_TRY(phy_h2d(ctx, ctx->d_rank_base.p, base.data(), (R + 1) * sizeof(uint64_t)));
    PHY_TRY(phy_ensure(ctx, ctx->d_qcount, nq + 1));
    if (nq) {
        rank_totals_range_kernel<<<(nq + 255) / 256, 256, 0, ctx->stream>>>(all.p, R, nq, q_lo, q_hi, ctx->d_qcount.p);
        ctx->launches++;
    }
    uint64_t total = 0;  // bound: every candidate of every rank (this rank's slice holds a part of them)
    for (uint32_t r = 0; r < R; r++) total += bound(r, R);
    PHY_TRY(phy_ensure(ctx, ctx->d_qoffs_c, nq + 2));
    PHY_TRY(phy_exscan(ctx, ctx->d_qcount.p, nq, ctx->d_qoffs_c.p, nullptr));
    PHY_TRY(phy_ensure(ctx, ctx->d_ckey, total + 1));
    PHY_TRY(phy_ensure(ctx, ctx->d_cval, total + 1));
    if (q_hi > q_lo && total) {
        unsigned blocks = (unsigned)std::min<uint64_t>(q_hi - q_lo, 148ull * 32);
        regroup_range_kernel<<<blocks, 128, 0, ctx->stream>>>(all.p, ctx->d_rank_base.p, R, nq, q_lo, q_hi, ctx->d_recv.p,
                                                             ctx->d_qoffs_c.p, ctx->d_ckey.p, ctx->d_cval.p);
        ctx->launches++;
        PHY_CUDA(ctx, cudaGetLastError());
    }
    ctx->merged_q_lo = q_lo;
    ctx->merged_q_hi = q_hi;
    return phy_merge_segments_bounded(ctx, top_n, total);
}
