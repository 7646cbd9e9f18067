// phy_internal.cuh -- shared declarations of libphylign_cuda.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/phylign_cuda.h"

#define PHY_MAX_DOCS (1u << 20)       // ref_rank field of the merge key: 20 bits
#define PHY_MAX_BATCH_RANK (1u << 12) // batch_rank field of the merge key: 12 bits
#define PHY_ROW_INVALID 0xFFFFFFFFu
#define PHY_FUSED_PLANES 10           // vertical counter planes of the fused kernel
#define PHY_FUSED_KMAX ((1u << PHY_FUSED_PLANES) - 1u)
#define PHY_SHORT_KMAX ((1u << 8) - 1u)   // 8 planes: reads up to 285 bp, pruning checkpoint every 16 rows
#define PHY_LONG_KMAX ((1u << 14) - 1u)   // 14 planes: fused up to 16383 k-mers (10 kbp reads)
#define PHY_LD_SLOTS 16               // page-locked 4 MB slots of the index file loader
#define PHY_CHUNK_BYTES 512u          // one warp-wide 128-bit load = 512 B of a row

// One resident COBS classic index, as the kernels see it.
struct DevIndex {
    const uint8_t* rows;       // signature_size x stride bytes (padding bytes are 0)
    const uint32_t* ref_rank;  // [n_docs]
    uint64_t sig;              // signature_size (< 2^32 - 1)
    uint64_t magic;            // UINT64_MAX / sig, for the exact fast modulo
    uint32_t stride;           // bytes per row in HBM (multiple of 16)
    uint32_t row_size;         // ceil(n_docs/8), bytes per row in the file
    uint32_t n_docs;
    uint32_t num_hashes;
    uint32_t batch_rank;
    uint32_t idx_id;
};

// a unit of the general (non-fused) path: k-mers [k0,k1) of query q added into score row `slot`
struct SlowItem {
    uint32_t slot, query, k0, k1;
};

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
};

struct HostIndex {
    DevIndex d{};
    std::string name;
    uint32_t term_size = 31;
    uint8_t canon = 1;
    bool alive = false, committed = false, active = true;  // active: takes part in phy_match_run
    uint64_t pushed = 0, body_bytes = 0, hbm_bytes = 0;
    uint8_t* rows_mut = nullptr;
    uint32_t* ref_rank_mut = nullptr;
    int lpr = 32;  // lanes per row chunk (1,2,4,8,16,32)
};

// host copies of the per-(query set, threshold, index set) launch tables of the gather pass
struct GatherTables {
    struct IdxClass { int c; uint32_t h; std::vector<uint32_t> ids; size_t off; };
    std::vector<IdxClass> classes;          // (lanes-per-row class, number of hash functions) -> index ids
    std::vector<uint32_t> fastq, slowq;     // fused-capable queries [10-plane | 8-plane | 14-plane]; K > 16383
    uint32_t n_fast8 = 0, n_fast10 = 0, n_fast14 = 0;
    double threshold = -1.0;
    uint32_t floor_mode = 0;
    uint64_t index_version = ~0ull;
};

struct phy_ctx {
    int device = 0, n_sm = 148;
    bool prune = true;    // PHY_NO_PRUNE=1 switches the exact threshold pruning of the ring kernel off
    bool sanitize_queries = false;  // phy_queries_set applies rule fix_query's base transform on the device
    bool pinned_results = true;  // phy_results / phy_merged blocks from the page-locked pool
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr, ev_ph[4] = {nullptr, nullptr, nullptr, nullptr};
    uint64_t budget = 0, used = 0;
    struct WsSlab { uint8_t* base; size_t cap, bump; };
    std::vector<WsSlab> ws_slabs;              // working-set arena (capi.cu: phy_ws_alloc)
    std::multimap<size_t, void*> ws_free;      // freed arena blocks by size
    std::map<void*, size_t> ws_size;
    std::string err;
    std::vector<HostIndex> idx;
    DevBuf<DevIndex> d_indexes;  // mirrors idx[] (dead slots zeroed)
    bool indexes_dirty = true;

    // pinned staging ring (index upload, query upload, result download)
    uint8_t* pin[2] = {nullptr, nullptr};
    uint8_t* d_stage[2] = {nullptr, nullptr};
    cudaEvent_t pin_ev[2] = {nullptr, nullptr};
    size_t pin_bytes = 0;
    int pin_cur = 0;

    // file loader (index_loader.cu): own upload stream, page-locked slot ring, device staging
    bool ld_ready = false;
    cudaStream_t up_stream = nullptr;
    uint8_t* ld_pin[PHY_LD_SLOTS] = {};
    cudaEvent_t ld_ev[PHY_LD_SLOTS] = {};
    uint8_t* ld_stage[2] = {nullptr, nullptr};

    // queries
    bool have_queries = false;
    uint32_t nq = 0, q_term_size = 0, q_canon = 0, q_num_hashes = 0;
    uint64_t total_bases = 0, total_kmers = 0;
    std::vector<uint64_t> h_qoffs, h_koffs;
    std::vector<uint32_t> h_nk;
    DevBuf<char> d_seq;
    DevBuf<uint64_t> d_qoffs, d_koffs, d_hashes, d_ioffs;
    DevBuf<uint32_t> d_nk, d_T, d_qlist, d_class;
    bool hashes_valid = false;
    bool hash_check_pending = false;  // K1's error word has not been read yet
    // per-query-set tables already on the device (rebuilt after phy_queries_set / a change of k, threshold, index set)
    bool kmer_tables_valid = false, ioffs_valid = false, gather_tables_valid = false;
    GatherTables gt;
    uint64_t n_hash_items = 0;                          // K1 work items of the current query set
    uint64_t index_version = 0;                         // bumped whenever the set of (active) indexes changes

    // match outputs (device resident until fetched)
    bool have_match = false, have_merged = false;
    DevBuf<phy_unit> d_units, d_units_sorted;
    DevBuf<uint32_t> d_unit_flag, d_unit_id;
    DevBuf<uint64_t> d_unit_pos;
    bool units_ordered = false;
    DevBuf<phy_hit> d_hits;
    DevBuf<unsigned long long> d_idx_bytes; // [index slot] row bytes gathered by the last match
    std::vector<unsigned long long> h_idx_bytes;
    DevBuf<unsigned long long> d_counters;  // [0] n_hits [1] n_units [2] error info
    DevBuf<uint32_t> d_qcount;              // kept hits per query (merge sizing)
    uint64_t n_units = 0, n_hits = 0;
    DevBuf<uint32_t> d_scores;              // general path: [slots][n_docs]
    DevBuf<SlowItem> d_items;
    DevBuf<uint32_t> d_slotq;
    // merge
    DevBuf<uint64_t> d_ckey, d_qoffs_c, d_foffs, d_scan_tmp;
    DevBuf<uint32_t> d_cval, d_qcursor, d_nfinal;
    DevBuf<phy_cand> d_final;
    uint64_t n_final = 0;
    uint32_t merge_top_n = 0;

    uint64_t h2d_bytes = 0, launches = 0, gathered_bytes = 0;
    float phase_ms[4] = {0, 0, 0, 0};
    DevBuf<uint8_t> d_flush;

    // multi-GPU
    DevBuf<uint64_t> d_foffs_all, d_rank_base;
    DevBuf<phy_cand> d_recv;
    void* nccl_comm = nullptr;
    int rank = 0, n_ranks = 1;
    bool merge_sharded = false;        // "merge_mode" 1: every rank finalises its slice of the queries
    bool shard_query_upload = false;   // every rank uploads 1/R of the bases, NCCL all-gather completes them
    uint32_t merged_q_lo = 0, merged_q_hi = 0;  // queries whose merged lists this context holds
};

// ---- error helpers ---------------------------------------------------------------
void phy_set_error(phy_ctx* ctx, const char* fmt, ...);
#define PHY_CUDA(ctx, call)                                                              \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            phy_set_error(ctx, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),   \
                          __FILE__, __LINE__);                                           \
            return PHY_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)
#define PHY_TRY(call)              \
    do {                           \
        int r_ = (call);           \
        if (r_ != PHY_OK) return r_; \
    } while (0)

int phy_dev_alloc(phy_ctx* ctx, void** p, size_t bytes, bool counted);
void phy_dev_free(phy_ctx* ctx, void* p, size_t bytes, bool counted);
int phy_ws_alloc(phy_ctx* ctx, void** p, size_t bytes);   // working buffers: carved from large slabs
void phy_ws_free(phy_ctx* ctx, void* p);
template <class T>
int phy_ensure(phy_ctx* ctx, DevBuf<T>& b, size_t n) {
    if (n <= b.cap && b.p) return PHY_OK;
    size_t want = n + n / 4 + 16;
    if (b.p) phy_ws_free(ctx, b.p);
    b.p = nullptr;
    b.cap = 0;
    void* p = nullptr;
    int r = phy_ws_alloc(ctx, &p, want * sizeof(T));
    if (r != PHY_OK) return r;
    b.p = (T*)p;
    b.cap = want;
    return PHY_OK;
}
template <class T>
void phy_release(phy_ctx* ctx, DevBuf<T>& b) {
    if (b.p) phy_ws_free(ctx, b.p);
    b.p = nullptr;
    b.cap = 0;
}

int phy_sync_indexes(phy_ctx* ctx);  // upload DevIndex table if dirty
int phy_h2d(phy_ctx* ctx, void* dst, const void* src, size_t bytes);  // via pinned ring
int phy_d2h(phy_ctx* ctx, void* dst, const void* src, size_t bytes);

// ---- kernels' host launchers (one per .cu) -----------------------------------------
int phy_launch_hash(phy_ctx* ctx);
int phy_launch_fix_bases(phy_ctx* ctx, uint8_t* d_bases, uint64_t n);
int phy_check_hash_error(phy_ctx* ctx);                          // sync + read K1's error word if pending
int phy_hash_error_of(phy_ctx* ctx, unsigned long long word);    // interpret a word that was already copied
int phy_launch_gather(phy_ctx* ctx, const phy_match_params* p);
int phy_launch_scores(phy_ctx* ctx, int idx_id, uint32_t* d_out_scores);
int phy_launch_sort_units(phy_ctx* ctx);
int phy_launch_merge(phy_ctx* ctx, uint32_t top_n);
int phy_nccl_merge(phy_ctx* ctx, uint32_t top_n);
int phy_nccl_allgather_inplace(phy_ctx* ctx, void* buf, size_t slice_bytes);
int phy_merge_segments(phy_ctx* ctx, uint32_t top_n);
int phy_exscan(phy_ctx* ctx, const uint32_t* d_in, uint64_t n, uint64_t* d_out, uint64_t* total_host);
int phy_lpr_for_stride(uint32_t stride);
int phy_order_units(phy_ctx* ctx, uint64_t cells);
int phy_restride_chunk(phy_ctx* ctx, HostIndex& ix, const uint8_t* d_src, uint64_t body_off,
                       uint64_t nbytes);
int phy_destride(phy_ctx* ctx, const HostIndex& ix, uint8_t* d_dst);

// ---- device helpers ----------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t phy_fastmod(uint64_t n, uint64_t d, uint64_t magic) {
    // exact n % d for every 64-bit n: q' = floor(n*magic/2^64) is floor(n/d) or one less
    uint64_t q = __umul64hi(n, magic);
    uint64_t r = n - q * d;
    if (r >= d) r -= d;
    return (uint32_t)r;
}
#endif
