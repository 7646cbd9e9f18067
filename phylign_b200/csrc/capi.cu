// capi.cu -- the C ABI of libphylign_cuda.so (include/phylign_cuda.h): context, HBM
// accounting, index store, query upload, match driver, result download.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>

#include <algorithm>

#include "phy_internal.cuh"

int phy_synth_build(phy_ctx* ctx, HostIndex& ix, const phy_synth_spec* spec);
int phy_synth_reads_dev(phy_ctx* ctx, const phy_synth_spec* d_specs, uint32_t n_specs, uint64_t reads_seed,
                        uint64_t first_read, uint32_t n_reads, uint32_t read_len, uint32_t random_q8,
                        uint32_t err_q16, char* d_out);

static thread_local std::string g_err;

void phy_set_error(phy_ctx* ctx, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    if (ctx) ctx->err = buf;
}

extern "C" const char* phy_last_error(const phy_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }
extern "C" int phy_abi_version(void) { return PHY_ABI_VERSION; }

extern "C" int phy_device_count(int* n) {
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (n) *n = e == cudaSuccess ? c : 0;
    if (e != cudaSuccess || c == 0) {
        phy_set_error(nullptr, "no CUDA device: %s", cudaGetErrorString(e));
        return PHY_ERR_CUDA;
    }
    return PHY_OK;
}

// ---- memory -----------------------------------------------------------------------
int phy_dev_alloc(phy_ctx* ctx, void** p, size_t bytes, bool counted) {
    if (bytes == 0) bytes = 16;
    if (counted && ctx->used + bytes > ctx->budget) {
        phy_set_error(ctx, "HBM budget exhausted: %llu + %llu > %llu bytes", (unsigned long long)ctx->used,
                      (unsigned long long)bytes, (unsigned long long)ctx->budget);
        return PHY_ERR_NOMEM;
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        phy_set_error(ctx, "cudaMalloc(%llu) failed: %s", (unsigned long long)bytes, cudaGetErrorString(e));
        return PHY_ERR_NOMEM;
    }
    if (counted) ctx->used += bytes;
    return PHY_OK;
}

void phy_dev_free(phy_ctx* ctx, void* p, size_t bytes, bool counted) {
    if (!p) return;
    cudaFree(p);
    if (counted) ctx->used -= std::min<uint64_t>(ctx->used, bytes ? bytes : 16);
}

// ---- working-set arena ---------------------------------------------------------------------------
// The ~25 working buffers of a match pass (hashes, unit tables, hit lists, merge keys ...) are carved
// out of a few large slabs instead of one cudaMalloc each: a cudaMalloc / cudaFree costs 2-7 ms here
// regardless of size (measured), which made the first pass of a fresh context ~100 ms slower than the
// steady state.  Freed blocks go to a size-ordered free list (buffers only ever grow).  Index rows keep
// their own allocations (they are evicted individually).
static const size_t WS_SLAB = 256u << 20;
int phy_ws_alloc(phy_ctx* ctx, void** out, size_t bytes) {
    bytes = (std::max<size_t>(bytes, 256) + 255) / 256 * 256;
    auto it = ctx->ws_free.lower_bound(bytes);
    if (it != ctx->ws_free.end() && it->first <= bytes * 2) {
        *out = it->second;
        ctx->ws_free.erase(it);
        return PHY_OK;
    }
    for (auto& sl : ctx->ws_slabs)
        if (sl.cap - sl.bump >= bytes) {
            *out = sl.base + sl.bump;
            sl.bump += bytes;
            ctx->ws_size[*out] = bytes;
            return PHY_OK;
        }
    const size_t cap = bytes > WS_SLAB / 2 ? bytes : WS_SLAB;
    void* p = nullptr;
    PHY_TRY(phy_dev_alloc(ctx, &p, cap, true));
    ctx->ws_slabs.push_back(phy_ctx::WsSlab{(uint8_t*)p, cap, bytes});
    ctx->ws_size[p] = bytes;
    *out = p;
    return PHY_OK;
}
void phy_ws_free(phy_ctx* ctx, void* p) {
    if (!p) return;
    auto it = ctx->ws_size.find(p);
    if (it != ctx->ws_size.end()) ctx->ws_free.insert({it->second, p});
}

static const size_t PIN_BYTES = 32u << 20;

extern "C" int phy_ctx_create(phy_ctx** out, int device, uint64_t hbm_budget) {
    if (!out) return PHY_ERR_ARG;
    *out = nullptr;
    int n = 0;
    PHY_TRY(phy_device_count(&n));
    if (device < 0 || device >= n) {
        phy_set_error(nullptr, "device %d out of range (%d visible)", device, n);
        return PHY_ERR_ARG;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
        phy_set_error(nullptr, "device %d is not sm_100 (this library has no other code path)", device);
        return PHY_ERR_CUDA;
    }
    phy_ctx* ctx = new phy_ctx();
    ctx->idx.reserve(PHY_MAX_BATCH_RANK);  // entries never move: a file load may run beside a match (index_loader.cu)
    ctx->device = device;
    ctx->n_sm = prop.multiProcessorCount;
    if (getenv("PHY_NO_PRUNE") && atoi(getenv("PHY_NO_PRUNE"))) ctx->prune = false;
    PHY_CUDA(ctx, cudaSetDevice(device));
    size_t fr = 0, tot = 0;
    PHY_CUDA(ctx, cudaMemGetInfo(&fr, &tot));
    const uint64_t margin = 2ull << 30;
    uint64_t avail = fr > margin ? fr - margin : fr / 2;
    ctx->budget = hbm_budget ? std::min<uint64_t>(hbm_budget, avail) : avail;
    PHY_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    PHY_CUDA(ctx, cudaEventCreate(&ctx->ev_t0));
    PHY_CUDA(ctx, cudaEventCreate(&ctx->ev_t1));
    for (int i = 0; i < 4; i++) PHY_CUDA(ctx, cudaEventCreate(&ctx->ev_ph[i]));
    ctx->pin_bytes = PIN_BYTES;
    for (int i = 0; i < 2; i++) {
        PHY_CUDA(ctx, cudaHostAlloc((void**)&ctx->pin[i], ctx->pin_bytes, cudaHostAllocDefault));
        PHY_CUDA(ctx, cudaMalloc((void**)&ctx->d_stage[i], ctx->pin_bytes));
        PHY_CUDA(ctx, cudaEventCreateWithFlags(&ctx->pin_ev[i], cudaEventDisableTiming));
    }
    *out = ctx;
    return PHY_OK;
}

void phy_nccl_shutdown(phy_ctx* ctx);
void phy_loader_destroy(phy_ctx* ctx);

extern "C" void phy_ctx_destroy(phy_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    phy_loader_destroy(ctx);
    phy_nccl_shutdown(ctx);  // no-op after phy_nccl_finalize; otherwise a local, non-blocking abort
    for (auto& ix : ctx->idx) {
        if (ix.rows_mut) cudaFree(ix.rows_mut);
        if (ix.ref_rank_mut) cudaFree(ix.ref_rank_mut);
    }
    for (auto& sl : ctx->ws_slabs) cudaFree(sl.base);  // every working buffer lives in these
    for (int i = 0; i < 2; i++) {
        if (ctx->pin[i]) cudaFreeHost(ctx->pin[i]);
        if (ctx->d_stage[i]) cudaFree(ctx->d_stage[i]);
        if (ctx->pin_ev[i]) cudaEventDestroy(ctx->pin_ev[i]);
    }
    cudaEventDestroy(ctx->ev_t0);
    cudaEventDestroy(ctx->ev_t1);
    for (int i = 0; i < 4; i++) cudaEventDestroy(ctx->ev_ph[i]);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

// ---- process-wide pool of pinned host blocks (results, caller buffers) -------------------------
// cudaHostAlloc/cudaFreeHost cost milliseconds; result buffers of consecutive steps have the
// same size class, so freed blocks are kept and handed out again.
#include <map>
#include <mutex>
static std::mutex g_pin_mu;
static std::multimap<size_t, void*> g_pin_free;      // capacity -> block
static std::map<void*, size_t> g_pin_cap;            // every live or cached block
static size_t pin_class(size_t n) {
    size_t c = 4096;
    while (c < n) c <<= 1;
    return c;
}
void* phy_pinned_alloc(size_t bytes) {
    const size_t cap = pin_class(bytes ? bytes : 1);
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        auto it = g_pin_free.find(cap);
        if (it != g_pin_free.end()) {
            void* p = it->second;
            g_pin_free.erase(it);
            return p;
        }
    }
    void* p = nullptr;
    if (cudaHostAlloc(&p, cap, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    std::lock_guard<std::mutex> lk(g_pin_mu);
    g_pin_cap[p] = cap;
    return p;
}
void phy_pinned_free(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_pin_mu);
    auto it = g_pin_cap.find(p);
    if (it == g_pin_cap.end()) return;
    size_t cached = 0;
    for (auto& kv : g_pin_free) cached += kv.first;
    if (cached + it->second > (size_t(4) << 30)) {   // keep at most 4 GiB parked
        cudaFreeHost(p);
        g_pin_cap.erase(it);
        return;
    }
    g_pin_free.insert({it->second, p});
}
// result blocks: page-locked from the pool (default) or plain malloc ("pinned_results" 0)
static void* result_alloc(phy_ctx* ctx, size_t bytes) {
    if (ctx->pinned_results) return phy_pinned_alloc(bytes);
    if (bytes < (4u << 20)) return malloc(bytes ? bytes : 1);
    // large one-shot blocks: 2 MB aligned + transparent huge pages, so the first touch by the
    // download costs one fault per 2 MB instead of one per 4 KB
    void* p = nullptr;
    const size_t huge = 2u << 20;
    if (posix_memalign(&p, huge, (bytes + huge - 1) / huge * huge) != 0) return nullptr;
    madvise(p, (bytes + huge - 1) / huge * huge, MADV_HUGEPAGE);
    return p;
}
static void result_free(void* p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        if (g_pin_cap.find(p) == g_pin_cap.end()) {
            free(p);
            return;
        }
    }
    phy_pinned_free(p);
}
static bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

extern "C" int phy_host_alloc(size_t bytes, void** out) {
    if (!out) return PHY_ERR_ARG;
    *out = phy_pinned_alloc(bytes);
    if (!*out) {
        phy_set_error(nullptr, "cannot allocate %llu bytes of pinned host memory", (unsigned long long)bytes);
        return PHY_ERR_NOMEM;
    }
    return PHY_OK;
}
extern "C" void phy_host_free(void* p) { phy_pinned_free(p); }

// host -> device; src is free again on return.  Pinned sources (phy_host_alloc) are DMA'd
// directly, pageable ones go through the pinned ring.
int phy_h2d(phy_ctx* ctx, void* dst, const void* src, size_t bytes) {
    const uint8_t* s = (const uint8_t*)src;
    uint8_t* d = (uint8_t*)dst;
    if (bytes >= (1u << 16) && is_pinned(src)) {
        PHY_CUDA(ctx, cudaMemcpyAsync(d, s, bytes, cudaMemcpyHostToDevice, ctx->stream));
        PHY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->h2d_bytes += bytes;
        return PHY_OK;
    }
    while (bytes) {
        size_t n = std::min(bytes, ctx->pin_bytes);
        int slot = ctx->pin_cur;
        ctx->pin_cur ^= 1;
        PHY_CUDA(ctx, cudaEventSynchronize(ctx->pin_ev[slot]));
        memcpy(ctx->pin[slot], s, n);
        PHY_CUDA(ctx, cudaMemcpyAsync(d, ctx->pin[slot], n, cudaMemcpyHostToDevice, ctx->stream));
        PHY_CUDA(ctx, cudaEventRecord(ctx->pin_ev[slot], ctx->stream));
        s += n; d += n; bytes -= n;
        ctx->h2d_bytes += n;
    }
    return PHY_OK;
}

// device -> host (pageable) through the pinned ring; synchronous
int phy_d2h(phy_ctx* ctx, void* dst, const void* src, size_t bytes) {
    uint8_t* d = (uint8_t*)dst;
    const uint8_t* s = (const uint8_t*)src;
    if (bytes >= (1u << 16) && is_pinned(dst)) {
        PHY_CUDA(ctx, cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        PHY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return PHY_OK;
    }
    // two slots in flight: copy chunk i+1 while chunk i is memcpy'd out
    size_t off = 0, pend_n[2] = {0, 0}, pend_off[2] = {0, 0};
    int slot = 0;
    PHY_CUDA(ctx, cudaEventSynchronize(ctx->pin_ev[0]));
    PHY_CUDA(ctx, cudaEventSynchronize(ctx->pin_ev[1]));
    while (off < bytes || pend_n[0] || pend_n[1]) {
        if (off < bytes && pend_n[slot] == 0) {
            size_t n = std::min(bytes - off, ctx->pin_bytes);
            PHY_CUDA(ctx, cudaMemcpyAsync(ctx->pin[slot], s + off, n, cudaMemcpyDeviceToHost, ctx->stream));
            PHY_CUDA(ctx, cudaEventRecord(ctx->pin_ev[slot], ctx->stream));
            pend_n[slot] = n; pend_off[slot] = off;
            off += n;
        }
        slot ^= 1;
        if (pend_n[slot]) {
            PHY_CUDA(ctx, cudaEventSynchronize(ctx->pin_ev[slot]));
            memcpy(d + pend_off[slot], ctx->pin[slot], pend_n[slot]);
            pend_n[slot] = 0;
        }
    }
    return PHY_OK;
}

// ---- index store --------------------------------------------------------------------
// Row stride in HBM: a power of two up to 128 B (a row then never straddles a 128-B line:
// measured 17.8 ms vs 15.3 ms for 83-B rows at stride 96 vs 128, profiles/r01_sweep_docs_*),
// a multiple of 32 B (the DRAM sector) above.
static uint32_t stride_for(uint32_t row_size) {
    if (row_size <= 128) {
        uint32_t s = 16;
        while (s < row_size) s <<= 1;
        return s;
    }
    return (row_size + 31u) / 32u * 32u;
}

static HostIndex* get_index(phy_ctx* ctx, int idx_id) {
    if (!ctx || idx_id < 0 || (size_t)idx_id >= ctx->idx.size() || !ctx->idx[idx_id].alive) {
        phy_set_error(ctx, "unknown index id %d", idx_id);
        return nullptr;
    }
    return &ctx->idx[idx_id];
}

extern "C" int phy_index_begin(phy_ctx* ctx, const char* batch_name, uint32_t term_size, uint8_t canonicalize,
                               uint64_t signature_size, uint64_t num_hashes, uint32_t n_docs, int* idx_id) {
    if (!ctx || !idx_id) return PHY_ERR_ARG;
    if (term_size == 0 || term_size > 31) {
        phy_set_error(ctx, "term_size %u unsupported (1..31)", term_size);
        return PHY_ERR_ARG;
    }
    if (n_docs == 0 || n_docs > PHY_MAX_DOCS || signature_size == 0 || signature_size >= 0xFFFFFFFFull ||
        num_hashes == 0 || num_hashes > 16) {
        phy_set_error(ctx, "index shape unsupported: docs=%u (max %u) signature_size=%llu (max 2^32-2) hashes=%llu",
                      n_docs, PHY_MAX_DOCS, (unsigned long long)signature_size, (unsigned long long)num_hashes);
        return PHY_ERR_ARG;
    }
    PHY_CUDA(ctx, cudaSetDevice(ctx->device));
    HostIndex ix;
    ix.name = batch_name ? batch_name : "";
    ix.term_size = term_size;
    ix.canon = canonicalize ? 1 : 0;
    ix.d.sig = signature_size;
    ix.d.magic = 0xFFFFFFFFFFFFFFFFull / signature_size;
    ix.d.row_size = (n_docs + 7u) / 8u;
    ix.d.stride = stride_for(ix.d.row_size);
    ix.d.n_docs = n_docs;
    ix.d.num_hashes = (uint32_t)num_hashes;
    ix.body_bytes = signature_size * ix.d.row_size;
    ix.hbm_bytes = signature_size * ix.d.stride + 512;  // tail pad: a masked-off lane never faults
    ix.lpr = phy_lpr_for_stride(ix.d.stride);
    void* p = nullptr;
    PHY_TRY(phy_dev_alloc(ctx, &p, ix.hbm_bytes, true));
    ix.rows_mut = (uint8_t*)p;
    ix.d.rows = ix.rows_mut;
    int r = phy_dev_alloc(ctx, &p, (size_t)n_docs * sizeof(uint32_t), true);
    if (r != PHY_OK) {
        phy_dev_free(ctx, ix.rows_mut, ix.hbm_bytes, true);
        return r;
    }
    ix.ref_rank_mut = (uint32_t*)p;
    ix.d.ref_rank = ix.ref_rank_mut;
    PHY_CUDA(ctx, cudaMemsetAsync(ix.rows_mut, 0, ix.hbm_bytes, ctx->stream));
    std::vector<uint32_t> ident(n_docs);
    for (uint32_t d = 0; d < n_docs; d++) ident[d] = d;
    PHY_TRY(phy_h2d(ctx, ix.ref_rank_mut, ident.data(), ident.size() * sizeof(uint32_t)));
    PHY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // rows are zero before any stream (push / file loader) fills them
    ix.alive = true;
    // reuse a dead slot if any
    size_t slot = ctx->idx.size();
    for (size_t i = 0; i < ctx->idx.size(); i++)
        if (!ctx->idx[i].alive) { slot = i; break; }
    if (slot == ctx->idx.size()) {
        if (ctx->idx.size() >= PHY_MAX_BATCH_RANK) {
            phy_dev_free(ctx, ix.rows_mut, ix.hbm_bytes, true);
            phy_dev_free(ctx, ix.ref_rank_mut, (size_t)n_docs * sizeof(uint32_t), true);
            phy_set_error(ctx, "more than %u indexes in one context", PHY_MAX_BATCH_RANK);
            return PHY_ERR_ARG;
        }
        ctx->idx.push_back(HostIndex());
    }
    ix.d.idx_id = (uint32_t)slot;
    ix.d.batch_rank = (uint32_t)slot % PHY_MAX_BATCH_RANK;
    ctx->idx[slot] = ix;
    ctx->indexes_dirty = true;
    ctx->index_version++;
    *idx_id = (int)slot;
    return PHY_OK;
}

extern "C" int phy_index_push(phy_ctx* ctx, int idx_id, const void* host_chunk, uint64_t nbytes) {
    HostIndex* ix = get_index(ctx, idx_id);
    if (!ix) return PHY_ERR_ARG;
    if (ix->committed || ix->pushed + nbytes > ix->body_bytes) {
        phy_set_error(ctx, "index %d: %llu bytes pushed beyond the body size %llu", idx_id,
                      (unsigned long long)(ix->pushed + nbytes), (unsigned long long)ix->body_bytes);
        return PHY_ERR_STATE;
    }
    PHY_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint8_t* s = (const uint8_t*)host_chunk;
    const bool pinned_src = nbytes >= (1u << 16) && is_pinned(host_chunk);
    while (nbytes) {
        size_t n = (size_t)std::min<uint64_t>(nbytes, ctx->pin_bytes);
        int slot = ctx->pin_cur;
        ctx->pin_cur ^= 1;
        PHY_CUDA(ctx, cudaEventSynchronize(ctx->pin_ev[slot]));
        if (pinned_src) {  // caller's buffer is page-locked (phy_host_alloc): DMA straight from it
            PHY_CUDA(ctx, cudaMemcpyAsync(ctx->d_stage[slot], s, n, cudaMemcpyHostToDevice, ctx->stream));
            PHY_CUDA(ctx, cudaEventRecord(ctx->ev_t1, ctx->stream));
            PHY_TRY(phy_restride_chunk(ctx, *ix, ctx->d_stage[slot], ix->pushed, n));
            PHY_CUDA(ctx, cudaEventRecord(ctx->pin_ev[slot], ctx->stream));
            PHY_CUDA(ctx, cudaEventSynchronize(ctx->ev_t1));  // the caller may refill its buffer now
            ix->pushed += n; s += n; nbytes -= n; ctx->h2d_bytes += n;
            continue;
        }
        memcpy(ctx->pin[slot], s, n);
        PHY_CUDA(ctx, cudaMemcpyAsync(ctx->d_stage[slot], ctx->pin[slot], n, cudaMemcpyHostToDevice, ctx->stream));
        PHY_TRY(phy_restride_chunk(ctx, *ix, ctx->d_stage[slot], ix->pushed, n));
        PHY_CUDA(ctx, cudaEventRecord(ctx->pin_ev[slot], ctx->stream));
        ix->pushed += n;
        s += n;
        nbytes -= n;
        ctx->h2d_bytes += n;
    }
    return PHY_OK;
}

extern "C" int phy_index_commit(phy_ctx* ctx, int idx_id) {
    HostIndex* ix = get_index(ctx, idx_id);
    if (!ix) return PHY_ERR_ARG;
    if (ix->pushed != ix->body_bytes) {
        phy_set_error(ctx, "index %d (%s): body holds %llu bytes, expected signature_size*row_size = %llu", idx_id,
                      ix->name.c_str(), (unsigned long long)ix->pushed, (unsigned long long)ix->body_bytes);
        return PHY_ERR_STATE;
    }
    PHY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->up_stream) PHY_CUDA(ctx, cudaStreamSynchronize(ctx->up_stream));
    ix->committed = true;
    ctx->indexes_dirty = true;
    ctx->index_version++;
    return PHY_OK;
}

extern "C" int phy_index_evict(phy_ctx* ctx, int idx_id) {
    HostIndex* ix = get_index(ctx, idx_id);
    if (!ix) return PHY_ERR_ARG;
    PHY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    phy_dev_free(ctx, ix->rows_mut, ix->hbm_bytes, true);
    phy_dev_free(ctx, ix->ref_rank_mut, (size_t)ix->d.n_docs * sizeof(uint32_t), true);
    *ix = HostIndex();
    ctx->indexes_dirty = true;
    ctx->index_version++;
    ctx->have_match = ctx->have_merged = false;
    return PHY_OK;
}

extern "C" int phy_index_set_ranks(phy_ctx* ctx, int idx_id, uint32_t batch_rank, const uint32_t* ref_rank) {
    HostIndex* ix = get_index(ctx, idx_id);
    if (!ix) return PHY_ERR_ARG;
    if (batch_rank >= PHY_MAX_BATCH_RANK) {
        phy_set_error(ctx, "batch_rank %u >= %u", batch_rank, PHY_MAX_BATCH_RANK);
        return PHY_ERR_ARG;
    }
    if (ref_rank) {
        for (uint32_t d = 0; d < ix->d.n_docs; d++)
            if (ref_rank[d] >= PHY_MAX_DOCS) {
                phy_set_error(ctx, "ref_rank[%u] out of range", d);
                return PHY_ERR_ARG;
            }
        PHY_TRY(phy_h2d(ctx, ix->ref_rank_mut, ref_rank, (size_t)ix->d.n_docs * sizeof(uint32_t)));
        PHY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    ix->d.batch_rank = batch_rank;
    ctx->indexes_dirty = true;
    ctx->index_version++;
    return PHY_OK;
}

extern "C" int phy_index_set_active(phy_ctx* ctx, int idx_id, int active) {
    HostIndex* ix = get_index(ctx, idx_id);
    if (!ix) return PHY_ERR_ARG;
    ctx->index_version++;
    ix->active = active != 0;  // takes effect with the next phy_match_run; fetched / fetchable results stay valid
    return PHY_OK;
}

extern "C" int phy_index_info_get(phy_ctx* ctx, int idx_id, phy_index_info* out) {
    HostIndex* ix = get_index(ctx, idx_id);
    if (!ix || !out) return PHY_ERR_ARG;
    out->signature_size = ix->d.sig;
    out->num_hashes = ix->d.num_hashes;
    out->hbm_bytes = ix->hbm_bytes;
    out->term_size = ix->term_size;
    out->n_docs = ix->d.n_docs;
    out->row_size = ix->d.row_size;
    out->row_stride = ix->d.stride;
    out->batch_rank = ix->d.batch_rank;
    out->canonicalize = ix->canon;
    out->committed = ix->committed;
    return PHY_OK;
}

extern "C" int phy_index_count(phy_ctx* ctx, int* n_resident) {
    if (!ctx || !n_resident) return PHY_ERR_ARG;
    int n = 0;
    for (auto& ix : ctx->idx) n += ix.alive && ix.committed;
    *n_resident = n;
    return PHY_OK;
}

extern "C" int phy_index_download(phy_ctx* ctx, int idx_id, void* host_out, uint64_t nbytes) {
    HostIndex* ix = get_index(ctx, idx_id);
    if (!ix || !host_out) return PHY_ERR_ARG;
    if (nbytes != ix->body_bytes) {
        phy_set_error(ctx, "index %d body is %llu bytes", idx_id, (unsigned long long)ix->body_bytes);
        return PHY_ERR_ARG;
    }
    PHY_CUDA(ctx, cudaSetDevice(ctx->device));
    void* tmp = nullptr;
    PHY_TRY(phy_dev_alloc(ctx, &tmp, nbytes, false));
    int r = phy_destride(ctx, *ix, (uint8_t*)tmp);
    if (r == PHY_OK) r = phy_d2h(ctx, host_out, tmp, nbytes);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(tmp);
    return r;
}

int phy_sync_indexes(phy_ctx* ctx) {
    if (!ctx->indexes_dirty && ctx->d_indexes.p) return PHY_OK;
    std::vector<DevIndex> tab(std::max<size_t>(ctx->idx.size(), 1));
    memset(tab.data(), 0, tab.size() * sizeof(DevIndex));
    for (size_t i = 0; i < ctx->idx.size(); i++)
        if (ctx->idx[i].alive) tab[i] = ctx->idx[i].d;
    PHY_TRY(phy_ensure(ctx, ctx->d_indexes, tab.size()));
    PHY_TRY(phy_h2d(ctx, ctx->d_indexes.p, tab.data(), tab.size() * sizeof(DevIndex)));
    ctx->indexes_dirty = false;
    return PHY_OK;
}

// ---- queries ---------------------------------------------------------------------------
extern "C" int phy_queries_set(phy_ctx* ctx, const char* seq_concat, const uint64_t* offs, uint32_t nq) {
    if (!ctx || (!seq_concat && nq && offs[nq] > 0) || !offs) return PHY_ERR_ARG;
    PHY_CUDA(ctx, cudaSetDevice(ctx->device));
    for (uint32_t q = 0; q < nq; q++)
        if (offs[q + 1] < offs[q]) {
            phy_set_error(ctx, "query offsets must be non-decreasing");
            return PHY_ERR_ARG;
        }
    ctx->nq = nq;
    ctx->h_qoffs.assign(offs, offs + nq + 1);
    ctx->total_bases = offs[nq] - offs[0];
    ctx->have_queries = true;
    ctx->kmer_tables_valid = ctx->ioffs_valid = ctx->gather_tables_valid = false;
    ctx->hashes_valid = false;
    ctx->have_match = ctx->have_merged = false;
    ctx->q_term_size = 0;  // k-mer tables are (re)built by phy_match_run for the resident indexes' k
    PHY_TRY(phy_ensure(ctx, ctx->d_qoffs, nq + 2));
    if (ctx->shard_query_upload && ctx->n_ranks > 1 && ctx->total_bases >= (1u << 20)) {
        // every rank of the job passes the same queries: upload 1/R of the bases over this GPU's PCIe
        // link and let the slices meet over NVLink (R x less host traffic per step)
        const uint64_t R = (uint64_t)ctx->n_ranks;
        const uint64_t slice = ((ctx->total_bases + R - 1) / R + 255) / 256 * 256;
        PHY_TRY(phy_ensure(ctx, ctx->d_seq, slice * R + 64));
        const uint64_t lo = std::min<uint64_t>(slice * (uint64_t)ctx->rank, ctx->total_bases);
        const uint64_t hi = std::min<uint64_t>(lo + slice, ctx->total_bases);
        if (hi > lo) PHY_TRY(phy_h2d(ctx, ctx->d_seq.p + lo, seq_concat + offs[0] + lo, hi - lo));
        PHY_TRY(phy_nccl_allgather_inplace(ctx, ctx->d_seq.p, slice));
    } else {
        PHY_TRY(phy_ensure(ctx, ctx->d_seq, ctx->total_bases + 64));
        if (ctx->total_bases) PHY_TRY(phy_h2d(ctx, ctx->d_seq.p, seq_concat + offs[0], ctx->total_bases));
    }
    if (offs[0] != 0)
        for (auto& o : ctx->h_qoffs) o -= offs[0];
    PHY_TRY(phy_h2d(ctx, ctx->d_qoffs.p, ctx->h_qoffs.data(), (nq + 1) * sizeof(uint64_t)));
    if (ctx->sanitize_queries) PHY_TRY(phy_launch_fix_bases(ctx, (uint8_t*)ctx->d_seq.p, ctx->total_bases));
    return PHY_OK;
}

extern "C" int phy_fix_bases(phy_ctx* ctx, char* bases, uint64_t n) {
    if (!ctx || (!bases && n)) return PHY_ERR_ARG;
    PHY_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf<char> tmp;
    PHY_TRY(phy_ensure(ctx, tmp, n + 64));
    int rc = phy_h2d(ctx, tmp.p, bases, n);
    if (rc == PHY_OK) rc = phy_launch_fix_bases(ctx, (uint8_t*)tmp.p, n);
    if (rc == PHY_OK) rc = phy_d2h(ctx, bases, tmp.p, n);
    cudaStreamSynchronize(ctx->stream);
    phy_release(ctx, tmp);
    return rc;
}

// k-mer offset tables + hashes for the term size / canonical flag of the resident indexes
static int prepare_hashes(phy_ctx* ctx, uint32_t k, uint32_t canon, uint32_t nh) {
    if (ctx->hashes_valid && ctx->q_term_size == k && ctx->q_canon == canon && ctx->q_num_hashes >= nh) return PHY_OK;
    const uint32_t nq = ctx->nq;
    // the k-mer offset tables depend on the query set and k only: built and uploaded once per
    // phy_queries_set, not once per pass (the hashes themselves are recomputed by every pass)
    if (!ctx->kmer_tables_valid || ctx->q_term_size != k) {
        ctx->h_koffs.resize(nq + 1);
        ctx->h_nk.resize(nq + 1);
        uint64_t run = 0;
        for (uint32_t q = 0; q < nq; q++) {
            uint64_t L = ctx->h_qoffs[q + 1] - ctx->h_qoffs[q];
            uint64_t K = L >= k ? L - k + 1 : 0;
            if (K > 0xFFFFFFF0ull) {
                phy_set_error(ctx, "query #%u too long", q);
                return PHY_ERR_ARG;
            }
            ctx->h_koffs[q] = run;
            ctx->h_nk[q] = (uint32_t)K;
            run += K;
        }
        ctx->h_koffs[nq] = run;
        ctx->h_nk[nq] = 0;
        ctx->total_kmers = run;
        ctx->q_term_size = k;
        PHY_TRY(phy_ensure(ctx, ctx->d_koffs, nq + 2));
        PHY_TRY(phy_ensure(ctx, ctx->d_nk, nq + 2));
        PHY_TRY(phy_h2d(ctx, ctx->d_koffs.p, ctx->h_koffs.data(), (nq + 1) * sizeof(uint64_t)));
        PHY_TRY(phy_h2d(ctx, ctx->d_nk.p, ctx->h_nk.data(), (nq + 1) * sizeof(uint32_t)));
        ctx->kmer_tables_valid = true;
        ctx->ioffs_valid = false;
        ctx->gather_tables_valid = false;
    }
    ctx->q_canon = canon;
    ctx->q_num_hashes = nh;
    return phy_launch_hash(ctx);
}

static int resident_shape(phy_ctx* ctx, uint32_t* k, uint32_t* canon, uint32_t* nh, int only_idx) {
    bool first = true;
    *nh = 0;
    for (size_t i = 0; i < ctx->idx.size(); i++) {
        const HostIndex& ix = ctx->idx[i];
        if (!ix.alive || !ix.committed) continue;
        if (only_idx >= 0 ? (int)i != only_idx : !ix.active) continue;
        if (first) { *k = ix.term_size; *canon = ix.canon; first = false; }
        else if (*k != ix.term_size || *canon != ix.canon) {
            phy_set_error(ctx, "resident indexes disagree on term_size/canonicalize (cobs requires them equal)");
            return PHY_ERR_STATE;
        }
        *nh = std::max(*nh, ix.d.num_hashes);
    }
    if (first) {
        phy_set_error(ctx, "no committed (and active) index resident");
        return PHY_ERR_STATE;
    }
    return PHY_OK;
}

// ---- match -----------------------------------------------------------------------------------
extern "C" int phy_match_run(phy_ctx* ctx, const phy_match_params* p, uint32_t merge_top_n) {
    if (!ctx || !p) return PHY_ERR_ARG;
    if (!ctx->have_queries) {
        phy_set_error(ctx, "phy_queries_set must be called before phy_match_run");
        return PHY_ERR_STATE;
    }
    if (!(p->threshold >= 0.0) || p->threshold > 1.0e9) {
        phy_set_error(ctx, "threshold must be a non-negative number");
        return PHY_ERR_ARG;
    }
    PHY_CUDA(ctx, cudaSetDevice(ctx->device));
    uint32_t k = 0, canon = 0, nh = 0;
    int n_active = 0;
    for (auto& ix : ctx->idx) n_active += ix.alive && ix.committed && ix.active;
    // a rank of a multi-GPU job may hold no index in some round: it still takes part in the merge
    const bool empty_rank = n_active == 0 && ctx->n_ranks > 1;
    if (!empty_rank) PHY_TRY(resident_shape(ctx, &k, &canon, &nh, -1));
    PHY_TRY(phy_sync_indexes(ctx));
    ctx->have_match = ctx->have_merged = false;
    ctx->hashes_valid = false;  // K1 is part of every match pass (never served from a cache)
    const uint64_t l0 = ctx->launches;
    PHY_CUDA(ctx, cudaEventRecord(ctx->ev_ph[0], ctx->stream));
    if (empty_rank) {
        ctx->h_nk.assign((size_t)ctx->nq + 1, 0);
        ctx->kmer_tables_valid = false;
        ctx->n_units = ctx->n_hits = 0;
        ctx->units_ordered = true;
        PHY_TRY(phy_ensure(ctx, ctx->d_qcount, ctx->nq + 1));
        PHY_CUDA(ctx, cudaMemsetAsync(ctx->d_qcount.p, 0, ((size_t)ctx->nq + 1) * sizeof(uint32_t), ctx->stream));
        PHY_CUDA(ctx, cudaEventRecord(ctx->ev_ph[1], ctx->stream));
    } else {
        PHY_TRY(prepare_hashes(ctx, k, canon, nh));
        PHY_CUDA(ctx, cudaEventRecord(ctx->ev_ph[1], ctx->stream));
        PHY_TRY(phy_launch_gather(ctx, p));
    }
    PHY_CUDA(ctx, cudaEventRecord(ctx->ev_ph[2], ctx->stream));
    PHY_TRY(phy_launch_sort_units(ctx));
    ctx->merge_top_n = merge_top_n;
    if (merge_top_n) {
        PHY_TRY(phy_launch_merge(ctx, merge_top_n));
        if (ctx->n_ranks > 1) PHY_TRY(phy_nccl_merge(ctx, merge_top_n));
        ctx->have_merged = true;
    }
    PHY_CUDA(ctx, cudaEventRecord(ctx->ev_ph[3], ctx->stream));
    PHY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 3; i++) cudaEventElapsedTime(&ctx->phase_ms[i], ctx->ev_ph[i], ctx->ev_ph[i + 1]);
    ctx->phase_ms[3] = (float)(ctx->launches - l0);
    ctx->have_match = true;
    return PHY_OK;
}

extern "C" int phy_results_fetch(phy_ctx* ctx, phy_results** out) {
    if (!ctx || !out) return PHY_ERR_ARG;
    *out = nullptr;
    if (!ctx->have_match) {
        phy_set_error(ctx, "phy_match_run must succeed before phy_results_fetch");
        return PHY_ERR_STATE;
    }
    PHY_CUDA(ctx, cudaSetDevice(ctx->device));
    phy_results* r = (phy_results*)calloc(1, sizeof(phy_results));
    if (!r) return PHY_ERR_NOMEM;
    r->n_queries = ctx->nq;
    int n_idx = 0;
    phy_index_count(ctx, &n_idx);
    r->n_indexes = (uint32_t)n_idx;
    r->n_units = ctx->n_units;
    r->n_hits = ctx->n_hits;
    r->units = (phy_unit*)result_alloc(ctx, std::max<uint64_t>(1, r->n_units) * sizeof(phy_unit));
    r->hits = (phy_hit*)result_alloc(ctx, std::max<uint64_t>(1, r->n_hits) * sizeof(phy_hit));
    uint32_t* nk = (uint32_t*)malloc((size_t)(ctx->nq + 1) * sizeof(uint32_t));
    if (!r->units || !r->hits || !nk) {
        result_free(r->units); result_free(r->hits); free(nk); free(r);
        phy_set_error(ctx, "host memory exhausted");
        return PHY_ERR_NOMEM;
    }
    memcpy(nk, ctx->h_nk.data(), (size_t)ctx->nq * sizeof(uint32_t));
    r->n_kmers = nk;
    int rc = PHY_OK;
    if (r->n_units) rc = phy_d2h(ctx, r->units, ctx->d_units.p, r->n_units * sizeof(phy_unit));
    if (rc == PHY_OK && r->n_hits) rc = phy_d2h(ctx, r->hits, ctx->d_hits.p, r->n_hits * sizeof(phy_hit));
    if (rc != PHY_OK) {
        phy_results_free(r);
        return rc;
    }
    if (!ctx->units_ordered)  // the device orders them unless the (index, query) table was too large
        std::sort(r->units, r->units + r->n_units, [](const phy_unit& a, const phy_unit& b) {
            return a.index != b.index ? a.index < b.index : a.query < b.query;
        });
    r->h2d_bytes = ctx->h2d_bytes;
    r->d2h_bytes = r->n_units * sizeof(phy_unit) + r->n_hits * sizeof(phy_hit);
    *out = r;
    return PHY_OK;
}

extern "C" void phy_results_free(phy_results* r) {
    if (!r) return;
    result_free(r->units);
    result_free(r->hits);
    free((void*)r->n_kmers);
    free(r);
}

extern "C" int phy_match(phy_ctx* ctx, const phy_match_params* p, phy_results** out) {
    PHY_TRY(phy_match_run(ctx, p, 0));
    return phy_results_fetch(ctx, out);
}

extern "C" int phy_scores(phy_ctx* ctx, int idx_id, uint32_t* host_scores) {
    HostIndex* ix = get_index(ctx, idx_id);
    if (!ix || !host_scores) return PHY_ERR_ARG;
    if (!ix->committed || !ctx->have_queries) {
        phy_set_error(ctx, "phy_scores needs a committed index and phy_queries_set");
        return PHY_ERR_STATE;
    }
    PHY_CUDA(ctx, cudaSetDevice(ctx->device));
    PHY_TRY(phy_sync_indexes(ctx));
    PHY_TRY(prepare_hashes(ctx, ix->term_size, ix->canon, ix->d.num_hashes));
    const size_t n = (size_t)ctx->nq * ix->d.n_docs;
    PHY_TRY(phy_ensure(ctx, ctx->d_scores, n + 1));
    PHY_TRY(phy_launch_scores(ctx, idx_id, ctx->d_scores.p));
    return phy_d2h(ctx, host_scores, ctx->d_scores.p, n * sizeof(uint32_t));
}

// ---- merged ---------------------------------------------------------------------------------------
extern "C" int phy_merged_fetch(phy_ctx* ctx, phy_merged** out) {
    if (!ctx || !out) return PHY_ERR_ARG;
    *out = nullptr;
    if (!ctx->have_merged) {
        phy_set_error(ctx, "phy_match_run(..., merge_top_n > 0) must succeed before phy_merged_fetch");
        return PHY_ERR_STATE;
    }
    PHY_CUDA(ctx, cudaSetDevice(ctx->device));
    phy_merged* m = (phy_merged*)calloc(1, sizeof(phy_merged));
    if (!m) return PHY_ERR_NOMEM;
    m->n_queries = ctx->nq;
    const bool holder = ctx->n_ranks == 1 || ctx->rank == 0 || ctx->merge_sharded;
    m->offs = (uint64_t*)result_alloc(ctx, ((size_t)ctx->nq + 1) * sizeof(uint64_t));
    if (!m->offs) {
        phy_merged_free(m);
        phy_set_error(ctx, "host memory exhausted");
        return PHY_ERR_NOMEM;
    }
    memset(m->offs, 0, ((size_t)ctx->nq + 1) * sizeof(uint64_t));
    int rc = PHY_OK;
    if (holder) rc = phy_d2h(ctx, m->offs, ctx->d_foffs.p, ((size_t)ctx->nq + 1) * sizeof(uint64_t));
    if (rc == PHY_OK) rc = cudaStreamSynchronize(ctx->stream) == cudaSuccess ? PHY_OK : PHY_ERR_CUDA;
    // the exact number of kept candidates is the last offset (ctx->n_final is only the bound the device
    // buffers were sized with)
    const uint64_t n = (holder && rc == PHY_OK) ? m->offs[ctx->nq] : 0;
    if (n > ctx->n_final) {
        phy_merged_free(m);
        phy_set_error(ctx, "internal: merged list longer than its bound");
        return PHY_ERR_STATE;
    }
    if (rc == PHY_OK) {
        m->cands = (phy_cand*)result_alloc(ctx, std::max<uint64_t>(1, n) * sizeof(phy_cand));
        if (!m->cands) {
            phy_merged_free(m);
            phy_set_error(ctx, "host memory exhausted");
            return PHY_ERR_NOMEM;
        }
        if (n) rc = phy_d2h(ctx, m->cands, ctx->d_final.p, n * sizeof(phy_cand));
    }
    if (rc != PHY_OK) {
        phy_merged_free(m);
        return rc;
    }
    m->d2h_bytes = holder ? ((uint64_t)ctx->nq + 1) * 8 + n * sizeof(phy_cand) : 0;
    *out = m;
    return PHY_OK;
}

int phy_merge_host_impl(phy_ctx* ctx, uint32_t nq, uint32_t top_n, const uint64_t* offs, const phy_cand* cands);

extern "C" int phy_merge_host(phy_ctx* ctx, uint32_t n_queries, uint32_t top_n, const uint64_t* offs,
                              const phy_cand* cands) {
    if (!ctx || !offs || (!cands && offs[n_queries])) return PHY_ERR_ARG;
    for (uint32_t q = 0; q < n_queries; q++)
        if (offs[q + 1] < offs[q]) {
            phy_set_error(ctx, "candidate offsets must be non-decreasing");
            return PHY_ERR_ARG;
        }
    for (uint64_t i = 0; i < offs[n_queries]; i++)
        if (cands[i].batch_rank >= PHY_MAX_BATCH_RANK || cands[i].ref_rank >= PHY_MAX_DOCS) {
            phy_set_error(ctx, "candidate %llu: rank out of range", (unsigned long long)i);
            return PHY_ERR_ARG;
        }
    PHY_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->have_queries = ctx->have_match = false;  // the query set of this ctx is replaced
    ctx->kmer_tables_valid = ctx->ioffs_valid = ctx->gather_tables_valid = false;
    ctx->have_merged = false;
    PHY_TRY(phy_merge_host_impl(ctx, n_queries, top_n, offs, cands));
    PHY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->have_merged = true;
    return PHY_OK;
}

extern "C" void phy_merged_free(phy_merged* m) {
    if (!m) return;
    result_free(m->offs);
    result_free(m->cands);
    free(m);
}

// ---- timing -----------------------------------------------------------------------------------------
extern "C" int phy_timer_start(phy_ctx* ctx) {
    if (!ctx) return PHY_ERR_ARG;
    PHY_CUDA(ctx, cudaSetDevice(ctx->device));
    PHY_CUDA(ctx, cudaEventRecord(ctx->ev_t0, ctx->stream));
    return PHY_OK;
}
extern "C" int phy_timer_stop(phy_ctx* ctx, float* ms) {
    if (!ctx || !ms) return PHY_ERR_ARG;
    PHY_CUDA(ctx, cudaEventRecord(ctx->ev_t1, ctx->stream));
    PHY_CUDA(ctx, cudaEventSynchronize(ctx->ev_t1));
    PHY_CUDA(ctx, cudaEventElapsedTime(ms, ctx->ev_t0, ctx->ev_t1));
    return PHY_OK;
}
extern "C" int phy_sync(phy_ctx* ctx) {
    if (!ctx) return PHY_ERR_ARG;
    PHY_CUDA(ctx, cudaSetDevice(ctx->device));
    PHY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PHY_OK;
}
extern "C" int phy_last_phase_ms(phy_ctx* ctx, float out[4]) {
    if (!ctx || !out) return PHY_ERR_ARG;
    for (int i = 0; i < 4; i++) out[i] = ctx->phase_ms[i];
    return PHY_OK;
}
extern "C" int phy_last_gather_bytes(phy_ctx* ctx, uint64_t* bytes) {
    if (!ctx || !bytes) return PHY_ERR_ARG;
    *bytes = ctx->gathered_bytes;
    return PHY_OK;
}
extern "C" int phy_merged_range(phy_ctx* ctx, uint32_t* q_lo, uint32_t* q_hi) {
    if (!ctx || !q_lo || !q_hi) return PHY_ERR_ARG;
    if (!ctx->have_merged) {
        phy_set_error(ctx, "no merged result");
        return PHY_ERR_STATE;
    }
    *q_lo = ctx->n_ranks > 1 ? ctx->merged_q_lo : 0;
    *q_hi = ctx->n_ranks > 1 ? ctx->merged_q_hi : ctx->nq;
    return PHY_OK;
}
extern "C" int phy_last_gather_bytes_of(phy_ctx* ctx, int idx_id, uint64_t* bytes) {
    if (!ctx || !bytes || idx_id < 0) return PHY_ERR_ARG;
    *bytes = (size_t)idx_id < ctx->h_idx_bytes.size() ? ctx->h_idx_bytes[idx_id] : 0;
    return PHY_OK;
}
extern "C" int phy_ctx_budget(phy_ctx* ctx, uint64_t* budget, uint64_t* used) {
    if (!ctx) return PHY_ERR_ARG;
    if (budget) *budget = ctx->budget;
    if (used) *used = ctx->used;
    return PHY_OK;
}
extern "C" int phy_ctx_set_option(phy_ctx* ctx, const char* name, int64_t value) {
    if (!ctx || !name) return PHY_ERR_ARG;
    if (!strcmp(name, "prune")) ctx->prune = value != 0;
    else if (!strcmp(name, "pinned_results")) ctx->pinned_results = value != 0;
    else if (!strcmp(name, "merge_mode") && (value == 0 || value == 1)) ctx->merge_sharded = value == 1;
    else if (!strcmp(name, "shard_query_upload")) ctx->shard_query_upload = value != 0;
    else if (!strcmp(name, "sanitize_queries")) ctx->sanitize_queries = value != 0;
    else {
        phy_set_error(ctx, "unknown option %s=%lld", name, (long long)value);
        return PHY_ERR_ARG;
    }
    return PHY_OK;
}
extern "C" int phy_flush_l2(phy_ctx* ctx) {
    if (!ctx) return PHY_ERR_ARG;
    const size_t n = 256u << 20;  // 2x the 126 MB L2
    PHY_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->d_flush.p) {
        void* p = nullptr;
        PHY_TRY(phy_ws_alloc(ctx, &p, n));
        ctx->d_flush.p = (uint8_t*)p;
        ctx->d_flush.cap = n;
    }
    PHY_CUDA(ctx, cudaMemsetAsync(ctx->d_flush.p, 0x5A, n, ctx->stream));
    return PHY_OK;
}

// ---- synthetic workload ---------------------------------------------------------------------------------
extern "C" int phy_index_synth(phy_ctx* ctx, int idx_id, const phy_synth_spec* spec) {
    HostIndex* ix = get_index(ctx, idx_id);
    if (!ix || !spec) return PHY_ERR_ARG;
    if (ix->committed || ix->pushed || spec->n_docs != ix->d.n_docs) {
        phy_set_error(ctx, "phy_index_synth needs a fresh index with n_docs == spec.n_docs");
        return PHY_ERR_STATE;
    }
    PHY_CUDA(ctx, cudaSetDevice(ctx->device));
    PHY_TRY(phy_synth_build(ctx, *ix, spec));
    PHY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ix->pushed = ix->body_bytes;
    ix->committed = true;
    ctx->indexes_dirty = true;
    ctx->index_version++;
    return PHY_OK;
}

int phy_insert_kmers(phy_ctx* ctx, HostIndex& ix, const uint32_t* d_doc_of_query);

extern "C" int phy_index_insert(phy_ctx* ctx, int idx_id, const uint32_t* doc_of_query) {
    HostIndex* ix = get_index(ctx, idx_id);
    if (!ix || !doc_of_query) return PHY_ERR_ARG;
    if (!ix->committed || !ctx->have_queries) {
        phy_set_error(ctx, "phy_index_insert needs a committed index and phy_queries_set");
        return PHY_ERR_STATE;
    }
    PHY_CUDA(ctx, cudaSetDevice(ctx->device));
    PHY_TRY(prepare_hashes(ctx, ix->term_size, ix->canon, ix->d.num_hashes));
    PHY_TRY(phy_check_hash_error(ctx));
    PHY_TRY(phy_ensure(ctx, ctx->d_slotq, ctx->nq + 1));
    PHY_TRY(phy_h2d(ctx, ctx->d_slotq.p, doc_of_query, (size_t)ctx->nq * sizeof(uint32_t)));
    PHY_TRY(phy_insert_kmers(ctx, *ix, ctx->d_slotq.p));
    PHY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->have_match = ctx->have_merged = false;
    return PHY_OK;
}

extern "C" int phy_synth_reads(phy_ctx* ctx, const phy_synth_spec* specs, uint32_t n_specs, uint64_t reads_seed,
                               uint64_t first_read, uint32_t n_reads, uint32_t read_len, uint32_t random_q8,
                               uint32_t err_q16, char* host_out) {
    if (!ctx || !host_out || (n_specs && !specs)) return PHY_ERR_ARG;
    PHY_CUDA(ctx, cudaSetDevice(ctx->device));
    void *d_specs = nullptr, *d_out = nullptr;
    const size_t nbytes = (size_t)n_reads * read_len;
    PHY_TRY(phy_dev_alloc(ctx, &d_specs, std::max<size_t>(1, n_specs) * sizeof(phy_synth_spec), false));
    int rc = phy_dev_alloc(ctx, &d_out, nbytes, false);
    if (rc == PHY_OK && n_specs) rc = phy_h2d(ctx, d_specs, specs, n_specs * sizeof(phy_synth_spec));
    if (rc == PHY_OK)
        rc = phy_synth_reads_dev(ctx, (const phy_synth_spec*)d_specs, n_specs, reads_seed, first_read, n_reads,
                                 read_len, random_q8, err_q16, (char*)d_out);
    if (rc == PHY_OK) rc = phy_d2h(ctx, host_out, d_out, nbytes);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_specs);
    if (d_out) cudaFree(d_out);
    return rc;
}
