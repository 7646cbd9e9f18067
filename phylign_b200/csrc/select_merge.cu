// select_merge.cu -- K3/K4: ordering of the per-(query,index) hit lists and the
// cross-index top-N + ties merge.
//
// Replaces cobs `counts_to_result` ordering (score desc; ties by document index, SURVEY
// 8(a) tie-order note) and /root/reference/scripts/filter_queries.py:123-150
// (`SingleQuery.add_matches/_housekeeping`): per query, every candidate whose score is >=
// the N-th largest score over all batches, ordered by (-kmers, batch, ref) (:135).
// The string order of (batch, ref) is carried as host-computed integer ranks packed into
// one 64-bit key:  [63:32] ~score  [31:20] batch_rank  [19:0] ref_rank.
#include "phy_internal.cuh"

#include <algorithm>

namespace {

constexpr int SORT_SMEM = 4096;   // hits of one unit sorted in shared memory by one block
constexpr int MERGE_SMEM = 4032;  // (key,val) pairs of one query: 12 B each, < 48 KB static

// Ascending bitonic network that tolerates n not being a power of two (missing
// elements behave as +inf and are never touched).  K = key array, V = payload or null.
template <class KT, class VT>
__device__ void bitonic_block(KT* key, VT* val, uint32_t n) {
    uint32_t n2 = 1;
    while (n2 < n) n2 <<= 1;
    for (uint32_t k = 2; k <= n2; k <<= 1) {
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            uint32_t j = i ^ (k - 1);
            if (j > i && j < n && key[i] > key[j]) {
                KT t = key[i]; key[i] = key[j]; key[j] = t;
                if (val) { VT u = val[i]; val[i] = val[j]; val[j] = u; }
            }
        }
        __syncthreads();
        for (uint32_t jj = k >> 2; jj > 0; jj >>= 1) {
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
                uint32_t j = i ^ jj;
                if (j > i && j < n && key[i] > key[j]) {
                    KT t = key[i]; key[i] = key[j]; key[j] = t;
                    if (val) { VT u = val[i]; val[i] = val[j]; val[j] = u; }
                }
            }
            __syncthreads();
        }
    }
}

// hits of one unit -> (score desc, doc asc).  One block per unit.
__global__ void __launch_bounds__(256) sort_units_kernel(const phy_unit* __restrict__ units,
                                                         uint64_t n_units, phy_hit* hits) {
    __shared__ uint64_t sk[SORT_SMEM];
    for (uint64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
        const uint32_t n = units[u].n_kept;
        if (n < 2) continue;
        uint64_t* g = reinterpret_cast<uint64_t*>(hits + units[u].offset);  // {doc, score} little-endian
        if (n <= SORT_SMEM) {
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
                uint64_t x = g[i];  // low = doc, high = score
                sk[i] = ((uint64_t)(~(uint32_t)(x >> 32)) << 32) | (uint32_t)x;
            }
            __syncthreads();
            bitonic_block<uint64_t, uint32_t>(sk, nullptr, n);
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
                uint64_t x = sk[i];
                g[i] = ((uint64_t)(~(uint32_t)(x >> 32)) << 32) | (uint32_t)x;
            }
            __syncthreads();
        } else {  // rare: more than 4096 kept hits in one unit -> sort in place in HBM
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
                uint64_t x = g[i];
                g[i] = ((uint64_t)(~(uint32_t)(x >> 32)) << 32) | (uint32_t)x;
            }
            __syncthreads();
            bitonic_block<uint64_t, uint32_t>(g, nullptr, n);
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
                uint64_t x = g[i];
                g[i] = ((uint64_t)(~(uint32_t)(x >> 32)) << 32) | (uint32_t)x;
            }
            __syncthreads();
        }
    }
}

// ---- exclusive scan uint32 -> uint64 offsets (out[n] = total) ----------------------------
constexpr int SCAN_TILE = 2048;
__global__ void __launch_bounds__(256) scan_tile_sums(const uint32_t* __restrict__ in, uint64_t n,
                                                      uint64_t* __restrict__ tile_sums) {
    __shared__ uint64_t sm[8];
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE, s = 0;
    for (uint32_t i = threadIdx.x; i < SCAN_TILE; i += 256)
        if (base + i < n) s += in[base + i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t t = 0;
        for (int i = 0; i < 8; i++) t += sm[i];
        tile_sums[blockIdx.x] = t;
    }
}
__global__ void scan_tile_offsets(uint64_t* tile_sums, uint64_t n_tiles, uint64_t* total_out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        uint64_t run = 0;
        for (uint64_t i = 0; i < n_tiles; i++) { uint64_t t = tile_sums[i]; tile_sums[i] = run; run += t; }
        *total_out = run;
    }
}
__global__ void __launch_bounds__(256) scan_tile_write(const uint32_t* __restrict__ in, uint64_t n,
                                                       const uint64_t* __restrict__ tile_offs,
                                                       uint64_t* __restrict__ out) {
    __shared__ uint64_t sm[256];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
    const uint32_t per = SCAN_TILE / 256;
    uint64_t loc[SCAN_TILE / 256], s = 0;
    for (uint32_t i = 0; i < per; i++) {
        uint64_t e = base + (uint64_t)threadIdx.x * per + i;
        loc[i] = s;
        s += e < n ? in[e] : 0;
    }
    sm[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t run = tile_offs[blockIdx.x];
        for (int i = 0; i < 256; i++) { uint64_t t = sm[i]; sm[i] = run; run += t; }
    }
    __syncthreads();
    for (uint32_t i = 0; i < per; i++) {
        uint64_t e = base + (uint64_t)threadIdx.x * per + i;
        if (e < n) out[e] = sm[threadIdx.x] + loc[i];
    }
}

// ---- scatter unit hits into per-query candidate segments ---------------------------------
__global__ void __launch_bounds__(256) scatter_cands_kernel(
    const phy_unit* __restrict__ units, uint64_t n_units, const phy_hit* __restrict__ hits,
    const DevIndex* __restrict__ indexes, const uint64_t* __restrict__ qoffs_c,
    uint32_t* __restrict__ qcursor, uint64_t* __restrict__ ckey, uint32_t* __restrict__ cval) {
    __shared__ uint32_t sm_base;
    for (uint64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
        const phy_unit pu = units[u];
        if (threadIdx.x == 0) sm_base = atomicAdd(&qcursor[pu.query], pu.n_kept);
        __syncthreads();
        const DevIndex& ix = indexes[pu.index];
        uint64_t dst = qoffs_c[pu.query] + sm_base;
        for (uint32_t i = threadIdx.x; i < pu.n_kept; i += blockDim.x) {
            phy_hit h = hits[pu.offset + i];
            ckey[dst + i] = ((uint64_t)(~h.score) << 32) | ((uint64_t)ix.batch_rank << 20) |
                            (uint64_t)ix.ref_rank[h.doc];
            cval[dst + i] = h.doc;
        }
        __syncthreads();
    }
}

// ---- per query: sort by key, cut at the N-th score, keep ties -----------------------------
__global__ void __launch_bounds__(256) merge_query_kernel(const uint64_t* __restrict__ qoffs_c,
                                                          uint32_t nq, uint32_t top_n,
                                                          uint64_t* ckey, uint32_t* cval,
                                                          uint32_t* __restrict__ n_final) {
    __shared__ uint64_t sk[MERGE_SMEM];
    __shared__ uint32_t sv[MERGE_SMEM];
    __shared__ uint32_t sm_cnt;
    for (uint32_t q = blockIdx.x; q < nq; q += gridDim.x) {
        const uint64_t o = qoffs_c[q];
        const uint64_t n64 = qoffs_c[q + 1] - o;
        const uint32_t n = (uint32_t)n64;
        if (n == 0) {
            if (threadIdx.x == 0) n_final[q] = 0;
            continue;
        }
        uint64_t* k = ckey + o;
        uint32_t* v = cval + o;
        const bool in_smem = n <= MERGE_SMEM;
        if (in_smem) {
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) { sk[i] = k[i]; sv[i] = v[i]; }
            __syncthreads();
            bitonic_block<uint64_t, uint32_t>(sk, sv, n);
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) { k[i] = sk[i]; v[i] = sv[i]; }
        } else {
            __syncthreads();
            bitonic_block<uint64_t, uint32_t>(k, v, n);
        }
        __syncthreads();
        // cut: all candidates whose score >= score of the top_n-th
        if (threadIdx.x == 0) sm_cnt = 0;
        __syncthreads();
        uint32_t keep = n;
        if (top_n != 0 && n > top_n) {
            const uint32_t cut_hi = (uint32_t)(k[top_n - 1] >> 32);  // ~score of the N-th
            uint32_t c = 0;
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) c += (uint32_t)(k[i] >> 32) <= cut_hi;
            atomicAdd(&sm_cnt, c);
            __syncthreads();
            keep = sm_cnt;
        }
        if (threadIdx.x == 0) n_final[q] = keep;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) compact_final_kernel(const uint64_t* __restrict__ qoffs_c,
                                                            const uint64_t* __restrict__ foffs,
                                                            uint32_t nq, const uint64_t* __restrict__ ckey,
                                                            const uint32_t* __restrict__ cval,
                                                            phy_cand* __restrict__ out) {
    for (uint32_t q = blockIdx.x; q < nq; q += gridDim.x) {
        const uint64_t src = qoffs_c[q], dst = foffs[q];
        const uint32_t n = (uint32_t)(foffs[q + 1] - dst);
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            uint64_t k = ckey[src + i];
            phy_cand c;
            c.score = ~(uint32_t)(k >> 32);
            c.batch_rank = (uint32_t)(k >> 20) & 0xFFFu;
            c.ref_rank = (uint32_t)k & 0xFFFFFu;
            c.doc = cval[src + i];
            out[dst + i] = c;
        }
    }
}

// units[] -> (index, query) order: cell c of the dense table goes to position scan[c]
__global__ void __launch_bounds__(256) order_units_kernel(const uint32_t* __restrict__ flag,
                                                          const uint32_t* __restrict__ id,
                                                          const uint64_t* __restrict__ pos, uint64_t cells,
                                                          const phy_unit* __restrict__ in, phy_unit* __restrict__ out) {
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += (uint64_t)gridDim.x * blockDim.x)
        if (flag[c]) out[pos[c]] = in[id[c]];
}

__global__ void __launch_bounds__(256) cands_to_keys_kernel(const phy_cand* __restrict__ c, uint64_t n,
                                                            uint64_t* __restrict__ ckey, uint32_t* __restrict__ cval) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        phy_cand x = c[i];
        ckey[i] = ((uint64_t)(~x.score) << 32) | ((uint64_t)(x.batch_rank & 0xFFFu) << 20) | (x.ref_rank & 0xFFFFFu);
        cval[i] = x.doc;
    }
}

}  // namespace

int phy_order_units(phy_ctx* ctx, uint64_t cells) {
    PHY_TRY(phy_ensure(ctx, ctx->d_unit_pos, cells + 2));
    uint64_t total = 0;
    PHY_TRY(phy_exscan(ctx, ctx->d_unit_flag.p, cells, ctx->d_unit_pos.p, &total));
    if (total != ctx->n_units) {
        phy_set_error(ctx, "internal: unit table holds %llu entries, expected %llu", (unsigned long long)total,
                      (unsigned long long)ctx->n_units);
        return PHY_ERR_STATE;
    }
    if (ctx->d_units_sorted.cap != ctx->d_units.cap) {  // twin of d_units: exactly the same capacity (they swap)
        phy_release(ctx, ctx->d_units_sorted);
        void* p = nullptr;
        PHY_TRY(phy_ws_alloc(ctx, &p, ctx->d_units.cap * sizeof(phy_unit)));
        ctx->d_units_sorted.p = (phy_unit*)p;
        ctx->d_units_sorted.cap = ctx->d_units.cap;
    }
    unsigned blocks = (unsigned)std::min<uint64_t>((cells + 255) / 256, 148ull * 16);
    order_units_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_unit_flag.p, ctx->d_unit_id.p, ctx->d_unit_pos.p, cells,
                                                       ctx->d_units.p, ctx->d_units_sorted.p);
    ctx->launches++;
    PHY_CUDA(ctx, cudaGetLastError());
    std::swap(ctx->d_units, ctx->d_units_sorted);
    return PHY_OK;
}

int phy_merge_segments_bounded(phy_ctx* ctx, uint32_t top_n, uint64_t max_total);

// candidates supplied by the host, already grouped per query (offs[nq+1])
int phy_merge_host_impl(phy_ctx* ctx, uint32_t nq, uint32_t top_n, const uint64_t* offs, const phy_cand* cands) {
    const uint64_t total = offs[nq];
    ctx->nq = nq;
    PHY_TRY(phy_ensure(ctx, ctx->d_qoffs_c, nq + 2));
    PHY_TRY(phy_h2d(ctx, ctx->d_qoffs_c.p, offs, ((size_t)nq + 1) * sizeof(uint64_t)));
    PHY_TRY(phy_ensure(ctx, ctx->d_ckey, total + 1));
    PHY_TRY(phy_ensure(ctx, ctx->d_cval, total + 1));
    PHY_TRY(phy_ensure(ctx, ctx->d_recv, total + 1));
    if (total) {
        PHY_TRY(phy_h2d(ctx, ctx->d_recv.p, cands, total * sizeof(phy_cand)));
        unsigned blocks = (unsigned)std::min<uint64_t>((total + 255) / 256, 148ull * 16);
        cands_to_keys_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_recv.p, total, ctx->d_ckey.p, ctx->d_cval.p);
        ctx->launches++;
        PHY_CUDA(ctx, cudaGetLastError());
    }
    return phy_merge_segments_bounded(ctx, top_n, total);
}

// out must hold n+1 entries; *total_host receives out[n]
int phy_exscan(phy_ctx* ctx, const uint32_t* d_in, uint64_t n, uint64_t* d_out, uint64_t* total_host) {
    const uint64_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    DevBuf<uint64_t>& ts = ctx->d_scan_tmp;
    PHY_TRY(phy_ensure(ctx, ts, n_tiles + 2));
    if (n_tiles) {
        scan_tile_sums<<<(unsigned)n_tiles, 256, 0, ctx->stream>>>(d_in, n, ts.p);
        scan_tile_offsets<<<1, 32, 0, ctx->stream>>>(ts.p, n_tiles, d_out + n);
        scan_tile_write<<<(unsigned)n_tiles, 256, 0, ctx->stream>>>(d_in, n, ts.p, d_out);
        ctx->launches += 3;
    } else {
        PHY_CUDA(ctx, cudaMemsetAsync(d_out, 0, sizeof(uint64_t), ctx->stream));
    }
    PHY_CUDA(ctx, cudaGetLastError());
    if (total_host) {  // callers that can size their buffers from a host-side bound pass nullptr: no host wait
        PHY_CUDA(ctx, cudaMemcpyAsync(total_host, d_out + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
        PHY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return PHY_OK;
}

int phy_launch_sort_units(phy_ctx* ctx) {
    if (ctx->n_units == 0) return PHY_OK;
    unsigned blocks = (unsigned)std::min<uint64_t>(ctx->n_units, 148ull * 64);
    sort_units_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_units.p, ctx->n_units, ctx->d_hits.p);
    ctx->launches++;
    PHY_CUDA(ctx, cudaGetLastError());
    return PHY_OK;
}

// sort + cut the per-query candidate segments [d_qoffs_c] in (d_ckey, d_cval); compacts into d_final.
// max_total = a host-side bound on the number of kept candidates (the number of candidates that went
// in): the output is sized from it, so nothing here waits for the device -- the exact count is read
// from d_foffs[nq] by phy_merged_fetch, which has to wait for the results anyway.
int phy_merge_segments_bounded(phy_ctx* ctx, uint32_t top_n, uint64_t max_total) {
    const uint32_t nq = ctx->nq;
    PHY_TRY(phy_ensure(ctx, ctx->d_nfinal, nq + 1));
    PHY_TRY(phy_ensure(ctx, ctx->d_foffs, nq + 2));
    unsigned blocks = (unsigned)std::min<uint64_t>(std::max<uint32_t>(nq, 1), 148ull * 32);
    merge_query_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_qoffs_c.p, nq, top_n, ctx->d_ckey.p,
                                                       ctx->d_cval.p, ctx->d_nfinal.p);
    ctx->launches++;
    PHY_CUDA(ctx, cudaGetLastError());
    PHY_TRY(phy_exscan(ctx, ctx->d_nfinal.p, nq, ctx->d_foffs.p, nullptr));
    PHY_TRY(phy_ensure(ctx, ctx->d_final, max_total + 1));
    if (max_total) {
        compact_final_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_qoffs_c.p, ctx->d_foffs.p, nq,
                                                             ctx->d_ckey.p, ctx->d_cval.p, ctx->d_final.p);
        ctx->launches++;
        PHY_CUDA(ctx, cudaGetLastError());
    }
    ctx->n_final = max_total;  // upper bound; exact = d_foffs[nq]
    return PHY_OK;
}

int phy_merge_segments(phy_ctx* ctx, uint32_t top_n) {
    uint64_t total = 0;  // candidates that go in = qoffs_c[nq]
    PHY_CUDA(ctx, cudaMemcpyAsync(&total, ctx->d_qoffs_c.p + ctx->nq, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    PHY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return phy_merge_segments_bounded(ctx, top_n, total);
}

int phy_launch_merge(phy_ctx* ctx, uint32_t top_n) {
    const uint32_t nq = ctx->nq;
    PHY_TRY(phy_ensure(ctx, ctx->d_qoffs_c, nq + 2));
    const uint64_t total = ctx->n_hits;  // every kept hit is a candidate: qcount sums to n_hits (no host wait)
    PHY_TRY(phy_exscan(ctx, ctx->d_qcount.p, nq, ctx->d_qoffs_c.p, nullptr));
    PHY_TRY(phy_ensure(ctx, ctx->d_ckey, total + 1));
    PHY_TRY(phy_ensure(ctx, ctx->d_cval, total + 1));
    PHY_TRY(phy_ensure(ctx, ctx->d_qcursor, nq + 1));
    PHY_CUDA(ctx, cudaMemsetAsync(ctx->d_qcursor.p, 0, (nq + 1) * sizeof(uint32_t), ctx->stream));
    if (ctx->n_units) {
        unsigned blocks = (unsigned)std::min<uint64_t>(ctx->n_units, 148ull * 32);
        scatter_cands_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_units.p, ctx->n_units, ctx->d_hits.p,
                                                             ctx->d_indexes.p, ctx->d_qoffs_c.p,
                                                             ctx->d_qcursor.p, ctx->d_ckey.p, ctx->d_cval.p);
        ctx->launches++;
        PHY_CUDA(ctx, cudaGetLastError());
    }
    return phy_merge_segments_bounded(ctx, top_n, total);
}
