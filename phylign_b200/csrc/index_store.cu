// index_store.cu -- layout of a COBS classic index in HBM, and the synthetic workload.
//
// File layout (what `cobs query --load-complete -i ...` reads,
// /root/reference/scripts/run_cobs_streaming.sh:24-29): signature_size rows of
// row_size = ceil(D/8) bytes, packed.  HBM layout: same rows at a stride that is a
// multiple of 16 B (32 B above 32 B) so every lane's 128-bit load is aligned and a
// 500-B row costs 16 sectors, not 17; padding bytes are zero.
//
// The synthetic builder restates `cobs classic-construct` (Appendix A.10) for procedural
// genomes (spec v1, mirrored by oracle/cobs_oracle.c: orc_synth_*): needed because the
// 90 GB - 1 TB indexes of BASELINE configs 3/4 cannot be built or shipped from the host.
#include "phy_internal.cuh"

namespace {

__global__ void __launch_bounds__(256) restride_kernel(const uint8_t* __restrict__ src, uint64_t body_off,
                                                       uint64_t nbytes, uint32_t row_size, uint32_t stride,
                                                       uint8_t* __restrict__ rows) {
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nbytes;
         t += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t o = body_off + t;
        uint64_t r = o / row_size;
        uint32_t c = (uint32_t)(o - r * row_size);
        rows[r * stride + c] = src[t];
    }
}

__global__ void __launch_bounds__(256) destride_kernel(const uint8_t* __restrict__ rows, uint64_t nbytes,
                                                       uint32_t row_size, uint32_t stride,
                                                       uint8_t* __restrict__ dst) {
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nbytes;
         t += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t r = t / row_size;
        uint32_t c = (uint32_t)(t - r * row_size);
        dst[t] = rows[r * stride + c];
    }
}

// ---- synthetic workload spec v1 ---------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ uint32_t sub_base(uint32_t b, uint64_t m) {
    return (b + 1u + (uint32_t)(((m >> 16) & 0xFFFFu) % 3u)) & 3u;
}
__host__ __device__ __forceinline__ uint32_t synth_base(const phy_synth_spec& s, uint32_t d, uint32_t pos) {
    uint32_t b = (uint32_t)(mix64(s.seed ^ ((uint64_t)pos * 0xD6E8FEB86659FD93ULL)) >> 62);
    uint64_t clade = d / (s.clade_size ? s.clade_size : 1u);
    uint64_t m1 = mix64((s.seed + (clade + 1) * 0x9E3779B97F4A7C15ULL) ^ ((uint64_t)pos * 0xC2B2AE3D27D4EB4FULL));
    if ((uint32_t)(m1 & 0xFFFFu) < s.clade_sub_q16) b = sub_base(b, m1);
    uint64_t m2 = mix64((s.seed + ((uint64_t)d + 0x100000001ULL) * 0xBF58476D1CE4E5B9ULL) ^
                        ((uint64_t)pos * 0x94D049BB133111EBULL));
    if ((uint32_t)(m2 & 0xFFFFu) < s.doc_sub_q16) b = sub_base(b, m2);
    return b;
}

constexpr uint64_t XP1 = 0x9E3779B185EBCA87ULL, XP2 = 0xC2B2AE3D27D4EB4FULL, XP3 = 0x165667B19E3779F9ULL,
                   XP4 = 0x85EBCA77C2B2AE63ULL, XP5 = 0x27D4EB2F165667C5ULL;
__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

// XXH64 of the 31 ASCII letters of the 2-bit packed k-mer v (MSB-first), seed j
__device__ __forceinline__ uint64_t xxh64_kmer31(uint64_t v, uint64_t seed) {
    uint64_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 31; i++) {
        uint32_t code = (uint32_t)(v >> (2 * (30 - i))) & 3u;
        uint64_t ch = (0x54474341u >> (8 * code)) & 0xFFu;
        w[i >> 3] |= ch << (8 * (i & 7));
    }
    uint64_t h = seed + XP5 + 31;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        h ^= rotl64(w[i] * XP2, 31) * XP1;
        h = rotl64(h, 27) * XP1 + XP4;
    }
    h ^= (w[3] & 0xFFFFFFFFULL) * XP1;
    h = rotl64(h, 23) * XP2 + XP3;
    uint64_t tail = w[3] >> 32;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        h ^= (tail & 0xFFULL) * XP5;
        h = rotl64(h, 11) * XP1;
        tail >>= 8;
    }
    h ^= h >> 33; h *= XP2; h ^= h >> 29; h *= XP3; h ^= h >> 32;
    return h;
}

// classic-construct for procedural genomes: thread = (document, run of SEG k-mer positions)
constexpr uint32_t SEG = 64;
__global__ void __launch_bounds__(256) synth_build_kernel(phy_synth_spec s, uint32_t canonicalize,
                                                          uint64_t sig, uint64_t magic, uint32_t num_hashes,
                                                          uint32_t stride, uint8_t* __restrict__ rows) {
    const uint32_t n_kmers = s.genome_len - 30u;
    const uint32_t segs = (n_kmers + SEG - 1) / SEG;
    const uint64_t total = (uint64_t)s.n_docs * segs;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (uint64_t)gridDim.x * blockDim.x) {
        // consecutive threads = consecutive documents of the same segment -> same index rows
        // (related genomes share most k-mers), which keeps the atomics inside a few sectors
        const uint32_t d = (uint32_t)(t % s.n_docs);
        const uint32_t p0 = (uint32_t)(t / s.n_docs) * SEG;
        const uint32_t p1 = min(p0 + SEG, n_kmers);
        uint64_t fwd = 0, rc = 0;
        const uint64_t mask = (1ULL << 62) - 1;
        for (uint32_t i = 0; i < 30; i++) {
            uint32_t b = synth_base(s, d, p0 + i);
            fwd = (fwd << 2) | b;
            rc = (rc >> 2) | ((uint64_t)(3u - b) << 60);
        }
        uint32_t* words = reinterpret_cast<uint32_t*>(rows);
        for (uint32_t p = p0; p < p1; p++) {
            uint32_t b = synth_base(s, d, p + 30);
            fwd = ((fwd << 2) | b) & mask;
            rc = (rc >> 2) | ((uint64_t)(3u - b) << 60);
            uint64_t v = (canonicalize && rc < fwd) ? rc : fwd;
            for (uint32_t j = 0; j < num_hashes; j++) {
                uint32_t row = phy_fastmod(xxh64_kmer31(v, j), sig, magic);
                // related genomes share most k-mers: lanes (= neighbouring documents) that hit
                // the same 32-bit word merge their bits and issue ONE atomic
                const uint64_t bit = (uint64_t)row * stride * 8ull + d;
                const unsigned long long word = bit >> 5;
                const unsigned peers = __match_any_sync(__activemask(), word);
                const uint32_t m = __reduce_or_sync(peers, 1u << (bit & 31));
                if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicOr(&words[word], m);
            }
        }
    }
}

// classic-construct for real sequences: OR every k-mer of query q into document doc_of_query[q]
__global__ void __launch_bounds__(256) insert_kmers_kernel(const uint64_t* __restrict__ hashes, uint64_t total_kmers,
                                                           const uint64_t* __restrict__ koffs, uint32_t nq,
                                                           const uint32_t* __restrict__ doc_of_query, uint64_t sig,
                                                           uint64_t magic, uint32_t num_hashes, uint32_t stride,
                                                           uint32_t n_docs, uint8_t* __restrict__ rows) {
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_kmers) return;
    uint32_t lo = 0, hi = nq;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (koffs[mid] <= g) lo = mid; else hi = mid;
    }
    const uint32_t d = doc_of_query[lo];
    if (d >= n_docs) return;
    uint32_t* words = reinterpret_cast<uint32_t*>(rows);
    for (uint32_t j = 0; j < num_hashes; j++) {
        const uint32_t row = phy_fastmod(hashes[(uint64_t)j * total_kmers + g], sig, magic);
        const uint64_t bit = (uint64_t)row * stride * 8ull + d;
        atomicOr(&words[bit >> 5], 1u << (bit & 31));
    }
}

__global__ void __launch_bounds__(256) synth_reads_kernel(const phy_synth_spec* __restrict__ specs,
                                                          uint32_t n_specs, uint64_t reads_seed,
                                                          uint64_t first_read, uint32_t n_reads,
                                                          uint32_t read_len, uint32_t random_q8,
                                                          uint32_t err_q16, char* __restrict__ out) {
    // one warp per read
    const uint32_t lane = threadIdx.x & 31;
    for (uint64_t ri = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); ri < n_reads;
         ri += (uint64_t)gridDim.x * (blockDim.x >> 5)) {
        const uint64_t r = first_read + ri;
        char* o = out + ri * read_len;
        const uint64_t u = mix64(reads_seed + r * 0x9E3779B97F4A7C15ULL);
        if ((uint32_t)(u & 0xFFu) < random_q8 || n_specs == 0) {
            for (uint32_t j = lane; j < read_len; j += 32)
                o[j] = "ACGT"[mix64(u + (uint64_t)j * 0xD6E8FEB86659FD93ULL) >> 62];
            continue;
        }
        const phy_synth_spec s = specs[(uint32_t)((u >> 8) & 0xFFFFFFu) % n_specs];
        const uint32_t d = (uint32_t)(u >> 32) % s.n_docs;
        const uint64_t u2 = mix64(u);
        const uint32_t span = s.genome_len >= read_len ? s.genome_len - read_len + 1 : 1;
        const uint32_t pos = (uint32_t)((u2 >> 1) % span);
        const uint32_t strand = (uint32_t)(u2 & 1u);
        const uint64_t u3 = mix64(u2);
        for (uint32_t j = lane; j < read_len; j += 32) {
            uint32_t b = synth_base(s, d, pos + j);
            uint64_t e = mix64(u3 + (uint64_t)j * 0xC2B2AE3D27D4EB4FULL);
            if ((uint32_t)(e & 0xFFFFu) < err_q16) b = sub_base(b, e);
            if (strand) o[read_len - 1 - j] = "ACGT"[3u - b];
            else o[j] = "ACGT"[b];
        }
    }
}

}  // namespace

int phy_restride_chunk_on(phy_ctx* ctx, HostIndex& ix, const uint8_t* d_src, uint64_t body_off, uint64_t nbytes,
                          cudaStream_t st) {
    if (nbytes == 0) return PHY_OK;
    if (ix.d.stride == ix.d.row_size) {
        PHY_CUDA(ctx, cudaMemcpyAsync(ix.rows_mut + body_off, d_src, nbytes, cudaMemcpyDeviceToDevice, st));
        return PHY_OK;
    }
    unsigned blocks = (unsigned)std::min<uint64_t>((nbytes + 255) / 256, 148ull * 16);
    restride_kernel<<<blocks, 256, 0, st>>>(d_src, body_off, nbytes, ix.d.row_size, ix.d.stride, ix.rows_mut);
    PHY_CUDA(ctx, cudaGetLastError());
    return PHY_OK;
}

int phy_restride_chunk(phy_ctx* ctx, HostIndex& ix, const uint8_t* d_src, uint64_t body_off, uint64_t nbytes) {
    ctx->launches++;
    return phy_restride_chunk_on(ctx, ix, d_src, body_off, nbytes, ctx->stream);
}

int phy_destride(phy_ctx* ctx, const HostIndex& ix, uint8_t* d_dst) {
    uint64_t nbytes = ix.body_bytes;
    if (nbytes == 0) return PHY_OK;
    unsigned blocks = (unsigned)std::min<uint64_t>((nbytes + 255) / 256, 148ull * 16);
    destride_kernel<<<blocks, 256, 0, ctx->stream>>>(ix.d.rows, nbytes, ix.d.row_size, ix.d.stride, d_dst);
    PHY_CUDA(ctx, cudaGetLastError());
    return PHY_OK;
}

int phy_synth_build(phy_ctx* ctx, HostIndex& ix, const phy_synth_spec* spec) {
    if (ix.term_size != 31 || spec->genome_len < 31) {
        phy_set_error(ctx, "synthetic builder needs term_size 31 and genome_len >= 31");
        return PHY_ERR_ARG;
    }
    uint64_t total = (uint64_t)spec->n_docs * ((spec->genome_len - 30u + SEG - 1) / SEG);
    unsigned blocks = (unsigned)std::min<uint64_t>((total + 255) / 256, 148ull * 64);
    synth_build_kernel<<<blocks, 256, 0, ctx->stream>>>(*spec, ix.canon, ix.d.sig, ix.d.magic,
                                                       ix.d.num_hashes, ix.d.stride, ix.rows_mut);
    PHY_CUDA(ctx, cudaGetLastError());
    return PHY_OK;
}

int phy_synth_reads_dev(phy_ctx* ctx, const phy_synth_spec* d_specs, uint32_t n_specs, uint64_t reads_seed,
                        uint64_t first_read, uint32_t n_reads, uint32_t read_len, uint32_t random_q8,
                        uint32_t err_q16, char* d_out) {
    if (n_reads == 0) return PHY_OK;
    unsigned blocks = (unsigned)std::min<uint64_t>(((uint64_t)n_reads + 7) / 8, 148ull * 32);
    synth_reads_kernel<<<blocks, 256, 0, ctx->stream>>>(d_specs, n_specs, reads_seed, first_read, n_reads,
                                                       read_len, random_q8, err_q16, d_out);
    PHY_CUDA(ctx, cudaGetLastError());
    return PHY_OK;
}

int phy_insert_kmers(phy_ctx* ctx, HostIndex& ix, const uint32_t* d_doc_of_query) {
    if (ctx->total_kmers == 0) return PHY_OK;
    const uint64_t nblk = (ctx->total_kmers + 255) / 256;
    insert_kmers_kernel<<<(unsigned)nblk, 256, 0, ctx->stream>>>(ctx->d_hashes.p, ctx->total_kmers, ctx->d_koffs.p,
                                                                ctx->nq, d_doc_of_query, ix.d.sig, ix.d.magic,
                                                                ix.d.num_hashes, ix.d.stride, ix.d.n_docs, ix.rows_mut);
    PHY_CUDA(ctx, cudaGetLastError());
    return PHY_OK;
}
